#!/bin/bash
# round 2, session N: deep tier with 16-bit key pairs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2n; mkdir -p $O
echo "== pytest deep / grid"; timeout 1200 python -m pytest tests/test_gpu_grid.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -k "deep or grid or cfg5 or beyond or moments or downsampl" > $O/pytest_deep.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_deep.log
echo "== cfg5"; timeout 900 python tools/bench_configs.py cfg5 > $O/configs.jsonl 2> $O/configs.err
NANOMOD_B200_NO_GRID=1 timeout 900 python tools/bench_configs.py cfg5 >> $O/configs.jsonl 2>> $O/configs.err
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s %.3f ms  %s  frac %.3f"%(d["config"][:60], d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"]))
PY
echo "== ncu deep"; timeout 900 ncu --set full --clock-control none -k regex:nm_deep_kernel -s 2 -c 1 -f -o /tmp/prof_deep python tools/bench_configs.py cfg5 > $O/ncu_deep.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py full /tmp/prof_deep.ncu-rep > $O/prof_deep.md 2>&1
