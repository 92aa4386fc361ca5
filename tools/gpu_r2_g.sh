#!/bin/bash
# round 2, session G: huge rows, device-resident head gather, plan pass; ncu of the unrolled deep kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2g; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
echo "== memcheck (huge rows, heads)"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "beyond_the_shared or sharded_device or ranking_head" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -4 $O/memcheck.log
echo "== configs"; timeout 900 python tools/bench_configs.py cfg5 cfg4 > $O/configs.jsonl 2> $O/configs.err
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s %.3f ms  %s  frac %.3f"%(d["config"][:60], d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"]))
PY
echo "== ncu deep"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_deep_kernel -s 2 -c 1 -f -o $O/prof_deep python tools/bench_configs.py cfg5 > $O/ncu_deep.log 2>&1; echo "rc=$?"
