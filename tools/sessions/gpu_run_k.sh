#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/k
O=gpurun_out/k
echo "== sanitizer (deep tier)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "deep_rows or cfg5" -p no:cacheprovider > $O/sanitizer.log 2>&1; echo "rc=$?"; tail -4 $O/sanitizer.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
timeout 1500 python tools/bench_configs.py ${CFGS:-cfg5} > $O/configs.jsonl 2> $O/configs.err; python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s step %.3f ms  kernels %s  frac %.3f  pos/s %.3g"%(d["config"], d["ms_per_step"], {k:round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"], d["positions_per_s"]))
PY
tail -3 $O/configs.err
