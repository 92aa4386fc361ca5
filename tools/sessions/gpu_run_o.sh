#!/bin/bash
# session O: four-chain KS walk vs two-chain (NM_WALK2), parity first
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/o; mkdir -p $O
echo "== pytest gpu (default)"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.log
VARIANTS="${VARIANTS:-base _w2}" bash tools/gpu_bench_variants.sh
timeout 1500 python tools/bench_configs.py cfg1 cfg2p cfg4 > $O/configs.jsonl 2> $O/configs.err; python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s step %.3f ms  kernels %s  frac %.3f  pos/s %.3g"%(d["config"], d["ms_per_step"], {k:round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"], d["positions_per_s"]))
PY
tail -3 $O/configs.err
