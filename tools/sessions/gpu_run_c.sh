#!/bin/bash
# GPU session C: lockstep variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in _pp _int_pp; do
  export NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  echo "== variant '$v'"
  timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_gpu$v.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu$v.log
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench$v.json 2> gpurun_out/bench$v.err; echo "bench rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench$v.json")); r=d["roofline"]
    print("value %.4g pos/s  lane %.3f ms  frac %.3f  other %s" % (d["value"], r["kernel_ms"], r["frac"], r["other_kernels_ms"]))
except Exception as e: print("bench parse failed", e)
PY
done
for v in _pp; do
  export NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 3 -c 1 -f -o gpurun_out/prof_lane_c$v python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_c$v.log 2>&1; echo "ncu rc=$?"
done
