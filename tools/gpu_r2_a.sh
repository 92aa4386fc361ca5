#!/bin/bash
# round 2, session A: microbench of CE building blocks, GPU tests of the default build (incl. the
# new golden / full-size tests), bench of the default build and of the float+IMAD key variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
echo "== ubench5"; timeout 300 tools/_build/ubench5 > $O/ubench5.txt 2>&1; cat $O/ubench5.txt
echo "== pytest gpu (default build)"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider --durations=12 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 $O/pytest_gpu.log
for v in "" _fi; do
  lib=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  [ -f $lib ] || continue
  for rep in 1 2; do
    NANOMOD_B200_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/bench$v.json 2> $O/bench$v.err
    python - <<PY
import json
l=[x for x in open("$O/bench$v.json") if x.startswith("{")][-1]; d=json.loads(l)
print("variant '$v' ms_per_step %.4f lane %.4f frac %.4f value %.4g"%(d["ms_per_step"], d["roofline"].get("kernel_ms", 0) or 0, d["roofline"]["frac"], d["value"]))
PY
  done
done
echo "== parity of the _fi variant"; NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200_fi.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_golden.py -m gpu -q --maxfail=5 -p no:cacheprovider > $O/pytest_fi.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_fi.log
