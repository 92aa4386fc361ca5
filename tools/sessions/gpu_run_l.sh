#!/bin/bash
# session L: multiple-of-4 network classes (N=100 exact) and warp-pair lock-step A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/l; mkdir -p $O
echo "== pytest gpu (default)"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
echo "== pytest gpu (pair sync)"; NANOMOD_B200_PAIR_SYNC=1 timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu_ps.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu_ps.log
for ps in 0 1; do
  NANOMOD_B200_PAIR_SYNC=$ps timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/bench_ps$ps.json 2> $O/bench_ps$ps.err
  python - <<PY
import json
l=[x for x in open("$O/bench_ps$ps.json") if x.startswith("{")][-1]; d=json.loads(l)
print("pair_sync=$ps ms_per_step %.4f lane %.4f frac %.4f value %.4g"%(d["ms_per_step"], d["roofline"].get("kernel_ms", 0) or 0, d["roofline"]["frac"], d["value"]))
PY
done
NANOMOD_B200_PAIR_SYNC=${PS:-0} timeout 1500 python tools/bench_configs.py cfg1 cfg3 cfg2p cfg4 > $O/configs.jsonl 2> $O/configs.err; python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s step %.3f ms  kernels %s  frac %.3f  pos/s %.3g"%(d["config"], d["ms_per_step"], {k:round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"], d["positions_per_s"]))
PY
tail -3 $O/configs.err
