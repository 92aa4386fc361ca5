#!/bin/bash
# session M: key conversion / pipelined claim / IMAD walk pointers + compare-exchange flavour mix A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/m; mkdir -p $O
echo "== pytest gpu (default)"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
for v in "" _m2 _m4 _m6; do
  lib=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  [ -f $lib ] || continue
  NANOMOD_B200_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/bench$v.json 2> $O/bench$v.err
  python - <<PY
import json
l=[x for x in open("$O/bench$v.json") if x.startswith("{")][-1]; d=json.loads(l)
print("variant '$v' ms_per_step %.4f lane %.4f frac %.4f value %.4g"%(d["ms_per_step"], d["roofline"].get("kernel_ms", 0) or 0, d["roofline"]["frac"], d["value"]))
PY
done
NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200${BEST}.so timeout 1500 python tools/bench_configs.py cfg3 cfg3b cfg2p cfg4 > $O/configs.jsonl 2> $O/configs.err; python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s step %.3f ms  kernels %s  frac %.3f  pos/s %.3g"%(d["config"], d["ms_per_step"], {k:round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"], d["positions_per_s"]))
PY
tail -3 $O/configs.err
