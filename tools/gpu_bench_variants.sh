#!/bin/bash
# session N: ablation of the lane-kernel micro-optimisations (one library variant each)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/n; mkdir -p $O
for v in ${VARIANTS:-"" _a _b _c _d _ab}; do
  [ "$v" = "base" ] && v=""
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
