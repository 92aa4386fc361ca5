#!/bin/bash
# round 2, session u: comparator mix of the packed 16-bit network (-DNM_CE_MIX_P16=m: one comparator in m as {min, max}
# on the ALU pipe, the others {min, 2 x IMAD}); default 3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2u; mkdir -p $O
for v in "" _mix1 _mix2 _mix4; do
  lib=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  [ -f $lib ] || continue
  NANOMOD_B200_LIB=$lib timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench$v.json 2> $O/bench$v.err
  python - <<PY
import json
try:
    l=[x for x in open("$O/bench$v.json") if x.startswith("{")][-1]; d=json.loads(l)
    print("variant '$v' ms_per_step %.4f lane %.4f frac %.4f grid_tiles %d"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["grid_tiles"]))
except Exception as e: print("variant '$v' failed", e)
PY
done
exit 0
