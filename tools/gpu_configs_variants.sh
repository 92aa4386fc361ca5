#!/bin/bash
# session X: bisect of the short-row configs (pre-binning library vs current)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/x; mkdir -p $O
for v in ${VARIANTS:-"" _pre}; do
  [ "$v" = "base" ] && v=""
  lib=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  [ -f $lib ] || continue
  echo "variant '$v'"
  NANOMOD_B200_LIB=$lib timeout 900 python tools/bench_configs.py ${CFGS:-cfg2h cfg4} 2> $O/configs$v.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('  %-58s step %.3f ms %s'%(d['config'], d['ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items()}))"
done
