#!/bin/bash
# session W: straight-line networks for N > 104 (no looped sort) on the binned Poisson config
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/w; mkdir -p $O
for v in "" _nl; do
  lib=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  [ -f $lib ] || continue
  echo "variant '$v'"
  NANOMOD_B200_LIB=$lib timeout 900 python tools/bench_configs.py cfg2p cfg2x 2> $O/configs$v.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('  %-58s step %.3f ms %s'%(d['config'], d['ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items()}))"
done
