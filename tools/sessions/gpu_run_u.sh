#!/bin/bash
# session U: class-binned lane tier -- parity, headline, variable-coverage configs (binned vs not)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/u; mkdir -p $O
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_gpu.log
VARIANTS="base" bash tools/gpu_bench_variants.sh
for ncs in 0 1; do
echo "NO_CLASS_SORT=$ncs"
NANOMOD_B200_NO_CLASS_SORT=$ncs timeout 900 python tools/bench_configs.py cfg2p cfg2o 2> $O/configs.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('  %-58s step %.3f ms %s'%(d['config'], d['ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items()}))"
done
tail -3 $O/configs.err
