#!/bin/bash
# session T: faster Student-t tail; sanitizer on the new kernels; all-variants config
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/t; mkdir -p $O
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_gpu.log
echo "== sanitizer (downsampling, ranking, mstd)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "downsampling or ranking or mstd" -p no:cacheprovider > $O/sanitizer.log 2>&1; echo "rc=$?"; tail -4 $O/sanitizer.log
timeout 900 python tools/bench_configs.py cfg3 cfg2p 2> $O/configs.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('  %-58s step %.3f ms %s'%(d['config'], d['ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items()}))"
