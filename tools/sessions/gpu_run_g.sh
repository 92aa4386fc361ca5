#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/g
O=gpurun_out/g
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
for v in ${VARIANTS:-""}; do
  vv=$v; [ "$v" = "default" ] && vv=""
  export NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200$vv.so
  echo "== variant '$vv'"
  timeout 600 python bench.py --no-e2e --no-cpu > $O/bench$vv.json 2> $O/bench$vv.err; python -c "
import json; d=json.load(open('$O/bench$vv.json')); r=d['roofline']; print('value %.4g lane %.3f ms frac %.3f step %.3f ms'%(d['value'], r['kernel_ms'], r['frac'], d['ms_per_step']))"
  timeout 1500 python tools/bench_configs.py ${CFGS:-cfg3 cfg2p cfg4} > $O/configs$vv.jsonl 2> $O/configs$vv.err; python -c "
import json
for l in open('$O/configs$vv.jsonl'):
    d=json.loads(l); print('%-60s step %.3f ms  kernels %s  frac %.3f  pos/s %.3g'%(d['config'], d['ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['tests_kernel_frac_of_measured_peak'], d['positions_per_s']))"; tail -2 $O/configs$vv.err
done
