#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/f
O=gpurun_out/f
for v in _old _flat ""; do
  export NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  echo "== variant '$v'"
  timeout 600 python bench.py --no-e2e --no-cpu > $O/bench$v.json 2> $O/bench$v.err; python -c "
import json; d=json.load(open('$O/bench$v.json')); r=d['roofline']; print('value %.4g lane %.3f ms frac %.3f step %.3f ms'%(d['value'], r['kernel_ms'], r['frac'], d['ms_per_step']))"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 3 -c 1 -f -o $O/prof$v python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu$v.log 2>&1; echo "ncu rc=$?"
done
