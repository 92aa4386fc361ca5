#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/e
O=gpurun_out/e
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --no-e2e --no-cpu > $O/bench.json 2> $O/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench.json')); r=d['roofline']; print('value %.4g lane %.3f ms frac %.3f step %.3f ms'%(d['value'], r['kernel_ms'], r['frac'], d['ms_per_step']))"; tail -3 $O/bench.err
echo "== configs"; timeout 1500 python tools/bench_configs.py ${CFGS:-cfg1 cfg3 cfg2p cfg4} > $O/configs.jsonl 2> $O/configs.err; echo "rc=$?"; python -c "
import json
for l in open('$O/configs.jsonl'):
    d=json.loads(l); print('%-60s step %.3f ms  kernels %s  tests-kernel frac %.3f  pos/s %.3g'%(d['config'], d['ms_per_step'], {k:round(v,3) for k,v in d['kernel_ms'].items()}, d['tests_kernel_frac_of_measured_peak'], d['positions_per_s']))"; tail -3 $O/configs.err
echo "== ncu cfg3 (U+t)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 1 -c 1 -f -o $O/prof_lane_cfg3 python tools/bench_configs.py cfg3 > $O/ncu_cfg3.log 2>&1; echo "rc=$?"
echo "== ncu cfg2p"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 1 -c 1 -f -o $O/prof_lane_cfg2p python tools/bench_configs.py cfg2p > $O/ncu_cfg2p.log 2>&1; echo "rc=$?"
