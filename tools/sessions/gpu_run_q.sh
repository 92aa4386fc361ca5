#!/bin/bash
# session Q: parity + headline bench + ncu full captures of the lane kernel on the headline
# (N=100 class, 8 warps/SM) and on the 2x50x config (N=52 class, 12 warps/SM) to compare stalls
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/q; mkdir -p $O
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
VARIANTS="base" bash tools/gpu_bench_variants.sh
timeout 900 python tools/bench_configs.py cfg2h cfg4 2> $O/configs.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('  %-58s step %.3f ms lane %.3f frac %.3f'%(d['config'], d['ms_per_step'], d['kernel_ms']['lane'], d['tests_kernel_frac_of_measured_peak']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 3 -c 1 -f -o $O/prof_lane100 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_100.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 3 -c 1 -f -o $O/prof_lane50 python tools/bench_configs.py cfg2h > $O/ncu_50.log 2>&1; echo "rc=$?"
