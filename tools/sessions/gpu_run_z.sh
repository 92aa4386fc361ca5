#!/bin/bash
# session Z: ncu --set full of the all-variants path (lane kernel with U + t, tails kernel) and of the deep tier
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/z; mkdir -p $O
timeout 900 ncu --set full --clock-control none -k regex:nm_lane_kernel\|nm_tails_kernel -s 8 -c 2 -f -o $O/prof_cfg3 python tools/bench_configs.py cfg3 > $O/ncu_cfg3.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:nm_deep_kernel -s 3 -c 1 -f -o $O/prof_cfg5 python tools/bench_configs.py cfg5 > $O/ncu_cfg5.log 2>&1; echo "rc=$?"
