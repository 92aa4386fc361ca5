#!/bin/bash
# session V: ncu on the class-binned Poisson config (which launch is slow, and why)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/v; mkdir -p $O
timeout 900 ncu --set full --clock-control none -k regex:nm_lane_kernel -s 6 -c 3 -f -o $O/prof_cfg2p python tools/bench_configs.py cfg2p > $O/ncu.log 2>&1; echo "rc=$?"
