#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/h
O=gpurun_out/h
for c in cfg3 cfg2p; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 1 -c 1 -f -o $O/prof_$c python tools/bench_configs.py $c > $O/ncu_$c.log 2>&1; echo "rc=$?"
done
