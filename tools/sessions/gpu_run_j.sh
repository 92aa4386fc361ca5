#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/j
O=gpurun_out/j
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_pair_kernel -s 3 -c 1 -f -o $O/prof_pair python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_pair.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_pair_kernel -s 1 -c 1 -f -o $O/prof_pair_cfg2p python tools/bench_configs.py cfg2p > $O/ncu_cfg2p.log 2>&1; echo "rc=$?"
