#!/bin/bash
# round 2, session K: ncu of the grid-key kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2k; mkdir -p $O
timeout 900 ncu --set full --clock-control none -k regex:nm_lane_grid_kernel -s 3 -c 1 -f -o /tmp/prof_grid python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-variants > $O/ncu_grid.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py full /tmp/prof_grid.ncu-rep > $O/prof_grid.md 2>&1
cat $O/prof_grid.md
