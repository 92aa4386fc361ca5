#!/bin/bash
# round 2, session t: ncu --set full of the small kernels of a step (plan, combine with the candidate list, head selection)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2t; mkdir -p $O
timeout 200 ncu --set full --clock-control none -k regex:nm_combine_kernel\|nm_plan_count\|nm_head_from_cands -s 18 -c 3 -f -o $O/prof_small python tools/time_armed_head.py > $O/ncu_small.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py full $O/prof_small.ncu-rep > $O/prof_small.md 2>&1
head -c 300 $O/prof_small.md
exit 0
