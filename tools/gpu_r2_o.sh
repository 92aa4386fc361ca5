#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2o; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_deepw_kernel -s 2 -c 1 -f -o /tmp/prof_deepw python tools/bench_configs.py cfg5 > $O/ncu_deepw.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py full /tmp/prof_deepw.ncu-rep > $O/prof_deepw.md 2>&1
ncu -i /tmp/prof_deepw.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $O/prof_deepw_source.csv.gz
cat $O/prof_deepw.md
