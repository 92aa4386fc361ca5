#!/bin/bash
# session R: issue utilisation of the lane kernel vs network size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none -k regex:nm_lane_kernel --csv --log-file $O/sweep.csv python tools/sweep_n.py 2000000 8 16 24 32 40 48 56 64 72 80 88 100 > $O/sweep.log 2>&1; echo "rc=$?"; tail -3 $O/sweep.log
python - <<PY
import csv
rows=list(csv.reader(open("$O/sweep.csv")))
st=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); h=rows[st]; ix={k:i for i,k in enumerate(h)}
import collections
d=collections.OrderedDict()
for r in rows[st+1:]:
    if len(r)<len(h): continue
    d.setdefault(r[ix["ID"]],{})[r[ix["Metric Name"]]]=r[ix["Metric Value"]]
    d[r[ix["ID"]]]["grid"]=r[ix["Grid Size"]]; d[r[ix["ID"]]]["k"]=r[ix["Kernel Name"]][:24]
ids=list(d)
for i in ids[2::3]:
    m=d[i]; print(m["k"], m["grid"], {k.split(".")[0][-24:]:v for k,v in m.items() if k not in("grid","k")})
PY
