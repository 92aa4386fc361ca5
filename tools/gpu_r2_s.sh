#!/bin/bash
# round 2, session s: what the armed head selection costs behind a detect call -- candidate list vs three passes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2s; mkdir -p $O
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "queued or peer or armed or sharded or candidate" > $O/pytest.log 2>&1; echo "rc=$?"; tail -4 $O/pytest.log
for v in 0 1; do
echo "== NO_HEAD_CANDS=$v"
NANOMOD_B200_NO_HEAD_CANDS=$v timeout 300 python tools/time_armed_head.py 2>&1 | tail -1
NANOMOD_B200_NO_HEAD_CANDS=$v timeout 300 python tools/time_armed_head.py 8055521 2>&1 | tail -1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nm_head -c 12 --csv --log-file $O/launches_0.csv python tools/time_armed_head.py > $O/ncu_0.log 2>&1
grep "nm_head" $O/launches_0.csv | tail -3 | cut -c1-400
exit 0
