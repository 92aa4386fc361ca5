#!/bin/bash
# round 2, session r: flat (straight-line) networks for the 112..128 classes of the grid-key kernel vs the looped ones
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2r; mkdir -p $O
echo "== default build"; timeout 300 python tools/bench_configs.py cfg2p cfg2x > $O/default.jsonl 2> $O/default.err; echo rc=$?
echo "== flat 112..128"; NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200_flat.so timeout 300 python tools/bench_configs.py cfg2p cfg2x > $O/flat.jsonl 2> $O/flat.err; echo rc=$?
python - <<PY
import json
for f in ("default","flat"):
    for l in open("$O/%s.jsonl"%f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["config"][:40], "ms %.3f"%d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, "frac %.3f"%d["tests_kernel_frac_of_measured_peak"])
PY
echo "== head select"; timeout 200 python tools/time_head_select.py 2>&1 | tail -1
exit 0
