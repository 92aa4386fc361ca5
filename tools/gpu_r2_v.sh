#!/bin/bash
# round 2, session v: the build with NM_CE_MIX_P16 = 2 -- grid-key / golden / deep parity tests, bench with variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2v; mkdir -p $O
timeout 110 python -m pytest tests/test_gpu_grid.py tests/test_gpu_golden.py tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "grid or golden or known_answer or combination_vectors or cfg1 or deep or degenerate" > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log
timeout 80 python bench.py --no-e2e --no-cpu > $O/bench.json 2> $O/bench.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
    print("ms/step %.4f lane %.4f frac %.4f"%(d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"]))
    for k,v in d["variants"].items():
        if "kernel_ms" in v: print(k, "%.3f ms"%v["ms_per_step"], {a: round(b,3) for a,b in v["kernel_ms"].items()}, "frac %.3f"%v["tests_kernel_frac_of_hbm_peak"])
except Exception as e: print("failed", e)
PY
exit 0
