#!/bin/bash
# round 2, session J: 16-bit grid-key path of the lane tier -- tests, A/B against the float32 sort
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2j; mkdir -p $O
echo "== pytest grid"; timeout 1200 python -m pytest tests/test_gpu_grid.py -m gpu -q -p no:cacheprovider > $O/pytest_grid.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_grid.log
b() { python - "$1" <<PY
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.4g ms/step %.4f lane %.4f frac %.4f grid %s/%s other %s step-kernel %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d.get("grid_tiles"),d.get("tiles"),d["roofline"]["other_kernels_ms"],d["roofline"]["step_minus_kernel_ms"]))
PY
}
echo "== bench grid"; timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench_grid.json 2> $O/bench_grid.err; b $O/bench_grid.json; tail -2 $O/bench_grid.err
echo "== bench grid data, float path"; NANOMOD_B200_NO_GRID=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench_nogrid.json 2> $O/bench_nogrid.err; b $O/bench_nogrid.json
echo "== bench off-grid data"; timeout 600 python bench.py --off-grid --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench_offgrid.json 2> $O/bench_offgrid.err; b $O/bench_offgrid.json
echo "== configs"; timeout 900 python tools/bench_configs.py cfg3 cfg2p cfg4 > $O/configs.jsonl 2> $O/configs.err
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s %.3f ms  %s  frac %.3f"%(d["config"][:60], d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"]))
PY
echo "== ncu grid"; timeout 900 ncu --set full --clock-control none -k regex:nm_lane_dense_kernel -s 6 -c 1 -f -o /tmp/prof_grid python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-variants > $O/ncu_grid.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py full /tmp/prof_grid.ncu-rep > $O/prof_grid.md 2>&1
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.log
