#!/bin/bash
# round 2, session M: the grid-key build -- all GPU tests, bench, other configs, ncu launch list + full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2m; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.log
b() { python - "$1" <<PY
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.4g ms/step %.4f lane %.4f frac %.4f grid %s/%s other %s step-kernel %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d.get("grid_tiles"),d.get("tiles"),d["roofline"]["other_kernels_ms"],d["roofline"]["step_minus_kernel_ms"]))
PY
}
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench.json 2> $O/bench.err; b $O/bench.json; tail -2 $O/bench.err
echo "== configs (grid data)"; timeout 900 python tools/bench_configs.py cfg3 cfg2p cfg5 > $O/configs.jsonl 2> $O/configs.err
echo "skip off-grid configs"
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s %.3f ms  %s  frac %.3f"%(d["config"][:60], d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"]))
PY
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-variants > $O/ncu_launch.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py launches $O/launches.csv > $O/launches.md 2>&1
echo "== ncu full (grid lane)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_grid_kernel -s 3 -c 1 -f -o /tmp/prof_grid python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-variants > $O/ncu_full.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py full /tmp/prof_grid.ncu-rep > $O/prof_grid.md 2>&1
ncu -i /tmp/prof_grid.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $O/prof_grid_source.csv.gz; ls -la $O/prof_grid_source.csv.gz
cat $O/launches.md | head -12
