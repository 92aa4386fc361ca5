#!/bin/bash
# The single-GPU check run under gpurun: GPU tests, smoke, bench (both arms), memcheck over the newest paths,
# ncu launch list and a --set full capture of the dominant kernel -> gpurun_out/final3/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/final3; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -14 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cut -c1-400 $O/bench_ref.json
echo "== bench (ours, default flags)"; S=$(date +%s); timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$? wall=$(( $(date +%s) - S ))s"; tail -3 $O/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
print("value %.4g ms/step %.4f (median %.4f best %.4f) lane %.4f frac %.4f path %s step-kernel %.3f launches %d"%(d["value"],d["ms_per_step"],d["ms_per_step_median"],d["ms_per_step_best"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d["code_path"],d["roofline"]["step_minus_kernel_ms"],d["gpu_launches"]))
print("e2e", {k: d["e2e"][k] for k in ("value","ms_per_step","pcie_GBps")}, "clocks", d["clocks"])
print("cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("all_tests_value"), d.get("cpu_baseline",{}).get("kind"))
for k,v in d["variants"].items():
    if "kernel_ms" in v: print(k, "%.3f ms"%v["ms_per_step"], {a: round(b,3) for a,b in v["kernel_ms"].items()}, "frac %.3f path %d"%(v["tests_kernel_frac_of_hbm_peak"], v["path"]))
    else: print(k, {a: v[a] for a in ("value","ms_per_step","pcie_GBps")})
PY
echo "== memcheck (queued calls, peer stores, armed selection)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "queued or peer or armed or sharded_device" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -4 $O/memcheck.log
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-variants > $O/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full (lane)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_grid_kernel -s 3 -c 1 -f -o $O/prof_lane python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-variants > $O/ncu_full.log 2>&1; echo "rc=$?"
python tools/summarize_profile.py launches $O/launches.csv > $O/launches.md 2>&1
python tools/summarize_profile.py full $O/prof_lane.ncu-rep > $O/prof_lane.md 2>&1
ls -la $O
exit 0
