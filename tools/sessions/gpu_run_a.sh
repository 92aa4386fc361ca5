#!/bin/bash
# First GPU session: smoke under compute-sanitizer, parity tests, microbench, bench, ncu.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== sanitizer (smoke)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== ubench"; timeout 120 tools/_build/ubench > gpurun_out/ubench.log 2>&1; cat gpurun_out/ubench.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full (lane kernel)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 3 -c 1 -f -o gpurun_out/prof_lane python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
