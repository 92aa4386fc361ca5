#!/bin/bash
# GPU session D: full check of the default build -- tests, bench line, ncu launch list + full
# capture of the lane kernel, and the other BASELINE configs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/final
O=gpurun_out/final
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (ours)"; timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cat $O/bench_ref.json
echo "== configs"; timeout 1500 python tools/bench_configs.py cfg1 cfg3 cfg2p cfg2o cfg2h cfg5 cfg4 > $O/configs.jsonl 2> $O/configs.err; echo "rc=$?"; cat $O/configs.jsonl; tail -3 $O/configs.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_kernel -s 3 -c 1 -f -o $O/prof_lane python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_full.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:nm_combine_kernel\|nm_plan -s 12 -c 4 -f -o $O/prof_small python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_small.log 2>&1; echo "rc=$?"
ls -la $O
