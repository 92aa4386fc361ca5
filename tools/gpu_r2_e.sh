#!/bin/bash
# round 2, session E: binned deep kernel -- tests, full default bench, deep configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2e; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
echo "== bench (ours, default flags)"; S=$(date +%s); timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$? wall=$(( $(date +%s) - S ))s"; cut -c1-6000 $O/bench.json; tail -5 $O/bench.err
echo "== deep configs"; timeout 900 python tools/bench_configs.py cfg5 > $O/configs.jsonl 2> $O/configs.err; cat $O/configs.jsonl | cut -c1-600
NANOMOD_B200_NO_DEEP2=1 timeout 900 python tools/bench_configs.py cfg5 > $O/configs_nodeep2.jsonl 2>> $O/configs.err; cat $O/configs_nodeep2.jsonl | cut -c1-600
echo "== memcheck deep"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "deep or cfg5 or coverage_1_to_140 or degenerate" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -4 $O/memcheck.log
