#!/bin/bash
# round 2, session D: new tests (sharded device path, golden), full default bench (e2e, cpu, variants), reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2d; mkdir -p $O
echo "== pytest gpu (golden + parity)"; timeout 2400 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (ours, default flags)"; /usr/bin/time -v timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cut -c1-3000 $O/bench.json; grep -E "Elapsed|Maximum resident" $O/bench.err; grep -v "Elapsed\|Maximum\|^\s" $O/bench.err | tail -5
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cut -c1-600 $O/bench_ref.json; tail -3 $O/bench_ref.err
