#!/bin/bash
# round 2, session C: GPU tests after the test fixes + head selection
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu.log
echo "== memcheck (dense path tests)"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "dense or speculative or cfg1_variants or bad_offsets or ranking_head" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -4 $O/memcheck.log
