#!/bin/bash
# session Y: compute-sanitizer over the GPU suite (memcheck) and racecheck on the staging paths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/y; mkdir -p $O
echo "== memcheck (whole GPU suite except the full-size test)"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not cfg2_full_size" -p no:cacheprovider > $O/memcheck.log 2>&1; echo "rc=$?"; tail -4 $O/memcheck.log
echo "== racecheck (binned launches, gaps, every coverage)"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "group_binned or gaps or every_coverage_1_to_140 or downsampling_branch" -p no:cacheprovider > $O/racecheck.log 2>&1; echo "rc=$?"; tail -6 $O/racecheck.log
