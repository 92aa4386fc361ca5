#!/bin/bash
# round 2, session B: dense path -- GPU tests, golden diagnostics, bench, other configs, memcheck, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2b; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu.log
echo "== golden diagnostics"; timeout 300 python tools/gpu_diag_golden.py > $O/diag_golden.txt 2>&1; cat $O/diag_golden.txt | cut -c1-400
for rep in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/bench.json 2> $O/bench.err
  python - <<PY
import json
l=[x for x in open("$O/bench.json") if x.startswith("{")][-1]; d=json.loads(l)
print("default: ms_per_step %.4f lane %.4f frac %.4f value %.4g other %s"%(d["ms_per_step"], d["roofline"].get("kernel_ms", 0) or 0, d["roofline"]["frac"], d["value"], d["roofline"]["other_kernels_ms"]))
PY
done
NANOMOD_B200_NO_DENSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/bench_nodense.json 2> $O/bench_nodense.err
python - <<PY
import json
l=[x for x in open("$O/bench_nodense.json") if x.startswith("{")][-1]; d=json.loads(l)
print("no-dense: ms_per_step %.4f lane %.4f frac %.4f value %.4g other %s"%(d["ms_per_step"], d["roofline"].get("kernel_ms", 0) or 0, d["roofline"]["frac"], d["value"], d["roofline"]["other_kernels_ms"]))
PY
echo "== configs"; timeout 1500 python tools/bench_configs.py cfg1 cfg3 cfg2p cfg2o cfg2h cfg5 cfg4 > $O/configs.jsonl 2> $O/configs.err; echo "rc=$?"
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s %.3f ms  %s  frac %.3f"%(d["config"][:60], d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"]))
PY
echo "== memcheck (dense path tests)"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "dense or speculative or cfg1_variants or bad_offsets" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -4 $O/memcheck.log
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_lane_dense_kernel -s 3 -c 1 -f -o $O/prof_lane python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_full.log 2>&1; echo "rc=$?"
ls -la $O
