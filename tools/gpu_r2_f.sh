#!/bin/bash
# round 2, session F: pipelined host entry, unrolled deep sort, dense path for mixed coverage
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2f; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
echo "== bench (ours, default flags)"; S=$(date +%s); timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$? wall=$(( $(date +%s) - S ))s"; tail -5 $O/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
print("value %.4g ms/step %.4f (median %.4f best %.4f) lane %.4f frac %.4f path %s"%(d["value"],d["ms_per_step"],d["ms_per_step_median"],d["ms_per_step_best"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d["code_path"]))
print("e2e", d["e2e"])
print("cpu", d.get("cpu_baseline"))
for k,v in d["variants"].items(): print(k, "%.3f ms"%v["ms_per_step"], {a: round(b,3) for a,b in v["kernel_ms"].items()}, "frac %.3f path %d"%(v["tests_kernel_frac_of_hbm_peak"], v["path"]))
PY
echo "== e2e without pipelining"; NANOMOD_B200_SLAB=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-variants > $O/bench_noslab.json 2> $O/bench_noslab.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_noslab.json") if l.startswith("{")][-1]); print("no-slab e2e", d["e2e"])
PY
for s in 65536 1048576; do NANOMOD_B200_SLAB=$s timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-variants > $O/bench_slab$s.json 2> $O/bench_slab$s.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_slab$s.json") if l.startswith("{")][-1]); print("slab $s e2e %.4g pos/s %.2f ms %.1f GB/s"%(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pcie_GBps"]))
PY
done
echo "== memcheck pipelined"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "pipelined or deep" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -4 $O/memcheck.log
