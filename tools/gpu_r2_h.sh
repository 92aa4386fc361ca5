#!/bin/bash
# round 2, session H: int16 transport format, full test suite, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2h; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
echo "== memcheck int16"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "int16 or pipelined" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -3 $O/memcheck.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (ours, default flags)"; S=$(date +%s); timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$? wall=$(( $(date +%s) - S ))s"; tail -5 $O/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
print("value %.4g ms/step %.4f (median %.4f best %.4f) lane %.4f frac %.4f path %s step-kernel %.3f"%(d["value"],d["ms_per_step"],d["ms_per_step_median"],d["ms_per_step_best"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d["code_path"],d["roofline"]["step_minus_kernel_ms"]))
print("e2e", {k: d["e2e"][k] for k in ("value","ms_per_step","pcie_GBps")})
print("cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("all_tests_value"))
for k,v in d["variants"].items():
    if "kernel_ms" in v: print(k, "%.3f ms"%v["ms_per_step"], {a: round(b,3) for a,b in v["kernel_ms"].items()}, "frac %.3f path %d"%(v["tests_kernel_frac_of_hbm_peak"], v["path"]))
    else: print(k, {a: v[a] for a in ("value","ms_per_step","pcie_GBps")})
PY
