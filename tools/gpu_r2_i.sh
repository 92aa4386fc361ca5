#!/bin/bash
# round 2, session I: int16 device path fix, templated combine kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2i; mkdir -p $O
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.log
echo "== memcheck int16"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "int16 or combination" > $O/memcheck.log 2>&1; echo "rc=$?"; tail -3 $O/memcheck.log
echo "== configs"; timeout 900 python tools/bench_configs.py cfg3 cfg4 > $O/configs.jsonl 2> $O/configs.err
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print("%-60s %.3f ms  %s  frac %.3f"%(d["config"][:60], d["ms_per_step"], {k: round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"]))
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
print("value %.4g ms/step %.4f lane %.4f other %s step-kernel %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["other_kernels_ms"],d["roofline"]["step_minus_kernel_ms"]))
PY
