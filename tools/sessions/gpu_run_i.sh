#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/i
O=gpurun_out/i
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== sanitizer (pair tier)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pair_tier_variants and (poisson100 or n65)" -p no:cacheprovider > $O/sanitizer.log 2>&1; echo "rc=$?"; tail -4 $O/sanitizer.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_gpu.log
for f in 0 1; do
  echo "== bench PAIR_TIER=$f"
  NANOMOD_B200_PAIR_TIER=$f timeout 600 python bench.py --no-e2e --no-cpu > $O/bench_$f.json 2> $O/bench_$f.err; python -c "
import json; d=json.load(open('$O/bench_$f.json')); r=d['roofline']; print('value %.4g lane %.3f ms frac %.3f step %.3f ms'%(d['value'], r['kernel_ms'], r['frac'], d['ms_per_step']))"; tail -2 $O/bench_$f.err
done
for f in 0 1; do echo "configs PAIR_TIER=$f"; NANOMOD_B200_PAIR_TIER=$f timeout 1500 python tools/bench_configs.py cfg2p cfg4 > $O/configs_$f.jsonl 2> $O/configs_$f.err; python - <<PY
import json
for l in open("$O/configs_$f.jsonl"):
    d=json.loads(l); print("%-60s step %.3f ms  kernels %s  frac %.3f  pos/s %.3g"%(d["config"], d["ms_per_step"], {k:round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"], d["positions_per_s"]))
PY
done
