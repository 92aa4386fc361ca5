#!/bin/bash
# session P: walk variants on the short-row config (cfg4, NMAX=64 kernel, 12 warps/SM)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/p; mkdir -p $O
for v in "" _w2 _w2d _d; do
  lib=$PWD/nanomod_b200/_C/libnanomod_b200$v.so
  [ -f $lib ] || continue
  echo "variant '$v'"
  NANOMOD_B200_LIB=$lib timeout 900 python tools/bench_configs.py cfg4 cfg2h > $O/configs$v.jsonl 2> $O/configs$v.err; python - <<PY
import json
for l in open("$O/configs$v.jsonl"):
    d=json.loads(l); print("  %-58s step %.3f ms  kernels %s  frac %.3f  pos/s %.3g"%(d["config"], d["ms_per_step"], {k:round(v,3) for k,v in d["kernel_ms"].items()}, d["tests_kernel_frac_of_measured_peak"], d["positions_per_s"]))
PY
done
