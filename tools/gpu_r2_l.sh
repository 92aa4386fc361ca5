#!/bin/bash
# round 2, session L: grid-key kernel variants (CTA lock-step, comparator mix), A/B on the bench workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2l; mkdir -p $O
b() { python - "$1" <<PY
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.4g ms/step %.4f lane %.4f frac %.4f grid %s/%s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d.get("grid_tiles"),d.get("tiles")))
PY
}
for v in _w _w0 _wf _w0f; do
  echo "== variant '$v'"
  NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench$v.json 2> $O/bench$v.err; b $O/bench$v.json; tail -2 $O/bench$v.err
  NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200$v.so timeout 600 python -m pytest tests/test_gpu_grid.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
done
for v in _w0f; do
  NANOMOD_B200_LIB=$PWD/nanomod_b200/_C/libnanomod_b200$v.so timeout 900 ncu --set full --clock-control none -k regex:nm_lane_grid_kernel -s 3 -c 1 -f -o /tmp/prof_grid$v python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-variants > $O/ncu_grid$v.log 2>&1; echo "rc=$?"
  python tools/summarize_profile.py full /tmp/prof_grid$v.ncu-rep > $O/prof_grid$v.md 2>&1
done
