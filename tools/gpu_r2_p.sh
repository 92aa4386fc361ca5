#!/bin/bash
# round 2, session p: queued calls + peer-memory head exchange -- new tests, 2-GPU check, bench at N = 1 and 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2p; mkdir -p $O
echo "== pytest (new tests)"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "queued or peer or armed or sharded or reference_seam or candidate" > $O/pytest_new.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_new.log
NGPU=2 CHECKS="2" NLIST="1 2" bash tools/gpu_scale.sh
mkdir -p $O/scale; cp gpurun_out/scale2/*_1.* gpurun_out/scale2/*_2.* $O/scale/ 2>/dev/null
if [ "$NCCL_TOO" = "1" ]; then
echo "== bench --gpus 2 --nccl-heads"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --nccl-heads > $O/bench_2_nccl.json 2> $O/bench_2_nccl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_2_nccl.json") if l.startswith("{")][-1]); print("nccl heads: n_gpus",d["n_gpus"],"ms/step %.3f"%d["ms_per_step"],"lane ms %.3f"%d["roofline"]["kernel_ms"], d.get("head_exchange"))
except Exception as e: print("failed", e)
PY
tail -3 $O/bench_2_nccl.err | grep -v "^W\|OMP\|\*\*\*"
fi
exit 0
