#!/bin/bash
# weak-scaling bench at N GPUs (run with gpurun --gpus N)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${NGPU:-8}
mkdir -p gpurun_out/scale
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
echo "== sharded check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/scale/check_$N.log 2>&1; grep "multi_gpu_check\|MISMATCH\|Error" gpurun_out/scale/check_$N.log | head -5
for n in ${NLIST:-$N}; do
echo "== bench --gpus $n"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > gpurun_out/scale/bench_$n.json 2> gpurun_out/scale/bench_$n.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale/bench_$n.json") if l.startswith("{")][-1]); print("n_gpus",d["n_gpus"],"value %.4g"%d["value"],"ms/step %.3f"%d["ms_per_step"],"lane ms %.3f"%d["roofline"]["kernel_ms"])
except Exception as e: print("failed", e)
PY
tail -2 gpurun_out/scale/bench_$n.err | grep -v "^W\|OMP\|\*\*\*"
done
