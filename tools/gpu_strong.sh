#!/bin/bash
# BASELINE configs[3] (chr20, 64 444 167 positions, 2x30x) split over N GPUs through the bench's queued steps:
# NGPU=N bash tools/gpu_strong.sh   (run with gpurun --gpus N) -> gpurun_out/strong/bench_N.json
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${NGPU:-1}
O=gpurun_out/strong; mkdir -p $O
P=$(( (64444167 + N - 1) / N ))
if [ "$N" = "1" ]; then
timeout 600 python bench.py --gpus 1 --positions $P --coverage 30 --steps 20 --warmup 3 --no-e2e --no-cpu --no-variants > $O/bench_$N.json 2> $O/bench_$N.err
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --positions $P --coverage 30 --steps 20 --warmup 3 --no-e2e > $O/bench_$N.json 2> $O/bench_$N.err
fi
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_$N.json") if l.startswith("{")][-1]); print("n_gpus",d["n_gpus"],"positions/GPU",d["config"]["positions_per_gpu"],"value %.4g"%d["value"],"ms/step %.3f (median %.3f best %.3f)"%(d["ms_per_step"],d["ms_per_step_median"],d["ms_per_step_best"]),"lane ms %.3f"%d["roofline"]["kernel_ms"], d["roofline"]["other_kernels_ms"], d.get("head_exchange"), d["code_path"])
except Exception as e: print("failed", e)
PY
tail -3 $O/bench_$N.err | grep -v "^W\|OMP\|\*\*\*"
exit 0
