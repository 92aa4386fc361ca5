#!/bin/bash
# sharded checks + scaling bench at N GPUs (run with gpurun --gpus N): NGPU=N [NLIST="1 2 .."] [CHR20=1]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${NGPU:-8}
O=gpurun_out/scale2; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
for n in ${CHECKS:-$N}; do
echo "== sharded check, $n GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py ${CHR20:+--chr20} > $O/check_$n.log 2>&1; echo "rc=$?"; grep "multi_gpu_check\|MISMATCH\|Error\|chr20_strong" $O/check_$n.log | head -8
done
for n in ${NLIST:-$N}; do
echo "== bench --gpus $n"
if [ "$n" = "1" ]; then
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --no-variants > $O/bench_$n.json 2> $O/bench_$n.err
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 > $O/bench_$n.json 2> $O/bench_$n.err
fi
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_$n.json") if l.startswith("{")][-1]); print("n_gpus",d["n_gpus"],"value %.4g"%d["value"],"ms/step %.3f (median %.3f best %.3f)"%(d["ms_per_step"],d["ms_per_step_median"],d["ms_per_step_best"]),"lane ms %.3f"%d["roofline"]["kernel_ms"], "e2e %.4g"%(d["e2e"]["value"] if d.get("e2e") else 0), "head rows", d.get("head_rows_exchanged"))
except Exception as e: print("failed", e)
PY
tail -3 $O/bench_$n.err | grep -v "^W\|OMP\|\*\*\*"
done
