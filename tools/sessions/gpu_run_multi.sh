#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${NGPU:-2}
mkdir -p gpurun_out/multi
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== sharded check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi/check.log 2>&1; grep "multi_gpu_check\|MISMATCH\|Error" gpurun_out/multi/check.log | head -5
for n in 1 $N; do
echo "== bench --gpus $n"
if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/multi/bench_1.json 2> gpurun_out/multi/bench_1.err
else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/multi/bench_$n.json 2> gpurun_out/multi/bench_$n.err; fi
echo "rc=$?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/multi/bench_$n.json') if l.startswith('{')][-1]); print('n_gpus',d['n_gpus'],'value %.4g'%d['value'],'ms/step %.3f'%d['ms_per_step'],'lane ms %.3f'%d['roofline']['kernel_ms'],'e2e',d['e2e'] and '%.4g'%d['e2e']['value'])"; tail -2 gpurun_out/multi/bench_$n.err
done
echo "== reference arm under torchrun"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
