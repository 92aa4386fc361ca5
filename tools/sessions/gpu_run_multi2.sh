#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${NGPU:-2}
mkdir -p gpurun_out/multi
run() { # name, env, extra args
  echo "== $1"
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e $3 > gpurun_out/multi/b_$1.json 2> gpurun_out/multi/b_$1.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/multi/b_$1.json') if l.startswith('{')][-1]); print('n_gpus',d['n_gpus'],'value %.4g'%d['value'],'ms/step %.3f'%d['ms_per_step'],'lane ms %.3f'%d['roofline']['kernel_ms'])" || tail -3 gpurun_out/multi/b_$1.err
}


run kern_r8 "NCCL_P2P_USE_CUDA_MEMCPY=0" "--sm-reserve 8"
run kern_r2 "NCCL_P2P_USE_CUDA_MEMCPY=0" "--sm-reserve 2"
