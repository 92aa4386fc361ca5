#!/usr/bin/env python3
"""Time the head selection of the multi-GPU exchange (three kernels) on one GPU: CUDA events around
ShardedDetector.gather_heads at world size 1 (no NCCL), bench workload."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nanomod_b200 as nm
from nanomod_b200.sharded import ShardedDetector
import bench

det = nm.Detector(0)
sd = ShardedDetector(det)
dev, _ = bench.make_device_workload(bench.GENOME, 100, 100, torch.device("cuda:0"))
opt = nm.DetectOptions(MinCoverage=5, neighborPvalues=3, WeightsDif=2.0, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
out = nm.alloc_device_table(opt, bench.GENOME, torch.device("cuda:0"))
res = sd.detect_shard(dev, 0, bench.GENOME, 0, opt, out)
for _ in range(3):
    sd.gather_heads(res, 1024, cap=4096)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    sd.gather_heads(res, 1024, cap=4096)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"head_select_ms": e0.elapsed_time(e1) / 20}))
