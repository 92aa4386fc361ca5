#!/usr/bin/env python3
"""Steps of the bench workload with the head selection armed (world size 1, no exchange): run under
`ncu --metrics gpu__time_duration.sum` to see what the selection kernels cost behind a detect call
(NANOMOD_B200_NO_HEAD_CANDS=1: the three-pass form instead of the combine kernel's candidate list)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nanomod_b200 as nm
from nanomod_b200.sharded import ShardedDetector
import bench

L = int(sys.argv[1]) if len(sys.argv) > 1 else bench.GENOME
det = nm.Detector(0)
sd = ShardedDetector(det)
dev, _ = bench.make_device_workload(L, 100, 100, torch.device("cuda:0"))
opt = nm.DetectOptions(MinCoverage=5, neighborPvalues=3, WeightsDif=2.0, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
out = nm.alloc_device_table(opt, L, torch.device("cuda:0"))
for _ in range(3):
    res = sd.detect_shard(dev, 10, L - 10, 0, opt, out, head_want=1024, head_cap=4096)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    res = sd.detect_shard(dev, 10, L - 10, 0, opt, out, head_want=1024, head_cap=4096)
e1.record()
torch.cuda.synchronize()
hdr = sd._head_mine[0][:48].cpu().numpy().view(nm.sharded.HEAD_REC)
print(json.dumps({"ms_per_step_sync": e0.elapsed_time(e1) / 10, "fired": res.head_slot is not None, "head_rows": int(hdr[0]["row"])}))
