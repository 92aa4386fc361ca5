#!/usr/bin/env python3
"""Lane-kernel sweep over uniform coverage n (2 x n reads, KS + Stouffer): one detect call per n
after a warm-up; meant to run under `ncu --metrics ...` to relate code footprint (network size)
to issue utilisation.  Usage: sweep_n.py [positions] n n n ..."""
import sys
import torch
sys.path.insert(0, ".")
import nanomod_b200 as nm
from bench import make_device_workload

L = int(sys.argv[1])
det = nm.Detector(0)
opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False)
for n in [int(a) for a in sys.argv[2:]]:
    dev, _ = make_device_workload(L, n, n, torch.device("cuda:0"))
    out = nm.alloc_device_table(opt, L, "cuda:0")
    for _ in range(3):
        det.detect_device(dev, opt, out)
    torch.cuda.synchronize()
    print("n", n, {k: round(v, 4) for k, v in det.handle.last_timings().items()}, flush=True)
    del dev, out
