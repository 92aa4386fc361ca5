#!/usr/bin/env python3
"""Time the other BASELINE.json configs (parity-test cases, not bench.py lines) on one GPU:
device-resident nm_detect_device, CUDA-event kernel times from nm_last_timings.
  cfg1  10 kb, 2x50, KS + Stouffer +-3 (latency only: fits in L2)
  cfg3  E. coli scale, 2x100x, U + t + KS, Fisher + Stouffer
  cfg4  chr20 scale (64.4 M positions), 2x30x, KS + Stouffer
  cfg5  50 kb plasmid, 2x2000x, KS + Stouffer (deep tier)
  cfg2p E. coli scale, coverage ~ Poisson(100) clipped to [5,128] (mixed network sizes)
Writes one JSON object per config to stdout."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nanomod_b200 as nm
import bench
from bench import make_device_workload, hbm_peak, to_grid_

if os.environ.get("NM_OFF_GRID") == "1":  # raw float32 normals instead of the reference's three-place decimals
    bench.GRID = False


def poisson_workload(length, mean, device, lo=5, hi=128, seed=7):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lam = torch.full((length,), float(mean), device=device)
    c0 = torch.poisson(lam, generator=g).clamp_(lo, hi).long()
    c1 = torch.poisson(lam, generator=g).clamp_(lo, hi).long()
    return workload_from_counts(c0, c1, device, g)


def workload_from_counts(c0, c1, device, g=None):
    from nanomod_b200._lib import padded_len
    if g is None:
        g = torch.Generator(device=device)
        g.manual_seed(11)
    length = c0.numel()
    off0 = torch.zeros(length + 1, dtype=torch.int64, device=device)
    off1 = torch.zeros(length + 1, dtype=torch.int64, device=device)
    off0[1:] = torch.cumsum(c0, 0)
    off1[1:] = torch.cumsum(c1, 0)
    v0 = torch.empty(padded_len(int(off0[-1])), dtype=torch.float32, device=device).normal_(generator=g)
    v1 = torch.empty(padded_len(int(off1[-1])), dtype=torch.float32, device=device).normal_(generator=g)
    if bench.GRID:
        to_grid_(v0)
        to_grid_(v1)
    pos = torch.arange(length, dtype=torch.int32, device=device)
    seg = torch.zeros(length, dtype=torch.int32, device=device)
    return nm.DevicePileup(v0, off0, v1, off1, pos, seg, length), int(off0[-1]) + int(off1[-1])


def run(name, dev, nvals, length, opt, bytes_out, steps=10, warmup=3):
    det = run.det
    out = nm.alloc_device_table(opt, length, "cuda:0")
    for _ in range(warmup):
        det.detect_device(dev, opt, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tms = {"plan": 0.0, "lane": 0.0, "deep": 0.0, "combine": 0.0}
    e0.record()
    for _ in range(steps):
        rows = det.detect_device(dev, opt, out)
        for k, v in det.handle.last_timings().items():
            tms[k] += v / steps
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    alg_bytes = 4 * nvals + (16 + bytes_out) * length
    peak, src = hbm_peak()
    main = max(tms["lane"], tms["deep"])
    print(json.dumps({"config": name, "positions": length, "rows": rows, "ms_per_step": ms, "positions_per_s": length / (ms * 1e-3),
                      "kernel_ms": tms, "algorithmic_bytes": alg_bytes,
                      "tests_kernel_GBps": alg_bytes / (main * 1e-3) / 1e9, "tests_kernel_frac_of_measured_peak": alg_bytes / (main * 1e-3) / 1e9 / peak,
                      "whole_step_GBps": alg_bytes / (ms * 1e-3) / 1e9, "peak_GBps": peak, "peak_source": src}), flush=True)
    del out


def main():
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg2p", "cfg5", "cfg4"]
    run.det = nm.Detector(0)
    d = torch.device("cuda:0")
    ks_st = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False)
    if "cfg1" in which:
        dev, _ = make_device_workload(10_000, 50, 50, d)
        run("cfg1 10kb 2x50 KS+Stouffer (L2-resident: latency only)", dev, 10_000 * 100, 10_000, ks_st, 28, steps=50)
    if "cfg3" in which:
        dev, _ = make_device_workload(4_600_000, 100, 100, d)
        allv = nm.DetectOptions(neighborPvalues=3, both_combinations=True, want_u=True, want_t=True)
        run("cfg3 E.coli 2x100x U+t+KS, Fisher+Stouffer", dev, 4_600_000 * 200, 4_600_000, allv, 76)
        ku = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=True, want_t=False)
        run("cfg3b E.coli 2x100x U+KS, Stouffer", dev, 4_600_000 * 200, 4_600_000, ku, 44)
        del dev
    if "cfg2p" in which:
        dev, nvals = poisson_workload(4_600_000, 100, d)
        run("cfg2p E.coli Poisson(100) clipped [5,128] KS+Stouffer", dev, nvals, 4_600_000, ks_st, 28)
        del dev
    if "cfg2o" in which:  # mostly 2x100x with 1 % of deeper positions (outliers set the old per-call class)
        g = torch.Generator(device=d).manual_seed(7)
        L = 4_600_000
        c = torch.full((L,), 100, dtype=torch.int64, device=d)
        c[torch.rand(L, generator=g, device=d) < 0.01] = 128
        dev, nvals = workload_from_counts(c, c.clone(), d)
        run("cfg2o E.coli 2x100x with 1% of 2x128x positions KS+Stouffer", dev, nvals, L, ks_st, 28)
        del dev
    if "cfg2x" in which:  # uniform 2x128x: the largest lane-tier class
        dev, _ = make_device_workload(2_000_000, 128, 128, d)
        run("cfg2x 2 Mb 2x128x KS+Stouffer", dev, 2_000_000 * 256, 2_000_000, ks_st, 28)
        del dev
    if "cfg2h" in which:
        dev, _ = make_device_workload(4_600_000, 50, 50, d)
        run("cfg2h E.coli 2x50x KS+Stouffer", dev, 4_600_000 * 100, 4_600_000, ks_st, 28)
        del dev
    if "cfg5" in which:
        dev, _ = make_device_workload(50_000, 2000, 2000, d)
        run("cfg5 50kb plasmid 2x2000x KS+Stouffer (deep tier)", dev, 50_000 * 4000, 50_000, ks_st, 28, steps=5, warmup=2)
        del dev
    if "cfg4" in which:
        torch.cuda.empty_cache()
        L = 64_444_167
        dev, _ = make_device_workload(L, 30, 30, d)
        run("cfg4 chr20 64.4Mb 2x30x KS+Stouffer (1 GPU)", dev, L * 60, L, ks_st, 28, steps=5, warmup=2)


if __name__ == "__main__":
    main()
