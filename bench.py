#!/usr/bin/env python3
"""bench.py -- positions tested per second (KS + weighted Stouffer) on synthetic pileups.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (N>1 under torchrun, one rank per
GPU) prints ONE JSON line on rank 0.  A "step" is one pass of the detection stage (coverage
filter, per-position KS test, window combination) over one synthetic pileup:
BASELINE.json configs[1] -- E. coli K-12 scale, 4.6 Mb, 2x100x, float32 Gaussian currents with
planted shifted sites -- per GPU (weak scaling: rank r holds its own 4.6 Mb shard
plus a halo of neighborPvalues positions per side; the only communication is the NCCL gather of
the 28-byte result records to rank 0, inside the timed region).

  value      whole-job positions/s with the pileup already resident in HBM (nm_detect_device)
  e2e        the same through Detector.detect(): pinned HOST buffers in, H2D + kernels + D2H
  roofline   nm_lane_kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the scalar oracle (the reference's scipy-1.2.1 arithmetic restated) on one host
             core, on a bounded sample of the same workload
``--impl reference`` times that oracle on all host cores instead (the reference itself is
Python-2-only and cannot run here -- DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "positions tested/sec (KS+Stouffer, 2x100x)"
UNIT = "positions/s"
GENOME = 4_600_000
COV = 100
NB = 3
WEIGHTS_DIF = 2.0
MIN_COV = 5
SEED = 20190131
BYTES_PER_POS = 4 * (COV + COV) + 16 + 28  # fp32 values + two int64 offsets + outputs
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def workload_config(n_gpus: int):
    return {"workload": "E. coli K-12 scale synthetic pileup, %d positions/GPU, 2x%dx, KS + weighted "
                        "Stouffer window +-%d" % (GENOME, COV, NB),
            "positions_per_gpu": GENOME, "coverage": [COV, COV], "neighborPvalues": NB, "WeightsDif": WEIGHTS_DIF,
            "MinCoverage": MIN_COV, "testMethod": "stouffer", "tests": "ks",
            "parallelism": "genome shards x%d, halo %d; NCCL gather of step k's result records to rank 0 overlaps the compute of step k+1" % (n_gpus, NB),
            "l2": "inputs (3.7 GB/GPU) are larger than the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------
def planted_shift_np(pos: np.ndarray) -> np.ndarray:
    from nanomod_b200.pileup import planted_shift
    return planted_shift(pos)


def make_device_workload(length: int, n0: int, n1: int, device, seed: int = SEED, pos0: int = 0):
    """Fixed-coverage synthetic pileup generated on the GPU (torch Philox generator).
    Returns (DevicePileup, per-position planted shift as a numpy array)."""
    import torch
    import nanomod_b200 as nm
    from nanomod_b200._lib import padded_len
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pos = np.arange(pos0, pos0 + length, dtype=np.int64)
    shift = planted_shift_np(pos)
    v0 = torch.empty(padded_len(length * n0), dtype=torch.float32, device=device)
    v1 = torch.empty(padded_len(length * n1), dtype=torch.float32, device=device)
    v0.normal_(generator=g)
    v1.normal_(generator=g)
    sh = torch.from_numpy(shift.astype(np.float32)).to(device)
    v1[: length * n1].view(length, n1).add_(sh[:, None])
    off0 = torch.arange(length + 1, dtype=torch.int64, device=device) * n0
    off1 = torch.arange(length + 1, dtype=torch.int64, device=device) * n1
    posd = torch.from_numpy(pos.astype(np.int32)).to(device)
    seg = torch.zeros(length, dtype=torch.int32, device=device)
    return nm.DevicePileup(v0, off0, v1, off1, posd, seg, length), shift


def host_sample_pileup(length: int, seed: int = SEED):
    import nanomod_b200 as nm
    return nm.synthetic_pileup(length, COV, COV, seed=seed)


# ---------------------------------------------------------------------------------------------
# CPU baseline: the scalar oracle (same scipy-level calls per position as myDetect.py:331-401)
# ---------------------------------------------------------------------------------------------
def _oracle_run(args):
    length, seed = args
    from oracle import nanomod_oracle as o
    p = host_sample_pileup(length, seed)
    d0, d1 = p.to_dicts()
    mo = o.default_moptions(MinCoverage=MIN_COV, neighborPvalues=NB, WeightsDif=WEIGHTS_DIF, testMethod="stouffer")
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    t0 = time.perf_counter()
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    return len(mo["sign_test"]), time.perf_counter() - t0


def cpu_baseline_one_core(sample: int = 20000):
    rows, sec = _oracle_run((sample, SEED))
    return {"value": rows / sec, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d positions of the same synthetic workload (2x%dx), scalar oracle: "
                      "mfilter_coverage + mtest2 (U, t, KS per position + Stouffer), %.1f s" % (sample, COV, sec)}


def run_reference_arm(args):
    """--impl reference: the oracle on every host core; each step = a bounded sample."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 1500
    ctx = mp.get_context("fork")
    times = []
    rows_total = 0
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_oracle_run, [(per_core, SEED + 17 * step + c) for c in range(cores)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
                rows_total += sum(r for r, _ in res)
    total = sum(times)
    value = rows_total / total
    sample = "%d positions per step (%d per core x %d cores) of the 2x%dx workload" % (per_core * cores, per_core, cores, COV)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = scalar CPU oracle (scipy-1.2.1 formulas); the Python-2 reference cannot run here"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------
def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def lane_traffic():
    """DRAM bytes per nm_lane_kernel launch on the bench workload, from the committed ncu
    --set full capture (profiles/lane_kernel_traffic.json); None if no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "lane_kernel_traffic.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), d.get("source")
    except Exception:
        return None, None


def lane_ncu_summary():
    """What actually bounds the lane kernel, from the same committed ncu capture: the path is
    HBM-bound by contract, but the kernel is limited by instruction issue on the half-rate integer
    pipes -- reported next to the HBM roofline so that `frac` is read correctly."""
    try:
        with open(os.path.join(ROOT, "profiles", "lane_kernel_traffic.json")) as f:
            d = json.load(f)
        return {k: d[k] for k in ("issue_active_pct", "alu_pipe_active_pct", "fmaheavy_pipe_active_pct",
                                  "warp_instructions", "registers_per_thread", "warps_active_pct") if k in d}
    except Exception:
        return None


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import nanomod_b200 as nm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    det = nm.Detector(local_rank)
    opt = nm.DetectOptions(MinCoverage=MIN_COV, neighborPvalues=NB, WeightsDif=WEIGHTS_DIF, testMethod="stouffer",
                           want_u=False, want_t=False)
    L = args.positions
    halo_lo = NB if rank > 0 else 0
    halo_hi = NB if rank < world - 1 else 0
    n_local = L + halo_lo + halo_hi
    dev, _shift = make_device_workload(n_local, COV, COV, device, seed=SEED + rank, pos0=rank * L - halo_lo)
    out = nm.alloc_device_table(opt, n_local, device)
    rec_cols = ["ks_dnum", "ks_p", "stouffer_stat", "stouffer_p"]  # the 28-byte result record
    outs = [out]
    gbufs = []
    pending = [None, None]
    if world > 1:
        # Multi-GPU step: every rank computes its shard, then the 28-byte result records go to
        # rank 0 over NCCL.  Outputs are double-buffered and the gather of step k runs on a side
        # stream while step k+1 is computed; the lane kernel leaves a few SMs free for NCCL's
        # copy kernels.  All gathers complete inside the timed region.
        comm = torch.cuda.Stream(device=device)
        det.handle.set_sm_limit(max(1, det.handle.sm_count - args.sm_reserve))
        outs.append(nm.alloc_device_table(opt, n_local, device))
        rec_w = [0, 4, 12, 20, 28]  # byte offsets of ks_dnum, ks_p, stouffer_stat, stouffer_p in a record
        recs = [torch.empty((L, 28), dtype=torch.uint8, device=device) for _b in range(2)]
        for _b in range(2):
            gbufs.append([torch.empty((L, 28), dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None)
    step_tm = {"plan": 0.0, "lane": 0.0, "deep": 0.0, "combine": 0.0}
    step_no = [0]

    def drain(b):
        if pending[b] is not None:
            for w in pending[b]:
                w.wait()
            pending[b] = None

    def step():
        b = step_no[0] & 1 if world > 1 else 0
        step_no[0] += 1
        if world > 1:
            drain(b)  # the buffer's previous gather (two steps ago) must have finished
        rows = det.detect_device(dev, opt, outs[b])  # returns with the results complete on the device
        for k, v in det.handle.last_timings().items():
            step_tm[k] = v
        if world > 1:
            # pack the four result columns into 28-byte records (one kernel), then ONE gather
            det.pack_records(outs[b], halo_lo, L, opt, recs[b])
            comm.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(comm):
                pending[b] = [dist.gather(recs[b], gbufs[b], dst=0, async_op=True)]
        return rows

    def fence():
        if world > 1:
            drain(0)
            drain(1)
            torch.cuda.synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = det.launch_count
    for _ in range(args.warmup):
        step()
    fence()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lane_ms, comb_ms, plan_ms = [], [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = det.launch_count
    ev0.record()
    rows = 0
    for _ in range(args.steps):
        rows = step()
        lane_ms.append(step_tm["lane"])
        comb_ms.append(step_tm["combine"])
        plan_ms.append(step_tm["plan"])
    if world > 1:  # the last two gathers belong to the timed region
        drain(0)
        drain(1)
    ev1.record()
    fence()
    ms_total = ev0.elapsed_time(ev1)
    launches = det.launch_count - launches1
    assert rows == n_local, (rows, n_local)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * L * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the public API (H2D + kernels + D2H every step)
    e2e = None
    if not args.no_e2e:
        pin = lambda x: x.cpu().pin_memory().numpy()
        hp = nm.Pileup(vals0=pin(dev.vals0), off0=pin(dev.off0), vals1=pin(dev.vals1), off1=pin(dev.off1),
                       pos=pin(dev.pos), seg=pin(dev.seg), base=np.zeros(n_local, np.uint8), seg_names=[("syn", "+")])
        from nanomod_b200.detect import _wanted_columns
        from nanomod_b200 import _lib
        cols = _wanted_columns(opt)
        tdt = {"int32": torch.int32, "int64": torch.int64, "float64": torch.float64, "uint8": torch.uint8}
        hout = {c: torch.empty(n_local, dtype=tdt[_lib.TABLE_DTYPES[c]]).pin_memory().numpy() for c in cols}
        e_steps = max(1, min(args.steps, args.e2e_steps))
        det.detect(hp, opt, out=hout)  # warm-up (allocates the staging buffers)
        fence()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            tbl = det.detect(hp, opt, out=hout)
            _ = float(tbl.stouffer_p[0])  # the result is on the host
        fence()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d = int(4 * (hp.off0[-1] + hp.off1[-1]) + 8 * 2 * (n_local + 1) + 4 * 2 * n_local)
        d2h = int(sum(hout[c].itemsize for c in cols) * n_local)
        e2e = {"value": world * L * e_steps / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps,
               "api": "nanomod_b200.Detector.detect (nm_detect_host): pinned host CSR in, result columns out"}
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = hbm_peak()
        lane_avg = sum(lane_ms) / len(lane_ms)
        achieved = BYTES_PER_POS * n_local / (lane_avg * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 keys, i32 ranks, f64 tails",
                "data": "synthetic", "config": workload_config(world),
                "roofline": {"bound": "hbm", "kernel": "nm_lane_kernel", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (lane_traffic()[0] if world == 1 and L == GENOME else None),
                             "traffic_source": lane_traffic()[1], "algorithmic_bytes_per_launch": BYTES_PER_POS * n_local,
                             "peak_source": peak_src,
                             "bytes_per_position": BYTES_PER_POS, "positions_per_launch": n_local,
                             "kernel_ms": lane_avg, "other_kernels_ms": {"plan": sum(plan_ms) / len(plan_ms),
                                                                         "combine": sum(comb_ms) / len(comb_ms)},
                             "frac_of_nominal_8TBs": achieved / 8000.0,
                             "practical_bound": "instruction issue on the ALU / FMA-heavy pipes (ncu, committed capture)",
                             "ncu": (lane_ncu_summary() if world == 1 and L == GENOME else None)},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "pct_hbm_peak_whole_step": 100.0 * (BYTES_PER_POS * value / world / 1e9) / peak}
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_one_core(args.cpu_sample)
            line["cpu_baseline"]["host_cores_available"] = os.cpu_count()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--positions", type=int, default=GENOME, help="positions per GPU (default: E. coli scale)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--sm-reserve", type=int, default=2, help="SMs left free for NCCL while computing (N > 1)")
    ap.add_argument("--cpu-sample", type=int, default=60000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
