#!/usr/bin/env python3
"""bench.py -- positions tested per second (KS + weighted Stouffer) on synthetic pileups.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (N>1 under torchrun, one rank per
GPU) prints ONE JSON line on rank 0.  A "step" is one pass of the detection stage (coverage
filter, per-position KS test, window combination) over one synthetic pileup:
BASELINE.json configs[1] -- E. coli K-12 scale, 4.6 Mb, 2x100x, float32 Gaussian currents with
planted shifted sites.

  value      whole-job positions/s with the pileup already resident in HBM.  Steps go through the
             queued entry (nm_detect_device_async / nm_detect_finish): the host stays one step ahead
             of the device, every step is validated (finished) inside the timed region.  N = 1:
             ``Detector.detect_device_async``.  N > 1 (weak scaling: every rank holds its own 4.6 Mb
             shard plus a halo): the product's sharded path, ``ShardedDetector.detect_shard_async`` /
             ``finish_shard`` -- the table stays sharded, each rank selects the head of its own ranking
             behind the step's kernels, and the selection kernels store it into every rank's buffer
             over NVLink (peer-mapped HBM; ``--nccl-heads``, or a node without peer mapping: one NCCL
             all-gather per step): the only communication, completed inside the timed region.
  e2e        the same through the host-facing API with pinned HOST buffers: H2D + kernels + D2H
             (+ the head exchange at N > 1) in the timed region
  roofline   the lane kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the reference's OWN code (oracle/_ref: bin/scripts/myDetect.py rendered to Python 3,
             scipy-1.2.1 call semantics) on one host core, on a bounded sample of the same
             workload, doing the same tests as the GPU arm (KS + combination; the reference's
             mannwhitneyu / ttest_ind calls are switched off for it, see ``tests``)
  variants   (N = 1) the other BASELINE configs through the same call, a few steps each:
             all tests (U + Welch t + KS, Fisher + Stouffer), Poisson coverage, chr20 2x30x, 2x2000x
``--impl reference`` times the reference's code on all host cores (one process per core), primary
value with the same tests as the GPU arm, the all-tests rate next to it.
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
HEAD_WANT = 1024           # rows of each rank's ranking head exchanged per step at N > 1
HEAD_CAP = 4096            # record capacity of the exchange buffer (48-byte records)
SM_RESERVE = 2             # NCCL heads only (--nccl-heads / no peer mapping): SMs the persistent lane kernel leaves to NCCL


def workload_config(n_gpus: int, positions: int = GENOME):
    return {"workload": ("E. coli K-12 scale synthetic pileup, %d positions/GPU, 2x%dx, KS + weighted Stouffer window +-%d"
                         if positions == GENOME and COV == 100 else
                         "synthetic pileup (non-default size: --positions / --coverage), %d positions/GPU, 2x%dx, KS + weighted "
                         "Stouffer window +-%d") % (positions, COV, NB),
            "positions_per_gpu": positions, "coverage": [COV, COV], "neighborPvalues": NB, "WeightsDif": WEIGHTS_DIF,
            "MinCoverage": MIN_COV, "testMethod": "stouffer",
            "tests": "KS test + weighted Stouffer combination per position (want_u = want_t = 0: the Mann-Whitney U "
                     "and Welch t columns of the reference's table are NOT computed in this configuration; "
                     "`variants.all_tests` is the run that fills them). Both arms do exactly this work.",
            "want_u": False, "want_t": False,
            "values": ("Gaussian currents (+ shifted means at planted sites) rounded to three decimals, as the reference "
                       "stores them (norm_mean = round(x, 3), myRefBaseSignalAnnotation.py:1108), float32 in HBM: the lane "
                       "tier checks every value on the device and sorts such tiles as packed 16-bit keys "
                       "(`grid_tiles`); `variants.off_grid_values` is the same run on raw float32 normals" if GRID else
                       "raw float32 Gaussian currents (--off-grid): not the three-place decimals the reference stores"),
            "parallelism": ("1 GPU" if n_gpus == 1 else
                            "genome shards x%d (weak scaling), halo of 10 candidates recomputed per side; the table stays "
                            "sharded; per step each rank's ranking head (>= %d rows) is selected on the device behind the step's own "
                            "kernels and exchanged by the selection kernels themselves, which store it into every rank's buffer over "
                            "NVLink (peer-mapped HBM; `head_exchange` says whether that or the NCCL all-gather fallback ran); sorting "
                            "the gathered heads into the global ranking is host-side ranking work and, like all ranking at N = 1, not "
                            "part of the timed step; every exchange completes inside the timed region" % (n_gpus, HEAD_WANT)),
            "queue": "steps are issued through nm_detect_device_async, the host one step ahead of the device (two result tables alternate)",
            "l2": "inputs (%.1f GB/GPU) are larger than the 126 MB L2; no explicit flush" % (positions * 2 * COV * 4 / 1e9)}


# ---------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------
def planted_shift_np(pos: np.ndarray) -> np.ndarray:
    from nanomod_b200.pileup import planted_shift
    return planted_shift(pos)


GRID = True  # values as the reference stores them: three-place decimals (see to_grid_); --off-grid: raw float32 normals


def to_grid_(v):
    """In place: every value becomes float32(round(float64(x), 3)) -- what the reference's annotation stage
    writes (norm_mean = round(x, 3), bin/scripts/myRefBaseSignalAnnotation.py:1108) and its packer casts."""
    import torch
    step = 1 << 26
    for lo in range(0, v.numel(), step):
        c = v[lo:lo + step]
        c.copy_((torch.round(c.double() * 1000.0) / 1000.0).float())
    return v


def make_device_workload(length: int, n0: int, n1: int, device, seed: int = SEED, pos0: int = 0, grid=None):
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
    if GRID if grid is None else grid:
        to_grid_(v0)
        to_grid_(v1)
    off0 = torch.arange(length + 1, dtype=torch.int64, device=device) * n0
    off1 = torch.arange(length + 1, dtype=torch.int64, device=device) * n1
    posd = torch.from_numpy(pos.astype(np.int32)).to(device)
    seg = torch.zeros(length, dtype=torch.int32, device=device)
    return nm.DevicePileup(v0, off0, v1, off1, posd, seg, length), shift


def host_sample_pileup(length: int, seed: int = SEED):
    import nanomod_b200 as nm
    return nm.synthetic_pileup(length, COV, COV, seed=seed, round_decimals=3 if GRID else None)


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref), else the restated oracle
# ---------------------------------------------------------------------------------------------
def _cpu_run(args):
    """one bounded sample: returns (rows, seconds, kind)"""
    length, seed, all_tests = args
    import copy
    p = host_sample_pileup(length, seed)
    d0, d1 = p.to_dicts()
    kw = dict(MinCoverage=MIN_COV, neighborPvalues=NB, WeightsDif=WEIGHTS_DIF, testMethod="stouffer")
    from oracle import ref_loader as rl
    if rl.available():
        rl.set_tests(all_tests)
        mo = rl.default_moptions(**kw)
        mo["g0"], mo["g1"] = d0, d1
        t0 = time.perf_counter()
        rl.run_detect(mo)
        return len(mo["sign_test"]), time.perf_counter() - t0, "reference"
    from oracle import nanomod_oracle as o
    mo = o.default_moptions(SaveTest=0, **kw)
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    if not all_tests:
        mo["_ks_only"] = 1
    t0 = time.perf_counter()
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    return len(mo["sign_test"]), time.perf_counter() - t0, "port"


def _cpu_what(kind: str) -> str:
    return ("the reference's own bin/scripts/myDetect.py (oracle/_ref: Python-3 rendering, scipy-1.2.1 call semantics): "
            "mfilter_coverage + mtest2" if kind == "reference" else
            "oracle port of myDetect.py:301-462 (oracle/_ref not built): mfilter_coverage + mtest2")


def cpu_baseline_one_core(sample: int):
    rows, sec, kind = _cpu_run((sample, SEED, False))
    rows_all, sec_all, _ = _cpu_run((max(sample // 4, 1000), SEED + 1, True))
    return {"value": rows / sec, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d positions of the same synthetic workload (2x%dx), %s, same tests as the GPU arm "
                      "(KS + Stouffer; U / t calls switched off), %.1f s" % (sample, COV, _cpu_what(kind), sec),
            "all_tests_value": rows_all / sec_all,
            "all_tests_note": "the reference as it actually runs (mannwhitneyu + ttest_ind + ks_2samp per position), "
                              "%d positions, %.1f s; compare with variants.all_tests" % (rows_all, sec_all)}


def run_reference_arm(args):
    """--impl reference: the reference's code on every host core; each step = a bounded sample."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 1500
    ctx = mp.get_context("fork")
    times, rows_total, kind = [], 0, "reference"
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_run, [(per_core, SEED + 17 * step + c, False) for c in range(cores)])
            dt = time.perf_counter() - t0
            kind = res[0][2]
            if step >= args.warmup:
                times.append(dt)
                rows_total += sum(r for r, _, _ in res)
        t0 = time.perf_counter()
        res_all = pool.map(_cpu_run, [(per_core // 3, SEED + 991 + c, True) for c in range(cores)])
        all_value = sum(r for r, _, _ in res_all) / (time.perf_counter() - t0)
    total = sum(times)
    value = rows_total / total
    sample = "%d positions per step (%d per core x %d cores) of the 2x%dx workload; %s" % (
        per_core * cores, per_core, cores, COV, _cpu_what(kind))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
            "ms_per_step_median": 1e3 * statistics.median(times), "ms_per_step_best": 1e3 * min(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "all_tests_value": all_value,
                             "all_tests_note": "the same code with its mannwhitneyu + ttest_ind calls left on (what "
                                               "NanoMod actually runs per position), all %d cores" % cores},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = NanoMod's own myDetect.py hot path executed from oracle/_ref on %d host cores "
                    "(one process per core; the reference itself is single-threaded)" % cores}
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


def lane_capture(kernel):
    """the committed ncu --set full capture of the lane kernel on this workload
    (profiles/lane_kernel_traffic.json, one entry per kernel): DRAM bytes per launch + what actually
    bounds the kernel"""
    try:
        with open(os.path.join(ROOT, "profiles", "lane_kernel_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def time_config(det, dev, opt, length, steps, warmup, nvals, bytes_out):
    """a few steps of another configuration through the same device-resident call"""
    import torch
    import nanomod_b200 as nm
    out = nm.alloc_device_table(opt, length, dev.vals0.device)
    for _ in range(warmup):
        det.detect_device(dev, opt, out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    tms = {"plan": 0.0, "lane": 0.0, "deep": 0.0, "combine": 0.0}
    ev[0].record()
    for k in range(steps):
        rows = det.detect_device(dev, opt, out)
        for name, v in det.handle.last_timings().items():
            tms[name] += v / steps
        ev[k + 1].record()
    torch.cuda.synchronize()
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    ms = sum(per) / steps
    alg = 4 * nvals + (16 + bytes_out) * length
    peak, _ = hbm_peak()
    main = max(tms["lane"], tms["deep"])
    del out
    return {"positions": length, "rows": rows, "steps": steps, "ms_per_step": ms, "ms_per_step_median": statistics.median(per),
            "ms_per_step_best": min(per), "positions_per_s": length / (ms * 1e-3), "kernel_ms": tms,
            "algorithmic_bytes": alg, "tests_kernel_frac_of_hbm_peak": alg / (main * 1e-3) / 1e9 / peak,
            "whole_step_frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak, "path": det.handle.last_path(),
            "grid_tiles": det.handle.last_grid_tiles(), "tiles": (length + 31) // 32}


def run_variants(det, device):
    """the other BASELINE configs (parity-test cases, not bench lines), N = 1 only"""
    import torch
    import nanomod_b200 as nm
    out = {}
    ks_st = nm.DetectOptions(neighborPvalues=NB, testMethod="stouffer", want_u=False, want_t=False, MinCoverage=MIN_COV)
    dev, _ = make_device_workload(GENOME, COV, COV, device)
    allv = nm.DetectOptions(neighborPvalues=NB, both_combinations=True, want_u=True, want_t=True, MinCoverage=MIN_COV)
    out["all_tests"] = time_config(det, dev, allv, GENOME, 5, 3, GENOME * 2 * COV, 76)
    out["all_tests"]["what"] = ("BASELINE configs[2]: the same pileup, U + Welch t + KS per position, Fisher AND Stouffer: "
                                "every column of the reference's table (892 algorithmic B/position)")
    del dev
    # Poisson coverage
    g = torch.Generator(device=device)
    g.manual_seed(7)
    lam = torch.full((GENOME,), float(COV), device=device)
    c0 = torch.poisson(lam, generator=g).clamp_(5, 128).long()
    c1 = torch.poisson(lam, generator=g).clamp_(5, 128).long()
    from nanomod_b200._lib import padded_len
    off0 = torch.zeros(GENOME + 1, dtype=torch.int64, device=device)
    off1 = torch.zeros(GENOME + 1, dtype=torch.int64, device=device)
    off0[1:] = torch.cumsum(c0, 0)
    off1[1:] = torch.cumsum(c1, 0)
    nv = int(off0[-1]) + int(off1[-1])
    v0 = torch.empty(padded_len(int(off0[-1])), dtype=torch.float32, device=device).normal_(generator=g)
    v1 = torch.empty(padded_len(int(off1[-1])), dtype=torch.float32, device=device).normal_(generator=g)
    if GRID:
        to_grid_(v0)
        to_grid_(v1)
    dev = nm.DevicePileup(v0, off0, v1, off1, torch.arange(GENOME, dtype=torch.int32, device=device),
                          torch.zeros(GENOME, dtype=torch.int32, device=device), GENOME)
    out["poisson_coverage"] = time_config(det, dev, ks_st, GENOME, 5, 3, nv, 28)
    out["poisson_coverage"]["what"] = "E. coli scale, coverage ~ Poisson(100) clipped to [5, 128] per group, KS + Stouffer"
    del dev, v0, v1, off0, off1, c0, c1, lam
    dev, _ = make_device_workload(50_000, 2000, 2000, device)
    out["deep_2x2000x"] = time_config(det, dev, ks_st, 50_000, 5, 2, 50_000 * 4000, 28)
    out["deep_2x2000x"]["what"] = "BASELINE configs[4]: 50 kb plasmid, 2x2000x, KS + Stouffer (deep tier)"
    del dev
    torch.cuda.empty_cache()
    L = 64_444_167
    dev, _ = make_device_workload(L, 30, 30, device)
    out["chr20_2x30x_1gpu"] = time_config(det, dev, ks_st, L, 3, 2, L * 60, 28)
    out["chr20_2x30x_1gpu"]["what"] = ("BASELINE configs[3] on ONE GPU: human chr20, 64 444 167 positions, 2x30x, KS + Stouffer "
                                       "(the sharded 2/4/8-GPU run of it: profiles/round2_cfg4_strong_scaling.json)")
    del dev
    torch.cuda.empty_cache()
    # last: a call on off-grid data makes the handle skip the grid-key attempt for its next 15 calls
    if GRID:
        dev, _ = make_device_workload(GENOME, COV, COV, device, grid=False)
        out["off_grid_values"] = time_config(det, dev, ks_st, GENOME, 5, 3, GENOME * 2 * COV, 28)
        out["off_grid_values"]["what"] = ("the headline workload with raw float32 normals instead of three-place decimals: "
                                          "every warp's first tiles fail the grid check, the float32 sort does the work")
        del dev
        torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import nanomod_b200 as nm
    from nanomod_b200.sharded import ShardedDetector, shard_halo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cpus = None
    if world > 1:
        # one process per GPU: stay on the CPUs (and so the host memory) of the GPU's own NUMA node
        from nanomod_b200.sharded import bind_to_gpu_cpus
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[local_rank]) if visible and visible.replace(",", "").isdigit() else local_rank
        cpus = bind_to_gpu_cpus(phys)
        dist.init_process_group("nccl", device_id=device)
    det = nm.Detector(local_rank)
    sd = ShardedDetector(det)
    opt = nm.DetectOptions(MinCoverage=MIN_COV, neighborPvalues=NB, WeightsDif=WEIGHTS_DIF, testMethod="stouffer",
                           want_u=False, want_t=False, SaveTest=0)
    L = args.positions
    halo = shard_halo(opt) if world > 1 else 0
    halo_lo = halo if rank > 0 else 0
    halo_hi = halo if rank < world - 1 else 0
    n_local = L + halo_lo + halo_hi
    dev, _shift = make_device_workload(n_local, COV, COV, device, seed=SEED + rank, pos0=rank * L - halo_lo)
    # Steps are queued through the asynchronous entry (nm_detect_device_async): the host stays one step ahead of
    # the device (step k is finished -- its summary looked at -- after step k+1 has been queued), so the device
    # never idles between steps; two result tables alternate.  At N > 1 the exchange of a step's ranking heads
    # is fused into their selection: the selection kernels, launched behind the step's own kernels, store the
    # head into every rank's buffer over NVLink (peer-mapped HBM, nanomod_b200.sharded.HeadExchange) -- no
    # collective kernel, no SMs set aside.  Without peer mapping the heads go through an NCCL all-gather instead.
    outs = [nm.alloc_device_table(opt, n_local, device) for _ in range(2)]
    step_tm = {"plan": 0.0, "lane": 0.0, "deep": 0.0, "combine": 0.0}
    gathered = [None]
    nccl_pending = [None, None]
    in_flight = []
    issued = [0]
    last_res = [None]
    rows_seen = [0]
    xchg = None
    if world > 1:
        xchg = None if args.nccl_heads else sd.peer_exchange(HEAD_CAP, device)
        if xchg is None or not xchg.ok:
            xchg = None
            det.handle.set_sm_limit(max(1, det.handle.sm_count - args.sm_reserve))
    exchange = "peer" if xchg is not None else "nccl"

    def nccl_drain(b):
        if nccl_pending[b] is not None:
            nccl_pending[b][1].wait()
            gathered[0] = nccl_pending[b][0]
            nccl_pending[b] = None

    def issue():
        b = issued[0] & 1
        issued[0] += 1
        if world == 1:
            in_flight.append((b, det.detect_device_async(dev, opt, outs[b])))
        else:
            nccl_drain(b)  # NCCL mode: this buffer pair's previous exchange (two steps ago)
            in_flight.append((b, sd.detect_shard_async(dev, halo_lo, halo_lo + L, rank * L - halo_lo, opt, outs[b],
                                                       head_want=HEAD_WANT, head_cap=HEAD_CAP, slot=b, peer=xchg is not None)))

    def complete():
        b, pend = in_flight.pop(0)
        if world == 1:
            rows_seen[0], _ = det.detect_finish(pend)
        else:
            res = sd.finish_shard(pend)
            rows_seen[0] = res.n_rows
            last_res[0] = res
            if xchg is None:
                nccl_pending[b] = sd.gather_heads(res, HEAD_WANT, cap=HEAD_CAP, slot=b, async_op=True)
            elif res.head_slot is None:  # the armed selection did not run with the step: select + store now
                sd.gather_heads(res, HEAD_WANT, cap=HEAD_CAP, slot=b, async_op=True, peer=True)
        for k, v in det.handle.last_timings().items():
            step_tm[k] = v

    def step():
        issue()
        if len(in_flight) > 1:
            complete()

    def flush():
        while in_flight:
            complete()
        nccl_drain(0)
        nccl_drain(1)

    def fence():
        flush()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the first calls establish the shape synchronously; then W warm-up steps through the queue
    for b in range(2):
        det.detect_device(dev, opt, outs[b])
    for _ in range(args.warmup):
        step()
    fence()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lane_ms, comb_ms, plan_ms = [], [], []
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches1 = det.launch_count
    evs[0].record()
    for k in range(args.steps):
        step()
        evs[k + 1].record()
        if k > 0:
            lane_ms.append(step_tm["lane"])
            comb_ms.append(step_tm["combine"])
            plan_ms.append(step_tm["plan"])
    flush()  # the last step's validation (and, NCCL mode, the last exchanges) belong to the timed region
    lane_ms.append(step_tm["lane"])
    comb_ms.append(step_tm["combine"])
    plan_ms.append(step_tm["plan"])
    if world > 1:
        evs[args.steps].record()
    fence()
    rows = rows_seen[0]
    if world > 1 and xchg is not None:
        # after the fence every rank's kernels are done: every section of the last step's slot carries its epoch
        slot = (issued[0] - 1) & 1
        ep = xchg.epochs(slot)
        assert np.all(ep == last_res[0].head_epoch), (ep, last_res[0].head_epoch)
        gathered[0] = xchg.gathered(slot).clone()
    per_step = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    ms_total = evs[0].elapsed_time(evs[args.steps])
    launches = det.launch_count - launches1
    path = det.handle.last_path()
    grid_tiles = det.handle.last_grid_tiles()
    assert rows == n_local, (rows, n_local)
    t = torch.tensor([ms_total, statistics.median(per_step), min(per_step)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_median, ms_best = (float(x) for x in t.tolist())
    value = world * L * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the host-facing API (H2D + kernels + D2H every step)
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

        hout_t = {c: torch.from_numpy(hout[c]) for c in cols}

        e2e_no = [0]

        def e2e_step():
            if world == 1:
                tbl = det.detect(hp, opt, out=hout)  # nm_detect_host: pinned host CSR in, result columns out
                _ = float(tbl.stouffer_p[0])         # the result is on the host
            else:
                # the sharded product path from host buffers: H2D of the shard, detect_shard, heads
                # all-gathered, this rank's rows back to (pinned) host memory
                d = nm.DevicePileup.from_host(hp, device)
                res = sd.detect_shard(d, halo_lo, halo_lo + L, rank * L - halo_lo, opt, outs[0], head_want=HEAD_WANT,
                                      head_cap=HEAD_CAP, slot=e2e_no[0] & 1, peer=xchg is not None)
                sd.gather_heads(res, HEAD_WANT, cap=HEAD_CAP, slot=e2e_no[0] & 1, peer=xchg is not None)
                e2e_no[0] += 1
                for c in cols:
                    hout_t[c][:res.n_core].copy_(res.core(c), non_blocking=True)
                torch.cuda.synchronize()
                _ = float(hout["stouffer_p"][0])

        e2e_step()  # warm-up (allocates the staging buffers)
        fence()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        fence()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d = int(4 * (hp.off0[-1] + hp.off1[-1]) + 8 * 2 * (n_local + 1) + 4 * 2 * n_local)
        d2h = int(sum(hout[c].itemsize for c in cols) * n_local)
        e2e = {"value": world * L * e_steps / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps, "ms_per_step": 1e3 * float(tt.item()) / e_steps,
               "pcie_GBps": (h2d + d2h) / (float(tt.item()) / e_steps) / 1e9,
               "api": ("nanomod_b200.Detector.detect (nm_detect_host): pinned host CSR in, result columns out" if world == 1 else
                       "DevicePileup.from_host (pinned) + ShardedDetector.detect_shard + gather_heads (%s) + the rank's rows to pinned host" % exchange)}
    # ---- the same through the 16-bit transport format (N = 1): the workload's values rounded to the
    # reference's 0.001 grid (norm_mean = round(x, 3)), shipped as int16 milli-units, expanded on the GPU
    e2e_i16 = None
    if not args.no_e2e and world == 1:
        k0 = torch.round(dev.vals0 * 1000.0).clamp_(-32767, 32767).to(torch.int16)
        k1 = torch.round(dev.vals1 * 1000.0).clamp_(-32767, 32767).to(torch.int16)
        pad8 = lambda t: torch.nn.functional.pad(t, (0, (-t.numel()) % 8 + 8))
        hp16 = nm.Pileup(vals0=hp.vals0, off0=hp.off0, vals1=hp.vals1, off1=hp.off1, pos=hp.pos, seg=hp.seg, base=hp.base,
                         seg_names=hp.seg_names, vals0_i16=pad8(k0).cpu().pin_memory().numpy(),
                         vals1_i16=pad8(k1).cpu().pin_memory().numpy(), i16_unit=0.001)
        del k0, k1
        det.detect(hp16, opt, out=hout)
        fence()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            tbl = det.detect(hp16, opt, out=hout)
            _ = float(tbl.stouffer_p[0])
        fence()
        dt16 = (time.perf_counter() - t0) / e_steps
        h2d16 = int(2 * (hp.off0[-1] + hp.off1[-1]) + 8 * 2 * (n_local + 1) + 4 * 2 * n_local)
        e2e_i16 = {"value": L / dt16, "unit": UNIT, "ms_per_step": 1e3 * dt16, "h2d_bytes_per_step": h2d16,
                   "d2h_bytes_per_step": d2h, "pcie_GBps": (h2d16 + d2h) / dt16 / 1e9,
                   "what": "the same workload with its values on the reference's 0.001 grid (norm_mean = round(x, 3)), "
                           "shipped as int16 milli-units (nm_pileup.vals*_i16) and expanded to float32 on the GPU: "
                           "bit-identical tables, half the PCIe bytes"}
    clocks = sampler.stop() if rank == 0 else None
    variants = None
    if rank == 0 and world == 1 and not args.no_variants and L == GENOME:
        del dev
        outs.clear()
        torch.cuda.empty_cache()
        variants = run_variants(det, device)

    if rank == 0:
        peak, peak_src = hbm_peak()
        lane_avg = sum(lane_ms) / len(lane_ms)
        achieved = BYTES_PER_POS * n_local / (lane_avg * 1e-3) / 1e9
        kernel = ("nm_lane_grid_kernel" if grid_tiles > 0 else "nm_lane_dense_kernel") if path in (1, 2, 3) else "nm_lane_kernel"
        cap = lane_capture(kernel) if (world == 1 and L == GENOME) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "ms_per_step_median": ms_median,
                "ms_per_step_best": ms_best, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": ("u16 keys (exact images of the float32 inputs, checked on the device), i32 ranks, f64 tails"
                          if kernel == "nm_lane_grid_kernel" else "f32 keys, i32 ranks, f64 tails"),
                "data": "synthetic", "config": workload_config(world, L),
                "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])) if cap else None,
                             "traffic_source": cap.get("source") if cap else None,
                             "algorithmic_bytes_per_launch": BYTES_PER_POS * n_local,
                             "peak_source": peak_src,
                             "bytes_per_position": BYTES_PER_POS, "positions_per_launch": n_local,
                             "kernel_ms": lane_avg, "kernel_ms_median": statistics.median(lane_ms), "kernel_ms_best": min(lane_ms),
                             "other_kernels_ms": {"plan": sum(plan_ms) / len(plan_ms), "combine": sum(comb_ms) / len(comb_ms)},
                             "step_minus_kernel_ms": ms_total / args.steps - lane_avg,
                             "frac_of_nominal_8TBs": achieved / 8000.0,
                             "practical_bound": "instruction issue on the half-rate ALU / FMA-heavy pipes (ncu, committed capture)",
                             "ncu": ({k: cap[k] for k in ("issue_active_pct", "alu_pipe_active_pct", "fmaheavy_pipe_active_pct",
                                                          "warp_instructions", "registers_per_thread", "warps_active_pct") if k in cap}
                                     if cap else None)},
                "e2e": e2e, "gpu_launches": launches, "grid_tiles": grid_tiles, "tiles": (n_local + 31) // 32, "code_path": {0: "general", 1: "dense", 2: "dense (speculative launch)",
                                                                    3: "dense (re-run)", 4: "general (re-run)"}.get(path, str(path)),
                "clocks": clocks,
                "pct_hbm_peak_whole_step": 100.0 * (BYTES_PER_POS * value / world / 1e9) / peak}
        if world > 1:
            heads = sd.heads_from_gathered(gathered[0], opt, HEAD_CAP)
            line["head_rows_exchanged"] = [int(h.rows.shape[0]) for h in heads]
            line["head_exchange"] = exchange
            line["rank0_cpu_affinity"] = ("%d CPUs of the GPU's NUMA node" % len(cpus)) if cpus else "unchanged"
        if variants is not None:
            line["variants"] = variants
        if e2e_i16 is not None:
            line.setdefault("variants", {})["e2e_int16_grid"] = e2e_i16
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_one_core(args.cpu_sample)
            line["cpu_baseline"]["host_cores_available"] = os.cpu_count()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global GRID, COV, BYTES_PER_POS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--positions", type=int, default=GENOME, help="positions per GPU (default: E. coli scale)")
    ap.add_argument("--coverage", type=int, default=COV,
                    help="reads per group and position (default 100; 30 with --positions 64444167/N is BASELINE configs[3] "
                         "split over N GPUs: the strong-scaling experiment of profiles/round2_cfg4_strong_scaling.json)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--sm-reserve", type=int, default=SM_RESERVE, help="N > 1, NCCL heads: SMs left free for NCCL")
    ap.add_argument("--nccl-heads", action="store_true",
                    help="N > 1: exchange the ranking heads by an NCCL all-gather instead of peer-memory stores")
    ap.add_argument("--off-grid", action="store_true",
                    help="raw float32 normals instead of the reference's three-place decimals")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=40000)
    args = ap.parse_args()
    GRID = not args.off_grid
    COV = args.coverage
    BYTES_PER_POS = 4 * (COV + COV) + 16 + 28
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
