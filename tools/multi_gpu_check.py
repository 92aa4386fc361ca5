#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU).  Checks of the sharded product path (NCCL):
  small   host pileup (gaps, two strands, mixed coverage): ShardedDetector.detect (gather of the whole
          table to rank 0) is byte-identical to the unsharded table; the sharded device path
          (detect_shard + heads) gives the same called sites on every rank; save_test writes one
          file identical to the single-GPU text
  chr20   (--chr20) BASELINE configs[3]: 64 444 167 positions, 2x30x, KS + Stouffer, the genome split
          across the ranks with halos -- STRONG scaling.  Every rank also computes the whole genome
          on its own GPU and compares its shard's rows byte for byte; prints ms per step at this N.
Prints one line per check on rank 0; exit code 1 on any mismatch."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nanomod_b200 as nm
from nanomod_b200.sharded import ShardedDetector, shard_halo

CHR20 = 64_444_167


def small(det, rank, world):
    p = nm.synthetic_pileup(200_000, 40, 40, drop_frac1=0.005, two_strands=True, poisson=True, clip=(3, 100), round_decimals=3)
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True, SaveTest=0)
    sd = ShardedDetector(det)
    table = sd.detect(p, opt)
    dev, core_lo, core_hi, cand_lo = sd.shard_device(p, opt, torch.device("cuda", torch.cuda.current_device()))
    res = sd.detect_shard(dev, core_lo, core_hi, cand_lo, opt)
    sites = sd.called_sites(res, p.seg_names)
    path = os.path.join(tempfile.gettempdir(), "nm_sharded_sign_test.txt")
    base_local = p.base[cand_lo:cand_lo + dev.n_pos]
    sd.save_test(res, p.seg_names, base_local, path)
    ok = True
    if rank == 0:
        full = det.detect(p, opt)
        for c in ("row_pos_index", "n0", "n1", "ks_dnum", "ks_p", "two_u", "u_p", "t_stat", "t_p", "stouffer_stat",
                  "stouffer_p", "fisher_stat", "fisher_p", "pos", "seg"):
            same = getattr(table, c).tobytes() == getattr(full, c).tobytes()
            ok &= same
            if not same:
                print("MISMATCH", c)
        text_same = open(path, "rb").read() == full.format_text()
        ok &= text_same and sites == full.called_sites()
        print("multi_gpu_check small world=%d rows=%d gathered_identical=%s sharded_called_sites_identical=%s "
              "sharded_text_identical=%s called=%s" % (world, len(table), ok, sites == full.called_sites(), text_same, sites[:3]))
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    return int(flag.item()) == 0


def peer_exchange(det, rank, world):
    """The peer-memory head exchange (selection kernels storing into every rank's buffer over NVLink) against
    the NCCL all-gather of the same heads: dense shards (armed selection, synchronous and queued steps) and
    shards with filtered rows (selection after the call)."""
    from nanomod_b200.sharded import HEAD_REC
    device = torch.device("cuda", torch.cuda.current_device())
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
    sd = ShardedDetector(det)
    cap, want = 2048, 300
    x = sd.peer_exchange(cap, device)
    if x is None or not x.ok:
        if rank == 0:
            print("multi_gpu_check peer world=%d SKIPPED: no peer mapping on this node" % world)
        return True

    def sections(buf):
        a = buf.cpu().numpy().view(HEAD_REC).reshape(world, cap + 1).copy()
        out = []
        for r in range(world):
            k = int(a[r, 0]["row"])
            hdr = a[r, 0].copy()
            hdr["reserved"] = 0
            out.append((hdr.tobytes(), np.sort(a[r, 1:k + 1], order=["row"]).tobytes()))
        return out

    bad = 0
    for drop in (0.0, 0.01):
        p = nm.synthetic_pileup(120_000, 70, 66, drop_frac1=drop, round_decimals=3, seed=5)
        dev, core_lo, core_hi, cand_lo = sd.shard_device(p, opt, device)
        outs = [nm.alloc_device_table(opt, dev.n_pos, device) for _ in range(2)]
        res = sd.detect_shard(dev, core_lo, core_hi, cand_lo, opt, outs[0])
        ref = sections(sd.gather_heads(res, want, cap=cap))                      # NCCL all-gather
        res = sd.detect_shard(dev, core_lo, core_hi, cand_lo, opt, outs[0], head_want=want, head_cap=cap, slot=0, peer=True)
        got = sections(sd.gather_heads(res, want, cap=cap, slot=0, peer=True))   # peer stores + barrier
        bad += 0 if got == ref else 1
        # queued steps: two in flight, then completion established by device synchronisation + barrier
        pend = []
        last = None
        for k in range(5):
            pend.append(sd.detect_shard_async(dev, core_lo, core_hi, cand_lo, opt, outs[k & 1], head_want=want, head_cap=cap,
                                              slot=k & 1, peer=True))
            if len(pend) > 1:
                last = sd.finish_shard(pend.pop(0))
                if last.head_slot is None:
                    sd.gather_heads(last, want, cap=cap, slot=(k - 1) & 1, async_op=True, peer=True)
        last = sd.finish_shard(pend.pop(0))
        if last.head_slot is None:
            sd.gather_heads(last, want, cap=cap, slot=0, async_op=True, peer=True)
        torch.cuda.synchronize()
        dist.barrier()
        ep = x.epochs(0)
        bad += 0 if np.all(ep == last.head_epoch) else 1
        bad += 0 if sections(x.gathered(0).clone()) == ref else 1
        dist.barrier()
    flag = torch.tensor([bad], device="cuda")
    dist.all_reduce(flag)
    ok = int(flag.item()) == 0
    if rank == 0:
        print("multi_gpu_check peer world=%d peer_exchange_identical_to_nccl_all_gather=%s" % (world, ok))
    return ok


def chr20(det, rank, world, steps=10):
    from bench import make_device_workload
    device = torch.device("cuda", torch.cuda.current_device())
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
    L, n = CHR20, 30
    whole, _ = make_device_workload(L, n, n, device)  # same seed on every rank: the same genome
    out_w = nm.alloc_device_table(opt, L, device)
    assert det.detect_device(whole, opt, out_w) == L
    halo = shard_halo(opt)
    lo, hi = rank * L // world, (rank + 1) * L // world
    hlo, hhi = max(0, lo - halo), min(L, hi + halo)
    from nanomod_b200._lib import padded_len
    nv = padded_len((hhi - hlo) * n)
    sl = nm.DevicePileup(whole.vals0[hlo * n:hlo * n + nv].clone(),
                         (whole.off0[hlo:hhi + 1] - whole.off0[hlo]).contiguous(),
                         whole.vals1[hlo * n:hlo * n + nv].clone(),
                         (whole.off1[hlo:hhi + 1] - whole.off1[hlo]).contiguous(),
                         whole.pos[hlo:hhi].contiguous(), whole.seg[hlo:hhi].contiguous(), hhi - hlo)
    sd = ShardedDetector(det)
    out_s = nm.alloc_device_table(opt, sl.n_pos, device)
    res = sd.detect_shard(sl, lo - hlo, hi - hlo, hlo, opt, out_s)
    bad = 0
    for c in ("n0", "n1", "ks_dnum", "ks_d", "ks_p", "stouffer_stat", "stouffer_p"):
        a, b = res.core(c), out_w[c][lo:hi]
        bad += int((a.view(torch.uint8) != b.view(torch.uint8)).any().item()) if a.dtype != torch.uint8 else int((a != b).any().item())
    sites = sd.called_sites(res, [("syn", "+")])
    from nanomod_b200.sharded import ShardResult, greedy_sites, merge_heads
    res_w = ShardResult(out_w, whole, L, 0, L, 0, opt)
    sites_1, _ = greedy_sites(merge_heads([sd.local_head(res_w, 4096)], False), opt, [("syn", "+")])
    bad += 0 if sites == sites_1 else 1
    # timing: the sharded step (detect on the shard + heads all-gathered), strong scaling
    for _ in range(3):
        res = sd.detect_shard(sl, lo - hlo, hi - hlo, hlo, opt, out_s)
        sd.gather_heads(res, 1024)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lane = 0.0
    for _ in range(steps):
        res = sd.detect_shard(sl, lo - hlo, hi - hlo, hlo, opt, out_s)
        lane += det.handle.last_timings()["lane"] / steps
        sd.gather_heads(res, 1024)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps, lane, float(bad)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = t[2].item() == 0
    if rank == 0:
        print(json.dumps({"check": "chr20_strong_scaling", "n_gpus": world, "positions": L, "coverage": [n, n],
                          "ms_per_step_max_over_ranks": t[0].item(), "lane_kernel_ms_max_over_ranks": t[1].item(),
                          "positions_per_s": L / (t[0].item() * 1e-3), "shard_rows_and_called_sites_identical_to_single_gpu": ok,
                          "halo_candidates": halo, "called_sites_head": [list(s) for s in sites[:5]],
                          "step": "detect_shard (plan + lane + combine on the shard) + nm_rank_head_device + NCCL all-gather of the heads"}))
    return ok


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    det = nm.Detector(lr)
    ok = small(det, rank, world)
    ok &= peer_exchange(det, rank, world)
    if "--chr20" in sys.argv:
        ok &= chr20(det, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
