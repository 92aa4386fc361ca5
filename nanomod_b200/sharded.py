"""Genome sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

The per-position tests are independent per position, and the window combination of a row only
looks at rows within ``neighborPvalues`` positions of it inside the same contiguous run
(pos_check, bin/scripts/myDetect.py:366-371, :383-389).  So the candidate list is cut into
contiguous ranges balanced by value count, every rank additionally computes a halo of ``nb``
candidates on each side (recomputed, never exchanged) and drops the halo rows afterwards.
A window slot that passes pos_check at row distance k is exactly k candidates away, hence a halo
of nb *candidates* is enough and the result is identical to the single-GPU one.  The only
communication is the final gather of the per-row result records to rank 0.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .detect import DetectOptions, SignTestTable
from .pileup import Pileup


def plan_shards(off0: np.ndarray, off1: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous candidate ranges [lo, hi), one per rank, balanced by the number of values
    (i.e. bytes streamed from HBM).  Ranges may be empty when world > n_pos."""
    n = off0.shape[0] - 1
    work = (off0 + off1).astype(np.int64)  # prefix sum of per-candidate value counts
    total = int(work[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        c = int(np.searchsorted(work, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_with_halo(pileup: Pileup, lo: int, hi: int, nb: int) -> Tuple[Pileup, int, int]:
    """Candidates [lo-nb, hi+nb) clipped to the pileup, plus the core range inside the slice."""
    hlo = max(0, lo - nb)
    hhi = min(pileup.n_pos, hi + nb)
    return pileup.slice_rows(hlo, hhi), lo - hlo, hi - hlo


def trim_table(t: SignTestTable, core_lo: int, core_hi: int, index_offset: int) -> SignTestTable:
    """Keep rows whose candidate index lies in the core range; re-base row_pos_index to the
    full pileup (index_offset = global index of the slice's first candidate)."""
    keep = (t.row_pos_index >= core_lo) & (t.row_pos_index < core_hi)
    kw = {}
    for name in _lib.TABLE_FIELDS:
        col = getattr(t, name)
        kw[name] = None if col is None else col[keep]
    kw["row_pos_index"] = kw["row_pos_index"] + np.int32(index_offset)
    return SignTestTable(options=t.options, seg_names=t.seg_names, seg=t.seg[keep], pos=t.pos[keep],
                         base=t.base[keep], **kw)


_META = ["seg", "pos", "base"]


def _columns(t: SignTestTable) -> List[str]:
    return [c for c in _lib.TABLE_FIELDS if getattr(t, c) is not None]


def pack_records(t: SignTestTable) -> np.ndarray:
    """Fixed-width byte records [rows, width] (all present columns + seg/pos/base)."""
    parts = [np.ascontiguousarray(getattr(t, c)).view(np.uint8).reshape(len(t), -1) for c in _columns(t)]
    parts += [np.ascontiguousarray(getattr(t, m)).view(np.uint8).reshape(len(t), -1) for m in _META]
    return np.ascontiguousarray(np.concatenate(parts, axis=1)) if len(t) else \
        np.zeros((0, sum(p.shape[1] for p in parts)), np.uint8)


def unpack_records(rec: np.ndarray, like: SignTestTable) -> SignTestTable:
    cols = _columns(like)
    kw = {}
    o = 0
    for c in cols:
        dt = np.dtype(_lib.TABLE_DTYPES[c])
        wd = _lib.TABLE_WIDTH.get(c, 1)
        kw[c] = np.ascontiguousarray(rec[:, o:o + wd * dt.itemsize]).view(dt).reshape((-1, wd) if wd > 1 else -1)
        o += wd * dt.itemsize
    meta = {}
    for m in _META:
        dt = getattr(like, m).dtype
        meta[m] = np.ascontiguousarray(rec[:, o:o + dt.itemsize]).view(dt).reshape(-1)
        o += dt.itemsize
    return SignTestTable(options=like.options, seg_names=like.seg_names, **meta, **kw)


def gather_tables(local: SignTestTable, group=None, device=None) -> Optional[SignTestTable]:
    """Gather every rank's trimmed table to rank 0 (concatenated in rank order == genome order).
    Uses the group's backend: NCCL moves device tensors over NVLink, gloo moves host tensors."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    backend = dist.get_backend(group)
    dev = torch.device("cpu") if backend == "gloo" else (device or torch.device("cuda", torch.cuda.current_device()))
    rec = torch.from_numpy(pack_records(local)).to(dev)
    width = rec.shape[1]
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = rec.shape[0]
    dist.all_reduce(counts, group=group)
    cmax = int(counts.max().item())
    padded = torch.zeros((cmax, width), dtype=torch.uint8, device=dev)
    padded[:rec.shape[0]] = rec
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    parts = [bufs[r][:int(counts[r].item())].cpu().numpy() for r in range(world)]
    return unpack_records(np.concatenate(parts, axis=0), local)


class ShardedDetector:
    """Runs the detection stage on this rank's shard and gathers the table on rank 0.

    ``engine`` is anything with ``detect(pileup, options) -> SignTestTable`` -- in the product
    always a ``nanomod_b200.Detector`` (CUDA); the CPU test-suite injects a checker engine to
    exercise the sharding / halo / gather logic under gloo."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group

    def detect(self, pileup: Pileup, options: DetectOptions) -> Optional[SignTestTable]:
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        options.validate()
        lo, hi = plan_shards(pileup.off0, pileup.off1, world)[rank]
        local = self.detect_range(pileup, lo, hi, options)
        if world == 1:
            return local
        return gather_tables(local, self.group)

    def detect_range(self, pileup: Pileup, lo: int, hi: int, options: DetectOptions) -> SignTestTable:
        nb = options.neighborPvalues if options.combine_mask() else 0
        sl, core_lo, core_hi = shard_with_halo(pileup, lo, hi, nb)
        t = self.engine.detect(sl, options)
        return trim_table(t, core_lo, core_hi, lo - core_lo)
