"""Genome sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

The per-position tests are independent per position, and the window combination of a row only
looks at rows within ``neighborPvalues`` positions of it inside the same contiguous run
(pos_check, bin/scripts/myDetect.py:366-371, :383-389).  So the candidate list is cut into
contiguous ranges balanced by value count, every rank additionally computes a halo of ``nb``
candidates on each side (recomputed, never exchanged) and drops the halo rows afterwards.
A window slot that passes pos_check at row distance k is exactly k candidates away, hence a halo
of nb *candidates* is enough and the result is identical to the single-GPU one.

What leaves a GPU.  The table stays sharded: every rank keeps (and, for ``save_test``, formats
and writes) its own rows.  The one thing that needs a global view is the called-site selection
(mboxplot / plot1, myDetect.py:279-297, :153-164), which walks the ranked list from the top and
stops after topN accepted sites -- so each rank contributes only the HEAD of its own ranking
(``nm_rank_head_device``: a few thousand rows found without a full sort), the heads are
all-gathered (NCCL over NVLink; gloo in the CPU tests) and merged on every rank.  The merged list
is exact up to the smallest "last key" any truncated head reports; the selection asks for longer
heads in the rare case it runs past that point.  ``gather_tables`` (everything to rank 0) remains
for callers that want the complete table in one place.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .detect import DetectOptions, SignTestTable
from .pileup import Pileup

HEAD_REC = np.dtype([("row", "<i8"), ("seg", "<i4"), ("pos", "<i4"), ("full_nbhd", "<i4"), ("reserved", "<i4"),
                     ("key", "<u8", (3,))])  # == nm_head_row (48 bytes)
HEAD_CAP = 16384   # records per rank in the exchange buffer (a longer head is cut: still a valid, incomplete head)


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


def bind_to_gpu_cpus(device_index: int) -> Optional[List[int]]:
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off (NVML's ideal CPU
    affinity).  One process per GPU on a multi-socket host: pinned staging buffers allocated afterwards
    are first-touched on that node, so the host<->device copies of the ranks do not all cross one
    socket's memory controller.  Returns the CPU list, or None when NVML / the platform does not
    offer it (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class _DevView:
    """a device address as a torch uint8 tensor (zero copy, through __cuda_array_interface__)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class HeadExchange:
    """Gathered-heads buffers in peer-visible HBM: every rank owns one buffer of two slots x world
    sections of (cap + 1) records, and maps every peer's buffer (nm_peer_alloc / nm_peer_open, CUDA
    IPC between the one-process-per-GPU ranks of a node).  A rank's head selection stores its
    records straight into its section of EVERY rank's buffer (nm_head_set_peers), so the exchange
    needs no collective kernel.  Building one is collective over the group."""

    def __init__(self, handle, group, cap: int, device):
        import torch
        import torch.distributed as dist
        self.handle, self.group, self.cap, self.device = handle, group, cap, device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.section = (cap + 1) * HEAD_REC.itemsize
        self.slot_bytes = self.world * self.section
        self.ptr, ipc = handle.peer_alloc(2 * self.slot_bytes)
        self.peer_ptrs = [self.ptr] * self.world
        self.ok = True
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, ipc, group=group)
            opened = []
            try:
                for r in range(self.world):
                    if r != self.rank:
                        self.peer_ptrs[r] = handle.peer_open(handles[r])
                        opened.append(self.peer_ptrs[r])
            except Exception:  # no peer mapping on this node: every rank falls back to the NCCL all-gather
                self.ok = False
            flag = torch.tensor([1 if self.ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            self.ok = bool(int(flag.item()))
            self._opened = opened
        self._local = torch.as_tensor(_DevView(self.ptr, 2 * self.slot_bytes), device=device)

    def close(self) -> None:
        """Unmaps the peers' buffers, meets the other ranks (nobody may still store into a buffer that is
        about to go), frees this rank's buffer.  Collective, like the constructor."""
        import torch
        import torch.distributed as dist
        if self.ptr is None:
            return
        torch.cuda.synchronize(self.device)
        for p in getattr(self, "_opened", []):
            try:
                self.handle.peer_close(p)
            except Exception:
                pass
        if self.world > 1:
            dist.barrier(group=self.group)
        self._local = None
        self.handle.peer_free(self.ptr)
        self.ptr, self.ok = None, False

    def bases(self, slot: int) -> List[int]:
        """this rank's section of slot ``slot`` in every rank's buffer"""
        off = (slot & 1) * self.slot_bytes + self.rank * self.section
        return [p + off for p in self.peer_ptrs]

    def gathered(self, slot: int):
        """this rank's copy of all sections of a slot (a view: valid once every rank's kernels are done)"""
        lo = (slot & 1) * self.slot_bytes
        return self._local[lo:lo + self.slot_bytes]

    def epochs(self, slot: int) -> np.ndarray:
        """the header epoch of every section of a slot, as it is on the device now"""
        hdr = self.gathered(slot).view(self.world, self.section)[:, :HEAD_REC.itemsize].contiguous().cpu().numpy()
        return hdr.view(HEAD_REC)["reserved"].reshape(-1).astype(np.int64)


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

    # ---- the table stays sharded (device-resident; only heads of the ranking are exchanged) ----
    def _world(self) -> Tuple[int, int]:
        import torch.distributed as dist
        if not dist.is_initialized():
            return 1, 0
        return dist.get_world_size(self.group), dist.get_rank(self.group)

    def shard_device(self, pileup: Pileup, options: DetectOptions, device):
        """This rank's slice of a host pileup (core range + halo of ``shard_halo`` candidates) on
        its GPU: (DevicePileup, core_lo, core_hi inside the slice, global index of its first candidate)."""
        from .detect import DevicePileup
        world, rank = self._world()
        lo, hi = plan_shards(pileup.off0, pileup.off1, world)[rank]
        sl, core_lo, core_hi = shard_with_halo(pileup, lo, hi, shard_halo(options))
        return DevicePileup.from_host(sl, device), core_lo, core_hi, lo - core_lo

    def _head_buffers(self, dev, cap: int, world: int):
        import torch
        key = (str(dev), cap, world)
        if getattr(self, "_head_key", None) != key:
            self._head_key = key
            self._head_mine = [torch.zeros((cap + 1) * HEAD_REC.itemsize, dtype=torch.uint8, device=dev) for _ in range(2)]
            self._head_all = [torch.zeros(world * (cap + 1) * HEAD_REC.itemsize, dtype=torch.uint8, device=dev) for _ in range(2)]

    def peer_exchange(self, cap: int, device):
        """The peer-memory head exchange of this group (built on first use -- collectively: every rank must
        make its first head-selecting call at the same point).  None when the engine has no device handle;
        ``.ok`` False when the node offers no peer mapping (the NCCL all-gather of ``gather_heads`` is used)."""
        handle = getattr(self.engine, "handle", None)
        if handle is None or not hasattr(handle, "peer_alloc"):
            return None
        x = getattr(self, "_xchg", None)
        if x is None or x.cap != cap:
            if x is not None:
                x.close()
            x = self._xchg = HeadExchange(handle, self.group, cap, device)
            self._epoch = 0
        return x

    def _arm_head(self, dev, out, core_lo: int, core_hi: int, options: DetectOptions, head_want: int, head_cap: int,
                  slot: int, peer: bool):
        """arms the head selection for the next detect call; returns the epoch of a peer-memory exchange (or 0)"""
        world, _ = self._world()
        self._head_buffers(dev.pos.device, head_cap, world)
        epoch = 0
        if peer:
            x = self.peer_exchange(head_cap, dev.pos.device)
            if x is not None and x.ok:
                self._epoch += 1
                epoch = self._epoch
                self.engine.handle.head_set_peers(x.bases(slot), epoch)
        self.engine.arm_head_select(out, core_lo, core_hi - core_lo, options, head_want, self._head_mine[slot & 1], head_cap,
                                    (None, dev.pos, dev.seg, core_lo, dev.n_pos, nearby_rows(options)))
        return epoch

    def _shard_result(self, dev, out, n_rows: int, core_lo: int, core_hi: int, cand_lo: int, options: DetectOptions,
                      fired: bool, slot: int, head_want: int, head_cap: int, epoch: int) -> "ShardResult":
        import torch
        if n_rows == dev.n_pos:  # nothing filtered: rows are the candidates
            r_lo, r_hi = core_lo, core_hi
        else:
            rpi = out["row_pos_index"][:n_rows]
            edges = torch.searchsorted(rpi, torch.tensor([core_lo, core_hi], dtype=rpi.dtype, device=rpi.device))
            r_lo, r_hi = int(edges[0].item()), int(edges[1].item())
        res = ShardResult(out, dev, n_rows, r_lo, r_hi, cand_lo, options)
        res.head_slot = (slot & 1, head_want, head_cap) if fired else None
        res.head_epoch = epoch  # > 0: a peer-memory exchange was armed under this epoch (stored iff head_slot)
        return res

    def detect_shard(self, dev, core_lo: int, core_hi: int, cand_lo: int, options: DetectOptions,
                     out: Optional[Dict[str, "object"]] = None, head_want: int = 0, head_cap: int = HEAD_CAP,
                     slot: int = 0, peer: bool = False) -> ShardResult:
        """Detection on a device-resident shard (``dev`` = core candidates [core_lo, core_hi) plus
        halo).  Results stay on the GPU; no communication.  With ``head_want`` the selection of this
        rank's ranking head (what ``gather_heads`` exchanges) is armed for the call: it is launched behind
        the call's own kernels and before its host wait -- when the coverage filter drops nothing, which is
        what the row range was armed for (``ShardResult.head_slot``; otherwise ``gather_heads`` selects).
        ``peer=True``: the armed selection also stores the head into every rank's gathered-heads buffer
        (``HeadExchange``), i.e. the exchange itself happens inside the selection kernels."""
        from .detect import alloc_device_table
        if out is None:
            out = alloc_device_table(options, dev.n_pos, dev.off0.device)
        epoch = 0
        if head_want > 0 and core_hi > core_lo:
            epoch = self._arm_head(dev, out, core_lo, core_hi, options, head_want, head_cap, slot, peer)
        n_rows = self.engine.detect_device(dev, options, out)
        fired = head_want > 0 and self.engine.handle.head_fired()
        return self._shard_result(dev, out, n_rows, core_lo, core_hi, cand_lo, options, fired, slot, head_want, head_cap, epoch)

    def detect_shard_async(self, dev, core_lo: int, core_hi: int, cand_lo: int, options: DetectOptions,
                           out: Dict[str, "object"], head_want: int = 0, head_cap: int = HEAD_CAP, slot: int = 0,
                           peer: bool = True):
        """``detect_shard`` without the host wait (nm_detect_device_async): the step is queued and a pending
        record returned for ``finish_shard``.  Up to two steps may be in flight, on different ``out`` tables and
        ``slot`` values; a caller that keeps one step queued ahead never leaves the GPU idle between steps."""
        epoch = 0
        if head_want > 0 and core_hi > core_lo:
            epoch = self._arm_head(dev, out, core_lo, core_hi, options, head_want, head_cap, slot, peer)
        ticket = self.engine.detect_device_async(dev, options, out)
        return (ticket, dev, out, core_lo, core_hi, cand_lo, options, slot, head_want, head_cap, epoch)

    def finish_shard(self, pending) -> ShardResult:
        """Completes a step queued by ``detect_shard_async`` (waits for that step only)."""
        ticket, dev, out, core_lo, core_hi, cand_lo, options, slot, head_want, head_cap, epoch = pending
        n_rows, fired = self.engine.detect_finish(ticket)
        return self._shard_result(dev, out, n_rows, core_lo, core_hi, cand_lo, options, fired and head_want > 0, slot,
                                  head_want, head_cap, epoch)

    def local_head(self, res: ShardResult, want: int) -> LocalHead:
        """Leading rows of this rank's own ranking of its core rows (nm_rank_head_device: three
        streaming passes + a few thousand records to the host)."""
        o = res.options
        use_p = o.rankUse == "pv"
        core = {c: v[res.r_lo:res.r_hi] for c, v in res.out.items() if v.dim() == 1}
        rpi = None if res.n_rows == res.dev.n_pos else res.out["row_pos_index"]
        rows = self.engine.rank_head_device(core, res.n_core, o, want,
                                            geometry=(rpi, res.dev.pos, res.dev.seg, res.r_lo, res.n_rows, nearby_rows(o)))
        return LocalHead(rows["row"].astype(np.int64), rows["seg"].astype(np.int32), rows["pos"].astype(np.int32),
                         np.ascontiguousarray(rows["key"]), rows["full_nbhd"] != 0, res.n_core, len(rows) == res.n_core)

    def gather_heads(self, res: ShardResult, want: int, cap: int = HEAD_CAP, slot: int = 0, async_op: bool = False,
                     peer: bool = False):
        """The multi-GPU exchange of a step, all on the device: three streaming kernels select the head of
        this rank's ranking into a record buffer (nm_rank_head_select_device) -- unless the detect call has
        done so already (``detect_shard(head_want=...)``) -- and every rank gets all heads.

        ``peer=False``: ONE NCCL all-gather.  Returns the gathered records as a CUDA tensor;
        ``heads_from_gathered`` parses it.  ``async_op=True`` returns (tensor, work): the collective runs
        beside whatever the caller launches next (``slot`` selects one of two buffer pairs);
        ``work.wait()`` before the tensor is read.

        ``peer=True`` (falls back to the above when the node has no peer mapping): the selection kernels
        store the head into every rank's buffer themselves (``HeadExchange``), no collective kernel runs.
        The call then waits for its own kernels, meets the other ranks at a barrier -- after which every
        head has arrived -- and returns a copy of the gathered records; with ``async_op=True`` it returns
        (view, None) at once and establishing completion (device synchronisation + barrier) is the caller's."""
        import torch
        import torch.distributed as dist
        o = res.options
        world, _ = self._world()
        dev = res.dev.pos.device
        self._head_buffers(dev, cap, world)
        x = self.peer_exchange(cap, dev) if peer else None
        use_peer = x is not None and x.ok
        mine, allv = self._head_mine[slot & 1], self._head_all[slot & 1]
        selected = getattr(res, "head_slot", None) == (slot & 1, want, cap)
        epoch = getattr(res, "head_epoch", 0)
        if use_peer and not (selected and epoch > 0):
            selected = False
            if epoch <= 0:  # every rank draws the same number: calls are collective
                self._epoch += 1
                epoch = res.head_epoch = self._epoch
            self.engine.handle.head_set_peers(x.bases(slot), epoch)
        if not selected:  # not selected by the detect call itself
            core = {c: res.out[c][res.r_lo:res.r_hi] for c in ("ks_p", "ks_d", "u_p", "u_stat", "fisher_p", "fisher_stat",
                                                               "stouffer_p", "stouffer_stat") if c in res.out}
            rpi = None if res.n_rows == res.dev.n_pos else res.out["row_pos_index"]
            self.engine.rank_head_select_device(core, res.n_core, o, want, mine, cap,
                                                geometry=(rpi, res.dev.pos, res.dev.seg, res.r_lo, res.n_rows, nearby_rows(o)))
            if use_peer:
                res.head_slot = (slot & 1, want, cap)
        if use_peer:
            if async_op:
                return x.gathered(slot), None
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier(group=self.group)
            got = x.gathered(slot).clone()
            ep = x.epochs(slot)
            if not np.all(ep == epoch):
                raise RuntimeError("head exchange: sections carry epochs %s, expected %d" % (ep.tolist(), epoch))
            return got
        if world == 1:
            return (mine, None) if async_op else mine
        work = dist.all_gather_into_tensor(allv, mine, group=self.group, async_op=async_op)
        return (allv, work) if async_op else allv

    def heads_from_gathered(self, gathered, options: DetectOptions, cap: int = HEAD_CAP) -> List[LocalHead]:
        """gathered device records -> every rank's head, each in its own ranking order"""
        buf = gathered.cpu().numpy().view(HEAD_REC)
        world = buf.shape[0] // (cap + 1)
        heads = []
        for h in unpack_heads(buf, world, cap):
            order = np.lexsort((h.rows if options.rankUse == "pv" else -h.rows, h.keys[:, 2], h.keys[:, 1], h.keys[:, 0])) \
                if h.rows.shape[0] else np.zeros(0, np.int64)
            heads.append(LocalHead(h.rows[order], h.seg[order], h.pos[order], h.keys[order], h.full_nbhd[order],
                                   h.n_core, h.complete))
        return heads

    def merged_head(self, res: ShardResult, want: int) -> MergedHead:
        heads = self.heads_from_gathered(self.gather_heads(res, want), res.options)
        return merge_heads(heads, res.options.rankUse != "pv")

    def _called_sites(self, head_fn, options: DetectOptions, seg_names, device=None, want: Optional[int] = None):
        if options.RegionRankbyST != 0:
            raise NotImplementedError("region ranking needs the whole table: use gather_tables / SignTestTable")
        want = want or max(64 * options.topN, 1024)
        while True:
            m = merge_heads(exchange_heads(head_fn(want), self.group, device), options.rankUse != "pv")
            sites, final = greedy_sites(m, options, seg_names)
            if final or m.complete:
                return sites
            want *= 8  # the walk ran past what the heads guarantee: ask every rank for more (same decision on all ranks)

    def called_sites(self, res: ShardResult, seg_names, want: Optional[int] = None) -> List[Tuple[str, str, int]]:
        """The reference's called-site list (same on every rank) from the sharded, device-resident table."""
        o = res.options
        if o.RegionRankbyST != 0:
            raise NotImplementedError("region ranking needs the whole table: use gather_tables / SignTestTable")
        want = want or max(64 * o.topN, 1024)
        cap = HEAD_CAP
        while True:
            m = merge_heads(self.heads_from_gathered(self.gather_heads(res, want, cap), o, cap), o.rankUse != "pv")
            sites, final = greedy_sites(m, o, seg_names)
            if final or m.complete:
                return sites
            want *= 8  # the walk ran past what the heads guarantee: longer heads (same decision on every rank)
            cap = max(cap, 2 * want)

    def called_sites_host(self, t: SignTestTable, core_lo: int, core_hi: int, want: Optional[int] = None):
        """Same from a host table of this rank's rows (core rows [core_lo, core_hi) + halo rows)."""
        return self._called_sites(lambda w: local_head_from_table(t, core_lo, core_hi, w), t.options, t.seg_names,
                                  None, want)

    def save_test(self, res: ShardResult, seg_names, base: np.ndarray, path: str, n_threads: int = 0) -> int:
        """`save_test` (myDetect.py:522-538) without gathering: every rank formats its own rows
        (native formatter) and writes them at its byte offset of ONE file, identical to the
        single-GPU table.  ``base``: uint8 base character of the shard's candidates.  Returns the
        bytes this rank wrote."""
        import torch
        import torch.distributed as dist
        world, rank = self._world()
        o = res.options
        cols = {c: res.core(c).cpu().numpy() for c in res.out if res.out[c].dim() == 1}
        cand = np.arange(res.r_lo, res.r_hi) if res.n_rows == res.dev.n_pos else cols["row_pos_index"]
        pos = res.dev.pos.cpu().numpy()[cand]
        seg = res.dev.seg.cpu().numpy()[cand]
        t = SignTestTable(options=o, seg_names=list(seg_names), seg=seg, pos=pos, base=np.asarray(base, np.uint8)[cand],
                          **{c: cols.get(c) for c in _lib.TABLE_FIELDS if c != "moments"})
        text = t.format_text(n_threads)
        sizes = [len(text)]
        if world > 1:
            backend = dist.get_backend(self.group)
            dev = torch.device("cpu") if backend == "gloo" else res.dev.pos.device
            mine = torch.tensor([len(text)], dtype=torch.int64, device=dev)
            allv = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine, group=self.group)
            sizes = [int(v.item()) for v in allv]
            if rank == 0:
                with open(path, "wb") as fh:
                    fh.truncate(sum(sizes))
            dist.barrier(group=self.group)
        else:
            with open(path, "wb") as fh:
                fh.truncate(sizes[0])
        with open(path, "r+b") as fh:
            fh.seek(sum(sizes[:rank]))
            fh.write(text)
        if world > 1:
            dist.barrier(group=self.group)
        return len(text)


# ---------------------------------------------------------------------------------------------
# the sharded table: heads of the ranking, merged; called sites; per-rank text output
# ---------------------------------------------------------------------------------------------
def nearby_rows(options: DetectOptions) -> int:
    """rows either side of a site that plot1 requires to be contiguous (myDetect.py:153-154)"""
    w = options.half_window + (1 if options.RegionRankbyST != 0 else 0)
    return 2 * w if options.RegionRankbyST == 1 else w


def shard_halo(options: DetectOptions) -> int:
    """candidates a shard needs beyond its core range: the combination window and the
    neighbourhood test of the called-site rule both look that far"""
    nb = options.neighborPvalues if options.combine_mask() else 0
    return max(nb, nearby_rows(options))


def key_image(x: np.ndarray) -> np.ndarray:
    """order-preserving uint64 image of float64 keys: NaN last, -0.0 == 0.0 (nm_rank_key, nm_rank.cu)"""
    x = np.asarray(x, np.float64) + 0.0
    b = x.view(np.uint64)
    img = np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))
    return np.where(np.isnan(x), np.uint64(0xFFFFFFFFFFFFFFFF), img)


@dataclass
class LocalHead:
    """The leading rows of one rank's own ranking, in ranking order."""
    rows: np.ndarray       # int64: row index inside the rank's core rows
    seg: np.ndarray        # int32
    pos: np.ndarray        # int32
    keys: np.ndarray       # uint64 [K, 3]: sort images of (combined, KS, U) -- ascending unsigned order is the ranking
                           # (already complemented for rankUse='st'); 0 where a key is absent
    full_nbhd: np.ndarray  # bool: rows r-nearby .. r+nearby are one contiguous run (plot1's requirement)
    n_core: int            # core rows this rank holds in total
    complete: bool         # the head holds every core row


def local_head_from_table(t: SignTestTable, core_lo: int, core_hi: int, want: int) -> LocalHead:
    """Head of a HOST table whose rows [core_lo, core_hi) are the rank's core rows (the rest is
    halo); numpy ranking.  The CUDA path builds the same thing with nm_rank_head_device."""
    o = t.options
    use_p = o.rankUse == "pv"
    comb = t.comb()
    cols = [None if comb is None else (comb[1] if use_p else comb[0]), t.ks_p if use_p else t.ks_d,
            None if t.u_p is None else (t.u_p if use_p else t.u_stat)]
    n = core_hi - core_lo
    img = []
    for c in cols:
        k = np.zeros(n, np.uint64) if c is None else key_image(c[core_lo:core_hi])
        img.append(k if (use_p or c is None) else ~k)
    order = np.lexsort((np.arange(n) if use_p else -np.arange(n), img[2], img[1], img[0]))
    k = min(n, max(want, 0))
    # keep whole groups of equal primary keys, as the device selection does (it keeps whole exponent bins)
    while 0 < k < n and img[0][order[k]] == img[0][order[k - 1]]:
        k += 1
    rows = order[:k].astype(np.int64)
    keys = np.stack([i[rows] for i in img], axis=1) if k else np.zeros((0, 3), np.uint64)
    return LocalHead(rows, t.seg[core_lo:core_hi][rows].astype(np.int32), t.pos[core_lo:core_hi][rows].astype(np.int32),
                     keys, neighbourhood_flags(t.seg, t.pos, rows + core_lo, nearby_rows(o)), n, k == n)


def neighbourhood_flags(seg: np.ndarray, pos: np.ndarray, rows: np.ndarray, nearby: int) -> np.ndarray:
    """plot1's test for each row r of a (local, halo included) row list: rows r-nearby..r+nearby
    exist and form one contiguous run.  Positions increase strictly inside a segment, so the run is
    contiguous iff its two ends are 2*nearby positions apart on the same segment."""
    n = seg.shape[0]
    lo, hi = rows - nearby, rows + nearby
    okk = (lo >= 0) & (hi <= n - 1)
    lo_c, hi_c = np.clip(lo, 0, max(n - 1, 0)), np.clip(hi, 0, max(n - 1, 0))
    if n == 0:
        return np.zeros(rows.shape[0], bool)
    return okk & (seg[lo_c] == seg[hi_c]) & (pos[hi_c].astype(np.int64) - pos[lo_c].astype(np.int64) == 2 * nearby)




def pack_head(local: LocalHead, cap: int = HEAD_CAP) -> np.ndarray:
    """[cap + 1] records: a header record followed by the head (cut at `cap`)"""
    k = min(local.rows.shape[0], cap)
    buf = np.zeros(cap + 1, dtype=HEAD_REC)
    buf["row"][0] = k
    buf["key"][0, 0] = local.n_core
    buf["key"][0, 1] = 1 if (local.complete and k == local.rows.shape[0]) else 0
    buf["row"][1:k + 1], buf["seg"][1:k + 1], buf["pos"][1:k + 1] = local.rows[:k], local.seg[:k], local.pos[:k]
    buf["full_nbhd"][1:k + 1] = local.full_nbhd[:k]
    buf["key"][1:k + 1] = local.keys[:k]
    return buf


def unpack_heads(buf: np.ndarray, world: int, cap: int = HEAD_CAP) -> List[LocalHead]:
    out = []
    for r in range(world):
        b = buf[r * (cap + 1):(r + 1) * (cap + 1)]
        k = int(b["row"][0])
        out.append(LocalHead(b["row"][1:k + 1].astype(np.int64), b["seg"][1:k + 1].copy(), b["pos"][1:k + 1].copy(),
                             b["key"][1:k + 1].copy(), b["full_nbhd"][1:k + 1] != 0, int(b["key"][0, 0]), bool(b["key"][0, 1])))
    return out


def exchange_heads(local: LocalHead, group=None, device=None, cap: int = HEAD_CAP, parse: bool = True):
    """all-gather of every rank's head: ONE collective of fixed-size records (NCCL over NVLink on
    GPUs, gloo on the CPU).  Returns the list of heads (or the raw gathered record array with
    ``parse=False``: the exchange is then complete and the host-side merge can happen later)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local] if parse else pack_head(local, cap)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cpu") if backend == "gloo" else (device or torch.device("cuda", torch.cuda.current_device()))
    mine = torch.from_numpy(pack_head(local, cap).view(np.uint8)).to(dev, non_blocking=True)
    gathered = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    buf = gathered.cpu().numpy().view(HEAD_REC)
    return unpack_heads(buf, world, cap) if parse else buf


@dataclass
class MergedHead:
    rank: np.ndarray       # owner of each row
    row: np.ndarray        # GLOBAL row index (rows of all ranks concatenated in rank order)
    seg: np.ndarray
    pos: np.ndarray
    keys: np.ndarray
    full_nbhd: np.ndarray
    n_exact: int           # the first n_exact rows are exactly the first rows of the global ranking
    complete: bool         # every row of every rank is in


def merge_heads(heads: Sequence[LocalHead], reverse: bool) -> MergedHead:
    """Global ranking of the gathered heads.  Stable order = global row order (ranks hold
    consecutive genome ranges), reversed as a whole for rankUse='st'.  Rows of a truncated head
    rank before everything that head left out, so the merged list is exact up to the smallest of
    the truncated heads' last rows."""
    base = np.concatenate([[0], np.cumsum([h.n_core for h in heads])])
    rank = np.concatenate([np.full(h.rows.shape[0], r, np.int32) for r, h in enumerate(heads)])
    row = np.concatenate([h.rows + base[r] for r, h in enumerate(heads)]).astype(np.int64)
    keys = np.concatenate([h.keys.reshape(-1, 3) for h in heads], axis=0).astype(np.uint64)
    order = np.lexsort((-row if reverse else row, keys[:, 2], keys[:, 1], keys[:, 0]))
    n_exact = order.shape[0]
    pos_in_order = np.empty(order.shape[0], np.int64)
    pos_in_order[order] = np.arange(order.shape[0])
    start = 0
    for h in heads:
        k = h.rows.shape[0]
        if not h.complete:
            # this rank's unseen rows all rank after its last head row: nothing past it is certain
            n_exact = min(n_exact, int(pos_in_order[start + k - 1]) + 1 if k else 0)
        start += k
    seg = np.concatenate([h.seg for h in heads])[order]
    pos = np.concatenate([h.pos for h in heads])[order]
    fl = np.concatenate([h.full_nbhd for h in heads])[order]
    return MergedHead(rank[order], row[order], seg, pos, keys[order], fl, n_exact, all(h.complete for h in heads))


def greedy_sites(m: MergedHead, options: DetectOptions, seg_names) -> Tuple[List[Tuple[str, str, int]], bool]:
    """mboxplot's walk (myDetect.py:279-297) over the exact part of a merged head.  Returns the
    sites and whether the answer is final (topN reached, or every row was available)."""
    closesize = options.neighborPvalues * 2
    if options.RegionRankbyST == 1:
        closesize = max(1, options.half_window + 1)
    out: List[Tuple[str, str, int]] = []
    acc: List[Tuple[int, int]] = []
    for i in range(m.n_exact):
        sg, ps = int(m.seg[i]), int(m.pos[i])
        if any(s == sg and abs(p - ps) < closesize for s, p in acc):
            continue
        if m.full_nbhd[i]:
            sk = seg_names[sg]
            out.append((sk[0], sk[1], ps))
            acc.append((sg, ps))
            if len(out) == options.topN:
                return out, True
    return out, m.complete and m.n_exact == m.row.shape[0]


@dataclass
class ShardResult:
    """One rank's part of the table, resident on its GPU (columns of ``out`` have capacity for the
    shard's candidates; rows [r_lo, r_hi) are the core rows, the others belong to the halo)."""
    out: Dict[str, "object"]
    dev: "object"          # DevicePileup of the shard (core + halo candidates)
    n_rows: int
    r_lo: int
    r_hi: int
    cand_lo: int           # global candidate index of the shard's first candidate
    options: DetectOptions
    head_slot: Optional[tuple] = None  # (slot, want, cap) when the detect call itself selected the ranking head
    head_epoch: int = 0                # epoch of the peer-memory exchange armed with it (0: none)

    @property
    def n_core(self) -> int:
        return self.r_hi - self.r_lo

    def core(self, name: str):
        return self.out[name][self.r_lo:self.r_hi]
