"""Host pileup packer: per-read event tables -> CSR pileup.

Replaces the reference's hot loop #0, ``ReadAllFast5`` / ``mReadSignalBase``
(bin/scripts/myDetect.py:33-127, :547-633), which appends every event of every read to
``moptions[ds]['norm_mean'][(chrom,strand)][pos]`` one Python list element at a time.  Here a
read is a ``ReadRecord`` (what ``myFast5.ReadMapInfoInRef`` and ``myFast5.ReadNanoraw_events``
return, myFast5.py:94-97, :119-126) and a whole group of reads is packed at once with numpy:
positions are computed per event (strand reversal as at :109-111), the read filters of :73-104
are applied, and a stable sort by (segment, position) turns the events into the CSR arrays the
C ABI consumes -- stable, so the values of a position keep the order in which the reference
would have appended them, and the base recorded for a position is that of the last read that
covers it (the reference overwrites ``['base'][...][pos]`` on every append, :122).

FAST5 input needs ``h5py``, which this image does not have: ``read_fast5`` imports it lazily
and is the only function that does.  ``save_reads_npz`` / ``load_reads_npz`` are the
interchange format for read tables produced elsewhere.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .pileup import Pileup

SegKey = Tuple[str, str]

# HDF5 locations of the Annotate stage's output, assembled from the same constants as the
# reference: myCom.py:35-56 (analyses_base, the NanoMod override of rawGenomeCorrected_base at
# :48-51 -- "NanomoCorrected_000", NOT nanoraw's "RawGenomeCorrected_000" --
# rawBaseCalled_template_base, rawAlignment_base, raw_event_base) and myFast5.py:92,113.
ANALYSES_BASE = "Analyses"
FAST5_MO_CORRECTION = "NanomoCorrected_000"          # myCom.py:48, :51
RAW_BASECALLED_TEMPLATE_BASE = "BaseCalled_template"  # myCom.py:52
RAW_ALIGNMENT_BASE = "Alignment"                      # myCom.py:53
RAW_EVENT_BASE = "Events"                             # myCom.py:56
FAST5_EVENTS = "/%s/%s/%s/%s" % (ANALYSES_BASE, FAST5_MO_CORRECTION, RAW_BASECALLED_TEMPLATE_BASE, RAW_EVENT_BASE)
FAST5_ALIGNMENT = "/%s/%s/%s/%s" % (ANALYSES_BASE, FAST5_MO_CORRECTION, RAW_BASECALLED_TEMPLATE_BASE, RAW_ALIGNMENT_BASE)


@dataclass
class ReadRecord:
    """One annotated read: where it maps and its per-base event table."""
    chrom: str
    strand: str            # '+' or '-'
    start: int             # mapped_start (0-based, leftmost reference coordinate)
    norm_mean: np.ndarray  # Events['norm_mean'], one value per read base, in READ order
    base: np.ndarray       # Events['base'] as uint8 character codes, same length

    def __post_init__(self):
        self.norm_mean = np.asarray(self.norm_mean, dtype=np.float64)
        b = self.base
        if isinstance(b, (bytes, str)):
            b = np.frombuffer(b.encode() if isinstance(b, str) else b, dtype=np.uint8)
        self.base = np.asarray(b, dtype=np.uint8)
        if self.norm_mean.shape != self.base.shape:
            raise ValueError("norm_mean and base differ in length")


@dataclass
class ReadFilter:
    """The read-level options of ``detect`` (NanoMod.py:377-388) as mReadSignalBase applies them."""
    min_lr: int = 500
    min_lr_nb: int = 0
    Chr: Optional[str] = None        # moptions['Chr']   (:73)
    Pos: Optional[int] = None        # with Pos2: reads must overlap [Pos, Pos2] (:75-77)
    Pos2: Optional[int] = None
    start_pos: Optional[int] = None  # with end_pos: reads must span it, events outside are dropped
    end_pos: Optional[int] = None    # (:79-86, :112-114)

    @classmethod
    def from_moptions(cls, mo: Dict) -> "ReadFilter":
        f = cls(min_lr=int(mo.get("min_lr", 500)), min_lr_nb=int(mo.get("min_lr_nb", 0)), Chr=mo.get("Chr"))
        if "Pos2" in mo:
            f.Pos, f.Pos2 = mo.get("Pos"), mo["Pos2"]
        elif "Pos" in mo and "window" in mo:  # ReadAllFast5 :550-557
            f.start_pos = max(0, mo["Pos"] - mo["window"])
            f.end_pos = mo["Pos"] + mo["window"]
        if "start_pos" in mo and "end_pos" in mo:
            f.start_pos, f.end_pos = mo["start_pos"], mo["end_pos"]
        return f

    def accepts(self, r: ReadRecord) -> bool:
        """``tocon`` of mReadSignalBase (myDetect.py:73-104) without the checkN early stop."""
        n = len(r.norm_mean)
        if self.Chr is not None and self.Chr != r.chrom:
            return False
        if self.Pos2 is not None and (r.start > self.Pos2 or r.start + n < self.Pos):
            return False
        if self.start_pos is not None and self.end_pos is not None:
            if r.start > self.start_pos or r.start + n < self.end_pos:
                return False
        nb = self.min_lr_nb
        if nb < 1:
            return n >= self.min_lr
        if not (self.min_lr - nb < n < self.min_lr + nb):
            return False

        def near(x):  # :100 -- read ends must sit near 0, 8000 or 16000
            return x < nb or 8000 - nb < x < 8000 + nb or 16000 - nb < x < 16000 + nb
        return near(r.start) and near(r.start + n)


@dataclass
class GroupEvents:
    """All accepted events of one group, flattened (one entry per event, in append order)."""
    seg_key: List[SegKey]
    seg: np.ndarray    # int32 index into seg_key
    pos: np.ndarray    # int64 reference coordinate
    val: np.ndarray    # float64 norm_mean
    base: np.ndarray   # uint8


def flatten_reads(reads: Iterable[ReadRecord], flt: Optional[ReadFilter] = None) -> GroupEvents:
    flt = flt or ReadFilter()
    keys: Dict[SegKey, int] = {}
    segs, poss, vals, bases = [], [], [], []
    for r in reads:
        if not flt.accepts(r):
            continue
        n = len(r.norm_mean)
        idx = np.arange(n, dtype=np.int64)
        p = r.start + idx if r.strand == "+" else r.start + n - 1 - idx  # :109-111
        keep = slice(None)
        if flt.start_pos is not None and flt.end_pos is not None:
            keep = (p >= flt.start_pos) & (p <= flt.end_pos)  # :112-114
        sid = keys.setdefault((r.chrom, r.strand), len(keys))
        pk = p[keep]
        segs.append(np.full(pk.shape[0], sid, dtype=np.int32))
        poss.append(pk)
        vals.append(r.norm_mean[keep])
        bases.append(r.base[keep])

    def cat(xs, dt):
        return np.concatenate(xs) if xs else np.zeros(0, dt)
    names = [k for k, _ in sorted(keys.items(), key=lambda kv: kv[1])]
    return GroupEvents(names, cat(segs, np.int32), cat(poss, np.int64), cat(vals, np.float64), cat(bases, np.uint8))


def _group_csr(ev: GroupEvents, order_of: Dict[SegKey, int]):
    """Events -> (keys [n,2] sorted, counts, values in append order per key, last base per key)."""
    gseg = np.array([order_of[k] for k in ev.seg_key], dtype=np.int64)[ev.seg] if len(ev.seg) else np.zeros(0, np.int64)
    order = np.lexsort((ev.pos, gseg))  # stable: append order survives inside a key
    s, p = gseg[order], ev.pos[order]
    new = np.ones(len(s), dtype=bool)
    new[1:] = (s[1:] != s[:-1]) | (p[1:] != p[:-1])
    starts = np.flatnonzero(new)
    counts = np.diff(np.append(starts, len(s)))
    last = np.append(starts[1:], len(s)) - 1
    return s[starts], p[starts], counts, ev.val[order], ev.base[order][last]


def pack_events(ev0: GroupEvents, ev1: GroupEvents) -> Pileup:
    """Two groups' events -> the candidate pileup: positions present in BOTH groups, ordered by
    sorted (chrom, strand) then position (mtest2, myDetect.py:421-431); base from group 1 (:436)."""
    both = sorted(set(ev0.seg_key) & set(ev1.seg_key))
    universe = sorted(set(ev0.seg_key) | set(ev1.seg_key))
    order_of = {k: i for i, k in enumerate(universe)}
    s0, p0, c0, v0, _b0 = _group_csr(ev0, order_of)
    s1, p1, c1, v1, b1 = _group_csr(ev1, order_of)
    # intersection of the two sorted key lists
    span = int(max(p0.max(initial=0), p1.max(initial=0))) + 1
    k0, k1 = s0 * span + p0, s1 * span + p1
    common, i0, i1 = np.intersect1d(k0, k1, assume_unique=True, return_indices=True)
    e0 = np.append(0, np.cumsum(c0))
    e1 = np.append(0, np.cumsum(c1))

    def take(vals, ends, counts, idx):
        if len(idx) == 0:
            return np.zeros(0, np.float32), np.zeros(1, np.int64)
        off = np.append(0, np.cumsum(counts[idx])).astype(np.int64)
        src = np.repeat(ends[idx] - off[:-1], counts[idx]) + np.arange(off[-1])
        return vals[src].astype(np.float32), off
    vals0, off0 = take(v0, e0, c0, i0)
    vals1, off1 = take(v1, e1, c1, i1)
    seg_of = {order_of[k]: i for i, k in enumerate(both)}
    seg = np.array([seg_of[int(s)] for s in s0[i0]], dtype=np.int32)
    return Pileup.from_arrays(vals0, off0, vals1, off1, p0[i0].astype(np.int32), seg, b1[i1], both)


def pack_reads(reads0: Iterable[ReadRecord], reads1: Iterable[ReadRecord],
               flt: Optional[ReadFilter] = None) -> Pileup:
    """ReadAllFast5 + the candidate rule of mtest2 in one step."""
    return pack_events(flatten_reads(reads0, flt), flatten_reads(reads1, flt))


# ---------------------------------------------------------------------------------------------
# input formats
# ---------------------------------------------------------------------------------------------
def read_fast5(path: str) -> Optional[ReadRecord]:
    """One annotated FAST5 (mReadSignalBase, myDetect.py:33-71).  Needs h5py."""
    try:
        import h5py  # noqa: WPS433 -- optional dependency, absent in the build image
    except ImportError as e:
        raise ImportError("reading FAST5 files needs h5py; use load_reads_npz for packed read tables") from e
    if not os.path.isfile(path):
        return None
    try:
        f = h5py.File(path, "r")
    except OSError:
        return None
    with f:
        if FAST5_ALIGNMENT not in f or FAST5_EVENTS not in f:
            return None
        at = dict(f[FAST5_ALIGNMENT].attrs.items())
        ev = f[FAST5_EVENTS][()]

        def text(x):
            return x.decode() if isinstance(x, bytes) else str(x)
        return ReadRecord(text(at["mapped_chrom"]), text(at["mapped_strand"]), int(at["mapped_start"]),
                          ev["norm_mean"], np.frombuffer(b"".join(ev["base"]), dtype=np.uint8))


def walk_fast5(folder: str, suffix: str = ".fast5") -> Iterable[str]:
    """The traversal of ReadAllFast5 / readsubfolder (:579-633): breadth first, skipping 'mall'."""
    level = [folder]
    while level:
        nxt: List[str] = []
        for sub in level:
            for name in os.listdir(sub):
                full = sub + "/" + name
                if name.endswith(suffix):
                    yield full
                elif os.path.isdir(full) and name != "mall":
                    nxt.append(full)
        level = nxt


def save_reads_npz(path: str, reads: Sequence[ReadRecord]) -> None:
    lens = np.array([len(r.norm_mean) for r in reads], dtype=np.int64)
    np.savez_compressed(
        path, chrom=np.array([r.chrom for r in reads]), strand=np.array([r.strand for r in reads]),
        start=np.array([r.start for r in reads], dtype=np.int64), length=lens,
        norm_mean=np.concatenate([r.norm_mean for r in reads]) if len(reads) else np.zeros(0),
        base=np.concatenate([r.base for r in reads]) if len(reads) else np.zeros(0, np.uint8))


def load_reads_npz(path: str) -> List[ReadRecord]:
    z = np.load(path)
    ends = np.cumsum(z["length"])
    out = []
    for i in range(len(ends)):
        lo, hi = int(ends[i] - z["length"][i]), int(ends[i])
        out.append(ReadRecord(str(z["chrom"][i]), str(z["strand"][i]), int(z["start"][i]),
                              z["norm_mean"][lo:hi], z["base"][lo:hi]))
    return out
