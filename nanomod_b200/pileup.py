"""CSR pileup: the device-friendly form of the reference's per-position signal lists.

The reference keeps ``moptions[ds]['norm_mean'][(chrom,strand)][pos] -> list[float]`` and
``['base'][...][pos] -> 'A'/'C'/'G'/'T'`` (bin/scripts/myDetect.py:569-572, filled at :108-124).
Here the two groups become two flat float32 arrays plus per-position int64 offsets, over a common
list of *candidate* positions in the reference's iteration order -- sorted (chrom,strand) tuples,
ascending position (:421,429).  Candidates are positions present in both groups; the coverage
filter (mfilter_coverage, :301-314) runs on the GPU.

float32 note: the reference's values are float64 on a 0.001 grid with |x| < ~10;
casting them to float32 preserves order and ties exactly, so D and U are unaffected.  The parity
rule is "the oracle consumes the identical float32 values upcast to float64".
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import padded_len

SegKey = Tuple[str, str]  # (chrom, strand)


@dataclass
class Pileup:
    vals0: np.ndarray          # float32, padded to padded_len(off0[-1])
    off0: np.ndarray           # int64 [n_pos + 1]
    vals1: np.ndarray
    off1: np.ndarray
    pos: np.ndarray            # int32 [n_pos], 0-based
    seg: np.ndarray            # int32 [n_pos], index into seg_names, non-decreasing
    base: np.ndarray           # uint8 [n_pos], ASCII of group 1's base (myDetect.py:436)
    seg_names: List[SegKey] = field(default_factory=list)
    # positions whose recorded base differs between the groups: (segment key, pos, base of group 1,
    # base of group 0) -- what mtest2 reports as 'Error not equal' (myDetect.py:432-434)
    base_mismatch: List[Tuple[SegKey, int, str, str]] = field(default_factory=list)
    # optional 16-bit transport format (see to_int16): value k stands for float32(k * i16_unit)
    vals0_i16: Optional[np.ndarray] = None
    vals1_i16: Optional[np.ndarray] = None
    i16_unit: float = 0.0

    @property
    def n_pos(self) -> int:
        return int(self.pos.shape[0])

    def counts(self) -> Tuple[np.ndarray, np.ndarray]:
        return np.diff(self.off0), np.diff(self.off1)

    def group(self, g: int, i: int) -> np.ndarray:
        """Values of group g at candidate i (float32 view)."""
        vals, off = (self.vals0, self.off0) if g == 0 else (self.vals1, self.off1)
        return vals[off[i]:off[i + 1]]

    def validate(self) -> None:
        n = self.n_pos
        assert self.off0.dtype == np.int64 and self.off1.dtype == np.int64
        assert self.off0.shape == (n + 1,) and self.off1.shape == (n + 1,)
        assert self.vals0.dtype == np.float32 and self.vals1.dtype == np.float32
        assert self.pos.dtype == np.int32 and self.seg.dtype == np.int32 and self.base.dtype == np.uint8
        assert self.off0[0] == 0 and self.off1[0] == 0
        assert np.all(np.diff(self.off0) >= 0) and np.all(np.diff(self.off1) >= 0)
        assert self.vals0.shape[0] >= padded_len(self.off0[-1])
        assert self.vals1.shape[0] >= padded_len(self.off1[-1])
        assert np.all(np.diff(self.seg) >= 0) and (n == 0 or self.seg.min() >= 0)
        # the device indexes per-segment arrays (seg_cov, chromosome names) with these ids
        assert n == 0 or int(self.seg.max()) < len(self.seg_names), "segment id without a seg_names entry"

    # ---- construction -------------------------------------------------------------------
    @staticmethod
    def _pad(values: np.ndarray) -> np.ndarray:
        out = np.zeros(padded_len(values.shape[0]), dtype=np.float32)
        out[:values.shape[0]] = values
        return out

    @classmethod
    def from_arrays(cls, vals0, off0, vals1, off1, pos, seg=None, base=None,
                    seg_names: Optional[List[SegKey]] = None) -> "Pileup":
        off0 = np.ascontiguousarray(off0, dtype=np.int64)
        off1 = np.ascontiguousarray(off1, dtype=np.int64)
        n = off0.shape[0] - 1
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        seg = np.zeros(n, np.int32) if seg is None else np.ascontiguousarray(seg, dtype=np.int32)
        if base is None:
            base = np.frombuffer(b"ACGT", dtype=np.uint8)[pos % 4].copy()
        v0 = np.asarray(vals0, dtype=np.float32)
        v1 = np.asarray(vals1, dtype=np.float32)
        if v0.shape[0] < padded_len(off0[-1]):
            v0 = cls._pad(v0[:off0[-1]])
        if v1.shape[0] < padded_len(off1[-1]):
            v1 = cls._pad(v1[:off1[-1]])
        p = cls(np.ascontiguousarray(v0), off0, np.ascontiguousarray(v1), off1, pos, seg,
                np.ascontiguousarray(base, dtype=np.uint8),
                list(seg_names) if seg_names is not None else
                [("syn", "+")] + [("syn%d" % k, "+") for k in range(1, int(seg.max()) + 1 if n else 1)])
        p.validate()
        return p

    @classmethod
    def from_dicts(cls, ds0: Dict, ds1: Dict) -> "Pileup":
        """From the reference's data model: ``dsK = {'norm_mean': {(chrom,strand): {pos: list}},
        'base': {(chrom,strand): {pos: 'A'}}}`` (what ReadAllFast5 builds, myDetect.py:569-572).
        Row order follows mtest2 (:421-431): sorted strand keys of ds0 that ds1 also has, then
        sorted positions of ds0 that ds1 also has."""
        seg_names: List[SegKey] = []
        pos: List[int] = []
        seg: List[int] = []
        base: List[int] = []
        c0: List[int] = []
        c1: List[int] = []
        chunks0: List[np.ndarray] = []
        chunks1: List[np.ndarray] = []
        mismatch: List[Tuple[SegKey, int, str, str]] = []
        for sk in sorted(ds0["norm_mean"].keys()):
            if sk not in ds1["norm_mean"]:
                continue
            d0, d1 = ds0["norm_mean"][sk], ds1["norm_mean"][sk]
            sid = len(seg_names)
            seg_names.append(sk)
            for pk in sorted(d0.keys()):
                if pk not in d1:
                    continue
                a = np.asarray(d0[pk], dtype=np.float32)
                b = np.asarray(d1[pk], dtype=np.float32)
                pos.append(pk)
                seg.append(sid)
                base.append(ord(ds1["base"][sk][pk]))
                if ds1["base"][sk][pk] != ds0["base"][sk][pk]:
                    mismatch.append((sk, pk, ds1["base"][sk][pk], ds0["base"][sk][pk]))
                c0.append(a.shape[0])
                c1.append(b.shape[0])
                chunks0.append(a)
                chunks1.append(b)
        off0 = np.zeros(len(pos) + 1, np.int64)
        off1 = np.zeros(len(pos) + 1, np.int64)
        np.cumsum(c0, out=off0[1:])
        np.cumsum(c1, out=off1[1:])
        v0 = np.concatenate(chunks0) if chunks0 else np.zeros(0, np.float32)
        v1 = np.concatenate(chunks1) if chunks1 else np.zeros(0, np.float32)
        p = cls.from_arrays(v0, off0, v1, off1, np.asarray(pos, np.int32),
                            np.asarray(seg, np.int32), np.asarray(base, np.uint8), seg_names)
        p.base_mismatch = mismatch
        return p

    def to_dicts(self) -> Tuple[Dict, Dict]:
        """Inverse of from_dicts (float32 values upcast to Python floats) -- used to feed the
        oracle the identical inputs."""
        out = []
        for g in (0, 1):
            nm: Dict = {}
            bs: Dict = {}
            vals, off = (self.vals0, self.off0) if g == 0 else (self.vals1, self.off1)
            for i in range(self.n_pos):
                if off[i + 1] == off[i]:
                    continue
                sk = self.seg_names[self.seg[i]]
                nm.setdefault(sk, {})[int(self.pos[i])] = [float(x) for x in vals[off[i]:off[i + 1]]]
                bs.setdefault(sk, {})[int(self.pos[i])] = chr(self.base[i])
            out.append({"norm_mean": nm, "base": bs})
        return out[0], out[1]

    def to_int16(self, unit: float = 0.001) -> "Pileup":
        """The same pileup with its values ALSO in the 16-bit transport format (half the bytes over
        PCIe): every value must be exactly float32(k * unit) for an integer |k| <= 32767 -- true for
        the reference's data (norm_mean = round(x, 3), myRefBaseSignalAnnotation.py:1108) cast to
        float32 with unit = 0.001.  ``Detector.detect`` then ships the int16 arrays and the library
        expands them on the GPU: results are bit-identical to the float32 pileup's."""
        out = []
        for vals, off in ((self.vals0, self.off0), (self.vals1, self.off1)):
            n = int(off[-1])
            k = np.rint(vals[:n].astype(np.float64) / unit)
            if n and (np.abs(k).max() > 32767 or not np.array_equal((k * unit).astype(np.float32), vals[:n])):
                raise ValueError("values are not on the %g grid within +-32767 units: no exact int16 form" % unit)
            a = np.zeros((n + 15) // 8 * 8, dtype=np.int16)
            a[:n] = k.astype(np.int16)
            out.append(a)
        return Pileup(self.vals0, self.off0, self.vals1, self.off1, self.pos, self.seg, self.base, self.seg_names,
                      self.base_mismatch, out[0], out[1], float(unit))

    def save_npz(self, path: str) -> None:
        np.savez(path, vals0=self.vals0[:self.off0[-1]], off0=self.off0, vals1=self.vals1[:self.off1[-1]],
                 off1=self.off1, pos=self.pos, seg=self.seg, base=self.base,
                 seg_names=np.asarray([list(s) for s in self.seg_names]))

    @classmethod
    def load_npz(cls, path: str) -> "Pileup":
        z = np.load(path)
        return cls.from_arrays(z["vals0"], z["off0"], z["vals1"], z["off1"], z["pos"], z["seg"],
                               z["base"], [tuple(s) for s in z["seg_names"].tolist()])

    def slice_rows(self, lo: int, hi: int) -> "Pileup":
        """Candidates [lo, hi) as a self-contained pileup (used for genome sharding)."""
        o0 = self.off0[lo:hi + 1] - self.off0[lo]
        o1 = self.off1[lo:hi + 1] - self.off1[lo]
        return Pileup.from_arrays(self.vals0[self.off0[lo]:self.off0[hi]], o0,
                                  self.vals1[self.off1[lo]:self.off1[hi]], o1, self.pos[lo:hi],
                                  self.seg[lo:hi], self.base[lo:hi], self.seg_names)


# -----------------------------------------------------------------------------------------
# synthetic pileups: Gaussian currents, planted shifted sites
# -----------------------------------------------------------------------------------------
SYN_SEED = 20190131


def planted_shift(pos: np.ndarray, period: int = 1000, phase: int = 500) -> np.ndarray:
    """Mean shift of group 1 at each position: +1.0 at pos % period == phase, +0.5 at +-1,
    +0.25 at +-2 (neighbour effect), else 0."""
    d = np.abs(((pos.astype(np.int64) - phase + period // 2) % period) - period // 2)
    return np.select([d == 0, d == 1, d == 2], [1.0, 0.5, 0.25], 0.0)


def synthetic_pileup(length: int, n0: int, n1: int, seed: int = SYN_SEED, *, round_decimals: Optional[int] = None,
                     drop_frac1: float = 0.0, poisson: bool = False, clip: Tuple[int, int] = (5, 160),
                     extra_shift: Optional[Dict[int, float]] = None, two_strands: bool = False) -> Pileup:
    """numpy (PCG64) generator for the small configs.  ``poisson``: per-position coverage
    ~ Poisson(n) clipped to ``clip``.  ``drop_frac1``: delete that fraction of positions from
    group 1 (gaps).  ``extra_shift``: {pos: shift} added on top of the planted pattern."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = np.arange(length, dtype=np.int32)
    if poisson:
        c0 = np.clip(rng.poisson(n0, length), clip[0], clip[1]).astype(np.int64)
        c1 = np.clip(rng.poisson(n1, length), clip[0], clip[1]).astype(np.int64)
    else:
        c0 = np.full(length, n0, np.int64)
        c1 = np.full(length, n1, np.int64)
    if drop_frac1 > 0:
        c1[rng.random(length) < drop_frac1] = 0
    off0 = np.concatenate([[0], np.cumsum(c0)]).astype(np.int64)
    off1 = np.concatenate([[0], np.cumsum(c1)]).astype(np.int64)
    v0 = rng.standard_normal(off0[-1]).astype(np.float32)
    shift = planted_shift(pos)
    if extra_shift:
        for p, s in extra_shift.items():
            shift[p] += s
    v1 = (rng.standard_normal(off1[-1]) + np.repeat(shift, c1)).astype(np.float32)
    if round_decimals is not None:
        v0 = np.round(v0.astype(np.float64), round_decimals).astype(np.float32)
        v1 = np.round(v1.astype(np.float64), round_decimals).astype(np.float32)
    if two_strands:
        half = length // 2
        seg = (np.arange(length) >= half).astype(np.int32)
        pos = np.where(seg == 0, pos, pos - half).astype(np.int32)
        names = [("syn", "+"), ("syn", "-")]
    else:
        seg = np.zeros(length, np.int32)
        names = [("syn", "+")]
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[pos % 4].copy()
    return Pileup.from_arrays(v0, off0, v1, off1, pos, seg, base, names)
