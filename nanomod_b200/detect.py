"""Host side of the detection stage: options, the per-position result table, ranking, the
called-site rule and the text writer -- everything of ``mtest2`` / ``save_test``
(bin/scripts/myDetect.py:416-545) that is not arithmetic.  The arithmetic runs in the CUDA
library behind ``nanomod_b200._lib`` (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .pileup import Pileup


class OptionError(ValueError):
    """Invalid option; the message is the reference's own text (NanoMod.py:40-97, 148-153)."""


@dataclass
class DetectOptions:
    """Options of ``NanoMod.py detect`` that the stage reads (names and defaults verbatim:
    NanoMod.py:348-366, 377-392).  ``window`` is the command-line value (21), stored by the
    reference as (W-1)/2 (:51) -- see ``half_window``."""
    wrkBase1: Optional[str] = None
    wrkBase2: Optional[str] = None
    outLevel: int = 2
    window: int = 21
    FileID: str = "mod"
    outFolder: str = "mRes/"
    MinCoverage: int = 5
    topN: int = 30
    neighborPvalues: int = 2
    WeightsDif: float = 2.0
    testMethod: str = "stouffer"
    rankUse: str = "pv"
    SaveTest: int = 1
    RegionRankbyST: int = 0
    percentile: float = 0.1
    WindOvlp: int = 0
    NA: str = ""
    Pos: str = ""
    mstd: bool = False
    plotType: str = "Density"
    min_lr: int = 500
    min_lr_nb: int = 0
    downsampling_quantile: float = 0.25
    downsampling: int = 100
    coverages: str = "0-0"
    # not a reference option: seed of the counter-based stream the down-sampling branch draws from
    # (the reference uses numpy's unseeded global generator, myDetect.py:351)
    seed: int = 20190131
    # not reference options: which of the reference's three per-position tests to compute.
    # The reference always computes all three (myDetect.py:331-341).
    want_u: bool = True
    want_t: bool = True
    # compute both combinations in one pass (BASELINE config 3); ranking still follows testMethod
    both_combinations: bool = False

    @property
    def half_window(self) -> int:
        return (self.window - 1) // 2  # NanoMod.py:51 (Python 2 integer division)

    def validate(self) -> None:
        """mCommonParam's checks (NanoMod.py:40-97); raises OptionError with the same text."""
        err = ""
        if self.half_window < 1:
            err += "\n\tWindow size (%d) is too small" % self.window
        if self.MinCoverage < 3:
            err += "\n\tThe coverage (%d) is too small" % self.MinCoverage
        if self.topN < 1:
            err += "\n\tThe topN (%d) is too smaller" % self.topN
        if self.neighborPvalues < 0:
            err += "\n\tThe neighborPvalues (%d) cannot be smaller than 0" % self.neighborPvalues
        if self.testMethod not in ("fisher", "stouffer", "ks"):
            err += "\n\ttestMethod must be one of fisher, stouffer, ks"
        if self.rankUse not in ("st", "pv"):
            err += "\n\trankUse must be one of st, pv"
        if self.WeightsDif < 1:
            self.WeightsDif = 1.0  # NanoMod.py:77-78
        if self.percentile < 0:
            self.percentile = 0.0
        if self.percentile >= 1:
            self.percentile = 0.99
        if any(c > 0 for c in self.coverage_pair()):
            if not 1 <= int(self.downsampling) <= _lib.NM_DS_MAX_TIMES:
                err += "\n\tdownsampling (%d) must be in [1, %d]" % (self.downsampling, _lib.NM_DS_MAX_TIMES)
            elif not 0 <= int(self.downsampling * self.downsampling_quantile) < int(self.downsampling):
                err += "\n\tdownsampling_quantile (%s) must be in [0, 1)" % self.downsampling_quantile
        if err:
            raise OptionError("Please provide correct parameters" + err)

    def coverage_pair(self) -> Tuple[int, int]:
        """moptions['coverages'] = [cov of '+' strands, cov of '-' strands] (NanoMod.py:174-176)."""
        cov = [int(x) for x in str(self.coverages).split("-")]
        if len(cov) == 1:
            cov = [cov[0], cov[0]]
        return cov[0], cov[1]

    def seg_cov(self, seg_names) -> Optional[np.ndarray]:
        """Per-segment down-sampling coverage (myDetect.py:339), None when it is off."""
        cp, cm = self.coverage_pair()
        if cp <= 0 and cm <= 0:
            return None
        return np.array([cp if sk[1] == "+" else cm for sk in seg_names], dtype=np.int32)

    def combine_mask(self) -> int:
        if self.both_combinations:
            return _lib.NM_COMBINE_FISHER | _lib.NM_COMBINE_STOUFFER
        return {"ks": _lib.NM_COMBINE_NONE, "fisher": _lib.NM_COMBINE_FISHER,
                "stouffer": _lib.NM_COMBINE_STOUFFER}[self.testMethod]

    def to_params(self) -> _lib.nm_params:
        return _lib.nm_params(int(self.MinCoverage), int(self.neighborPvalues), float(self.WeightsDif),
                              int(self.combine_mask()), int(bool(self.want_u)), int(bool(self.want_t)), 0,
                              int(self.downsampling), int(self.downsampling * self.downsampling_quantile),
                              int(self.seed))

    @classmethod
    def from_moptions(cls, moptions: Dict) -> "DetectOptions":
        o = cls()
        for k in ("outLevel", "FileID", "outFolder", "MinCoverage", "topN", "neighborPvalues",
                  "WeightsDif", "testMethod", "rankUse", "SaveTest", "RegionRankbyST", "percentile",
                  "WindOvlp", "NA", "mstd", "downsampling", "downsampling_quantile", "seed"):
            if k in moptions:
                setattr(o, k, moptions[k])
        if "window" in moptions:  # moptions stores the half window
            o.window = 2 * int(moptions["window"]) + 1
        if "coverages" in moptions:
            o.coverages = "-".join(str(int(c)) for c in moptions["coverages"])
        return o


@dataclass
class SignTestTable:
    """The per-position result table == ``moptions['sign_test']`` in SoA form (row order
    identical).  Columns not computed are None."""
    options: DetectOptions
    seg_names: List[Tuple[str, str]]
    seg: np.ndarray
    pos: np.ndarray
    base: np.ndarray
    n0: np.ndarray
    n1: np.ndarray
    ks_dnum: np.ndarray
    ks_d: np.ndarray
    ks_p: np.ndarray
    two_u: Optional[np.ndarray] = None
    u_stat: Optional[np.ndarray] = None
    u_p: Optional[np.ndarray] = None
    t_stat: Optional[np.ndarray] = None
    t_p: Optional[np.ndarray] = None
    fisher_stat: Optional[np.ndarray] = None
    fisher_p: Optional[np.ndarray] = None
    stouffer_stat: Optional[np.ndarray] = None
    stouffer_p: Optional[np.ndarray] = None
    flags: Optional[np.ndarray] = None
    row_pos_index: Optional[np.ndarray] = None
    moments: Optional[np.ndarray] = None  # [rows, 4]: mean0, var0, mean1, var1 (ddof=1); --mstd

    def __len__(self) -> int:
        return int(self.pos.shape[0])

    # ---- the combined column selected by testMethod (4th tuple of a sign_test row) ----------
    def comb(self) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        m = self.options.testMethod
        if m == "fisher":
            return self.fisher_stat, self.fisher_p
        if m == "stouffer":
            return self.stouffer_stat, self.stouffer_p
        return None

    def to_sign_test(self) -> List:
        """``moptions['sign_test']`` exactly as mtest2 builds it (myDetect.py:436, :377)."""
        comb = self.comb()
        zeros = np.zeros(len(self))
        u_s = self.u_stat if self.u_stat is not None else zeros
        u_p = self.u_p if self.u_p is not None else zeros
        t_s = self.t_stat if self.t_stat is not None else zeros
        t_p = self.t_p if self.t_p is not None else zeros
        out = []
        for r in range(len(self)):
            sk = self.seg_names[self.seg[r]]
            tests = [(float(u_s[r]), float(u_p[r])), (float(t_s[r]), float(t_p[r])),
                     (float(self.ks_d[r]), float(self.ks_p[r]))]
            if comb is not None:
                tests.append((float(comb[0][r]), float(comb[1][r])))
            out.append(((sk[0], sk[1], int(self.pos[r]), chr(self.base[r]), int(self.n0[r]),
                         int(self.n1[r])), tests))
        return out

    # ---- save_test (myDetect.py:522-538) ---------------------------------------------------
    def format_lines(self) -> List[str]:
        comb = self.comb()
        with_comb = self.options.neighborPvalues > 0 and comb is not None
        zeros = np.zeros(len(self))
        u_s = self.u_stat if self.u_stat is not None else zeros
        u_p = self.u_p if self.u_p is not None else zeros
        t_s = self.t_stat if self.t_stat is not None else zeros
        t_p = self.t_p if self.t_p is not None else zeros
        lines = []
        for r in range(len(self)):
            sk = self.seg_names[self.seg[r]]
            s = "%s %s %d %s %d %d %.3f %.3E %.3f %.3E %.3f %.3E" % (
                sk[0], sk[1], self.pos[r] + 1, chr(self.base[r]), self.n0[r], self.n1[r],
                u_s[r], u_p[r], t_s[r], t_p[r], self.ks_d[r], self.ks_p[r])
            if with_comb:
                s += " %.3f %.3E\n" % (comb[0][r], comb[1][r])
            else:
                s += "\n"
            lines.append(s)
        return lines

    def format_text(self, n_threads: int = 0) -> bytes:
        """The same text as ``''.join(format_lines())``, produced by the library's multithreaded
        native writer (nm_format_sign_test): the reference formats and flushes one line at a time."""
        return self._format_text_array(n_threads).tobytes()

    def _format_text_array(self, n_threads: int = 0) -> np.ndarray:
        comb = self.comb()
        with_comb = self.options.neighborPvalues > 0 and comb is not None
        return _lib.format_sign_test(self.seg_names, self.seg, self.pos, self.base, self.n0, self.n1, self.u_stat,
                                     self.u_p, self.t_stat, self.t_p, self.ks_d, self.ks_p,
                                     comb[0] if with_comb else None, comb[1] if with_comb else None, n_threads)

    def save_test(self, path: Optional[str] = None) -> Optional[str]:
        """Write ``<outFolder>/<FileID>_sign_test.txt`` when SaveTest != 0."""
        if self.options.SaveTest == 0:
            return None
        if path is None:
            os.makedirs(self.options.outFolder, exist_ok=True)
            path = self.options.outFolder + "/" + self.options.FileID + "_sign_test.txt"
        with open(path, "wb") as f:
            f.write(self._format_text_array())
        self.save_meanstd()  # the reference writes both files from save_test (:540-545)
        return path

    # ---- --mstd output (myDetect.py:437-438, :540-545) --------------------------------------
    def mean_std(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """np.mean and np.std (ddof=0) of each group per row, from the device's two-pass moments
        (var is ddof=1 there: std0 = sqrt(var * (n-1) / n))."""
        if self.moments is None:
            raise ValueError("the table was computed without mstd")
        m = self.moments
        n0 = self.n0.astype(np.float64)
        n1 = self.n1.astype(np.float64)
        return (m[:, 0], np.sqrt(m[:, 1] * (n0 - 1.0) / n0), m[:, 2], np.sqrt(m[:, 3] * (n1 - 1.0) / n1))

    def sign_test_mstd(self) -> Dict:
        """``moptions['sign_test_mstd']`` as mtest2 fills it (myDetect.py:437-438)."""
        m0, s0, m1, s1 = self.mean_std()
        out = {}
        for r in range(len(self)):
            sk = self.seg_names[self.seg[r]]
            out[(sk[0], sk[1], int(self.pos[r]))] = [[float(m0[r]), float(s0[r])], [float(m1[r]), float(s1[r])]]
        return out

    def meanstd_lines(self) -> List[str]:
        """Lines of ``<FileID>_meanstd.cvs`` (myDetect.py:540-545; position is 0-based there)."""
        m0, s0, m1, s1 = self.mean_std()
        lines = []
        for r in range(len(self)):
            sk = self.seg_names[self.seg[r]]
            lines.append("%s %s %d %s %.3f %.3f %.3f %.3f\n" % (sk[0], sk[1], self.pos[r], chr(self.base[r]),
                                                             m0[r], s0[r], m1[r], s1[r]))
        return lines

    def save_meanstd(self, path: Optional[str] = None) -> Optional[str]:
        if not self.options.mstd or self.moments is None:
            return None
        if path is None:
            os.makedirs(self.options.outFolder, exist_ok=True)
            path = self.options.outFolder + "/" + self.options.FileID + "_meanstd.cvs"
        with open(path, "w") as f:
            f.writelines(self.meanstd_lines())
        return path

    # ---- ranking (myDetect.py:447-462, RegionRankbyST == 0) --------------------------------
    def ranked(self) -> np.ndarray:
        """Row indices in the order of ``moptions['sorted_sign_test']``: stable sort by
        (combined, KS, U) on the p-value ('pv') or the statistic ('st', then reversed)."""
        use_p = self.options.rankUse == "pv"
        comb = self.comb()
        ks = self.ks_p if use_p else self.ks_d
        if self.u_p is not None:
            u = self.u_p if use_p else self.u_stat
        else:
            u = np.zeros(len(self))
        keys = [u, ks]
        if comb is not None:
            keys.append(comb[1] if use_p else comb[0])
        order = np.lexsort(tuple(keys))  # last key is primary; lexsort is stable
        if not use_p:
            order = order[::-1]
        return order

    # ---- region ranking (myDetect.py:463-515, RegionRankbyST == 1) --------------------------
    def region_ranked(self) -> np.ndarray:
        """Row indices of the window centres in the order of ``moptions['sorted_sign_test']`` when
        RegionRankbyST is 1: windows of half-width w = half_window + 1 (the reference increments
        moptions['window'], :465) centred every w positions (every position with WindOvlp=1) that
        are fully covered inside one (chrom, strand); each window is scored by the percentile-th
        smallest value of its (optionally base-filtered, --NA) combined p-values or statistics;
        ties by the distance of the window's minimum from the centre; with WindOvlp=1 a window
        closer than w to a better-ranked one is dropped (:503-511)."""
        o = self.options
        w = o.half_window + 1
        move = 1 if o.WindOvlp == 1 else w
        use_p = o.rankUse == "pv"
        comb = self.comb()
        val = (comb[1] if use_p else comb[0]) if comb is not None else (self.ks_p if use_p else self.ks_d)
        val = np.asarray(val, dtype=np.float64)
        na = ord(o.NA) if (o.NA is not None and len(o.NA) > 0) else None
        n = len(self)
        cent, score, dist = [], [], []
        bounds = np.flatnonzero(np.diff(self.seg)) + 1
        for lo, hi in zip(np.append(0, bounds), np.append(bounds, n)):
            pos = self.pos[lo:hi].astype(np.int64)
            if hi - lo < 2 * w + 1:
                continue
            pmin, pmax = pos[0], pos[-1]
            r = np.arange(w, hi - lo - w)  # rows that have w rows on either side
            ok = (pos[r + w] - pos[r - w] == 2 * w) & (pos[r + w] < pmax) & ((pos[r] - pmin) % move == 0) & (pos[r] < pmax)
            r = r[ok]
            if len(r) == 0:
                continue
            win = r[:, None] + np.arange(-w, w + 1)[None, :]
            v = val[lo:hi][win]
            keep = np.ones_like(v, dtype=bool) if na is None else (self.base[lo:hi][win] == na)
            cnt = keep.sum(axis=1)
            vm = np.where(keep, v, np.inf)
            first_min = np.argmin(vm, axis=1)  # first occurrence, window order
            srt = np.sort(vm, axis=1)
            good = cnt > 5
            k = (o.percentile * (cnt - 1) + 0.5).astype(np.int64)
            k = np.clip(k, 0, 2 * w)
            sc = srt[np.arange(len(r)), k]
            before = np.cumsum(keep, axis=1) - keep  # filtered elements before each window slot
            ds = np.abs(w - before[np.arange(len(r)), first_min])
            cent.append((r + lo)[good])
            score.append(sc[good])
            dist.append(ds[good])
        if not cent:
            return np.zeros(0, dtype=np.int64)
        cent, score, dist = np.concatenate(cent), np.concatenate(score), np.concatenate(dist)
        order = cent[np.lexsort((dist, score))]
        if o.WindOvlp != 1:
            return order
        out = []
        taken: Dict[int, np.ndarray] = {}
        for r in order:
            sg, ps = int(self.seg[r]), int(self.pos[r])
            m = taken.get(sg)
            if m is None:
                sel = self.seg == sg
                m = taken[sg] = np.zeros(int(self.pos[sel].max()) + 2 * w + 2, dtype=bool)
            a, b = max(0, ps - w + 1), ps + w
            if m[a:b].any():
                continue
            m[ps] = True
            out.append(r)
        return np.asarray(out, dtype=np.int64)

    def sorted_rows(self) -> np.ndarray:
        """Row order of ``moptions['sorted_sign_test']`` for the table's options."""
        return self.region_ranked() if self.options.RegionRankbyST != 0 else self.ranked()

    # ---- called sites (mboxplot :279-297 + plot1 :153-164) ---------------------------------
    def called_sites(self) -> List[Tuple[str, str, int]]:
        o = self.options
        closesize = o.neighborPvalues * 2
        # region mode: mtest2 leaves moptions['window'] incremented by one (:465); mboxplot then
        # uses closesize = window (:280-282) and plot1 2*window rows either side (:153-154)
        nearby = o.half_window + (1 if o.RegionRankbyST != 0 else 0)
        if o.RegionRankbyST == 1:
            closesize = max(1, nearby)
            nearby = 2 * nearby
        n = len(self)
        out: List[Tuple[str, str, int]] = []
        acc_seg: List[int] = []
        acc_pos: List[int] = []
        for r in self.sorted_rows():
            r = int(r)
            sg, ps = int(self.seg[r]), int(self.pos[r])
            if any(s == sg and abs(p - ps) < closesize for s, p in zip(acc_seg, acc_pos)):
                continue
            lo, hi = r - nearby, r + nearby
            okk = lo >= 0 and hi <= n - 1
            if okk:
                w = slice(lo, hi + 1)
                okk = bool(np.all(self.seg[w] == sg) and
                           np.all(self.pos[w].astype(np.int64) - ps == np.arange(-nearby, nearby + 1)))
            if okk:
                sk = self.seg_names[sg]
                out.append((sk[0], sk[1], ps))
                acc_seg.append(sg)
                acc_pos.append(ps)
            if len(out) == self.options.topN:
                break
        return out


_TABLE_OPTIONAL = {"two_u": "want_u", "u_stat": "want_u", "u_p": "want_u", "t_stat": "want_t",
                   "t_p": "want_t"}


def _wanted_columns(opt: DetectOptions) -> List[str]:
    cols = ["row_pos_index", "n0", "n1", "ks_dnum", "ks_d", "ks_p", "flags"]
    if opt.want_u:
        cols += ["two_u", "u_stat", "u_p"]
    if opt.want_t:
        cols += ["t_stat", "t_p"]
    mask = opt.combine_mask()
    if mask & _lib.NM_COMBINE_FISHER:
        cols += ["fisher_stat", "fisher_p"]
    if mask & _lib.NM_COMBINE_STOUFFER:
        cols += ["stouffer_stat", "stouffer_p"]
    if opt.mstd:
        cols += ["moments"]
    return cols


def _col_shape(c: str, n: int):
    w = _lib.TABLE_WIDTH.get(c, 1)
    return (n, w) if w > 1 else (n,)


# numpy view of the 28-byte records of Detector.pack_records
RECORD_DTYPE = np.dtype({"names": ["ks_dnum", "ks_p", "comb_stat", "comb_p"], "formats": ["<i4", "<f8", "<f8", "<f8"],
                         "offsets": [0, 4, 12, 20], "itemsize": 28})


class Detector:
    """One GPU's detection engine.  ``detect`` is the host-buffer call (numpy in, numpy out,
    H2D/D2H inside); ``detect_device`` works on torch CUDA tensors that are already resident."""

    def __init__(self, device: int = 0):
        self.handle = _lib.Handle(device)
        self.device = int(device)

    @property
    def launch_count(self) -> int:
        return self.handle.launch_count

    def detect(self, pileup: Pileup, options: Optional[DetectOptions] = None,
               out: Optional[Dict[str, np.ndarray]] = None) -> SignTestTable:
        opt = options or DetectOptions()
        opt.validate()
        n = pileup.n_pos
        cols = _wanted_columns(opt)
        if out is None:
            out = {c: np.empty(_col_shape(c, n), dtype=_lib.TABLE_DTYPES[c]) for c in cols}
        tb = _lib.nm_table(**{c: out[c].ctypes.data for c in cols})
        seg_cov = opt.seg_cov(pileup.seg_names)
        i16 = pileup.vals0_i16 is not None and pileup.vals1_i16 is not None
        pl = _lib.nm_pileup(None if i16 else pileup.vals0.ctypes.data, pileup.off0.ctypes.data,
                            None if i16 else pileup.vals1.ctypes.data,
                            pileup.off1.ctypes.data, pileup.pos.ctypes.data, pileup.seg.ctypes.data, n,
                            None if seg_cov is None else seg_cov.ctypes.data,
                            0 if seg_cov is None else len(seg_cov),
                            pileup.vals0_i16.ctypes.data if i16 else None, pileup.vals1_i16.ctypes.data if i16 else None,
                            pileup.i16_unit if i16 else 0.0, int(pileup.off0[-1]), int(pileup.off1[-1]))
        n_rows = self.handle.detect_host(pl, opt.to_params(), tb)
        res = {c: out[c][:n_rows] for c in cols}
        if n_rows == n:  # nothing filtered: rows are the candidates themselves
            seg, pos, base = pileup.seg, pileup.pos, pileup.base
        else:
            idx = res["row_pos_index"]
            seg, pos, base = pileup.seg[idx], pileup.pos[idx], pileup.base[idx]
        return SignTestTable(options=opt, seg_names=pileup.seg_names, seg=seg, pos=pos, base=base,
                             **{c: res[c] for c in cols})

    def rank(self, table: SignTestTable) -> np.ndarray:
        """Row order of ``moptions['sorted_sign_test']`` (myDetect.py:459-461) computed on the GPU:
        same result as ``SignTestTable.ranked()`` (numpy), which it replaces for large tables."""
        use_p = table.options.rankUse == "pv"
        comb = table.comb()
        cols = [None if comb is None else (comb[1] if use_p else comb[0]),
                table.ks_p if use_p else table.ks_d,
                None if table.u_p is None else (table.u_p if use_p else table.u_stat)]
        cols = [None if c is None else np.ascontiguousarray(c, dtype=np.float64) for c in cols]
        order = np.empty(len(table), dtype=np.int32)
        if len(table):
            self.handle.rank_host(*[None if c is None else c.ctypes.data for c in cols], len(table),
                                  not use_p, order.ctypes.data)
        return order

    def rank_device(self, out: Dict[str, "object"], n_rows: int, options: DetectOptions,
                    stream: Optional[int] = None):
        """Ranking of a device-resident table (``out`` as filled by ``detect_device``): returns an
        int32 CUDA tensor of row indices."""
        import torch
        use_p = options.rankUse == "pv"
        m = options.testMethod
        comb = None if m == "ks" else out[("fisher" if m == "fisher" else "stouffer") + ("_p" if use_p else "_stat")]
        ks = out["ks_p" if use_p else "ks_d"]
        u = out.get("u_p" if use_p else "u_stat")
        order = torch.empty(n_rows, dtype=torch.int32, device=ks.device)
        if stream is None:
            stream = torch.cuda.current_stream(ks.device).cuda_stream
        self.handle.rank_device(None if comb is None else comb.data_ptr(), ks.data_ptr(),
                                None if u is None else u.data_ptr(), n_rows, not use_p, order.data_ptr(), stream)
        return order

    def rank_head_device(self, out: Dict[str, "object"], n_rows: int, options: DetectOptions, want: int,
                         stream: Optional[int] = None, geometry=None) -> np.ndarray:
        """The first rows (>= ``want``) of ``rank_device``'s order without sorting the whole
        table.  Returns a structured array (``_lib.HEAD_ROW_DTYPE``): ``row`` indices in ranking
        order, plus -- with ``geometry = (row_pos_index | None, pos, seg, row_offset, n_rows_total,
        nearby)`` (device tensors) -- each row's segment, position and plot1's neighbourhood flag.
        Every row that ties with the cut's exponent bin is included, so rows not returned rank
        strictly after all rows returned."""
        import torch
        use_p = options.rankUse == "pv"
        m = options.testMethod
        comb = None if m == "ks" else out[("fisher" if m == "fisher" else "stouffer") + ("_p" if use_p else "_stat")]
        ks = out["ks_p" if use_p else "ks_d"]
        u = out.get("u_p" if use_p else "u_stat") if options.want_u else None
        if stream is None:
            stream = torch.cuda.current_stream(ks.device).cuda_stream
        geo = None
        if geometry is not None:
            rpi, pos, seg, row_offset, n_total, nearby = geometry
            geo = _lib.nm_head_geometry(None if rpi is None else rpi.data_ptr(), pos.data_ptr(), seg.data_ptr(),
                                        int(row_offset), int(n_total), int(nearby), 0)
        cap = max(4 * want, 16384)
        while True:
            rows = np.empty(cap, dtype=_lib.HEAD_ROW_DTYPE)
            try:
                n = self.handle.rank_head_device(None if comb is None else comb.data_ptr(), ks.data_ptr(),
                                                 None if u is None else u.data_ptr(), n_rows, not use_p, want, geo,
                                                 rows.ctypes.data, cap, stream)
                return rows[:n]
            except _lib.NmError as e:  # more ties at the cut than `cap`
                if e.code != _lib.NM_ERR_BAD_ARG or cap >= n_rows:
                    raise
                cap = min(n_rows, cap * 8)

    def rank_head_select_device(self, out: Dict[str, "object"], n_rows: int, options: DetectOptions, want: int,
                                records, cap: int, geometry=None, stream: Optional[int] = None) -> None:
        """``rank_head_device`` without leaving the GPU and without waiting: fills ``records`` (uint8
        CUDA tensor of (cap + 1) * 48 bytes, ``sharded.HEAD_REC`` entries; entry 0 is a header) with
        the UNSORTED rows of the ranking's head -- what a rank hands to the all-gather."""
        import torch
        use_p = options.rankUse == "pv"
        m = options.testMethod
        comb = None if m == "ks" else out[("fisher" if m == "fisher" else "stouffer") + ("_p" if use_p else "_stat")]
        ks = out["ks_p" if use_p else "ks_d"]
        u = out.get("u_p" if use_p else "u_stat") if options.want_u else None
        if stream is None:
            stream = torch.cuda.current_stream(ks.device).cuda_stream
        geo = None
        if geometry is not None:
            rpi, pos, seg, row_offset, n_total, nearby = geometry
            geo = _lib.nm_head_geometry(None if rpi is None else rpi.data_ptr(), pos.data_ptr(), seg.data_ptr(),
                                        int(row_offset), int(n_total), int(nearby), 0)
        self.handle.rank_head_select_device(None if comb is None else comb.data_ptr(), ks.data_ptr(),
                                            None if u is None else u.data_ptr(), n_rows, not use_p, want, geo,
                                            records.data_ptr(), cap, stream)

    def arm_head_select(self, out: Dict[str, "object"], row_lo: int, n_rows: int, options: DetectOptions, want: int,
                        records, cap: int, geometry) -> None:
        """Arm ``rank_head_select_device`` over rows [row_lo, row_lo + n_rows) of ``out`` for the NEXT
        ``detect_device`` call: that call launches the selection behind its own kernels, before its host wait,
        if its rows turn out to be its candidates (``handle.head_fired()`` tells)."""
        use_p = options.rankUse == "pv"
        m = options.testMethod
        sl = slice(row_lo, row_lo + n_rows)
        comb = None if m == "ks" else out[("fisher" if m == "fisher" else "stouffer") + ("_p" if use_p else "_stat")][sl]
        ks = out["ks_p" if use_p else "ks_d"][sl]
        u = out.get("u_p" if use_p else "u_stat")[sl] if options.want_u else None
        rpi, pos, seg, row_offset, n_total, nearby = geometry
        geo = _lib.nm_head_geometry(None if rpi is None else rpi.data_ptr(), pos.data_ptr(), seg.data_ptr(),
                                    int(row_offset), int(n_total), int(nearby), 0)
        self.handle.arm_head_select(None if comb is None else comb.data_ptr(), ks.data_ptr(),
                                    None if u is None else u.data_ptr(), n_rows, not use_p, want, geo,
                                    records.data_ptr(), cap)

    def pack_records(self, out: Dict[str, "object"], row_lo: int, n: int, options: DetectOptions, records,
                     stream: Optional[int] = None) -> None:
        """Device-resident table -> 28-byte records {ks_dnum, ks_p, comb stat, comb p} of rows
        [row_lo, row_lo + n) in ``records`` (uint8 CUDA tensor of 28*n bytes): what a rank sends to
        rank 0 in multi-GPU runs.  Record dtype: ``RECORD_DTYPE``."""
        import torch
        which = _lib.NM_COMBINE_FISHER if options.testMethod == "fisher" else _lib.NM_COMBINE_STOUFFER
        cols = [c for c in _wanted_columns(options) if c in out]
        tb = _lib.nm_table(**{c: out[c].data_ptr() for c in cols})
        if stream is None:
            stream = torch.cuda.current_stream(records.device).cuda_stream
        self.handle.pack_records_device(tb, row_lo, n, which, records.data_ptr(), stream)

    def detect_device(self, dev: "DevicePileup", options: DetectOptions, out: Dict[str, "object"],
                      stream: Optional[int] = None, _async: bool = False) -> int:
        """Device-resident call.  ``out`` maps column name -> torch CUDA tensor with capacity
        n_pos (see ``alloc_device_table``).  Returns n_rows; results are complete on return."""
        import torch
        cols = _wanted_columns(options)
        tb = _lib.nm_table(**{c: out[c].data_ptr() for c in cols})
        seg_cov = getattr(dev, "seg_cov", None)
        i16 = getattr(dev, "vals0_i16", None) is not None
        pl = _lib.nm_pileup(None if i16 else dev.vals0.data_ptr(), dev.off0.data_ptr(),
                            None if i16 else dev.vals1.data_ptr(),
                            dev.off1.data_ptr(), dev.pos.data_ptr(), dev.seg.data_ptr(), dev.n_pos,
                            None if seg_cov is None else seg_cov.data_ptr(),
                            0 if seg_cov is None else int(seg_cov.numel()),
                            dev.vals0_i16.data_ptr() if i16 else None, dev.vals1_i16.data_ptr() if i16 else None,
                            float(dev.i16_unit) if i16 else 0.0, int(dev.i16_total0) if i16 else 0,
                            int(dev.i16_total1) if i16 else 0)
        if stream is None:
            stream = torch.cuda.current_stream(dev.off0.device).cuda_stream
        if _async:
            return self.handle.detect_device_async(pl, options.to_params(), tb, stream)
        return self.handle.detect_device(pl, options.to_params(), tb, stream)

    def detect_device_async(self, dev: "DevicePileup", options: DetectOptions, out: Dict[str, "object"],
                            stream: Optional[int] = None) -> int:
        """``detect_device`` without the host wait (nm_detect_device_async): queues the call and returns a ticket
        for ``detect_finish``.  Up to two calls may be in flight, writing different ``out`` tables; ``dev`` and
        ``out`` must stay alive and untouched until the call is finished."""
        return self.detect_device(dev, options, out, stream, _async=True)

    def detect_finish(self, ticket: int):
        """Waits for the call behind ``ticket`` (re-running it if the device refused the assumed shape);
        returns (n_rows, head_fired)."""
        return self.handle.detect_finish(ticket)


@dataclass
class DevicePileup:
    """CSR pileup resident in HBM (torch CUDA tensors; vals padded per nm_padded_len)."""
    vals0: "object"
    off0: "object"
    vals1: "object"
    off1: "object"
    pos: "object"
    seg: "object"
    n_pos: int
    seg_cov: "object" = None  # optional int32 CUDA tensor [n_seg]: down-sampling coverage per segment
    # optional 16-bit transport format resident on the device (vals0 / vals1 may then be None)
    vals0_i16: "object" = None
    vals1_i16: "object" = None
    i16_unit: float = 0.0
    i16_total0: int = 0
    i16_total1: int = 0

    @classmethod
    def from_host(cls, p: Pileup, device) -> "DevicePileup":
        import torch
        t = lambda a: torch.from_numpy(a).to(device)
        return cls(t(p.vals0), t(p.off0), t(p.vals1), t(p.off1), t(p.pos), t(p.seg), p.n_pos)


def alloc_device_table(options: DetectOptions, n_pos: int, device) -> Dict[str, "object"]:
    import torch
    tdt = {"int32": torch.int32, "int64": torch.int64, "float64": torch.float64, "uint8": torch.uint8}
    return {c: torch.empty(_col_shape(c, n_pos), dtype=tdt[_lib.TABLE_DTYPES[c]], device=device)
            for c in _wanted_columns(options)}
