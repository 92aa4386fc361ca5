"""Vectorised numpy variant of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see nanomod_oracle.py
for who may import it and for the parity-pinning statement).

Same formulas as ``nanomod_oracle`` (scipy-1.2.1 ks_2samp / mannwhitneyu / ttest_ind /
combine_pvalues, reference call sites bin/scripts/myDetect.py:331,335,341,393,401), evaluated for
many positions at once so that 10^5..10^6-row pileups can be checked in seconds.
tests/test_oracle.py validates it against the scalar oracle row by row.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import scipy.special as sc

from .nanomod_oracle import FLOAT_MAX, FLOAT_MIN, stouffer_weights


def _padded(vals: np.ndarray, off: np.ndarray, rows: np.ndarray, width: int) -> np.ndarray:
    """[len(rows), width] float64, row r = vals[off[r]:off[r+1]] padded with +inf."""
    n = (off[rows + 1] - off[rows]).astype(np.int64)
    out = np.full((rows.shape[0], width), np.inf)
    col = np.arange(width)[None, :]
    mask = col < n[:, None]
    idx = off[rows][:, None] + col
    out[mask] = vals[idx[mask]].astype(np.float64)
    return out


def tests_block(A: np.ndarray, n0: np.ndarray, B: np.ndarray, n1: np.ndarray) -> Dict[str, np.ndarray]:
    """Per-position statistics for a block of rows.  A [R, M0], B [R, M1] are +inf padded."""
    R, M0 = A.shape
    M1 = B.shape[1]
    n0 = n0.astype(np.int64)
    n1 = n1.astype(np.int64)
    pooled = np.concatenate([A, B], axis=1)
    tag = np.concatenate([np.zeros(M0, np.int64), np.ones(M1, np.int64)])
    order = np.argsort(pooled, axis=1, kind="stable")
    sv = np.take_along_axis(pooled, order, axis=1)
    st = tag[order]
    fin = np.isfinite(sv)
    c0 = np.cumsum(st == 0, axis=1)
    c1 = np.cumsum(st == 1, axis=1)
    nxt = np.concatenate([sv[:, 1:], np.full((R, 1), np.inf)], axis=1)
    endg = fin & (nxt > sv)
    d = np.abs(c0 * n1[:, None] - c1 * n0[:, None])
    dnum = np.max(np.where(endg, d, 0), axis=1)
    # ---- KS tail (scipy 1.2.1)
    D = np.max(np.where(endg, np.abs(c0 / (1.0 * n0[:, None]) - c1 / (1.0 * n1[:, None])), 0.0), axis=1)
    en = np.sqrt(n0 * n1 / (n0 + n1).astype(np.float64))
    pks = sc.kolmogorov((en + 0.12 + 0.11 / en) * D)
    # ---- average ranks: group [lo, hi) in 0-based sorted order -> avg rank (lo + 1 + hi) / 2
    M = M0 + M1
    idx = np.broadcast_to(np.arange(M)[None, :], (R, M))
    start = np.concatenate([np.ones((R, 1), bool), sv[:, 1:] != sv[:, :-1]], axis=1)
    lo = np.maximum.accumulate(np.where(start, idx, 0), axis=1)
    end_idx = np.where(endg | ~fin, idx + 1, M + 1)
    hi = np.minimum.accumulate(end_idx[:, ::-1], axis=1)[:, ::-1]
    two_rank = lo + hi + 1
    r2 = np.sum(np.where(fin & (st == 0), two_rank, 0), axis=1)
    t = hi - lo
    tie = np.sum(np.where(fin, t * t - 1, 0), axis=1)
    two_u1 = 2 * n0 * n1 + n0 * (n0 + 1) - r2
    two_u2 = 2 * n0 * n1 - two_u1
    two_u = np.minimum(two_u1, two_u2)
    two_big = np.maximum(two_u1, two_u2)
    n = (n0 + n1).astype(np.float64)
    T = 1.0 - tie / (n ** 3 - n)
    with np.errstate(divide="ignore", invalid="ignore"):
        sd = np.sqrt(T * n0 * n1 * (n + 1) / 12.0)
        z = (0.5 * two_big - (n0 * n1 / 2.0 + 0.5)) / sd
        pu = sc.ndtr(-np.abs(z))
    pu = np.where(T == 0, np.nan, pu)
    # ---- Welch
    fa = np.isfinite(A)
    fb = np.isfinite(B)
    A0 = np.where(fa, A, 0.0)
    B0 = np.where(fb, B, 0.0)
    m0 = A0.sum(axis=1) / n0
    m1 = B0.sum(axis=1) / n1
    v0 = np.where(fa, (A0 - m0[:, None]) ** 2, 0.0).sum(axis=1) / (n0 - 1)
    v1 = np.where(fb, (B0 - m1[:, None]) ** 2, 0.0).sum(axis=1) / (n1 - 1)
    vn0 = v0 / n0
    vn1 = v1 / n1
    with np.errstate(divide="ignore", invalid="ignore"):
        df = (vn0 + vn1) ** 2 / (vn0 ** 2 / (n0 - 1) + vn1 ** 2 / (n1 - 1))
        df = np.where(np.isnan(df), 1, df)
        tt = (m0 - m1) / np.sqrt(vn0 + vn1)
        pt = sc.stdtr(df, -np.abs(tt)) * 2
    clamp_p = lambda p: np.where(p < FLOAT_MIN, FLOAT_MIN, p)
    clamp_s = lambda s: np.where(s > FLOAT_MAX, FLOAT_MAX, s)
    return {"dnum": dnum.astype(np.int64), "D": clamp_s(D), "pks": clamp_p(pks),
            "twoU": two_u.astype(np.int64), "U": clamp_s(0.5 * two_u), "pu": clamp_p(pu),
            "uflag": (T == 0).astype(np.uint8), "t": clamp_s(tt), "pt": clamp_p(pt)}


def combine(ks_p: np.ndarray, pos: np.ndarray, seg: np.ndarray, nb: int, weights_dif: float,
            method: str) -> Dict[str, np.ndarray]:
    """get_combin_pvalue over all rows (myDetect.py:379-404) for method 'fisher'|'stouffer'."""
    n = ks_p.shape[0]
    if nb == 0:
        raise ValueError("nb == 0 returns the KS tuple itself; handle in the caller")
    P = np.ones((n, 2 * nb + 1))
    ii = np.arange(n)
    for k in range(-nb, nb + 1):
        j = ii + k
        okj = (j >= 0) & (j < n)
        jj = np.clip(j, 0, n - 1)
        okk = okj & (seg[jj] == seg) & ((pos[jj].astype(np.int64) - pos.astype(np.int64)) == k)
        P[:, k + nb] = np.where(okk, ks_p[jj], 1.0)
    clamp_p = lambda p: np.where(p < FLOAT_MIN, FLOAT_MIN, p)
    clamp_s = lambda s: np.where(s > FLOAT_MAX, FLOAT_MAX, s)
    if method == "fisher":
        x2 = -2 * np.sum(np.log(P), axis=1)
        return {"stat": clamp_s(x2), "p": clamp_p(sc.chdtrc(2 * (2 * nb + 1), x2))}
    w = np.asarray(stouffer_weights(nb, weights_dif))
    with np.errstate(invalid="ignore"):
        zi = -sc.ndtri(P)
        z = (zi * w[None, :]).sum(axis=1) / np.linalg.norm(w)
    return {"stat": clamp_s(z), "p": clamp_p(sc.ndtr(-z))}


def detect(vals0, off0, vals1, off1, pos, seg, min_coverage: int = 5, nb: int = 2,
           weights_dif: float = 2.0, methods=("stouffer",), chunk: int = 20000,
           rows: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """Filter + tests + combination for a CSR pileup.  ``rows`` (indices into the kept rows)
    restricts the per-position tests to a sample; combination then needs all rows, so it is
    only computed when rows is None."""
    off0 = np.asarray(off0, np.int64)
    off1 = np.asarray(off1, np.int64)
    c0 = np.diff(off0)
    c1 = np.diff(off1)
    keep = np.nonzero((c0 >= min_coverage) & (c1 >= min_coverage))[0]
    sel = keep if rows is None else keep[rows]
    out: Dict[str, list] = {}
    for s in range(0, sel.shape[0], chunk):
        r = sel[s:s + chunk]
        n0 = c0[r]
        n1 = c1[r]
        A = _padded(vals0, off0, r, int(n0.max()))
        B = _padded(vals1, off1, r, int(n1.max()))
        blk = tests_block(A, n0, B, n1)
        blk["n0"] = n0
        blk["n1"] = n1
        for k, v in blk.items():
            out.setdefault(k, []).append(v)
    res = {k: np.concatenate(v) for k, v in out.items()} if out else {}
    res["row_pos_index"] = sel
    if rows is None and sel.shape[0] > 0:
        for m in methods:
            if m == "ks":
                continue
            if nb == 0:
                res[m + "_stat"], res[m + "_p"] = res["D"], res["pks"]
            else:
                c = combine(res["pks"], np.asarray(pos)[sel], np.asarray(seg)[sel], nb, weights_dif, m)
                res[m + "_stat"], res[m + "_p"] = c["stat"], c["p"]
    return res
