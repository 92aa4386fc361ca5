"""GPU parity against the vectors the REFERENCE ITSELF produced (tests/golden/ref_cases.*, see
tests/golden/make_ref_golden.py) and against every SURVEY 8c / tests/golden/golden.json
known-answer case -- all through the C ABI (`nm_detect_host` via `Detector.detect`).

Bar (BASELINE north_star): row set / order, n0, n1, Dnum, 2U bit-exact; every statistic and
p-value within relative 1e-6 (+-inf, NaN and the DBL_MIN / DBL_MAX clamps equal exactly); ranking
identical up to ties of the ranking keys; called-site list identical; table text identical except
where a printed digit sits on a 1e-6 rounding edge.
"""
import json
import os

import numpy as np
import pytest

import nanomod_b200 as nm
from nanomod_b200 import myDetect

import golden_ref as G

pytestmark = pytest.mark.gpu
RTOL = 1e-6
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def det():
    return nm.Detector(0)


@pytest.mark.parametrize("name", G.CASE_NAMES)
def test_cuda_path_reproduces_reference_golden(det, name, tmp_path):
    case, p, opt = G.REF["cases"][name], G.case_pileup(name), G.case_options(name)
    t = det.detect(p, opt)
    rows, want = case["rows"], G.stats_columns(name)
    assert len(t) == len(rows)
    names = t.seg_names
    assert [[names[s][0], names[s][1], int(ps), chr(b), int(a), int(c)] for s, ps, b, a, c in
            zip(t.seg, t.pos, t.base, t.n0, t.n1)] == rows
    assert np.array_equal(t.two_u, np.round(2 * want[:, 0]).astype(np.int64))
    assert np.array_equal(t.ks_dnum.astype(np.int64),
                          np.round(want[:, 4] * t.n0.astype(np.float64) * t.n1).astype(np.int64))
    cols = [("u_stat", 0), ("u_p", 1), ("t_stat", 2), ("t_p", 3), ("ks_d", 4), ("ks_p", 5)]
    if opt.testMethod != "ks":
        cols += [(opt.testMethod + "_stat", 6), (opt.testMethod + "_p", 7)]
    for attr, c in cols:
        got = getattr(t, attr)
        for r in range(len(rows)):
            assert G.same_float(got[r], want[r, c], RTOL), (attr, r, got[r], want[r, c])
    # ranking: the device's radix passes equal the host's lexsort on the same numbers, and with the
    # last-bit noise of both sides rounded away the order is the reference's, row for row
    if opt.RegionRankbyST == 0:
        assert np.array_equal(det.rank(t), t.ranked())
    assert [int(r) for r in G.snapped(t).sorted_rows()] == [int(r) for r in G.snapped(G.table_from_fixture(name)).sorted_rows()]
    # called sites: the reference's own list, except for testMethod='ks', whose discrete p-values tie
    # massively and are ordered in the reference by the last-bit noise of its D -- there the
    # comparison is made after rounding that noise away on both sides
    if opt.testMethod != "ks":
        assert [list(s) for s in t.called_sites()] == case["called_sites"]
    assert G.snapped(t).called_sites() == G.snapped(G.table_from_fixture(name)).called_sites()
    got_lines, want_lines = t.format_lines(), case["sign_test_txt"].splitlines(keepends=True)
    assert len(got_lines) == len(want_lines)
    assert all(same_line(a, b) for a, b in zip(got_lines, want_lines))
    assert sum(a == b for a, b in zip(got_lines, want_lines)) >= 0.98 * len(want_lines)
    if case["meanstd_cvs"] is not None:
        got_m, want_m = t.meanstd_lines(), case["meanstd_cvs"].splitlines(keepends=True)
        assert len(got_m) == len(want_m) and all(same_line(a, b) for a, b in zip(got_m, want_m))


def same_line(a, b):
    """two table lines agree: identical text, or every numeric field within one unit of its last
    printed digit.  (D = Dnum/(n0*n1) is one rounding here and |c0/n0 - c1/n1| -- two divisions
    and a subtraction -- in scipy: at exact decimal ties such as 0.0125 the `%.3f` of the two
    doubles, one ulp apart, prints differently.)"""
    if a == b:
        return True
    fa, fb = a.split(), b.split()
    if len(fa) != len(fb):
        return False
    for x, y in zip(fa, fb):
        if x == y:
            continue
        try:
            vx, vy = float(x), float(y)
        except ValueError:
            return False
        if "E" in x:
            if abs(vx - vy) > 1.001e-3 * 10.0 ** int(y.split("E")[1]):
                return False
        elif abs(vx - vy) > 1.001e-3:
            return False
    return True


def _single_position_pileup(pairs):
    """every (a, b) pair becomes one position of one pileup (positions 0, 2, 4, ...: no neighbours)"""
    v0 = np.concatenate([np.asarray(a, np.float32) for a, _ in pairs])
    v1 = np.concatenate([np.asarray(b, np.float32) for _, b in pairs])
    off0 = np.concatenate([[0], np.cumsum([len(a) for a, _ in pairs])])
    off1 = np.concatenate([[0], np.cumsum([len(b) for _, b in pairs])])
    return nm.Pileup.from_arrays(v0, off0, v1, off1, 2 * np.arange(len(pairs)))


def test_known_answer_vectors_through_the_c_abi(det):
    """SURVEY 8c K1..K4 (as returned by the reference's own getKStest) and the 15 scipy-made cases
    of tests/golden/golden.json, each as one position of a pileup through nm_detect_host."""
    with open(os.path.join(HERE, "golden", "golden.json")) as fh:
        gold = json.load(fh)
    known = G.REF["known"]
    pairs = [(k["a"], k["b"]) for k in known.values()] + [(c["a"], c["b"]) for c in gold["cases"]]
    t = det.detect(_single_position_pileup(pairs), nm.DetectOptions(MinCoverage=3, testMethod="ks", SaveTest=0))
    assert len(t) == len(pairs)
    for r, k in enumerate(known.values()):
        (u, pu), (tt, pt), (d, pks) = k["result"]
        assert t.two_u[r] == round(2 * u) and t.ks_dnum[r] == round(d * len(k["a"]) * len(k["b"]))
        for g, w in ((t.u_stat[r], u), (t.u_p[r], pu), (t.t_stat[r], tt), (t.t_p[r], pt), (t.ks_d[r], d), (t.ks_p[r], pks)):
            assert G.same_float(g, w, RTOL), (r, g, w)
    fmin = np.finfo(np.float64).tiny
    for r, c in enumerate(gold["cases"], start=len(known)):
        assert t.ks_dnum[r] == c["dnum"] and t.two_u[r] == c["twoU"]
        for g, w in ((t.ks_p[r], max(c["pks"], fmin)), (t.u_p[r], max(c["pu"], fmin)), (t.t_stat[r], c["t"]),
                     (t.t_p[r], max(c["pt"], fmin))):
            assert G.same_float(g, w, RTOL), (r, g, w)
    # the reference-seam mirror gives the same numbers for K1..K4
    mo = {"_detector": det}
    for k in known.values():
        got = myDetect.getKStest(mo, k["a"], k["b"], "+")
        for g, w in zip([x for tup in got for x in tup], [x for tup in k["result"] for x in tup]):
            assert G.same_float(g, w, RTOL)


def test_combination_vectors_through_the_c_abi(det):
    """golden.json 'combos' (scipy.stats.combine_pvalues on given p-value windows) and the SURVEY
    8c Stouffer / Fisher vectors: rows are built whose KS p-values reproduce the window only
    approximately, so this checks the combine kernel on the GPU's own p-values against the same
    closed forms the fixtures were made with."""
    import scipy.stats as st
    p = nm.synthetic_pileup(400, 30, 30, round_decimals=3, drop_frac1=0.02)
    for nb, wd in ((2, 2.0), (3, 2.0), (1, 1.5), (5, 3.0)):
        for method in ("stouffer", "fisher"):
            t = det.detect(p, nm.DetectOptions(neighborPvalues=nb, WeightsDif=wd, testMethod=method, SaveTest=0))
            stat, pv = t.comb()
            w = [100.0]
            for _ in range(nb):
                w.insert(0, w[0] / wd)
                w.append(w[-1] / wd)
            for r in range(0, len(t), 7):
                win = []
                for j in range(r - nb, r + nb + 1):
                    okk = 0 <= j < len(t) and t.seg[j] == t.seg[r] and (j - r) == int(t.pos[j]) - int(t.pos[r])
                    win.append(float(t.ks_p[j]) if okk else 1.0)
                with np.errstate(divide="ignore", invalid="ignore"):
                    ws, wp = (st.combine_pvalues(win, method="stouffer", weights=w) if method == "stouffer"
                              else st.combine_pvalues(win))
                assert G.same_float(stat[r], ws, RTOL) or abs(stat[r] - ws) < 1e-9, (nb, method, r, stat[r], ws)
                assert G.same_float(pv[r], max(wp, np.finfo(np.float64).tiny), RTOL), (nb, method, r, pv[r], wp)
