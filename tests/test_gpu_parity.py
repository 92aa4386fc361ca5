"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on
the same float32 inputs.

Parity bar (BASELINE.json north_star / SURVEY.md 8d):
  exact      row set and order, n0, n1, the integer KS numerator Dnum, 2U
  rel 1e-6   every p-value and statistic (KS, U, t, Fisher X^2, Stouffer Z); -inf == -inf;
             values clamped to DBL_MIN / DBL_MAX equal bit for bit
  identical  called-site list
"""
import numpy as np
import pytest

import nanomod_b200 as nm
from nanomod_b200 import myDetect
from nanomod_b200.sharded import ShardedDetector, plan_shards
from oracle import nanomod_oracle as o
from oracle import nanomod_oracle_vec as ov

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # stated tolerance of the north star for p-values and z-scores


@pytest.fixture(scope="module")
def det():
    return nm.Detector(0)


def close(got, want, rtol=RTOL, atol=0.0):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    with np.errstate(invalid="ignore"):
        ok = same | (np.abs(got - want) <= rtol * np.abs(want) + atol)
    return bool(np.all(ok)), int(np.argmin(ok)) if not np.all(ok) else -1


def assert_table_matches(t, res, opt, check_ut=True):
    assert len(t) == len(res["dnum"])
    assert np.array_equal(t.row_pos_index, res["row_pos_index"])
    assert np.array_equal(t.n0, res["n0"]) and np.array_equal(t.n1, res["n1"])
    assert np.array_equal(t.ks_dnum, res["dnum"]), "KS numerator must be bit-exact"
    for name, got, want, atol in (("ks_d", t.ks_d, res["D"], 0), ("ks_p", t.ks_p, res["pks"], 0)):
        ok, i = close(got, want, atol=atol)
        assert ok, (name, i, got[i], want[i])
    if check_ut and opt.want_u:
        assert np.array_equal(t.two_u, res["twoU"]), "2U must be bit-exact"
        assert np.array_equal(t.flags & 1, res["uflag"])
        for name, got, want in (("u_stat", t.u_stat, res["U"]), ("u_p", t.u_p, res["pu"])):
            ok, i = close(got, want)
            assert ok, (name, i, got[i], want[i])
    if check_ut and opt.want_t:
        ok, i = close(t.t_stat, res["t"], atol=1e-12)  # t crosses 0: absolute floor for |t| < 1e-6
        assert ok, ("t", i, t.t_stat[i], res["t"][i])
        ok, i = close(t.t_p, res["pt"])
        assert ok, ("t_p", i, t.t_p[i], res["pt"][i])
    for m in ("stouffer", "fisher"):
        if getattr(t, m + "_p") is not None and (m + "_p") in res:
            gs, gp = getattr(t, m + "_stat"), getattr(t, m + "_p")
            assert np.array_equal(np.isneginf(gs), np.isneginf(res[m + "_stat"]))
            ok, i = close(gs, res[m + "_stat"], atol=1e-9)
            assert ok, (m + "_stat", i, gs[i], res[m + "_stat"][i])
            ok, i = close(gp, res[m + "_p"])
            assert ok, (m + "_p", i, gp[i], res[m + "_p"][i])


def vec(p, opt, methods=("stouffer", "fisher")):
    return ov.detect(p.vals0, p.off0, p.vals1, p.off1, p.pos, p.seg, opt.MinCoverage, opt.neighborPvalues,
                     opt.WeightsDif, methods)


def scalar_moptions(p, opt):
    d0, d1 = p.to_dicts()
    mo = o.default_moptions(MinCoverage=opt.MinCoverage, neighborPvalues=opt.neighborPvalues,
                            WeightsDif=opt.WeightsDif, testMethod=opt.testMethod, rankUse=opt.rankUse,
                            topN=opt.topN, window=opt.half_window)
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    return mo


# ---------------------------------------------------------------------------------------------
# BASELINE config 1: 10 kb, 2x50, KS + weighted Stouffer +-3 -- against the SCALAR oracle
# ---------------------------------------------------------------------------------------------
def test_cfg1_scalar_oracle_full(det):
    p = nm.synthetic_pileup(10000, 50, 50)
    opt = nm.DetectOptions(neighborPvalues=3, WeightsDif=2.0, testMethod="stouffer", topN=30)
    t = det.detect(p, opt)
    mo = scalar_moptions(p, opt)
    st_ = mo["sign_test"]
    assert len(t) == len(st_) == 10000
    for r in (0, 1, 2, 3, 499, 500, 501, 5000, 9996, 9997, 9998, 9999):
        key, tests = st_[r]
        assert (key[2], key[4], key[5]) == (t.pos[r], t.n0[r], t.n1[r])
    col = lambda f: np.array([f(m) for m in st_])
    for name, got, want in (("U", t.u_stat, col(lambda m: m[1][0][0])), ("pU", t.u_p, col(lambda m: m[1][0][1])),
                            ("pt", t.t_p, col(lambda m: m[1][1][1])), ("D", t.ks_d, col(lambda m: m[1][2][0])),
                            ("pks", t.ks_p, col(lambda m: m[1][2][1])), ("Z", t.stouffer_stat, col(lambda m: m[1][3][0])),
                            ("pZ", t.stouffer_p, col(lambda m: m[1][3][1]))):
        ok, i = close(got, want, atol=1e-9 if name == "Z" else 0)
        assert ok, (name, i, got[i], want[i])
    ok, i = close(t.t_stat, col(lambda m: m[1][1][0]), atol=1e-12)
    assert ok
    assert np.array_equal(t.two_u, np.round(2 * col(lambda m: m[1][0][0])).astype(np.int64))
    # called sites and the table text
    assert t.called_sites() == o.called_sites(mo)
    want_lines = o.save_test_lines(mo)
    got_lines = t.format_lines()
    same = sum(a == b for a, b in zip(got_lines, want_lines))
    assert same >= 0.999 * len(want_lines), same  # a %.3E digit may flip on a 1e-13 difference
    assert int(np.sum(np.isneginf(t.stouffer_stat))) == 6  # first/last 3 rows: Z = -inf, p = 1


@pytest.mark.parametrize("variant", ["ties3", "ties1", "gaps", "two_strands_gaps", "poisson", "mincov3", "integers"])
def test_cfg1_variants(det, variant):
    kw = {"ties3": dict(round_decimals=3), "ties1": dict(round_decimals=1), "gaps": dict(drop_frac1=0.01),
          "two_strands_gaps": dict(drop_frac1=0.02, two_strands=True, round_decimals=2),
          "poisson": dict(poisson=True, clip=(2, 128), round_decimals=3),
          "mincov3": dict(poisson=True, clip=(1, 12)), "integers": dict(round_decimals=0)}[variant]
    n = 50 if variant not in ("mincov3",) else 6
    p = nm.synthetic_pileup(10000, n, n, seed=nm.SYN_SEED + 1, **kw)
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True, MinCoverage=3 if variant == "mincov3" else 5)
    t = det.detect(p, opt)
    assert_table_matches(t, vec(p, opt), opt)


@pytest.mark.parametrize("nb,wd,method", [(0, 2.0, "stouffer"), (1, 1.0, "stouffer"), (2, 2.0, "fisher"),
                                          (5, 1.5, "stouffer"), (32, 2.0, "fisher"), (2, 0.3, "stouffer")])
def test_combination_parameters(det, nb, wd, method):
    p = nm.synthetic_pileup(4000, 20, 24, drop_frac1=0.01, two_strands=True)
    opt = nm.DetectOptions(neighborPvalues=nb, WeightsDif=wd, testMethod=method, want_u=False, want_t=False)
    t = det.detect(p, opt)
    opt_eff = nm.DetectOptions(neighborPvalues=nb, WeightsDif=max(wd, 1.0), testMethod=method)
    res = vec(p, opt_eff, (method,))
    assert_table_matches(t, res, opt, check_ut=False)
    assert t.u_p is None and t.t_p is None


def test_ks_only_and_column_subsets(det):
    p = nm.synthetic_pileup(3000, 30, 30, round_decimals=2)
    res = vec(p, nm.DetectOptions(), ())
    for want_u, want_t in ((False, False), (True, False), (False, True)):
        opt = nm.DetectOptions(testMethod="ks", want_u=want_u, want_t=want_t)
        t = det.detect(p, opt)
        assert t.stouffer_p is None and t.fisher_p is None
        assert_table_matches(t, res, opt)


def test_every_coverage_1_to_140(det):
    """Coverage sweep through every network size and across the lane/deep tier boundary."""
    rng = np.random.default_rng(9)
    c0 = np.concatenate([np.arange(1, 141), rng.integers(3, 141, 400)]).astype(np.int64)
    c1 = np.concatenate([np.arange(1, 141)[::-1], rng.integers(3, 141, 400)]).astype(np.int64)
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    v0 = np.round(rng.normal(0, 1, off0[-1]), 2).astype(np.float32)
    v1 = np.round(rng.normal(0.4, 1, off1[-1]), 2).astype(np.float32)
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, np.arange(len(c0), dtype=np.int32))
    opt = nm.DetectOptions(MinCoverage=3, neighborPvalues=2, both_combinations=True)
    assert_table_matches(det.detect(p, opt), vec(p, opt), opt)


def test_deep_rows_mixed_with_lane_rows(det):
    """Deep pileups (block-per-position tier) interleaved with ordinary rows; a +4 sigma deep
    site drives the KS p-value below DBL_MIN so that the clamp (myDetect.py:317-320) is hit."""
    rng = np.random.default_rng(21)
    L = 600
    c0 = np.full(L, 40, np.int64)
    c1 = np.full(L, 40, np.int64)
    deep = {10: (2000, 2000), 11: (2000, 1500), 12: (129, 5), 13: (5, 129), 300: (4000, 3000), 301: (257, 255), 599: (1024, 1024)}
    for i, (a, b) in deep.items():
        c0[i], c1[i] = a, b
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    v0 = np.round(rng.normal(0, 1, off0[-1]), 3).astype(np.float32)
    shift = np.zeros(L)
    shift[10] = 4.0
    shift[300] = 0.1
    v1 = np.round(rng.normal(0, 1, off1[-1]) + np.repeat(shift, c1), 3).astype(np.float32)
    v1[off1[11]:off1[12]] += 50.0  # disjoint deep groups
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, np.arange(L, dtype=np.int32))
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    t = det.detect(p, opt)
    res = vec(p, opt)
    assert_table_matches(t, res, opt)
    assert t.ks_p[11] == o.FLOAT_MIN and t.ks_dnum[11] == 2000 * 1500
    # the scalar oracle agrees on the deep rows too
    for i in deep:
        ref = o.per_position(p.group(0, i).astype(np.float64), p.group(1, i).astype(np.float64))
        assert t.ks_dnum[i] == ref["dnum"] and t.two_u[i] == ref["twoU"]
        assert abs(t.ks_p[i] - ref["pks"]) <= RTOL * ref["pks"]


def test_cfg5_deep_plasmid_slice(det):
    """BASELINE config 5 shape (2x2000x), 300 positions of it, with planted +4 sigma sites."""
    L, n = 300, 2000
    rng = np.random.default_rng(5)
    off = (np.arange(L + 1) * n).astype(np.int64)
    v0 = rng.normal(0, 1, L * n).astype(np.float32)
    shift = np.zeros(L)
    shift[[50, 150, 250]] = 4.0
    v1 = (rng.normal(0, 1, L * n) + np.repeat(shift, n)).astype(np.float32)
    p = nm.Pileup.from_arrays(v0, off, v1, off, np.arange(L, dtype=np.int32))
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer")
    t = det.detect(p, opt)
    assert_table_matches(t, vec(p, opt, ("stouffer",)), opt)
    assert np.all(t.ks_p[[50, 150, 250]] == o.FLOAT_MIN)
    assert abs(t.stouffer_stat[150] - 37.5193793471445 * 100 / np.linalg.norm(o.stouffer_weights(3, 2.0))) < 1.0


def test_degenerate_positions(det):
    rows = [([1.0] * 6, [1.0] * 7), ([1.0] * 5, [2.0] * 5), ([2.0] * 5, [1.0] * 5),
            ([0.0, -0.0, 0.0, 1.0, 2.0], [-0.0, 0.0, 1.0, 2.0, 3.0]), ([1.0, 2.0, 3.0], [1.0, 2.0, 3.0])]
    c0 = np.array([len(a) for a, _ in rows])
    c1 = np.array([len(b) for _, b in rows])
    p = nm.Pileup.from_arrays(np.concatenate([a for a, _ in rows]), np.concatenate([[0], np.cumsum(c0)]),
                              np.concatenate([b for _, b in rows]), np.concatenate([[0], np.cumsum(c1)]),
                              np.arange(len(rows), dtype=np.int32))
    t = det.detect(p, nm.DetectOptions(MinCoverage=3, testMethod="ks"))
    assert t.flags[0] & 1 and np.isnan(t.u_p[0]) and t.ks_dnum[0] == 0 and t.ks_p[0] == 1.0 and t.two_u[0] == 42
    assert np.isnan(t.t_stat[0]) and np.isnan(t.t_p[0])
    assert t.t_stat[1] == -np.inf and t.t_p[1] == o.FLOAT_MIN and t.ks_dnum[1] == 25 and t.two_u[1] == 0
    assert t.t_stat[2] == o.FLOAT_MAX
    ref = o.per_position(np.array(rows[3][0]), np.array(rows[3][1]))
    assert t.ks_dnum[3] == ref["dnum"] and t.two_u[3] == ref["twoU"]
    assert t.ks_dnum[4] == 0 and t.two_u[4] == 9


def test_empty_and_all_filtered(det):
    p = nm.synthetic_pileup(100, 4, 4)
    t = det.detect(p, nm.DetectOptions(MinCoverage=5))
    assert len(t) == 0 and t.called_sites() == []
    empty = nm.Pileup.from_arrays(np.zeros(0, np.float32), np.zeros(1, np.int64), np.zeros(0, np.float32),
                                  np.zeros(1, np.int64), np.zeros(0, np.int32))
    assert len(det.detect(empty, nm.DetectOptions())) == 0


def test_error_codes(det):
    p = nm.synthetic_pileup(64, 8, 8)
    from nanomod_b200 import _lib
    pl = _lib.nm_pileup(p.vals0.ctypes.data, p.off0.ctypes.data, p.vals1.ctypes.data, p.off1.ctypes.data,
                        p.pos.ctypes.data, p.seg.ctypes.data, p.n_pos)
    out = {c: np.empty(64, dtype=_lib.TABLE_DTYPES[c]) for c in ("row_pos_index", "n0", "n1", "ks_dnum", "ks_p")}
    tb = _lib.nm_table(**{c: a.ctypes.data for c, a in out.items()})
    for prm, code in ((_lib.nm_params(2, 2, 2.0, 0, 0, 0, 0), 2), (_lib.nm_params(5, -1, 2.0, 0, 0, 0, 0), 2),
                      (_lib.nm_params(5, 33, 2.0, 2, 0, 0, 0), 2), (_lib.nm_params(5, 2, 2.0, 8, 0, 0, 0), 2),
                      (_lib.nm_params(5, 2, 2.0, 2, 0, 0, 0), 1),   # stouffer outputs missing
                      (_lib.nm_params(5, 2, 2.0, 0, 1, 0, 0), 1)):  # U outputs missing
        with pytest.raises(nm.NmError) as e:
            det.handle.detect_host(pl, prm, tb)
        assert e.value.code == code, (prm.min_coverage, prm.nb, e.value)
    assert det.handle.detect_host(pl, _lib.nm_params(5, 2, 2.0, 0, 0, 0, 0), tb) == 64


def test_rows_beyond_the_shared_memory_deep_tier(det):
    """The reference has no depth limit (getKStest takes any two lists, myDetect.py:327-343): rows
    too long for the deep tier's shared memory go through the global-memory path (nm_huge.cu)
    instead of failing the call -- here next to ordinary and deep rows, with ties."""
    rng = np.random.default_rng(17)
    c0 = np.array([30, 40000, 25, 3000, 70000, 30, 30], np.int64)
    c1 = np.array([28, 30000, 25, 2500, 9000, 31, 20000], np.int64)
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    v0 = np.round(rng.normal(0, 1, off0[-1]), 2).astype(np.float32)
    v1 = np.round(rng.normal(0.02, 1, off1[-1]), 2).astype(np.float32)
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, np.arange(len(c0), dtype=np.int32))
    opt = nm.DetectOptions(neighborPvalues=2, both_combinations=True, mstd=True)
    t = det.detect(p, opt)
    assert len(t) == len(c0)
    for r in range(len(c0)):
        ref = o.per_position(p.group(0, r).astype(np.float64), p.group(1, r).astype(np.float64))
        assert t.ks_dnum[r] == ref["dnum"] and t.two_u[r] == ref["twoU"], r
        for got, want in ((t.ks_p[r], ref["pks"]), (t.u_p[r], ref["pu"]), (t.t_stat[r], ref["t"]), (t.t_p[r], ref["pt"]),
                          (t.ks_d[r], ref["D"])):
            assert got == want or abs(got - want) <= RTOL * abs(want), (r, got, want)
        a = p.group(0, r).astype(np.float64)
        assert abs(t.moments[r, 0] - a.mean()) <= 1e-9 and abs(t.moments[r, 1] - a.var(ddof=1)) <= 1e-9 * a.var(ddof=1)
    res = vec(p, opt)
    ok, i = close(t.stouffer_p, res["stouffer_p"])
    assert ok, i


def test_reference_seam_mirror(det):
    """myDetect.mfilter_coverage / mtest2 / getKStest on the reference's own data model."""
    p = nm.synthetic_pileup(1500, 14, 16, drop_frac1=0.02, two_strands=True, round_decimals=3)
    d0, d1 = p.to_dicts()
    mo = o.default_moptions(neighborPvalues=2, testMethod="stouffer", SaveTest=0, topN=5)
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    mo["_detector"] = det
    ref = scalar_moptions(p, nm.DetectOptions(neighborPvalues=2, topN=5))
    myDetect.mfilter_coverage(mo)
    myDetect.mtest2(mo)
    assert len(mo["sign_test"]) == len(ref["sign_test"])
    for got, want in zip(mo["sign_test"], ref["sign_test"]):
        assert got[0] == want[0]
        for (gs, gp), (ws, wp) in zip(got[1], want[1]):
            assert gs == ws or abs(gs - ws) <= RTOL * abs(ws) + 1e-12
            assert gp == wp or abs(gp - wp) <= RTOL * abs(wp)
    assert [m[0] for m in mo["sorted_sign_test"][:20]] == [m[0] for m in ref["sorted_sign_test"][:20]]
    assert myDetect.called_sites(mo) == o.called_sites(ref)
    a, b = [.1, .2, .2, .3, .3, .3], [.2, .3, .3, .4, .4, .5, .6]  # K2 of SURVEY 8c
    (u, pu), (tt, pt), (d, pks) = myDetect.getKStest(mo, a, b, "+")
    want = o.getKStest(o.default_moptions(), np.float32(a).astype(np.float64), np.float32(b).astype(np.float64), "+")
    for g, w in zip((u, pu, tt, pt, d, pks), [x for pair in want for x in pair]):
        assert abs(g - w) <= RTOL * abs(w)
    assert abs(pks - 0.15504417912365295) < 1e-6 and u == 7.0


def test_mstd_moments_and_meanstd_file(det, tmp_path):
    """--mstd (myDetect.py:437-438, :540-545): per-group mean / std(ddof=0) from the device's
    moments, for lane-tier and deep-tier rows, with and without the t-test, through the seam
    mirror, and the text of <FileID>_meanstd.cvs against the oracle's."""
    rng = np.random.default_rng(5)
    L = 400
    c0 = rng.integers(5, 90, L).astype(np.int64)
    c1 = rng.integers(5, 90, L).astype(np.int64)
    c0[7], c1[7] = 700, 300
    c0[200], c1[200] = 130, 6
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    v0 = np.round(rng.normal(0.3, 1.5, off0[-1]), 3).astype(np.float32)
    v1 = np.round(rng.normal(-0.2, 0.7, off1[-1]), 3).astype(np.float32)
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, np.arange(L, dtype=np.int32))
    for want_t in (True, False):
        opt = nm.DetectOptions(neighborPvalues=2, mstd=True, want_t=want_t, want_u=want_t)
        t = det.detect(p, opt)
        m0, s0, m1, s1 = t.mean_std()
        for r in range(L):
            a = p.group(0, r).astype(np.float64)
            b = p.group(1, r).astype(np.float64)
            for got, want in ((m0[r], np.mean(a)), (s0[r], np.std(a)), (m1[r], np.mean(b)), (s1[r], np.std(b))):
                assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (r, got, want)
        if want_t:
            assert_table_matches(t, vec(p, opt), opt)
    # through the seam mirror, against the scalar oracle's dictionary and file text
    d0, d1 = p.to_dicts()
    mo = o.default_moptions(neighborPvalues=2, mstd=1, SaveTest=1, outFolder=str(tmp_path), FileID="ms")
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    mo["_detector"] = det
    myDetect.mfilter_coverage(mo)
    myDetect.mtest2(mo)
    ref = o.default_moptions(neighborPvalues=2, mstd=1)
    ref["ds2"] = ["g0", "g1"]
    ref["g0"], ref["g1"] = p.to_dicts()
    o.mfilter_coverage(ref)
    o.mtest2(ref, strict=False)
    assert set(mo["sign_test_mstd"]) == set(ref["sign_test_mstd"])
    for k, want in ref["sign_test_mstd"].items():
        got = mo["sign_test_mstd"][k]
        assert np.allclose(np.array(got), np.array(want), rtol=1e-12, atol=1e-12)
    assert open(str(tmp_path) + "/ms_meanstd.cvs").readlines() == o.save_meanstd_lines(ref)


@pytest.mark.parametrize("coverages,times,quantile", [("20-30", 100, 0.25), ("25", 37, 0.5), ("0-12", 64, 0.0)])
def test_downsampling_branch(det, coverages, times, quantile):
    """getKStest's down-sampling branch (myDetect.py:345-361) against the scalar oracle, which
    restates the library's seeded resampling stream: bit-exact KS numerators, p within 1e-6,
    U / t untouched, combination built on the down-sampled p-values; lane- and deep-tier rows."""
    rng = np.random.default_rng(77)
    L = 240
    c0 = rng.integers(6, 70, L).astype(np.int64)
    c1 = rng.integers(6, 70, L).astype(np.int64)
    c0[5], c1[5] = 200, 40
    c0[6], c1[6] = 256, 256
    c0[100], c1[100] = 19, 140
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    shift = np.where(np.arange(L) % 40 == 7, 1.0, 0.0)
    v0 = np.round(rng.normal(0, 1, off0[-1]), 2).astype(np.float32)
    v1 = np.round(rng.normal(0, 1, off1[-1]) + np.repeat(shift, c1), 2).astype(np.float32)
    seg = (np.arange(L) >= L // 2).astype(np.int32)
    pos = np.concatenate([np.arange(L // 2), np.arange(L - L // 2)]).astype(np.int32)
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, pos, seg, seg_names=[("chrA", "+"), ("chrA", "-")])
    opt = nm.DetectOptions(neighborPvalues=2, coverages=coverages, downsampling=times,
                           downsampling_quantile=quantile, seed=424242)
    t = det.detect(p, opt)
    d0, d1 = p.to_dicts()
    mo = o.default_moptions(neighborPvalues=2, coverages=list(opt.coverage_pair()), downsampling=times,
                            downsampling_quantile=quantile, seed=424242)
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    assert len(t) == len(mo["sign_test"]) == L
    n_ds = 0
    for r, (key, tests) in enumerate(mo["sign_test"]):
        cov = opt.coverage_pair()[0 if key[1] == "+" else 1]
        m0 = min(int(t.n0[r]), cov) if cov > 0 else int(t.n0[r])
        m1 = min(int(t.n1[r]), cov) if cov > 0 else int(t.n1[r])
        n_ds += (m0, m1) != (int(t.n0[r]), int(t.n1[r]))
        (u, pu), (tt, pt), (d, pks), (z, pz) = tests
        assert t.ks_dnum[r] == int(round(d * m0 * m1)), (r, key)
        assert abs(t.ks_d[r] - d) <= 1e-12 and abs(t.ks_p[r] - pks) <= RTOL * pks
        assert abs(t.u_p[r] - pu) <= RTOL * pu and abs(t.t_p[r] - pt) <= RTOL * pt
        assert (z == t.stouffer_stat[r]) or abs(t.stouffer_stat[r] - z) <= RTOL * abs(z) + 1e-12
        assert abs(t.stouffer_p[r] - pz) <= RTOL * pz
    assert n_ds > 50
    # same seed, same answer; another seed, another stream
    t2 = det.detect(p, opt)
    assert np.array_equal(t2.ks_dnum, t.ks_dnum)
    opt3 = nm.DetectOptions(neighborPvalues=2, coverages=coverages, downsampling=times,
                            downsampling_quantile=quantile, seed=7)
    assert not np.array_equal(det.detect(p, opt3).ks_dnum, t.ks_dnum)
    # shards see the same stream: it is keyed on (segment, position), not on the row index
    lo = 60
    ts = det.detect(p.slice_rows(lo, L), opt)
    assert np.array_equal(ts.ks_dnum, t.ks_dnum[lo:])


def test_downsampling_limits(det):
    p = nm.synthetic_pileup(50, 300, 20)
    assert len(det.detect(p, nm.DetectOptions(coverages="50"))) == 50  # 300 reads: the block kernel's (round 1: refused)
    with pytest.raises(nm.OptionError):
        det.detect(p, nm.DetectOptions(coverages="50", downsampling=5000))
    # beyond the block kernel: a threshold above 1024 at a position with more than 256 reads -- refused from the plan pass
    q = nm.synthetic_pileup(20, 1500, 20)
    with pytest.raises(nm.NmError) as e:
        det.detect(q, nm.DetectOptions(coverages="1100"))
    assert e.value.code == 5  # NM_ERR_TOO_DEEP
    assert "1024" in str(e.value)


@pytest.mark.parametrize("coverages,times,quantile", [("20-30", 100, 0.25), ("300", 40, 0.5)])
def test_downsampling_branch_deep_positions(det, coverages, times, quantile):
    """Positions to be down-sampled with MORE than 256 reads in a group (nm_downsample_deep_kernel: one CTA per
    position, same seeded stream): bit-exact KS numerators against the scalar oracle, next to positions the
    warp kernel takes and positions that are not down-sampled at all."""
    rng = np.random.default_rng(78)
    L = 48
    c0 = rng.integers(6, 70, L).astype(np.int64)
    c1 = rng.integers(6, 70, L).astype(np.int64)
    deep = {3: (1500, 40), 4: (35, 2100), 9: (3000, 2600), 10: (257, 10), 20: (600, 600), 30: (4097, 280), 31: (200, 256),
            40: (513, 512)}
    for k, (a, b) in deep.items():
        c0[k], c1[k] = a, b
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    shift = np.where(np.arange(L) % 10 == 0, 0.5, 0.0)
    v0 = np.round(rng.normal(0, 1, off0[-1]), 2).astype(np.float32)
    v1 = np.round(rng.normal(0, 1, off1[-1]) + np.repeat(shift, c1), 2).astype(np.float32)
    seg = (np.arange(L) >= L // 2).astype(np.int32)
    pos = np.concatenate([np.arange(L // 2), np.arange(L - L // 2)]).astype(np.int32)
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, pos, seg, seg_names=[("chrA", "+"), ("chrA", "-")])
    opt = nm.DetectOptions(neighborPvalues=2, coverages=coverages, downsampling=times,
                           downsampling_quantile=quantile, seed=99)
    t = det.detect(p, opt)
    d0, d1 = p.to_dicts()
    mo = o.default_moptions(neighborPvalues=2, coverages=list(opt.coverage_pair()), downsampling=times,
                            downsampling_quantile=quantile, seed=99)
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    assert len(t) == len(mo["sign_test"]) == L
    n_deep_ds = 0
    for r, (key, tests) in enumerate(mo["sign_test"]):
        cov = opt.coverage_pair()[0 if key[1] == "+" else 1]
        n0, n1 = int(t.n0[r]), int(t.n1[r])
        m0, m1 = min(n0, cov), min(n1, cov)
        n_deep_ds += (m0, m1) != (n0, n1) and max(n0, n1) > 256
        (u, pu), (tt, pt), (d, pks), (z, pz) = tests
        assert t.ks_dnum[r] == int(round(d * m0 * m1)), (r, key, n0, n1)
        assert abs(t.ks_d[r] - d) <= 1e-12 and abs(t.ks_p[r] - pks) <= RTOL * pks
        assert abs(t.u_p[r] - pu) <= RTOL * pu and abs(t.t_p[r] - pt) <= RTOL * pt
        assert abs(t.stouffer_p[r] - pz) <= RTOL * pz
    assert n_deep_ds >= 6
    assert np.array_equal(det.detect(p, opt).ks_dnum, t.ks_dnum)


@pytest.mark.parametrize("rank_use", ["pv", "st"])
@pytest.mark.parametrize("method", ["stouffer", "fisher", "ks"])
def test_device_ranking_equals_host_ranking(det, rank_use, method):
    """nm_rank_host / nm_rank_device (mtest2 :459-461) against numpy's stable lexsort on the same
    table: heavily tied keys (1-decimal data, coverage 6), NaN U p-values, both directions."""
    import torch
    p = nm.synthetic_pileup(30000, 6, 7, round_decimals=1, drop_frac1=0.01, two_strands=True)
    p.vals0[p.off0[100]:p.off0[101]] = 0.5  # all-identical position: U p-value is NaN
    p.vals1[p.off1[100]:p.off1[101]] = 0.5
    opt = nm.DetectOptions(neighborPvalues=2, testMethod=method, rankUse=rank_use)
    t = det.detect(p, opt)
    want = t.ranked()
    got = det.rank(t)
    assert got.dtype == np.int32 and np.array_equal(got, want)
    # the scalar oracle's sorted list agrees on the leading rows (with testMethod 'ks' the keys
    # are so heavily tied on this data that last-bit differences of U p-values between oracle
    # and GPU decide the order; the combined methods have distinct primary keys)
    if method != "ks":
        ref = scalar_moptions(p, opt)
        st = t.to_sign_test()
        assert [st[int(r)][0] for r in got[:50]] == [m[0] for m in ref["sorted_sign_test"][:50]]
    # device-resident variant
    dev = nm.DevicePileup.from_host(p, "cuda:0")
    out = nm.alloc_device_table(opt, p.n_pos, "cuda:0")
    n_rows = det.detect_device(dev, opt, out)
    order = det.rank_device(out, n_rows, opt)
    torch.cuda.synchronize()
    assert np.array_equal(order.cpu().numpy(), want)
    # without the U column
    opt2 = nm.DetectOptions(neighborPvalues=2, testMethod=method, rankUse=rank_use, want_u=False, want_t=False)
    t2 = det.detect(p, opt2)
    assert np.array_equal(det.rank(t2), t2.ranked())


@pytest.mark.parametrize("want_all", [False, True])
def test_group_binned_launches(det, want_all):
    """Mixed coverage with a dominant size group: the call is split into one launch per group
    (rows partitioned by group, tiles staged as spans or row by row).  Results must not depend
    on it: same table as the oracle's and as the un-binned run of the same library."""
    import os
    rng = np.random.default_rng(123)
    L = 6000
    c0 = rng.integers(20, 40, L).astype(np.int64)
    c1 = rng.integers(20, 40, L).astype(np.int64)
    pick = rng.random(L)
    mid = pick < 0.04
    big = (pick >= 0.04) & (pick < 0.06)
    c0[mid], c1[mid] = rng.integers(70, 104, mid.sum()), rng.integers(5, 104, mid.sum())
    c0[big], c1[big] = rng.integers(105, 129, big.sum()), rng.integers(90, 129, big.sum())
    c0[rng.random(L) < 0.03] = 2          # filtered by MinCoverage
    c0[1000], c1[1000] = 300, 20          # deep rows
    c0[4000:4040], c1[4000:4040] = 100, 100   # a run of mid rows: contiguous tiles inside a group
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    v0 = np.round(rng.normal(0, 1, off0[-1]), 2).astype(np.float32)
    v1 = np.round(rng.normal(0.2, 1, off1[-1]), 2).astype(np.float32)
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, np.arange(L, dtype=np.int32))
    opt = nm.DetectOptions(neighborPvalues=2, both_combinations=want_all, want_u=want_all, want_t=want_all, mstd=want_all)
    before = det.handle.launch_count
    t = det.detect(p, opt)
    launches = det.handle.launch_count - before
    assert launches >= 3 + 3 + 3 + 1  # plan, group keys + partition, three lane launches, combine
    assert_table_matches(t, vec(p, opt), opt)
    os.environ["NANOMOD_B200_NO_CLASS_SORT"] = "1"
    try:
        plain = nm.Detector(0).detect(p, opt)
    finally:
        del os.environ["NANOMOD_B200_NO_CLASS_SORT"]
    for c in ("ks_dnum", "ks_p", "stouffer_p") + (("two_u", "u_p", "t_stat", "t_p", "fisher_p", "moments") if want_all else ()):
        assert np.array_equal(getattr(t, c), getattr(plain, c), equal_nan=True), c


@pytest.mark.parametrize("method", ["stouffer", "fisher"])
def test_pack_records(det, method):
    """nm_pack_records_device: the 28-byte records a rank sends to rank 0."""
    import torch
    from nanomod_b200.detect import RECORD_DTYPE
    p = nm.synthetic_pileup(1000, 12, 15, round_decimals=2)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod=method, want_u=False, want_t=False)
    dev = nm.DevicePileup.from_host(p, "cuda:0")
    out = nm.alloc_device_table(opt, p.n_pos, "cuda:0")
    n_rows = det.detect_device(dev, opt, out)
    lo, n = 3, n_rows - 7
    rec = torch.empty(28 * n, dtype=torch.uint8, device="cuda:0")
    det.pack_records(out, lo, n, opt, rec)
    torch.cuda.synchronize()
    r = rec.cpu().numpy().view(RECORD_DTYPE)
    assert r.shape == (n,)
    assert np.array_equal(r["ks_dnum"], out["ks_dnum"][lo:lo + n].cpu().numpy())
    assert np.array_equal(r["ks_p"], out["ks_p"][lo:lo + n].cpu().numpy())
    assert np.array_equal(r["comb_stat"], out[method + "_stat"][lo:lo + n].cpu().numpy(), equal_nan=True)
    assert np.array_equal(r["comb_p"], out[method + "_p"][lo:lo + n].cpu().numpy())


def test_device_resident_entry_and_unaligned_offsets(det):
    """nm_detect_device on torch tensors; rows whose slices start at odd element offsets."""
    import torch
    p = nm.synthetic_pileup(5000, 50, 50, poisson=True, clip=(5, 128), seed=77)
    opt = nm.DetectOptions(neighborPvalues=3, want_u=True, want_t=True)
    dev = nm.DevicePileup.from_host(p, "cuda:0")
    out = nm.alloc_device_table(opt, p.n_pos, "cuda:0")
    n_rows = det.detect_device(dev, opt, out)
    res = vec(p, opt, ("stouffer",))
    assert n_rows == len(res["dnum"])
    assert np.array_equal(out["ks_dnum"][:n_rows].cpu().numpy(), res["dnum"])
    assert np.array_equal(out["two_u"][:n_rows].cpu().numpy(), res["twoU"])
    ok, i = close(out["stouffer_p"][:n_rows].cpu().numpy(), res["stouffer_p"])
    assert ok
    tm = det.handle.last_timings()
    assert tm["lane"] > 0 and tm["combine"] > 0 and det.launch_count >= 5


def test_sharded_equals_single(det):
    """Halo-recompute sharding (2, 3 and 8 shards on one GPU) is identical to the single run."""
    p = nm.synthetic_pileup(6000, 20, 20, drop_frac1=0.01, two_strands=True, poisson=True, clip=(3, 60))
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    full = det.detect(p, opt)
    sd = ShardedDetector(det)
    for world in (2, 3, 8):
        parts = [sd.detect_range(p, lo, hi, opt) for lo, hi in plan_shards(p.off0, p.off1, world)]
        for name in ("row_pos_index", "ks_dnum", "ks_p", "two_u", "u_p", "t_stat", "t_p", "stouffer_stat",
                     "stouffer_p", "fisher_stat", "fisher_p"):
            cat = np.concatenate([getattr(t, name) for t in parts])
            assert cat.tobytes() == getattr(full, name).tobytes(), (world, name)


# ---------------------------------------------------------------------------------------------
# BASELINE config 2 at full size (4.6 Mb, 2x100x): size-independent properties + sampled oracle
# ---------------------------------------------------------------------------------------------
def test_cfg2_full_size_properties(det):
    import torch
    from bench import make_device_workload
    L, n = 4_600_000, 100
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False)
    dev, host_shift = make_device_workload(L, n, n, torch.device("cuda:0"))
    out = nm.alloc_device_table(opt, L, "cuda:0")
    assert det.detect_device(dev, opt, out) == L
    dnum = out["ks_dnum"].clone()
    pks = out["ks_p"].clone()
    z = out["stouffer_stat"].clone()
    # (1) sampled rows against the vectorised oracle (every planted site +-10 and a 0.5 % sample)
    rng = np.random.default_rng(1)
    planted = np.nonzero(host_shift > 0)[0]
    near = np.unique(np.clip(planted[:, None] + np.arange(-10, 11)[None, :], 0, L - 1))[::7]
    rows = np.unique(np.concatenate([rng.choice(L, 23000, replace=False), near, [0, 1, L - 2, L - 1]]))
    r_t = torch.from_numpy(rows).cuda()
    A = dev.vals0[: L * n].view(L, n)[r_t].double().cpu().numpy()
    B = dev.vals1[: L * n].view(L, n)[r_t].double().cpu().numpy()
    blk = ov.tests_block(A, np.full(len(rows), n), B, np.full(len(rows), n))
    assert np.array_equal(dnum[r_t].cpu().numpy(), blk["dnum"])
    ok, i = close(pks[r_t].cpu().numpy(), blk["pks"])
    assert ok
    # (2) combination of the GPU's own p-values over a window of rows == oracle combine
    lo = 123_000
    seg = np.zeros(4000, np.int32)
    c = ov.combine(pks[lo:lo + 4000].cpu().numpy(), np.arange(lo, lo + 4000, dtype=np.int32), seg, 3, 2.0, "stouffer")
    ok, i = close(z[lo + 3:lo + 3997].cpu().numpy(), c["stat"][3:-3], atol=1e-9)
    assert ok
    assert int(torch.isinf(z).sum().item()) == 6  # only the first/last nb rows of the single run
    # (3) symmetry: swapping the groups leaves Dnum and the p-value unchanged
    swapped = nm.DevicePileup(dev.vals1, dev.off1, dev.vals0, dev.off0, dev.pos, dev.seg, L)
    out2 = nm.alloc_device_table(opt, L, "cuda:0")
    det.detect_device(swapped, opt, out2)
    assert torch.equal(out2["ks_dnum"], dnum) and torch.equal(out2["ks_p"], pks)
    # (4) order statistics only: an exact strictly increasing map (x -> 2x) leaves Dnum unchanged
    mono = nm.DevicePileup(dev.vals0 * 2.0, dev.off0, dev.vals1 * 2.0, dev.off1, dev.pos, dev.seg, L)
    det.detect_device(mono, opt, out2)
    assert torch.equal(out2["ks_dnum"], dnum)
    # (5) idempotence / determinism
    det.detect_device(dev, opt, out2)
    assert torch.equal(out2["ks_dnum"], dnum) and torch.equal(out2["stouffer_stat"].nan_to_num(neginf=-1e300), z.nan_to_num(neginf=-1e300))
    # (6) the planted sites are what gets called
    t_pos = torch.argsort(out["stouffer_p"])[:200].cpu().numpy()
    assert np.mean(host_shift[t_pos] > 0) > 0.95


# ---------------------------------------------------------------------------------------------
# dense path (rows == candidates) vs general path: same tables, and the speculative launch
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def det_general():
    import os
    os.environ["NANOMOD_B200_NO_DENSE"] = "1"
    try:
        d = nm.Detector(0)
    finally:
        del os.environ["NANOMOD_B200_NO_DENSE"]
    return d


def _tables_identical(a, b):
    assert len(a) == len(b)
    for name in ("row_pos_index", "n0", "n1", "ks_dnum", "two_u", "flags"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for name in ("ks_d", "ks_p", "u_stat", "u_p", "t_stat", "t_p", "fisher_stat", "fisher_p", "stouffer_stat", "stouffer_p"):
        x, y = getattr(a, name), getattr(b, name)
        assert (x is None) == (y is None), name
        if x is not None:
            assert np.array_equal(x, y, equal_nan=True), name  # bit for bit: same arithmetic on both paths


@pytest.mark.parametrize("n", [9, 50, 64, 65, 100, 104, 128])
def test_dense_path_equals_general_path(det, det_general, n):
    p = nm.synthetic_pileup(3000 + n, n, max(5, n - 3), round_decimals=2)
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    t = det.detect(p, opt)
    assert det.handle.last_path() in (1, 2, 3)  # dense: directly, speculatively, or after a refused guess
    g = det_general.detect(p, opt)
    assert det_general.handle.last_path() == 0
    _tables_identical(t, g)
    assert_table_matches(t, vec(p, opt), opt)
    ks = nm.DetectOptions(neighborPvalues=2, testMethod="fisher", want_u=False, want_t=False)
    _tables_identical(det.detect(p, ks), det_general.detect(p, ks))
    nb0 = nm.DetectOptions(neighborPvalues=0, testMethod="stouffer")
    _tables_identical(det.detect(p, nb0), det_general.detect(p, nb0))


def test_speculative_dense_launch_and_refusal(det):
    """After a dense-shaped call the next one is launched without waiting for its plan summary;
    when its shape differs the device refuses and the call is re-run -- results are right either way."""
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer")
    a = nm.synthetic_pileup(2000, 40, 40)
    b = nm.synthetic_pileup(2000, 40, 40, seed=7)
    c = nm.synthetic_pileup(2000, 100, 90)                      # another network class
    d = nm.synthetic_pileup(2000, 40, 40, drop_frac1=0.05)      # filtered rows: not dense at all
    e2 = nm.synthetic_pileup(400, 40, 40)                       # dense again
    fresh = nm.Detector(0)
    paths = []
    for p in (a, b, c, d, e2, a):
        t = fresh.detect(p, opt)
        paths.append(fresh.handle.last_path())
        assert_table_matches(t, vec(p, opt), opt)
    assert paths == [1, 2, 3, 4, 1, 2], paths


def test_bad_offsets_and_segment_ids_are_rejected(det):
    p = nm.synthetic_pileup(500, 20, 20)
    off = p.off0.copy()
    off[100] = off[101] + 5  # not monotonic
    bad = nm.Pileup(vals0=p.vals0, off0=off, vals1=p.vals1, off1=p.off1, pos=p.pos, seg=p.seg, base=p.base,
                    seg_names=p.seg_names)
    with pytest.raises(nm.NmError) as e:
        det.detect(bad, nm.DetectOptions())
    assert e.value.code == 1
    t = det.detect(p, nm.DetectOptions())  # the handle is still usable
    assert len(t) == 500


# ---------------------------------------------------------------------------------------------
# head of the ranking without a full sort (what the called-site rule and a sharded run need)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rank_use,method,want", [("pv", "stouffer", 50), ("st", "stouffer", 50), ("pv", "ks", 300),
                                                   ("st", "fisher", 5000), ("pv", "fisher", 1)])
def test_ranking_head_is_a_prefix_of_the_full_ranking(det, rank_use, method, want):
    import torch
    p = nm.synthetic_pileup(40000, 12, 14, round_decimals=1)  # heavy ties: KS D takes few values
    opt = nm.DetectOptions(neighborPvalues=2, testMethod=method, rankUse=rank_use)
    dev = nm.DevicePileup.from_host(p, "cuda:0")
    out = nm.alloc_device_table(opt, p.n_pos, "cuda:0")
    n_rows = det.detect_device(dev, opt, out)
    full = det.rank_device(out, n_rows, opt).cpu().numpy()
    head = det.rank_head_device(out, n_rows, opt, want)
    assert len(head) >= min(want, n_rows)
    assert np.array_equal(head["row"], full[:len(head)].astype(np.int64))
    # with geometry: segment / position of each row and plot1's neighbourhood test (rows are candidates here
    # unless something was filtered; the flag is checked against the host implementation)
    rpi = out["row_pos_index"][:n_rows]
    nearby = 10
    hg = det.rank_head_device(out, n_rows, opt, want, geometry=(rpi, dev.pos, dev.seg, 0, n_rows, nearby))
    assert np.array_equal(hg["row"], head["row"])
    idx = rpi.cpu().numpy()[hg["row"]]
    assert np.array_equal(hg["pos"], p.pos[idx]) and np.array_equal(hg["seg"], p.seg[idx])
    from nanomod_b200.sharded import neighbourhood_flags
    want_flags = neighbourhood_flags(p.seg[rpi.cpu().numpy()], p.pos[rpi.cpu().numpy()], hg["row"], nearby)
    assert np.array_equal(hg["full_nbhd"] != 0, want_flags)
    # asking for everything returns the whole ranking
    if want == 5000:
        everything = det.rank_head_device(out, n_rows, opt, n_rows)
        assert np.array_equal(everything["row"], full.astype(np.int64))


# ---------------------------------------------------------------------------------------------
# sharded, device-resident table: per-shard heads merged == the single table's called sites; text
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method,rank_use", [("stouffer", "pv"), ("fisher", "st"), ("ks", "pv")])
def test_sharded_device_heads_and_text(det, method, rank_use, tmp_path):
    from nanomod_b200.sharded import greedy_sites, merge_heads, shard_halo, shard_with_halo
    p = nm.synthetic_pileup(30000, 14, 12, drop_frac1=0.004, two_strands=True, round_decimals=2)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod=method, rankUse=rank_use, topN=15, SaveTest=0)
    full = det.detect(p, opt)
    sd = ShardedDetector(det)
    world = 5
    heads, texts = [], []
    for lo, hi in plan_shards(p.off0, p.off1, world):
        sl, core_lo, core_hi = shard_with_halo(p, lo, hi, shard_halo(opt))
        res = sd.detect_shard(nm.DevicePileup.from_host(sl, "cuda:0"), core_lo, core_hi, lo - core_lo, opt)
        heads.append(sd.local_head(res, 200))
        path = str(tmp_path / ("part_%d.txt" % lo))
        sd.save_test(res, p.seg_names, sl.base, path)
        texts.append(open(path, "rb").read())
    assert sum(h.n_core for h in heads) == len(full)
    m = merge_heads(heads, rank_use != "pv")
    assert np.array_equal(m.row[:m.n_exact], det.rank(full)[:m.n_exact])
    sites, final = greedy_sites(m, opt, p.seg_names)
    assert final and sites == full.called_sites()
    assert b"".join(texts) == full.format_text()
    # one "rank" holding everything: the public entry point
    res = sd.detect_shard(nm.DevicePileup.from_host(p, "cuda:0"), 0, p.n_pos, 0, opt)
    assert sd.called_sites(res, p.seg_names) == full.called_sites()


@pytest.mark.parametrize("drop", [0.0, 0.01])
def test_head_selection_armed_for_the_detect_call(det, drop):
    """detect_shard(head_want=...) arms the head selection so that the detect call launches it behind its own
    kernels, before its host wait: same records as selecting afterwards.  With filtered rows the armed range does
    not apply, nothing fires, and gather_heads selects as before."""
    import torch
    from nanomod_b200.sharded import HEAD_REC, shard_halo, shard_with_halo
    p = nm.synthetic_pileup(20000, 70, 66, drop_frac1=drop, round_decimals=3, seed=12)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
    sd = ShardedDetector(det)
    lo, hi = 4000, 16000
    sl, core_lo, core_hi = shard_with_halo(p, lo, hi, shard_halo(opt))
    dev = nm.DevicePileup.from_host(sl, "cuda:0")
    for _ in range(2):  # second round: the speculative dense launch
        plain = sd.detect_shard(dev, core_lo, core_hi, lo - core_lo, opt)
        assert plain.head_slot is None
        want_rec = sd.gather_heads(plain, 300, cap=1024).cpu().numpy().view(HEAD_REC).copy()
        armed = sd.detect_shard(dev, core_lo, core_hi, lo - core_lo, opt, head_want=300, head_cap=1024)
        assert (armed.head_slot is not None) == (drop == 0.0)
        got_rec = sd.gather_heads(armed, 300, cap=1024).cpu().numpy().view(HEAD_REC).copy()
        n = int(want_rec[0]["row"])
        assert n >= 300 and int(got_rec[0]["row"]) == n
        a = np.sort(want_rec[1:n + 1], order=["row"])
        b = np.sort(got_rec[1:n + 1], order=["row"])
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("kind", ["null", "identical_groups", "all_shifted"])
def test_armed_head_from_the_combine_kernels_candidate_list(det, kind):
    """With p-value ranking the armed selection works from the candidate list the combine kernel leaves (rows
    whose combined p lies below a power of two chosen from want / n) instead of three passes over all rows:
    same header and records as selecting afterwards.  When the list cannot give the head (identical groups:
    every p is 1, no candidates) the call reports the selection as not run and the caller selects as before."""
    from nanomod_b200.sharded import HEAD_REC
    p = nm.synthetic_pileup(30000, 70, 66, round_decimals=3, seed=31)
    if kind == "identical_groups":
        p = nm.Pileup(vals0=p.vals0, off0=p.off0, vals1=p.vals0.copy(), off1=p.off0.copy(), pos=p.pos, seg=p.seg,
                      base=p.base, seg_names=p.seg_names)
    elif kind == "all_shifted":
        p.vals1[:] = np.round(p.vals1 + 1.5, 3)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
    sd = ShardedDetector(det)
    dev = nm.DevicePileup.from_host(p, "cuda:0")
    lo, hi = 10, 29990
    for _ in range(2):
        plain = sd.detect_shard(dev, lo, hi, 0, opt)
        want_rec = sd.gather_heads(plain, 400, cap=2048).cpu().numpy().view(HEAD_REC).copy()
        armed = sd.detect_shard(dev, lo, hi, 0, opt, head_want=400, head_cap=2048)
        assert (armed.head_slot is not None) == (kind != "identical_groups")
        got_rec = sd.gather_heads(armed, 400, cap=2048).cpu().numpy().view(HEAD_REC).copy()
        n = int(want_rec[0]["row"])
        assert int(got_rec[0]["row"]) == n and got_rec[0]["key"].tobytes() == want_rec[0]["key"].tobytes()
        assert np.sort(want_rec[1:n + 1], order=["row"]).tobytes() == np.sort(got_rec[1:n + 1], order=["row"]).tobytes()


def _table_bytes(out, n):
    return {c: out[c][:n].cpu().numpy().tobytes() for c in out if out[c].dim() == 1}


def test_queued_calls_equal_synchronous_calls(det):
    """nm_detect_device_async / nm_detect_finish: two calls in flight on the assumed (dense) shape give the
    tables of the synchronous call; a call whose shape is not the previous one's is refused on the device and
    re-run by finish; calls that are not dense at all run synchronously behind the same interface."""
    import torch
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
    pa = nm.synthetic_pileup(30000, 100, 100, round_decimals=3, seed=21)
    pb = nm.synthetic_pileup(30000, 100, 100, round_decimals=3, seed=22)
    pc = nm.synthetic_pileup(30000, 40, 44, round_decimals=3, seed=23)            # dense, another network class
    pd = nm.synthetic_pileup(30000, 100, 100, drop_frac1=0.01, round_decimals=3, seed=24)  # filtered rows: general path
    devs = [nm.DevicePileup.from_host(p, "cuda:0") for p in (pa, pb, pc, pd)]
    want = []
    for d in devs:
        o = nm.alloc_device_table(opt, d.n_pos, "cuda:0")
        n = det.detect_device(d, opt, o)
        want.append((n, _table_bytes(o, n)))
    det.detect_device(devs[0], opt, nm.alloc_device_table(opt, devs[0].n_pos, "cuda:0"))  # establishes the dense shape
    outs = [nm.alloc_device_table(opt, 30000, "cuda:0") for _ in range(2)]
    order = [0, 1, 0, 2, 2, 0, 3, 1, 1, 0]
    paths = []
    pend = []
    for k, i in enumerate(order):
        pend.append((i, k & 1, det.detect_device_async(devs[i], opt, outs[k & 1])))
        if len(pend) > 1:
            i0, b0, t0 = pend.pop(0)
            n, _ = det.detect_finish(t0)
            paths.append(det.handle.last_path())
            assert n == want[i0][0]
            assert _table_bytes(outs[b0], n) == want[i0][1], "queued call %d differs from the synchronous call" % i0
    i0, b0, t0 = pend.pop(0)
    n, _ = det.detect_finish(t0)
    assert n == want[i0][0] and _table_bytes(outs[b0], n) == want[i0][1]
    assert 2 in paths, "no call was launched on the assumed shape"
    # a third call in flight, and two calls writing one table, are refused
    t1 = det.detect_device_async(devs[0], opt, outs[0])
    with pytest.raises(nm.NmError):
        det.detect_device_async(devs[1], opt, outs[0])
    t2 = det.detect_device_async(devs[1], opt, outs[1])
    with pytest.raises(nm.NmError):
        det.detect_device_async(devs[0], opt, nm.alloc_device_table(opt, 30000, "cuda:0"))
    assert det.detect_finish(t1)[0] == 30000 and det.detect_finish(t2)[0] == 30000
    with pytest.raises(nm.NmError):
        det.detect_finish(t1)


@pytest.mark.parametrize("drop", [0.0, 0.01])
def test_head_selection_stores_into_peer_buffers(det, drop):
    """nm_head_set_peers: the selection kernels store the header (with the epoch) and every record into each
    peer section as well -- here two sections in this GPU's own memory stand in for two ranks' buffers.  Through
    ShardedDetector (world 1): the armed, queued and after-the-call selections all land in the exchange buffer."""
    import torch
    from nanomod_b200.sharded import HEAD_REC, shard_halo, shard_with_halo
    p = nm.synthetic_pileup(20000, 70, 66, drop_frac1=drop, round_decimals=3, seed=12)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
    sd = ShardedDetector(det)
    lo, hi = 4000, 16000
    sl, core_lo, core_hi = shard_with_halo(p, lo, hi, shard_halo(opt))
    dev = nm.DevicePileup.from_host(sl, "cuda:0")
    cap, want = 1024, 300
    plain = sd.detect_shard(dev, core_lo, core_hi, lo - core_lo, opt)
    ref = sd.gather_heads(plain, want, cap=cap).cpu().numpy().view(HEAD_REC).copy()
    n = int(ref[0]["row"])
    assert n >= want

    def same(buf, epoch):
        b = buf.cpu().numpy().view(HEAD_REC)
        assert int(b[0]["row"]) == n and int(b[0]["reserved"]) == epoch
        assert b[0]["key"].tobytes() == ref[0]["key"].tobytes()
        assert np.sort(b[1:n + 1], order=["row"]).tobytes() == np.sort(ref[1:n + 1], order=["row"]).tobytes()

    # C ABI: two peer sections
    peers = [torch.zeros((cap + 1) * HEAD_REC.itemsize, dtype=torch.uint8, device="cuda:0") for _ in range(2)]
    mine = torch.zeros((cap + 1) * HEAD_REC.itemsize, dtype=torch.uint8, device="cuda:0")
    det.handle.head_set_peers([t.data_ptr() for t in peers], 7)
    core = {c: plain.out[c][plain.r_lo:plain.r_hi] for c in ("ks_p", "stouffer_p")}
    rpi = None if plain.n_rows == dev.n_pos else plain.out["row_pos_index"]
    from nanomod_b200.sharded import nearby_rows
    det.rank_head_select_device(core, plain.n_core, opt, want, mine, cap,
                                geometry=(rpi, dev.pos, dev.seg, plain.r_lo, plain.n_rows, nearby_rows(opt)))
    torch.cuda.synchronize()
    same(mine, 7)
    for t in peers:
        same(t, 7)
    # one shot: the next selection stores nowhere else
    for t in peers:
        t.zero_()
    det.rank_head_select_device(core, plain.n_core, opt, want, mine, cap,
                                geometry=(rpi, dev.pos, dev.seg, plain.r_lo, plain.n_rows, nearby_rows(opt)))
    torch.cuda.synchronize()
    assert all(int(t.sum().item()) == 0 for t in peers)
    # product path at world size 1
    res = sd.detect_shard(dev, core_lo, core_hi, lo - core_lo, opt, head_want=want, head_cap=cap, slot=0, peer=True)
    assert (res.head_slot is not None) == (drop == 0.0) and res.head_epoch > 0
    same(sd.gather_heads(res, want, cap=cap, slot=0, peer=True), res.head_epoch)
    outs = [nm.alloc_device_table(opt, dev.n_pos, "cuda:0") for _ in range(2)]
    pend = [sd.detect_shard_async(dev, core_lo, core_hi, lo - core_lo, opt, outs[k], head_want=want, head_cap=cap, slot=k,
                                  peer=True) for k in range(2)]
    for k in range(2):
        r = sd.finish_shard(pend[k])
        assert r.n_rows == plain.n_rows
        same(sd.gather_heads(r, want, cap=cap, slot=k, peer=True), r.head_epoch)
    # another capacity: the old exchange buffer is closed, a new one built
    x_old = sd.peer_exchange(cap, dev.pos.device)
    x_new = sd.peer_exchange(2 * cap, dev.pos.device)
    assert x_old.ptr is None and x_new.ptr is not None and x_new.cap == 2 * cap
    x_new.close()


def test_peer_buffer_alloc_and_view(det):
    import torch
    from nanomod_b200.sharded import _DevView
    ptr, ipc = det.handle.peer_alloc(4096)
    assert ptr != 0 and len(ipc) == 64
    t = torch.as_tensor(_DevView(ptr, 4096), device="cuda:0")
    assert int(t.sum().item()) == 0
    t[:16] = 3
    assert int(torch.as_tensor(_DevView(ptr, 4096), device="cuda:0").sum().item()) == 48
    del t
    det.handle.peer_free(ptr)


# ---------------------------------------------------------------------------------------------
# pipelined host entry (slabs with halos, copies overlapped with compute) == the one-piece call
# ---------------------------------------------------------------------------------------------
def _detector_with_env(**env):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return nm.Detector(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize("slab", [257, 1000, 4096])
def test_pipelined_host_entry_equals_one_piece(slab):
    whole = _detector_with_env(NANOMOD_B200_SLAB=0)
    piped = _detector_with_env(NANOMOD_B200_SLAB=slab)
    cases = [
        (nm.synthetic_pileup(9000, 30, 28, round_decimals=2), nm.DetectOptions(neighborPvalues=3, both_combinations=True)),
        (nm.synthetic_pileup(9000, 12, 14, drop_frac1=0.03, two_strands=True, poisson=True, clip=(2, 60), round_decimals=3),
         nm.DetectOptions(neighborPvalues=5, testMethod="fisher", mstd=True)),
        (nm.synthetic_pileup(9000, 20, 20, drop_frac1=0.5), nm.DetectOptions(neighborPvalues=2, testMethod="ks", want_u=False, want_t=False)),
        (nm.synthetic_pileup(9000, 40, 40), nm.DetectOptions(neighborPvalues=0, testMethod="stouffer")),
    ]
    for p, opt in cases:
        a, b = whole.detect(p, opt), piped.detect(p, opt)
        _tables_identical(a, b)
        if opt.mstd:
            assert np.array_equal(a.moments, b.moments)
    # deep rows and the down-sampling branch inside slabs
    rng = np.random.default_rng(3)
    L = 3000
    c0 = np.full(L, 20, np.int64)
    c1 = np.full(L, 24, np.int64)
    for i in (5, 256, 257, 1000, 1001, 2999):
        c0[i], c1[i] = 700, 300
    c0[1500], c1[1500] = 150, 180
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    p = nm.Pileup.from_arrays(np.round(rng.normal(0, 1, off0[-1]), 3), off0, np.round(rng.normal(0.1, 1, off1[-1]), 3), off1,
                              np.arange(L, dtype=np.int32))
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    _tables_identical(whole.detect(p, opt), piped.detect(p, opt))
    keep = np.ones(L, bool)
    keep[[5, 256, 257, 1000, 1001, 2999]] = False
    q = nm.Pileup.from_arrays(p.vals0[:off0[-1]][np.repeat(keep, c0)], np.concatenate([[0], np.cumsum(c0[keep])]),
                              p.vals1[:off1[-1]][np.repeat(keep, c1)], np.concatenate([[0], np.cumsum(c1[keep])]),
                              np.arange(L, dtype=np.int32)[keep])
    ds = nm.DetectOptions(neighborPvalues=2, testMethod="stouffer", coverages="100-100", downsampling=64)
    _tables_identical(whole.detect(q, ds), piped.detect(q, ds))


# ---------------------------------------------------------------------------------------------
# 16-bit transport format: half the bytes in, bit-identical tables out
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("slab", [0, 1500])
def test_int16_transport_format(slab):
    import torch
    det = _detector_with_env(NANOMOD_B200_SLAB=slab)
    p = nm.synthetic_pileup(7000, 35, 31, round_decimals=3, drop_frac1=0.01, two_strands=True)  # the reference's 0.001 grid
    q = p.to_int16(0.001)
    assert q.vals0_i16.dtype == np.int16 and q.vals0_i16.nbytes < p.vals0.nbytes * 0.52
    for opt in (nm.DetectOptions(neighborPvalues=3, both_combinations=True, mstd=True),
                nm.DetectOptions(neighborPvalues=2, testMethod="ks", want_u=False, want_t=False)):
        a, b = det.detect(p, opt), det.detect(q, opt)
        _tables_identical(a, b)
        if opt.mstd:
            assert np.array_equal(a.moments, b.moments)
    # device-resident int16 pileup
    dev = nm.DevicePileup.from_host(p, "cuda:0")
    dev16 = nm.DevicePileup(None, dev.off0, None, dev.off1, dev.pos, dev.seg, p.n_pos,
                            vals0_i16=torch.from_numpy(q.vals0_i16).cuda(), vals1_i16=torch.from_numpy(q.vals1_i16).cuda(),
                            i16_unit=0.001, i16_total0=int(p.off0[-1]), i16_total1=int(p.off1[-1]))
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer")
    o1, o2 = nm.alloc_device_table(opt, p.n_pos, "cuda:0"), nm.alloc_device_table(opt, p.n_pos, "cuda:0")
    n1, n2 = det.detect_device(dev, opt, o1), det.detect_device(dev16, opt, o2)
    assert n1 == n2
    for c in ("ks_dnum", "ks_p", "two_u", "t_stat", "stouffer_stat", "stouffer_p"):
        assert torch.equal(o1[c][:n1].view(torch.uint8), o2[c][:n2].view(torch.uint8)), c
    # values off the grid have no exact int16 form
    with pytest.raises(ValueError):
        nm.synthetic_pileup(100, 10, 10).to_int16(0.001)
