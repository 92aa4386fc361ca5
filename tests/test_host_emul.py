"""The product's host/device headers (nm_math.cuh, nm_lane.cuh, nm_deep.cuh), compiled for the
CPU by tests/host_emul, against the oracle: lane tier (sorting network + merge walk), deep tier
(rank counts), window combination and the fp64 tails.  No GPU involved; the GPU tests repeat the
same comparisons through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest
import scipy.special as sc

from oracle import nanomod_oracle as o
from oracle import nanomod_oracle_vec as ov
from conftest import RowOut, rel_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
FP = ctypes.POINTER(ctypes.c_float)


def run(fn, a, b, want_u=1, want_t=1):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    r = RowOut()
    rc = fn(a.ctypes.data_as(FP), len(a), b.ctypes.data_as(FP), len(b), want_u, want_t, ctypes.byref(r))
    assert rc == 0
    return r


def check(r, ref, tol=1e-9):
    assert r.dnum == ref["dnum"] and r.two_u == ref["twoU"]
    assert r.flags == ref["uflag"]
    for got, want in ((r.ks_p, ref["pks"]), (r.ks_d, ref["D"]), (r.u_p, ref["pu"]), (r.t_p, ref["pt"])):
        assert rel_err(got, want) < tol, (got, want)
    assert rel_err(r.t_stat, ref["t"]) < tol or abs(r.t_stat - ref["t"]) < 1e-13


@pytest.mark.parametrize("tier", ["lane", "deep"])
def test_tiers_random_ties_unequal_n(emul, tier):
    fn = emul.emul_lane_position if tier == "lane" else emul.emul_deep_position
    rng = np.random.default_rng(11)
    for trial in range(600):
        n0, n1 = int(rng.integers(3, 129)), int(rng.integers(3, 129))
        a = rng.normal(0, 1, n0)
        b = rng.normal(rng.choice([0, 0.5, 1, 3]), 1, n1)
        dec = [None, 1, 3, 0][trial % 4]
        if dec is not None:
            a, b = np.round(a, dec), np.round(b, dec)
        a, b = a.astype(np.float32), b.astype(np.float32)
        check(run(fn, a, b), o.per_position(a.astype(np.float64), b.astype(np.float64)))


def test_lane_every_network_size_boundary(emul):
    rng = np.random.default_rng(5)
    for n in list(range(3, 18)) + [23, 24, 25, 31, 32, 33, 63, 64, 65, 96, 97, 100, 104, 120, 121, 127, 128]:
        for m in (3, n):
            a = np.round(rng.normal(0, 1, n), 2).astype(np.float32)
            b = np.round(rng.normal(0.3, 1, m), 2).astype(np.float32)
            check(run(emul.emul_lane_position, a, b), o.per_position(a.astype(np.float64), b.astype(np.float64)))


def test_golden_cases(emul):
    for c in GOLD["cases"]:
        a, b = np.asarray(c["a"], np.float32), np.asarray(c["b"], np.float32)
        fns = [emul.emul_deep_position]
        if max(len(a), len(b)) <= 128:
            fns.append(emul.emul_lane_position)
        for fn in fns:
            r = run(fn, a, b)
            assert r.dnum == c["dnum"] and r.two_u == c["twoU"]
            assert rel_err(r.ks_p, max(c["pks"], o.FLOAT_MIN)) < 1e-9
            assert rel_err(r.u_p, max(c["pu"], o.FLOAT_MIN)) < 1e-9 and rel_err(r.t_stat, c["t"]) < 1e-9
            assert rel_err(r.t_p, max(c["pt"], o.FLOAT_MIN)) < 1e-9


def test_degenerate_positions(emul):
    for fn in (emul.emul_lane_position, emul.emul_deep_position):
        r = run(fn, [1.0] * 6, [1.0] * 7)  # all identical: reference raises; we flag
        assert r.flags == 1 and np.isnan(r.u_p) and r.dnum == 0 and r.ks_p == 1.0 and r.two_u == 42
        assert np.isnan(r.t_stat) and np.isnan(r.t_p)
        r = run(fn, [1.0] * 5, [2.0] * 5)  # zero variance, different means: t = -inf, p = 0 -> clamp
        assert r.t_stat == -np.inf and r.t_p == o.FLOAT_MIN and r.dnum == 25 and r.two_u == 0
        r = run(fn, [2.0] * 5, [1.0] * 5)  # +inf statistic is clamped to DBL_MAX (m_max_float)
        assert r.t_stat == o.FLOAT_MAX
        r = run(fn, [0.0, -0.0, 0.0, 1.0, 2.0], [-0.0, 0.0, 1.0, 2.0, 3.0])  # -0.0 ties 0.0
        check(r, o.per_position(np.array([0.0, -0.0, 0.0, 1.0, 2.0]), np.array([-0.0, 0.0, 1.0, 2.0, 3.0])))


def test_deep_clamped_pvalue(emul):
    a = (np.arange(2000) / 2000.0).astype(np.float32)
    b = (10 + np.arange(2000) / 2000.0).astype(np.float32)
    r = run(emul.emul_deep_position, a, b)
    assert r.dnum == 2000 * 2000 and r.ks_p == o.FLOAT_MIN and r.two_u == 0
    ref = o.per_position(a.astype(np.float64), b.astype(np.float64))
    assert rel_err(r.u_p, ref["pu"]) < 1e-9 and rel_err(r.t_p, ref["pt"]) < 1e-9


def test_combine_window(emul):
    import nanomod_b200 as nm
    p = nm.synthetic_pileup(1500, 20, 20, drop_frac1=0.03, two_strands=True)
    res = ov.detect(p.vals0, p.off0, p.vals1, p.off1, p.pos, p.seg, 5, 3, 2.0, ("stouffer", "fisher"))
    n = len(res["pks"])
    pos = np.ascontiguousarray(p.pos[res["row_pos_index"]], np.int32)
    seg = np.ascontiguousarray(p.seg[res["row_pos_index"]], np.int32)
    pks = np.ascontiguousarray(res["pks"])
    pks[::97] = o.FLOAT_MIN  # exercise the clamp inside the window
    for nb, wd in ((3, 2.0), (2, 2.0), (1, 1.0), (5, 1.5)):
        want_s = ov.combine(pks, pos, seg, nb, wd, "stouffer")
        want_f = ov.combine(pks, pos, seg, nb, wd, "fisher")
        outs = [np.zeros(n) for _ in range(4)]
        dp = ctypes.POINTER(ctypes.c_double)
        emul.emul_combine(pks.ctypes.data_as(dp), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                          seg.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_int64(n), nb,
                          ctypes.c_double(wd), 1, 1, *[x.ctypes.data_as(dp) for x in outs])
        fs, fpv, ss, sp = outs
        assert np.array_equal(np.isneginf(ss), np.isneginf(want_s["stat"]))
        fin = np.isfinite(want_s["stat"])
        # host build uses a Newton inverse-normal (libm has none); the device uses normcdfinv
        assert np.max(np.abs(ss[fin] - want_s["stat"][fin]) / np.maximum(np.abs(want_s["stat"][fin]), 1e-12)) < 1e-7
        assert np.all(sp[~fin] == 1.0)
        assert np.max(np.abs(np.log(sp[fin]) - np.log(want_s["p"][fin]))) < 1e-5
        assert np.max(np.abs(fs - want_f["stat"]) / np.maximum(np.abs(want_f["stat"]), 1e-300)) < 1e-12
        assert np.max(np.abs(fpv - want_f["p"]) / want_f["p"]) < 1e-9


def test_special_functions_against_golden_and_scipy(emul):
    for x, want in GOLD["special"]["kolmogorov"]:
        assert rel_err(emul.emul_kolmogorov_sf(x), want) < 1e-13
    xs = np.concatenate([np.linspace(0.01, 3, 3000), np.linspace(3, 30, 300)])
    got = np.array([emul.emul_kolmogorov_sf(x) for x in xs])
    want = sc.kolmogorov(xs)
    nz = want > 0
    assert np.max(np.abs(got[nz] - want[nz]) / want[nz]) < 1e-13 and np.all(got[~nz] == 0)
    for t, df, want in GOLD["special"]["t_two_sided"]:
        assert rel_err(emul.emul_student_t_two_sided(t, df), want) < 1e-10
    for df in [2.0, 3.3, 10.04, 93.8, 500.0, 3998.0, 9000.5]:
        for t in [1e-3, 0.3, 1.0, 1.6, 1.8, 2.5, 6.0, 25.0, 200.0, 1e4]:
            want = 2 * sc.stdtr(df, -t)
            if want > 1e-300:
                assert rel_err(emul.emul_student_t_two_sided(t, df), want) < 1e-9, (t, df)
    assert rel_err(emul.emul_student_t_two_sided(40.0, 198.0), 2 * GOLD["survey"]["edges"]["stdtr_198_-40"]) < 1e-10
    for x, k, want in GOLD["special"]["chdtrc"]:
        assert rel_err(emul.emul_chi2_sf_even(x, k), want) < 1e-11
    for z, want in GOLD["special"]["ndtr"]:
        assert rel_err(emul.emul_norm_sf(z), want) < 1e-13


@pytest.mark.parametrize("chains", [2, 4])
def test_fast_walks_match_oracle(emul, chains):
    """The KS-only fast walks of the lane kernel (two chains; four chains with a merge-path split),
    in their plain-C++ statement, against the oracle's searchsorted form: ties inside and across
    groups, every split position, trip counts taken from a longer row of the same warp."""
    import ctypes
    rng = np.random.default_rng(9 + chains)
    f = emul.emul_fast_walk
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int,
                  ctypes.c_int, ctypes.c_int]
    ran = 0
    for case in range(1500):
        n0, n1 = int(rng.integers(3, 129)), int(rng.integers(3, 129))
        dec = int(rng.choice([0, 1, 2, 6]))
        a = np.sort(np.round(rng.normal(0, 1, n0), dec).astype(np.float32))
        b = np.sort(np.round(rng.normal(rng.choice([0, 0.5, 3.0, -9.0]), 1, n1), dec).astype(np.float32))
        if case % 7 == 0:
            b[:] = a[rng.integers(0, n0, n1)]
            b.sort()
        tmax = n0 + n1 + int(rng.choice([0, 0, 1, 5, 40]))
        got = f(a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n0, b.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n1,
                tmax, chains)
        if got < 0:
            continue
        ran += 1
        want = o.per_position(a.astype(np.float64), b.astype(np.float64))["dnum"]
        assert got == want, (case, n0, n1, tmax, got, want)
    assert ran > 1000


# ---------------------------------------------------------------------------------------------
# grid keys (nm_lane.cuh): the 16-bit image of three-place decimals and the packed sort
# ---------------------------------------------------------------------------------------------
def test_grid_key_statement_on_every_float32(emul):
    """nm_grid_bits against ALL 2^32 float32 patterns: a value passes the check iff it is
    fl32(fl64(k / 1000)) with |k| <= 32766 (numpy's cast of the reference's round(x, 3)) or -0.0, and
    then both key patterns carry k + 32768.  Hence the key map is injective and monotone on whatever
    passes: order and ties of the keys are order and ties of the values."""
    if emul.variant != "float_imad":
        pytest.skip("same source in every build; checked once")
    n = ctypes.c_longlong()
    viol = emul.emul_grid_exhaustive(ctypes.byref(n), os.cpu_count() or 4)
    assert viol == 0 and n.value == 65534


def _run_grid(emul, a, b, want_u, want_t, walk):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    r = RowOut()
    rc = emul.emul_lane_position_grid(a.ctypes.data_as(FP), len(a), b.ctypes.data_as(FP), len(b), want_u, want_t, walk,
                                      ctypes.byref(r))
    return rc, r


def test_grid_lane_matches_oracle(emul):
    """The packed path -- one network pass over (group 0 | group 1 << 16) key pairs, walks over the 16-bit
    columns -- against the oracle: every network size boundary, unequal n, heavy ties, both signs."""
    rng = np.random.default_rng(23)
    sizes = list(range(3, 18)) + [23, 24, 25, 31, 32, 33, 63, 64, 65, 96, 97, 100, 101, 104, 105, 112, 120, 121, 127, 128]
    for trial in range(500):
        n0 = int(rng.choice(sizes)) if trial % 2 else int(rng.integers(3, 129))
        n1 = int(rng.choice([n0, 3, int(rng.integers(3, 129))]))
        dec = [3, 3, 1, 2, 0][trial % 5]
        a = np.round(rng.normal(0, [1, 3, 8][trial % 3], n0), dec)
        b = np.round(rng.normal(rng.choice([0, 0.5, 1, 3, -9]), 1, n1), dec)
        if trial % 11 == 0:
            b[:] = a[rng.integers(0, n0, n1)]
        a, b = np.clip(a, -32.766, 32.766).astype(np.float32), np.clip(b, -32.766, 32.766).astype(np.float32)
        ref = o.per_position(a.astype(np.float64), b.astype(np.float64))
        rc, r = _run_grid(emul, a, b, 1, 1, 0)
        assert rc == 0
        check(r, ref)
        for walk in (2, 4):
            rc, r = _run_grid(emul, a, b, 0, 0, walk)
            if rc == 0:
                assert r.dnum == ref["dnum"], (trial, n0, n1, walk)


def test_grid_lane_refuses_values_off_the_grid(emul):
    a = np.round(np.random.default_rng(1).normal(0, 1, 40), 3).astype(np.float32)
    b = a.copy()
    assert _run_grid(emul, a, b, 1, 1, 0)[0] == 0
    for bad in (np.nextafter(np.float32(0.417), np.float32(1)), np.float32(0.4175), np.float32(32.767), np.float32(-40.0),
                np.float32(1e10), np.float32(np.nan), np.float32(np.inf), np.float32(1e-30)):
        for grp in (0, 1):
            x, y = a.copy(), b.copy()
            (x if grp == 0 else y)[17] = bad
            assert _run_grid(emul, x, y, 1, 1, 0)[0] == 2, bad
    x = a.copy()
    x[3], x[4], x[5] = -0.0, 32.766, -32.766   # grid values all three
    rc, r = _run_grid(emul, x, b, 1, 1, 0)
    assert rc == 0
    check(r, o.per_position(x.astype(np.float64), b.astype(np.float64)))
