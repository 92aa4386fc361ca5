"""The oracle against (a) the known-answer vectors of SURVEY.md 8c, (b) golden fixtures produced
by calling scipy.stats directly (tests/golden/make_golden.py), (c) modern scipy on the code paths
that are algebraically unchanged since the pinned 1.2.1, (d) itself (vectorised vs scalar)."""
import json
import os

import numpy as np
import pytest
import scipy.stats as st

from oracle import nanomod_oracle as o
from oracle import nanomod_oracle_vec as ov
from conftest import rel_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.mark.parametrize("name", ["K1", "K2", "K3", "K4"])
def test_survey_known_answers(name):
    k = GOLD["survey"][name]
    if name == "K4":
        r = np.random.RandomState(1)
        a = np.round(r.normal(0, 1, 50), 3)
        b = np.round(r.normal(1, 1, 50), 3)
    else:
        a, b = k["a"], k["b"]
    res = o.per_position(a, b)
    assert res["dnum"] == k["dnum"] and res["twoU"] == k["twoU"]
    for key in ("pks", "pu", "t", "pt"):
        assert rel_err(res[key], k[key]) < 1e-12, key


def test_survey_combination_vectors():
    s5, s7 = GOLD["survey"]["stouffer5"], GOLD["survey"]["stouffer7"]
    for s, nb in ((s5, 2), (s7, 3)):
        assert o.stouffer_weights(nb, 2.0) == [float(x) for x in s["w"]]
        z, pz = o.combine_pvalues_stouffer(s["p"], s["w"])
        x2, px = o.combine_pvalues_fisher(s["p"])
        assert rel_err(z, s["Z"]) < 1e-12 and rel_err(pz, s["pz"]) < 1e-10
        assert rel_err(x2, s["X2"]) < 1e-12 and rel_err(px, s["px"]) < 1e-10
    assert o.combine_pvalues_stouffer([1.0, 0.01, 0.5], [50, 100, 50]) == (-np.inf, 1.0)
    e = GOLD["survey"]["edges"]
    z, _ = o.combine_pvalues_stouffer([o.FLOAT_MIN], [100.0])
    assert rel_err(z, e["isf_min"]) < 1e-12


def test_golden_cases_from_scipy_stats():
    for c in GOLD["cases"]:
        res = o.per_position(np.asarray(c["a"]), np.asarray(c["b"]))
        assert res["dnum"] == c["dnum"] and res["twoU"] == c["twoU"]
        assert rel_err(res["D"], c["D"]) < 1e-14
        for key in ("pks", "pu", "t", "pt"):
            want = max(c[key], o.FLOAT_MIN) if key != "t" else c[key]
            assert rel_err(res[key], want) < 1e-10, (key, res[key], want)


def test_golden_combos_from_scipy_stats():
    for c in GOLD["combos"]:
        w = o.stouffer_weights(c["nb"], c["WeightsDif"])
        z, pz = o.combine_pvalues_stouffer(c["p"], w)
        x2, px = o.combine_pvalues_fisher(c["p"])
        assert rel_err(z, c["stouffer"][0]) < 1e-12 and rel_err(pz, c["stouffer"][1]) < 1e-10
        assert rel_err(x2, c["fisher"][0]) < 1e-12 and rel_err(px, c["fisher"][1]) < 1e-10


def test_modern_scipy_cross_check_random():
    rng = np.random.default_rng(3)
    for trial in range(200):
        n0, n1 = int(rng.integers(5, 150)), int(rng.integers(5, 150))
        a = rng.normal(0, 1, n0)
        b = rng.normal(rng.choice([0.0, 0.5, 2.0]), 1, n1)
        if trial % 2:
            a, b = np.round(a, 2), np.round(b, 2)
        res = o.per_position(a, b)
        mw = st.mannwhitneyu(a, b, method="asymptotic")
        tt = st.ttest_ind(a, b, equal_var=False)
        assert rel_err(res["pu"], mw.pvalue / 2) < 1e-10
        assert rel_err(res["t"], tt.statistic) < 1e-12 and rel_err(res["pt"], max(tt.pvalue, o.FLOAT_MIN)) < 1e-10
        # modern 'asymp' KS uses a different p formula; D itself must agree exactly
        assert rel_err(res["D"], st.ks_2samp(a, b, method="asymp").statistic) < 1e-14


def test_clamps_and_degenerate():
    assert o.m_min_float(0.0) == o.FLOAT_MIN and o.m_min_float(1e-400) == o.FLOAT_MIN
    assert np.isnan(o.m_min_float(float("nan")))
    assert o.m_max_float(float("inf")) == o.FLOAT_MAX and o.m_max_float(-float("inf")) == -float("inf")
    with pytest.raises(ValueError):
        o.mannwhitneyu_legacy([1.0] * 6, [1.0] * 7)
    u, p, two_u, flag = o.mannwhitneyu_legacy([1.0] * 6, [1.0] * 7, strict=False)
    assert flag == 1 and np.isnan(p) and two_u == 42
    # disjoint deep groups: KS p underflows to 0 -> clamp
    res = o.per_position(np.arange(2000) / 2000.0, 10 + np.arange(2000) / 2000.0)
    assert res["pks"] == o.FLOAT_MIN and res["dnum"] == 2000 * 2000


def _mopts(pileup, **kw):
    d0, d1 = pileup.to_dicts()
    mo = o.default_moptions(**kw)
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    return mo


def test_vectorised_matches_scalar():
    import nanomod_b200 as nm
    p = nm.synthetic_pileup(800, 30, 45, poisson=True, drop_frac1=0.02, round_decimals=2, two_strands=True)
    mo = _mopts(p, neighborPvalues=3, testMethod="stouffer")
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    res = ov.detect(p.vals0, p.off0, p.vals1, p.off1, p.pos, p.seg, 5, 3, 2.0, ("stouffer", "fisher"))
    assert len(mo["sign_test"]) == len(res["dnum"])
    for r, (key, tests) in enumerate(mo["sign_test"]):
        assert key[2] == p.pos[res["row_pos_index"][r]] and (key[4], key[5]) == (res["n0"][r], res["n1"][r])
        pairs = [(tests[0][1], res["pu"][r]), (tests[1][0], res["t"][r]), (tests[1][1], res["pt"][r]),
                 (tests[2][1], res["pks"][r]), (tests[3][0], res["stouffer_stat"][r]),
                 (tests[3][1], res["stouffer_p"][r])]
        for a, b in pairs:
            assert rel_err(a, b) < 1e-9 or abs(a - b) < 1e-13
        assert int(round(2 * tests[0][0])) == res["twoU"][r]


def test_mtest2_row_order_gaps_and_table_text():
    import nanomod_b200 as nm
    p = nm.synthetic_pileup(300, 12, 12, drop_frac1=0.05, two_strands=True)
    mo = _mopts(p, neighborPvalues=2, testMethod="stouffer")
    o.mfilter_coverage(mo)
    o.mtest2(mo)
    st_ = mo["sign_test"]
    keys = [(k[0][0], k[0][1], k[0][2]) for k in st_]
    assert keys == sorted(keys)  # '+' < '-' is false in ASCII ('+'=43 < '-'=45 is true): sorted order
    # every row within nb of a gap or segment end has Z = -inf, p = 1 (hard part 4)
    for i, row in enumerate(st_):
        full = all(o.pos_check(st_, i, j) for j in range(i - 2, i + 3))
        if not full:
            assert row[1][3] == (-np.inf, 1.0)
        else:
            assert np.isfinite(row[1][3][0])
    line = o.save_test_lines(mo)[0].split()
    assert len(line) == 14 and line[0] == "syn" and line[2] == str(st_[0][0][2] + 1)
    mo["testMethod"] = "ks"
    o.mtest2(mo)
    assert len(o.save_test_lines(mo)[0].split()) == 12
    sites = o.called_sites(mo)
    assert len(sites) <= mo["topN"]


def test_meanstd_lines_format():
    """myDetect.py:540-545: 0-based position, np.std with ddof=0, three decimals."""
    mo = o.default_moptions(mstd=1, MinCoverage=3)
    mo["ds2"] = ["a", "b"]
    mo["a"] = {"norm_mean": {("chr1", "+"): {4: [0.1, 0.2, 0.4], 5: [1.0, 1.0, 1.0]}}, "base": {("chr1", "+"): {4: "A", 5: "C"}}}
    mo["b"] = {"norm_mean": {("chr1", "+"): {4: [0.5, 0.7, 0.9], 5: [2.0, 2.5, 3.0]}}, "base": {("chr1", "+"): {4: "A", 5: "C"}}}
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    lines = o.save_meanstd_lines(mo)
    assert lines[0] == "chr1 + 4 A 0.233 0.125 0.700 0.163\n"
    assert lines[1] == "chr1 + 5 C 1.000 0.000 2.500 0.408\n"


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = o.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want
    idx = o.downsample_indices(20190131, 3, 1234, 7, 1, 1000, 37)
    assert idx.min() >= 0 and idx.max() < 37 and len(np.unique(idx)) == 37


def test_downsampling_branch_matches_reference_procedure_statistically():
    """The reference resamples with np.random.choice on an unseeded generator (myDetect.py:345-361).
    Its procedure, run literally with a seeded numpy generator, and the oracle's branch (the
    library's Philox stream) must give the same DISTRIBUTION of the selected p-value."""
    rs = np.random.RandomState(3)
    a = np.round(rs.normal(0, 1, 80), 3)
    b = np.round(rs.normal(0.6, 1, 70), 3)
    cov, times, q = 25, 100, 0.25

    def literal(rng):
        p_array = np.zeros(times)
        for i in range(times):
            _st, pks, _ = o.ks_2samp_legacy(rng.choice(a, cov), rng.choice(b, cov))
            p_array[i] = o.m_min_float(pks)
        return p_array[np.argsort(p_array)[int(times * q)]]

    ref = np.array([literal(np.random.RandomState(1000 + k)) for k in range(120)])
    mo = o.default_moptions(coverages=[cov, cov], downsampling=times, downsampling_quantile=q)
    ours = []
    for k in range(120):
        mo["seed"] = 555 + k
        mo["_ds_site"] = (0, k)
        ours.append(o.getKStest(mo, a, b, "+", strict=False)[2][1])
    ours = np.array(ours)
    assert st.ks_2samp(np.log(ref), np.log(ours)).pvalue > 1e-3
    assert abs(np.median(np.log(ref)) - np.median(np.log(ours))) < 0.35
    # U and t are computed on the full groups (myDetect.py:331-337)
    full = o.getKStest(o.default_moptions(), a, b, "+", strict=False)
    mo["_ds_site"] = (0, 0)
    ds = o.getKStest(mo, a, b, "+", strict=False)
    assert ds[0] == full[0] and ds[1] == full[1] and ds[2] != full[2]
