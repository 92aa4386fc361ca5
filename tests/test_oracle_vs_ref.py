"""Pin the oracle to the reference's OWN code (CPU only).

tests/golden/ref_cases.* were produced by the reference's `myDetect.py` (rendered to Python 3 by
oracle/build_ref.py, scipy calls bound to the 1.2.1 semantics of oracle/scipy_legacy.py).  Here:
  * the restated oracle must reproduce every fixture: rows, every statistic, the ranking, the
    called sites and the `save_test` text;
  * where /root/reference is present (this container, not the GPU box) the reference is executed
    live -- against the fixtures (they are not stale) and against the oracle on BASELINE cfg1 at
    full size (10 000 positions, 2x50, continuous float32 values);
  * the host side of the product (`SignTestTable`: ranking, region ranking, called sites, text)
    is checked against the same fixtures, fed with the fixture's numbers;
  * the packer is checked against the reference's own `mReadSignalBase` through an in-memory h5py.
"""
import copy
import sys
import tempfile
import types

import numpy as np
import pytest

import nanomod_b200 as nm
from nanomod_b200 import packer
from nanomod_b200.detect import SignTestTable
from oracle import nanomod_oracle as o
from oracle import nanomod_oracle_vec as ov
from oracle import ref_loader as rl

import golden_ref as G

live = pytest.mark.skipif(not rl.available(), reason="reference tree / oracle/_ref not present")


def run_oracle(name):
    mo = G.case_moptions(name, o.default_moptions)
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=True)
    return mo


def assert_sign_test_equal(st, case, rtol):
    assert [[m[0][0], m[0][1], int(m[0][2]), m[0][3], int(m[0][4]), int(m[0][5])] for m in st] == case["rows"]
    for m, want in zip(st, case["stats"]):
        got = [x for tup in m[1] for x in tup]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert G.same_float(g, w, rtol), (m[0], got, want)


@pytest.mark.parametrize("name", G.CASE_NAMES)
def test_oracle_reproduces_reference_golden(name):
    case = G.REF["cases"][name]
    mo = run_oracle(name)
    st = mo["sign_test"]
    assert_sign_test_equal(st, case, 1e-12)
    index_of = {m[0][:3]: i for i, m in enumerate(st)}
    assert [index_of[m[0][:3]] for m in mo["sorted_sign_test"]] == case["sorted"]
    assert [list(s) for s in o.called_sites(mo)] == case["called_sites"]
    assert "".join(o.save_test_lines(mo)) == case["sign_test_txt"]
    if case["meanstd_cvs"] is not None:
        assert "".join(o.save_meanstd_lines(mo)) == case["meanstd_cvs"]


@pytest.mark.parametrize("name", G.CASE_NAMES)
def test_vectorised_oracle_reproduces_reference_golden(name):
    case = G.REF["cases"][name]
    p, opt = G.case_pileup(name), G.case_options(name)
    methods = () if opt.testMethod == "ks" else (opt.testMethod,)
    res = ov.detect(p.vals0, p.off0, p.vals1, p.off1, p.pos, p.seg, opt.MinCoverage, opt.neighborPvalues,
                    opt.WeightsDif, methods)
    want = G.stats_columns(name)
    assert len(res["dnum"]) == len(want)
    idx = res["row_pos_index"]
    assert [int(x) for x in p.pos[idx]] == [r[2] for r in case["rows"]]
    assert [int(x) for x in res["n0"]] == [r[4] for r in case["rows"]]
    assert np.array_equal(res["twoU"], np.round(2 * want[:, 0]).astype(np.int64))
    assert np.array_equal(res["dnum"], np.round(want[:, 4] * res["n0"] * res["n1"]).astype(np.int64))
    cols = [("U", 0), ("pu", 1), ("t", 2), ("pt", 3), ("D", 4), ("pks", 5)]
    if methods:
        cols += [(methods[0] + "_stat", 6), (methods[0] + "_p", 7)]
    for key, c in cols:
        for g, w in zip(res[key], want[:, c]):
            assert G.same_float(g, w, 1e-11), (key, g, w)


def test_known_answer_vectors_through_reference_getKStest():
    """SURVEY 8c K1..K4 were computed in the survey with restated formulas; the fixture holds what
    the reference's getKStest returns for them, and both must agree with the oracle."""
    survey = {"K1": (0.6, 0.2089848305751669, 3.0, 0.030051402969433157, -2.5000000000000004, 0.03694203771362409),
              "K2": (0.5714285714285714, 0.15504417912365295, 7.0, 0.022835621469841673, -2.506433439961463,
                     0.031015852000315595),
              "K3": (1.0, 7.26220915473708e-06, 0.0, 4.3669632941789134e-05, -14.849242404917499,
                     2.9120185192417662e-12),
              "K4": (0.58, 3.761754930361375e-08, 444.0, 1.4042661075413371e-08, -6.586679477347367,
                     2.5818774981476798e-09)}
    for name, k in G.REF["known"].items():
        (u, pu), (t, pt), (d, pks) = k["result"]
        got = o.getKStest(o.default_moptions(), k["a"], k["b"], "+")
        for g, w in zip([x for tup in got for x in tup], (u, pu, t, pt, d, pks)):
            assert G.same_float(g, w, 1e-12)
        D, PKS, U, PU, T, PT = survey[name]
        # the survey's numbers are for float64 inputs, the fixture's for their float32 images
        for g, w, tol in ((d, D, 1e-12), (pks, PKS, 1e-9), (u, U, 0), (pu, PU, 1e-9), (t, T, 1e-6), (pt, PT, 1e-6)):
            assert G.same_float(g, w, tol), (name, g, w)


# ------------------------------------------------------------------------------------------
# the reference executed live (this container only)
# ------------------------------------------------------------------------------------------
def run_reference(name):
    from golden.make_ref_golden import to_groups  # the fixture's own dict builder
    case = G.REF["cases"][name]
    inp = {k: G._NPZ["%s/%s" % (name, k)] for k in ("vals0", "off0", "vals1", "off1", "pos", "seg", "base")}
    groups, bases = to_groups(inp, case["seg_names"])
    mo = rl.moptions_from_groups(groups, bases, **case["options"])
    return rl.run_detect(mo)


@live
@pytest.mark.parametrize("name", G.CASE_NAMES)
def test_live_reference_matches_committed_golden(name):
    case = G.REF["cases"][name]
    mo = run_reference(name)
    assert_sign_test_equal(mo["sign_test"], case, 0.0)
    index_of = {m[0][:3]: i for i, m in enumerate(mo["sign_test"])}
    assert [index_of[m[0][:3]] for m in mo["sorted_sign_test"]] == case["sorted"]
    assert [list(s) for s in rl.called_sites(mo)] == case["called_sites"]
    with tempfile.TemporaryDirectory() as td:
        text, mstd = rl.save_test_text(mo, td)
    assert text == case["sign_test_txt"] and mstd == case["meanstd_cvs"]


@live
def test_live_reference_equals_oracle_on_cfg1_full_size():
    """BASELINE configs[0] at full size: 10 kb, 2x50 reads, KS + weighted Stouffer +-3."""
    p = nm.synthetic_pileup(10000, 50, 50)
    d0, d1 = p.to_dicts()
    kw = dict(MinCoverage=5, neighborPvalues=3, WeightsDif=2.0, testMethod="stouffer", rankUse="pv", topN=30, window=10)
    mo_o = o.default_moptions(SaveTest=0, **kw)
    mo_o["ds2"] = ["g0", "g1"]
    mo_o["g0"], mo_o["g1"] = copy.deepcopy(d0), copy.deepcopy(d1)
    o.mfilter_coverage(mo_o)
    o.mtest2(mo_o, strict=True)
    mo_r = rl.default_moptions(**kw)
    mo_r["g0"], mo_r["g1"] = d0, d1
    rl.run_detect(mo_r)
    assert len(mo_o["sign_test"]) == len(mo_r["sign_test"]) == 10000
    for a, b in zip(mo_o["sign_test"], mo_r["sign_test"]):
        assert a[0] == b[0]
        for g, w in zip([x for tup in a[1] for x in tup], [x for tup in b[1] for x in tup]):
            assert G.same_float(g, w, 1e-12)
    assert [m[0] for m in mo_o["sorted_sign_test"]] == [m[0] for m in mo_r["sorted_sign_test"]]
    assert o.called_sites(mo_o) == rl.called_sites(mo_r)
    with tempfile.TemporaryDirectory() as td:
        text, _ = rl.save_test_text(mo_r, td)
    assert "".join(o.save_test_lines(mo_o)) == text


# ------------------------------------------------------------------------------------------
# product host logic against the reference's vectors (no GPU: the table is filled from the fixture)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", G.CASE_NAMES)
def test_product_host_logic_on_reference_vectors(name):
    case = G.REF["cases"][name]
    t = G.table_from_fixture(name)
    assert [int(r) for r in t.sorted_rows()] == case["sorted"]
    assert [list(s) for s in t.called_sites()] == case["called_sites"]
    assert "".join(t.format_lines()) == case["sign_test_txt"]
    # (the GPU test compares the order of ITS numbers with the order of these after rounding the
    # ranking keys of both to 11 digits: the reference breaks exact-D ties by its own last-bit noise)
    assert sorted(int(r) for r in G.snapped(t).sorted_rows()) == sorted(case["sorted"])


# ------------------------------------------------------------------------------------------
# packer vs the reference's own read loop (mReadSignalBase through an in-memory h5py)
# ------------------------------------------------------------------------------------------
def _reads(rng, n_reads, chroms=("chrA", "chrB"), lo=30, hi=90, span=400):
    out = []
    for _ in range(n_reads):
        n = int(rng.integers(lo, hi))
        out.append({"chrom": chroms[int(rng.integers(len(chroms)))], "strand": "+-"[int(rng.integers(2))],
                    "start": int(rng.integers(0, span)), "norm_mean": np.round(rng.normal(0, 1, n), 3),
                    "base": "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))})
    return out


@live
def test_fast5_paths_are_the_reference_constants():
    f5 = rl.load().myFast5
    assert packer.FAST5_EVENTS == f5.raw_event_ful            # myFast5.py:92
    assert packer.FAST5_ALIGNMENT == f5.rawAlignment_full     # myFast5.py:113
    assert "NanomoCorrected_000" in packer.FAST5_EVENTS       # myCom.py:48-51


@live
@pytest.mark.parametrize("min_lr", [20, 60])
def test_packer_equals_reference_mReadSignalBase(min_lr):
    rng = np.random.default_rng(5)
    reads = [_reads(rng, 60), _reads(rng, 70)]
    mo = rl.default_moptions(min_lr=min_lr, min_lr_nb=0, MinCoverage=3)
    for g in (0, 1):
        rl.read_reads(mo, g, reads[g])
    ref = nm.Pileup.from_dicts(mo["g0"], mo["g1"])
    flt = packer.ReadFilter(min_lr=min_lr)
    recs = [[packer.ReadRecord(r["chrom"], r["strand"], r["start"], r["norm_mean"], r["base"]) for r in grp]
            for grp in reads]
    got = packer.pack_reads(recs[0], recs[1], flt)
    assert got.seg_names == ref.seg_names
    for k in ("off0", "off1", "pos", "seg", "base"):
        assert np.array_equal(getattr(got, k), getattr(ref, k)), k
    assert np.array_equal(got.vals0[:got.off0[-1]], ref.vals0[:ref.off0[-1]])
    assert np.array_equal(got.vals1[:got.off1[-1]], ref.vals1[:ref.off1[-1]])


def test_read_fast5_opens_the_nanomod_group(monkeypatch, tmp_path):
    """read_fast5 with a stand-in h5py: the file only has the NanoMod-annotated group."""
    ev = np.zeros(4, dtype=[("norm_mean", "f8"), ("base", "S1")])
    ev["norm_mean"] = [0.1, 0.2, 0.3, 0.4]
    ev["base"] = [b"A", b"C", b"G", b"T"]

    class DS:
        def __init__(self, value=None, attrs=None):
            self.value, self.attrs = value, attrs or {}

        def __getitem__(self, k):
            return self.value

    class F(dict):
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    store = F({"/Analyses/NanomoCorrected_000/BaseCalled_template/Events": DS(ev),
               "/Analyses/NanomoCorrected_000/BaseCalled_template/Alignment":
                   DS(attrs={"mapped_chrom": "chrA", "mapped_strand": "+", "mapped_start": 12})})
    fake = types.ModuleType("h5py")
    fake.File = lambda path, mode="r": store
    monkeypatch.setitem(sys.modules, "h5py", fake)
    f = tmp_path / "read.fast5"
    f.write_bytes(b"")
    r = packer.read_fast5(str(f))
    assert r is not None and (r.chrom, r.strand, r.start) == ("chrA", "+", 12)
    assert np.allclose(r.norm_mean, [0.1, 0.2, 0.3, 0.4]) and bytes(r.base) == b"ACGT"
