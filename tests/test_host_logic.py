"""Host-side logic that needs no GPU: pileup packing, option validation (reference messages),
table text, ranking and the called-site rule against the oracle, and the C-ABI library's
symbols + its refusal to run without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import nanomod_b200 as nm
from nanomod_b200 import _lib
from nanomod_b200.detect import SignTestTable
from oracle import nanomod_oracle as o
from oracle import nanomod_oracle_vec as ov

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_table(p, opt):
    """A SignTestTable filled from the vectorised oracle (what the GPU would return)."""
    res = ov.detect(p.vals0, p.off0, p.vals1, p.off1, p.pos, p.seg, opt.MinCoverage, opt.neighborPvalues,
                    opt.WeightsDif, ("stouffer", "fisher"))
    idx = res["row_pos_index"]
    return SignTestTable(options=opt, seg_names=p.seg_names, seg=p.seg[idx], pos=p.pos[idx], base=p.base[idx],
                         n0=res["n0"], n1=res["n1"], ks_dnum=res["dnum"], ks_d=res["D"], ks_p=res["pks"],
                         two_u=res["twoU"], u_stat=res["U"], u_p=res["pu"], t_stat=res["t"], t_p=res["pt"],
                         fisher_stat=res.get("fisher_stat"), fisher_p=res.get("fisher_p"),
                         stouffer_stat=res.get("stouffer_stat"), stouffer_p=res.get("stouffer_p"))


def oracle_moptions(p, opt, **over):
    d0, d1 = p.to_dicts()
    mo = o.default_moptions(MinCoverage=opt.MinCoverage, neighborPvalues=opt.neighborPvalues,
                            WeightsDif=opt.WeightsDif, testMethod=opt.testMethod, rankUse=opt.rankUse,
                            topN=opt.topN, window=opt.half_window, **over)
    mo["ds2"] = ["g0", "g1"]
    mo["g0"], mo["g1"] = d0, d1
    o.mfilter_coverage(mo)
    o.mtest2(mo, strict=False)
    return mo


def table_from_sign_test(p, opt, mo):
    """SignTestTable carrying the scalar oracle's exact numbers (so ranking ties break alike)."""
    st_ = mo["sign_test"]
    seg_id = {sk: i for i, sk in enumerate(p.seg_names)}
    col = lambda f: np.array([f(m) for m in st_])
    kw = {}
    if opt.testMethod != "ks":
        kw = {opt.testMethod + "_stat": col(lambda m: m[1][3][0]), opt.testMethod + "_p": col(lambda m: m[1][3][1])}
    return SignTestTable(options=opt, seg_names=p.seg_names,
                         seg=col(lambda m: seg_id[(m[0][0], m[0][1])]).astype(np.int32),
                         pos=col(lambda m: m[0][2]).astype(np.int32),
                         base=col(lambda m: ord(m[0][3])).astype(np.uint8),
                         n0=col(lambda m: m[0][4]), n1=col(lambda m: m[0][5]),
                         ks_dnum=np.zeros(len(st_), np.int32), ks_d=col(lambda m: m[1][2][0]),
                         ks_p=col(lambda m: m[1][2][1]), u_stat=col(lambda m: m[1][0][0]),
                         u_p=col(lambda m: m[1][0][1]), t_stat=col(lambda m: m[1][1][0]),
                         t_p=col(lambda m: m[1][1][1]), **kw)


def test_pileup_roundtrip_and_order():
    p = nm.synthetic_pileup(400, 9, 11, drop_frac1=0.1, two_strands=True, round_decimals=3)
    p.validate()
    d0, d1 = p.to_dicts()
    q = nm.Pileup.from_dicts(d0, d1)
    c0, c1 = p.counts()
    both = (c0 > 0) & (c1 > 0)
    assert q.n_pos == int(both.sum())
    assert np.array_equal(q.pos, p.pos[both]) and np.array_equal(q.seg, p.seg[both])
    assert q.seg_names == [("syn", "+"), ("syn", "-")]
    assert np.array_equal(q.vals0[:q.off0[-1]], np.concatenate([p.group(0, i) for i in np.nonzero(both)[0]]))
    assert q.vals0.shape[0] >= _lib.padded_len(q.off0[-1]) and q.vals0.ctypes.data % 16 == 0
    s = p.slice_rows(100, 250)
    assert s.n_pos == 150 and np.array_equal(s.group(1, 0), p.group(1, 100))


def test_pileup_npz_roundtrip(tmp_path):
    p = nm.synthetic_pileup(50, 6, 7, two_strands=True)
    f = str(tmp_path / "p.npz")
    p.save_npz(f)
    q = nm.Pileup.load_npz(f)
    assert np.array_equal(q.vals1[:q.off1[-1]], p.vals1[:p.off1[-1]]) and q.seg_names == p.seg_names


def test_option_defaults_match_reference_cli():
    d = nm.DetectOptions()
    assert (d.window, d.FileID, d.outFolder, d.MinCoverage, d.topN, d.neighborPvalues, d.WeightsDif) == \
        (21, "mod", "mRes/", 5, 30, 2, 2.0)
    assert (d.testMethod, d.rankUse, d.SaveTest, d.RegionRankbyST, d.percentile, d.WindOvlp, d.NA) == \
        ("stouffer", "pv", 1, 0, 0.1, 0, "")
    assert (d.min_lr, d.min_lr_nb, d.downsampling_quantile, d.downsampling, d.coverages, d.outLevel) == \
        (500, 0, 0.25, 100, "0-0", 2)
    assert d.half_window == 10


def test_option_validation_messages():
    with pytest.raises(nm.OptionError, match=r"The coverage \(2\) is too small"):
        nm.DetectOptions(MinCoverage=2).validate()
    with pytest.raises(nm.OptionError, match=r"The neighborPvalues \(-1\) cannot be smaller than 0"):
        nm.DetectOptions(neighborPvalues=-1).validate()
    with pytest.raises(nm.OptionError, match=r"Window size \(2\) is too small"):
        nm.DetectOptions(window=2).validate()
    with pytest.raises(nm.OptionError, match="topN"):
        nm.DetectOptions(topN=0).validate()
    d = nm.DetectOptions(WeightsDif=0.5)
    d.validate()
    assert d.WeightsDif == 1.0  # NanoMod.py:77-78
    nm.DetectOptions(coverages="50-50").validate()  # the down-sampling branch is supported
    assert nm.DetectOptions(coverages="40").coverage_pair() == (40, 40)  # NanoMod.py:174-176
    assert list(nm.DetectOptions(coverages="0-12").seg_cov([("c", "+"), ("c", "-"), ("d", "+")])) == [0, 12, 0]
    assert nm.DetectOptions().seg_cov([("c", "+")]) is None
    with pytest.raises(nm.OptionError, match="downsampling"):
        nm.DetectOptions(coverages="50", downsampling=0).validate()
    with pytest.raises(nm.OptionError, match="downsampling_quantile"):
        nm.DetectOptions(coverages="50", downsampling_quantile=1.0).validate()
    p = nm.DetectOptions(coverages="50", downsampling=100, downsampling_quantile=0.25, seed=9).to_params()
    assert (p.ds_times, p.ds_index, p.ds_seed) == (100, 25, 9)


@pytest.mark.parametrize("method,rank", [("stouffer", "pv"), ("fisher", "pv"), ("ks", "pv"), ("stouffer", "st")])
def test_table_text_ranking_called_sites_match_oracle(method, rank):
    p = nm.synthetic_pileup(3000, 25, 25, drop_frac1=0.004, two_strands=True)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod=method, rankUse=rank, topN=8)
    mo = oracle_moptions(p, opt)
    t = table_from_sign_test(p, opt, mo)
    assert t.format_lines() == o.save_test_lines(mo)
    assert oracle_table(p, opt).format_lines() == o.save_test_lines(mo)  # vectorised oracle, 4 digits
    want_order = [(m[0][0], m[0][1], m[0][2]) for m in mo["sorted_sign_test"]]
    got_order = [(t.seg_names[t.seg[r]][0], t.seg_names[t.seg[r]][1], int(t.pos[r])) for r in t.ranked()]
    assert got_order == want_order
    assert t.called_sites() == o.called_sites(mo)
    assert len(t.called_sites()) == 8
    st_ = t.to_sign_test()
    assert st_[5][0] == mo["sign_test"][5][0] and len(st_[5][1]) == len(mo["sign_test"][5][1])


def test_called_sites_find_planted_positions():
    p = nm.synthetic_pileup(10000, 50, 50)
    opt = nm.DetectOptions(neighborPvalues=3, topN=6)
    sites = oracle_table(p, opt).called_sites()
    assert len(sites) == 6
    assert all(abs((s[2] % 1000) - 500) <= 2 for s in sites)


def test_save_test_file(tmp_path):
    p = nm.synthetic_pileup(40, 6, 6)
    opt = nm.DetectOptions(outFolder=str(tmp_path), FileID="abc")
    t = oracle_table(p, opt)
    path = t.save_test()
    assert path.endswith("abc_sign_test.txt")
    lines = open(path).read().splitlines()
    assert len(lines) == 40
    assert re.match(r"^syn \+ 1 A 6 6 \d+\.\d{3} \d\.\d{3}E[-+]\d+ -?\d+\.\d{3} ", lines[0])
    assert lines[0].endswith("-inf 1.000E+00")  # first nb rows: missing neighbours -> Z = -inf, p = 1
    opt.SaveTest = 0
    assert t.save_test() is None


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "nanomod_b200.h")).read()
    declared = set(re.findall(r"\b(nm_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert os.path.exists(_lib.LIB_PATH), "run `python -m nanomod_b200.build` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.nm_version.restype = ctypes.c_int
    assert lib.nm_version() == 120
    assert _lib.load().nm_padded_len(5) == _lib.padded_len(5) == 12


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors vs the real header, measured by a C program compiled from include/."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nanomod_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(nm_params), sizeof(nm_pileup),'
                   ' sizeof(nm_table), offsetof(nm_table, flags), offsetof(nm_table, moments));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(_lib.nm_params), ctypes.sizeof(_lib.nm_pileup), ctypes.sizeof(_lib.nm_table),
                   _lib.nm_table.flags.offset, _lib.nm_table.moments.offset]
    assert got[:3] == [48, 112, 17 * 8]
    # the head-exchange structures (nm_head_row = HEAD_REC of nanomod_b200.sharded, nm_head_geometry, nm_head_peers)
    from nanomod_b200.sharded import HEAD_REC
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nanomod_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %d %d\\n", sizeof(nm_head_row), offsetof(nm_head_row, key),'
                   ' offsetof(nm_head_row, reserved), sizeof(nm_head_geometry), sizeof(nm_head_peers),'
                   ' offsetof(nm_head_peers, epoch), offsetof(nm_head_peers, base), NM_MAX_PEERS, NM_IPC_HANDLE_BYTES);return 0;}\n')
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [HEAD_REC.itemsize, HEAD_REC.fields["key"][1], HEAD_REC.fields["reserved"][1],
                   ctypes.sizeof(_lib.nm_head_geometry), ctypes.sizeof(_lib.nm_head_peers), _lib.nm_head_peers.epoch.offset,
                   _lib.nm_head_peers.base.offset, _lib.NM_MAX_PEERS, _lib.NM_IPC_HANDLE_BYTES]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nm.NmError) as e:
        nm.Detector(0)
    assert e.value.code == 6  # NM_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nanomod_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


# ---------------------------------------------------------------------------------------------
# host packer against the oracle's literal restatement of mReadSignalBase
# ---------------------------------------------------------------------------------------------
def _synthetic_reads(rng, n_reads, chroms=("chrA", "chrB"), lo=80, hi=400, span=3000):
    from nanomod_b200.packer import ReadRecord
    reads = []
    for _ in range(n_reads):
        n = int(rng.integers(lo, hi))
        reads.append(ReadRecord(str(rng.choice(chroms)), str(rng.choice(["+", "-"])), int(rng.integers(0, span)),
                                np.round(rng.normal(0, 1, n), 3), rng.choice(np.frombuffer(b"ACGT", np.uint8), n)))
    return reads


def _oracle_pileup(reads0, reads1, **mo_over):
    mo = o.default_moptions(**dict({"min_lr": 100, "min_lr_nb": 0}, **mo_over))
    mo["ds2"] = ["g0", "g1"]
    for name, reads in (("g0", reads0), ("g1", reads1)):
        mo[name] = {"norm_mean": {}, "base": {}}
        mo["cur_wrkBase"] = name
        for r in reads:
            o.mReadSignalBase_events(mo, r.chrom, r.start, r.strand, r.norm_mean, [chr(c) for c in r.base])
    return nm.Pileup.from_dicts(mo["g0"], mo["g1"])


@pytest.mark.parametrize("mode", ["plain", "region", "pos2_chr", "length_band"])
def test_packer_matches_reference_read_loop(mode):
    from nanomod_b200.packer import ReadFilter, pack_reads
    rng = np.random.default_rng(11)
    reads0, reads1 = _synthetic_reads(rng, 120), _synthetic_reads(rng, 110)
    over, flt = {}, ReadFilter(min_lr=100)
    if mode == "region":
        over = {"start_pos": 1500, "end_pos": 1520}
        flt = ReadFilter(min_lr=100, start_pos=1500, end_pos=1520)
    elif mode == "pos2_chr":
        over = {"Chr": "chrB", "Pos": 900, "Pos2": 1400}
        flt = ReadFilter(min_lr=100, Chr="chrB", Pos=900, Pos2=1400)
    elif mode == "length_band":
        for r in reads0[:30] + reads1[:30]:  # reads that start and end near 0 / 8000
            r.start = int(rng.integers(0, 40))
            n = int(rng.integers(7975, 8020)) - r.start
            r.norm_mean = np.round(rng.normal(0, 1, n), 3)
            r.base = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
        over = {"min_lr": 8000, "min_lr_nb": 50}
        flt = ReadFilter(min_lr=8000, min_lr_nb=50)
    want = _oracle_pileup(reads0, reads1, **over)
    assert flt.__dict__ == ReadFilter.from_moptions(dict({"min_lr": 100, "min_lr_nb": 0}, **over)).__dict__
    got = pack_reads(reads0, reads1, flt)
    assert got.n_pos == want.n_pos and got.n_pos > 0
    assert got.seg_names == want.seg_names
    for f in ("vals0", "off0", "vals1", "off1", "pos", "seg", "base"):
        a, b = getattr(got, f), getattr(want, f)
        n = int(got.off0[-1]) if f == "vals0" else int(got.off1[-1]) if f == "vals1" else len(a)
        assert np.array_equal(a[:n], b[:n]), f


def test_packer_npz_roundtrip_and_missing_h5py(tmp_path):
    from nanomod_b200 import packer
    rng = np.random.default_rng(2)
    reads = _synthetic_reads(rng, 9)
    packer.save_reads_npz(str(tmp_path / "r.npz"), reads)
    back = packer.load_reads_npz(str(tmp_path / "r.npz"))
    assert len(back) == 9
    for a, b in zip(reads, back):
        assert (a.chrom, a.strand, a.start) == (b.chrom, b.strand, b.start)
        assert np.array_equal(a.norm_mean, b.norm_mean) and np.array_equal(a.base, b.base)
    (tmp_path / "d" / "mall").mkdir(parents=True)
    (tmp_path / "d" / "x").mkdir()
    for f in ("d/a.fast5", "d/x/b.fast5", "d/mall/skip.fast5", "d/x/c.txt"):
        (tmp_path / f).write_text("")
    assert sorted(os.path.basename(f) for f in packer.walk_fast5(str(tmp_path / "d"))) == ["a.fast5", "b.fast5"]
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            packer.read_fast5(str(tmp_path / "d" / "a.fast5"))


@pytest.mark.parametrize("ovlp,na,rank,method", [(0, "", "pv", "stouffer"), (1, "", "pv", "stouffer"), (0, "A", "pv", "fisher"),
                                                 (1, "C", "st", "stouffer"), (0, "", "pv", "ks")])
def test_region_rank_mode_matches_oracle(ovlp, na, rank, method):
    """RegionRankbyST=1 (myDetect.py:463-515): windowed percentile ranking, --NA base filter,
    overlap suppression -- the table's vectorised version against the literal restatement."""
    p = nm.synthetic_pileup(2500, 9, 9, drop_frac1=0.01, two_strands=True, round_decimals=1)
    opt = nm.DetectOptions(neighborPvalues=2, testMethod=method, rankUse=rank, RegionRankbyST=1, WindOvlp=ovlp,
                           NA=na, percentile=0.1, window=25 if na else 9)
    mo = oracle_moptions(p, opt, RegionRankbyST=1, WindOvlp=ovlp, NA=na, percentile=0.1)
    assert mo["window"] == opt.half_window + 1
    t = table_from_sign_test(p, opt, mo)
    rows = t.region_ranked()
    got = [(t.seg_names[t.seg[r]][0], t.seg_names[t.seg[r]][1], int(t.pos[r])) for r in rows]
    want = [(m[0][0], m[0][1], m[0][2]) for m in mo["sorted_sign_test"]]
    assert len(want) > (5 if na else 20)
    assert got == want


@pytest.mark.parametrize("method,nb", [("stouffer", 2), ("fisher", 3), ("ks", 2), ("stouffer", 0)])
def test_native_table_text_equals_python_format(method, nb, tmp_path):
    """nm_format_sign_test (host threads in the library) against SignTestTable.format_lines, which
    is the reference's '%' formatting (myDetect.py:522-538) -- including the values that test
    printf: infinities, NaN, -0.0, DBL_MIN, DBL_MAX, denormals, exact halves."""
    rng = np.random.default_rng(4)
    n = 5000
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 2.2250738585072014e-308, 1.7976931348623157e308, 5e-324,
                        0.0005, 0.0015, 0.0025, 1.0005, 2.5, 0.125, 1e-5, 9.9995, 9.9995e-7, 123456.7895, -1e15])

    def col():
        x = rng.normal(0, 1, n) * 10.0 ** rng.integers(-12, 6, n)
        idx = rng.choice(n, 400, replace=False)
        x[idx] = rng.choice(special, 400)
        return x
    opt = nm.DetectOptions(neighborPvalues=nb, testMethod=method, outFolder=str(tmp_path), FileID="nat")
    kw = {}
    if method != "ks":
        kw = {method + "_stat": col(), method + "_p": np.abs(col())}
    t = SignTestTable(options=opt, seg_names=[("chrI", "+"), ("chrI", "-"), ("a_very_long_contig_name_0123456789", "+")],
                      seg=np.sort(rng.integers(0, 3, n)).astype(np.int32), pos=rng.integers(0, 2**31 - 2, n).astype(np.int32),
                      base=rng.choice(np.frombuffer(b"ACGT", np.uint8), n), n0=rng.integers(3, 5000, n).astype(np.int32),
                      n1=rng.integers(3, 5000, n).astype(np.int32), ks_dnum=np.zeros(n, np.int32), ks_d=np.abs(col()),
                      ks_p=np.abs(col()), u_stat=col(), u_p=np.abs(col()), t_stat=col(), t_p=np.abs(col()), **kw)
    want = "".join(t.format_lines()).encode()
    for threads in (1, 3, 16):
        assert t.format_text(threads) == want
    path = t.save_test()
    assert open(path, "rb").read() == want
    # columns that were not computed print as zeros, as in format_lines
    t2 = SignTestTable(options=opt, seg_names=t.seg_names, seg=t.seg, pos=t.pos, base=t.base, n0=t.n0, n1=t.n1,
                       ks_dnum=t.ks_dnum, ks_d=t.ks_d, ks_p=t.ks_p, **kw)
    assert t2.format_text() == "".join(t2.format_lines()).encode()
    empty = SignTestTable(options=opt, seg_names=t.seg_names, seg=t.seg[:0], pos=t.pos[:0], base=t.base[:0], n0=t.n0[:0],
                          n1=t.n1[:0], ks_dnum=t.ks_dnum[:0], ks_d=t.ks_d[:0], ks_p=t.ks_p[:0])
    assert empty.format_text() == b""


def test_bench_and_entry_scripts_compile_and_parse_their_arguments():
    """bench.py / __graft_entry__.py / the tools are run by the driver on a GPU box: a syntax error or a bad
    argument table there would only show at round end."""
    import py_compile
    import subprocess
    import sys
    for rel in ("bench.py", "__graft_entry__.py", "tools/multi_gpu_check.py", "tools/bench_configs.py",
                "tools/time_armed_head.py", "tools/time_head_select.py", "tools/summarize_profile.py"):
        py_compile.compile(os.path.join(ROOT, rel), doraise=True)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, check=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--positions", "--coverage", "--nccl-heads"):
        assert flag in out
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.workload_config(1)
    assert cfg["positions_per_gpu"] == 4_600_000 and cfg["coverage"] == [100, 100] and "workload" in cfg
    assert "peer" in bench.workload_config(8)["parallelism"]
