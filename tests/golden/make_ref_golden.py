#!/usr/bin/env python3
"""Generate tests/golden/ref_cases.{npz,json}: golden vectors produced BY THE REFERENCE ITSELF.

`oracle/build_ref.py` renders the reference's `bin/scripts/myDetect.py` to Python 3 (three
mechanical rewrites) and `oracle/ref_loader.py` runs its own `mfilter_coverage`, `mtest2`
(getKStest, pos_check, get_combin_pvalue, ranking, region mode), `save_test` and the called-site
selection of `mboxplot`/`plot1` on seeded inputs, with the four scipy calls bound to the pinned
scipy-1.2.1 semantics (`oracle/scipy_legacy.py`).  The reference tree does not exist on the GPU
box, so inputs and outputs are committed here as small fixtures:

  ref_cases.npz   inputs  : <case>/vals0, off0, vals1, off1, pos, seg, base (CSR pileup, float32)
  ref_cases.json  outputs : per case the options, segment names, the `sign_test` rows, the order of
                            `sorted_sign_test`, the called sites, the text `save_test` writes

Consumers: tests/test_oracle_vs_ref.py (oracle == fixtures, CPU; and live reference == fixtures
when /root/reference is present) and tests/test_gpu_golden.py (CUDA through the C ABI == fixtures).

Run (in the build container, where /root/reference exists):  python tests/golden/make_ref_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402

SEED = 20190131


def make_inputs(name, length, n0, n1, *, decimals=3, poisson=False, clip=(3, 60), drop=0.0, two_strands=False,
                shift_sites=(), rng=None):
    """seeded pileup on the reference's 0.001 grid (norm_mean = round(x, 3),
    myRefBaseSignalAnnotation.py:1108), stored as float32"""
    pos = np.arange(length, dtype=np.int32)
    if poisson:
        c0 = np.clip(rng.poisson(n0, length), clip[0], clip[1]).astype(np.int64)
        c1 = np.clip(rng.poisson(n1, length), clip[0], clip[1]).astype(np.int64)
    else:
        c0 = np.full(length, n0, np.int64)
        c1 = np.full(length, n1, np.int64)
    if drop > 0:
        c1[rng.random(length) < drop] = 0
        c0[rng.random(length) < drop / 2] = 0
    shift = np.zeros(length)
    for s in shift_sites:
        for d, v in ((0, 1.0), (1, 0.5), (2, 0.25)):
            for q in {s - d, s + d}:
                if 0 <= q < length:
                    shift[q] = max(shift[q], v)
    off0 = np.concatenate([[0], np.cumsum(c0)]).astype(np.int64)
    off1 = np.concatenate([[0], np.cumsum(c1)]).astype(np.int64)
    v0 = np.round(rng.standard_normal(off0[-1]), decimals).astype(np.float32)
    v1 = np.round(rng.standard_normal(off1[-1]) + np.repeat(shift, c1), decimals).astype(np.float32)
    if two_strands:
        half = length // 2
        seg = (np.arange(length) >= half).astype(np.int32)
        pos = np.where(seg == 0, pos, pos - half + 7).astype(np.int32)  # second strand starts at 7
        names = [["chrS", "+"], ["chrS", "-"]]
    else:
        seg = np.zeros(length, np.int32)
        names = [["chrS", "+"]]
    base = np.frombuffer(b"ACGT", dtype=np.uint8)[(pos * 7 + 3) % 4].copy()
    return dict(vals0=v0, off0=off0, vals1=v1, off1=off1, pos=pos, seg=seg, base=base), names


def to_groups(inp, names):
    groups = [{}, {}]
    bases = {}
    for g, (vals, off) in enumerate(((inp["vals0"], inp["off0"]), (inp["vals1"], inp["off1"]))):
        for i in range(len(inp["pos"])):
            if off[i + 1] == off[i]:
                continue
            sk = tuple(names[inp["seg"][i]])
            groups[g].setdefault(sk, {})[int(inp["pos"][i])] = [float(x) for x in vals[off[i]:off[i + 1]]]
            bases.setdefault(sk, {})[int(inp["pos"][i])] = chr(inp["base"][i])
    return groups, bases


CASES = [
    # name, input kwargs, option overrides
    ("cfg1_small", dict(length=1200, n0=50, n1=50, shift_sites=(300, 700, 1100)),
     dict(neighborPvalues=3, testMethod="stouffer")),
    ("ties_1decimal", dict(length=400, n0=12, n1=15, decimals=1, shift_sites=(100, 250)),
     dict(neighborPvalues=2, testMethod="stouffer")),
    ("gaps_two_strands_fisher", dict(length=700, n0=18, n1=22, poisson=True, drop=0.04, two_strands=True,
                                     shift_sites=(120, 500)),
     dict(neighborPvalues=2, testMethod="fisher")),
    ("ks_only", dict(length=400, n0=25, n1=20, shift_sites=(200,)), dict(testMethod="ks")),
    ("rank_by_statistic", dict(length=400, n0=20, n1=20, shift_sites=(100, 300)),
     dict(neighborPvalues=2, testMethod="stouffer", rankUse="st")),
    ("nb0_fisher", dict(length=300, n0=16, n1=16, shift_sites=(150,)), dict(neighborPvalues=0, testMethod="fisher")),
    ("weights_nb4", dict(length=400, n0=30, n1=28, shift_sites=(77, 301)),
     dict(neighborPvalues=4, WeightsDif=1.5, testMethod="stouffer")),
    ("mincov3_sparse", dict(length=500, n0=8, n1=9, poisson=True, clip=(1, 20), shift_sites=(250,)),
     dict(MinCoverage=3, neighborPvalues=2, testMethod="stouffer")),
    ("region_rank", dict(length=600, n0=20, n1=20, shift_sites=(150, 420)),
     dict(neighborPvalues=2, testMethod="stouffer", RegionRankbyST=1, window=3, topN=5)),
    ("region_rank_ovlp_na", dict(length=900, n0=20, n1=20, drop=0.004, shift_sites=(150, 420, 700)),
     dict(neighborPvalues=2, testMethod="fisher", RegionRankbyST=1, window=14, WindOvlp=1, NA="A", percentile=0.2,
          topN=5)),
    ("mstd", dict(length=200, n0=14, n1=11, shift_sites=(100,)), dict(neighborPvalues=2, testMethod="stouffer", mstd=1)),
    ("long_rows_140", dict(length=120, n0=140, n1=131, shift_sites=(60,)),
     dict(neighborPvalues=2, testMethod="stouffer")),
]

# SURVEY.md section 8c known-answer inputs, pushed through the reference's getKStest
_r = np.random.RandomState(1)
KNOWN = {
    "K1": ([.1, .2, .3, .4, .5], [.35, .45, .55, .65, .75]),
    "K2": ([.1, .2, .2, .3, .3, .3], [.2, .3, .3, .4, .4, .5, .6]),
    "K3": (list(np.arange(10) / 10), list(2 + np.arange(12) / 10)),
    "K4": (list(np.round(_r.normal(0, 1, 50), 3)), list(np.round(_r.normal(1, 1, 50), 3))),
}


def main():
    rng = np.random.Generator(np.random.PCG64(SEED))
    arrays = {}
    out = {"generator": "tests/golden/make_ref_golden.py", "reference": "bin/scripts/myDetect.py via oracle/_ref",
           "scipy_semantics": "1.2.1 (oracle/scipy_legacy.py)", "cases": {}, "known": {}}
    md = rl.load().myDetect
    for name, ikw, okw in CASES:
        inp, names = make_inputs(name, rng=rng, **ikw)
        for k, v in inp.items():
            arrays["%s/%s" % (name, k)] = v
        groups, bases = to_groups(inp, names)
        mo = rl.moptions_from_groups(groups, bases, **okw)
        rl.run_detect(mo)
        st = mo["sign_test"]
        index_of = {m[0][:3]: i for i, m in enumerate(st)}
        with tempfile.TemporaryDirectory() as td:
            text, mstd_text = rl.save_test_text(mo, td)
        sites = rl.called_sites(mo)
        out["cases"][name] = {
            "options": okw, "seg_names": names,
            "rows": [[m[0][0], m[0][1], int(m[0][2]), m[0][3], int(m[0][4]), int(m[0][5])] for m in st],
            "stats": [[float(x) for tup in m[1] for x in tup] for m in st],  # U pU t pt D pks [comb_st comb_p]
            "sorted": [index_of[m[0][:3]] for m in mo["sorted_sign_test"]],
            "called_sites": [[c, s, int(p)] for c, s, p in sites],
            "sign_test_txt": text, "meanstd_cvs": mstd_text,
        }
        print("%-26s rows %5d  called %2d  sorted %5d" % (name, len(st), len(sites), len(mo["sorted_sign_test"])))
    for name, (a, b) in KNOWN.items():
        a32 = [float(np.float32(x)) for x in a]
        b32 = [float(np.float32(x)) for x in b]
        res = md.getKStest({"coverages": [0, 0]}, a32, b32, "+")
        out["known"][name] = {"a": a32, "b": b32, "result": [[float(x) for x in tup] for tup in res]}
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **arrays)
    with open(os.path.join(HERE, "ref_cases.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote ref_cases.npz (%d arrays) and ref_cases.json" % len(arrays))


if __name__ == "__main__":
    main()
