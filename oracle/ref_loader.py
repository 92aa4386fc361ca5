"""Import and drive the reference's own modules from ``oracle/_ref/``.  TEST INFRASTRUCTURE ONLY.

``oracle/build_ref.py`` renders `/root/reference/bin/scripts/{myCom,myFast5,myDetect}.py` to
Python 3 in ``oracle/_ref/``.  This loader imports that rendering with
  * ``h5py``, ``rpy2`` and ``pkg_resources`` stubbed (HDF5 files are served from memory, the R
    plotting calls do nothing -- the called-site SELECTION of mboxplot/plot1 still runs);
  * ``mannwhitneyu / ttest_ind / ks_2samp / combine_pvalues`` rebound to the scipy-1.2.1
    semantics of ``oracle/scipy_legacy.py`` (the reference pins scipy 1.2.1).
Everything else that runs is the reference's own text: ``mReadSignalBase``, ``mfilter_coverage``,
``getKStest``, ``pos_check``, ``get_combin_pvalue``, ``mtest2`` (row order, ranking, region mode),
``save_test``, ``mboxplot`` / ``plot1``.

Only ``tests/``, ``tests/golden/make_ref_golden.py``, ``__graft_entry__`` and ``bench.py``'s CPU
legs may import this module.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types
from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import build_ref, scipy_legacy

_CACHE: Optional[types.SimpleNamespace] = None


class RefUnavailable(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------
# stubs
# ------------------------------------------------------------------------------------------
class _Anything:
    """accepts any attribute / call / item access (R objects, ggplot imports)"""

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __getitem__(self, key):
        return _Anything()

    def __setitem__(self, key, value):
        pass


class _FakeDataset:
    def __init__(self, value=None, attrs=None):
        self.value = value  # h5py < 3 spelling used by myFast5.ReadNanoraw_events
        self.attrs = attrs or {}

    def __getitem__(self, key):
        return self.value


class FakeFast5(dict):
    """In-memory stand-in for ``h5py.File``: maps HDF5 paths to ``_FakeDataset``."""

    def __contains__(self, path):
        return dict.__contains__(self, path)


_FAKE_FILES: Dict[str, FakeFast5] = {}


def _fake_h5py_module():
    m = types.ModuleType("h5py")

    def File(fn, mode="r"):
        if fn not in _FAKE_FILES:
            raise IOError("no such fake fast5: %s" % fn)
        return _FAKE_FILES[fn]

    m.File = File
    return m


def _stub_modules() -> Dict[str, types.ModuleType]:
    rpy2 = types.ModuleType("rpy2")
    robjects_mod = types.ModuleType("rpy2.robjects")
    for name in ("r", "StrVector", "FloatVector", "IntVector", "FactorVector", "DataFrame"):
        setattr(robjects_mod, name, _Anything())
    robjects_mod.globalenv = _Anything()
    packages = types.ModuleType("rpy2.robjects.packages")
    packages.importr = lambda *a, **k: _Anything()
    rpy2.robjects = robjects_mod
    robjects_mod.packages = packages
    pkg = types.ModuleType("pkg_resources")
    pkg.resource_string = lambda *a, **k: b""
    return {"h5py": _fake_h5py_module(), "rpy2": rpy2, "rpy2.robjects": robjects_mod,
            "rpy2.robjects.packages": packages, "pkg_resources": pkg}


def available() -> bool:
    return build_ref.built() or build_ref.reference_available()


def load() -> types.SimpleNamespace:
    """the rendered reference modules: ns.myCom, ns.myFast5, ns.myDetect"""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    d = build_ref.build()
    if d is None:
        raise RefUnavailable("oracle/_ref is not built and %s is absent" % build_ref.REF_ROOT)
    saved = {k: sys.modules.get(k) for k in
             ("h5py", "rpy2", "rpy2.robjects", "rpy2.robjects.packages", "pkg_resources", "myCom", "myFast5")}
    mods = {}
    try:
        sys.modules.update(_stub_modules())
        for name in ("myCom", "myFast5", "myDetect"):
            spec = importlib.util.spec_from_file_location("_nanomod_ref_" + name, os.path.join(d, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            if name != "myDetect":
                sys.modules[name] = mod  # `from myCom import *`, `import myFast5` inside the reference
            with contextlib.redirect_stdout(io.StringIO()):
                spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    md = mods["myDetect"]
    md.mannwhitneyu = scipy_legacy.mannwhitneyu
    md.ttest_ind = scipy_legacy.ttest_ind
    md.ks_2samp = scipy_legacy.ks_2samp
    md.combine_pvalues = scipy_legacy.combine_pvalues
    _CACHE = types.SimpleNamespace(**mods)
    return _CACHE


# ------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------
def default_moptions(**over) -> Dict:
    """the options ``detect`` reads (defaults of NanoMod.py:347-392 after mCommonParam/detect parsing)"""
    mo = {"outLevel": 3, "MinCoverage": 5, "coverages": [0, 0], "downsampling": 100, "downsampling_quantile": 0.25,
          "neighborPvalues": 2, "WeightsDif": 2.0, "testMethod": "stouffer", "rankUse": "pv", "mstd": 0,
          "SaveTest": 0, "outFolder": ".", "FileID": "mod", "RegionRankbyST": 0, "window": 10, "percentile": 0.1,
          "WindOvlp": 0, "NA": "", "topN": 30, "plotType": "Density", "min_lr": 500, "min_lr_nb": 0,
          "ds2": ["g0", "g1"]}
    mo.update(over)
    return mo


def moptions_from_groups(groups: Sequence[Dict], bases: Dict, **over) -> Dict:
    """groups[g][(chrom, strand)][pos] -> list of values; bases[(chrom, strand)][pos] -> 'A'.."""
    mo = default_moptions(**over)
    for name, grp in zip(mo["ds2"], groups):
        nm = defaultdict(lambda: defaultdict(list))
        bs = defaultdict(lambda: defaultdict(str))
        bd = defaultdict(lambda: defaultdict(lambda: defaultdict(int)))
        for sk, posd in grp.items():
            for pk, vals in posd.items():
                nm[sk][pk] = [float(v) for v in vals]
                b = bases[sk][pk] if not isinstance(bases[sk][pk], (tuple, list)) else bases[sk][pk][mo["ds2"].index(name)]
                bs[sk][pk] = b
                bd[sk][pk][b] += len(vals)
        mo[name] = {"norm_mean": nm, "base": bs, "basedict": bd}
    return mo


def run_detect(mo: Dict) -> Dict:
    """``mfilter_coverage(mo); mtest2(mo)`` -- the reference's own call pair (myDetect.py:639-641)"""
    md = load().myDetect
    with contextlib.redirect_stdout(io.StringIO()):
        md.mfilter_coverage(mo)
        md.mtest2(mo)
    return mo


def called_sites(mo: Dict) -> List[Tuple[str, str, int]]:
    """sites the reference would plot: run ``mboxplot`` (myDetect.py:258-299) with the R calls
    stubbed and record every ranked row for which ``plot1`` (:130-256) reports enough neighbours."""
    md = load().myDetect
    accepted: List[Tuple[str, str, int]] = []
    orig = md.plot1

    def recording_plot1(moptions, significant_pos, curn):
        noenough = orig(moptions, significant_pos, curn)
        if not noenough:
            accepted.append((significant_pos[0][0], significant_pos[0][1], significant_pos[0][2]))
        return noenough

    md.plot1 = recording_plot1
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            md.mboxplot(mo)
    finally:
        md.plot1 = orig
    return accepted


def save_test_text(mo: Dict, tmpdir: str) -> Tuple[str, Optional[str]]:
    """text of ``<FileID>_sign_test.txt`` (and ``_meanstd.cvs`` with --mstd) as ``save_test`` writes it"""
    md = load().myDetect
    mo2 = dict(mo)
    mo2["SaveTest"] = 1
    mo2["outFolder"] = tmpdir
    with contextlib.redirect_stdout(io.StringIO()):
        md.save_test(mo2)
    with open(os.path.join(tmpdir, mo2["FileID"] + "_sign_test.txt")) as fh:
        text = fh.read()
    mstd = None
    if mo2["mstd"] != 0:
        with open(os.path.join(tmpdir, mo2["FileID"] + "_meanstd.cvs")) as fh:
            mstd = fh.read()
    return text, mstd


def read_reads(mo: Dict, group_index: int, reads: Sequence[Dict]) -> None:
    """feed in-memory reads through the reference's ``mReadSignalBase`` (myDetect.py:33-127).
    Each read: {'chrom','strand','start','norm_mean','base'}."""
    ns = load()
    md, f5 = ns.myDetect, ns.myFast5
    name = mo["ds2"][group_index]
    mo["cur_wrkBase"] = name
    if name not in mo:
        mo[name] = {"base": defaultdict(lambda: defaultdict(str)),
                    "norm_mean": defaultdict(lambda: defaultdict(list)),
                    "basedict": defaultdict(lambda: defaultdict(lambda: defaultdict(int)))}
    real_os, real_h5py = md.os, md.h5py
    md.h5py = _fake_h5py_module()
    try:
        md.os = types.SimpleNamespace(path=types.SimpleNamespace(isfile=lambda fn: fn in _FAKE_FILES))
        for k, rd in enumerate(reads):
            fn = "fake://%s/%d.fast5" % (name, k)
            ev = np.zeros(len(rd["norm_mean"]), dtype=[("norm_mean", "f8"), ("base", "U1")])
            ev["norm_mean"] = rd["norm_mean"]
            ev["base"] = list(rd["base"])
            _FAKE_FILES[fn] = FakeFast5({
                f5.rawAlignment_full: _FakeDataset(attrs={ns.myCom.map_chr_str: rd["chrom"],
                                                          ns.myCom.map_start_str: rd["start"],
                                                          ns.myCom.map_strand_str: rd["strand"]}),
                f5.raw_event_ful: _FakeDataset(value=ev)})
            mo["fast5filename"] = fn
            with contextlib.redirect_stdout(io.StringIO()):
                md.mReadSignalBase(mo)
            del _FAKE_FILES[fn]
    finally:
        md.os, md.h5py = real_os, real_h5py


def set_tests(all_tests: bool) -> None:
    """bench.py's like-for-like switch: with ``all_tests=False`` the reference's calls of
    ``mannwhitneyu`` and ``ttest_ind`` (myDetect.py:331,335) return at once, so that its driver code
    does the work of a GPU run with want_u = want_t = False (KS test + combination).  The
    reference itself has no such option -- it always computes all three."""
    md = load().myDetect
    if all_tests:
        md.mannwhitneyu = scipy_legacy.mannwhitneyu
        md.ttest_ind = scipy_legacy.ttest_ind
    else:
        md.mannwhitneyu = lambda a, b: (0.0, 1.0)
        md.ttest_ind = lambda a, b, equal_var=False: (0.0, 1.0)
