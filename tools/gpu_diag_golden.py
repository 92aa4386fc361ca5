#!/usr/bin/env python3
"""Diagnostics on the reference-made golden cases: max relative error of every GPU column against
the fixture, and the first ranking differences with their keys."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nanomod_b200 as nm  # noqa: E402
import golden_ref as G  # noqa: E402

det = nm.Detector(0)
for name in G.CASE_NAMES:
    case, p, opt = G.REF["cases"][name], G.case_pileup(name), G.case_options(name)
    t = det.detect(p, opt)
    want = G.stats_columns(name)
    cols = [("u_stat", 0), ("u_p", 1), ("t_stat", 2), ("t_p", 3), ("ks_d", 4), ("ks_p", 5)]
    if opt.testMethod != "ks":
        cols += [(opt.testMethod + "_stat", 6), (opt.testMethod + "_p", 7)]
    errs = []
    for attr, c in cols:
        g, w = np.asarray(getattr(t, attr), np.float64), want[:, c]
        fin = np.isfinite(w) & (w != 0)
        with np.errstate(invalid="ignore", divide="ignore"):
            e = np.abs(g[fin] - w[fin]) / np.abs(w[fin])
        errs.append("%s %.1e" % (attr, e.max() if e.size else 0.0))
    print(name, " ".join(errs))
    got = [int(r) for r in t.sorted_rows()]
    if got != case["sorted"]:
        k = next(i for i, (a, b) in enumerate(zip(got, case["sorted"])) if a != b)
        print("   first ranking difference at", k, "got", got[k:k + 4], "want", case["sorted"][k:k + 4])
        col = 7 if opt.testMethod != "ks" else 5
        for r in sorted(set(got[k:k + 3] + case["sorted"][k:k + 3])):
            lo, hi = max(0, r - 5), min(len(want), r + 6)
            gp = getattr(t, (opt.testMethod if opt.testMethod != "ks" else "ks") + "_p")
            print("   row", r, "pos", case["rows"][r][2], "fixture p[r-5..r+5]", ["%.17g" % x for x in want[lo:hi, col]])
            print("   row", r, "gpu     p[r-5..r+5]", ["%.17g" % x for x in gp[lo:hi]])
