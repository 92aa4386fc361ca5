#!/usr/bin/env python3
"""Generate tests/golden/golden.json: input/output vectors for the per-position path.

The reference itself cannot run here (Python 2 + h5py + rpy2) and ships no
fixtures, so the golden OUTPUTS are produced by the scipy that IS importable in this container
(1.18.1) on the code paths that are algebraically identical to the pinned scipy 1.2.1:
  U p   = stats.mannwhitneyu(a, b, use_continuity=True, method='asymptotic').pvalue / 2
  t, pt = stats.ttest_ind(a, b, equal_var=False)
  KS p  = stats.kstwobign.sf((en + 0.12 + 0.11/en) * D), D from the definition (searchsorted)
  comb  = stats.combine_pvalues(p[, 'stouffer', weights])
i.e. by calling scipy.stats directly -- NOT by calling oracle/.  tests/test_oracle.py then
requires the oracle to reproduce these numbers, and the GPU tests require the CUDA path to.
The K1..K4 / Stouffer / Fisher / edge vectors of SURVEY.md section 8c are included verbatim.

Run: python tests/golden/make_golden.py     (rewrites golden.json next to this file)
"""
import json
import os

import numpy as np
import scipy.special as sc
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))


def one_case(a, b):
    a64 = np.asarray(a, np.float32).astype(np.float64)
    b64 = np.asarray(b, np.float32).astype(np.float64)
    n0, n1 = len(a64), len(b64)
    pooled = np.concatenate([a64, b64])
    c0 = np.searchsorted(np.sort(a64), pooled, side="right")
    c1 = np.searchsorted(np.sort(b64), pooled, side="right")
    dnum = int(np.max(np.abs(c0 * n1 - c1 * n0)))
    D = float(np.max(np.abs(c0 / n0 - c1 / n1)))
    en = np.sqrt(n0 * n1 / float(n0 + n1))
    pks = float(st.kstwobign.sf((en + 0.12 + 0.11 / en) * D))
    mw = st.mannwhitneyu(a64, b64, use_continuity=True, alternative="two-sided", method="asymptotic")
    u1 = float(mw.statistic)
    two_u = int(round(2 * min(u1, n0 * n1 - u1)))
    tt = st.ttest_ind(a64, b64, equal_var=False)
    return {"a": [float(x) for x in np.asarray(a, np.float32)], "b": [float(x) for x in np.asarray(b, np.float32)],
            "dnum": dnum, "D": D, "pks": pks, "twoU": two_u, "pu": float(mw.pvalue) / 2,
            "t": float(tt.statistic), "pt": float(tt.pvalue)}


def main():
    rng = np.random.Generator(np.random.PCG64(7))
    cases = []
    for n0, n1, shift, dec in [(5, 5, 0.0, None), (5, 9, 1.0, 1), (12, 7, 0.5, 3), (30, 30, 0.0, None),
                               (30, 30, 1.0, 3), (50, 50, 1.0, 3), (100, 100, 0.0, None),
                               (100, 100, 0.25, 3), (100, 100, 1.0, 2), (128, 128, 0.5, 1),
                               (97, 128, 2.0, None), (160, 40, 0.3, 3), (300, 200, 0.2, 2),
                               (2000, 2000, 0.1, 3), (2000, 1500, 4.0, None)]:
        a = rng.normal(0, 1, n0)
        b = rng.normal(shift, 1, n1)
        if dec is not None:
            a, b = np.round(a, dec), np.round(b, dec)
        cases.append(one_case(a, b))
    combos = []
    for nb, wd in [(2, 2.0), (3, 2.0), (1, 1.0), (4, 1.5)]:
        k = 2 * nb + 1
        for _ in range(3):
            p = np.clip(10.0 ** (-rng.uniform(0, 12, k)), 2.2250738585072014e-308, 1.0)
            w = [100.0]
            for _i in range(nb):
                w.insert(0, w[0] / wd)
                w.append(w[-1] / wd)
            s = st.combine_pvalues(p, method="stouffer", weights=w)
            f = st.combine_pvalues(p, method="fisher")
            combos.append({"nb": nb, "WeightsDif": wd, "p": [float(x) for x in p],
                           "stouffer": [float(s.statistic), float(s.pvalue)],
                           "fisher": [float(f.statistic), float(f.pvalue)]})
    survey = {
        "K1": {"a": [.1, .2, .3, .4, .5], "b": [.35, .45, .55, .65, .75], "dnum": 15, "pks": 0.2089848305751669,
               "twoU": 6, "pu": 0.030051402969433157, "t": -2.5000000000000004, "pt": 0.03694203771362409},
        "K2": {"a": [.1, .2, .2, .3, .3, .3], "b": [.2, .3, .3, .4, .4, .5, .6], "dnum": 24,
               "pks": 0.15504417912365295, "twoU": 14, "pu": 0.022835621469841673, "t": -2.506433439961463,
               "pt": 0.031015852000315595},
        "K3": {"a": list(np.arange(10) / 10), "b": list(2 + np.arange(12) / 10), "dnum": 120,
               "pks": 7.26220915473708e-06, "twoU": 0, "pu": 4.3669632941789134e-05, "t": -14.849242404917499,
               "pt": 2.9120185192417662e-12},
        "K4": {"seed": 1, "dnum": 1450, "pks": 3.761754930361375e-08, "twoU": 888, "pu": 1.4042661075413371e-08,
               "t": -6.586679477347367, "pt": 2.5818774981476798e-09},
        "stouffer5": {"p": [0.5, 0.04, 1e-3, 1e-8, 2e-2], "w": [25, 50, 100, 50, 25], "Z": 5.714834526221459,
                      "pz": 5.490541316467834e-09, "X2": 66.3049640675816, "px": 2.279584895841645e-10},
        "stouffer7": {"p": [0.5, 0.04, 1e-3, 1e-8, 2e-2, 0.3, 0.9], "w": [12.5, 25, 50, 100, 50, 25, 12.5],
                      "Z": 6.67667306242052, "pz": 1.2221369460069533e-11, "X2": 68.92363070754912,
                      "px": 3.0215524820249258e-09},
        "edges": {"isf_min": 37.5193793471445, "kolmogorov_7.2": 1.8766276545661244e-45,
                  "stdtr_198_-40": 4.1953717271829827e-97},
    }
    special = {
        "kolmogorov": [[float(x), float(sc.kolmogorov(x))] for x in
                       [0.05, 0.17, 0.18, 0.2, 0.3, 0.5, 0.8, 0.82, 0.83, 1.0, 1.5, 2.0, 3.0, 5.0, 7.2, 10.0, 15.0, 19.0, 26.0]],
        "t_two_sided": [[float(t), float(df), float(2 * sc.stdtr(df, -t))] for df in [2.0, 8.0, 10.04, 93.8, 198.0, 3998.0]
                        for t in [0.5, 1.0, 2.0, 5.0, 14.8, 40.0]],
        "chdtrc": [[float(x), int(k), float(sc.chdtrc(2 * k, x))] for k in [3, 5, 7] for x in [0.5, 10.0, 66.3, 500.0, 1300.0]],
        "ndtr": [[float(z), float(sc.ndtr(-z))] for z in [-3.0, 0.0, 1.0, 5.7, 20.0, 37.0]],
        "ndtri": [[float(p), float(-sc.ndtri(p))] for p in [2.2250738585072014e-308, 1e-100, 1e-8, 0.02, 0.3, 0.9]],
    }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"scipy": __import__("scipy").__version__, "cases": cases, "combos": combos,
                   "survey": survey, "special": special}, f)
    print("wrote golden.json:", len(cases), "cases,", len(combos), "combos")


if __name__ == "__main__":
    main()
