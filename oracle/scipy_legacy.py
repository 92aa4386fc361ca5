"""scipy==1.2.1 call semantics of the four ``scipy.stats`` functions the reference calls.
TEST INFRASTRUCTURE ONLY (imported by ``oracle/ref_loader.py``).

The reference pins scipy 1.2.1 (env.py27nanomod.yml:99) and calls (myDetect.py:331,335,341,393,401)
``mannwhitneyu(a, b)``, ``ttest_ind(a, b, equal_var=False)``, ``ks_2samp(a, b)`` and
``combine_pvalues(p[, method='stouffer', weights=w])``.  scipy 1.2.1 itself is not installable
here (no network, Python 2 era), and the scipy that IS here (1.18) changed the defaults of two of
them (``ks_2samp`` -> exact / kstwo p-values, ``mannwhitneyu`` -> two-sided, U of the first
sample).  These shims restate the 1.2.1 bodies through ``scipy.stats.distributions`` objects
(``norm``, ``t``, ``chi2``, ``kstwobign``) -- deliberately a different route from
``oracle/nanomod_oracle.py``, which goes through ``scipy.special`` primitives -- so that running
the reference's own driver code on top of them is a second, independent statement of the path.

THIS is what remains unpinned: the 1.2.1 function bodies are transcribed from the published
release, not executed.  tests/test_oracle.py cross-checks them against modern scipy on the code
paths that are algebraically unchanged.
"""
from __future__ import annotations

import numpy as np
from scipy.stats import distributions, rankdata, tiecorrect


def mannwhitneyu(x, y, use_continuity=True, alternative=None):
    """scipy 1.2.1 stats.py ``mannwhitneyu`` with the deprecated default ``alternative=None``:
    returns (min(u1, u2), one-sided p)."""
    x = np.asarray(x)
    y = np.asarray(y)
    n1 = len(x)
    n2 = len(y)
    ranked = rankdata(np.concatenate((x, y)))
    rankx = ranked[0:n1]
    u1 = n1 * n2 + (n1 * (n1 + 1)) / 2.0 - np.sum(rankx, axis=0)
    u2 = n1 * n2 - u1
    T = tiecorrect(ranked)
    if T == 0:
        raise ValueError("All numbers are identical in mannwhitneyu")
    sd = np.sqrt(T * n1 * n2 * (n1 + n2 + 1) / 12.0)
    meanrank = n1 * n2 / 2.0 + 0.5 * use_continuity
    if alternative is None or alternative == "two-sided":
        bigu = max(u1, u2)
    elif alternative == "less":
        bigu = u1
    elif alternative == "greater":
        bigu = u2
    else:
        raise ValueError("alternative should be None, 'less', 'greater' or 'two-sided'")
    z = (bigu - meanrank) / sd
    if alternative is None:
        p = distributions.norm.sf(abs(z))
    elif alternative == "two-sided":
        p = 2 * distributions.norm.sf(abs(z))
    else:
        p = distributions.norm.sf(z)
    u = u2
    if alternative is None:
        u = min(u1, u2)
    return u, p


def ttest_ind(a, b, axis=0, equal_var=True):
    """scipy 1.2.1 ``ttest_ind`` (1-D inputs; Welch when ``equal_var=False``)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    v1 = np.var(a, axis, ddof=1)
    v2 = np.var(b, axis, ddof=1)
    n1 = a.shape[axis]
    n2 = b.shape[axis]
    if equal_var:
        df = n1 + n2 - 2.0
        svar = ((n1 - 1) * v1 + (n2 - 1) * v2) / df
        denom = np.sqrt(svar * (1.0 / n1 + 1.0 / n2))
    else:
        vn1 = v1 / n1
        vn2 = v2 / n2
        with np.errstate(divide="ignore", invalid="ignore"):
            df = (vn1 + vn2) ** 2 / (vn1 ** 2 / (n1 - 1) + vn2 ** 2 / (n2 - 1))
        df = np.where(np.isnan(df), 1, df)
        denom = np.sqrt(vn1 + vn2)
    d = np.mean(a, axis) - np.mean(b, axis)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.divide(d, denom)
    prob = distributions.t.sf(np.abs(t), df) * 2
    return t[()] if isinstance(t, np.ndarray) else t, prob[()] if isinstance(prob, np.ndarray) else prob


def ks_2samp(data1, data2):
    """scipy 1.2.1 ``ks_2samp``: asymptotic Kolmogorov p with the Stephens correction."""
    data1 = np.sort(data1)
    data2 = np.sort(data2)
    n1 = data1.shape[0]
    n2 = data2.shape[0]
    data_all = np.concatenate([data1, data2])
    cdf1 = np.searchsorted(data1, data_all, side="right") / (1.0 * n1)
    cdf2 = np.searchsorted(data2, data_all, side="right") / (1.0 * n2)
    d = np.max(np.absolute(cdf1 - cdf2))
    en = np.sqrt(n1 * n2 / float(n1 + n2))
    try:
        prob = distributions.kstwobign.sf((en + 0.12 + 0.11 / en) * d)
    except Exception:
        prob = 1.0
    return d, prob


def combine_pvalues(pvalues, method="fisher", weights=None):
    """scipy 1.2.1 ``combine_pvalues`` (Fisher / weighted Stouffer)."""
    pvalues = np.asarray(pvalues)
    if pvalues.ndim != 1:
        raise ValueError("pvalues is not 1-D")
    if method == "fisher":
        Xsq = -2 * np.sum(np.log(pvalues))
        pval = distributions.chi2.sf(Xsq, 2 * len(pvalues))
        return (Xsq, pval)
    elif method == "stouffer":
        if weights is None:
            weights = np.ones_like(pvalues)
        elif len(weights) != len(pvalues):
            raise ValueError("pvalues and weights must be of the same size.")
        weights = np.asarray(weights)
        if weights.ndim != 1:
            raise ValueError("weights is not 1-D")
        with np.errstate(invalid="ignore"):
            Zi = distributions.norm.isf(pvalues)
            Z = np.dot(weights, Zi) / np.linalg.norm(weights)
        pval = distributions.norm.sf(Z)
        return (Z, pval)
    raise ValueError("Invalid method '%s'. Options are 'fisher' or 'stouffer'" % method)
