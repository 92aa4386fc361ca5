"""Randomised GPU parity: pileups with random shapes (coverage distributions, gaps, outliers, deep
rows, strands, tie density) and random options, against the vectorised oracle -- same bar as
tests/test_gpu_parity.py.  Aimed at the code that picks a path per call or per tile: network
size, size-group launches, contiguous / span / row-by-row staging, tier boundaries."""
import numpy as np
import pytest

import nanomod_b200 as nm
from test_gpu_parity import assert_table_matches, vec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det():
    return nm.Detector(0)


def random_pileup(rng):
    # the large sizes give every warp of the persistent lane kernel several tiles (tile queue,
    # metadata pipeline, copies issued a tile ahead)
    L = int(rng.choice([1, 31, 32, 33, 200, 777, 2048, 5000, 60000, 130000]))
    shape = rng.choice(["uniform", "poisson", "bimodal", "outliers", "ramp", "wide"])
    mean = int(rng.choice([6, 17, 30, 52, 64, 65, 90, 100, 104, 105, 120, 128]))
    if L > 10000:
        mean = min(mean, int(rng.choice([30, 66, 104])))
    if shape == "uniform":
        c0 = np.full(L, mean)
        c1 = np.full(L, max(5, mean - int(rng.integers(0, 4))))
    elif shape == "poisson":
        c0, c1 = rng.poisson(mean, L), rng.poisson(mean, L)
    elif shape == "bimodal":
        hi = rng.random(L) < rng.choice([0.02, 0.3, 0.7])
        c0 = np.where(hi, mean, max(5, mean // 3))
        c1 = np.where(hi, mean, max(5, mean // 3 + 1))
    elif shape == "outliers":
        c0 = np.full(L, min(mean, 100))
        c1 = np.full(L, min(mean, 100))
        out = rng.random(L) < 0.03
        c0[out] = rng.integers(101, 129, out.sum())
        c1[out] = rng.integers(5, 129, out.sum())
    elif shape == "ramp":
        c0 = np.linspace(5, mean, L).astype(np.int64) + rng.integers(0, 3, L)
        c1 = np.linspace(mean, 5, L).astype(np.int64) + rng.integers(0, 3, L)
    else:
        c0, c1 = rng.integers(1, 141, L), rng.integers(1, 141, L)
    c0 = np.clip(c0, 0, 140).astype(np.int64)
    c1 = np.clip(c1, 0, 140).astype(np.int64)
    if rng.random() < 0.5:  # positions lost to the coverage filter
        drop = rng.random(L) < rng.choice([0.01, 0.2])
        c0[drop] = rng.integers(0, 3, drop.sum())
    if rng.random() < 0.4 and L > 40:  # a few deep rows
        for i in rng.choice(L, 3, replace=False):
            c0[i], c1[i] = int(rng.integers(129, 900)), int(rng.integers(5, 900))
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    dec = int(rng.choice([1, 2, 3, 6]))
    v0 = np.round(rng.normal(0, 1, off0[-1]), dec).astype(np.float32)
    v1 = np.round(rng.normal(rng.choice([0.0, 0.3]), 1, off1[-1]), dec).astype(np.float32)
    if rng.random() < 0.3:
        v0[rng.random(v0.shape[0]) < 0.05] = -0.0
        v1[rng.random(v1.shape[0]) < 0.05] = 0.0
    pos = np.arange(L, dtype=np.int32)
    seg = None
    names = None
    if rng.random() < 0.5 and L > 4:
        cut = int(rng.integers(1, L))
        seg = (np.arange(L) >= cut).astype(np.int32)
        pos = np.concatenate([np.arange(cut), np.arange(L - cut)]).astype(np.int32)
        names = [("chr1", "+"), ("chr1", "-")]
    if rng.random() < 0.3 and L > 10:  # holes in the coordinate: the window rule must see them
        pos = pos + np.cumsum(rng.random(L) < 0.02).astype(np.int32) * (1 if seg is None else 0)
    return nm.Pileup.from_arrays(v0, off0, v1, off1, pos, seg, seg_names=names)


@pytest.mark.parametrize("seed", range(120))
def test_random_pileups(det, seed):
    rng = np.random.default_rng(1000 + seed)
    p = random_pileup(rng)
    full = bool(rng.random() < 0.5)
    opt = nm.DetectOptions(MinCoverage=int(rng.choice([3, 5, 8])), neighborPvalues=int(rng.choice([0, 1, 2, 3, 5])),
                           WeightsDif=float(rng.choice([1.0, 2.0, 3.5])), both_combinations=True,
                           want_u=full, want_t=full, mstd=full)
    t = det.detect(p, opt)
    res = vec(p, opt)
    if "dnum" not in res:  # nothing passes the coverage filter
        assert len(t) == 0 and len(res["row_pos_index"]) == 0
        return
    assert_table_matches(t, res, opt)
    if full:
        m0, s0, m1, s1 = t.mean_std()
        for r in range(0, len(t), max(1, len(t) // 25)):
            i = int(t.row_pos_index[r])
            a, b = p.group(0, i).astype(np.float64), p.group(1, i).astype(np.float64)
            assert abs(m0[r] - a.mean()) <= 1e-12 * max(1.0, abs(a.mean())) and abs(s1[r] - b.std()) <= 1e-12 * max(1.0, b.std())
