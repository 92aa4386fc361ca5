"""GPU tests of the lane tier's 16-bit grid-key path (nm_lane.cuh "Grid keys"), through the C ABI.

Values that are float32 images of three-place decimals -- what the reference stores (norm_mean =
round(x, 3), bin/scripts/myRefBaseSignalAnnotation.py:1108) -- are sorted as packed 16-bit keys,
both groups of a position in one network pass; every value is checked on the device and a tile with
any other value takes the float32 path.  The tests hold the path to three statements:
  * on grid data the tables are bit-identical to the float32 path's and match the oracle;
  * the path is actually taken (nm_last_grid_tiles) on grid data and never on other data;
  * values built to fool a weaker check (the float next to a grid value in the same position, values
    beyond the 16-bit range, -0.0, NaN-free extremes) change nothing: the oracle's numbers come out.
"""
import os

import numpy as np
import pytest

import nanomod_b200 as nm
from oracle import nanomod_oracle_vec as ov
from test_gpu_parity import assert_table_matches, vec, _tables_identical

pytestmark = pytest.mark.gpu


def _detector(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return nm.Detector(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def det():
    return nm.Detector(0)


@pytest.fixture(scope="module")
def det_float():
    return _detector(NANOMOD_B200_NO_GRID="1")


@pytest.fixture(scope="module")
def det_grid_u():
    """by default calls that want the rank statistics stay on the float32 kernel (its walk is the faster one);
    this detector takes the grid-key kernel for them too, so that its U / tie walk stays tested"""
    return _detector(NANOMOD_B200_GRID_U="1")


@pytest.fixture(scope="module")
def det_general():
    return _detector(NANOMOD_B200_NO_DENSE="1")


def n_tiles(p, t):
    return (len(t) + 31) // 32


@pytest.mark.parametrize("n", [9, 31, 50, 64, 65, 100, 103, 104, 112, 128])
def test_grid_path_taken_and_identical_to_float_path(det, det_grid_u, det_float, det_general, n):
    p = nm.synthetic_pileup(2500 + n, n, max(5, n - 3), round_decimals=3)
    for opt in (nm.DetectOptions(neighborPvalues=3, both_combinations=True),
                nm.DetectOptions(neighborPvalues=3, both_combinations=True, want_u=False),
                nm.DetectOptions(neighborPvalues=2, testMethod="stouffer", want_u=False, want_t=False)):
        t = det_grid_u.detect(p, opt)
        # rows of <= 64 reads stay on the float32 sort: there the check costs what the packed sort saves
        assert det_grid_u.handle.last_grid_tiles() == (n_tiles(p, t) if n > 64 else 0), "every tile of grid data takes the grid path"
        d = det.detect(p, opt)
        assert det.handle.last_grid_tiles() == (n_tiles(p, t) if n > 64 and not opt.want_u else 0)
        _tables_identical(t, d)
        f = det_float.detect(p, opt)
        assert det_float.handle.last_grid_tiles() == 0
        _tables_identical(t, f)
        g = det_general.detect(p, opt)  # the general path (filtered rows, size-group launches) sorts float32 only
        assert det_general.handle.last_path() == 0 and det_general.handle.last_grid_tiles() == 0
        _tables_identical(t, g)
        assert_table_matches(t, vec(p, opt), opt)


@pytest.mark.parametrize("kw", [dict(poisson=True, clip=(2, 128), round_decimals=3),
                                dict(poisson=True, clip=(5, 128), round_decimals=1, drop_frac1=0.02),
                                dict(round_decimals=0), dict(round_decimals=2, two_strands=True, drop_frac1=0.01)])
def test_grid_path_ragged_rows_gaps_heavy_ties(det_grid_u, det_float, kw):
    det = det_grid_u
    p = nm.synthetic_pileup(6000, 90, 77, seed=5, **kw)
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    t = det.detect(p, opt)
    # (gaps put a call on the general path, which sorts float32 only)
    assert (det.handle.last_grid_tiles() > 0) == (det.handle.last_path() not in (0, 4))
    _tables_identical(t, det_float.detect(p, opt))
    assert_table_matches(t, vec(p, opt), opt)


def test_off_grid_data_never_takes_the_grid_path(det):
    p = nm.synthetic_pileup(3000, 100, 100)  # float32 normals: essentially none is a three-place decimal
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer")
    t = det.detect(p, opt)
    assert det.handle.last_grid_tiles() == 0
    assert_table_matches(t, vec(p, opt), opt)


def _poison(p, rows, make):
    """copy of pileup p with values of the given positions replaced by make(row_values0, row_values1)"""
    v0, v1 = p.vals0.copy(), p.vals1.copy()
    for r in rows:
        a, b = v0[p.off0[r]:p.off0[r + 1]], v1[p.off1[r]:p.off1[r + 1]]
        make(a, b)
    return nm.Pileup(vals0=v0, off0=p.off0, vals1=v1, off1=p.off1, pos=p.pos, seg=p.seg, base=p.base,
                     seg_names=p.seg_names)


def test_values_built_to_fool_the_check(det_grid_u, det_float):
    det = det_grid_u
    base = nm.synthetic_pileup(3200, 100, 100, seed=3, round_decimals=3)
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    up = lambda x: np.nextafter(np.float32(x), np.float32(np.inf))
    dn = lambda x: np.nextafter(np.float32(x), np.float32(-np.inf))

    def neighbour_floats(a, b):   # the floats next to a grid value, next to that value itself: a false
        a[:] = np.float32(0.417)  # tie would change D, U and the tie correction
        b[:50] = up(0.417)
        b[50:] = dn(0.417)

    def near_ties(a, b):          # one group on the grid, the other one ulp away from each of its values
        b[:] = up(a[:len(b)])

    def out_of_range(a, b):       # on the grid but beyond 16 bits: 32.767 and up; a key that would wrap
        a[0], a[1], a[2], b[0], b[1] = 32.767, -32.767, 40.0, 65.536, -65.536

    def extremes(a, b):           # the largest keys the path accepts, both signs, and -0.0 / 0.0 ties
        a[0], a[1], a[2], a[3] = 32.766, -32.766, -0.0, 0.0
        b[0], b[1], b[2], b[3] = -32.766, 32.766, 0.0, -0.0

    def huge(a, b):               # values whose product with 1000 is not even an integer-valued float
        a[0], b[0], b[1] = 1e10, -3e38, 1e-30

    for k, make in enumerate((neighbour_floats, near_ties, out_of_range, extremes, huge)):
        rows = [7 + 32 * k, 1000 + k, 3199 - 32 * k]
        p = _poison(base, rows, make)
        t = det.detect(p, opt)
        tiles = det.handle.last_grid_tiles()
        if make is extremes:
            assert tiles == 100, "the extremes of the key range and -0.0 are grid values"
        else:
            assert 100 - len(rows) <= tiles < 100, (make.__name__, tiles)
        _tables_identical(t, det_float.detect(p, opt))
        assert_table_matches(t, vec(p, opt), opt)


def test_off_grid_call_gives_up_and_later_calls_skip_the_attempt(det_float):
    """Data that are not three-place decimals: every warp of the grid-key launch fails its first tiles, the
    give-up flag stops the launch, the float32 launch does the work.  The handle then skips the attempt for
    NM_GRID_SKIP_CALLS calls and tries again afterwards.  Tables are the float path's throughout."""
    fresh = nm.Detector(0)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False)
    big = nm.synthetic_pileup(80000, 100, 100, seed=4)          # 2500 tiles for 1184 resident warps
    t = fresh.detect(big, opt)
    assert fresh.handle.last_grid_tiles() == 0
    _tables_identical(t, det_float.detect(big, opt))
    small = nm.synthetic_pileup(3200, 100, 100, seed=5, round_decimals=3)
    want = det_float.detect(small, opt)
    seen = []
    for _ in range(17):
        _tables_identical(fresh.detect(small, opt), want)
        seen.append(fresh.handle.last_grid_tiles())
    assert seen[:15] == [0] * 15 and seen[15:] == [100, 100], seen


def test_half_off_grid_call(det):
    """Tiles that are on the grid take the grid path, the others the float32 path, in one call."""
    p = nm.synthetic_pileup(19200, 100, 100, seed=9)
    v0, v1 = p.vals0.copy(), p.vals1.copy()
    v0[: 100 * 9600] = np.round(v0[: 100 * 9600].astype(np.float64), 3).astype(np.float32)
    v1[: 100 * 9600] = np.round(v1[: 100 * 9600].astype(np.float64), 3).astype(np.float32)
    q = nm.Pileup(vals0=v0, off0=p.off0, vals1=v1, off1=p.off1, pos=p.pos, seg=p.seg, base=p.base, seg_names=p.seg_names)
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False)
    t = det.detect(q, opt)
    assert det.handle.last_grid_tiles() == 300
    assert_table_matches(t, vec(q, opt), opt)


def test_deep_tier_grid_keys(det, det_float):
    """Deep rows (block-per-position tier): positions whose values are three-place decimals sort ONE array of
    16-bit key pairs (checked block-wide); others sort float32.  Unequal group sizes either way round, a row
    beyond the 16-bit range, a row with the float next to a grid value -- identical to the float32 path and to
    the oracle, KS only and with the rank statistics."""
    rng = np.random.default_rng(33)
    L = 400
    c0 = np.full(L, 30, np.int64)
    c1 = np.full(L, 30, np.int64)
    deep = {5: (2000, 2000), 6: (3000, 600), 7: (600, 3000), 8: (129, 131), 9: (5, 140), 200: (2048, 2047), 201: (1000, 257),
            202: (2000, 2000), 203: (2000, 2000), 399: (513, 512)}
    for i, (a, b) in deep.items():
        c0[i], c1[i] = a, b
    off0 = np.concatenate([[0], np.cumsum(c0)])
    off1 = np.concatenate([[0], np.cumsum(c1)])
    v0 = np.round(rng.normal(0, 1, off0[-1]), 3).astype(np.float32)
    shift = np.zeros(L)
    shift[5], shift[200] = 0.2, 4.0
    v1 = np.round(rng.normal(0, 1, off1[-1]) + np.repeat(shift, c1), 3).astype(np.float32)
    v1[off1[202]:off1[203]] += 40.0                                   # beyond the 16-bit range: float32 sorts
    v0[off0[203]:off0[204]] = np.float32(0.25)                       # a false tie would change D and U here
    v1[off1[203]:off1[203] + 1000] = np.nextafter(np.float32(0.25), np.float32(1))
    v1[off1[203] + 1000:off1[204]] = np.nextafter(np.float32(0.25), np.float32(-1))
    p = nm.Pileup.from_arrays(v0, off0, v1, off1, np.arange(L, dtype=np.int32))
    for opt in (nm.DetectOptions(neighborPvalues=3, both_combinations=True),
                nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False)):
        t = det.detect(p, opt)
        f = det_float.detect(p, opt)
        # bit for bit, except the Welch columns of deep rows: the two kernels sum the moments in different orders
        for name in ("row_pos_index", "n0", "n1", "ks_dnum", "two_u", "flags", "ks_d", "ks_p", "u_stat", "u_p",
                     "fisher_stat", "fisher_p", "stouffer_stat", "stouffer_p"):
            x, y = getattr(t, name), getattr(f, name)
            assert (x is None) == (y is None) and (x is None or np.array_equal(x, y, equal_nan=True)), name
        if opt.want_t:
            assert np.allclose(t.t_stat, f.t_stat, rtol=1e-11, atol=1e-13) and np.allclose(t.t_p, f.t_p, rtol=1e-9, atol=0)
        assert_table_matches(t, vec(p, opt), opt)
    assert t.ks_dnum[203] == 2000 * 1000


def test_grid_check_statement_on_the_device(det):
    """nm_grid_selftest: the device evaluates nm_grid_bits on all 2^32 float32 patterns -- a pattern passes
    iff it is fl32(fl64(k / 1000)), |k| <= 32766 (or -0.0), and the packed key halves are k + 32768."""
    viol, n_pass = det.handle.grid_selftest()
    assert viol == 0 and n_pass == 65534
