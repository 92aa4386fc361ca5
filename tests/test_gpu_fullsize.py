"""Parity AT THE SIZES BASELINE.json quotes (configs[2], [3], [4]); configs[1] is covered by
tests/test_gpu_parity.py::test_cfg2_full_size_properties.  Device-resident synthetic pileups
(the bench's generator), the CUDA path through `nm_detect_device`, the vectorised oracle on the
same values read back from the device.

  cfg3  E. coli 4.6 Mb, 2x100x, ALL variants: U, Welch t, KS, Fisher AND Stouffer.  Per-position
        tests on a 1 % sample + every planted site +-10; both window combinations on EVERY row
        (oracle combine over the GPU's 4.6 M KS p-values).
  cfg4  human chr20, 64 444 167 positions, 2x30x, KS + Stouffer: per-position tests on a 0.25 %
        sample + planted sites, combination on four 1 M-row windows (incl. both ends).
  cfg5  50 kb plasmid, 2x2000x (deep tier), KS + Stouffer: EVERY one of the 50 000 positions
        against the oracle, with +4 sigma sites driving p below DBL_MIN (clamp, myDetect.py:317).
Reference call sites: myDetect.py:331,335,341 (tests), :393,401 (combination).
"""
import numpy as np
import pytest

import nanomod_b200 as nm
from oracle import nanomod_oracle as o
from oracle import nanomod_oracle_vec as ov
from test_gpu_parity import close

pytestmark = pytest.mark.gpu
CHR20 = 64_444_167


@pytest.fixture(scope="module")
def det():
    return nm.Detector(0)


def sample_rows(L, host_shift, frac, rng, near_step=1):
    planted = np.nonzero(host_shift >= 1.0)[0]
    near = np.unique(np.clip(planted[:, None] + np.arange(-10, 11)[None, :], 0, L - 1))[::near_step]
    return np.unique(np.concatenate([rng.choice(L, int(L * frac), replace=False), near, [0, 1, 2, L - 3, L - 2, L - 1]]))


def oracle_block(dev, rows, L, n0, n1, chunk=20000):
    import torch
    out = {}
    for s in range(0, len(rows), chunk):
        r_t = torch.from_numpy(rows[s:s + chunk]).to(dev.vals0.device)
        A = dev.vals0[: L * n0].view(L, n0)[r_t].double().cpu().numpy()
        B = dev.vals1[: L * n1].view(L, n1)[r_t].double().cpu().numpy()
        blk = ov.tests_block(A, np.full(len(r_t), n0), B, np.full(len(r_t), n1))
        for k, v in blk.items():
            out.setdefault(k, []).append(v)
    return {k: np.concatenate(v) for k, v in out.items()}


def check_tests(out, rows, blk, want_ut):
    import torch
    r_t = torch.from_numpy(rows).cuda()
    g = lambda c: out[c][r_t].cpu().numpy()
    assert np.array_equal(g("ks_dnum"), blk["dnum"]), "KS numerator must be bit-exact"
    for col, key, atol in (("ks_p", "pks", 0), ("ks_d", "D", 0)):
        ok, i = close(g(col), blk[key], atol=atol)
        assert ok, (col, rows[i], g(col)[i], blk[key][i])
    if want_ut:
        assert np.array_equal(g("two_u"), blk["twoU"]), "2U must be bit-exact"
        for col, key, atol in (("u_stat", "U", 0), ("u_p", "pu", 0), ("t_stat", "t", 1e-12), ("t_p", "pt", 0)):
            ok, i = close(g(col), blk[key], atol=atol)
            assert ok, (col, rows[i], g(col)[i], blk[key][i])


def check_combine(out, lo, hi, nb, wd, methods, pos0=0, whole=False):
    """combination of the GPU's own KS p-values over rows [lo, hi) == oracle combine; rows within
    nb of a cut that is not a real end of the run are skipped"""
    pks = out["ks_p"][lo:hi].cpu().numpy()
    n = hi - lo
    pos = np.arange(pos0 + lo, pos0 + hi, dtype=np.int32)
    seg = np.zeros(n, np.int32)
    a = 0 if (whole or lo == 0) else nb
    for m in methods:
        c = ov.combine(pks, pos, seg, nb, wd, m)
        b = n if whole else n - nb
        gs = out[m + "_stat"][lo:hi].cpu().numpy()
        gp = out[m + "_p"][lo:hi].cpu().numpy()
        assert np.array_equal(np.isneginf(gs[a:b]), np.isneginf(c["stat"][a:b]))
        ok, i = close(gs[a:b], c["stat"][a:b], atol=1e-9)
        assert ok, (m, "stat", lo + a + i, gs[a + i], c["stat"][a + i])
        ok, i = close(gp[a:b], c["p"][a:b])
        assert ok, (m, "p", lo + a + i, gp[a + i], c["p"][a + i])


def test_cfg3_all_variants_full_size(det):
    import torch
    from bench import make_device_workload
    L, n = 4_600_000, 100
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", both_combinations=True, SaveTest=0)
    dev, host_shift = make_device_workload(L, n, n, torch.device("cuda:0"))
    out = nm.alloc_device_table(opt, L, "cuda:0")
    assert det.detect_device(dev, opt, out) == L
    rows = sample_rows(L, host_shift, 0.01, np.random.default_rng(3))
    assert len(rows) >= 46000 + 4600 * 21 * 0.9
    check_tests(out, rows, oracle_block(dev, rows, L, n, n), want_ut=True)
    check_combine(out, 0, L, 3, 2.0, ("stouffer", "fisher"), whole=True)
    assert int(torch.isneginf(out["stouffer_stat"]).sum().item()) == 6
    assert int((out["flags"] != 0).sum().item()) == 0


def test_cfg4_chr20_full_size(det):
    import torch
    from bench import make_device_workload
    L, n = CHR20, 30
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False, SaveTest=0)
    dev, host_shift = make_device_workload(L, n, n, torch.device("cuda:0"))
    out = nm.alloc_device_table(opt, L, "cuda:0")
    assert det.detect_device(dev, opt, out) == L
    rows = sample_rows(L, host_shift, 0.0025, np.random.default_rng(4), near_step=5)
    assert len(rows) > 400_000
    check_tests(out, rows, oracle_block(dev, rows, L, n, n, chunk=100000), want_ut=False)
    for lo in (0, 21_000_000, 43_000_123, L - 1_000_000):
        check_combine(out, lo, lo + 1_000_000, 3, 2.0, ("stouffer",))
    assert int(torch.isneginf(out["stouffer_stat"]).sum().item()) == 6
    # every planted site is called: the 200 smallest combined p-values sit on planted positions
    top = torch.argsort(out["stouffer_p"])[:200].cpu().numpy()
    assert np.mean(host_shift[top] > 0) > 0.95


def test_cfg5_deep_plasmid_every_position(det):
    import torch
    from bench import make_device_workload
    L, n = 50_000, 2000
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", SaveTest=0)
    dev, _shift = make_device_workload(L, n, n, torch.device("cuda:0"))
    hot = np.array([77, 5000, 10001, 15002, 20003, 25004, 30005, 35006, 40007, 49990])
    dev.vals1[: L * n].view(L, n)[torch.from_numpy(hot).cuda()] += 4.0  # lambda > 18.8: p underflows
    out = nm.alloc_device_table(opt, L, "cuda:0")
    assert det.detect_device(dev, opt, out) == L
    rows = np.arange(L)
    blk = oracle_block(dev, rows, L, n, n, chunk=1000)
    check_tests(out, rows, blk, want_ut=True)
    assert np.all(out["ks_p"][torch.from_numpy(hot).cuda()].cpu().numpy() == o.FLOAT_MIN)
    check_combine(out, 0, L, 3, 2.0, ("stouffer",), whole=True)
