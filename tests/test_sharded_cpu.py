"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups exercising shard planning, halo
recomputation, trimming and the final gather.  The per-shard compute is injected: a checker
engine built on the vectorised oracle stands in for nanomod_b200.Detector (which needs a GPU), so
these tests cover the host-side distributed logic only -- the GPU tests check the CUDA engine
through the same ShardedDetector.detect_range code path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nanomod_b200 as nm
from nanomod_b200.detect import SignTestTable
from nanomod_b200.sharded import (ShardedDetector, greedy_sites, local_head_from_table, merge_heads, pack_records,
                                  plan_shards, shard_halo, shard_with_halo, unpack_records)
from oracle import nanomod_oracle_vec as ov


class OracleEngine:
    """Test-only stand-in for Detector: same detect(pileup, options) -> SignTestTable contract."""

    def detect(self, p, opt):
        res = ov.detect(p.vals0, p.off0, p.vals1, p.off1, p.pos, p.seg, opt.MinCoverage, opt.neighborPvalues,
                        opt.WeightsDif, ("stouffer", "fisher"))
        idx = res["row_pos_index"].astype(np.int32)
        n = len(idx)
        g = lambda k, dt: np.ascontiguousarray(res[k], dtype=dt) if n else np.zeros(0, dt)
        return SignTestTable(options=opt, seg_names=p.seg_names, seg=p.seg[idx], pos=p.pos[idx], base=p.base[idx],
                             row_pos_index=idx, n0=g("n0", np.int32), n1=g("n1", np.int32),
                             ks_dnum=g("dnum", np.int32), ks_d=g("D", np.float64), ks_p=g("pks", np.float64),
                             two_u=g("twoU", np.int64), u_stat=g("U", np.float64), u_p=g("pu", np.float64),
                             t_stat=g("t", np.float64), t_p=g("pt", np.float64),
                             fisher_stat=g("fisher_stat", np.float64) if n else np.zeros(0),
                             fisher_p=g("fisher_p", np.float64) if n else np.zeros(0),
                             stouffer_stat=g("stouffer_stat", np.float64) if n else np.zeros(0),
                             stouffer_p=g("stouffer_p", np.float64) if n else np.zeros(0),
                             flags=g("uflag", np.uint8))


def make_pileup():
    return nm.synthetic_pileup(3000, 12, 14, drop_frac1=0.02, two_strands=True, poisson=True, clip=(2, 40),
                               round_decimals=2)


def test_plan_shards_balanced_and_covering():
    p = make_pileup()
    for world in (1, 2, 3, 8, 64):
        sh = plan_shards(p.off0, p.off1, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == p.n_pos
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        work = [(p.off0[hi] - p.off0[lo]) + (p.off1[hi] - p.off1[lo]) for lo, hi in sh]
        assert max(work) - min(work) <= 2 * 80  # within one position's worth of values
    tiny = nm.synthetic_pileup(3, 6, 6)
    sh = plan_shards(tiny.off0, tiny.off1, 8)
    assert sum(hi - lo for lo, hi in sh) == 3


def test_halo_slices_and_record_roundtrip():
    p = make_pileup()
    sl, lo, hi = shard_with_halo(p, 100, 200, 3)
    assert sl.n_pos == 106 and (lo, hi) == (3, 103) and np.array_equal(sl.pos, p.pos[97:203])
    sl, lo, hi = shard_with_halo(p, 0, 50, 3)
    assert sl.n_pos == 53 and (lo, hi) == (0, 50)
    t = OracleEngine().detect(p, nm.DetectOptions(neighborPvalues=3, both_combinations=True))
    back = unpack_records(pack_records(t), t)
    for c in ("row_pos_index", "ks_dnum", "ks_p", "two_u", "stouffer_stat", "fisher_p", "flags", "pos", "seg", "base"):
        assert getattr(back, c).tobytes() == getattr(t, c).tobytes(), c


def test_sharded_serial_equals_single():
    """Every shard computed with its halo, trimmed and concatenated == the unsharded result."""
    p = make_pileup()
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    eng = OracleEngine()
    full = eng.detect(p, opt)
    sd = ShardedDetector(eng)
    for world in (2, 5):
        parts = [sd.detect_range(p, lo, hi, opt) for lo, hi in plan_shards(p.off0, p.off1, world)]
        for c in ("row_pos_index", "ks_dnum", "ks_p", "stouffer_stat", "stouffer_p", "fisher_stat", "fisher_p"):
            assert np.concatenate([getattr(t, c) for t in parts]).tobytes() == getattr(full, c).tobytes(), (world, c)
    # neighborPvalues larger than a shard: the halo must still suffice
    opt2 = nm.DetectOptions(neighborPvalues=8, testMethod="stouffer")
    full2 = eng.detect(p, opt2)
    parts = [sd.detect_range(p, lo, hi, opt2) for lo, hi in plan_shards(p.off0, p.off1, 7)]
    assert np.concatenate([t.stouffer_stat for t in parts]).tobytes() == full2.stouffer_stat.tobytes()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = make_pileup()
        opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
        table = ShardedDetector(OracleEngine()).detect(p, opt)
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), row=table.row_pos_index, dnum=table.ks_dnum,
                     z=table.stouffer_stat, fp=table.fisher_p, pos=table.pos, seg=table.seg,
                     sites=np.array([s[2] for s in table.called_sites()]))
        else:
            assert table is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_world_gather(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    p = make_pileup()
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    full = OracleEngine().detect(p, opt)
    assert np.array_equal(got["row"], full.row_pos_index) and np.array_equal(got["dnum"], full.ks_dnum)
    assert got["z"].tobytes() == full.stouffer_stat.tobytes() and got["fp"].tobytes() == full.fisher_p.tobytes()
    assert np.array_equal(got["pos"], full.pos) and np.array_equal(got["seg"], full.seg)
    assert list(got["sites"]) == [s[2] for s in full.called_sites()]


# ---------------------------------------------------------------------------------------------
# the table stays sharded: heads of the per-shard rankings, merged; called sites without a gather
# ---------------------------------------------------------------------------------------------
def _shard_tables(p, opt, world):
    """every rank's host table over core + halo rows, with the core row range inside it"""
    eng = OracleEngine()
    out = []
    for lo, hi in plan_shards(p.off0, p.off1, world):
        sl, core_lo, core_hi = shard_with_halo(p, lo, hi, shard_halo(opt))
        t = eng.detect(sl, opt)
        r_lo = int(np.searchsorted(t.row_pos_index, core_lo))
        r_hi = int(np.searchsorted(t.row_pos_index, core_hi))
        out.append((t, r_lo, r_hi))
    return out


@pytest.mark.parametrize("rank_use,method", [("pv", "stouffer"), ("st", "stouffer"), ("pv", "ks"), ("st", "fisher")])
@pytest.mark.parametrize("world", [1, 3, 8])
def test_merged_heads_give_the_single_table_called_sites(world, rank_use, method):
    p = nm.synthetic_pileup(6000, 10, 12, drop_frac1=0.01, two_strands=True, round_decimals=1)
    opt = nm.DetectOptions(neighborPvalues=2, testMethod=method, rankUse=rank_use, topN=12)
    full = OracleEngine().detect(p, opt)
    want_sites = full.called_sites()
    ranked = full.ranked()
    shards = _shard_tables(p, opt, world)
    for want in (5, 40, 400, 100000):
        heads = [local_head_from_table(t, lo, hi, want) for t, lo, hi in shards]
        m = merge_heads(heads, rank_use != "pv")
        # the exact part of the merged head IS the head of the single-table ranking
        assert np.array_equal(m.row[:m.n_exact], ranked[:m.n_exact])
        assert m.n_exact >= min(want, len(full)) or m.complete
        sites, final = greedy_sites(m, opt, p.seg_names)
        assert sites == want_sites[:len(sites)]
        if final:
            assert sites == want_sites
    assert final  # with every row in, the answer is final


def _head_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = make_pileup()
        opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", topN=9)
        t, r_lo, r_hi = _shard_tables(p, opt, world)[rank]
        # a deliberately short first request: the loop has to come back for longer heads
        sites = ShardedDetector(OracleEngine()).called_sites_host(t, r_lo, r_hi, want=3)
        np.save(os.path.join(out_dir, "sites%d.npy" % rank), np.array([s[2] for s in sites]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_called_sites_without_gather(tmp_path, world):
    mp.spawn(_head_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p = make_pileup()
    opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", topN=9)
    want = [s[2] for s in OracleEngine().detect(p, opt).called_sites()]
    assert len(want) > 0
    for r in range(world):
        assert list(np.load(os.path.join(str(tmp_path), "sites%d.npy" % r))) == want
