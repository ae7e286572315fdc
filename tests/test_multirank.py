"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU path — shard ranges,
role weights, the unique-id broadcast plumbing, and the property the whole scheme rests on:
int64 fixed-point partial forces of disjoint i-block shards all-reduce to the single-rank result
bit for bit."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from mdpy_b200 import multigpu


def test_shard_ranges_partition_the_residues():
    for weights in ([1, 1], [1, 1, 1, 1], [1, 0.25], [0.9, 1, 1, 1, 1, 1, 1, 0.2], [1e-9, 1]):
        r = multigpu.shard_ranges(weights)
        m = multigpu.RESIDUES_PER_RANK * len(weights)
        assert r[0][0] == 0 and r[-1][1] == m
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        assert all(hi > lo for (lo, hi), w in zip(r, weights) if w > 0)
        widths = np.array([hi - lo for lo, hi in r], dtype=float)
        assert np.abs(widths / m - np.asarray(weights) / np.sum(weights)).max() < 1.5 / m + 1e-9 or min(weights) < 1e-6
    with pytest.raises(ValueError):
        multigpu.shard_ranges([0, 0])


def test_shard_ranges_cover_every_residue_exactly_once_for_any_weights():
    """Property test: whatever the (non-negative) weights — zeros for ranks that get no pair work included —
    the ranges tile [0, 32 N) without gap or overlap, which is what makes every i-block owned exactly once."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.one_of(st.just(0.0), st.floats(min_value=1e-6, max_value=1e3)), min_size=1, max_size=8))
    def check(weights):
        if sum(weights) <= 0:
            with pytest.raises(ValueError):
                multigpu.shard_ranges(weights)
            return
        r = multigpu.shard_ranges(weights)
        m = multigpu.RESIDUES_PER_RANK * len(weights)
        owner = np.zeros(m, dtype=int)
        for lo, hi in r:
            assert 0 <= lo <= hi <= m
            owner[lo:hi] += 1
        assert (owner == 1).all()
        assert all((hi > lo) == (w > 0) or (w > 0 and hi > lo) for (lo, hi), w in zip(r, weights) if w > 0)
        assert all(hi == lo for (lo, hi), w in zip(r, weights) if w == 0)
    check()


def test_role_weights_equalise_rank_times():
    for n, pair, pme, bonded in ((2, 70.0, 54.0, 16.0), (8, 3000.0, 350.0, 40.0), (4, 50.0, 80.0, 5.0)):
        x = multigpu.role_weights(n, pair, pme, bonded)
        assert x.sum() == pytest.approx(1) and (x >= 0).all()
        t = pair * x; t[-1] += pme; t[0] += bonded
        busy = t[x > 0]
        assert busy.max() - busy.min() < 1e-6 * pair
        assert t.max() <= max(pme, bonded, busy.max()) + 1e-9


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    from oracle import cpu_oracle as ora
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # 1. unique-id plumbing with a stand-in device object
        class FakeDev:
            def comm_unique_id(self):
                return bytes(range(128))
        uid = multigpu.broadcast_unique_id(dist, FakeDev(), rank)
        ok_uid = uid == bytes(range(128))
        # 2. per-rank timings differ in their last digits: the weights every rank uses must be rank 0's
        #    (weights that differ between ranks give overlapping or missing shard ranges)
        class FakeCtxDev:
            def comm_init(self, *a): pass
            def set_shard(self, lo, hi, mod): self.shard = (lo, hi, mod)
            comm_unique_id = FakeDev.comm_unique_id
        class FakeCtx:
            dev = FakeCtxDev()
        noisy = multigpu.role_weights(world, 70.0 + 3.0 * rank, 30.0 - 2.0 * rank, 0.0)
        fc = FakeCtx()
        multigpu.attach(fc, dist, rank, world, noisy)
        mine_t = torch.tensor([fc.shard[0], fc.shard[1]]); both = [torch.zeros(2, dtype=mine_t.dtype) for _ in range(world)]
        dist.all_gather(both, mine_t)
        ranges = [tuple(int(v) for v in t) for t in both]
        ok_uid = ok_uid and ranges == multigpu.shard_ranges(multigpu.role_weights(world, 70.0, 30.0, 0.0))
        # 3. sharded partial forces -> fixed point -> all_reduce == single-rank result, bit for bit
        g = load_golden('mix_small_f64')
        n = g['positions'].shape[0]
        n_blocks = (n + 31) // 32
        lo, hi = multigpu.shard_ranges(multigpu.role_weights(world, 70.0, 30.0, 5.0))[rank]
        mod = multigpu.RESIDUES_PER_RANK * world
        mine = [b for b in range(n_blocks) if lo <= b % mod < hi]
        part = np.zeros((n, 3))
        for b in mine:
            t = ora.nonbonded_bruteforce(g['positions'], g['box'], g['lj_table'], g['charges'], g['bonded'], g['scaling'],
                                         rc_lj=9.0, i_range=(32 * b, min(n, 32 * b + 32)))
            part[32 * b:32 * b + 32] = t['f_lj'][32 * b:32 * b + 32]
        fix = torch.from_numpy(np.rint(part * 2.0 ** 40).astype(np.int64))
        dist.all_reduce(fix)
        total = np.zeros((n, 3))
        if rank == 0:
            full = ora.nonbonded_bruteforce(g['positions'], g['box'], g['lj_table'], g['charges'], g['bonded'],
                                            g['scaling'], rc_lj=9.0)
            total = np.rint(full['f_lj'] * 2.0 ** 40).astype(np.int64)
            q.put((ok_uid, bool(np.array_equal(fix.numpy(), total)), len(mine)))
        else:
            q.put((ok_uid, True, len(mine)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_fixed_point_allreduce_is_exact():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert all(r[0] and r[1] for r in res)
    assert sum(r[2] for r in res) == (2701 + 31) // 32      # every i-block owned exactly once
