"""CPU-only (world_size 2 over gloo where ranks are involved): the host-side logic of the domain-decomposed
multi-GPU path — the domain grid, the pair-ownership rule across domain boundaries (host mirror of the list
builder's), the rendezvous plumbing (unique id, communicator, domain) and the halo-force protocol: int64
fixed-point partial forces returned to the owner add up to the all-reduced sum bit for bit."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from mdpy_b200 import multigpu


def test_domain_grid_cuts_the_longest_domains_first():
    assert multigpu.domain_grid(1, [50, 60, 70]) == (1, 1, 1)
    assert multigpu.domain_grid(2, [108.86, 108.86, 77.76]) in ((2, 1, 1), (1, 2, 1))
    assert multigpu.domain_grid(2, [50, 60, 70]) == (1, 1, 2)
    assert sorted(multigpu.domain_grid(4, [216.8] * 3)) == [1, 2, 2]
    assert multigpu.domain_grid(8, [216.8] * 3) == (2, 2, 2)           # never 8 slabs thinner than their halo
    assert multigpu.domain_grid(4, [30, 30, 200]) == (1, 1, 4)
    assert int(np.prod(multigpu.domain_grid(6, [60, 60, 60]))) == 6
    with pytest.raises(ValueError):
        multigpu.domain_grid(11, [60, 60, 60])
    with pytest.raises(ValueError):
        multigpu.domain_grid(16, [10, 10, 400])                        # would need more than 4 domains on one axis


def test_every_block_pair_is_listed_by_exactly_one_rank_and_the_seam_work_is_shared():
    """The pair-ownership rule (host mirror of the kernel's): over any contiguous block ownership every unordered
    pair of blocks is listed exactly once, and the pairs across a domain boundary are split evenly between the two
    sides (a plain half shell by index would give them all to the lower domain)."""
    rng = np.random.default_rng(3)
    for ranks in (2, 3, 8):
        n_blocks = 97
        cuts = np.concatenate([[0], np.sort(rng.choice(np.arange(1, n_blocks), size=ranks - 1, replace=False)), [n_blocks]])
        owner = np.searchsorted(cuts, np.arange(n_blocks), side='right') - 1
        listed = np.zeros((n_blocks, n_blocks), dtype=int)
        seam = np.zeros(ranks, dtype=int)
        for b in range(n_blocks):
            lo, hi = cuts[owner[b]], cuts[owner[b] + 1]
            for bj in range(n_blocks):
                if multigpu.lists_pair(b, bj, lo, hi):
                    listed[min(b, bj), max(b, bj)] += 1
                    if owner[bj] != owner[b]:
                        seam[owner[b]] += 1
        iu = np.triu_indices(n_blocks, 1)
        assert np.all(listed[iu] == 1) and listed.trace() == 0
        for r in range(ranks):      # each rank lists half (to rounding) of the pairs between its blocks and the others'
            mine = cuts[r + 1] - cuts[r]
            assert abs(seam[r] - mine * (n_blocks - mine) / 2) <= mine


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # 1. rendezvous: unique id from rank 0, communicator, then the rank's domain of the grid — with stand-in
        #    device / context objects that record what they are told
        class FakeDev:
            calls = []
            def comm_unique_id(self): return bytes(range(128))
            def comm_init(self, r, w, uid): self.calls.append(('comm_init', r, w, uid))
            def dd_init(self, r, w, grid, local_group=-1): self.calls.append(('dd_init', r, w, tuple(grid), local_group))
        class FakeState:
            pbc_matrix = np.diag([60.0, 80.0, 50.0])
        class FakeEns:
            state = FakeState()
        class FakeCtx:
            dev = FakeDev(); ensemble = FakeEns(); _pos_rev = 'stale'
        ctx = FakeCtx()
        grid = multigpu.attach(ctx, dist, rank, world)
        ok = grid == (1, 2, 1) and ctx._pos_rev is None
        ok = ok and ctx.dev.calls[0] == ('comm_init', rank, world, bytes(range(128))) and ctx.dev.calls[1] == ('dd_init', rank, world, (1, 2, 1), -1)
        # 2. the halo protocol with integers: each rank owns half of the atoms and holds fixed-point partial forces
        #    for ALL atoms (its own work touches the other's halo); the halo rows go back to their owner and are added
        #    there — the owner's rows then equal the all-reduced sum bit for bit, whoever computed what
        n = 1000
        own = slice(0, 480) if rank == 0 else slice(480, n)
        other = slice(480, n) if rank == 0 else slice(0, 480)
        rng = np.random.default_rng(100 + rank)
        partial = torch.from_numpy(rng.integers(-2 ** 50, 2 ** 50, size=(n, 3)))
        halo_out = partial[other].clone().contiguous()
        halo_in = torch.zeros_like(partial[own]).contiguous()
        reqs = [dist.isend(halo_out, 1 - rank), dist.irecv(halo_in, 1 - rank)]
        for r in reqs: r.wait()
        mine = partial[own] + halo_in
        total = partial.clone()
        dist.all_reduce(total)
        ok = ok and bool(torch.equal(mine, total[own]))
        # 3. state broadcast helper
        arr = multigpu.broadcast_array(dist, np.arange(5.0) + rank, rank)
        ok = ok and np.array_equal(arr, np.arange(5.0))
        q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_rendezvous_and_halo_protocol():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert all(res)
