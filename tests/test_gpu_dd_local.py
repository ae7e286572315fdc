"""Spatial domain decomposition with halo exchange (csrc/mdk_dd.cu) on ONE GPU: several device contexts of this
process act as the ranks of the job ("local group": the decomposition, ownership, halo index lists, pack / unpack
kernels, PME sub-mesh traffic, state gather and rebuild logic are exactly those of a multi-process NCCL run; only the
transfers are device-to-device copies instead of ncclSend / ncclRecv).  Runs on the single-GPU box the driver tests
on; tests/test_gpu_multi.py holds the same checks over NCCL for boxes with several GPUs.

Gate: the decomposed job reproduces the single-domain result — forces to 5e-6 relative RMS (the int64 accumulation
makes the sum independent of who evaluates what; what differs is the float32 partial sum inside a work unit, since
the lists group the pairs differently, and the PME charge mesh, whose sub-meshes travel as float32 and are added on
the mesh rank: measured 1.8e-6), energies to 1e-6 of the terms, and a Langevin trajectory with list rebuilds and
atom migration between domains to 1e-3 A."""
import numpy as np
import pytest

from conftest import rel_rms
from mdpy_b200 import _native, synthetic
from mdpy_b200.integrator import LangevinIntegrator
from mdpy_b200.unit import KB, Quantity, default_energy_unit, kelvin

pytestmark = pytest.mark.gpu
KT = float((Quantity(300, kelvin) * KB).convert_to(default_energy_unit).value)


def make_system():
    return synthetic.solvated_protein_box(20002, (60.0, 60.0, 60.0), protein_fraction=0.1, seed=3)


def make_ensemble(s):
    return s.ensemble(cutoff=10.0, switch=8.0, pme=True, grid=(60, 60, 60))


@pytest.fixture(scope='module')
def single():
    """Single-domain reference: forces / energies at the start, then 12 + 60 Langevin steps."""
    s = make_system()
    ens = make_ensemble(s)
    ens.update()
    ctx = _native.context_of(ens)
    out = dict(system=s, f0=ens.forces.copy(), e0=ctx.dev.last_energies().copy(), pot0=ens.potential_energy)
    dev = ctx.dev
    terms = 0
    for c in ens.constraints:
        terms |= c.terms
    dev.step_langevin(0.5, KT, 0.05, 5, 12, terms)          # off the lattice clashes
    dev.step_langevin(1.0, KT, 0.01, 5, 60, terms)
    out.update(x=dev.download_positions(unwrapped=True), v=dev.download_velocities(), e=dev.last_energies().copy(),
               rebuilds=dev.timing()['rebuilds'])
    dev.close()
    return out


@pytest.mark.parametrize('grid', [(2, 1, 1), (2, 2, 1), (2, 2, 2), (1, 1, 3)])
def test_decomposed_job_equals_single_domain(single, grid):
    s = single['system']
    n = int(np.prod(grid))
    group = _native.LocalGroup([make_ensemble(s) for _ in range(n)], grid)
    e0, forces = group.compute()
    scale = np.abs(single['e0'][:10]).sum()
    for r, f in enumerate(forces):          # every rank ends up with all forces
        assert rel_rms(f, single['f0']) < 5e-6, (grid, r)
    assert np.abs(e0[:10] - single['e0'][:10]).max() < 1e-6 * scale, (e0[:10], single['e0'][:10])
    stats = [c.dev.dd_stats() for c in group.ctxs]
    own = sorted((st['own_lo'], st['own_hi']) for st in stats)
    assert own[0][0] == 0 and own[-1][1] == s.num_particles
    assert all(a[1] == b[0] for a, b in zip(own[:-1], own[1:]))               # the domains tile the atoms
    assert all(st['halo_atoms_in'] > 0 and st['halo_atoms_out'] > 0 and st['ranks'] == n for st in stats)
    assert sum(st['halo_atoms_in'] for st in stats) == sum(st['halo_atoms_out'] for st in stats)
    sizes = [b - a for a, b in own]
    assert max(sizes) < 1.35 * min(sizes)                                     # equal-volume domains of a uniform box
    # a second evaluation reuses the lists (halo exchange only) and is bitwise identical
    e1, forces1 = group.compute()
    assert np.array_equal(forces1[0], forces[0]) and np.array_equal(e1, e0)
    # trajectory: the same noise (Philox counter = atom id, step), rebuilds with migration between domains
    group.step_langevin(0.5, KT, 0.05, 5, 12)
    e = group.step_langevin(1.0, KT, 0.01, 5, 60)
    stats = [c.dev.dd_stats() for c in group.ctxs]
    assert all(st['rebuilds'] >= 3 for st in stats), stats                    # initial + at least two during the run
    box = s.box
    for c in group.ctxs:                     # every rank returns the complete state
        d = c.dev.download_positions(unwrapped=True) - single['x']
        d -= box * np.round(d / box)
        assert np.abs(d).max() < 1e-3, grid
        assert np.abs(c.dev.download_velocities() - single['v']).max() < 1e-4
    assert np.abs(e[:11] - single['e'][:11]).max() < 1e-4 * np.abs(single['e'][:11]).sum()
    for c in group.ctxs:
        c.dev.close()


def test_decomposed_pair_sets_partition_the_single_domain_set(single):
    """The pairs the ranks evaluate (production pair kernel, emitted per rank) are disjoint and their union is
    the canonical pair set: nothing is lost or counted twice at a domain boundary."""
    from oracle import cpu_oracle as ora
    s = single['system']
    grid = (2, 2, 1)
    group = _native.LocalGroup([make_ensemble(s) for _ in range(4)], grid)
    group.compute()
    keys = []
    for c in group.ctxs:
        p = c.dev.pairs(production=True)
        keys.append(p[:, 0].astype(np.int64) * (1 << 32) + p[:, 1])
    allk = np.concatenate(keys)
    assert len(np.unique(allk)) == len(allk)                                  # no pair on two ranks
    x = group.ctxs[0].dev.download_positions()
    topo = group.ensembles[0].topology
    want = ora.pair_set_f32(x, np.float32(s.box), 10.0, topo.bonded_particles, threads=8)
    assert np.array_equal(np.sort(allk), want[:, 0].astype(np.int64) * (1 << 32) + want[:, 1])
    share = np.array([len(k) for k in keys]) / len(allk)
    assert share.max() < 1.25 * share.min(), share                            # the seam pairs are shared out evenly
    for c in group.ctxs:
        c.dev.close()


def test_weighted_domains_shrink_the_mesh_rank_and_change_nothing_else(single):
    """mdk_dd_set_weights: the domains are a recursive bisection whose volumes follow the ranks' work weights (the rank
    that also runs the PME mesh chain gets less pair work); forces and energies are those of the single domain."""
    s = single['system']
    weights = [1.1, 1.1, 1.1, 1.1, 1.1, 1.1, 1.1, 0.3]
    group = _native.LocalGroup([make_ensemble(s) for _ in range(8)], (2, 2, 2), weights=weights)
    e0, forces = group.compute()
    assert rel_rms(forces[0], single['f0']) < 5e-6 and rel_rms(forces[7], single['f0']) < 5e-6
    assert np.abs(e0[:10] - single['e0'][:10]).max() < 1e-6 * np.abs(single['e0'][:10]).sum()
    stats = [c.dev.dd_stats() for c in group.ctxs]
    sizes = np.array([st['own_hi'] - st['own_lo'] for st in stats])
    assert sizes.sum() == s.num_particles
    assert sizes[7] < 0.6 * sizes[:7].mean(), sizes            # cut at cell granularity: 8 cells per axis here
    e = group.step_langevin(0.5, KT, 0.05, 5, 30)
    assert np.isfinite(e).all()
    for c in group.ctxs:
        c.dev.close()
