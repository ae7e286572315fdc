"""Rigid three-site waters (SETTLE behind the reference's is_SHAKE flag, forcefield/charmm_forcefield.py:24,32;
no reference code exists, parity unpinned): the device's constrained G-JF step against the host restatement
(oracle/settle.py pinned against iterative SHAKE + oracle/cpu_oracle.py:gjf_rigid_water_call), the constraints over a run
at the benchmark's 2 fs, equipartition over the remaining degrees of freedom, and the list-rebuild interval."""
import numpy as np
import pytest

from conftest import rel_rms
from mdpy_b200 import _native, synthetic
from mdpy_b200.integrator import LangevinIntegrator
from mdpy_b200.unit import KB, Quantity, default_energy_unit, kelvin
from oracle import cpu_oracle as ora

pytestmark = pytest.mark.gpu
KT = float((Quantity(300, kelvin) * KB).convert_to(default_energy_unit).value)
D_OH = 0.9572
D_HH = 2 * D_OH * np.sin(np.deg2rad(104.52) / 2)


def whole(x, trip, box):
    """Molecules made whole: hydrogens as the images next to their oxygen."""
    x = x.copy()
    for t in (1, 2):
        d = x[trip[:, t]] - x[trip[:, 0]]
        x[trip[:, t]] -= box * np.round(d / box)
    return x


def bond_lengths(x, trip, box):
    x = whole(x, trip, box)
    d = lambda i, j: np.linalg.norm(x[trip[:, i]] - x[trip[:, j]], axis=1)
    return d(0, 1), d(0, 2), d(1, 2)


def test_rigid_water_step_matches_host_restatement_and_holds_the_constraints():
    s = synthetic.water_box(2000, 5, box=np.full(3, 39.2))
    ens = s.ensemble(cutoff=9.0, pme=True, grid=(40, 40, 40), rigid_water=True)
    assert ens.num_constraints == 2                      # LJ + PME: the waters' bond / angle terms are gone
    trip = s.water_triplets()
    ctx = _native.context_of(ens)
    dev = ctx.dev
    LangevinIntegrator(0.5, 300, 0.05, seed=3).integrate(ens, 300)       # off the lattice clashes, constrained from the start
    box = s.box
    for length, want in zip(bond_lengths(dev.download_positions(unwrapped=True), trip, box), (D_OH, D_OH, D_HH)):
        assert np.abs(length - want).max() < 1e-9
    # one step against the host restatement, with the device's own forces as input
    dt, gamma, seed = 2.0, 0.01, 77
    integ = LangevinIntegrator(dt, 300, gamma, seed=seed)
    ens.update()
    x0 = whole(dev.download_positions(unwrapped=True), trip, box)
    v0 = ens.state.velocities.astype(np.float64)
    f0 = dev.forces(np.float64)
    integ.integrate(ens, 1)
    x_dev = dev.download_positions(unwrapped=True)
    f_dev = dev.forces(np.float64)
    masses = np.asarray(ens.topology.masses, dtype=np.float64).reshape(-1)
    x1, v1, _ = ora.gjf_rigid_water_call(x0, v0, f0, lambda _x: f_dev, masses, trip, D_OH, D_HH, dt, gamma, KT, seed, 0)
    d = x_dev - x1
    d -= box * np.round(d / box)
    assert np.abs(d).max() < 1e-6 * np.abs(x1 - x0).max() + 1e-9
    assert rel_rms(ens.state.velocities, v1) < 2e-6       # v0 went in as float32
    # a run at the benchmark's time step: constraints to rounding, no velocity along a bond
    integ.integrate(ens, 500)
    x = dev.download_positions(unwrapped=True)
    for length, want in zip(bond_lengths(x, trip, box), (D_OH, D_OH, D_HH)):
        assert np.abs(length - want).max() < 1e-9
    xw = whole(x, trip, box)
    v = ens.state.velocities.astype(np.float64)
    for i, j in ((0, 1), (0, 2), (1, 2)):
        along = ((v[trip[:, j]] - v[trip[:, i]]) * (xw[trip[:, j]] - xw[trip[:, i]])).sum(1)
        assert np.abs(along).max() < 1e-6
    assert np.isfinite(ens.total_energy)


def test_rigid_water_equipartition_and_rebuild_interval():
    s = synthetic.water_box(2000, 5, box=np.full(3, 39.2))
    rebuilds, temps = {}, {}
    for rigid in (True, False):
        ens = s.ensemble(cutoff=9.0, pme=True, grid=(40, 40, 40), rigid_water=rigid)
        dev = _native.context_of(ens).dev
        LangevinIntegrator(0.5, 300, 0.05, seed=3).integrate(ens, 400)
        integ = LangevinIntegrator(2.0 if rigid else 0.5, 300, 0.01, seed=11)
        integ.integrate(ens, 1500)
        before = dev.timing()['rebuilds']
        t = []
        for _ in range(30):
            integ.integrate(ens, 100)
            dof = 6 * 2000 if rigid else 9 * 2000           # 3 constraints per molecule removed
            t.append(2 * ens.kinetic_energy / dof / (KT / 300.0))
        rebuilds[rigid] = (dev.timing()['rebuilds'] - before) / 3000.0 * (2.0 if rigid else 0.5)   # per fs... per step scaled below
        temps[rigid] = float(np.mean(t))
        dev.close()
    assert abs(temps[True] - 300.0) < 9.0, temps          # +- 3 %: 12 000 degrees of freedom, G-JF at 2 fs
    assert abs(temps[False] - 300.0) < 9.0, temps
