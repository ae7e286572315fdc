"""Host side of the device integrators with the device stubbed out (no GPU): what
LangevinIntegrator.integrate hands to mdk_step_langevin_host and what it does with the result
(mdpy/integrator/integrator.py:14-50, langevin_integrator.py:37-72 are the reference's versions)."""
import numpy as np
import pytest

import mdpy_b200 as md
from mdpy_b200 import _native, synthetic
from mdpy_b200.integrator import LangevinIntegrator


class StubDevice:
    """Records the calls an EnsembleContext / integrator makes; step_langevin_host shifts x by +0.01."""

    def __init__(self, device=None):
        self.n = 0
        self.calls = []

    def set_atoms(self, q, m):
        self.n = int(np.asarray(q).reshape(-1).size)

    def __getattr__(self, name):
        if name.startswith('set_') or name == 'reset_integrator':
            return lambda *a, **k: self.calls.append(name)
        raise AttributeError(name)

    def pinned_empty(self, shape, dtype=np.float32):
        return np.zeros(shape, dtype=dtype)

    def step_langevin_host(self, x_in, v_in, x_out, v_out, dt, kT, gamma, seed, nsteps, terms):
        assert x_in.dtype == np.float32 and v_in.dtype == np.float32 and x_in.flags.c_contiguous
        assert x_out is not x_in and v_out is not v_in
        self.calls.append(('step', id(x_in), id(x_out), nsteps, terms, dt))
        x_out[...] = x_in + np.float32(0.01) * nsteps
        v_out[...] = v_in + np.float32(1.0)
        e = np.zeros(_native.NUM_ENERGIES)
        e[_native.E_LJ], e[_native.E_COUL_DIRECT], e[_native.E_PME_RECIP] = -3.0, -2.0, -1.0
        e[_native.E_BOND], e[_native.E_KINETIC] = 0.5, 7.0
        return e


@pytest.fixture
def stub_ensemble(monkeypatch):
    monkeypatch.setattr(_native, 'Device', StubDevice)
    s = synthetic.water_box(300, 4, box=np.full(3, 30.0))
    return s, s.ensemble(cutoff=9.0, pme=True, grid=(32, 32, 32))


def test_integrate_hands_the_host_state_in_and_installs_what_comes_back(stub_ensemble):
    s, ens = stub_ensemble
    dev = _native.context_of(ens).dev
    integ = LangevinIntegrator(2.0, 300, 1e-3, seed=1)
    x0 = ens.state.positions.copy()
    rev0 = ens.state.revision
    integ.integrate(ens, 3)
    step = [c for c in dev.calls if isinstance(c, tuple)][-1]
    assert step[3] == 3 and step[5] == 2.0
    terms = 0
    for c in ens.constraints:
        terms |= c.terms
    assert step[4] == terms
    assert np.allclose(ens.state.positions, x0 + 0.03, atol=1e-6)
    assert ens.state.positions.dtype == np.float32 and ens.state.velocities.dtype == np.float32
    assert ens.state.revision == rev0 + 1
    assert integ.is_cached
    # energies: potential = sum over the constraints' slots, kinetic separately
    assert ens.kinetic_energy == 7.0
    assert ens.potential_energy == pytest.approx(-3.0 - 2.0 - 1.0 + 0.5)
    assert ens.total_energy == pytest.approx(ens.potential_energy + 7.0)
    assert ens.constraints[0].potential_energy == -3.0


def test_state_arrays_are_never_overwritten_while_referenced(stub_ensemble):
    """The arrays a call publishes as the new State come from a pool of page-locked blocks; a block is reused only
    when nobody holds it any more, so kept frames stay intact (the reference allocates a fresh array per
    set_positions, state.py:58-60)."""
    s, ens = stub_ensemble
    ctx = _native.context_of(ens)
    dev = ctx.dev
    integ = LangevinIntegrator(2.0, 300, 1e-3, seed=1)
    integ.integrate(ens, 1)
    first_out = ens.state.positions
    first_copy = first_out.copy()
    integ.integrate(ens, 1)
    steps = [c for c in dev.calls if isinstance(c, tuple)]
    assert steps[1][1] == id(first_out)               # the State array itself goes in: no copy, no revision check
    assert ens.state.positions is not first_out       # ... and is not the buffer being written
    frames = [ens.state.positions]
    for _ in range(5):                                # a dumper that keeps references
        integ.integrate(ens, 1)
        frames.append(ens.state.positions)
    assert np.array_equal(first_out, first_copy)
    for a, b in zip(frames[:-1], frames[1:]):
        assert np.allclose(b, a + 0.01, atol=1e-6)    # every kept frame is still its own step
    # views count as references too
    view = ens.state.positions[::2]
    keep = view.copy()
    for _ in range(3):
        integ.integrate(ens, 1)
    assert np.array_equal(view, keep)
    # once nobody holds the old arrays the blocks are reused: the pool stops growing
    del frames, first_out, view
    for _ in range(4):
        integ.integrate(ens, 1)
    size = len(ctx._state_pool)
    used = set()
    for _ in range(6):
        integ.integrate(ens, 1)
        used.add(id(ens.state.positions))
    assert len(ctx._state_pool) == size <= ctx.MAX_STATE_BUFFERS
    assert len(used) == 2                             # steady state: two blocks take turns
    # in-place edits of the State arrays are part of the next call's input
    ens.state.positions[0, 0] = 5.0
    integ.integrate(ens, 1)
    assert ens.state.positions[0, 0] == pytest.approx(5.01, abs=1e-6)
    # a caller that keeps more frames than the pool holds gets plain arrays, still never overwritten
    hoard = []
    for _ in range(ctx.MAX_STATE_BUFFERS + 3):
        integ.integrate(ens, 1)
        hoard.append((ens.state.positions, ens.state.positions.copy()))
    assert all(np.array_equal(a, b) for a, b in hoard)


def test_a_new_integrator_object_drops_the_device_step_caches(stub_ensemble):
    s, ens = stub_ensemble
    dev = _native.context_of(ens).dev
    a, b = LangevinIntegrator(2.0, 300, 1e-3), LangevinIntegrator(1.0, 300, 1e-3)
    a.integrate(ens, 1)
    n_reset = dev.calls.count('reset_integrator')
    a.integrate(ens, 1)
    assert dev.calls.count('reset_integrator') == n_reset          # same owner: caches kept
    b.integrate(ens, 1)
    assert dev.calls.count('reset_integrator') == n_reset + 1      # like a fresh reference integrator (integrator.py:20-22)
    a.erase_cache()
    assert not a.is_cached


def test_foreign_constraints_are_refused(stub_ensemble):
    s, ens = stub_ensemble

    class Foreign:
        is_native = False
        cutoff_radius = 1.0
        def bind_ensemble(self, e): pass
    ens._constraints.append(Foreign())
    with pytest.raises(TypeError):
        LangevinIntegrator(2.0, 300, 1e-3).integrate(ens, 1)


def test_double_precision_state_is_converted_at_the_boundary(monkeypatch):
    monkeypatch.setattr(_native, 'Device', StubDevice)
    md.env.set_precision('DOUBLE')
    s = synthetic.water_box(300, 4, box=np.full(3, 30.0))
    ens = s.ensemble(cutoff=9.0, pme=False)
    assert ens.state.positions.dtype == np.float64
    x0 = ens.state.positions.copy()
    LangevinIntegrator(2.0, 300, 1e-3).integrate(ens, 2)
    assert ens.state.positions.dtype == np.float64 and ens.state.velocities.dtype == np.float64
    assert np.allclose(ens.state.positions, x0 + 0.02, atol=1e-5)


def test_ensemble_update_fuses_native_constraints_into_one_evaluation(monkeypatch):
    """Ensemble.update (mdpy/ensemble.py:53-61): with every constraint native, one device evaluation of the
    union of the terms; fused=False keeps the reference's per-constraint protocol.  Energies are split by
    each constraint's slots either way."""
    class Dev(StubDevice):
        def compute(self, terms):
            self.calls.append(('compute', terms))
            e = np.zeros(_native.NUM_ENERGIES)
            e[_native.E_LJ], e[_native.E_BOND], e[_native.E_ANGLE] = -1.0, 0.25, 0.125
            return e

        def forces(self, dtype=np.float32):
            return np.full((self.n, 3), 0.5, dtype=dtype)

        def upload_positions(self, x):
            self.calls.append('upload')

    monkeypatch.setattr(_native, 'Device', Dev)
    s = synthetic.water_box(300, 4, box=np.full(3, 30.0))
    ens = s.ensemble(cutoff=9.0, pme=True, grid=(32, 32, 32))
    dev = _native.context_of(ens).dev
    ens.state.set_velocities(np.full((900, 3), 0.01, dtype=np.float32))
    ens.update()
    computes = [c for c in dev.calls if isinstance(c, tuple) and c[0] == 'compute']
    union = 0
    for c in ens.constraints:
        union |= c.terms
    assert computes == [('compute', union)]
    assert ens.forces.dtype == np.float64 and ens.forces.shape == (900, 3) and np.all(ens.forces == 0.5)
    assert ens.potential_energy == pytest.approx(-1.0 + 0.25 + 0.125)
    m = np.asarray(ens.topology.masses, dtype=np.float64).reshape(-1)
    assert ens.kinetic_energy > 0 and ens.total_energy == pytest.approx(ens.potential_energy + ens.kinetic_energy)
    # each constraint's own forces: evaluated on first access, with its own terms, energy kept
    lj = ens.constraints[0]
    e_lj = lj.potential_energy
    f = lj.forces
    computes = [c for c in dev.calls if isinstance(c, tuple) and c[0] == 'compute']
    assert computes[-1] == ('compute', lj.terms) and f.shape == (900, 3) and lj.potential_energy == e_lj
    assert lj.forces is f and len([c for c in dev.calls if isinstance(c, tuple) and c[0] == 'compute']) == len(computes)
    ens.update()
    ens.state.set_positions(ens.state.positions + np.float32(0.01))
    with pytest.raises(RuntimeError):
        ens.constraints[1].forces          # positions moved on since the fused evaluation
    computes = [c for c in dev.calls if isinstance(c, tuple) and c[0] == 'compute']
    n_before = len(computes)
    ens.update(fused=False)
    computes = [c for c in dev.calls if isinstance(c, tuple) and c[0] == 'compute']
    assert len(computes) - n_before == ens.num_constraints          # one evaluation per constraint
    assert np.all(ens.forces == 0.5 * ens.num_constraints)
    assert ens.potential_energy == pytest.approx(-1.0 + 0.25 + 0.125)
