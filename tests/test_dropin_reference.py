"""The drop-in claim, exercised on the REAL reference objects: the unmodified mdpy package (baseline/_ref, the
offline install DESIGN.md records; git-ignored, shipped to the GPU box by gpurun) builds its own Topology / Ensemble /
VerletIntegrator, and the mdpy_b200 constraint classes are bound to that Ensemble in place of the reference's
(the two class names CharmmForcefield.create_ensemble instantiates, forcefield/charmm_forcefield.py:96-112).

  * GPU test: mdpy.Ensemble.update() (ensemble.py:53-61) and the reference's own host-side
    VerletIntegrator.integrate (verlet_integrator.py:20-50) run 5 steps on forces that come from libmdpyb200
    through the C ABI, and land on the trajectory the unmodified reference produced with its own kernels
    (tests/golden/verlet_small_f64.npz).
  * CPU test: the same wiring with the device stubbed out (no GPU): what a reference Ensemble hands to the
    boundary and what it gets back.
Nothing here reads /root/reference.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_rms

REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def import_reference():
    """`import mdpy` from baseline/_ref with the import shims of SURVEY 8c (oracle/refshim)."""
    if not os.path.isdir(os.path.join(REF_DIR, 'mdpy')):
        pytest.skip('baseline/_ref/mdpy is not installed (python -m pip install --no-deps --target baseline/_ref /root/reference)')
    spec = importlib.util.spec_from_file_location('refshim_sitecustomize', os.path.join(ROOT, 'oracle', 'refshim', 'sitecustomize.py'))
    spec.loader.exec_module(importlib.util.module_from_spec(spec))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import mdpy
    assert os.path.realpath(mdpy.__file__).startswith(os.path.realpath(REF_DIR))
    return mdpy


def reference_ensemble(mdpy, s):
    """A reference Topology / Ensemble of the synthetic system, built with the reference's own API
    (the recipe of oracle/make_golden.py:ref_topology)."""
    from mdpy.core import Particle, Topology
    from mdpy.ensemble import Ensemble
    t = Topology()
    t.add_particles([Particle(particle_id=i, particle_type=tp, particle_name=tp, molecule_type='SYN', mass=float(m), charge=float(q))
                     for i, (tp, m, q) in enumerate(zip(s.types, s.masses, s.charges))])
    for b in s.bonds: t.add_bond([int(x) for x in b])
    for a in s.angles: t.add_angle([int(x) for x in a])
    for d in s.dihedrals: t.add_dihedral([int(x) for x in d])
    for d in s.impropers: t.add_improper([int(x) for x in d])
    t.join()
    return Ensemble(t, np.diag(s.box))


def small_system():
    from mdpy_b200 import synthetic
    return synthetic.solvated_protein_box(1471, (24.5, 24.5, 24.5), protein_fraction=0.068, seed=12, n_res=10)


@pytest.mark.gpu
def test_reference_ensemble_and_integrator_run_on_the_dropin_constraints():
    mdpy = import_reference()
    mdpy.env.set_precision('DOUBLE')      # the reference's host arithmetic in float64, like the golden run
    import mdpy_b200
    from mdpy_b200.constraint import CharmmNonbondedConstraint, ElectrostaticConstraint
    from mdpy.integrator import VerletIntegrator
    mdpy_b200.env.set_precision('DOUBLE')  # forces handed back as float64 arrays (the kernels stay float32 + int64 sums)
    g = load_golden('verlet_small_f64')
    s = small_system()
    ens = reference_ensemble(mdpy, s)
    assert type(ens).__module__ == 'mdpy.ensemble'
    lj = CharmmNonbondedConstraint(s.lj_parameters, cutoff_radius=float(g['rc']))
    el = ElectrostaticConstraint()
    ens.add_constraints(lj, el)            # reference ensemble.py:40-51: bind_ensemble + cutoff negotiation
    assert ens.num_constraints == 2 and lj.parent_ensemble is ens and lj.force_id == 0 and el.force_id == 1
    ens.state.set_positions(s.positions.astype(np.float64))
    assert np.array_equal(np.asarray(ens.topology.bonded_particles), g['bonded'])
    # single point through the reference's Ensemble.update
    ens.update()
    assert rel_rms(lj.forces, g['lj_forces']) < 1e-5 and rel_rms(el.forces, g['coul_forces']) < 1e-5
    assert ens.potential_energy == pytest.approx(float(g['lj_energy']) + float(g['coul_energy']), rel=1e-6)
    assert rel_rms(ens.forces, g['lj_forces'] + g['coul_forces']) < 1e-5
    # the reference's own integrator loop: Ensemble.update -> our constraints -> numpy update -> State.set_positions
    ens.state.set_velocities(g['verlet_v0'].astype(np.float64))
    integ = VerletIntegrator(float(g['verlet_dt']))
    integ.integrate(ens, int(g['verlet_steps']))
    assert np.abs(integ.cur_positions - g['verlet_cur']).max() < 2e-5
    assert np.abs(ens.state.positions - g['verlet_positions']).max() < 2e-5
    assert np.abs(ens.state.velocities - g['verlet_velocities']).max() < 1e-5
    # the device list was reused between the steps (the reference rebuilds its cell list on every set_positions)
    assert lj._ctx.dev.timing()['rebuilds'] <= 2


def test_reference_ensemble_binds_the_dropin_constraints_cpu(monkeypatch):
    """Same wiring, device stubbed: the boundary receives the reference Topology's tables and the reference State's
    positions, and Ensemble.update sums what the constraints hand back."""
    mdpy = import_reference()
    import mdpy_b200
    from mdpy_b200 import _native
    from mdpy_b200.constraint import CharmmNonbondedConstraint, ElectrostaticConstraint
    seen = {}

    class Dev:
        def __init__(self, device=None):
            self.n = 0

        def set_box(self, box): seen['box'] = np.array(box)
        def set_atoms(self, q, m): self.n = int(np.asarray(q).size); seen['q'] = np.array(q).reshape(-1)
        def set_exclusions(self, b, s): seen['bonded'] = np.array(b); seen['scaling'] = np.array(s)
        def set_lj(self, table, rc, rs=None): seen['lj'] = (np.array(table), rc, rs)
        def set_coulomb(self, k, a=0.0, rc=0.0): seen['k_e'] = k
        def upload_positions(self, x): seen['x'] = np.array(x); seen['uploads'] = seen.get('uploads', 0) + 1
        def compute(self, terms):
            e = np.zeros(_native.NUM_ENERGIES); e[_native.E_LJ] = -1.5; e[_native.E_COUL_BARE] = 0.25
            seen.setdefault('terms', []).append(terms)
            return e
        def forces(self, dtype=np.float32): return np.full((self.n, 3), 0.5, dtype=dtype)

    monkeypatch.setattr(_native, 'Device', Dev)
    s = small_system()
    ens = reference_ensemble(mdpy, s)
    lj = CharmmNonbondedConstraint(s.lj_parameters, cutoff_radius=9.0)
    el = ElectrostaticConstraint()
    ens.add_constraints(lj, el)
    ens.state.set_positions(s.positions.astype(mdpy.env.NUMPY_FLOAT))
    ens.update()
    assert seen['terms'] == [_native.TERM_LJ, _native.TERM_COUL_BARE]       # the reference's per-constraint protocol
    assert seen['uploads'] == 1                                             # positions uploaded once, not per constraint
    assert np.allclose(seen['box'], s.box) and np.allclose(seen['q'], s.charges)
    assert np.array_equal(seen['bonded'], np.asarray(ens.topology.bonded_particles))
    assert seen['lj'][0].shape == (s.num_particles, 4) and seen['lj'][1] == 9.0
    assert np.allclose(seen['x'], ens.state.positions)
    assert np.all(ens.forces == 1.0) and ens.potential_energy == pytest.approx(-1.25)
    ens.state.set_positions(ens.state.positions + 0.01)                     # a new positions array -> re-upload
    ens.update()
    assert seen['uploads'] == 2
