"""CPU-only: the host-side mirror of the reference interface (names, argument meaning, error
behaviour), the C-ABI library's export list, and the no-fallback rule."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import mdpy_b200 as md
from conftest import GOLDEN, ROOT, load_golden
from mdpy_b200 import _native, synthetic
from mdpy_b200.constraint import CharmmNonbondedConstraint, ElectrostaticConstraint, ElectrostaticPMEConstraint
from mdpy_b200.constraint.electrostatic_constraint import ewald_alpha, fft_size
from mdpy_b200.core import Particle, Topology
from mdpy_b200.error import (ArrayDimError, CellListPoorDefinedError, ConstraintConflictError, EnvironmentVariableError,
                             ModifyJoinedTopologyError, NonBoundedError, ParticleConflictError, ParticleLossError,
                             UnitDimensionDismatchedError)
from mdpy_b200.unit import (EPSILON0, KB, NA, Quantity, RMIN_TO_SIGMA_FACTOR, angstrom, coulomb_constant,
                            default_energy_unit, default_length_unit, femtosecond, kelvin, kilocalorie_permol,
                            kilojoule_permol, nanometer)
from mdpy_b200.utils import unwrap_vec, wrap_positions


def test_environment_matches_reference_protocol():
    """mdpy/test/test_environment.py:16-25 + the documented platform difference."""
    env = md.env
    assert env.precision == 'SINGLE' and env.NUMPY_FLOAT == np.float32 and env.NUMPY_INT == np.int32
    env.set_precision('double')
    assert env.NUMPY_FLOAT == np.float64 and env.NUMPY_INT == np.int64
    with pytest.raises(EnvironmentVariableError):
        env.set_precision('HALF')
    with pytest.raises(EnvironmentVariableError):
        env.set_platform('OPENCL')
    env.set_platform('CPU')           # accepted name ...
    assert env.platform == 'CPU'
    env.set_default()
    assert env.platform == 'CUDA'     # ... but the default (and only runnable) platform is the GPU


def test_constants_equal_the_reference_float32_values():
    ref = json.load(open(os.path.join(GOLDEN, 'reference_constants.json')))
    assert float(EPSILON0.value) == ref['EPSILON0']
    assert float(KB.value) == ref['KB'] and float(NA.value) == ref['NA']
    # the fixture was written in DOUBLE mode (float64 conversions); SINGLE mode rounds the same number to float32
    assert Quantity(1, kilocalorie_permol).convert_to(default_energy_unit).value == np.float32(ref['kcal_permol'])
    md.env.set_precision('DOUBLE')
    assert float(Quantity(1, kilocalorie_permol).convert_to(default_energy_unit).value) == pytest.approx(ref['kcal_permol'], rel=1e-14)
    md.env.set_precision('SINGLE')
    assert float(Quantity(1, kilojoule_permol).convert_to(default_energy_unit).value) == pytest.approx(ref['kj_permol'], rel=1e-6)
    assert float(RMIN_TO_SIGMA_FACTOR) == ref['RMIN_TO_SIGMA_FACTOR']
    assert float((Quantity(300, kelvin) * KB).convert_to(default_energy_unit).value) == pytest.approx(ref['kbt_300'], rel=1e-6)
    assert float(Quantity(0.91, nanometer).convert_to(default_length_unit).value) == pytest.approx(ref['nm_091_in_A'], rel=1e-6)
    assert synthetic.KCAL == float(np.float32(ref['kcal_permol']))
    assert coulomb_constant() == pytest.approx(1 / (4 * np.pi * ref['EPSILON0']), rel=1e-15)
    with pytest.raises(UnitDimensionDismatchedError):
        Quantity(1, angstrom).convert_to(femtosecond)


def test_topology_join_reproduces_reference_tables():
    """Same partner rules as topology.py:122-134,155-167,188-200 — checked against the tables the
    reference built for the same seeded system."""
    g = load_golden('mix_small_f64')
    s = synthetic.solvated_protein_box(2701, (30.0, 30.0, 30.0), protein_fraction=0.037, seed=11, n_res=10)
    t = s.topology()
    assert np.array_equal(t.bonded_particles, g['bonded'])
    assert np.array_equal(t.scaling_particles, g['scaling'])
    assert np.allclose(t.charges, g['charges']) and np.allclose(t.masses, g['masses'])
    assert np.allclose(s.lj_table(), g['lj_table'], rtol=1e-6)
    # the incremental API gives the same answer on the first molecules
    t2 = Topology()
    m = 130
    t2.add_particles([Particle(particle_id=i, particle_type=s.types[i], mass=s.masses[i], charge=s.charges[i]) for i in range(m)])
    for b in s.bonds[(s.bonds < m).all(1)]: t2.add_bond(list(b))
    for a in s.angles[(s.angles < m).all(1)]: t2.add_angle(list(a))
    for d in s.dihedrals[(s.dihedrals < m).all(1)]: t2.add_dihedral(list(d))
    t2.join()
    w = t2.bonded_particles.shape[1]
    assert np.array_equal(t2.bonded_particles[:100], g['bonded'][:100, :w])
    assert np.array_equal(t2.scaling_particles[:100], g['scaling'][:100, :w])


def test_topology_error_behaviour():
    t = Topology()
    t.add_particles([Particle(particle_id=i, particle_type='CA', mass=12, charge=0) for i in range(4)])
    t.add_bond([0, 1])
    with pytest.raises(ParticleConflictError):
        t.add_bond([1, 0])              # particle.py:53-58
    with pytest.raises(ParticleConflictError):
        t.add_bond([0, 7])
    with pytest.raises(ParticleConflictError):
        t.add_angle([0, 1, 0])
    t.add_dihedral([0, 1, 2, 3]); t.add_dihedral([0, 2, 1, 3])   # duplicates are silently merged (particle.py:71-80)
    t.join()
    assert list(t.scaling_particles[0]) == [3]
    with pytest.raises(ModifyJoinedTopologyError):
        t.add_bond([2, 3])
    with pytest.raises(ParticleConflictError):
        Topology.from_arrays(['A'] * 3, [1] * 3, [0] * 3, bonds=[[0, 1], [1, 0]])


def test_state_wrap_and_errors():
    t = Topology(); t.add_particles([Particle(particle_id=i, particle_type='CA', mass=12, charge=0) for i in range(2)])
    ens = md.Ensemble(t, np.eye(3) * 30)
    ens.state.set_positions(np.array([[0, 0, 0], [0, 21, 0]], dtype=np.float64))
    assert ens.state.positions.dtype == np.float32 and np.allclose(ens.state.positions[1], [0, -9, 0])
    with pytest.raises(ParticleLossError):
        ens.state.set_positions(np.array([[0, 0, 0], [0, 46.0, 0]]))
    with pytest.raises(ArrayDimError):
        ens.state.set_positions(np.zeros((3, 3)))
    with pytest.raises(TypeError):
        ens.state.set_positions([[0, 0, 0], [1, 1, 1]])
    with pytest.raises(CellListPoorDefinedError):
        md.Ensemble(t, np.eye(3) * 20)          # state.py:28: the default 12 A list needs L >= 24
    ens.state.set_velocities(np.ones((2, 3)))
    ens.update()
    assert ens.kinetic_energy == pytest.approx(0.5 * 24 * 3) and ens.potential_energy == 0   # test_ensemble.py:88-96
    pbc = np.eye(3) * 10.0
    assert np.allclose(unwrap_vec(np.array([6.0, -6.0, 4.0]), pbc, np.linalg.inv(pbc)), [-4, 4, 4])
    assert np.allclose(wrap_positions(np.array([[14.0, 0, 0]]), pbc, np.linalg.inv(pbc)), [[4, 0, 0]])


def test_constraint_protocol_without_a_gpu():
    lj = CharmmNonbondedConstraint({'CA': [1e-4, 3.0]}, cutoff_radius=Quantity(0.9, nanometer))
    assert lj.cutoff_radius == pytest.approx(9.0) and lj.force_id == 0 and lj.force_group == 0
    assert lj.forces is None and lj.potential_energy is None and lj.parent_ensemble is None
    assert repr(lj) == '<mdpy_b200.constraint.CharmmNonbondedConstraint object>'
    assert lj == lj and lj != CharmmNonbondedConstraint({'CA': [1e-4, 3.0]})
    for c in (lj, ElectrostaticConstraint(), ElectrostaticPMEConstraint()):
        with pytest.raises(NonBoundedError):
            c.update()
    lj.set_cutoff_radius(12)
    assert lj.cutoff_radius == 12
    pme = ElectrostaticPMEConstraint(cutoff_radius=12, ewald_error=1e-6)
    from math import erfc
    assert erfc(pme.alpha * 12) / 12 == pytest.approx(1e-6, rel=1e-6)
    assert pme.grid_for((61.7, 61.7, 61.7)) == (64, 64, 64) and fft_size(108.86) == 112 and fft_size(77.76) == 80


def test_duplicate_constraint_is_rejected_before_any_device_work():
    t = Topology(); t.add_particles([Particle(particle_id=0, particle_type='CA', mass=12, charge=0)])
    ens = md.Ensemble(t, np.eye(3) * 30)

    class Dummy:
        cutoff_radius = 0
        def bind_ensemble(self, ensemble): self.bound = ensemble
    d = Dummy()
    ens.add_constraints(d)
    with pytest.raises(ConstraintConflictError):   # test_ensemble.py:85-86
        ens.add_constraints(d)


def test_library_exports_every_symbol_the_header_declares():
    header = open(os.path.join(ROOT, 'include', 'mdpy_b200.h')).read()
    declared = sorted(set(re.findall(r'MDK_API\s+[\w\s\*]+?\b(mdk_\w+)\s*\(', header)))
    assert len(declared) >= 30
    assert sorted(_native.EXPORTS) == declared
    assert os.path.exists(_native.LIB_PATH), 'build with make -C mdpy_b200/csrc'
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    _native.load_library()      # argtypes for every export resolve


def test_no_cpu_fallback():
    """Without a GPU (this container) every compute entry fails loudly; nothing routes to the oracle."""
    import subprocess, sys
    code = ("import numpy as np, mdpy_b200 as md\n"
            "from mdpy_b200.core import Particle, Topology\n"
            "from mdpy_b200.constraint import CharmmNonbondedConstraint\n"
            "t=Topology(); t.add_particles([Particle(particle_id=0, particle_type='CA', mass=12, charge=0)])\n"
            "e=md.Ensemble(t, np.eye(3)*30)\n"
            "try:\n  e.add_constraints(CharmmNonbondedConstraint({'CA':[1e-4,3.0]}))\n  print('BOUND')\n"
            "except RuntimeError as ex:\n  print('RAISED', ex)\n"
            "import sys; print('ORACLE' if any(m.startswith('oracle') for m in sys.modules) else 'CLEAN')\n")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, '-c', code], cwd=ROOT, env=env, capture_output=True, text=True).stdout
    assert 'RAISED' in out and 'no CPU fallback' in out and 'CLEAN' in out
    # and the package sources never mention the oracle
    for root, _, files in os.walk(os.path.join(ROOT, 'mdpy_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                assert 'import oracle' not in open(os.path.join(root, f)).read() and 'from oracle' not in open(os.path.join(root, f)).read()


def test_synthetic_configs_have_the_survey_sizes():
    w = synthetic.water_box()
    assert w.num_particles == 23556 and w.box[0] == pytest.approx(61.7, abs=0.05)
    p = synthetic.solvated_protein_box()
    assert p.num_particles == 92224 and abs(float(p.charges.sum())) < 1e-3
    assert len(p.dihedrals) > 0 and len(p.impropers) > 0
    t = p.topology()
    assert t.bonded_particles.shape[1] <= 13
