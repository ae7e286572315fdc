"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the drop-in
Constraint / Integrator classes and therefore through the C ABI, against
  (1) the golden vectors the unmodified reference produced (tests/golden, oracle/make_golden.py),
  (2) the CPU oracle on the same seeded inputs,
  (3) size-independent properties at the benchmark sizes.
Tolerances are the north star's: bit-exact pair set, force relative RMS <= 1e-5, energy <= 1e-6.
"""
import json
import os

import numpy as np
import pytest

import mdpy_b200 as md
from conftest import GOLDEN, load_golden, rel_rms
from mdpy_b200 import _native, synthetic
from mdpy_b200.constraint import (CharmmAngleConstraint, CharmmBondConstraint, CharmmDihedralConstraint,
                                  CharmmImproperConstraint, CharmmNonbondedConstraint, CharmmVDWConstraint,
                                  ElectrostaticConstraint, ElectrostaticPMEConstraint)
from mdpy_b200.core import Particle, Topology
from mdpy_b200.error import CellListPoorDefinedError, NonBoundedError, ParticleLossError
from mdpy_b200.integrator import LangevinIntegrator, VerletIntegrator
from mdpy_b200.unit import (EPSILON0, KB, Quantity, RMIN_TO_SIGMA_FACTOR, angstrom, coulomb_constant,
                            default_energy_unit, default_force_unit, e, kilocalorie_permol, nanometer)
from mdpy_b200.utils import get_unit_vec
from oracle import cpu_oracle as ora
from oracle import spme

pytestmark = pytest.mark.gpu
FORCE_TOL, ENERGY_TOL = 1e-5, 1e-6
K_E = coulomb_constant()


def ensemble_from_golden(g, n_types=None):
    n = g['positions'].shape[0]
    topo = Topology.from_tables(['X'] * n, g['masses'], g['charges'], g['bonded'], g['scaling'])
    ens = md.Ensemble(topo, np.diag(g['box']))
    ens.state.set_positions(g['positions'].astype(np.float32))
    return ens


# ---------------------------------------------------------------------------------------------
# the reference's own known-answer tests, unchanged in substance
class TestCharmmNonbondedConstraintKAT:
    """mdpy/test/test_charmm_nonbonded_constraint.py:26-126."""

    def setup_method(self):
        g = load_golden('kat_f64')
        self.parameters = {k: v for k, v in json.loads(str(g['lj_param_json'])).items()}
        ps = [Particle(particle_id=i, particle_name=n, particle_type=t, molecule_type='ASN', mass=m, charge=0)
              for i, (n, t, m) in enumerate([('C', 'CA', 12), ('N', 'NY', 14), ('CA', 'CPT', 1), ('C', 'CA', 12)])]
        t = Topology(); t.add_particles(ps)
        self.p = np.array([[0, 0, 0], [0, 10, 0], [0, 21, 0], [0, 11, 0]], dtype=np.float32)
        self.ensemble = md.Ensemble(t, np.eye(3) * 30)
        self.ensemble.state.cell_list.set_cutoff_radius(5)
        self.ensemble.state.set_positions(self.p)
        self.constraint = CharmmNonbondedConstraint(self.parameters)
        self.golden = g

    def test_exceptions(self):
        with pytest.raises(NonBoundedError):
            self.constraint._check_bound_state()
        with pytest.raises(NonBoundedError):
            self.constraint.update()

    def test_bind_ensemble(self):
        self.ensemble.add_constraints(self.constraint)
        assert self.constraint._parent_ensemble.num_constraints == 1
        assert self.constraint._parameters_list[0, 0] == Quantity(0.07, kilocalorie_permol).convert_to(default_energy_unit).value
        assert self.constraint._parameters_list[1, 1] == pytest.approx(np.float32(1.85 * RMIN_TO_SIGMA_FACTOR * 2))
        self.constraint._check_bound_state()

    def test_update(self):
        self.constraint.set_cutoff_radius(Quantity(0.91, nanometer))
        self.ensemble.add_constraints(self.constraint)
        self.constraint.update()
        forces = self.constraint.forces
        assert forces.dtype == np.float32 and forces.shape == (4, 3)
        assert forces.sum() == pytest.approx(0, abs=1e-8)
        epsilon = np.sqrt(Quantity(0.07, kilocalorie_permol).convert_to(default_energy_unit).value *
                          Quantity(0.099, kilocalorie_permol).convert_to(default_energy_unit).value)
        sigma = (1.9924 + 1.86) * RMIN_TO_SIGMA_FACTOR
        r = 9
        scaled_r = sigma / r
        force_val = - 24 * epsilon / r * (2 * scaled_r**12 - scaled_r**6)
        force = force_val * -get_unit_vec(self.p[2, :] - self.p[0, :])
        for a in range(3):
            assert forces[0, a] == pytest.approx(force[a], abs=1e-8)
        epsilon = np.sqrt(Quantity(0.07, kilocalorie_permol).convert_to(default_energy_unit).value *
                          Quantity(0.2, kilocalorie_permol).convert_to(default_energy_unit).value)
        sigma = (1.9924 + 1.85) * RMIN_TO_SIGMA_FACTOR
        energy_ref = 4 * epsilon * (sigma**12 - sigma**6)
        assert self.constraint.potential_energy == pytest.approx(np.float32(energy_ref), abs=1e-3)
        # and the reference's own numbers
        assert rel_rms(forces, self.golden['lj_forces']) < FORCE_TOL
        assert self.constraint.potential_energy == pytest.approx(float(self.golden['lj_energy']), rel=ENERGY_TOL)


class TestElectrostaticConstraintKAT:
    """mdpy/test/test_electrostatic_constraint.py:26-95."""

    def setup_method(self):
        ps = [Particle(particle_id=i, particle_type=t, particle_name=n, molecule_type='ASN', mass=m, charge=q)
              for i, (t, n, m, q) in enumerate([('C', 'CA', 12, 1), ('N', 'NY', 14, 2), ('CA', 'CPT', 1, 0), ('C', 'CA', 12, 0)])]
        t = Topology(); t.add_particles(ps)
        self.pbc = np.diag(np.ones(3) * 100)
        self.p = np.array([[0, 0, 0], [0, 10, 0], [0, 21, 0], [0, 11, 0]], dtype=np.float64)
        self.ensemble = md.Ensemble(t, np.eye(3) * 30)
        self.ensemble.state.cell_list.set_cutoff_radius(12)
        self.ensemble.state.set_positions(self.p)
        self.constraint = ElectrostaticConstraint()

    def test_exceptions(self):
        with pytest.raises(NonBoundedError):
            self.constraint._check_bound_state()

    def test_update(self):
        self.ensemble.state.set_pbc_matrix(self.pbc)
        self.ensemble.add_constraints(self.constraint)
        assert self.constraint._parent_ensemble.num_constraints == 1
        self.constraint.update()
        forces = self.constraint.forces
        assert forces[2, 0] == 0 and forces[3, 1] == 0
        k = Quantity(4 * np.pi) * EPSILON0
        force_val = - Quantity(1, e) * Quantity(2, e) / k / Quantity(10, angstrom)**2
        force1 = force_val.convert_to(default_force_unit).value * get_unit_vec(np.array([0, 10, 0], dtype=np.float32))
        for a in range(3):
            assert forces[0, a] == pytest.approx(force1[a])
            assert forces[1, a] == pytest.approx(-force1[a])
        energy = Quantity(1, e) * Quantity(2, e) / k / Quantity(10, angstrom)
        assert self.constraint.potential_energy == pytest.approx(energy.convert_to(default_energy_unit).value)
        g = load_golden('kat_f64')
        assert rel_rms(forces, g['coul_forces']) < FORCE_TOL
        assert self.constraint.potential_energy == pytest.approx(float(g['coul_energy']), rel=ENERGY_TOL)


def test_error_behaviour_matches_reference():
    t = Topology(); t.add_particles([Particle(particle_id=i, particle_type='CA', mass=12, charge=0) for i in range(2)])
    ens = md.Ensemble(t, np.eye(3) * 30)
    with pytest.raises(CellListPoorDefinedError):          # test_cell_list.py:31-39
        ens.state.cell_list.set_cutoff_radius(0)
    with pytest.raises(CellListPoorDefinedError):
        ens.state.cell_list.set_cutoff_radius(24)
    with pytest.raises(ParticleLossError):                 # test_pbc.py
        ens.state.set_positions(np.array([[0, 0, 0], [0, 61.0, 0]], dtype=np.float32))
    lj = CharmmNonbondedConstraint({'CA': [1e-4, 3.0]}, cutoff_radius=24)
    with pytest.raises(CellListPoorDefinedError):          # ensemble.py:49-50 -> cell_list.py:64-69
        ens.add_constraints(lj)


# ---------------------------------------------------------------------------------------------
# many-body parity against the reference's golden outputs (DOUBLE mode)
@pytest.mark.parametrize('name,lj_key,el_key', [('mix_small_f64', 'lj', 'coul'),
                                                ('config1_f64', 'CharmmNonbondedConstraint', 'ElectrostaticConstraint')])
def test_lj_and_bare_coulomb_match_reference(name, lj_key, el_key):
    g = load_golden(name)
    ens = ensemble_from_golden(g)
    lj = CharmmNonbondedConstraint(g['lj_table'], cutoff_radius=float(g['rc']))
    el = ElectrostaticConstraint()
    ens.add_constraints(lj, el)
    # float64 oracle on the float32 positions the device sees (the oracle itself equals the reference to
    # 1e-10 on the reference's float64 inputs, tests/test_oracle.py); its |pair energy| sums give the scale
    # a cancelling total has to be judged on (DESIGN.md "tolerances")
    t = ora.nonbonded_bruteforce(ens.state.positions, g['box'], g['lj_table'], g['charges'], g['bonded'], g['scaling'],
                                 rc_lj=float(g['rc']), coul_mode=2, k_e=K_E)
    lj.update()
    assert rel_rms(lj.forces, g[lj_key + '_forces']) < FORCE_TOL
    assert abs(lj.potential_energy - float(g[lj_key + '_energy'])) < ENERGY_TOL * t['e_lj_abs']
    if name == 'config1_f64':   # clash-dominated total: plain relative error works too
        assert lj.potential_energy == pytest.approx(float(g[lj_key + '_energy']), rel=ENERGY_TOL)
    el.update()
    # the bare minimum-image sum is discontinuous where a pair sits at L/2 (exact ties in the PDB's
    # 3-decimal coordinates of config 1): compare on identical inputs, i.e. the oracle at fp32 positions
    assert rel_rms(el.forces, t['f_coul']) < FORCE_TOL
    assert abs(el.potential_energy - t['e_coul']) < ENERGY_TOL * t['e_coul_abs']
    if name == 'mix_small_f64':  # random coordinates: no ties, the reference's own output is reproduced
        assert rel_rms(el.forces, g[el_key + '_forces']) < FORCE_TOL
        assert el.potential_energy == pytest.approx(float(g[el_key + '_energy']), rel=ENERGY_TOL)
    # Ensemble.update sums the same thing (fused single evaluation) — ensemble.py:53-61
    ens.update()
    assert rel_rms(ens.forces, t['f_lj'] + t['f_coul']) < FORCE_TOL
    assert abs(ens.potential_energy - t['e_lj'] - t['e_coul']) < ENERGY_TOL * (t['e_lj_abs'] + t['e_coul_abs'])


def test_config1_bonded_terms_match_reference():
    g = load_golden('config1_f64')
    n = g['positions'].shape[0]
    topo = Topology.from_tables(['X'] * n, g['masses'], g['charges'], g['bonded'], g['scaling'],
                                bonds=g['CharmmBondConstraint_idx'], angles=g['CharmmAngleConstraint_idx'],
                                dihedrals=g['CharmmDihedralConstraint_idx'], impropers=g['CharmmImproperConstraint_idx'])
    ens = md.Ensemble(topo, np.diag(g['box']))
    ens.state.set_positions(g['positions'].astype(np.float32))
    cs = dict(CharmmBondConstraint=CharmmBondConstraint(g['CharmmBondConstraint_par']),
              CharmmAngleConstraint=CharmmAngleConstraint(g['CharmmAngleConstraint_par']),
              CharmmDihedralConstraint=CharmmDihedralConstraint(g['CharmmDihedralConstraint_par']),
              CharmmImproperConstraint=CharmmImproperConstraint(g['CharmmImproperConstraint_par']))
    ens.add_constraints(*cs.values())
    # The gates (1e-5 force RMS, 1e-6 energy) are taken on IDENTICAL inputs: oracle/bonded.py (equal to the reference
    # to 1e-10 on the reference's float64 inputs, tests/test_oracle.py) evaluated at the float32 positions the
    # device holds.  The PDB's 3-decimal coordinates are not float32 numbers, and a 1e-6 A shift of a stiff
    # bond moves its force by ~1e-4 of itself; against the golden values themselves the distance is recorded.
    from oracle import bonded
    x32 = ens.state.positions.astype(np.float64)
    fns = dict(CharmmBondConstraint=bonded.bonds, CharmmAngleConstraint=bonded.angles, CharmmImproperConstraint=bonded.impropers)
    for name, c in cs.items():
        c.update()
        if name == 'CharmmDihedralConstraint':   # the reference's dihedral force is not the gradient (DESIGN Q12)
            e_ref = bonded.dihedral_energy(x32, g['box'], g[name + '_idx'], g[name + '_par'])
            assert c.potential_energy == pytest.approx(e_ref, rel=ENERGY_TOL), name
            assert c.potential_energy == pytest.approx(float(g[name + '_energy']), rel=1e-5), name
            continue
        f_ref, e_ref = fns[name](x32, g['box'], g[name + '_idx'], g[name + '_par'])
        assert rel_rms(c.forces, f_ref) < FORCE_TOL, name
        assert c.potential_energy == pytest.approx(e_ref, rel=ENERGY_TOL), name
        # and the reference's own numbers (float64 inputs): within what the input rounding explains
        assert rel_rms(c.forces, g[name + '_forces']) < 1e-4, name
        assert c.potential_energy == pytest.approx(float(g[name + '_energy']), rel=1e-5), name


def test_pair_set_is_bit_exact():
    """Bit-exact neighbour pair set vs the canonical fp32 criterion, on a system with
    exclusions and on the case where the reference itself drops pairs."""
    for name in ('mix_small_f64', 'q1_case_f64'):
        g = load_golden(name)
        n = g['positions'].shape[0]
        if 'bonded' not in g:
            g['bonded'] = -np.ones((n, 1), dtype=np.int32); g['scaling'] = g['bonded']
            g['masses'] = np.ones(n); g['charges'] = np.zeros(n)
        ens = ensemble_from_golden(g)
        lj = CharmmNonbondedConstraint(g['lj_table'], cutoff_radius=float(g['rc']))
        ens.add_constraints(lj)
        got = lj.neighbor_pairs()
        want = ora.pair_set_f32(ens.state.positions, np.float32(g['box']), float(g['rc']), g['bonded'])
        assert got.shape == want.shape, name
        assert np.array_equal(got, want), name


@pytest.mark.parametrize('opts', [dict(far_split=0), dict(far_flush=32), dict(pair_units_per_warp=2), dict(far_flush=64, unit_waves=1)])
def test_pair_set_and_forces_do_not_depend_on_list_options(opts):
    """The list order (skin-shell atoms last, early flushes of their staging buffer), the work-unit granularity and the
    block lifetime of the pair kernel are execution options: same pair set bit for bit, same forces to float32 summation order."""
    g = load_golden('mix_small_f64')
    def build(options):
        ens = ensemble_from_golden(g)
        lj = CharmmNonbondedConstraint(g['lj_table'], cutoff_radius=float(g['rc']))
        ens.add_constraints(lj)
        dev = _native.context_of(ens).dev
        dev.set_nlist(4.0)            # a thick skin: a third of every list is far class
        for k, v in options.items():
            dev.set_option(k, v)
        ens.update()
        return lj.neighbor_pairs(), lj.forces.copy(), lj.potential_energy
    p0, f0, e0 = build({})
    p1, f1, e1 = build(opts)
    want = ora.pair_set_f32(g['positions'].astype(np.float32), np.float32(g['box']), float(g['rc']), g['bonded'])
    assert np.array_equal(p0, want) and np.array_equal(p1, want)
    # (the LJ total of this box is a small difference of large pair terms: 2e-7 of it is float32 summation order)
    assert rel_rms(f1, f0) < 1e-6 and e1 == pytest.approx(e0, rel=1e-5)


def test_q1_case_matches_bruteforce_truth_not_the_reference_defect():
    g = load_golden('q1_case_f64')
    n = g['positions'].shape[0]
    g['bonded'] = -np.ones((n, 1), dtype=np.int32); g['scaling'] = g['bonded']
    g['masses'] = np.ones(n); g['charges'] = np.zeros(n)
    ens = ensemble_from_golden(g)
    lj = CharmmNonbondedConstraint(g['lj_table'], cutoff_radius=float(g['rc']))
    ens.add_constraints(lj)
    lj.update()
    t = ora.nonbonded_bruteforce(ens.state.positions, g['box'], g['lj_table'], g['charges'], g['bonded'], g['scaling'],
                                 rc_lj=float(g['rc']))
    assert rel_rms(lj.forces, t['f_lj']) < FORCE_TOL
    assert lj.potential_energy == pytest.approx(t['e_lj'], rel=ENERGY_TOL)
    # the reference misses pairs here (its energy differs by more than the tolerance)
    assert abs(float(g['lj_energy']) - t['e_lj']) / abs(t['e_lj']) > ENERGY_TOL


# ---------------------------------------------------------------------------------------------
# new physics (parity unpinned by the reference): float64 oracle on the same inputs
def test_switched_lj_matches_oracle():
    g = load_golden('mix_small_f64')
    ens = ensemble_from_golden(g)
    lj = CharmmVDWConstraint(g['lj_table'], cutoff_radius=12.0, switch_radius=10.0)
    ens.add_constraints(lj)
    lj.update()
    t = ora.nonbonded_bruteforce(ens.state.positions, g['box'], g['lj_table'], g['charges'], g['bonded'], g['scaling'],
                                 rc_lj=12.0, r_on=10.0)
    assert rel_rms(lj.forces, t['f_lj']) < FORCE_TOL
    assert abs(lj.potential_energy - t['e_lj']) < ENERGY_TOL * t['e_lj_abs']


@pytest.mark.parametrize('order,grid,alpha', [(4, (32, 32, 32), 0.30), (6, (48, 48, 48), 0.36)])
def test_pme_matches_float64_spme_restatement(order, grid, alpha):
    """CUDA PME vs the float64 SPME restatement with identical alpha / mesh / order."""
    g = load_golden('mix_small_f64')
    ens = ensemble_from_golden(g)
    pme = ElectrostaticPMEConstraint(cutoff_radius=12.0, alpha=alpha, grid=grid, order=order)
    ens.add_constraints(pme)
    pme.update()
    f, en = spme.pme_total(ens.state.positions, g['charges'], g['box'], g['bonded'], grid, order, alpha, 12.0, K_E)
    assert rel_rms(pme.forces, f) < FORCE_TOL
    # the PME total is a small difference of large terms (self, excluded pairs): the 1e-6 is taken
    # against the magnitude of the terms, and each device term is checked on its own as well
    scale = sum(abs(en[k]) for k in ('direct', 'excl', 'recip', 'self_bg'))
    assert abs(pme.potential_energy - en['total']) < ENERGY_TOL * scale
    e = pme._ctx.dev.last_energies() if False else pme._ctx.compute(pme.terms)
    assert e[2] == pytest.approx(en['recip'], rel=ENERGY_TOL)
    assert e[4] == pytest.approx(en['excl'], rel=ENERGY_TOL)
    assert e[3] == pytest.approx(en['self_bg'], rel=ENERGY_TOL)
    assert abs(e[1] - en['direct']) < ENERGY_TOL * scale


def test_pme_converges_to_exact_ewald():
    """Tight parameters: the whole CUDA electrostatics against the converged Ewald sum."""
    s = synthetic.solvated_protein_box(1471, (24.5, 24.5, 24.5), protein_fraction=0.068, seed=12, n_res=10)
    ens = s.ensemble(cutoff=12.0, pme=False, bonded=False)
    pme = ElectrostaticPMEConstraint(cutoff_radius=12.0, alpha=0.42, grid=(64, 64, 64), order=8)
    ens.add_constraints(pme)
    pme.update()
    topo = ens.topology
    f_ex, e_ex = ora.ewald_exact(ens.state.positions, s.charges, s.box, topo.bonded_particles, K_E)
    assert rel_rms(pme.forces, f_ex) < FORCE_TOL
    e = pme._ctx.compute(pme.terms)
    assert abs(pme.potential_energy - e_ex) < ENERGY_TOL * np.abs(e[1:5]).sum()


def test_small_mesh_fft_kernels_match_cufft():
    """Power-of-two meshes up to 64 per axis run the fused mesh kernels (mdk_pme.cu: k_mesh_*); the cuFFT
    chain on the same mesh is the check (option pme_cufft), on a non-cubic mesh so that every axis has its
    own length."""
    g = load_golden('mix_small_f64')
    res = []
    for use_cufft in (0, 1):
        ens = ensemble_from_golden(g)
        pme = ElectrostaticPMEConstraint(cutoff_radius=9.0, alpha=0.32, grid=(32, 64, 16), order=4)
        ens.add_constraints(pme)
        pme._ctx.dev.set_option('pme_cufft', use_cufft)
        pme.update()
        res.append((pme.forces.astype(np.float64), pme.potential_energy, pme._ctx.dev.last_energies()[2]))
    assert rel_rms(res[0][0], res[1][0]) < 2e-6
    assert res[0][2] == pytest.approx(res[1][2], rel=1e-6)
    assert res[0][1] == pytest.approx(res[1][1], rel=1e-6, abs=1e-6 * abs(res[1][2]))


def test_results_are_bitwise_reproducible():
    g = load_golden('mix_small_f64')
    out = []
    for _ in range(2):
        ens = ensemble_from_golden(g)
        lj = CharmmNonbondedConstraint(g['lj_table'], cutoff_radius=9.0)
        pme = ElectrostaticPMEConstraint(cutoff_radius=9.0, grid=(32, 32, 32))
        ens.add_constraints(lj, pme)
        ens.update()
        out.append((ens.forces.copy(), ens.potential_energy))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


# ---------------------------------------------------------------------------------------------
# integrators
def test_verlet_matches_reference_trajectory():
    g = load_golden('verlet_small_f64')
    ens = ensemble_from_golden(g)
    ens.add_constraints(CharmmNonbondedConstraint(g['lj_table'], cutoff_radius=float(g['rc'])), ElectrostaticConstraint())
    ens.state.set_velocities(g['verlet_v0'].astype(np.float32))
    integ = VerletIntegrator(float(g['verlet_dt']))
    integ.integrate(ens, int(g['verlet_steps']))
    assert np.abs(integ.cur_positions - g['verlet_cur']).max() < 2e-5
    assert np.abs(ens.state.positions - g['verlet_positions']).max() < 2e-5
    assert np.abs(ens.state.velocities - g['verlet_velocities']).max() < 1e-5   # sic: half the true velocity (Q4)
    # continuing in a second call uses the cached history like the reference (integrator.py:43-49)
    assert integ.is_cached


def test_verlet_nve_energy_is_bounded_short():
    """Bounded NVE drift on a 23k-atom water box (own kinetic energy, textbook velocities; Q4)."""
    s = synthetic.water_box(7852, 20260001)
    ens = s.ensemble(cutoff=9.0, pme=True, grid=(64, 64, 64))
    # take the lattice start off its clashes first
    LangevinIntegrator(0.25, 300, 0.05, seed=3).integrate(ens, 400)
    integ = VerletIntegrator(0.5, reference_quirks=False)
    e_tot, ke = [], []
    for _ in range(20):
        integ.integrate(ens, 50)
        e_tot.append(ens.total_energy); ke.append(ens.kinetic_energy)
    e_tot = np.array(e_tot)
    assert np.isfinite(e_tot).all()
    assert np.abs(e_tot - e_tot[0]).max() < 1e-2 * np.mean(ke)


def test_langevin_equipartition_quick():
    s = synthetic.water_box(2000, 5, box=np.full(3, 39.2))
    ens = s.ensemble(cutoff=9.0, pme=True, grid=(40, 40, 40))
    integ = LangevinIntegrator(0.5, 300, 0.02, seed=11)
    integ.integrate(ens, 1500)
    temps = []
    kb = float((Quantity(1, md.unit.kelvin) * KB).convert_to(default_energy_unit).value)
    for _ in range(10):
        integ.integrate(ens, 100)
        temps.append(2 * ens.kinetic_energy / (3 * s.num_particles) / kb)
    assert 270 < np.mean(temps) < 330


def _small_water(seed=5):
    s = synthetic.water_box(2000, seed, box=np.full(3, 39.2))
    ens = s.ensemble(cutoff=9.0, pme=True, grid=(40, 40, 40))
    LangevinIntegrator(0.25, 300, 0.05, seed=3).integrate(ens, 300)   # off the lattice clashes
    return s, ens


def test_langevin_host_state_calls_continue_the_trajectory():
    """integrate(ens, 1) x 12 (host State in and out at every call, mdk_step_langevin_host) walks the same
    trajectory as integrate(ens, 12): the device keeps its float64 state while the host copy is unchanged."""
    _, ens_a = _small_water()
    _, ens_b = _small_water()
    assert np.array_equal(ens_a.state.positions, ens_b.state.positions)
    ia, ib = LangevinIntegrator(1.0, 300, 0.001, seed=7), LangevinIntegrator(1.0, 300, 0.001, seed=7)
    ia.integrate(ens_a, 12)
    for _ in range(12):
        ib.integrate(ens_b, 1)
    assert np.abs(ens_a.state.positions - ens_b.state.positions).max() < 1e-4
    assert np.abs(ens_a.state.velocities - ens_b.state.velocities).max() < 1e-5
    assert ens_a.potential_energy == pytest.approx(ens_b.potential_energy, rel=1e-6)
    assert ens_b.state.positions.dtype == np.float32 and ens_b.state.positions.shape == (6000, 3)
    # the energies the step call reports are those of a fresh evaluation at the returned positions
    e_step = ens_b.potential_energy
    ens_b.update()
    assert ens_b.potential_energy == pytest.approx(e_step, rel=1e-6)


def test_langevin_host_state_edits_are_honoured():
    """In-place edits of ensemble.state arrays (no set_positions call, so no revision bump) reach the device."""
    s, ens = _small_water()
    integ = LangevinIntegrator(0.5, 300, 0.001, seed=9)
    integ.integrate(ens, 3)
    x = ens.state.positions
    moved = x[0].copy() + np.float32([0.4, 0.0, 0.0])
    x[0] = moved
    ens.state.velocities[...] = 0
    integ.integrate(ens, 1)
    assert np.abs(ens.state.positions[0] - moved).max() < 0.2
    assert np.abs(ens.state.velocities).max() < 0.05          # one 0.5 fs step from rest
    # an atom two box lengths away is lost, as in utils/pbc.py:29-34
    ens.state.positions[5, 1] = 2.5 * 39.2
    with pytest.raises(ParticleLossError):
        integ.integrate(ens, 1)


# benchmark sizes (configs 1-4): tests/test_gpu_benchmark_parity.py
