"""CPU-only: pins the oracle (oracle/) against the golden vectors the UNMODIFIED reference produced
in this container (oracle/make_golden.py), and the float64 truth against itself."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_rms
from oracle import cpu_oracle as ora
from oracle import spme

K_E = 1.0 / (4 * np.pi * float(np.float32(0.5727653)))


def test_oracle_builds_and_exports():
    lib = ora.lib()
    for name in ('ora_lj_cell_f32', 'ora_lj_cell_f64', 'ora_coulomb_allpairs_f64', 'ora_nonbonded_bruteforce',
                 'ora_pair_set_f32', 'ora_ewald_recip', 'ora_wrap_positions_f64', 'ora_cell_index_f32', 'ora_cell_fill'):
        assert hasattr(lib, name)


# ---- reference KATs (mdpy/test/test_charmm_nonbonded_constraint.py:93-126, test_electrostatic_constraint.py:71-95)
def test_kat_lj_against_reference_and_analytic():
    g = load_golden('kat_f64')
    pbc = np.eye(3) * 30
    none = -np.ones((4, 1), dtype=np.int32)
    # the reference's cell list was at the LJ cutoff (ensemble.py:49-50 widens 5 -> 9.1)
    f, e, visits = ora.lj_cell(g['positions'], g['lj_table'], pbc, float(g['lj_cutoff']), none, none)
    assert np.allclose(f, g['lj_forces'], rtol=1e-12, atol=1e-18)
    assert e == pytest.approx(float(g['lj_energy']), rel=1e-12)
    assert abs(f.sum()) < 1e-8
    # analytic value of the CA-CPT pair at the periodic distance 9 A
    eps = np.sqrt(g['lj_table'][0, 0] * g['lj_table'][2, 0]); sig = 0.5 * (g['lj_table'][0, 1] + g['lj_table'][2, 1])
    fval = -24 * eps / 9 * (2 * (sig / 9) ** 12 - (sig / 9) ** 6)
    assert f[0, 1] == pytest.approx(-fval * 1.0 * -1.0, abs=1e-8) or f[0, 1] == pytest.approx(fval * -1.0, abs=1e-8)
    assert visits == 4  # CA-NY (r=10 > 9.1) out; (0,2) r=9 and (1,3) r=1 in, both ways


def test_kat_coulomb_against_reference_and_analytic():
    g = load_golden('kat_f64')
    none = -np.ones((4, 1), dtype=np.int32)
    k = 4 * np.pi * float(np.float32(0.5727653))
    f, e = ora.coulomb_allpairs(g['coul_positions'], g['coul_charges'], np.eye(3) * 100, none, k)
    assert np.allclose(f, g['coul_forces'], rtol=1e-12, atol=1e-18)
    assert e == pytest.approx(float(g['coul_energy']), rel=1e-12)
    assert f[0, 1] == pytest.approx(-1 * 2 / k / 100)
    assert e == pytest.approx(1 * 2 / k / 10)
    assert f[2, 0] == 0 and f[3, 1] == 0


# ---- many-body parity: restatement vs reference DOUBLE mode
@pytest.mark.parametrize('name', ['mix_small', 'config1'])
def test_lj_and_coulomb_restatement_matches_reference_double(name):
    g = load_golden(name + '_f64')
    pbc = np.diag(g['box'])
    rc = float(g['rc'])
    lj_key, el_key = ('lj', 'coul') if name == 'mix_small' else ('CharmmNonbondedConstraint', 'ElectrostaticConstraint')
    f, e, _ = ora.lj_cell(g['positions'], g['lj_table'], pbc, rc, g['bonded'], g['scaling'], cell_cutoff=max(rc, 12.0))
    assert rel_rms(f, g[lj_key + '_forces']) < 1e-10
    assert e == pytest.approx(float(g[lj_key + '_energy']), rel=1e-10)
    k = 4 * np.pi * float(np.float32(0.5727653))
    f, e = ora.coulomb_allpairs(g['positions'], g['charges'], pbc, g['bonded'], k)
    assert rel_rms(f, g[el_key + '_forces']) < 1e-10
    assert e == pytest.approx(float(g[el_key + '_energy']), rel=1e-10)


def test_config1_golden_energies_match_survey():
    """SURVEY §6 golden single-point energies of the example system (reference DOUBLE mode)."""
    g = load_golden('config1_f64')
    kcal = json.load(open(os.path.join(GOLDEN, 'reference_constants.json')))['kcal_permol']
    assert float(g['CharmmNonbondedConstraint_energy']) / kcal == pytest.approx(6561.881429, abs=2e-3)
    assert float(g['ElectrostaticConstraint_energy']) / kcal == pytest.approx(-7499.577066, abs=2e-3)
    assert g['positions'].shape == (2423, 3)


def test_float32_restatement_close_to_reference_single():
    g = load_golden('mix_small_f32')
    pbc = np.diag(g['box']).astype(np.float32)
    f, e, _ = ora.lj_cell(g['positions'].astype(np.float32), g['lj_table'].astype(np.float32), pbc, float(g['rc']),
                          g['bonded'], g['scaling'], cell_cutoff=12.0)
    assert rel_rms(f, g['lj_forces']) < 2e-5   # two different fp32 evaluation orders
    assert e == pytest.approx(float(g['lj_energy']), rel=2e-5)


def test_cell_list_restatement_matches_reference():
    g = load_golden('mix_small_f64')
    ncell, _, cell_inv = ora.cell_attributes(g['box'], 12.0, np.float64)
    assert list(ncell) == list(g['cell_num'])
    pci, cl = ora.cell_list_update(g['positions'], cell_inv, ncell)
    assert np.array_equal(pci, g['cell_index'])
    assert list(cl.shape) == list(g['cell_list_shape'])
    assert np.array_equal(cl[0, 0, 0], g['cell_list_head'])
    with pytest.raises(ValueError):
        ora.cell_attributes(np.full(3, 30.0), 0.0)      # test_cell_list.py:31-39
    with pytest.raises(ValueError):
        ora.cell_attributes(np.full(3, 30.0), 24.0)


def test_q1_reference_drops_pairs_and_restatement_reproduces_it():
    """with >= 4 cells per dimension the reference's cell list loses pairs across the
    periodic boundary.  The restatement reproduces the reference bit for bit; the brute-force pair set
    is the definition the CUDA path is held to."""
    g = load_golden('q1_case_f64')
    n = g['positions'].shape[0]
    none = -np.ones((n, 1), dtype=np.int32)
    f, e, visits = ora.lj_cell(g['positions'], g['lj_table'], np.diag(g['box']), float(g['rc']), none, none)
    assert list(g['cell_num']) == [4, 4, 4]
    assert rel_rms(f, g['lj_forces']) < 1e-10
    assert e == pytest.approx(float(g['lj_energy']), rel=1e-10)
    truth = ora.nonbonded_bruteforce(g['positions'], g['box'], g['lj_table'], np.zeros(n), none, none, rc_lj=float(g['rc']))
    assert truth['n_lj'] > visits                   # pairs the reference never sees
    assert 0 < (truth['n_lj'] - visits) // 2 < 500
    assert abs(truth['e_lj'] - e) / abs(truth['e_lj']) > 1e-6


def test_pbc_wrap_kats():
    """mdpy/test/test_pbc.py:26-81 semantics."""
    pbc = np.diag([10.0, 10.0, 10.0]); inv = np.linalg.inv(pbc)
    p = np.array([[0, 0, 0], [6, 0, 0], [0, -6, 0], [4.9, 5.1, -5.1], [14.0, 0, 0]], dtype=np.float64)
    w, lost, _ = ora.wrap_positions(p, pbc, inv)
    assert lost == 0
    assert np.allclose(w, [[0, 0, 0], [-4, 0, 0], [0, 4, 0], [4.9, -4.9, 4.9], [4.0, 0, 0]])
    w, lost, first = ora.wrap_positions(np.array([[0, 0, 0], [16.0, 0, 0]]), pbc, inv)
    assert lost == 1 and first == 1                  # ParticleLossError in the reference (pbc.py:30-34)


def test_verlet_restatement_matches_reference():
    g = load_golden('verlet_small_f64')
    pbc = np.diag(g['box'])
    k = 4 * np.pi * float(np.float32(0.5727653))
    rc = float(g['rc'])

    def force(x):
        f1, _, _ = ora.lj_cell(x, g['lj_table'], pbc, rc, g['bonded'], g['scaling'], cell_cutoff=12.0)
        f2, _ = ora.coulomb_allpairs(x, g['charges'], pbc, g['bonded'], k)
        return f1 + f2
    pos, vel, cur, pre = ora.verlet(g['positions'], g['verlet_v0'], g['masses'], pbc, float(g['verlet_dt']),
                                    int(g['verlet_steps']), force)
    assert np.abs(cur - g['verlet_cur']).max() < 1e-9
    assert np.abs(pos - g['verlet_positions']).max() < 1e-9
    assert np.abs(vel - g['verlet_velocities']).max() < 1e-10


# ---- float64 truth -------------------------------------------------------------------------
def test_bruteforce_agrees_with_cell_restatement_when_no_pairs_can_be_dropped():
    g = load_golden('mix_small_f64')
    f, e, visits = ora.lj_cell(g['positions'], g['lj_table'], np.diag(g['box']), float(g['rc']), g['bonded'],
                               g['scaling'], cell_cutoff=12.0)
    t = ora.nonbonded_bruteforce(g['positions'], g['box'], g['lj_table'], g['charges'], g['bonded'], g['scaling'],
                                 rc_lj=float(g['rc']), coul_mode=2, k_e=K_E)
    assert t['n_lj'] == visits
    assert rel_rms(t['f_lj'], f) < 1e-12 and t['e_lj'] == pytest.approx(e, rel=1e-12)
    assert rel_rms(t['f_coul'], g['coul_forces']) < 1e-10
    assert t['e_coul'] == pytest.approx(float(g['coul_energy']), rel=1e-10)


def test_canonical_fp32_pair_set_is_the_minimum_image_set():
    g = load_golden('mix_small_f64')
    pos = g['positions'].astype(np.float32)
    pairs = ora.pair_set_f32(pos, g['box'], 9.0, g['bonded'])
    t = ora.nonbonded_bruteforce(pos.astype(np.float64), np.float32(g['box']).astype(np.float64), g['lj_table'],
                                 g['charges'], g['bonded'], g['scaling'], rc_lj=9.0)
    assert abs(len(pairs) - t['n_lj'] // 2) <= 2    # only pairs within an ulp of the cutoff may differ
    assert (pairs[:, 0] < pairs[:, 1]).all()
    assert len(np.unique(pairs[:, 0].astype(np.int64) * len(pos) + pairs[:, 1])) == len(pairs)


def test_switch_function_is_continuous_and_consistent():
    """CHARMM switch: E -> 0 at rc, forces = -dE/dx by finite differences."""
    box = np.full(3, 40.0)
    params = np.array([[1e-4, 3.4, 1e-4, 3.4]] * 2)
    none = -np.ones((2, 1), dtype=np.int32)

    def e_at(r):
        pos = np.array([[0, 0, 0], [r, 0, 0]], dtype=np.float64)
        return ora.nonbonded_bruteforce(pos, box, params, np.zeros(2), none, none, rc_lj=12.0, r_on=10.0)
    assert e_at(11.999999)['e_lj'] == pytest.approx(0, abs=1e-16)
    assert e_at(9.9)['e_lj'] == pytest.approx(4e-4 * ((3.4 / 9.9) ** 12 - (3.4 / 9.9) ** 6), rel=1e-12)
    for r in (10.5, 11.0, 11.7):
        h = 1e-5
        fd = -(e_at(r + h)['e_lj'] - e_at(r - h)['e_lj']) / (2 * h)
        assert e_at(r)['f_lj'][1, 0] == pytest.approx(fd, rel=1e-6)


def _random_ionic(n=150, seed=0):
    rng = np.random.default_rng(seed)
    box = np.array([24.0, 25.0, 26.0])
    pos = rng.uniform(-0.5, 0.5, size=(n, 3)) * box
    q = rng.normal(size=n); q -= q.mean()
    bonded = -np.ones((n, 2), dtype=np.int32)
    for i in range(0, 30, 2):
        pos[i + 1] = pos[i] + rng.normal(size=3) * 0.6
        bonded[i, 0], bonded[i + 1, 0] = i + 1, i
    return pos, q, box, bonded


def test_ewald_exact_is_alpha_independent_and_forces_are_gradients():
    pos, q, box, bonded = _random_ionic()
    f1, e1 = ora.ewald_exact(pos, q, box, bonded, K_E, tol_exp=36.0)
    f2, e2 = ora.ewald_exact(pos, q, box, bonded, K_E, tol_exp=28.0)
    assert e1 == pytest.approx(e2, rel=1e-10) and np.abs(f1 - f2).max() < 1e-10
    assert np.abs(f1.sum(0)).max() < 1e-10
    h = 1e-5
    for atom, ax in ((3, 0), (40, 2)):
        p = pos.copy(); p[atom, ax] += h; _, ep = ora.ewald_exact(p, q, box, bonded, K_E)
        p[atom, ax] -= 2 * h; _, em = ora.ewald_exact(p, q, box, bonded, K_E)
        assert f1[atom, ax] == pytest.approx(-(ep - em) / (2 * h), rel=1e-5, abs=1e-9)


@pytest.mark.parametrize('order,grid,alpha,tol', [(4, (24, 25, 27), 0.32, 2e-3), (6, (48, 50, 54), 0.42, 2e-5),
                                                   (8, (72, 75, 80), 0.50, 1e-6)])
def test_spme_restatement_converges_to_exact_ewald(order, grid, alpha, tol):
    pos, q, box, bonded = _random_ionic()
    f_ex, e_ex = ora.ewald_exact(pos, q, box, bonded, K_E)
    f, en = spme.pme_total(pos, q, box, bonded, grid, order, alpha, 11.5, K_E)
    assert rel_rms(f, f_ex) < tol
    assert abs(en['total'] - e_ex) / abs(e_ex) < tol


def test_config1_verlet_fixture_is_the_reference_trajectory_without_the_dihedral_term():
    """tests/golden/config1_verlet_f64.npz (oracle/make_golden.py --only config1_verlet): 100 steps of the
    reference's own VerletIntegrator on the example system, dt 0.05 fs from rest, every force-field term but
    the dihedral one (whose reference force is not the gradient of its energy, DESIGN Q12).  The GPU parity
    test on it belongs to the next round; here the fixture is checked for what it claims to be."""
    g = load_golden('config1_verlet_f64')
    c1 = load_golden('config1_f64')
    assert np.array_equal(g['positions0'], c1['positions']) and np.array_equal(g['box'], c1['box'])
    assert int(g['steps']) == 100 and float(g['dt']) == 0.05
    names = [str(x) for x in g['constraints']]
    assert 'CharmmDihedralConstraint' not in names and 'CharmmNonbondedConstraint' in names and 'ElectrostaticConstraint' in names
    steps = [int(k) for k in g['snapshot_steps']]
    assert steps == sorted(steps) and steps[-1] == 100 and g['snapshots'].shape == (len(steps), 2423, 3)
    # from rest, x(t) - x(0) = a t^2 / 2 to leading order (the reference's first step uses a dt^2, Q4): the
    # early snapshots must grow quadratically with the step count
    d1 = np.abs(g['snapshots'][steps.index(5)] - g['positions0']).max()
    d2 = np.abs(g['snapshots'][steps.index(10)] - g['positions0']).max()
    assert 3.0 < d2 / d1 < 4.6
    assert np.isfinite(g['final_velocities']).all() and np.abs(g['snapshots'][-1] - g['positions0']).max() < 2.0


def test_restatement_matches_reference_at_benchmark_size():
    """tests/golden/config2_full_f64.npz (oracle/make_golden.py --only config2_full): the unmodified reference
    on the 23 556-atom water box of the benchmark (config 2) — LJ over its 27-cell list at rc 9 A (5 cells per
    edge: the pair loss Q1 is in the numbers) and bare all-pairs Coulomb, DOUBLE mode.  The box is regenerated
    from its seed; the fixture stores every 16th atom's forces, the energies and checksums."""
    from mdpy_b200 import synthetic
    from mdpy_b200.utils import wrap_positions
    g = load_golden('config2_full_f64')
    s = synthetic.CONFIGS['water_23k']()
    n = s.num_particles
    assert n == int(g['n']) == 23556 and np.allclose(s.box, g['box'])
    pbc = np.diag(s.box)
    pos = wrap_positions(s.positions.astype(np.float64), pbc, np.linalg.inv(pbc))
    assert np.allclose([pos.sum(), (pos ** 2).sum()], g['position_checksum'], rtol=1e-12)
    topo = s.topology()
    rows = {k: (list(v) + list(v) if len(v) == 2 else list(v)) for k, v in s.lj_parameters.items()}
    table = np.array([rows[t] for t in s.types], dtype=np.float64)
    charges = np.asarray(topo.charges, dtype=np.float64)
    stride = int(g['stride'])
    threads = os.cpu_count() or 1
    f, e, _ = ora.lj_cell(pos, table, pbc, 9.0, topo.bonded_particles, topo.scaling_particles, cell_cutoff=12.0, threads=threads)
    assert rel_rms(f[::stride], g['lj_forces_strided']) < 1e-9
    assert (f ** 2).sum() == pytest.approx(float(g['lj_force_sumsq']), rel=1e-9)
    assert e == pytest.approx(float(g['lj_energy']), rel=1e-9)
    k = 4 * np.pi * float(np.float32(0.5727653))
    f, e = ora.coulomb_allpairs(pos, charges, pbc, topo.bonded_particles, k, threads=threads)
    assert rel_rms(f[::stride], g['coul_forces_strided']) < 1e-9
    assert (f ** 2).sum() == pytest.approx(float(g['coul_force_sumsq']), rel=1e-9)
    assert e == pytest.approx(float(g['coul_energy']), rel=1e-8)


def test_langevin_oracle_is_pinned_to_published_philox_vectors_and_gjf_limits():
    """oracle/cpu_oracle.py:philox4x32_10 against the known-answer vectors published with Random123
    (kat_vectors: philox4x32 10 rounds), and the G-JF restatement against its analytic limits: gamma = 0 is
    velocity Verlet; a free particle's velocity variance after many steps is kT/m."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = ora.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(x) for x in got) == want
    z = ora.langevin_noise(7, 100000, 3)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and not np.array_equal(z, ora.langevin_noise(7, 100000, 4))
    assert np.array_equal(z, ora.langevin_noise(7, 100000, 3))          # counter based: reproducible
    # gamma = 0: x' = x + dt v + dt^2 f / 2m, v' = v + dt (f + f') / 2m with a harmonic force
    rng = np.random.default_rng(0)
    x, v, m, k, dt = rng.normal(size=(5, 3)), rng.normal(size=(5, 3)), np.array([1., 2, 3, 4, 5]), 0.3, 0.1
    xn, vn, fn = ora.gjf_step(x, v, -k * x, lambda y: -k * y, m, dt, 0.0, 1.0, 1, 0)
    assert np.allclose(xn, x + dt * v + 0.5 * dt * dt * (-k * x) / m[:, None])
    assert np.allclose(vn, v + 0.5 * dt * (-k * x - k * xn) / m[:, None])
    # free particles in a bath: <v^2> -> kT / m
    n, kT, gamma = 20000, 0.5, 0.5
    x = np.zeros((n, 3)); v = np.zeros((n, 3)); f = np.zeros((n, 3)); mm = np.full(n, 2.0)
    for step in range(60):
        x, v, f = ora.gjf_step(x, v, f, lambda y: np.zeros_like(y), mm, 0.5, gamma, kT, 11, step)
    assert (v ** 2).mean() == pytest.approx(kT / 2.0, rel=0.02)


def test_bonded_restatement_matches_reference_on_config1():
    """oracle/bonded.py against the unmodified reference's bonded constraints on the example system (DOUBLE mode)."""
    from oracle import bonded
    g = load_golden('config1_f64')
    x, box = g['positions'], g['box']
    for name, fn in (('CharmmBondConstraint', bonded.bonds), ('CharmmAngleConstraint', bonded.angles),
                     ('CharmmImproperConstraint', bonded.impropers)):
        f, e = fn(x, box, g[name + '_idx'], g[name + '_par'])
        assert rel_rms(f, g[name + '_forces']) < 1e-10, name
        assert e == pytest.approx(float(g[name + '_energy']), rel=1e-10), name
    e = bonded.dihedral_energy(x, box, g['CharmmDihedralConstraint_idx'], g['CharmmDihedralConstraint_par'])
    assert e == pytest.approx(float(g['CharmmDihedralConstraint_energy']), rel=1e-10)


def test_oracle_reproduces_the_reference_config1_trajectory():
    """100 steps of the reference's VerletIntegrator on the example system (tests/golden/config1_verlet_f64.npz)
    against oracle.verlet + the restated force terms, from the golden's own float64 start: the checker the GPU
    trajectory test uses is the reference's trajectory to 1e-8 A."""
    g = load_golden('config1_verlet_f64')
    c1 = load_golden('config1_f64')
    steps = [int(v) for v in g['snapshot_steps']]
    _, vel, _, _, snaps = ora.verlet(g['positions0'], np.zeros_like(g['positions0']), c1['masses'], np.diag(c1['box']),
                                     float(g['dt']), steps[-1], ora.config1_force_fn(c1, threads=os.cpu_count() or 1),
                                     snapshot_steps=steps)
    for k, step in enumerate(steps):
        assert np.abs(snaps[step] - g['snapshots'][k]).max() < 1e-8, step
    assert np.abs(vel - g['final_velocities']).max() < 1e-8


def test_settle_oracle_is_pinned_against_iterative_shake():
    """oracle/settle.py: the analytic SETTLE positions equal what an independent iterative SHAKE converges to, the
    constraints hold to rounding, the centre of mass does not move, and the velocity stage leaves no velocity along a bond."""
    from oracle import settle as st
    rng = np.random.default_rng(1)
    W, d_oh, ang = 500, 0.9572, np.deg2rad(104.52)
    d_hh = 2 * d_oh * np.sin(ang / 2)
    local = np.array([[0, 0, 0], [d_oh * np.sin(ang / 2), 0, d_oh * np.cos(ang / 2)], [-d_oh * np.sin(ang / 2), 0, d_oh * np.cos(ang / 2)]])
    q = rng.normal(size=(W, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    x_old = np.einsum('nij,kj->nki', R, local) + rng.uniform(-20, 20, size=(W, 1, 3))
    x_new = x_old + rng.normal(size=(W, 3, 3)) * 0.05
    mo, mh = 15.9994, 1.008
    a = st.settle(x_old, x_new, mo, mh, d_oh, d_hh)
    assert np.abs(a - st.shake(x_old, x_new, mo, mh, d_oh, d_hh)).max() < 1e-12
    dist = lambda p, i, j: np.linalg.norm(p[:, i] - p[:, j], axis=1)
    assert np.abs(dist(a, 0, 1) - d_oh).max() < 1e-13 and np.abs(dist(a, 0, 2) - d_oh).max() < 1e-13 and np.abs(dist(a, 1, 2) - d_hh).max() < 1e-13
    m = np.array([mo, mh, mh])
    assert np.abs((m[None, :, None] * (a - x_new)).sum(1)).max() < 1e-12
    v = rng.normal(size=(W, 3, 3)) * 0.01
    v2 = st.rattle_velocities(a, v, mo, mh)
    for i, j in ((0, 1), (0, 2), (1, 2)):
        assert np.abs(((v2[:, j] - v2[:, i]) * (a[:, j] - a[:, i])).sum(1)).max() < 1e-15
    assert np.abs((m[None, :, None] * (v2 - v)).sum(1)).max() < 1e-15
