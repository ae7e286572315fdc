"""oracle/make_golden.py — runs the UNMODIFIED reference (mdpy v0.2.x, /root/reference) in this
container and writes golden input/output vectors to tests/golden/.

TEST INFRASTRUCTURE ONLY.  The reference is Python and cannot travel to the GPU box, so its
results are frozen here as small fixtures (the committed .npz/.json files); this script is the
recipe that made them:

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=oracle/refshim:/root/reference \
        python oracle/make_golden.py [--precision DOUBLE|SINGLE] [--only NAME]

Each fixture stores the inputs exactly as the reference's kernels saw them (wrapped positions,
per-atom LJ table, -1 padded exclusion tables, box) and the reference's outputs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='DOUBLE')
ap.add_argument('--only', default=None)
args = ap.parse_args()

import mdpy as md  # the reference (via refshim)

md.env.set_precision(args.precision)  # must precede every Particle/Topology/constraint
from mdpy.constraint import (CharmmAngleConstraint, CharmmBondConstraint, CharmmDihedralConstraint,  # noqa: E402
                             CharmmImproperConstraint, CharmmNonbondedConstraint, ElectrostaticConstraint)
from mdpy.core import Particle, Topology  # noqa: E402
from mdpy.ensemble import Ensemble  # noqa: E402
from mdpy.forcefield import CharmmForcefield  # noqa: E402
from mdpy.integrator import VerletIntegrator  # noqa: E402
from mdpy.io import CharmmTopparParser  # noqa: E402
from mdpy.unit import *  # noqa: E402,F401,F403
from mdpy.utils import wrap_positions, unwrap_vec  # noqa: E402

from mdpy_b200 import synthetic  # noqa: E402  (only the seeded generators: pure numpy)

SUFFIX = '_f64' if args.precision.upper() == 'DOUBLE' else '_f32'
F = md.env.NUMPY_FLOAT


def want(name):
    return args.only is None or args.only == name


def save(name, **arrays):
    path = os.path.join(GOLDEN, name + SUFFIX + '.npz')
    np.savez_compressed(path, **arrays)
    print('wrote %s (%.1f kB)' % (path, os.path.getsize(path) / 1e3))


def ref_topology(types, masses, charges, bonds, angles, dihedrals, impropers=()):
    t = Topology()
    t.add_particles([Particle(particle_id=i, particle_type=tp, particle_name=tp, molecule_type='SYN',
                              mass=float(m), charge=float(q))
                     for i, (tp, m, q) in enumerate(zip(types, masses, charges))])
    for b in bonds: t.add_bond([int(x) for x in b])
    for a in angles: t.add_angle([int(x) for x in a])
    for d in dihedrals: t.add_dihedral([int(x) for x in d])
    for d in impropers: t.add_improper([int(x) for x in d])
    t.join()
    return t


def eval_constraint(c):
    t0 = time.time()
    c.update()
    return np.array(c.forces), float(c.potential_energy), time.time() - t0


# ---------------------------------------------------------------------------------------------
# constants the hot path depends on
if want('constants') and SUFFIX == '_f64':
    from mdpy.io.charmm_toppar_parser import RMIN_TO_SIGMA_FACTOR
    consts = dict(
        EPSILON0=float(EPSILON0.value), EPSILON0_repr=repr(EPSILON0.value),
        KB=float(KB.value), NA=float(NA.value),
        kcal_permol=float(Quantity(1, kilocalorie_permol).convert_to(default_energy_unit).value),
        kj_permol=float(Quantity(1, kilojoule_permol).convert_to(default_energy_unit).value),
        RMIN_TO_SIGMA_FACTOR=float(RMIN_TO_SIGMA_FACTOR),
        kbt_300=float((Quantity(300, kelvin) * KB).convert_to(default_energy_unit).value),
        nm_091_in_A=float(Quantity(0.91, nanometer).convert_to(default_length_unit).value),
    )
    with open(os.path.join(GOLDEN, 'reference_constants.json'), 'w') as f:
        json.dump(consts, f, indent=1)
    print(consts)

# ---------------------------------------------------------------------------------------------
# known-answer tests of the reference's own suite, with the reference's outputs
if want('kat'):
    data = os.path.join(REF, 'mdpy', 'test', 'data')
    prm = CharmmTopparParser(os.path.join(data, 'toppar_water_ions_namd.str'), os.path.join(data, 'par_all36_prot.prm'),
                             os.path.join(data, 'top_all36_na.rtf')).parameters
    # test_charmm_nonbonded_constraint.py:26-67,93-126
    types = ['CA', 'NY', 'CPT', 'CA']
    t = ref_topology(types, [12, 14, 1, 12], [0, 0, 0, 0], [], [], [])
    pos = np.array([[0, 0, 0], [0, 10, 0], [0, 21, 0], [0, 11, 0]], dtype=F)
    ens = Ensemble(t, np.eye(3) * 30)
    ens.state.cell_list.set_cutoff_radius(5)
    ens.state.set_positions(pos)
    lj = CharmmNonbondedConstraint(prm['nonbonded'])
    lj.set_cutoff_radius(Quantity(0.91, nanometer))
    ens.add_constraints(lj)
    ens.state.cell_list.update(ens.state.positions)
    f_lj, e_lj, _ = eval_constraint(lj)
    lj_params = {k: [float(x) for x in prm['nonbonded'][k]] for k in ('CA', 'NY', 'CPT')}
    # test_electrostatic_constraint.py:26-95
    t2 = ref_topology(['C', 'N', 'CA', 'C'], [12, 14, 1, 12], [1, 2, 0, 0], [], [], [])
    ens2 = Ensemble(t2, np.eye(3) * 30)
    ens2.state.cell_list.set_cutoff_radius(12)
    ens2.state.set_positions(pos)
    ens2.state.set_pbc_matrix(np.diag(np.ones(3) * 100))
    el = ElectrostaticConstraint()
    ens2.add_constraints(el)
    f_el, e_el, _ = eval_constraint(el)
    save('kat', positions=pos, lj_forces=f_lj, lj_energy=e_lj, lj_cutoff=float(lj.cutoff_radius),
         lj_table=np.array(lj._parameters_list), lj_param_json=json.dumps(lj_params),
         coul_positions=np.array(ens2.state.positions), coul_forces=f_el, coul_energy=e_el,
         coul_charges=np.array(t2.charges))

# ---------------------------------------------------------------------------------------------
def run_synthetic(name, sys_, rc, do_verlet=0, dt=0.5):
    t = ref_topology(sys_.types, sys_.masses, sys_.charges, sys_.bonds, sys_.angles, sys_.dihedrals, sys_.impropers)
    ens = Ensemble(t, np.diag(sys_.box))
    lj = CharmmNonbondedConstraint(sys_.lj_parameters, cutoff_radius=rc)
    el = ElectrostaticConstraint()
    ens.add_constraints(lj, el)
    ens.state.set_positions(sys_.positions.astype(F))
    f_lj, e_lj, t_lj = eval_constraint(lj)
    f_el, e_el, t_el = eval_constraint(el)
    print('%s: N=%d LJ %.6f (%.1fs) Coulomb %.6f (%.1fs)' % (name, t.num_particles, e_lj, t_lj, e_el, t_el))
    cl = ens.state.cell_list
    out = dict(box=sys_.box, positions=np.array(ens.state.positions), charges=np.array(t.charges),
               masses=np.array(t.masses), lj_table=np.array(lj._parameters_list), rc=rc,
               bonded=np.array(t.bonded_particles), scaling=np.array(t.scaling_particles),
               lj_forces=f_lj, lj_energy=e_lj, coul_forces=f_el, coul_energy=e_el,
               cell_num=np.array(cl.num_cell_vec), cell_index=np.array(cl.particle_cell_index),
               cell_list_shape=np.array(cl.cell_list.shape), cell_list_head=np.array(cl.cell_list[0, 0, 0]),
               ref_seconds=np.array([t_lj, t_el]))
    if do_verlet:
        rng = np.random.default_rng(7)
        v0 = (rng.normal(size=sys_.positions.shape) * 0.002).astype(F)
        ens.state.set_velocities(v0)
        integ = VerletIntegrator(dt)
        t0 = time.time()
        integ.integrate(ens, do_verlet)
        print('  verlet %d steps %.1fs' % (do_verlet, time.time() - t0))
        out.update(verlet_v0=v0, verlet_dt=dt, verlet_steps=do_verlet,
                   verlet_positions=np.array(ens.state.positions), verlet_velocities=np.array(ens.state.velocities),
                   verlet_cur=np.array(integ.cur_positions), verlet_pre=np.array(integ.pre_positions))
    save(name, **out)


if want('mix_small'):
    # 1 helix (100 atoms) + 867 waters in a 30 A box, rc 9: 3x3x3 cells, so the reference's cell list
    # cannot drop pairs and its result is the plain minimum-image sum
    s = synthetic.solvated_protein_box(2701, (30.0, 30.0, 30.0), protein_fraction=0.037, seed=11, n_res=10)
    run_synthetic('mix_small', s, 9.0)

if want('verlet_small'):
    # the reference's State insists on a box >= 2 x 12 A (state.py:28 -> cell_list.py:20,64-69)
    s = synthetic.solvated_protein_box(1471, (24.5, 24.5, 24.5), protein_fraction=0.068, seed=12, n_res=10)
    run_synthetic('verlet_small', s, 9.0, do_verlet=5, dt=0.5)

if want('q1_case'):
    # >= 4 cells per dimension -> the reference misses pairs across the periodic boundary
    rng = np.random.default_rng(5)
    n, L, rc = 1500, 50.0, 12.0
    m = int(np.ceil(n ** (1 / 3)))
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing='ij'), -1).reshape(-1, 3)[:n]
    pos = ((g + 0.5) * (L / m) - L / 2 + rng.uniform(-1.2, 1.2, size=(n, 3))).astype(F)
    t = ref_topology(['CT1'] * n, [12.0] * n, [0.0] * n, [], [], [])
    ens = Ensemble(t, np.eye(3) * L)
    lj = CharmmNonbondedConstraint({'CT1': synthetic.LJ_PARAMETERS['CT1']}, cutoff_radius=rc)
    ens.add_constraints(lj)
    ens.state.set_positions(pos)
    f, e, tt = eval_constraint(lj)
    print('q1_case: LJ %.8f (%.1fs), cells %s' % (e, tt, ens.state.cell_list.num_cell_vec))
    save('q1_case', box=np.full(3, L), positions=np.array(ens.state.positions), lj_table=np.array(lj._parameters_list),
         rc=rc, lj_forces=f, lj_energy=e, cell_num=np.array(ens.state.cell_list.num_cell_vec))

# ---------------------------------------------------------------------------------------------
# config 1: example/charmm system (SURVEY §8c "Config-1 loading without MDAnalysis", Appendix A5)
def parse_psf(path):
    lines = open(path).read().split('\n')
    sec = {}
    i = 0
    while i < len(lines):
        ln = lines[i]
        if '!N' in ln:
            key = ln.split('!')[1].split(':')[0].split()[0]
            cnt = int(ln.split()[0])
            sec[key] = (i + 1, cnt)
        i += 1
    s, n = sec['NATOM']
    atoms = [lines[s + k].split() for k in range(n)]

    def block(key, width):
        s, cnt = sec[key]
        vals = []
        k = s
        while len(vals) < cnt * width:
            vals += [int(x) for x in lines[k].split()]
            k += 1
        return np.array(vals[:cnt * width]).reshape(cnt, width) - 1
    return atoms, block('NBOND', 2), block('NTHETA', 3), block('NPHI', 4), block('NIMPHI', 4)


if want('config1'):
    base = os.path.join(REF, 'example', 'charmm', 'str')
    atoms, bonds, angles, dihedrals, impropers = parse_psf(os.path.join(base, '6PO6_ionized.psf'))
    pdb = open(os.path.join(base, '6PO6_ionized.pdb')).read().split('\n')
    cryst = [l for l in pdb if l.startswith('CRYST1')][0]
    box = np.array([float(cryst[6:15]), float(cryst[15:24]), float(cryst[24:33])])
    xyz = np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in pdb if l.startswith(('ATOM', 'HETATM'))])
    assert len(xyz) == len(atoms)
    t = Topology()
    t.add_particles([Particle(particle_id=i, particle_type=a[5], particle_name=a[4], molecule_id=int(a[2]),
                              molecule_type=a[3], chain_id=a[1], mass=float(a[7]), charge=float(a[6]))
                     for i, a in enumerate(atoms)])
    for b in bonds: t.add_bond([int(x) for x in b])
    for a in angles: t.add_angle([int(x) for x in a])
    for d in dihedrals: t.add_dihedral([int(x) for x in d])
    for d in impropers: t.add_improper([int(x) for x in d])
    t.join()
    ff = CharmmForcefield(t, np.diag(box))
    ff.set_param_files(os.path.join(REF, 'data', 'charmm', 'par_all36_prot.prm'),
                       os.path.join(REF, 'data', 'charmm', 'toppar_water_ions_namd.str'))
    ens = ff.create_ensemble()
    ens.state.set_positions(xyz.astype(F))
    out = dict(box=box, positions=np.array(ens.state.positions), charges=np.array(t.charges), masses=np.array(t.masses),
               bonded=np.array(t.bonded_particles), scaling=np.array(t.scaling_particles),
               types=np.array([p.particle_type for p in t.particles]))
    kcal = Quantity(1, kilocalorie_permol).convert_to(default_energy_unit).value
    for c in ens.constraints:
        name = type(c).__name__
        f, e, tt = eval_constraint(c)
        print('config1 %-28s %14.6f kcal/mol (%.1fs)' % (name, e / kcal, tt))
        out[name + '_forces'] = f
        out[name + '_energy'] = e
        if hasattr(c, '_parameters_list'):
            out['lj_table'] = np.array(c._parameters_list)
            out['rc'] = float(c.cutoff_radius)
        if hasattr(c, '_int_parameters') and len(np.shape(c._int_parameters)) == 2:
            out[name + '_idx'] = np.array(c._int_parameters)
            out[name + '_par'] = np.array(c._float_parameters)
    save('config1', **out)

# ---------------------------------------------------------------------------------------------
# config 1, trajectory: the example system advanced by the reference's own VerletIntegrator
# (example/charmm/charmm.py:45: dt = 0.05 fs; zero start velocities) with every constraint of the
# force field EXCEPT the dihedral one, whose force is not the gradient of its energy (DESIGN Q12), so a
# drop-in with the analytic gradient cannot follow it.  ~10 s per step on one core: run with
#   --only config1_verlet [--verlet-steps 100]
if want('config1_verlet') and args.only == 'config1_verlet':
    from mdpy.constraint import CharmmDihedralConstraint as _Dih
    g = np.load(os.path.join(GOLDEN, 'config1' + SUFFIX + '.npz'))
    base = os.path.join(REF, 'example', 'charmm', 'str')
    atoms, bonds, angles, dihedrals, impropers = parse_psf(os.path.join(base, '6PO6_ionized.psf'))
    t = Topology()
    t.add_particles([Particle(particle_id=i, particle_type=a[5], particle_name=a[4], molecule_id=int(a[2]),
                              molecule_type=a[3], chain_id=a[1], mass=float(a[7]), charge=float(a[6]))
                     for i, a in enumerate(atoms)])
    for b in bonds: t.add_bond([int(x) for x in b])
    for a in angles: t.add_angle([int(x) for x in a])
    for d in dihedrals: t.add_dihedral([int(x) for x in d])
    for d in impropers: t.add_improper([int(x) for x in d])
    t.join()
    ff = CharmmForcefield(t, np.diag(g['box']))
    ff.set_param_files(os.path.join(REF, 'data', 'charmm', 'par_all36_prot.prm'),
                       os.path.join(REF, 'data', 'charmm', 'toppar_water_ions_namd.str'))
    full = ff.create_ensemble()
    ens = Ensemble(t, np.diag(g['box']))
    kept = [c for c in full.constraints if not isinstance(c, _Dih)]
    for c in kept:
        c._parent_ensemble = None
    ens.add_constraints(*kept)
    ens.state.set_positions(np.array(g['positions'], dtype=F))
    nsteps = int(os.environ.get('VERLET_STEPS', '100'))
    dt = 0.05
    integ = VerletIntegrator(dt)
    t0 = time.time()
    snaps = {}
    for s in range(nsteps):
        integ.integrate(ens, 1)
        if (s + 1) in (1, 5, 10, 25, 50, 100, nsteps):
            snaps[s + 1] = np.array(integ.cur_positions)
            print('  config1 verlet step %d (%.0fs)' % (s + 1, time.time() - t0), flush=True)
    save('config1_verlet', box=g['box'], positions0=np.array(g['positions']), dt=dt, steps=nsteps,
         constraints=np.array([type(c).__name__ for c in kept]),
         snapshot_steps=np.array(sorted(snaps)), snapshots=np.stack([snaps[k] for k in sorted(snaps)]),
         final_positions=np.array(ens.state.positions), final_velocities=np.array(ens.state.velocities))

# ---------------------------------------------------------------------------------------------
# config 2 at full size: the reference's own LJ (27-cell list, rc 9 A — 5 cells per edge, so its pair
# loss Q1 is in the numbers) and bare all-pairs Coulomb on the 23 556-atom water box of the benchmark.
# The box is regenerated from its seed (mdpy_b200.synthetic.water_box(7852, 20260001)), so only a
# strided subset of the per-atom forces, the energies and a position checksum are stored.  ~6 min on one
# core:  --only config2_full
if want('config2_full') and args.only == 'config2_full':
    s = synthetic.CONFIGS['water_23k']()
    t = ref_topology(s.types, s.masses, s.charges, s.bonds, s.angles, s.dihedrals, s.impropers)
    ens = Ensemble(t, np.diag(s.box))
    lj = CharmmNonbondedConstraint(s.lj_parameters, cutoff_radius=9.0)
    el = ElectrostaticConstraint()
    ens.add_constraints(lj, el)
    ens.state.set_positions(s.positions.astype(F))
    f_lj, e_lj, t_lj = eval_constraint(lj)
    print('config2_full LJ %.9f (%.1fs)' % (e_lj, t_lj), flush=True)
    f_el, e_el, t_el = eval_constraint(el)
    print('config2_full Coulomb %.9f (%.1fs)' % (e_el, t_el), flush=True)
    stride = 16
    cl = ens.state.cell_list
    save('config2_full', box=s.box, n=t.num_particles, rc=9.0, stride=stride,
         position_checksum=np.array([np.asarray(ens.state.positions, dtype=np.float64).sum(),
                                     (np.asarray(ens.state.positions, dtype=np.float64) ** 2).sum()]),
         lj_forces_strided=f_lj[::stride], lj_energy=e_lj, lj_force_sumsq=float((f_lj.astype(np.float64) ** 2).sum()),
         coul_forces_strided=f_el[::stride], coul_energy=e_el, coul_force_sumsq=float((f_el.astype(np.float64) ** 2).sum()),
         cell_num=np.array(cl.num_cell_vec), ref_seconds=np.array([t_lj, t_el]))
