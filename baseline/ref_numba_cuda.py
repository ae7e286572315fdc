#!/usr/bin/env python
"""baseline/ref_numba_cuda.py — times the UNMODIFIED reference (mdpy v0.2.x, installed into the
git-ignored baseline/_ref by `pip install --no-deps --target baseline/_ref /root/reference`) on its own
numba.cuda path on one B200: the "reference numba.cuda path" baseline of BASELINE.md §3.

    python baseline/ref_numba_cuda.py [--config water_23k|protein_92k] [--platform CUDA|CPU] [--atoms N]

One MD step of the reference = State.set_positions (host cell-list rebuild, state.py:56-61) +
CharmmNonbondedConstraint.update() + ElectrostaticConstraint.update() (11 + 9 cuda.to_device copies,
one kernel and two copy_to_host each, charmm_nonbonded_constraint.py:196-226,
electrostatic_constraint.py:148-174).  The reference does LESS physics than the product path: plain
cutoff LJ and bare 27-cell Coulomb, no PME, no bonded terms timed here.  Prints one JSON line.
MEASUREMENT INFRASTRUCTURE ONLY: nothing in mdpy_b200/ imports this or baseline/_ref.
"""
import argparse
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument('--config', default='protein_92k')
ap.add_argument('--platform', default='CUDA')
ap.add_argument('--atoms', type=int, default=0, help='generate a smaller box of the same kind (testing)')
ap.add_argument('--evals', type=int, default=3)
args = ap.parse_args()

# import shims of SURVEY 8c (stubs for MDAnalysis / h5py / matplotlib, np.round_)
spec = importlib.util.spec_from_file_location('refshim_sitecustomize', os.path.join(ROOT, 'oracle', 'refshim', 'sitecustomize.py'))
spec.loader.exec_module(importlib.util.module_from_spec(spec))
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
try:
    import mdpy as md
    from mdpy.constraint import CharmmNonbondedConstraint, ElectrostaticConstraint
    from mdpy.core import Particle, Topology
    from mdpy.ensemble import Ensemble
except Exception as e:  # noqa: BLE001
    print(json.dumps(dict(impl='reference_numba_cuda', unavailable='import failed: %r' % (e,))))
    sys.exit(0)

from mdpy_b200 import synthetic   # seeded generators only (pure numpy)

CUTOFF = dict(water_23k=9.0, protein_92k=12.0, protein_1m=12.0)[args.config]
t0 = time.perf_counter()
if args.atoms:
    s = synthetic.water_box(args.atoms // 3, 7, box=np.full(3, (args.atoms / 0.1002) ** (1 / 3)))
else:
    s = synthetic.CONFIGS[args.config]()
n = s.num_particles
try:
    md.env.set_platform(args.platform)
    topo = Topology()
    topo.add_particles([Particle(particle_id=i, particle_type=t, particle_name=t, molecule_type='SYN', mass=float(m), charge=float(q))
                        for i, (t, m, q) in enumerate(zip(s.types, s.masses, s.charges))])
    for b in s.bonds: topo.add_bond([int(x) for x in b])
    for a in s.angles: topo.add_angle([int(x) for x in a])
    for d in s.dihedrals: topo.add_dihedral([int(x) for x in d])
    topo.join()
    ens = Ensemble(topo, np.diag(s.box))
    lj = CharmmNonbondedConstraint(s.lj_parameters, cutoff_radius=CUTOFF)
    coul = ElectrostaticConstraint()
    ens.add_constraints(lj, coul)
    t_setup = time.perf_counter() - t0
    pos = s.positions.astype(md.env.NUMPY_FLOAT)
    t_cell, t_lj, t_c = [], [], []
    for it in range(args.evals + 1):      # first pass = JIT compile + warm-up, not timed
        a = time.perf_counter(); ens.state.set_positions(pos)
        b = time.perf_counter(); lj.update()
        c = time.perf_counter(); coul.update()
        d = time.perf_counter()
        if it:
            t_cell.append(b - a); t_lj.append(c - b); t_c.append(d - c)
    step = float(np.mean(t_cell) + np.mean(t_lj) + np.mean(t_c))
    f = np.asarray(lj.forces, dtype=np.float64) + np.asarray(coul.forces, dtype=np.float64)
    print(json.dumps(dict(
        impl='reference_numba_cuda' if args.platform == 'CUDA' else 'reference_numba_cpu', workload=args.config, atoms=n,
        cutoff_A=CUTOFF, seconds_per_step=step, atom_steps_per_s=n / step, ns_per_day_at_2fs=86400 * 2.0 * 1e-6 / step,
        cell_list_s=float(np.mean(t_cell)), lj_update_s=float(np.mean(t_lj)), coulomb_update_s=float(np.mean(t_c)),
        setup_s=t_setup, evals=args.evals, lj_energy=float(lj.potential_energy), coulomb_energy=float(coul.potential_energy),
        force_rms=float(np.sqrt((f ** 2).mean())),
        note='reference physics: plain-cutoff LJ + bare 27-cell Coulomb (no PME, no bonded terms, no integrator); '
             'per step = State.set_positions (host cell list) + the two Constraint.update() calls incl. their H2D/D2H')))
except Exception as e:  # noqa: BLE001
    import traceback
    traceback.print_exc()
    print(json.dumps(dict(impl='reference_numba_cuda', workload=args.config, atoms=n, unavailable='%s: %s' % (type(e).__name__, str(e)[:300]))))
