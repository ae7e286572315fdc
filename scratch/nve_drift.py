#!/usr/bin/env python
"""NVE energy drift of the device Verlet integrator on the 23 556-atom water box (SURVEY 8d gate):
10^4 steps at dt = 0.5 fs after a damped relaxation, total energy from the integrator's own kinetic
energy (textbook velocities, Q4), reported as max |E(t) - E(0)| / <KE> and as the least-squares slope in
kT per atom per ns.  Needs a B200:  python scratch/nve_drift.py [--steps 10000] [--dt 0.5]
(not run in round 1; tests/test_gpu_parity.py::test_verlet_nve_energy_is_bounded is the 10^3-step gate)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdpy_b200 import synthetic
from mdpy_b200.integrator import LangevinIntegrator, VerletIntegrator
from mdpy_b200.unit import KB, Quantity, default_energy_unit, kelvin

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=10000)
ap.add_argument('--dt', type=float, default=0.5)
ap.add_argument('--block', type=int, default=100)
args = ap.parse_args()

s = synthetic.water_box(7852, 20260001)
ens = s.ensemble(cutoff=9.0, pme=True, grid=(64, 64, 64))
for rdt, gamma, steps in ((0.1, 0.2, 200), (0.25, 0.05, 400), (0.5, 0.01, 2000)):
    LangevinIntegrator(rdt, 300, gamma, seed=3).integrate(ens, steps)
integ = VerletIntegrator(args.dt, reference_quirks=False)
e_tot, ke = [], []
for _ in range(args.steps // args.block):
    integ.integrate(ens, args.block)
    e_tot.append(ens.total_energy); ke.append(ens.kinetic_energy)
e_tot, ke = np.array(e_tot), np.array(ke)
t_ns = np.arange(1, len(e_tot) + 1) * args.block * args.dt * 1e-6
kT = float((Quantity(300, kelvin) * KB).convert_to(default_energy_unit).value)
slope = np.polyfit(t_ns, e_tot, 1)[0] / kT / s.num_particles
print(json.dumps(dict(steps=args.steps, dt_fs=args.dt, atoms=s.num_particles,
                      max_abs_dE_over_mean_KE=float(np.abs(e_tot - e_tot[0]).max() / ke.mean()),
                      drift_kT_per_atom_per_ns=float(slope), mean_T_K=float(2 * ke.mean() / (3 * s.num_particles) / kT * 300),
                      finite=bool(np.isfinite(e_tot).all()))))
