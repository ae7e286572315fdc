import sys, numpy as np
sys.path.insert(0, '.')
import mdpy_b200 as md
from mdpy_b200 import synthetic, _native
from mdpy_b200.constraint import *
from mdpy_b200.unit import coulomb_constant
from oracle import cpu_oracle as ora, spme
np.set_printoptions(precision=10)
s = synthetic.solvated_protein_box(1471, (24.5, 24.5, 24.5), protein_fraction=0.068, seed=12, n_res=10)
ens = md.Ensemble(s.topology(), np.diag(s.box))
pme = ElectrostaticPMEConstraint(cutoff_radius=12.0, alpha=0.30, grid=(32, 32, 32), order=4)
el = ElectrostaticConstraint()
ens.add_constraints(pme)
ens.state.set_positions(s.positions)
pme._configure()
e = pme._ctx.compute(pme.terms)
print('gpu energies', e[:6])
topo = ens.topology
f, en = spme.pme_total(ens.state.positions, s.charges, s.box, topo.bonded_particles, (32,32,32), 4, 0.30, 12.0, coulomb_constant())
print('oracle', en)
ens2 = md.Ensemble(s.topology(), np.diag(s.box)); ens2.add_constraints(el); ens2.state.set_positions(s.positions)
el._configure(); e2 = el._ctx.compute(el.terms); print('bare', e2[:8])
t = ora.nonbonded_bruteforce(ens.state.positions, s.box, s.lj_table(), s.charges, topo.bonded_particles, topo.scaling_particles, rc_lj=0.0, coul_mode=2, k_e=coulomb_constant())
print('bare oracle', t['e_coul'])
