import sys, numpy as np
sys.path.insert(0, '.')
import mdpy_b200 as md
from mdpy_b200.core import Particle, Topology
from mdpy_b200.constraint import *
ps = [Particle(particle_id=i, particle_type=t, particle_name=n, molecule_type='ASN', mass=m, charge=q)
      for i, (t, n, m, q) in enumerate([('C', 'CA', 12, 1), ('N', 'NY', 14, 2), ('CA', 'CPT', 1, 0), ('C', 'CA', 12, 0)])]
t = Topology(); t.add_particles(ps)
p = np.array([[0, 0, 0], [0, 10, 0], [0, 21, 0], [0, 11, 0]], dtype=np.float64)
ens = md.Ensemble(t, np.eye(3) * 30)
ens.state.set_positions(p)
ens.state.set_pbc_matrix(np.diag(np.ones(3) * 100))
c = ElectrostaticConstraint()
ens.add_constraints(c)
c._configure()
for k in range(3):
    e = c._ctx.compute(c.terms)
    print(k, e[:8], c._ctx.dev.forces()[:2], c._ctx.dev.timing())
c.update(); print(c.potential_energy, c.forces)
