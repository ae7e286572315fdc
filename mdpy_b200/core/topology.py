"""Topology — particle list plus bonds/angles/dihedrals/impropers, and `join()` which turns
them into the dense arrays the kernels consume (mdpy/core/topology.py:59-80):
masses, charges float[N,1]; bonded_particles (1-2 u 1-3) and scaling_particles (1-4) as
-1-padded int[N, maxBonded] (both use the *bonded* width, topology.py:74-76).

`Topology.from_arrays` is the bulk path for synthetic systems with 10^5-10^7 atoms, where
one Python object per atom is not an option; it produces the same joined arrays.
"""
import numpy as np

from ..environment import env
from ..error import GeomtryDimError, ModifyJoinedTopologyError, ParticleConflictError
from .particle import Particle


def _padded_table(rows_i, rows_j, n, width=None, dtype=np.int32):
    """rows: pairs (i -> j) in insertion order -> -1 padded [n, width] table."""
    rows_i = np.asarray(rows_i, dtype=np.int64); rows_j = np.asarray(rows_j, dtype=np.int64)
    counts = np.bincount(rows_i, minlength=n) if rows_i.size else np.zeros(n, dtype=np.int64)
    w = int(counts.max()) if counts.size and counts.max() > 0 else 0
    width = w if width is None else max(width, w)
    table = -np.ones((n, width), dtype=dtype)
    if rows_i.size:
        order = np.argsort(rows_i, kind='stable')
        ri, rj = rows_i[order], rows_j[order]
        start = np.concatenate([[0], np.cumsum(counts)[:-1]])
        col = np.arange(ri.size) - start[ri]
        table[ri, col] = rj
    return table


class Topology:
    def __init__(self):
        self._particles = []
        self._bonds, self._angles, self._dihedrals, self._impropers = [], [], [], []
        self._is_joined = False
        self._masses = self._charges = self._bonded_particles = self._scaling_particles = []
        self._bulk = None

    def __repr__(self):
        return '<mdpy_b200.core.Toplogy object: %d particles at %x>' % (self.num_particles, id(self))

    # ---- bulk construction -------------------------------------------------------------
    @classmethod
    def from_arrays(cls, particle_types, masses, charges, bonds=(), angles=(), dihedrals=(), impropers=()):
        """particle_types: sequence of N type names; bonds [nb,2], angles [na,3], dihedrals [nd,4]
        int arrays of matrix ids.  Applies the same partner rules as add_bond / add_angle /
        add_dihedral (topology.py:122-134,155-167,188-200) and joins immediately."""
        t = cls()
        n = len(particle_types)
        as_idx = lambda a, k: np.asarray(a, dtype=np.int64).reshape(-1, k)
        bonds, angles, dihedrals, impropers = as_idx(bonds, 2), as_idx(angles, 3), as_idx(dihedrals, 4), as_idx(impropers, 4)
        # insertion order of the reference: bond by bond (both directions), then angle by angle
        both = lambda a, b: np.stack([a, b], 1).reshape(-1)
        bi = np.concatenate([both(bonds[:, 0], bonds[:, 1]), both(angles[:, 0], angles[:, 2])])
        bj = np.concatenate([both(bonds[:, 1], bonds[:, 0]), both(angles[:, 2], angles[:, 0])])
        key = bi * n + bj
        if np.unique(key).size != key.size or (bi == bj).any():
            raise ParticleConflictError('a particle pair appears twice in the bonded (1-2 / 1-3) lists')
        si = both(dihedrals[:, 0], dihedrals[:, 3]); sj = both(dihedrals[:, 3], dihedrals[:, 0])
        if si.size:  # repeats are dropped, first occurrence kept
            _, first = np.unique(si * n + sj, return_index=True)
            first.sort()
            si, sj = si[first], sj[first]
        t._bulk = dict(types=list(particle_types), n=n)
        t._bonds, t._angles, t._dihedrals, t._impropers = bonds, angles, dihedrals, impropers
        t._masses = np.asarray(masses, dtype=env.NUMPY_FLOAT).reshape(n, 1)
        t._charges = np.asarray(charges, dtype=env.NUMPY_FLOAT).reshape(n, 1)
        t._bonded_particles = _padded_table(bi, bj, n, dtype=env.NUMPY_INT)
        t._scaling_particles = _padded_table(si, sj, n, width=t._bonded_particles.shape[1], dtype=env.NUMPY_INT)
        t._is_joined = True
        return t

    @classmethod
    def from_tables(cls, particle_types, masses, charges, bonded_particles, scaling_particles,
                    bonds=(), angles=(), dihedrals=(), impropers=()):
        """A joined Topology straight from the dense arrays a reference Topology.join() produced
        (golden fixtures): no partner rules are re-derived."""
        t = cls()
        n = len(particle_types)
        as_idx = lambda a, k: np.asarray(a, dtype=np.int64).reshape(-1, k)
        t._bulk = dict(types=list(particle_types), n=n)
        t._bonds, t._angles, t._dihedrals, t._impropers = as_idx(bonds, 2), as_idx(angles, 3), as_idx(dihedrals, 4), as_idx(impropers, 4)
        t._masses = np.asarray(masses, dtype=env.NUMPY_FLOAT).reshape(n, 1)
        t._charges = np.asarray(charges, dtype=env.NUMPY_FLOAT).reshape(n, 1)
        t._bonded_particles = np.ascontiguousarray(bonded_particles, dtype=env.NUMPY_INT).reshape(n, -1)
        t._scaling_particles = np.ascontiguousarray(scaling_particles, dtype=env.NUMPY_INT).reshape(n, -1)
        t._is_joined = True
        return t

    @property
    def particle_types(self):
        if self._bulk is not None:
            return self._bulk['types']
        return [p.particle_type for p in self._particles]

    # ---- incremental construction (reference API) ------------------------------------------
    def _editable(self):
        if self._is_joined:
            raise ModifyJoinedTopologyError('%s has been joined. No change can be made.' % self)

    def _check_ids(self, *ids):
        for k, i in enumerate(ids):
            if i >= self.num_particles:
                raise ParticleConflictError('Matrix id %d beyonds the range of particles contain in toplogy' % i)
            if i in ids[k + 1:]:
                raise ParticleConflictError('Particle appears twice in a topology connection')

    def add_particles(self, particles):
        self._editable()
        for p in particles:
            if not isinstance(p, Particle):
                raise TypeError('mdpy_b200.core.Particle type is excepted, while %s provided' % type(p))
            p.change_matrix_id(len(self._particles))
            self._particles.append(p)

    def _connection(self, ids, size, name):
        self._editable()
        if len(ids) != size:
            raise GeomtryDimError('%s should be a matrix id list of %d Particles, instead of %d' % (name, size, len(ids)))
        self._check_ids(*ids)

    def add_bond(self, bond):
        self._connection(bond, 2, 'Bond')
        a, b = bond
        self._bonds.append(list(bond))
        self._particles[a].add_bonded_particle(b); self._particles[b].add_bonded_particle(a)

    def add_angle(self, angle):
        self._connection(angle, 3, 'Angle')
        a, _, c = angle
        self._angles.append(list(angle))
        self._particles[a].add_bonded_particle(c); self._particles[c].add_bonded_particle(a)

    def add_dihedral(self, dihedral, scaling_factor=1):
        self._connection(dihedral, 4, 'Dihedral')
        a, d = dihedral[0], dihedral[3]
        self._dihedrals.append(list(dihedral))
        self._particles[a].add_scaling_particle(d, scaling_factor); self._particles[d].add_scaling_particle(a, scaling_factor)

    def add_improper(self, improper):
        self._connection(improper, 4, 'Improper')
        self._impropers.append(list(improper))

    def join(self):
        if self._bulk is not None:
            return
        n = self.num_particles
        self._masses = np.array([[p.mass] for p in self._particles], dtype=env.NUMPY_FLOAT).reshape(n, 1)
        self._charges = np.array([[p.charge] for p in self._particles], dtype=env.NUMPY_FLOAT).reshape(n, 1)
        width = max([p.num_bonded_particles for p in self._particles], default=0)
        self._bonded_particles = -np.ones((n, width), dtype=env.NUMPY_INT)
        self._scaling_particles = -np.ones((n, width), dtype=env.NUMPY_INT)
        for i, p in enumerate(self._particles):
            self._bonded_particles[i, :p.num_bonded_particles] = p.bonded_particles
            self._scaling_particles[i, :p.num_scaling_particles] = p.scaling_particles
        self._is_joined = True

    def split(self):
        if self._bulk is None:
            self._masses = self._charges = self._bonded_particles = self._scaling_particles = []
            self._is_joined = False

    masses = property(lambda self: self._masses)
    charges = property(lambda self: self._charges)
    bonded_particles = property(lambda self: self._bonded_particles)
    scaling_particles = property(lambda self: self._scaling_particles)
    particles = property(lambda self: self._particles)
    num_particles = property(lambda self: self._bulk['n'] if self._bulk is not None else len(self._particles))
    bonds = property(lambda self: self._bonds)
    angles = property(lambda self: self._angles)
    dihedrals = property(lambda self: self._dihedrals)
    impropers = property(lambda self: self._impropers)
    num_bonds = property(lambda self: len(self._bonds))
    num_angles = property(lambda self: len(self._angles))
    num_dihedrals = property(lambda self: len(self._dihedrals))
    num_impropers = property(lambda self: len(self._impropers))
    is_joined = property(lambda self: self._is_joined)
