"""Particle record — the subset of mdpy/core/particle.py:14-143 the hot path reads:
type (LJ table lookup, charmm_nonbonded_constraint.py:52-53), mass, charge, and the 1-2/1-3
("bonded") and 1-4 ("scaling") partner lists that become the exclusion tables."""
from ..error import ParticleConflictError
from ..unit import check_quantity_value, default_charge_unit, default_mass_unit


class Particle:
    def __init__(self, particle_id=None, particle_type=None, particle_name=None, matrix_id=None,
                 molecule_id=None, molecule_type=None, chain_id=None, mass=None, charge=None):
        self.particle_id, self.particle_type, self.particle_name = particle_id, particle_type, particle_name
        self.matrix_id = matrix_id
        self.molecule_id, self.molecule_type, self.chain_id = molecule_id, molecule_type, chain_id
        self.mass = check_quantity_value(mass, default_mass_unit)
        self.charge = check_quantity_value(charge, default_charge_unit)
        self.bonded_particles = []     # 1-2 and 1-3 partners (matrix ids)
        self.scaling_particles = []    # 1-4 partners
        self.scaling_factors = []      # stored, never read by a kernel

    def __repr__(self):
        return '<mdpy_b200.core.Particle object: %s-%s at %x>' % (self.particle_name, self.particle_id, id(self))

    def __eq__(self, other):
        return self is other

    __hash__ = object.__hash__

    def change_matrix_id(self, matrix_id):
        self.matrix_id = matrix_id

    def add_bonded_particle(self, other_id):
        # particle.py:53-62: duplicates and self-bonds are errors
        if other_id in self.bonded_particles:
            raise ParticleConflictError('Particle %d has been added twice to the bonded_particles of Particle %d'
                                        % (other_id, self.matrix_id))
        if other_id == self.matrix_id:
            raise ParticleConflictError('Particle itself can not be added to the bonded_particle list.')
        self.bonded_particles.append(other_id)

    def del_bonded_particle(self, other_id):
        if other_id in self.bonded_particles:
            self.bonded_particles.remove(other_id)

    def add_scaling_particle(self, other_id, factor=1):
        # particle.py:71-80: self is an error, repeats are dropped silently (benzene case)
        if other_id == self.matrix_id:
            raise ParticleConflictError('Particle itself can not be added to the scaling_particle list.')
        if other_id not in self.scaling_particles:
            self.scaling_particles.append(other_id)
            self.scaling_factors.append(factor)

    def del_scaling_particle(self, other_id):
        if other_id in self.scaling_particles:
            k = self.scaling_particles.index(other_id)
            del self.scaling_particles[k], self.scaling_factors[k]

    num_bonded_particles = property(lambda self: len(self.bonded_particles))
    num_scaling_particles = property(lambda self: len(self.scaling_particles))
