"""Ensemble — binds constraints to a (topology, state) pair and sums their forces/energies:
the caller of the drop-in boundary (mdpy/ensemble.py:16-102).

`update()` keeps the reference's contract (every constraint's update(), float64 sums,
ensemble.py:53-61).  When all bound constraints are native and share one device context
the per-constraint passes are fused into a single mdk_compute; the
per-constraint `.forces` then come from the one shared accumulator and only their *sum* is
meaningful, which is exactly what Ensemble exposes.
"""
import numpy as np

from .core import State, Topology
from .error import ConstraintConflictError
from .unit import Quantity, default_energy_unit, default_mass_unit, default_velocity_unit


class Ensemble:
    def __init__(self, topology: Topology, pbc_matrix):
        if not topology.is_joined:
            topology.join()
        self._topology = topology
        self._state = State(topology, pbc_matrix)
        self._matrix_shape = self._state.matrix_shape
        self._forces = np.zeros(self._matrix_shape)
        self._total_energy = self._potential_energy = self._kinetic_energy = 0
        self._constraints = []
        self._native = None  # shared device context, created by the first native constraint

    def __repr__(self):
        return '<mdpy_b200.Ensemble object: %d constraints at %x>' % (self.num_constraints, id(self))

    def add_constraints(self, *constraints):
        for constraint in constraints:
            if any(constraint is c for c in self._constraints):
                raise ConstraintConflictError('%s has added twice to %s' % (constraint, self))
            self._constraints.append(constraint)
            constraint.bind_ensemble(self)
            if constraint.cutoff_radius > self._state.cell_list.cutoff_radius:
                self._state.cell_list.set_cutoff_radius(constraint.cutoff_radius)

    def update(self, fused=True):
        self._forces = np.zeros(self._matrix_shape)
        self._potential_energy = 0
        native = [c for c in self._constraints if getattr(c, 'is_native', False)]
        if fused and self._native is not None and len(native) == len(self._constraints) and len(native) > 1:
            forces, energy = self._native.compute_fused(native)
            self._forces += forces
            self._potential_energy += energy
        else:
            for constraint in self._constraints:
                constraint.update()
                self._forces += constraint.forces
                self._potential_energy += constraint.potential_energy
        self._update_kinetic_energy()
        self._total_energy = self._potential_energy + self._kinetic_energy

    def _update_kinetic_energy(self):
        v = np.asarray(self._state.velocities, dtype=np.float64)
        m = np.asarray(self._topology.masses, dtype=np.float64).reshape(-1)
        ke = 0.5 * float(((v ** 2).sum(1) * m).sum())
        self._kinetic_energy = Quantity(ke, default_velocity_unit ** 2 * default_mass_unit).convert_to(default_energy_unit).value

    topology = property(lambda self: self._topology)
    state = property(lambda self: self._state)
    forces = property(lambda self: self._forces)
    total_energy = property(lambda self: self._total_energy)
    potential_energy = property(lambda self: self._potential_energy)
    kinetic_energy = property(lambda self: self._kinetic_energy)
    constraints = property(lambda self: self._constraints)
    num_constraints = property(lambda self: len(self._constraints))
