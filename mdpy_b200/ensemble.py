"""Ensemble: a topology, its State and the constraints acting on it — the caller of the drop-in
boundary.  Interface of mdpy/ensemble.py:16-102 (add_constraints, update, forces / energies,
constraints / num_constraints), kept so that reference code driving an Ensemble keeps working.

What differs from the reference is where the sum over constraints happens.  The reference loops over
its constraints and adds their float arrays on the host (ensemble.py:53-61).  Here, when every bound
constraint is native and they share one device context, `update()` issues ONE force evaluation for the
union of their terms; each constraint keeps its own energy, and its own `.forces`
(ensemble.py:56-59 leaves every constraint with its own array) are evaluated lazily, on first access,
for the same positions.  Foreign constraints (anything without `is_native`) fall back to the reference's
per-constraint protocol.
"""
import numpy as np

from .core import State, Topology
from .error import ConstraintConflictError
from .unit import Quantity, default_energy_unit, default_mass_unit, default_velocity_unit

_KE_TO_ENERGY = None  # conversion factor Da (A/fs)^2 -> internal energy unit, resolved on first use


def _kinetic_energy_of(velocities, masses):
    """sum m v^2 / 2 in the internal energy unit, accumulated in float64."""
    global _KE_TO_ENERGY
    if _KE_TO_ENERGY is None:
        _KE_TO_ENERGY = Quantity(1.0, default_velocity_unit ** 2 * default_mass_unit).convert_to(default_energy_unit).value
    v = np.asarray(velocities, dtype=np.float64)
    m = np.asarray(masses, dtype=np.float64).reshape(-1)
    return 0.5 * float(np.einsum('i,ij,ij->', m, v, v)) * float(_KE_TO_ENERGY)


class Ensemble:
    def __init__(self, topology: Topology, pbc_matrix):
        if not topology.is_joined:
            topology.join()
        self._topology = topology
        self._state = State(topology, pbc_matrix)
        self._matrix_shape = self._state.matrix_shape
        self._constraints = []
        self._native = None   # device context shared by the native constraints, created by the first one bound
        self._forces = np.zeros(self._matrix_shape)
        self._potential_energy = self._kinetic_energy = self._total_energy = 0

    def __repr__(self):
        return '<mdpy_b200.Ensemble object: %d constraints at %x>' % (len(self._constraints), id(self))

    # ---- constraints ----------------------------------------------------------------------------------
    def add_constraints(self, *constraints):
        """Bind constraints in order; the same object twice is a ConstraintConflictError (ensemble.py:40-46)
        and the State's cutoff guard follows the largest cutoff bound so far (ensemble.py:49-50)."""
        guard = self._state.cell_list
        for new in constraints:
            for bound in self._constraints:
                if bound is new:
                    raise ConstraintConflictError('%s has added twice to %s' % (new, self))
            self._constraints.append(new)
            new.bind_ensemble(self)
            if new.cutoff_radius > guard.cutoff_radius:
                guard.set_cutoff_radius(new.cutoff_radius)

    def _all_native(self):
        return self._native is not None and all(getattr(c, 'is_native', False) for c in self._constraints)

    # ---- one force / energy evaluation ----------------------------------------------------------------
    def update(self, fused=True):
        """Forces (float64 [N,3]) and potential energy summed over the constraints, kinetic energy from the
        State's velocities, total = potential + kinetic."""
        if fused and len(self._constraints) > 1 and self._all_native():
            forces, potential = self._native.compute_fused(self._constraints)
            total_force = np.array(forces, dtype=np.float64)
        else:
            total_force = np.zeros(self._matrix_shape)
            potential = 0
            for c in self._constraints:
                c.update()
                total_force += c.forces
                potential += c.potential_energy
        self._forces = total_force
        self._potential_energy = potential
        self._kinetic_energy = _kinetic_energy_of(self._state.velocities, self._topology.masses)
        self._total_energy = self._potential_energy + self._kinetic_energy

    # ---- read-only views ------------------------------------------------------------------------------
    @property
    def topology(self):
        return self._topology

    @property
    def state(self):
        return self._state

    @property
    def constraints(self):
        return self._constraints

    @property
    def num_constraints(self):
        return len(self._constraints)

    @property
    def forces(self):
        return self._forces

    @property
    def potential_energy(self):
        return self._potential_energy

    @property
    def kinetic_energy(self):
        return self._kinetic_energy

    @property
    def total_energy(self):
        return self._total_energy
