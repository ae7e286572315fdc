"""Constraint base class — the drop-in boundary on the Python side
(mdpy/constraint/constraint.py:16-73): bind_ensemble / update / forces / potential_energy /
cutoff_radius / set_cutoff_radius / force_id / force_group / parent_ensemble, equality by
identity, NonBoundedError before binding."""
import numpy as np

from .. import _native
from ..environment import env
from ..error import NonBoundedError
from ..unit import check_quantity_value, default_length_unit


class Constraint:
    is_native = False
    terms = 0  # MDK_TERM_* bits this constraint evaluates

    def __init__(self, parameters, force_id: int = 0, force_group: int = 0):
        self._parameters = parameters
        self._force_id = force_id
        self._force_group = force_group
        self._parent_ensemble = None
        self._forces = None
        self._potential_energy = None
        self._cutoff_radius = env.NUMPY_FLOAT(0)
        self._ctx = None
        self._lazy_stamp = None   # set by a fused Ensemble.update: own forces are evaluated on first access

    def __repr__(self):
        return '<mdpy_b200.constraint.Constraint class>'

    __str__ = lambda self: repr(self)

    def __eq__(self, other):
        return self is other

    __hash__ = object.__hash__

    def _check_bound_state(self):
        if self._parent_ensemble is None:
            raise NonBoundedError('%s has not been bounded to any Ensemble instance' % self)

    def _attach(self, ensemble):
        """Common part of bind_ensemble (constraint.py:36 + e.g. charmm_nonbonded_constraint.py:48-50)."""
        self._parent_ensemble = ensemble
        self._force_id = list(ensemble.constraints).index(self) if self in list(ensemble.constraints) else self._force_id
        self._ctx = _native.context_of(ensemble)
        return self._ctx

    def bind_ensemble(self, ensemble):
        raise NotImplementedError('The subclass of Constraint should overload bind_ensemble method')

    def _configure(self):
        """Push this constraint's parameters into the shared device context (idempotent)."""

    def _energy_from(self, energies):
        slots = self.__dict__.get('_energy_slots')
        if slots is None or slots[0] != self.terms:
            slots = self._energy_slots = (self.terms, [s for t, sl in _native.TERM_ENERGY_SLOTS.items() if self.terms & t for s in sl])
        return float(sum(energies[s] for s in slots[1]))

    def update(self):
        """Constraint.update: one device evaluation of this constraint's terms; forces come back as
        env.NUMPY_FLOAT [N,3] in matrix_id order, the energy as float64."""
        self._check_bound_state()
        self._configure()
        e = self._ctx.compute(self.terms)
        self._forces = np.ascontiguousarray(self._ctx.dev.forces(np.float64 if env.NUMPY_FLOAT == np.float64 else np.float32),
                                            dtype=env.NUMPY_FLOAT)
        self._potential_energy = self._energy_from(e)
        self._lazy_stamp = None

    def _defer_forces(self, stamp):
        """Called by the fused Ensemble.update: the sum over all constraints came out of ONE device evaluation;
        this constraint's own forces (ensemble.py:56-59 leaves each constraint with its own) are evaluated when
        somebody reads `.forces`, against the same positions (`stamp` = the State revision they belong to)."""
        self._forces = None
        self._lazy_stamp = stamp

    @property
    def forces(self):
        if self._forces is None and self._lazy_stamp is not None:
            ctx = self._ctx
            if ctx._revision(ctx.ensemble.state) != self._lazy_stamp:
                raise RuntimeError('%s: the State changed since Ensemble.update(); call update() on the constraint '
                                   'for forces at the new positions' % self)
            energy = self._potential_energy
            self.update()
            self._potential_energy = energy   # the value of the fused evaluation (same positions, same term)
        return self._forces

    def set_cutoff_radius(self, val):
        self._cutoff_radius = check_quantity_value(val, default_length_unit)

    force_id = property(lambda self: self._force_id)
    force_group = property(lambda self: self._force_group)
    parent_ensemble = property(lambda self: self._parent_ensemble)
    potential_energy = property(lambda self: self._potential_energy)
    cutoff_radius = property(lambda self: self._cutoff_radius)
