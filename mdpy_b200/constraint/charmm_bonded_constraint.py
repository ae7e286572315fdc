"""CHARMM bonded terms on the device: bond, angle (+Urey-Bradley), dihedral,
improper, same constructor `(parameters, force_id=0, force_group=0)` and parameter-dictionary keys
('A-B', 'A-B-C', 'A-B-C-D' -> lists in internal units) as mdpy/constraint/charmm_bond_constraint.py:32-51,
charmm_angle_constraint.py:32-53, charmm_dihedral_constraint.py:33-56, charmm_improper_constraint.py:32-55.
Energies follow the reference's formulas; forces are the analytic gradients of those energies (the
reference's dihedral force at charmm_dihedral_constraint.py:80 is not — see DESIGN.md quirk Q12).
`parameters` may also be a ready [n_terms, n_par] array (bulk synthetic systems).
"""
import numpy as np

from .. import _native
from ..environment import env
from .constraint import Constraint


class _BondedConstraint(Constraint):
    is_native = True
    _kind, _n_idx, _n_par, _attr, _name = 0, 2, 2, 'bonds', 'Bond'

    def __repr__(self):
        return '<mdpy_b200.constraint.Charmm%sConstraint object>' % self._name

    def bind_ensemble(self, ensemble):
        self._attach(ensemble)
        topo = ensemble.topology
        idx = np.asarray(getattr(topo, self._attr), dtype=np.int64).reshape(-1, self._n_idx)
        if isinstance(self._parameters, dict):
            types = topo.particle_types if hasattr(topo, 'particle_types') else [p.particle_type for p in topo.particles]
            rows = [self._lookup('-'.join(types[i] for i in term)) for term in idx]
            par = np.array(rows, dtype=np.float64).reshape(len(rows), -1)
        else:
            par = np.asarray(self._parameters, dtype=np.float64).reshape(idx.shape[0], -1)
        if par.shape[1] < self._n_par:  # e.g. angles without Urey-Bradley columns
            par = np.concatenate([par, np.zeros((par.shape[0], self._n_par - par.shape[1]))], axis=1)
        self._int_parameters = idx.astype(env.NUMPY_INT)
        self._float_parameters = par[:, :self._n_par].astype(env.NUMPY_FLOAT)
        self._configured = False

    def _lookup(self, key):
        return list(self._parameters[key])

    def _configure(self):
        if not self._configured:
            self._ctx.dev.set_bonded(self._kind, self._int_parameters, self._float_parameters)
            self._configured = True


class CharmmBondConstraint(_BondedConstraint):
    terms = _native.TERM_BOND


class CharmmAngleConstraint(_BondedConstraint):
    terms = _native.TERM_ANGLE
    _kind, _n_idx, _n_par, _attr, _name = 1, 3, 4, 'angles', 'Angle'


class CharmmDihedralConstraint(_BondedConstraint):
    terms = _native.TERM_DIHEDRAL
    _kind, _n_idx, _n_par, _attr, _name = 2, 4, 3, 'dihedrals', 'Dihedral'


class CharmmImproperConstraint(_BondedConstraint):
    terms = _native.TERM_IMPROPER
    _kind, _n_idx, _n_par, _attr, _name = 3, 4, 2, 'impropers', 'Improper'
