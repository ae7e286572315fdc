"""CHARMM Lennard-Jones — drop-in for mdpy/constraint/charmm_nonbonded_constraint.py:21-229.

Same constructor `(parameters, cutoff_radius=12, force_id=0, force_group=0)`, same per-atom
[eps, sigma, eps14, sigma14] table (:48-62), same physics: plain inclusive cutoff, 1-2/1-3
exclusion, 1-4 alternate parameters, Lorentz-Berthelot mixing (:83-107).  The evaluation is one
launch of the sm_100a tile-list pair kernel (csrc/mdk_pair.cu) instead of the numba kernels.
`switch_radius` (keyword only) turns on the CHARMM energy switch — the north star's
"CharmmVDWConstraint"; left at None it reproduces the reference exactly.
"""
import numpy as np

from .. import _native
from ..environment import env
from ..unit import check_quantity_value, default_length_unit
from .constraint import Constraint


class CharmmNonbondedConstraint(Constraint):
    is_native = True
    terms = _native.TERM_LJ

    def __init__(self, parameters, cutoff_radius=12, force_id: int = 0, force_group: int = 0, *, switch_radius=None):
        super().__init__(parameters, force_id=force_id, force_group=force_group)
        self._cutoff_radius = check_quantity_value(cutoff_radius, default_length_unit)
        self._switch_radius = check_quantity_value(switch_radius, default_length_unit)
        self._parameters_list = []
        self._num_nonbonded_pairs = 0

    def __repr__(self):
        return '<mdpy_b200.constraint.CharmmNonbondedConstraint object>'

    def bind_ensemble(self, ensemble):
        self._attach(ensemble)
        topo = ensemble.topology
        if isinstance(self._parameters, np.ndarray):  # ready per-atom [N,4] table (bulk systems, fixtures)
            self._parameters_list = np.ascontiguousarray(self._parameters, dtype=env.NUMPY_FLOAT).reshape(topo.num_particles, 4)
            self._configured = None
            return
        types = topo.particle_types if hasattr(topo, 'particle_types') else [p.particle_type for p in topo.particles]
        # one table row per distinct type, then a gather: [eps, sigma] -> [eps, sigma, eps, sigma] (:54-59)
        rows = {}
        for name in set(types):
            p = list(self._parameters[name])
            rows[name] = p + p if len(p) == 2 else p
        self._parameters_list = np.array([rows[name] for name in types], dtype=env.NUMPY_FLOAT).reshape(len(types), 4)
        self._configured = None

    def set_cutoff_radius(self, val):
        super().set_cutoff_radius(val)
        self._configured = None

    def _configure(self):
        key = (float(self._cutoff_radius), None if self._switch_radius is None else float(self._switch_radius))
        if getattr(self, '_configured', None) != key:
            self._ctx.dev.set_lj(self._parameters_list, key[0], key[1])
            self._configured = key

    @property
    def num_nonbonded_pairs(self):
        return self._num_nonbonded_pairs

    def neighbor_pairs(self):
        """The in-cutoff, non-excluded pair set the tile list yields, sorted (i<j) rows of matrix ids."""
        self._check_bound_state()
        self._configure()
        self._ctx.sync_positions()
        pairs = self._ctx.dev.pairs()
        self._num_nonbonded_pairs = len(pairs)
        return pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]


class CharmmVDWConstraint(CharmmNonbondedConstraint):
    """CHARMM LJ with the energy switch on (switch_radius, cutoff_radius]."""

    def __init__(self, parameters, cutoff_radius=12, switch_radius=10, force_id: int = 0, force_group: int = 0):
        super().__init__(parameters, cutoff_radius, force_id, force_group, switch_radius=switch_radius)

    def __repr__(self):
        return '<mdpy_b200.constraint.CharmmVDWConstraint object>'
