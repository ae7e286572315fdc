"""Exception classes of the hot path, same names as mdpy/error.py (:10, :37, :52, :69, :78,
:87, :113, :123, :203) so reference-side `except` clauses and tests keep working."""


class EnvironmentVariableError(Exception):
    """Unsupported precision / platform name (mdpy/environment.py:39,49)."""


class UnitDimensionDismatchedError(Exception):
    """Conversion between units of different dimension."""


class ArrayDimError(Exception):
    """A position / velocity / pbc array has the wrong shape (mdpy/core/state.py:36-45)."""


class ParticleConflictError(Exception):
    """A particle was registered twice as a bonded partner (mdpy/core/particle.py:53-62)."""


class ConstraintConflictError(Exception):
    """A constraint was added twice to one Ensemble (mdpy/ensemble.py:42-46)."""


class ModifyJoinedTopologyError(Exception):
    """Topology edited after join() (mdpy/core/topology.py:53-57)."""


class NonBoundedError(Exception):
    """Constraint used before bind_ensemble (mdpy/constraint/constraint.py:39-43)."""


class PBCPoorDefinedError(Exception):
    """Singular periodic box (mdpy/utils/pbc.py:22-25)."""


class CellListPoorDefinedError(Exception):
    """Cutoff of 0 or larger than half the box (mdpy/core/cell_list.py:58-69)."""


class GeomtryDimError(Exception):
    """Bond / angle / dihedral with the wrong number of particles (mdpy/core/topology.py:125)."""


class ParticleLossError(Exception):
    """An atom moved two or more periodic images away (mdpy/utils/pbc.py:30-34)."""
