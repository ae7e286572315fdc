"""Force terms behind mdpy's Constraint protocol (mdpy/constraint/__init__.py).

CharmmNonbondedConstraint / ElectrostaticConstraint keep the reference's names, constructor
signatures and results (the names CharmmForcefield.create_ensemble instantiates,
forcefield/charmm_forcefield.py:101-102); CharmmVDWConstraint / ElectrostaticPMEConstraint are the
north star's names for the switched LJ and the PME electrostatics.
"""
import numpy as np

from ..environment import env
from .constraint import Constraint

NUM_NEIGHBOR_CELLS = 27
NEIGHBOR_CELL_TEMPLATE = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)],
                                  dtype=env.NUMPY_INT)
LONG_RANGE_SOLVER = ['PME', 'PPPM']

from .charmm_nonbonded_constraint import CharmmNonbondedConstraint, CharmmVDWConstraint  # noqa: E402
from .electrostatic_constraint import ElectrostaticConstraint, ElectrostaticPMEConstraint  # noqa: E402
from .charmm_bonded_constraint import (CharmmAngleConstraint, CharmmBondConstraint,  # noqa: E402
                                       CharmmDihedralConstraint, CharmmImproperConstraint)

__all__ = ['Constraint', 'ElectrostaticConstraint', 'ElectrostaticPMEConstraint', 'CharmmNonbondedConstraint',
           'CharmmVDWConstraint', 'CharmmBondConstraint', 'CharmmAngleConstraint', 'CharmmDihedralConstraint',
           'CharmmImproperConstraint']
