"""Host-side helpers of the hot path: periodic wrapping and minimum image
(drop-ins for mdpy/utils/pbc.py:16-44 and mdpy/utils/geometry.py:17-20)."""
import numpy as np

from .error import ArrayDimError, ParticleLossError, PBCPoorDefinedError
from .unit import check_quantity, check_quantity_value  # noqa: F401  (re-exported like mdpy.utils)

SPATIAL_DIM = 3


def check_pbc_matrix(pbc_matrix):
    """mdpy/utils/pbc.py:16-26: 3x3 and non-singular."""
    pbc_matrix = np.asarray(pbc_matrix)
    if pbc_matrix.shape != (SPATIAL_DIM, SPATIAL_DIM):
        raise ArrayDimError('The pbc matrix should have shape [3, 3], while matrix %s is provided'
                            % (list(pbc_matrix.shape),))
    if np.linalg.det(pbc_matrix) == 0:
        raise PBCPoorDefinedError('PBC is poor defined. Two or more column vectors are linear corellated')
    return pbc_matrix


def wrap_positions(positions, pbc_matrix, pbc_inv):
    """mdpy/utils/pbc.py:28-36: x - round(x . pbc_inv) . pbc, ParticleLossError at >= 2 images."""
    shift = -np.round(positions @ pbc_inv)
    lost = np.abs(shift) >= 2
    if lost.any():
        raise ParticleLossError('Atom(s) with matrix id: %s moved beyond 2 PBC image.'
                                % np.unique(np.nonzero(lost)[0]))
    return positions + shift @ pbc_matrix


def unwrap_vec(vec, pbc_matrix, pbc_inv):
    """mdpy/utils/pbc.py:38-44: minimum-image vector(s)."""
    s = np.asarray(vec) @ pbc_inv
    s = s - np.round(s)
    return s @ pbc_matrix


def get_unit_vec(vec):
    """mdpy/utils/geometry.py:17-20."""
    vec = np.asarray(vec)
    norm = np.linalg.norm(vec)
    return vec / norm if norm != 0 else vec
