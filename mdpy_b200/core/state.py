"""State — positions, velocities and the periodic box (drop-in for mdpy/core/state.py:18-107).

The reference's State owns a host CellList that is rebuilt on every set_positions
(state.py:56-61).  Here the neighbour structure lives on the device inside the native
context (tile list, rebuilt only when an atom has moved skin/2); State keeps the host
arrays, wraps like the reference, and counts a `revision` so bound constraints know when
the device copy is stale.  `cell_list` survives as a thin object carrying the cutoff that
Ensemble.add_constraints negotiates (ensemble.py:49-50) and raising the same
CellListPoorDefinedError (cell_list.py:56-69).
"""
import numpy as np

from ..environment import env
from ..error import ArrayDimError, CellListPoorDefinedError
from ..unit import (KB, Quantity, check_quantity, check_quantity_value, default_length_unit, default_mass_unit,
                    default_temperature_unit, default_velocity_unit)
from ..utils import SPATIAL_DIM, check_pbc_matrix, wrap_positions


class CutoffGuard:
    """What is left of mdpy.core.CellList on the host: the cutoff bookkeeping."""

    def __init__(self, pbc_matrix, cutoff_radius=12):
        self._pbc_diag = np.asarray(pbc_matrix).diagonal().astype(np.float64)
        self._cutoff_radius = env.NUMPY_FLOAT(0)
        self.set_cutoff_radius(cutoff_radius)

    def set_pbc_matrix(self, pbc_matrix):
        self._pbc_diag = np.asarray(check_pbc_matrix(check_quantity_value(pbc_matrix, default_length_unit))).diagonal().astype(np.float64)

    def set_cutoff_radius(self, cutoff_radius):
        rc = check_quantity_value(cutoff_radius, default_length_unit)
        if rc == 0:
            raise CellListPoorDefinedError('Cutoff radius is poor defined, current value %.3f' % rc)
        if (np.floor(self._pbc_diag / float(rc)) < 2).any():
            raise CellListPoorDefinedError('The cutoff_radius is too large to create cell list')
        self._cutoff_radius = rc

    def update(self, positions):  # kept for call-site compatibility; the tile list is device side
        return None

    cutoff_radius = property(lambda self: self._cutoff_radius)


class State:
    def __init__(self, topology, pbc_matrix):
        self._num_particles = topology.num_particles
        self._masses = topology.masses
        self._matrix_shape = [self._num_particles, SPATIAL_DIM]
        self._positions = np.zeros(self._matrix_shape, dtype=env.NUMPY_FLOAT)
        self._velocities = np.zeros(self._matrix_shape, dtype=env.NUMPY_FLOAT)
        self.revision = 0  # bumped whenever positions or the box change on the host
        self.set_pbc_matrix(pbc_matrix)
        # like the reference (state.py:28 -> cell_list.py:20) the guard starts at 12 A, so a box edge
        # under 24 A raises CellListPoorDefinedError right here
        self._cell_list = CutoffGuard(self._pbc_matrix, 12)

    def __repr__(self):
        return '<mdpy_b200.core.State object with %d particles at %x>' % (self._num_particles, id(self))

    def _check_matrix_shape(self, matrix):
        if not isinstance(matrix, np.ndarray):
            raise TypeError('Matrix should be numpy.ndarray, instead of %s' % type(matrix))
        if matrix.ndim != 2 or list(matrix.shape) != self._matrix_shape:
            raise ArrayDimError('The dimension of array should be [%d, %d], while array %s is provided'
                                % (self._matrix_shape[0], self._matrix_shape[1], list(matrix.shape)))

    def set_pbc_matrix(self, pbc_matrix):
        pbc_matrix = check_pbc_matrix(check_quantity_value(pbc_matrix, default_length_unit))
        self._pbc_matrix = np.ascontiguousarray(pbc_matrix, dtype=env.NUMPY_FLOAT)
        self._pbc_inv = np.ascontiguousarray(np.linalg.inv(self._pbc_matrix), dtype=env.NUMPY_FLOAT)
        if hasattr(self, '_cell_list'):
            self._cell_list.set_pbc_matrix(self._pbc_matrix)  # keep the guard's box current
        self.revision += 1

    def set_positions(self, positions):
        self._check_matrix_shape(positions)
        self._positions = wrap_positions(positions.astype(env.NUMPY_FLOAT), self._pbc_matrix, self._pbc_inv)
        self.revision += 1

    def set_velocities(self, velocities):
        self._check_matrix_shape(velocities)
        self._velocities = velocities.astype(env.NUMPY_FLOAT)

    def set_velocities_to_temperature(self, temperature, seed=None):
        """state.py:66-80: uniform in [-w, w] with w = sqrt(3 kB T / m) per component."""
        temperature = check_quantity(temperature, default_temperature_unit)
        factor = (Quantity(3) * KB * temperature / default_mass_unit).convert_to(default_velocity_unit ** 2).value
        width = np.sqrt(np.float64(factor) / np.asarray(self._masses, dtype=np.float64).reshape(-1, 1))
        rng = np.random.default_rng(seed)
        self.set_velocities(((rng.random(self._matrix_shape) * 2 - 1) * width).astype(env.NUMPY_FLOAT))

    positions = property(lambda self: self._positions)
    velocities = property(lambda self: self._velocities)
    matrix_shape = property(lambda self: self._matrix_shape)
    pbc_matrix = property(lambda self: self._pbc_matrix)
    pbc_inv = property(lambda self: self._pbc_inv)
    pbc_info = property(lambda self: (self._pbc_matrix, self._pbc_inv))
    cell_list = property(lambda self: self._cell_list)
