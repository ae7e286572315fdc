"""Electrostatics behind the Constraint API.

ElectrostaticConstraint — drop-in for mdpy/constraint/electrostatic_constraint.py:23-174 with the
semantics of its CPU kernel (:52-79): every pair i<j not in bonded_particles, minimum
image, bare q_i q_j / (4 pi eps0 r), no cutoff.  O(N^2) by definition; float64 on the device.

ElectrostaticPMEConstraint — the north star's PME electrostatics [not in the reference tree]:
erfc direct space over the tile list + smooth PME reciprocal space + self / background /
excluded-pair terms.  alpha from erfc(alpha rc)/rc = ewald_error (the reference's only hint is the
ewald_error=1e-6 default of CharmmForcefield, forcefield/charmm_forcefield.py:23).
"""
import functools
import math

import numpy as np

from .. import _native
from ..environment import env
from ..unit import check_quantity_value, coulomb_constant, default_length_unit
from .constraint import Constraint


class ElectrostaticConstraint(Constraint):
    is_native = True
    terms = _native.TERM_COUL_BARE

    def __init__(self, parameters=None, force_id: int = 0, force_group: int = 0):
        super().__init__(parameters, force_id=force_id, force_group=force_group)

    def __repr__(self):
        return '<mdpy_b200.constraint.ElectrostaticConstraint object>'

    def bind_ensemble(self, ensemble):
        self._attach(ensemble)
        self._configured = False

    def _configure(self):
        if not getattr(self, '_configured', False):
            owner = getattr(self._ctx, 'coulomb_owner', None)
            if owner is None or owner is self:
                self._ctx.dev.set_coulomb(coulomb_constant(), 0.0, 0.0)
                self._ctx.coulomb_owner = self
            self._configured = True


@functools.lru_cache(maxsize=64)
def ewald_alpha(cutoff_radius, ewald_error):
    """alpha such that erfc(alpha rc) / rc == ewald_error (bisection)."""
    lo, hi = 0.0, 10.0 / cutoff_radius
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if math.erfc(mid * cutoff_radius) / cutoff_radius > ewald_error:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def fft_size(minimum):
    """Smallest even 2^a 3^b 5^c 7^d >= minimum (cuFFT-friendly)."""
    n = max(4, int(math.ceil(minimum)))
    while True:
        m = n
        for p in (2, 3, 5, 7):
            while m % p == 0:
                m //= p
        if m == 1 and n % 2 == 0:
            return n
        n += 1


class ElectrostaticPMEConstraint(Constraint):
    is_native = True
    terms = _native.TERM_COUL_DIRECT | _native.TERM_PME_RECIP

    def __init__(self, cutoff_radius=12, ewald_error=1e-6, alpha=None, grid=None, order=4, grid_spacing=1.0,
                 force_id: int = 0, force_group: int = 0):
        super().__init__(None, force_id=force_id, force_group=force_group)
        self._cutoff_radius = check_quantity_value(cutoff_radius, default_length_unit)
        self._ewald_error = float(ewald_error)
        self._alpha = None if alpha is None else float(alpha)
        self._grid = None if grid is None else tuple(int(g) for g in np.broadcast_to(grid, 3))
        self._order = int(order)
        self._grid_spacing = float(grid_spacing)

    def __repr__(self):
        return '<mdpy_b200.constraint.ElectrostaticPMEConstraint object>'

    alpha = property(lambda self: self._alpha if self._alpha is not None else ewald_alpha(float(self._cutoff_radius), self._ewald_error))

    def grid_for(self, box):
        if self._grid is not None:
            return self._grid
        return tuple(fft_size(L / self._grid_spacing) for L in box)

    def bind_ensemble(self, ensemble):
        self._attach(ensemble)
        self._configured = None

    def set_cutoff_radius(self, val):
        super().set_cutoff_radius(val)
        self._configured = None

    def _configure(self):
        box = tuple(self._ctx.check_box())
        key = (float(self._cutoff_radius), self.alpha, self.grid_for(box), self._order)
        if getattr(self, '_configured', None) != key:
            self._ctx.dev.set_coulomb(coulomb_constant(), key[1], key[0])
            self._ctx.dev.set_pme(key[2], key[3])
            self._ctx.coulomb_owner = self
            self._configured = key
