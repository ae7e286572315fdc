"""A small unit system with the reference's default units and constants.

Only what the hot path touches of mdpy/unit/ (1130 lines there): `Quantity(value, unit)`,
`.convert_to(unit).value`, products / quotients / powers of units, the default units
(angstrom, femtosecond, dalton, e, kelvin — mdpy/unit/__init__.py:31-36), the energy units the
CHARMM parameter files use, and KB / NA / EPSILON0.

SURVEY Q7: the reference's constants are rounded to float32 at import
(EPSILON0.value == 0.5727653, kcal/mol == 4.1840284e-4 internal); the same float32 values
are reproduced here so that energies agree to the last digit (pinned by
tests/golden/reference_constants.json).
"""
import numpy as np

from .error import UnitDimensionDismatchedError

_DIMS = ('length', 'mass', 'time', 'temperature', 'charge', 'mol')


class Unit:
    __slots__ = ('dim', 'factor')

    def __init__(self, dim, factor):
        self.dim = tuple(dim)
        self.factor = float(factor)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit([a + b for a, b in zip(self.dim, other.dim)], self.factor * other.factor)
        return Quantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit([a - b for a, b in zip(self.dim, other.dim)], self.factor / other.factor)
        return Quantity(1.0 / other, self)

    def __rtruediv__(self, other):
        return Quantity(other, Unit([-a for a in self.dim], 1.0 / self.factor))

    def __pow__(self, p):
        return Unit([a * p for a in self.dim], self.factor ** p)

    def __eq__(self, other):
        return (isinstance(other, Unit) and self.dim == other.dim
                and abs(self.factor / other.factor - 1) < 1e-6)

    def __hash__(self):
        return hash(self.dim)

    def __repr__(self):
        return '<Unit %s x %g>' % (dict(zip(_DIMS, self.dim)), self.factor)


def _base(**kw):
    return [kw.get(k, 0) for k in _DIMS]


no_unit = Unit(_base(), 1)
meter = Unit(_base(length=1), 1)
nanometer = Unit(_base(length=1), 1e-9)
angstrom = Unit(_base(length=1), 1e-10)
kilogram = Unit(_base(mass=1), 1)
dalton = amu = Unit(_base(mass=1), 1.66053904e-27)
second = Unit(_base(time=1), 1)
nanosecond = Unit(_base(time=1), 1e-9)
picosecond = Unit(_base(time=1), 1e-12)
femtosecond = Unit(_base(time=1), 1e-15)
kelvin = Unit(_base(temperature=1), 1)
coulomb = Unit(_base(charge=1), 1)
e = Unit(_base(charge=1), 1.602176634e-19)
mol = Unit(_base(mol=1), 1)
_energy = _base(mass=1, length=2, time=-2)
joule = Unit(_energy, 1)
kilojoule_permol = Unit(_energy, 1e3 / 6.0221e23)
kilocalorie_permol = Unit(_energy, 4.184e3 / 6.0221e23)
newton = Unit(_base(mass=1, length=1, time=-2), 1)
kilocalorie_permol_over_angstrom = kilocalorie_permol / angstrom
kilojoule_permol_over_nanometer = kilojoule_permol / nanometer

default_length_unit = angstrom
default_mass_unit = dalton
default_time_unit = femtosecond
default_temperature_unit = kelvin
default_charge_unit = e
default_mol_unit = mol
default_velocity_unit = default_length_unit / default_time_unit
default_energy_unit = default_mass_unit * default_length_unit ** 2 / default_time_unit ** 2
default_force_unit = default_energy_unit / default_length_unit


class Quantity:
    """value * unit; values are stored as env.NUMPY_FLOAT like mdpy/unit/quantity.py:33-37."""

    def __init__(self, value, unit: Unit = no_unit):
        from .environment import env
        self.unit = unit
        self.value = np.asarray(value, dtype=env.NUMPY_FLOAT) if np.ndim(value) else env.NUMPY_FLOAT(value)

    def convert_to(self, target: Unit):
        if tuple(self.unit.dim) != tuple(target.dim):
            raise UnitDimensionDismatchedError('%s can not be converted to %s' % (self.unit, target))
        return Quantity(np.float64(self.value) * (self.unit.factor / target.factor), target)

    def _binary(self, other, op):
        if isinstance(other, Quantity):
            return Quantity(op(np.float64(self.value), np.float64(other.value)), op(self.unit, other.unit))
        if isinstance(other, Unit):
            return Quantity(self.value, op(self.unit, other))
        return Quantity(op(np.float64(self.value), other), self.unit)

    def __mul__(self, other):
        return self._binary(other, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self._binary(other, lambda a, b: a / b)

    def __rtruediv__(self, other):
        return Quantity(other / np.float64(self.value), no_unit / self.unit)

    def __pow__(self, p):
        return Quantity(np.float64(self.value) ** p, self.unit ** p)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __repr__(self):
        return '<Quantity %s %s>' % (self.value, self.unit)


def check_quantity_value(val, target_unit: Unit):
    """mdpy/utils/check_quantity.py:20-26."""
    if val is None:
        return None
    if isinstance(val, Quantity):
        return val.convert_to(target_unit).value
    return Quantity(val, target_unit).value


def check_quantity(val, target_unit: Unit):
    """mdpy/utils/check_quantity.py:12-18."""
    if val is None:
        return None
    if isinstance(val, Quantity):
        return val.convert_to(target_unit)
    return Quantity(val, target_unit)


# Constants, float32-rounded exactly like the reference's (mdpy/unit/__init__.py:54-58).
KB = Quantity(np.float32(1.38064852e-23), Unit(_base(mass=1, length=2, time=-2, temperature=-1), 1))
NA = Quantity(np.float32(6.0221e23), Unit(_base(mol=-1), 1))
EPSILON0 = Quantity(np.float32(0.5727653),
                    default_time_unit ** 2 * default_charge_unit ** 2 / default_length_unit ** 3 / default_mass_unit)
RMIN_TO_SIGMA_FACTOR = np.float32(2 ** (-1 / 6))  # mdpy/io/charmm_toppar_parser.py:16


def coulomb_constant():
    """k_e = 1 / (4 pi eps0) in internal units, from the float32-rounded EPSILON0 the same way
    the reference's kernels see it (electrostatic_constraint.py:21,60)."""
    return 1.0 / (4.0 * np.pi * float(np.float32(EPSILON0.value)))
