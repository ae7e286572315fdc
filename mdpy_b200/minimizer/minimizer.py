"""Minimizer base — drop-in for mdpy/minimizer/minimizer.py:13-33 (constructor, _energy2str, the
NotImplementedError of the base class)."""
from ..unit import Quantity, default_energy_unit, kilojoule_permol


class Minimizer:
    def __init__(self, output_unit=kilojoule_permol, output_unit_label='kj/mol', is_verbose=False, log_freq=5):
        self._output_unit = output_unit
        self._output_unit_label = output_unit_label
        self._is_verbose = is_verbose
        self._log_freq = log_freq

    def minimize(self, ensemble, energy_tolerance=0.001, max_iterations: int = 1000):
        raise NotImplementedError('The subclass of mdpy.minimizer.Minimizer class should overload minimize method')

    def _energy2str(self, energy):
        return '%.5f %s' % (Quantity(energy, default_energy_unit).convert_to(self._output_unit).value, self._output_unit_label)
