"""Precision / platform switch — the drop-in for mdpy/environment.py:14-82.

Same singleton, same method names, same exceptions.  Differences that follow from the
north star (no CPU fallback, B200 only):
  * the default platform is 'CUDA'; 'CPU' is still an accepted *name* (so reference scripts
    that call env.set_platform('CPU') do not raise at that line) but the native constraints
    refuse to run on it and say where the CPU path lives (the reference itself);
  * 'SINGLE': fp32 pair arithmetic with int64 fixed-point force/energy accumulation and fp64
    integrator state; 'DOUBLE' only widens the host-side arrays (NUMPY_FLOAT = float64), the
    kernels are the same — see DESIGN.md "precision".
"""
import numpy as np

from .error import EnvironmentVariableError


class MDPYEnvironment:
    _PRECISIONS = {'SINGLE': (np.float32, np.int32), 'DOUBLE': (np.float64, np.int64)}
    _PLATFORMS = ('CPU', 'CUDA')

    def __init__(self):
        self.set_default()

    def set_precision(self, precision: str):
        key = str(precision).upper()
        if key not in self._PRECISIONS:
            raise EnvironmentVariableError(
                'Precision %s is not supported. Check supported precision with '
                '`mdpy_b200.env.supported_precisions`' % key)
        self._precision = key
        self.NUMPY_FLOAT, self.NUMPY_INT = self._PRECISIONS[key]

    def set_platform(self, platform: str):
        key = str(platform).upper()
        if key not in self._PLATFORMS:
            raise EnvironmentVariableError(
                'Platform %s is not supported. Check supported platform with '
                '`mdpy_b200.env.supported_platforms`' % key)
        self._platform = key

    def set_default(self):
        self.set_precision('SINGLE')
        self.set_platform('CUDA')

    precision = property(lambda self: self._precision)
    platform = property(lambda self: self._platform)
    supported_precisions = property(lambda self: list(self._PRECISIONS))
    supported_presisions = supported_precisions  # the reference's spelling (environment.py:58)
    supported_platforms = property(lambda self: list(self._PLATFORMS))
    default_precision = property(lambda self: 'SINGLE')
    default_platform = property(lambda self: 'CUDA')


env = MDPYEnvironment()
