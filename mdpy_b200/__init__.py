"""mdpy_b200 — B200-native (sm_100a) nonbonded hot path behind mdpy's Constraint / Integrator API.

Layout: csrc/ (CUDA kernels + C ABI -> libmdpyb200.so), _native.py (ctypes), and a host-side
mirror of the reference interface for this path only: env, unit, core (Particle/Topology/State),
Ensemble, constraint.*, integrator.*.  io / forcefield parsing / dumpers / analysers stay in mdpy.
"""
SPATIAL_DIM = 3

from .environment import env  # noqa: E402
from . import unit, utils, error, core, constraint, integrator, minimizer  # noqa: E402,F401
from .ensemble import Ensemble  # noqa: E402,F401

__version__ = '0.1.0'
