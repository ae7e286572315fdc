from .integrator import Integrator
from .verlet_integrator import VerletIntegrator
from .langevin_integrator import LangevinIntegrator

__all__ = ['Integrator', 'VerletIntegrator', 'LangevinIntegrator']
