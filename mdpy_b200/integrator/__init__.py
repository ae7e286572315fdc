"""Device-resident integrators behind the reference's Integrator API (mdpy/integrator): the whole
integrate(ensemble, num_steps) loop — forces, position / velocity update, wrap, neighbour-list upkeep —
runs on the GPU through mdk_step_verlet / mdk_step_langevin_host."""
from .integrator import Integrator
from .langevin_integrator import LangevinIntegrator
from .verlet_integrator import VerletIntegrator

__all__ = ['Integrator', 'LangevinIntegrator', 'VerletIntegrator']
