from .particle import Particle
from .topology import Topology
from .state import State

__all__ = ['Particle', 'Topology', 'State']
