"""Energy minimizers behind the reference's Minimizer API (mdpy/minimizer), device resident."""
from .minimizer import Minimizer
from .steepest_descent_minimizer import SteepestDescentMinimizer

__all__ = ['Minimizer', 'SteepestDescentMinimizer']
