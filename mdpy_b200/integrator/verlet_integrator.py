"""Position-Verlet — drop-in for mdpy/integrator/verlet_integrator.py:15-50.

reference_quirks=True (default) reproduces the reference's recurrences exactly, including the two
SURVEY Q4 points: the first step uses a dt^2 (not a dt^2 / 2) and the velocity written back to
State is minimg(x_n - x_n-1) / (2 dt).  reference_quirks=False gives the textbook start and
v_n = (x_n - x_n-1)/dt + a_n dt/2.
"""
from .integrator import Integrator


class VerletIntegrator(Integrator):
    def __init__(self, time_step, reference_quirks=True):
        super().__init__(time_step)
        self._time_step_square = self._time_step ** 2
        self._reference_quirks = bool(reference_quirks)

    def integrate(self, ensemble, num_steps: int = 1):
        ctx, terms = self._prepare(ensemble)
        ctx.dev.step_verlet(float(self._time_step), int(num_steps), terms, self._reference_quirks)
        self._publish(ensemble, ctx, terms)
