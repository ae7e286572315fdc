"""Langevin dynamics — same constructor and coefficients as
mdpy/integrator/langevin_integrator.py:18-35 (a = (1 - g dt/2)/(1 + g dt/2), b = 1/(1 + g dt/2),
sigma = sqrt(2 kT g)), textbook G-JF update with one force evaluation per step.  The reference's own
update has the force sign and state carry-over wrong and draws numpy noise, so there is
no trajectory parity to keep; this integrator is validated by equipartition instead.
Noise: Philox4x32-10 keyed by `seed`, counter = (atom, step) — reproducible and order independent.
"""
import numpy as np

from ..unit import KB, Quantity, check_quantity_value, default_energy_unit, default_temperature_unit, default_time_unit
from .integrator import Integrator


class LangevinIntegrator(Integrator):
    def __init__(self, time_step, temperature, friction_rate, seed=0):
        super().__init__(time_step)
        self._temperature = check_quantity_value(temperature, default_temperature_unit)
        self._gamma = check_quantity_value(friction_rate, 1 / default_time_unit)
        self._kbt = (Quantity(self._temperature, default_temperature_unit) * KB).convert_to(default_energy_unit).value
        self._sigma = np.sqrt(2 * self._kbt * self._gamma)
        half = self._gamma * self._time_step / 2
        self._a = (1 - half) / (1 + half)
        self._b = 1 / (1 + half)
        self._seed = int(seed)

    def integrate(self, ensemble, num_steps: int = 1):
        """One call = the host State goes in, num_steps steps run on the device, the new host State comes
        out (mdk_step_langevin_host).  The State is uploaded on every call — in-place edits of its arrays
        are honoured — and the device keeps its float64 trajectory wherever the host value still equals
        what the previous call handed out.  The arrays published as the new State are never written again
        while anybody holds a reference to them (EnsembleContext.state_buffers)."""
        ctx, terms = self._bind(ensemble)
        state = ensemble.state
        x_out, v_out = ctx.state_buffers()
        pooled = x_out is not None
        if not pooled:   # every page-locked block is still referenced (kept frames): plain arrays
            x_out = np.empty((ctx.dev.n, 3), dtype=np.float32)
            v_out = np.empty((ctx.dev.n, 3), dtype=np.float32)
        e = ctx.dev.step_langevin_host(self._host_f32(state.positions), self._host_f32(state.velocities), x_out, v_out,
                                       float(self._time_step), float(self._kbt), float(self._gamma), self._seed,
                                       int(num_steps), terms)
        self._publish_state(ensemble, ctx, x_out, v_out, e)
