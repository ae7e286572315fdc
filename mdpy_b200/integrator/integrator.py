"""Integrator base — drop-in for mdpy/integrator/integrator.py:14-50, device resident.

`integrate(ensemble, num_steps)` runs the whole loop (force evaluation + position update +
wrap + neighbour-list upkeep) on the GPU through mdk_step_*; positions and velocities come back
to `ensemble.state` once per call instead of once per step.  The step cache
(`is_cached` / `erase_cache`) lives in the device context; uploading new host positions drops it,
as building a new reference integrator would.
"""
import numpy as np

from .. import _native
from ..environment import env
from ..unit import check_quantity_value, default_time_unit


class Integrator:
    def __init__(self, time_step):
        self._time_step = check_quantity_value(time_step, default_time_unit)
        self._cur_positions = None
        self._pre_positions = None
        self._ctx = None

    def erase_cache(self):
        self._cur_positions = None
        self._pre_positions = None
        if self._ctx is not None:
            self._ctx.dev.reset_integrator()

    # ---- helpers shared by the concrete integrators ---------------------------------------
    def _prepare(self, ensemble):
        """Device context of the ensemble, parameters pushed, state mirrored.  Returns (ctx, terms)."""
        ctx = _native.context_of(ensemble)
        terms = 0
        for c in ensemble.constraints:
            if not getattr(c, 'is_native', False):
                raise TypeError('%s is not a native constraint: the device integrators need every force '
                                'term on the GPU' % c)
            c._configure()
            terms |= c.terms
        state = ensemble.state
        fresh = ctx._pos_rev != ctx._revision(state) or ctx.integrator_owner is not self
        if fresh:
            ctx._pos_rev = None
            ctx.sync_positions()
            ctx.dev.upload_velocities(state.velocities)
            ctx.dev.reset_integrator()
            ctx.integrator_owner = self
            self._cur_positions = None
        self._ctx = ctx
        return ctx, terms

    def _publish(self, ensemble, ctx, terms):
        """Bring the final state and energies back to the host objects."""
        state = ensemble.state
        state._positions = ctx.dev.download_positions().astype(env.NUMPY_FLOAT)
        state._velocities = ctx.dev.download_velocities().astype(env.NUMPY_FLOAT)
        if hasattr(state, 'revision'):
            state.revision += 1
        ctx.mark_positions_current()
        self._cur_positions = ctx.dev.download_positions(unwrapped=True)
        e = ctx.dev.last_energies()
        pot = 0.0
        for c in ensemble.constraints:
            c._potential_energy = c._energy_from(e)
            pot += c._potential_energy
        ensemble._potential_energy = pot
        ensemble._kinetic_energy = float(e[_native.E_KINETIC])
        ensemble._total_energy = pot + ensemble._kinetic_energy

    def integrate(self, ensemble, num_steps: int = 1):
        raise NotImplementedError('The subclass of Integrator should overload integrate method')

    @property
    def time_step(self):
        return self._time_step

    @time_step.setter
    def time_step(self, time_step):
        self._time_step = check_quantity_value(time_step, default_time_unit)

    cur_positions = property(lambda self: self._cur_positions)
    pre_positions = property(lambda self: self._pre_positions)

    @property
    def is_cached(self):
        return self._cur_positions is not None
