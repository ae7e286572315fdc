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
        self._cached = False
        self._ctx = None

    def erase_cache(self):
        self._cur_positions = None
        self._pre_positions = None
        self._cached = False
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

    def _bind(self, ensemble):
        """Device context with every constraint configured and the box current; a change of integrator
        object drops the device-side step caches (a new reference integrator starts uncached too).
        Returns (ctx, terms)."""
        ctx = _native.context_of(ensemble)
        terms = 0
        for c in ensemble.constraints:
            if not getattr(c, 'is_native', False):
                raise TypeError('%s is not a native constraint: the device integrators need every force '
                                'term on the GPU' % c)
            c._configure()
            terms |= c.terms
        box = ctx.check_box()
        if ctx._box_rev is None or not np.array_equal(box, ctx._box_rev):
            ctx.dev.set_box(box)
            ctx._box_rev = box
        if ctx.integrator_owner is not self:
            ctx.dev.reset_integrator()
            ctx.integrator_owner = self
            self._cur_positions = None
        self._ctx = ctx
        return ctx, terms

    @staticmethod
    def _host_f32(a):
        return np.ascontiguousarray(a, dtype=np.float32)

    def _publish_state(self, ensemble, ctx, x, v, e):
        """Install the arrays a host-state step call filled (page-locked, owned by the context and
        reused every other call) as the new State, and the energies of the last step."""
        state = ensemble.state
        native_float = np.dtype(env.NUMPY_FLOAT) == np.float32
        state._positions = x if native_float else x.astype(env.NUMPY_FLOAT)
        state._velocities = v if native_float else v.astype(env.NUMPY_FLOAT)
        if hasattr(state, 'revision'):
            state.revision += 1
        ctx.mark_positions_current()
        self._cur_positions = None
        self._cached = True
        pot = 0.0
        for c in ensemble.constraints:
            c._potential_energy = c._energy_from(e)
            pot += c._potential_energy
        ensemble._potential_energy = pot
        ensemble._kinetic_energy = float(e[_native.E_KINETIC])
        ensemble._total_energy = pot + ensemble._kinetic_energy

    def _publish(self, ensemble, ctx, terms):
        """Bring the final state and energies back to the host objects."""
        state = ensemble.state
        state._positions = ctx.dev.download_positions().astype(env.NUMPY_FLOAT)
        state._velocities = ctx.dev.download_velocities().astype(env.NUMPY_FLOAT)
        if hasattr(state, 'revision'):
            state.revision += 1
        ctx.mark_positions_current()
        self._cur_positions = None
        self._cached = True
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

    @property
    def cur_positions(self):
        """Unwrapped float64 positions of the device state (fetched on demand)."""
        if self._cur_positions is None and self._cached and self._ctx is not None:
            self._cur_positions = self._ctx.dev.download_positions(unwrapped=True)
        return self._cur_positions

    pre_positions = property(lambda self: self._pre_positions)

    @property
    def is_cached(self):
        return self._cached
