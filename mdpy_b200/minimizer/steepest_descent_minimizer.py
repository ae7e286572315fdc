"""Steepest descent — drop-in for mdpy/minimizer/steepest_descent_minimizer.py:15-53, device resident.

Same constructor `(alpha=0.01, output_unit, output_unit_label, is_verbose, log_freq)` and the same iteration:
every atom moves `alpha` along its own unit force vector, `x_i += alpha F_i / |F_i|` (:38-41), until the relative
change of the potential energy between two iterations is below `energy_tolerance` (:44,48) or `max_iterations`
is reached.  The reference runs one host numpy update, one wrap, one cell-list rebuild and one Ensemble.update
per iteration; here the whole loop is one call (mdk_minimize_sd): forces, move, wrap and list upkeep stay on the
GPU and only the energy of each iteration crosses to the host for the stopping test.
"""
from .. import _native
from ..environment import env
from ..unit import kilojoule_permol
from .minimizer import Minimizer


class SteepestDescentMinimizer(Minimizer):
    def __init__(self, alpha=0.01, output_unit=kilojoule_permol, output_unit_label='kj/mol', is_verbose=False, log_freq=5):
        super().__init__(output_unit=output_unit, output_unit_label=output_unit_label, is_verbose=is_verbose, log_freq=log_freq)
        self._alpha = alpha
        self.num_iterations = 0

    def minimize(self, ensemble, energy_tolerance=0.001, max_iterations: int = 1000):
        ctx = _native.context_of(ensemble)
        terms = 0
        for c in ensemble.constraints:
            if not getattr(c, 'is_native', False):
                raise TypeError('%s is not a native constraint: the device minimizer needs every force term on the GPU' % c)
            c._configure()
            terms |= c.terms
        ctx._pos_rev = None
        ctx.sync_positions()
        ctx.dev.reset_integrator()
        ctx.integrator_owner = None
        print('Start energy minimization with steepest decent method')
        it, (e_first, e_prev, e_last), e = ctx.dev.minimize_sd(float(self._alpha), float(energy_tolerance), int(max_iterations), terms)
        print('Initial potential energy: %s' % self._energy2str(e_first))
        self.num_iterations = it
        state = ensemble.state
        state._positions = ctx.dev.download_positions().astype(env.NUMPY_FLOAT)
        if hasattr(state, 'revision'):
            state.revision += 1
        ctx.mark_positions_current()
        pot = 0.0
        for c in ensemble.constraints:
            c._potential_energy = c._energy_from(e)
            pot += c._potential_energy
        ensemble._potential_energy = pot
        error = abs((e_last - e_prev) / e_prev) if e_prev != 0 else 0.0
        if it < max_iterations or error < energy_tolerance:
            print('Penultimate potential energy %s' % self._energy2str(e_prev))
            print('Final potential energy %s' % self._energy2str(e_last))
            print('Energy error: %e < %e' % (error, energy_tolerance))
        else:
            print('Final potential energy: %s' % self._energy2str(e_last))
