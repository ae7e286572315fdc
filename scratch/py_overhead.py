"""Host-side cost of LangevinIntegrator.integrate(ens, 1) with the device stubbed out (no GPU here)."""
import sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, '.')
import mdpy_b200 as md
from mdpy_b200 import _native, synthetic
from mdpy_b200.integrator import LangevinIntegrator

class StubDev:
    def __init__(self, device=None): self.n = 0
    def set_box(self, b): pass
    def set_atoms(self, q, m): self.n = len(np.asarray(q).reshape(-1))
    def __getattr__(self, name):
        if name.startswith('set_') or name in ('reset_integrator',):
            return lambda *a, **k: None
        raise AttributeError(name)
    def pinned_empty(self, shape, dtype=np.float32): return np.zeros(shape, dtype)
    def step_langevin_host(self, x, v, xo, vo, *a):
        return np.zeros(16)
_native.Device = StubDev
s = synthetic.CONFIGS['water_23k']()
ens = s.ensemble(cutoff=9.0, pme=True, grid=(64, 64, 64))
integ = LangevinIntegrator(2.0, 300, 1e-3, seed=1)
for _ in range(10): integ.integrate(ens, 1)
t0 = time.perf_counter()
for _ in range(2000): integ.integrate(ens, 1)
print('us per call', (time.perf_counter() - t0) / 2000 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): integ.integrate(ens, 1)
pr.disable(); pstats.Stats(pr).sort_stats('cumtime').print_stats(14)
