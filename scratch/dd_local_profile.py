#!/usr/bin/env python
"""Kernel-level view of a decomposed step on ONE GPU: N in-process ranks (local group) of the 1M box, a few steps.
Meant to run under `ncu --metrics gpu__time_duration.sum` (launch list): the per-rank kernels of a rebuild and of a step."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mdpy_b200 import _native, multigpu
name = sys.argv[1] if len(sys.argv) > 1 else 'protein_1m'
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
cfg = bench.CONFIGS[name]
system = bench.build_system(cfg)
ens = [system.ensemble(cutoff=cfg['cutoff'], switch=cfg['switch'], pme=True, ewald_error=1e-6, grid=cfg['grid'], order=4, bonded=True)
       for _ in range(nd)]
grid = multigpu.domain_grid(nd, system.box)
group = _native.LocalGroup(ens, grid)
kT = 300 * 8.31446e-7
group.step_langevin(0.1, kT, 0.2, 1, 10)
group.step_langevin(0.5, kT, 0.05, 1, 10)
group.step_langevin(1.0, kT, 0.01, 1, steps)
print([c.dev.dd_stats() for c in group.ctxs])
