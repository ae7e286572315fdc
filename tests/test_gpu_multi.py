"""Multi-GPU parity (needs >= 2 B200s; skipped on a single-GPU box): the sharded, all-reduced
forces of an N-rank job equal the single-GPU forces bit for bit, and a sharded Langevin run
reproduces the single-GPU trajectory — bit for bit with the canonical per-pair minimum image (forces
are then independent of when the tile list was rebuilt), to rounding with the hoisted one (whose last
bits depend on the i-block frames, i.e. on the rebuild history, which differs between the two paths)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, numpy as np
sys.path.insert(0, os.environ['MDPY_ROOT'])
import torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
os.environ['MDPY_B200_DEVICE'] = str(local)
from mdpy_b200 import synthetic, _native, multigpu
from mdpy_b200.integrator import LangevinIntegrator
s = synthetic.solvated_protein_box(20002, (60.0, 60.0, 60.0), protein_fraction=0.1, seed=3)
def run(shard, canonical, pme_ms=40.0):
    ens = s.ensemble(cutoff=10.0, switch=8.0, pme=True, grid=(60, 60, 60))
    ctx = _native.context_of(ens)
    ctx.dev.set_option('canonical_min_image', 1 if canonical else 0)
    if shard:
        multigpu.attach(ctx, dist, rank, world, multigpu.role_weights(world, 100.0, pme_ms, 10.0))
    ens.update()
    f0, e0 = ens.forces.copy(), ens.potential_energy
    LangevinIntegrator(1.0, 300, 0.01, seed=5).integrate(ens, 25)
    return f0, e0, ens.state.positions.copy(), ens.total_energy
a = run(False, True)
b = run(True, True)
ok = (np.array_equal(a[0], b[0]), a[1] == b[1], np.array_equal(a[2], b[2]), a[3] == b[3])
c = run(False, False)
d = run(True, False)
dx = c[2] - d[2]; dx -= 60.0 * np.round(dx / 60.0)
ok2 = (np.array_equal(c[0], d[0]), c[1] == d[1], bool(np.abs(dx).max() < 1e-3), bool(abs(c[3] - d[3]) < 1e-5 * abs(c[3])))
# a PME rank so loaded that it gets no pair work at all (what an 8-GPU run of the 1M box does)
z = run(True, True, pme_ms=4000.0)
ok3 = (np.array_equal(a[0], z[0]), a[1] == z[1], np.array_equal(a[2], z[2]), a[3] == z[3])
if rank == 0:
    print('MULTI', ok, float(np.abs(a[0] - b[0]).max()), a[1], b[1])
    print('HOIST', ok2, float(np.abs(dx).max()), c[3], d[3])
    print('ZEROW', ok3, float(np.abs(a[2] - z[2]).max()), a[3], z[3])
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_forces_and_trajectory_equal_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MDPY_ROOT=ROOT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(min(n, 4)),
                          '--master-addr', '127.0.0.1', '--master-port', '29631', str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('MULTI')][0]
    assert '(True, True, True, True)' in line, line
    line = [l for l in out.stdout.splitlines() if l.startswith('HOIST')][0]
    assert '(True, True, True, True)' in line, line
    line = [l for l in out.stdout.splitlines() if l.startswith('ZEROW')][0]
    assert '(True, True, True, True)' in line, line
