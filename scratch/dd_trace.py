#!/usr/bin/env python
"""Per-phase wall times of the domain-decomposed step (mdk_dd_trace), one line per rank.  Run under torchrun:
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scratch/dd_trace.py [--config protein_1m] [--steps 60]
With the trace on every phase ends in a stream synchronisation, so the sum is larger than an untraced step; the
SHARES show where a step spends its time and which rank waits for which."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument('--config', default='protein_1m')
ap.add_argument('--steps', type=int, default=60)
ap.add_argument('--mesh-ms', type=float, default=-1.0, help='weight the domains: cost of the mesh chain on the mesh rank (ms); < 0 = equal domains')
ap.add_argument('--pme-pair-blocks', type=int, default=0, help='persistent pair blocks per SM on the mesh rank (0 = default)')
a = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
import torch, torch.distributed as dist
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
os.environ['MDPY_B200_DEVICE'] = str(local)
import bench
from mdpy_b200 import multigpu

class A: pass
args = A(); args.relax = 0.3; args.no_graph = False
w = bench.Workload(a.config, args)
weights = multigpu.domain_weights(world, 3.3, a.mesh_ms) if a.mesh_ms >= 0 else None
grid = multigpu.attach(w.ctx, dist, rank, world, weights=weights)
if a.pme_pair_blocks and rank == world - 1:
    w.dev.set_option('pair_blocks_per_sm', a.pme_pair_blocks)
w.integ.integrate(w.ens, 5)
w.step(20)
# untraced timing first
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
w.step(a.steps)
torch.cuda.synchronize(); dist.barrier(); untraced = (time.perf_counter() - t0) * 1e3 / a.steps
w.dev.dd_trace(True)
w.step(a.steps)
t = w.dev.dd_trace(False)
steps = max(t.pop('steps'), 1.0)
line = dict(rank=rank, grid=list(grid), weights=None if weights is None else np.round(weights, 3).tolist(), untraced_ms_per_step=round(untraced, 4), traced_ms_per_step=round(sum(t.values()) / steps, 4),
            phases_ms_per_step={k: round(v / steps, 4) for k, v in t.items()}, stats=w.dev.dd_stats())
for r in range(world):
    dist.barrier()
    if r == rank:
        print(json.dumps(line), flush=True)
dist.destroy_process_group()
