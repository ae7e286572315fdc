"""Multi-GPU host layer: one process per GPU (torchrun), NCCL over NVLink/NVSwitch.

The reference is single-process, single-device; this is new design:

* positions are replicated; every rank integrates all atoms (O(N), deterministic, so all ranks
  stay bit-identical);
* the i-blocks of the tile list are dealt to ranks by `block % modulus in [lo, hi)` — tile order is
  spatial, so this is a fine spatial interleave that also balances the half-shell list lengths;
  each rank builds and evaluates only its own blocks' work units;
* the PME mesh runs on the last rank (its pair range is narrowed by the weights so that it still
  finishes with the others); bonded and excluded-pair terms are dealt evenly, in contiguous ranges;
* every force evaluation ends with ONE `ncclAllReduce(sum)` of the int64 fixed-point force
  accumulator, issued by libmdpyb200 on its own stream (mdk_comm.cu).  Integer sums are exact and
  order independent: the N-GPU forces equal the 1-GPU forces bit for bit.

torch.distributed is used for the rendezvous only (broadcast of the 128-byte NCCL unique id).
"""
import numpy as np

RESIDUES_PER_RANK = 32   # interleave period = 32 * world i-blocks: fine enough to balance the half-shell lists


def shard_ranges(weights, modulus=None):
    """Split residues [0, modulus) into len(weights) consecutive ranges with widths proportional to
    weights (largest-remainder rounding, every positive weight gets at least one residue).
    Returns a list of (lo, hi)."""
    w = np.asarray(weights, dtype=np.float64)
    if modulus is None:
        modulus = RESIDUES_PER_RANK * len(w)
    if (w < 0).any() or w.sum() <= 0:
        raise ValueError('weights must be non-negative with a positive sum')
    if len(w) > modulus:
        raise ValueError('more ranks than residues')
    ideal = w / w.sum() * modulus
    width = np.floor(ideal).astype(int)
    width[(w > 0) & (width == 0)] = 1
    # distribute what is left (or take back what the minimum-one rule overspent) by largest remainder
    while width.sum() < modulus:
        k = int(np.argmax(ideal - width)); width[k] += 1
    while width.sum() > modulus:
        k = int(np.argmax(np.where(width > 1, width - ideal, -np.inf))); width[k] -= 1
    hi = np.cumsum(width)
    lo = hi - width
    return [(int(a), int(b)) for a, b in zip(lo, hi)]


def role_weights(nranks, pair_ms, pme_ms, bonded_ms=0.0):
    """Pair-work weights that equalise rank times when the last rank also runs the PME mesh
    (pme_ms) and rank 0 the bonded / excluded-pair terms (bonded_ms); pair_ms is the single-GPU
    pair-kernel time.  Solves  pair_ms * x_r + extra_r = T  with  sum x_r = 1."""
    extra = np.zeros(nranks)
    extra[-1] += pme_ms
    extra[0] += bonded_ms
    if nranks == 1:
        return np.ones(1)
    active = np.ones(nranks, dtype=bool)
    for _ in range(nranks):
        T = (pair_ms + extra[active].sum()) / active.sum()
        x = np.where(active, (T - extra) / pair_ms, 0.0)
        if (x[active] >= 0).all():
            break
        active &= x > 0       # a rank whose extra work already exceeds T gets no pair work
    x = np.clip(x, 0, None)
    if x.sum() <= 0:
        x = np.ones(nranks)
    return x / x.sum()


def broadcast_unique_id(dist, dev, rank):
    """Rank 0 asks NCCL (through libmdpyb200) for a unique id and ships it with torch.distributed."""
    import torch
    if dist.get_backend() == 'nccl':
        buf = torch.zeros(128, dtype=torch.uint8, device='cuda')
    else:
        buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = dev.comm_unique_id()
        buf.copy_(torch.frombuffer(bytearray(uid), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def attach(ctx, dist, rank, world, weights=None):
    """Join this rank's device context to the job: communicator + i-block shard."""
    import os
    import sys
    dev = ctx.dev
    uid = broadcast_unique_id(dist, dev, rank)
    # NCCL may print its version banner on stdout at the first communicator; keep stdout clean for
    # callers that emit machine-readable output there (bench.py's single JSON line)
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        dev.comm_init(rank, world, uid)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    # The shard ranges MUST be identical on every rank (a block owned twice is counted twice, a block owned
    # by nobody is lost): weights derived from per-rank timings differ in their last digits, so rank 0's
    # copy is the one everybody uses.
    w = np.ones(world) if weights is None else np.asarray(weights, dtype=np.float64)
    set_weights(ctx, rank, world, broadcast_array(dist, w, rank))


def broadcast_array(dist, values, rank, src=0):
    """float64 array from rank `src` to everybody (torch.distributed, either backend)."""
    import torch
    t = torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64).copy())
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def set_weights(ctx, rank, world, weights):
    modulus = RESIDUES_PER_RANK * world
    lo, hi = shard_ranges(weights, modulus)[rank]
    ctx.dev.set_shard(lo, hi, modulus)
    ctx.shard = (lo, hi, modulus)
