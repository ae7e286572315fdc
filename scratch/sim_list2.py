"""Slot efficiency vs i-cluster size (exact per-cluster j filter), j padded to 32 per cluster."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from mdpy_b200 import synthetic
from scipy.spatial import cKDTree
name = sys.argv[1]; rc = float(sys.argv[2]); skin = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
sysm = synthetic.CONFIGS[name]()
box = np.asarray(sysm.box, dtype=np.float64)
x = np.asarray(sysm.positions, dtype=np.float64)
rng = np.random.default_rng(0)
x = x + rng.normal(0, 0.5, x.shape); x -= box * np.round(x / box)
n = len(x); rho = n / box.prod(); R = rc + skin
def order_for(atoms_per_cell_pair, mode):
    cyz = np.cbrt(atoms_per_cell_pair / rho)
    target = np.array([0.5 * cyz, cyz, cyz])
    nc = np.maximum(1, np.floor(box / target)).astype(int); cw = box / nc
    c = np.clip(np.floor((x + 0.5 * box) / cw).astype(int), 0, nc - 1)
    key = (c[:, 2] * nc[1] + c[:, 1]) * nc[0] + c[:, 0]
    if mode == 1:
        f = (x + 0.5 * box) / cw - c
        sy = (f[:, 1] >= 0.5).astype(int); sz = (f[:, 2] >= 0.5).astype(int)
        key = key * 4 + sz * 2 + sy
    return np.argsort(key, kind='stable')
t0 = cKDTree(x + 0.5 * box, boxsize=box)
npairs = (t0.count_neighbors(t0, rc) - n) // 2
print(name, n, 'pairs/atom', npairs / n, 'rc', rc, 'R', R)
def sim(order, isz, nsample=400, pad=32):
    xs = x[order]; t = cKDTree(xs + 0.5 * box, boxsize=box)
    nb = n // isz
    slots = 0; pairs = 0; chunks = 0
    for b in rng.choice(nb - 1, nsample, replace=False):
        ii = np.arange(b * isz, (b + 1) * isz); xi = xs[ii]
        cand = set()
        for lst in t.query_ball_point(xi + 0.5 * box, R): cand.update(lst)
        js = np.array(sorted(j for j in cand if j >= (b + 1) * isz), dtype=int)
        d = xs[js][None] - xi[:, None]; d -= box * np.round(d / box)
        pairs += ((d ** 2).sum(-1) <= rc * rc).sum()
        slots += isz * ((len(js) + pad - 1) // pad) * pad
        chunks += (len(js) + pad - 1) // pad
        dd = xi[None] - xi[:, None]; dd -= box * np.round(dd / box)
        pairs += (((dd ** 2).sum(-1) <= rc * rc).sum() - isz) // 2
        slots += isz * pad; chunks += 1    # diagonal tile, padded
    return pairs / slots, chunks / nsample * nb
for apc, mode in ((32, 0), (32, 1), (16, 0), (8, 0), (64, 1)):
    o = order_for(apc, mode)
    print('cell pair holds %d atoms, subsort %d:' % (apc, mode), end=' ')
    for isz in (4, 8, 16, 32):
        e, ch = sim(o, isz)
        print('i=%d eff %.3f (%.0fk chunks)' % (isz, e, ch / 1e3), end=' | ')
    print()
