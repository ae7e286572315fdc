"""Estimate pair-slot efficiency of list geometries on the synthetic boxes (numpy, CPU)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from mdpy_b200 import synthetic
from scipy.spatial import cKDTree

name = sys.argv[1] if len(sys.argv) > 1 else 'protein_92k'
rc = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
skin = 2.0
sysm = synthetic.CONFIGS[name]()
box = np.asarray(sysm.box, dtype=np.float64)
x = np.asarray(sysm.positions, dtype=np.float64)
rng = np.random.default_rng(0)
x = x + rng.normal(0, 0.5, x.shape)   # thermal disorder instead of the lattice start
x -= box * np.round(x / box)
n = len(x)
rho = n / box.prod()
R = rc + skin
print(name, n, 'rho', rho)

def order_for(sub):
    cyz = np.cbrt(32.0 / rho)
    target = np.array([0.5 * cyz, cyz, cyz])
    nc = np.maximum(1, np.floor(box / target)).astype(int)
    cw = box / nc
    c = np.clip(np.floor((x + 0.5 * box) / cw).astype(int), 0, nc - 1)
    key = (c[:, 2] * nc[1] + c[:, 1]) * nc[0] + c[:, 0]
    if sub:
        f = (x + 0.5 * box) / cw - c     # fractional position in cell
        sy = (f[:, 1] >= 0.5).astype(int); sz = (f[:, 2] >= 0.5).astype(int)
        if sub == 1:
            key = key * 4 + sz * 2 + sy
        elif sub == 2:   # pair two x-cells: cell pair (even x, odd x), then subcube order
            cx2 = c[:, 0] // 2
            nx2 = (nc[0] + 1) // 2
            k2 = (c[:, 2] * nc[1] + c[:, 1]) * nx2 + cx2
            key = k2 * 8 + sz * 4 + sy * 2 + (c[:, 0] & 1)
    return np.argsort(key, kind='stable')

tree_pairs = None
def count_pairs():
    t = cKDTree(x + 0.5 * box, boxsize=box)
    return t.count_neighbors(t, rc) - n

npairs = count_pairs() // 2
print('pairs in rc', npairs, 'per atom', npairs / n)

def mi(d):
    return d - box * np.round(d / box)

def simulate(order, nblk_sample=300, isub=8, jsub=4):
    xs = x[order]
    nb = (n + 31) // 32
    t = cKDTree(xs + 0.5 * box, boxsize=box)
    blocks = rng.choice(nb - 1, nblk_sample, replace=False)
    slots_tile = 0; slots_sub = 0; slots_sub_skin = 0; pairs = 0; slots_sub_exact = 0
    for b in blocks:
        ii = np.arange(b * 32, b * 32 + 32)
        xi = xs[ii]
        # j candidates: tile index > block end, within R of any i atom
        cand = set()
        for lst in t.query_ball_point(xi + 0.5 * box, R):
            cand.update(lst)
        js = np.array(sorted(j for j in cand if j >= (b + 1) * 32))
        # reference frame: block centre
        c0 = xi[0]
        xi_r = mi(xi - c0)
        xj_r = mi(xs[js] - c0)
        d2 = ((xi_r[:, None, :] - xj_r[None, :, :]) ** 2).sum(-1)
        pairs += (d2 <= rc * rc).sum()
        nj = len(js)
        njp = ((nj + 31) // 32) * 32
        slots_tile += 32 * njp
        # sub-tiles
        for a in range(0, 32, isub):
            bi_lo = xi_r[a:a + isub].min(0); bi_hi = xi_r[a:a + isub].max(0)
            for q in range(0, nj, jsub):
                xj = xj_r[q:q + jsub]
                bj_lo = xj.min(0); bj_hi = xj.max(0)
                gap = np.maximum(0, np.maximum(bi_lo - bj_hi, bj_lo - bi_hi))
                g2 = (gap ** 2).sum()
                if g2 <= rc * rc: slots_sub += isub * jsub
                if g2 <= R * R: slots_sub_skin += isub * jsub
                if (d2[a:a + isub, q:q + jsub] <= rc * rc).any(): slots_sub_exact += isub * jsub
        # diagonal tile: 32*32/2 useful-ish; count full tile
        dd = ((xi_r[:, None, :] - xi_r[None, :, :]) ** 2).sum(-1)
        pairs += ((dd <= rc * rc).sum() - 32) // 2
        slots_tile += 1024; slots_sub += 1024; slots_sub_skin += 1024; slots_sub_exact += 1024
    return pairs / slots_tile, pairs / slots_sub, pairs / slots_sub_skin, pairs / slots_sub_exact

for sub in (0, 1, 2):
    o = order_for(sub)
    for (isub, jsub) in ((8, 4), (4, 8), (8, 8), (16, 2)):
        t0 = time.time()
        e = simulate(o, 150, isub, jsub)
        print('sub-sort %d  %dx%d: tile eff %.3f | bbox-pruned@rc %.3f | bbox-pruned@R %.3f | exact-any@rc %.3f  (%.0fs)' % (sub, isub, jsub, *e, time.time() - t0))
