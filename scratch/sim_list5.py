"""Hybrid cost model on distance-sorted j lists: per chunk, ring (14 / 80 per rotation step) or filter-then-compute
(288 + 125 per iteration of the longest lane + 336 scatter; measured from k_pair5), whichever is cheaper; and the
grouped job-queue model (G steps, fixed 33 per step + 57 per body of 32 jobs, empty groups 16 per step)."""
import sys, numpy as np
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
cyz = np.cbrt(32 / rho); target = np.array([0.5 * cyz, cyz, cyz])
nc = np.maximum(1, np.floor(box / target)).astype(int); cw = box / nc
c = np.clip(np.floor((x + 0.5 * box) / cw).astype(int), 0, nc - 1)
order = np.argsort((c[:, 2] * nc[1] + c[:, 1]) * nc[0] + c[:, 0], kind='stable')
xs = x[order]; t = cKDTree(xs + 0.5 * box, boxsize=box)
nb = n // 32
ar = np.arange(32)
def d2(xi, xj):
    d = xj[None] - xi[:, None]; d -= box * np.round(d / box)
    return (d ** 2).sum(-1)
def chunks(m):
    pad = (-m.shape[1]) % 32
    m = np.concatenate([m, np.zeros((32, pad), bool)], 1)
    return [m[:, c * 32:(c + 1) * 32] for c in range(m.shape[1] // 32)]
def ring(tile):
    ne = sum(tile[ar, (ar + k) % 32].any() for k in range(32))
    return 32 * 14 + ne * 66
def v5(tile):
    mp = tile.sum(1).max()
    return 288 + 125 * mp + (336 if mp else 0) + 60
def queue(tile, G):
    cost = 0
    for g0 in range(0, 32, G):
        J = sum(tile[ar, (ar + k) % 32].sum() for k in range(g0, g0 + G))
        cost += G * 16 if J == 0 else G * 33 + 57 * -(-J // 32) + 8
    return cost
res = dict(ring_row=0, ring_sorted=0, hybrid=0, q2=0, q4=0, q8=0, q4row=0)
dens = []
for b in rng.choice(nb - 2, 100, replace=False):
    ii = np.arange(b * 32, b * 32 + 32)
    cand = set()
    for lst in t.query_ball_point(xs[ii] + 0.5 * box, R): cand.update(lst)
    js = np.array(sorted(j for j in cand if j >= (b + 1) * 32), dtype=int)
    dd = d2(xs[ii], xs[js]); m = dd <= rc * rc
    o = np.argsort(dd.min(0), kind='stable')
    for tile in chunks(m): res['ring_row'] += ring(tile); res['q4row'] += queue(tile, 4)
    for tile in chunks(m[:, o]):
        r, v = ring(tile), v5(tile)
        res['ring_sorted'] += r; res['hybrid'] += min(r, v); dens.append(tile.mean())
        res['q2'] += queue(tile, 2); res['q4'] += queue(tile, 4); res['q8'] += queue(tile, 8)
print(name, rc, R)
for k, v in res.items(): print('%-12s %10d  speed-up %.3f' % (k, v, res['ring_row'] / v))
print('chunk density histogram (sorted order):', np.histogram(dens, bins=[0, 1e-9, .05, .1, .2, .3, .5, .8, 1.01])[0])
