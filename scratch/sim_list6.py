"""8 x 4 sub-tile mapping (lane = (i % 8, j % 4); 32 sub-tile steps per 32 x 32 chunk; a step whose 32 slots are all outside
the cutoff leaves after the distance test) against the rotation ring, on the two-class (near / far) list order.
Variants of the atom order inside an i-block: as sorted by cell (ties: particle id), or re-sorted by a finer key."""
import sys, numpy as np
sys.path.insert(0, '.')
from mdpy_b200 import synthetic
from scipy.spatial import cKDTree
name = sys.argv[1]; rc = float(sys.argv[2]); skin = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
fine = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sysm = synthetic.CONFIGS[name]()
box = np.asarray(sysm.box, dtype=np.float64)
x = np.asarray(sysm.positions, dtype=np.float64)
rng = np.random.default_rng(0)
x = x + rng.normal(0, 0.5, x.shape); x -= box * np.round(x / box)
n = len(x); rho = n / box.prod(); R = rc + skin
cyz = np.cbrt(32 / rho); target = np.array([0.5 * cyz, cyz, cyz])
nc = np.maximum(1, np.floor(box / target)).astype(int); cw = box / nc
u = (x + 0.5 * box) / cw
c = np.clip(np.floor(u).astype(int), 0, nc - 1)
key = ((c[:, 2] * nc[1] + c[:, 1]) * nc[0] + c[:, 0]).astype(np.int64)
if fine:   # inside a cell: 2 x 2 sub-cells in y, z (x is already half width), then x
    f = np.clip(((u - c) * 2).astype(int), 0, 1)
    key = key * 4 + f[:, 2] * 2 + f[:, 1]
order = np.argsort(key, kind='stable')
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
def sub84(tile):
    t4 = tile.reshape(4, 8, 8, 4)            # [ic, il, jg, jl]
    hit = t4.any(axis=(1, 3))               # [ic, jg]
    jg_hit = hit.any(0).sum()
    return 32 * 13 + hit.sum() * 66 + jg_hit * 22 + 30, hit.sum(), t4.sum()
res = dict(ring=0, sub84=0); nh = nl = 0
for b in rng.choice(nb - 2, 100, replace=False):
    ii = np.arange(b * 32, b * 32 + 32)
    cand = set()
    for lst in t.query_ball_point(xs[ii] + 0.5 * box, R): cand.update(lst)
    js = np.array(sorted(j for j in cand if j >= (b + 1) * 32), dtype=int)
    dd = d2(xs[ii], xs[js]); m = dd <= rc * rc
    near = dd.min(0) <= rc * rc
    for part in (m[:, near], m[:, ~near]):
        for tile in chunks(part):
            res['ring'] += ring(tile)
            c8, h, l = sub84(tile); res['sub84'] += c8; nh += h; nl += l
print(name, rc, R, 'fine' if fine else 'cell order')
for k, v in res.items(): print('%-8s %10d  speed-up over the ring %.3f' % (k, v, res['ring'] / v))
print('lanes active in executed sub-tile steps: %.1f of 32' % (nl / nh))
