"""Warp-step counts: 32x32 rotation tiles (current) vs 8-atom i-clusters x 4 groups of 8 j (8 rotation steps)."""
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
tot = dict(cur_steps=0, cur_nonempty=0, cur_lanes=0, c8_steps=0, c8_nonempty=0, c8_lanes=0, pairs=0)
def inr(xi, xj):
    d = xj[None] - xi[:, None]; d -= box * np.round(d / box)
    return (d ** 2).sum(-1) <= rc * rc
for b in rng.choice(nb - 2, 120, replace=False):
    ii = np.arange(b * 32, b * 32 + 32); xi = xs[ii]
    # current: j > block, within R of any of the 32
    cand = set()
    for lst in t.query_ball_point(xi + 0.5 * box, R): cand.update(lst)
    js = np.array(sorted(j for j in cand if j >= (b + 1) * 32), dtype=int)
    pad = (-len(js)) % 32
    m = inr(xi, xs[js]); m = np.concatenate([m, np.zeros((32, pad), bool)], 1)
    tot['pairs'] += m.sum()
    for ch in range(m.shape[1] // 32):
        tile = m[:, ch * 32:(ch + 1) * 32]
        for k in range(32):
            lanes = tile[np.arange(32), (np.arange(32) + k) % 32].sum()
            tot['cur_steps'] += 1; tot['cur_nonempty'] += lanes > 0; tot['cur_lanes'] += lanes
    # cluster-8: for each of the 4 clusters: j >= cluster start + 32 within R of any of its 8 (the first 32 following atoms form the diagonal chunk: ignore here and in 'cur' alike)
    for a in range(4):
        i8 = ii[a * 8:(a + 1) * 8]; x8 = xs[i8]
        cand = set()
        for lst in t.query_ball_point(x8 + 0.5 * box, R): cand.update(lst)
        j8 = np.array(sorted(j for j in cand if j >= (b + 1) * 32), dtype=int)
        pad = (-len(j8)) % 32
        m8 = inr(x8, xs[j8]); m8 = np.concatenate([m8, np.zeros((8, pad), bool)], 1)
        for ch in range(m8.shape[1] // 32):
            tile = m8[:, ch * 32:(ch + 1) * 32].reshape(8, 4, 8)   # [ii, g, slot]
            for r in range(8):
                lanes = tile[np.arange(8)[:, None], np.arange(4)[None, :], (np.arange(8)[:, None] + r) % 8].sum()
                tot['c8_steps'] += 1; tot['c8_nonempty'] += lanes > 0; tot['c8_lanes'] += lanes
print(name, 'rc', rc, 'R', R)
print('current : warp-steps %d nonempty %.3f  lanes/nonempty-step %.2f' % (tot['cur_steps'], tot['cur_nonempty'] / tot['cur_steps'], tot['cur_lanes'] / tot['cur_nonempty']))
print('cluster8: warp-steps %d nonempty %.3f  lanes/nonempty-step %.2f' % (tot['c8_steps'], tot['c8_nonempty'] / tot['c8_steps'], tot['c8_lanes'] / tot['c8_nonempty']))
cheap, exp = 13, 62
cur = tot['cur_steps'] * cheap + tot['cur_nonempty'] * exp
c8 = tot['c8_steps'] * cheap + tot['c8_nonempty'] * exp
print('instruction model (13 cheap + 62 expensive per non-empty step): current %d  cluster8 %d  ratio %.3f' % (cur, c8, cur / c8))
