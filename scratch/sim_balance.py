"""Work per i-block as a function of its position in tile order (half shell by tile index): how evenly would
contiguous spatial domains be loaded?  numpy model on the 92k box."""
import sys, numpy as np
sys.path.insert(0, '.')
from mdpy_b200 import synthetic
from scipy.spatial import cKDTree
s = synthetic.CONFIGS['protein_92k'](); box = np.asarray(s.box, float)
x = np.asarray(s.positions, float); rng = np.random.default_rng(0)
x = x + rng.normal(0, 0.5, x.shape); x -= box * np.round(x / box)
n = len(x); rho = n / box.prod(); R = 14.0
cyz = np.cbrt(32 / rho); nc = np.maximum(1, np.floor(box / np.array([0.5 * cyz, cyz, cyz]))).astype(int); cw = box / nc
c = np.clip(np.floor((x + 0.5 * box) / cw).astype(int), 0, nc - 1)
order = np.argsort((c[:, 2] * nc[1] + c[:, 1]) * nc[0] + c[:, 0], kind='stable'); xs = x[order]
t = cKDTree(xs + 0.5 * box, boxsize=box); nb = n // 32
blocks = np.arange(0, nb, 12); work = []
for b in blocks:
    cand = set()
    for lst in t.query_ball_point(xs[b * 32:b * 32 + 32] + 0.5 * box, R): cand.update(lst)
    work.append(sum(1 for j in cand if j >= (b + 1) * 32))
work = np.array(work, float); frac = blocks / nb
for lo in np.arange(0, 1, 0.125):
    m = (frac >= lo) & (frac < lo + 0.125)
    print('tile-order octile %.3f-%.3f (z slab %4.1f..%4.1f A): mean j per block %.0f' % (lo, lo + 0.125, lo * box[2], (lo + 0.125) * box[2], work[m].mean()))
print('mean %.0f  max/mean over octiles %.2f' % (work.mean(), max(work[(frac >= lo) & (frac < lo + 0.125)].mean() for lo in np.arange(0, 1, 0.125)) / work.mean()))
