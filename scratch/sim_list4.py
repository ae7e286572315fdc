"""Ring-kernel cost model when the j-atoms of an i-block's list are ORDERED before they are cut into chunks of 32:
(a) cell-row order (current), (b) two classes: j with >= 1 i-atom within rc + d_near first, the rest ("far": skin shell only)
last, (c) fully sorted by minimum distance to the block's atoms.  A rotation step whose 32 slots are all outside rc
costs 14 warp instructions, any other step 80 (measured from the SASS of k_pair<1,1,1,1,1,1,0>)."""
import sys, numpy as np
sys.path.insert(0, '.')
from mdpy_b200 import synthetic
from scipy.spatial import cKDTree
name = sys.argv[1]; rc = float(sys.argv[2]); skin = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
drift = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
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
# positions the kernel sees some steps after the build
xk = xs + rng.normal(0, drift / np.sqrt(3), xs.shape) if drift > 0 else xs
nb = n // 32
CHEAP, FULL = 14, 80
def d2(xi, xj):
    d = xj[None] - xi[:, None]; d -= box * np.round(d / box)
    return (d ** 2).sum(-1)
def cost(m):
    pad = (-m.shape[1]) % 32
    m = np.concatenate([m, np.zeros((32, pad), bool)], 1)
    steps = nonempty = 0
    ar = np.arange(32)
    for ch in range(m.shape[1] // 32):
        tile = m[:, ch * 32:(ch + 1) * 32]
        for k in range(32):
            steps += 1; nonempty += tile[ar, (ar + k) % 32].any()
    return steps, nonempty
tot = {k: np.zeros(2) for k in ('row', 'two', 'two1', 'two+.3', 'two+.6', 'sorted')}
pairs = slots = zero = 0
for b in rng.choice(nb - 2, 100, replace=False):
    ii = np.arange(b * 32, b * 32 + 32)
    cand = set()
    for lst in t.query_ball_point(xs[ii] + 0.5 * box, R): cand.update(lst)
    js = np.array(sorted(j for j in cand if j >= (b + 1) * 32), dtype=int)
    dmin = np.sqrt(d2(xs[ii], xs[js]).min(0))          # at build time
    m = d2(xk[ii], xk[js]) <= rc * rc                   # what the kernel finds
    pairs += m.sum(); slots += m.size; zero += (dmin > rc).sum() / len(js)
    tot['row'] += cost(m)
    near = dmin <= rc
    tot['two'] += np.add(cost(m[:, near]), cost(m[:, ~near]))
    near1 = dmin <= rc - 1.0
    tot['two1'] += np.add(cost(m[:, near1]), cost(m[:, ~near1]))
    for mg, key in ((0.3, 'two+.3'), (0.6, 'two+.6')):
        nr = dmin <= rc + mg
        tot[key] += np.add(cost(m[:, nr]), cost(m[:, ~nr]))
    o = np.argsort(dmin, kind='stable')
    tot['sorted'] += cost(m[:, o])
print(name, 'rc', rc, 'R', R, 'drift', drift, 'density %.3f' % (pairs / slots), 'zero-hit j fraction %.3f' % (zero / 100))
base = None
for k, (s, ne) in tot.items():
    ins = s * CHEAP + ne * (FULL - CHEAP)
    base = base or ins
    print('%-7s steps %7d nonempty %.3f  instr %.0f  speed-up %.3f' % (k, s, ne / s, ins, base / ins))
