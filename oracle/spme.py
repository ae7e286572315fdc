"""oracle/spme.py — float64 numpy restatement of the smooth-PME reciprocal step.

TEST INFRASTRUCTURE ONLY.  PME has no code in the reference tree (only the names at
mdpy/constraint/__init__.py:22 and forcefield/charmm_forcefield.py:23-24): PARITY UNPINNED by any
reference output.  This restates the published algorithm (Essmann et al., J. Chem. Phys. 103, 8577,
1995) exactly as SURVEY §8c specifies it, and is itself held against the converged Ewald sum of
oracle/cpu_oracle.py:ewald_exact in tests/test_oracle.py.

    u_i = (x_i / L + 1/2) n,  Q(k) = sum_i q_i prod_a M_p(u_ia - k_a)   (periodic)
    E_rec = 1/2 sum_{m != 0} G(m) |F[Q](m)|^2,
    G(m) = k_e exp(-pi^2 m~^2 / alpha^2) / (pi V m~^2) / prod_a |b_a(m_a)|^-2 ... i.e. times B(m)
    F_i = -q_i sum_k grad_i[theta_i(k)] phi(k),  phi = N ifft(G fft(Q))
"""
import numpy as np


def bspline_weights(w, order):
    """M_p(w + p - 1 - j), j = 0..p-1, and d/dw of it, for fractional offsets w in [0,1).
    Cardinal B-spline recursion M_n(u) = [u M_{n-1}(u) + (n - u) M_{n-1}(u - 1)] / (n - 1)."""
    w = np.asarray(w, dtype=np.float64)

    def M(n, u):
        if n == 2:
            return np.where((u >= 0) & (u <= 2), 1.0 - np.abs(u - 1.0), 0.0)
        return (u * M(n - 1, u) + (n - u) * M(n - 1, u - 1.0)) / (n - 1)

    j = np.arange(order)
    u = w[:, None] + (order - 1 - j)[None, :]
    theta = M(order, u)
    dtheta = M(order - 1, u) - M(order - 1, u - 1.0)
    return theta, dtheta


def bspline_moduli(n, order):
    """|b(m)|^-2 = |sum_{k=0}^{p-2} M_p(k+1) exp(2 pi i m k / n)|^2."""
    k = np.arange(order - 1)
    theta, _ = bspline_weights(np.zeros(1), order)       # theta[0, j] = M_p(p - 1 - j)
    Mk = theta[0, ::-1][1:order]                         # M_p(1), ..., M_p(p-1)
    m = np.arange(n)
    s = (Mk[None, :] * np.exp(2j * np.pi * m[:, None] * k[None, :] / n)).sum(1)
    mod = np.abs(s) ** 2
    for i in np.nonzero(mod < 1e-7)[0]:
        mod[i] = 0.5 * (mod[(i - 1) % n] + mod[(i + 1) % n])
    return mod


def influence_function(box, grid, order, alpha, k_e):
    nx, ny, nz = grid
    V = float(np.prod(box))
    mx = np.fft.fftfreq(nx, 1.0 / nx) / box[0]
    my = np.fft.fftfreq(ny, 1.0 / ny) / box[1]
    mz = np.fft.fftfreq(nz, 1.0 / nz) / box[2]
    # fftfreq puts the Nyquist index at -n/2; the sign does not matter for m^2
    m2 = mx[:, None, None] ** 2 + my[None, :, None] ** 2 + mz[None, None, :] ** 2
    bmod = (bspline_moduli(nx, order)[:, None, None] * bspline_moduli(ny, order)[None, :, None]
            * bspline_moduli(nz, order)[None, None, :])
    with np.errstate(divide='ignore', invalid='ignore'):
        G = k_e * np.exp(-np.pi ** 2 * m2 / alpha ** 2) / (np.pi * V * m2 * bmod)
    G[0, 0, 0] = 0.0
    return G


def _spline_terms(pos, box, grid, order):
    th, dth, idx = [], [], []
    for a in range(3):
        u = (pos[:, a] / box[a] + 0.5) * grid[a]
        k0 = np.floor(u)
        t, d = bspline_weights(u - k0, order)
        th.append(t); dth.append(d)
        idx.append((k0.astype(np.int64)[:, None] - order + 1 + np.arange(order)[None, :]) % grid[a])
    return th, dth, idx


def spme_reciprocal(positions, charges, box, grid, order, alpha, k_e, atoms=None, chunk=65536):
    """Returns (forces [N,3], E_rec).  positions may be anywhere (wrapped internally).
    atoms: index array — forces are gathered for these atoms only (rows of the others stay zero); the mesh
    always carries every charge.  Atoms are processed in chunks so that million-atom boxes fit in memory."""
    pos = np.asarray(positions, dtype=np.float64)
    q = np.asarray(charges, dtype=np.float64).reshape(-1)
    box = np.asarray(box, dtype=np.float64).reshape(3)
    grid = tuple(int(g) for g in grid)
    n = pos.shape[0]
    pos = pos - box * np.round(pos / box)
    nx, ny, nz = grid
    Q = np.zeros(nx * ny * nz)
    for lo in range(0, n, chunk):
        sl = slice(lo, min(n, lo + chunk))
        th, _, idx = _spline_terms(pos[sl], box, grid, order)
        flat = ((idx[0][:, :, None, None] * ny + idx[1][:, None, :, None]) * nz + idx[2][:, None, None, :])
        wgt = q[sl, None, None, None] * th[0][:, :, None, None] * th[1][:, None, :, None] * th[2][:, None, None, :]
        Q += np.bincount(flat.ravel(), weights=wgt.ravel(), minlength=Q.size)
    Q = Q.reshape(grid)
    G = influence_function(box, grid, order, alpha, k_e)
    FQ = np.fft.fftn(Q)
    e_rec = 0.5 * float((G * np.abs(FQ) ** 2).sum())
    phi = (np.real(np.fft.ifftn(G * FQ)) * Q.size).ravel()
    forces = np.zeros((n, 3))
    which = np.arange(n) if atoms is None else np.asarray(atoms, dtype=np.int64)
    scale = np.array(grid) / box
    for lo in range(0, len(which), chunk):
        sel = which[lo:lo + chunk]
        th, dth, idx = _spline_terms(pos[sel], box, grid, order)
        flat = ((idx[0][:, :, None, None] * ny + idx[1][:, None, :, None]) * nz + idx[2][:, None, None, :])
        p = phi[flat]
        fx = (dth[0][:, :, None, None] * th[1][:, None, :, None] * th[2][:, None, None, :] * p).sum((1, 2, 3))
        fy = (th[0][:, :, None, None] * dth[1][:, None, :, None] * th[2][:, None, None, :] * p).sum((1, 2, 3))
        fz = (th[0][:, :, None, None] * th[1][:, None, :, None] * dth[2][:, None, None, :] * p).sum((1, 2, 3))
        forces[sel] = -q[sel, None] * np.stack([fx, fy, fz], 1) * scale[None, :]
    return forces, e_rec


def pme_total(positions, charges, box, bonded, grid, order, alpha, rc, k_e, threads=1):
    """Full PME electrostatics in float64 (direct erfc inside rc + reciprocal + self + background +
    excluded-pair correction) — what ElectrostaticPMEConstraint evaluates.
    Returns (forces, dict of energy terms)."""
    from . import cpu_oracle as ora
    n = np.asarray(positions).shape[0]
    d = ora.nonbonded_bruteforce(positions, box, np.zeros((n, 4)), charges, bonded, -np.ones((n, 1), dtype=np.int32),
                                 rc_lj=0.0, coul_mode=1, k_e=k_e, alpha=alpha, rc_coul=rc, threads=threads)
    f_rec, e_rec = spme_reciprocal(positions, charges, box, grid, order, alpha, k_e)
    q = np.asarray(charges, dtype=np.float64).reshape(-1)
    V = float(np.prod(np.asarray(box, dtype=np.float64)))
    e_self = -k_e * alpha / np.sqrt(np.pi) * float((q ** 2).sum())
    e_bg = -k_e * np.pi * float(q.sum()) ** 2 / (2 * V * alpha ** 2)
    en = dict(direct=d['e_coul'], excl=d['e_excl'], recip=e_rec, self_bg=e_self + e_bg)
    en['total'] = sum(en.values())
    return d['f_coul'] + f_rec, en
