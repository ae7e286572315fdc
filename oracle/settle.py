"""oracle/settle.py — rigid three-site water constraints restated in float64 numpy.

TEST INFRASTRUCTURE ONLY.  The reference has no constraint code (only the `is_SHAKE` flag,
mdpy/forcefield/charmm_forcefield.py:24,32): PARITY UNPINNED by any reference output.  Restated from the published
algorithms: SETTLE (Miyamoto & Kollman, J. Comput. Chem. 13, 952, 1992 — the analytic solution of SHAKE for a rigid
triatomic) for the positions, RATTLE's velocity stage (Andersen, J. Comput. Phys. 52, 24, 1983) for the velocities.
Pinned in tests/test_oracle.py against an independent iterative SHAKE solver (`shake`, converged to 1e-13): both
must give the same constrained positions, which pins signs and frames of the analytic solution.
"""
import numpy as np


def settle(x_old, x_new, m_o, m_h, d_oh, d_hh):
    """x_old [W,3,3]: constrained positions (O, H1, H2) at the start of the step; x_new [W,3,3]: positions after the
    unconstrained update.  Returns the positions that satisfy |OH1| = |OH2| = d_oh, |H1H2| = d_hh and differ from
    x_new by displacements along the OLD bond directions weighted by 1/m (SHAKE's condition)."""
    b0, b1, b2 = x_old[:, 0], x_old[:, 1], x_old[:, 2]
    c0, c1, c2 = x_new[:, 0], x_new[:, 1], x_new[:, 2]
    wohh = m_o + 2 * m_h
    rc = 0.5 * d_hh
    t = np.sqrt(d_oh * d_oh - rc * rc)
    ra = 2 * m_h * t / wohh
    rb = t - ra
    com = (m_o * c0 + m_h * c1 + m_h * c2) / wohh
    xb0, xc0 = b1 - b0, b2 - b0
    xa1, xb1, xc1 = c0 - com, c1 - com, c2 - com
    zax = np.cross(xb0, xc0)
    xax = np.cross(xa1, zax)
    yax = np.cross(zax, xax)
    unit = lambda v: v / np.linalg.norm(v, axis=1, keepdims=True)
    xax, yax, zax = unit(xax), unit(yax), unit(zax)
    dot = lambda a, b: (a * b).sum(1)
    b0d = np.stack([dot(xb0, xax), dot(xb0, yax)], 1)
    c0d = np.stack([dot(xc0, xax), dot(xc0, yax)], 1)
    a1z = dot(xa1, zax)
    b1d = np.stack([dot(xb1, xax), dot(xb1, yax), dot(xb1, zax)], 1)
    c1d = np.stack([dot(xc1, xax), dot(xc1, yax), dot(xc1, zax)], 1)
    sinphi = a1z / ra
    cosphi = np.sqrt(1 - sinphi ** 2)
    sinpsi = (b1d[:, 2] - c1d[:, 2]) / (2 * rc * cosphi)
    cospsi = np.sqrt(1 - sinpsi ** 2)
    a2y = ra * cosphi
    b2x = -rc * cospsi
    t1 = -rb * cosphi
    t2 = rc * sinpsi * sinphi
    b2y, c2y = t1 - t2, t1 + t2
    alpa = b2x * (b0d[:, 0] - c0d[:, 0]) + b0d[:, 1] * b2y + c0d[:, 1] * c2y
    beta = b2x * (c0d[:, 1] - b0d[:, 1]) + b0d[:, 0] * b2y + c0d[:, 0] * c2y
    gama = b0d[:, 0] * b1d[:, 1] - b1d[:, 0] * b0d[:, 1] + c0d[:, 0] * c1d[:, 1] - c1d[:, 0] * c0d[:, 1]
    al2be2 = alpa ** 2 + beta ** 2
    sinthe = (alpa * gama - beta * np.sqrt(al2be2 - gama ** 2)) / al2be2
    costhe = np.sqrt(1 - sinthe ** 2)
    a3 = np.stack([-a2y * sinthe, a2y * costhe, a1z], 1)
    b3 = np.stack([b2x * costhe - b2y * sinthe, b2x * sinthe + b2y * costhe, b1d[:, 2]], 1)
    c3 = np.stack([-b2x * costhe - c2y * sinthe, -b2x * sinthe + c2y * costhe, c1d[:, 2]], 1)
    back = lambda v: com + v[:, 0:1] * xax + v[:, 1:2] * yax + v[:, 2:3] * zax
    return np.stack([back(a3), back(b3), back(c3)], 1)


def shake(x_old, x_new, m_o, m_h, d_oh, d_hh, tol=1e-14, max_iter=500):
    """Independent check: classic iterative SHAKE on the same three constraints."""
    x = x_new.copy()
    inv_m = np.array([1 / m_o, 1 / m_h, 1 / m_h])
    pairs = [(0, 1, d_oh), (0, 2, d_oh), (1, 2, d_hh)]
    for _ in range(max_iter):
        worst = 0.0
        for i, j, d in pairs:
            r = x[:, j] - x[:, i]
            r_old = x_old[:, j] - x_old[:, i]
            diff = (r * r).sum(1) - d * d
            worst = max(worst, np.abs(diff).max() / (d * d))
            g = diff / (2 * (inv_m[i] + inv_m[j]) * (r * r_old).sum(1))
            x[:, i] += (inv_m[i] * g)[:, None] * r_old
            x[:, j] -= (inv_m[j] * g)[:, None] * r_old
        if worst < tol:
            break
    return x


def rattle_velocities(x, v, m_o, m_h):
    """Remove the velocity components along the three bonds (d/dt of every constrained distance = 0) with impulses
    along the bonds: solves the 3 x 3 system for the multipliers."""
    inv_m = np.array([1 / m_o, 1 / m_h, 1 / m_h])
    pairs = [(0, 1), (0, 2), (1, 2)]
    W = x.shape[0]
    e = [x[:, j] - x[:, i] for i, j in pairs]
    A = np.zeros((W, 3, 3)); rhs = np.zeros((W, 3))
    for k, (i, j) in enumerate(pairs):
        rhs[:, k] = ((v[:, j] - v[:, i]) * e[k]).sum(1)
        for l, (p, q) in enumerate(pairs):
            # impulse lam_l along e_l: v_p += inv_m_p lam e_l, v_q -= inv_m_q lam e_l
            coef = np.zeros(W)
            for atom, sign in ((p, 1.0), (q, -1.0)):
                if atom == j: coef += sign * inv_m[atom] * (e[l] * e[k]).sum(1)
                if atom == i: coef -= sign * inv_m[atom] * (e[l] * e[k]).sum(1)
            A[:, k, l] = coef
    lam = np.linalg.solve(A, -rhs[:, :, None])[:, :, 0]
    out = v.copy()
    for l, (p, q) in enumerate(pairs):
        out[:, p] += (inv_m[p] * lam[:, l])[:, None] * e[l]
        out[:, q] -= (inv_m[q] * lam[:, l])[:, None] * e[l]
    return out
