"""oracle/bonded.py — float64 numpy restatement of the reference's CHARMM bonded kernels.

TEST INFRASTRUCTURE ONLY (only tests/ imports it).  Vectorised over the terms, formula by formula as the
reference's njit loops compute them:
    bonds      mdpy/constraint/charmm_bond_constraint.py:53-73      E = k (r - r0)^2
    angles     mdpy/constraint/charmm_angle_constraint.py:55-96     E = k (th - th0)^2 + ku (r13 - u0)^2
    impropers  mdpy/constraint/charmm_improper_constraint.py:57-94  E = k (psi - psi0)^2
    dihedrals  mdpy/constraint/charmm_dihedral_constraint.py:59-95  E = k (1 + cos(n phi - delta))  (energy only:
               the reference's force there is not the gradient of this energy, DESIGN Q12)
    torsion angle: mdpy/utils/geometry.py:84-96 (atan2 convention); minimum image: utils/pbc.py:38-44.
Pinned against the unmodified reference's outputs on the example system (tests/golden/config1_f64.npz,
tests/test_oracle.py).  Its purpose: the same terms evaluated at the float32-rounded positions the device holds,
so that the GPU comparison is on identical inputs (the golden run used the PDB's decimal coordinates in float64).
"""
import numpy as np


def _mi(d, box):
    return d - box * np.round(d / box)


def _unit(v):
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def _scatter(n, ids, f):
    out = np.zeros((n, 3))
    for k in range(3):
        out[:, k] = np.bincount(ids, weights=f[:, k], minlength=n)
    return out


def bonds(pos, box, idx, par):
    pos = np.asarray(pos, dtype=np.float64); box = np.asarray(box, dtype=np.float64)
    i1, i2 = idx[:, 0], idx[:, 1]
    k, r0 = par[:, 0], par[:, 1]
    vec = _mi(pos[i2] - pos[i1], box)
    r = np.linalg.norm(vec, axis=1)
    f = (2 * k * (r - r0))[:, None] * vec / r[:, None]
    n = pos.shape[0]
    return _scatter(n, i1, f) - _scatter(n, i2, f), float((k * (r - r0) ** 2).sum())


def angles(pos, box, idx, par):
    pos = np.asarray(pos, dtype=np.float64); box = np.asarray(box, dtype=np.float64)
    i1, i2, i3 = idx[:, 0], idx[:, 1], idx[:, 2]
    k, th0, ku, u0 = par[:, 0], par[:, 1], par[:, 2], par[:, 3]
    r21 = _mi(pos[i1] - pos[i2], box); l21 = np.linalg.norm(r21, axis=1)
    r23 = _mi(pos[i3] - pos[i2], box); l23 = np.linalg.norm(r23, axis=1)
    th = np.arccos((r21 * r23).sum(1) / (l21 * l23))
    fv = -2 * k * (th - th0)
    nrm = np.cross(r21, r23)
    v1 = _unit(np.cross(r21, nrm)) / l21[:, None]
    v3 = _unit(np.cross(-r23, nrm)) / l23[:, None]
    n = pos.shape[0]
    f = _scatter(n, i1, fv[:, None] * v1) - _scatter(n, i2, fv[:, None] * (v1 + v3)) + _scatter(n, i3, fv[:, None] * v3)
    e = float((k * (th - th0) ** 2).sum())
    r13 = _mi(pos[i3] - pos[i1], box); l13 = np.linalg.norm(r13, axis=1)
    fu = (2 * ku * (l13 - u0))[:, None] * r13 / l13[:, None]
    f += _scatter(n, i1, fu) - _scatter(n, i3, fu)
    e += float((ku * (l13 - u0) ** 2).sum())
    return f, e


def torsion_angle(pos, box, idx):
    p = [np.asarray(pos, dtype=np.float64)[idx[:, a]] for a in range(4)]
    r1, r2, r3 = _mi(p[1] - p[0], box), _mi(p[2] - p[1], box), _mi(p[3] - p[2], box)
    n1, n2 = np.cross(r1, r2), np.cross(r2, r3)
    x = (np.linalg.norm(r2, axis=1)[:, None] * r1 * n2).sum(1)
    return np.arctan2(x, (n1 * n2).sum(1)), (r1, r2, r3)


def impropers(pos, box, idx, par):
    box = np.asarray(box, dtype=np.float64)
    k, psi0 = par[:, 0], par[:, 1]
    psi, (vab, vbc, vcd) = torsion_angle(pos, box, idx)
    fv = -2 * k * (psi - psi0)
    lab, lbc, lcd = (np.linalg.norm(v, axis=1) for v in (vab, vbc, vcd))
    voc, loc = vbc / 2, lbc / 2
    th_abc = np.arccos((-vab * vbc).sum(1) / (lab * lbc))
    th_bcd = np.arccos((-vbc * vcd).sum(1) / (lbc * lcd))
    fa = (fv / (lab * np.sin(th_abc)))[:, None] * _unit(np.cross(-vab, vbc))
    fd = (fv / (lcd * np.sin(th_bcd)))[:, None] * _unit(np.cross(vcd, -vbc))
    fc = np.cross(-(np.cross(voc, fd) + np.cross(vcd, fd) / 2 + np.cross(-vab, fa) / 2), voc) / (loc ** 2)[:, None]
    fb = -(fa + fc + fd)
    n = np.asarray(pos).shape[0]
    f = sum(_scatter(n, idx[:, a], ff) for a, ff in enumerate((fa, fb, fc, fd)))
    return f, float((k * (psi - psi0) ** 2).sum())


def dihedral_energy(pos, box, idx, par):
    phi, _ = torsion_angle(pos, np.asarray(box, dtype=np.float64), idx)
    return float((par[:, 0] * (1 + np.cos(par[:, 1] * phi - par[:, 2]))).sum())
