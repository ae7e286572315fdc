"""Seeded synthetic systems for the benchmark configs of BASELINE.json / SURVEY §8d
(there is no network for real structures): TIP3P water boxes and "solvated protein" boxes
built from poly-alanine helices, with CHARMM36 atom types, charges and LJ parameters
(published values of par_all36_prot.prm / toppar_water_ions.str, converted with the
reference's rules: eps -> |eps| kcal/mol in internal units, sigma = Rmin/2 * 2 * 2^(-1/6),
mdpy/io/charmm_toppar_parser.py:209-222).

Bond / angle equilibrium values of the helices are measured from the generated geometry, so
a fresh box starts free of bonded strain; waters sit on a jittered lattice with random
orientations.  Everything is numpy-vectorised: the 1.07 M-atom box builds in seconds.
"""
from dataclasses import dataclass, field

import numpy as np

from .core import Topology
from .unit import RMIN_TO_SIGMA_FACTOR

KCAL = float(np.float32(4.1840284e-4))  # kcal/mol in Da A^2/fs^2 (float32-rounded like the reference)
WATER_DENSITY = 0.0334                  # molecules / A^3


def _lj(eps_kcal, rmin_half, eps14=None, rmin14=None):
    row = [abs(eps_kcal) * KCAL, rmin_half * 2 * float(RMIN_TO_SIGMA_FACTOR)]
    if eps14 is not None:
        row += [abs(eps14) * KCAL, rmin14 * 2 * float(RMIN_TO_SIGMA_FACTOR)]
    return row


LJ_PARAMETERS = {
    'OT': _lj(-0.1521, 1.7682), 'HT': _lj(-0.046, 0.2245),
    'NH1': _lj(-0.20, 1.85, -0.20, 1.55), 'H': _lj(-0.046, 0.2245),
    'CT1': _lj(-0.032, 2.0, -0.01, 1.9), 'HB1': _lj(-0.022, 1.32),
    'CT3': _lj(-0.078, 2.04, -0.01, 1.9), 'HA3': _lj(-0.024, 1.34),
    'C': _lj(-0.11, 2.0), 'O': _lj(-0.12, 1.7, -0.12, 1.4),
}
MASS = {'OT': 15.9994, 'HT': 1.008, 'NH1': 14.007, 'H': 1.008, 'CT1': 12.011, 'HB1': 1.008, 'CT3': 12.011,
        'HA3': 1.008, 'C': 12.011, 'O': 15.9994}

# alanine residue: name, type, charge (CHARMM36 ALA, neutral)
_ALA = [('N', 'NH1', -0.47), ('HN', 'H', 0.31), ('CA', 'CT1', 0.07), ('HA', 'HB1', 0.09), ('CB', 'CT3', -0.27),
        ('HB1', 'HA3', 0.09), ('HB2', 'HA3', 0.09), ('HB3', 'HA3', 0.09), ('C', 'C', 0.51), ('O', 'O', -0.51)]


@dataclass
class SyntheticSystem:
    box: np.ndarray
    positions: np.ndarray
    types: list
    masses: np.ndarray
    charges: np.ndarray
    bonds: np.ndarray
    angles: np.ndarray
    dihedrals: np.ndarray
    impropers: np.ndarray
    bond_par: np.ndarray
    angle_par: np.ndarray
    dihedral_par: np.ndarray
    improper_par: np.ndarray
    lj_parameters: dict = field(default_factory=lambda: dict(LJ_PARAMETERS))
    name: str = 'synthetic'

    @property
    def num_particles(self):
        return self.positions.shape[0]

    def topology(self):
        return Topology.from_arrays(self.types, self.masses, self.charges, self.bonds, self.angles,
                                    self.dihedrals, self.impropers)

    def lj_table(self):
        """[N,4] eps, sigma, eps14, sigma14 — what CharmmNonbondedConstraint.bind_ensemble builds."""
        rows = {k: (v + v if len(v) == 2 else v) for k, v in self.lj_parameters.items()}
        return np.array([rows[t] for t in self.types], dtype=np.float32)

    def water_triplets(self):
        """(O, H, H) matrix ids of the TIP3P molecules (the waters are the trailing OT HT HT triplets)."""
        types = np.asarray(self.types)
        first = int(np.argmax(types == 'OT')) if (types == 'OT').any() else len(types)
        o = np.arange(first, len(types), 3)
        return np.stack([o, o + 1, o + 2], 1).astype(np.int32)

    def ensemble(self, cutoff=12.0, switch=None, pme=True, ewald_error=1e-6, grid=None, order=4, bonded=True, rigid_water=False):
        """Ensemble with the native constraints bound (needs a B200).  rigid_water: the waters are held rigid by
        SETTLE (what the reference's is_SHAKE flag asks for) and their bond / angle terms are left out."""
        from . import Ensemble
        from .constraint import (CharmmAngleConstraint, CharmmBondConstraint, CharmmDihedralConstraint,
                                 CharmmImproperConstraint, CharmmNonbondedConstraint, ElectrostaticPMEConstraint)
        ens = Ensemble(self.topology(), np.diag(self.box))
        cs = [CharmmNonbondedConstraint(self.lj_parameters, cutoff, switch_radius=switch)]
        if pme:
            cs.append(ElectrostaticPMEConstraint(cutoff, ewald_error=ewald_error, grid=grid, order=order))
        keep_b = keep_a = slice(None)
        if rigid_water:
            water_atoms = self.water_triplets().reshape(-1)
            keep_b = ~np.isin(self.bonds[:, 0], water_atoms) if len(self.bonds) else slice(None)
            keep_a = ~np.isin(self.angles[:, 1], water_atoms) if len(self.angles) else slice(None)
            # exclusions come from the full bond graph (O-H, H-H stay excluded); the force terms lose the waters' rows
            ens.topology._bonds = self.bonds[keep_b]
            ens.topology._angles = self.angles[keep_a]
        if bonded:
            if len(self.bonds[keep_b]): cs.append(CharmmBondConstraint(self.bond_par[keep_b]))
            if len(self.angles[keep_a]): cs.append(CharmmAngleConstraint(self.angle_par[keep_a]))
            if len(self.dihedrals): cs.append(CharmmDihedralConstraint(self.dihedral_par))
            if len(self.impropers): cs.append(CharmmImproperConstraint(self.improper_par))
        ens.add_constraints(*cs)
        ens.state.set_positions(self.positions.astype(np.float32))
        if rigid_water:
            from . import _native
            _native.context_of(ens).dev.set_rigid_waters(self.water_triplets(), _R_OH, 2 * _R_OH * np.sin(_ANG_HOH / 2))
        return ens


# ---------------------------------------------------------------------------------------
def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)


_R_OH, _ANG_HOH = 0.9572, np.deg2rad(104.52)
_WATER_LOCAL = np.array([[0.0, 0.0, 0.0],
                         [_R_OH * np.sin(_ANG_HOH / 2), 0.0, _R_OH * np.cos(_ANG_HOH / 2)],
                         [-_R_OH * np.sin(_ANG_HOH / 2), 0.0, _R_OH * np.cos(_ANG_HOH / 2)]])


def _lattice_sites(box, spacing):
    counts = np.maximum(1, np.floor(box / spacing).astype(int))
    axes = [(np.arange(c) + 0.5) * (box[a] / c) - box[a] / 2 for a, c in enumerate(counts)]
    g = np.stack(np.meshgrid(*axes, indexing='ij'), -1).reshape(-1, 3)
    return g


def _place_waters(rng, box, n_waters, obstacles=None, jitter=0.15):
    """n_waters oxygen sites on a lattice that avoids `obstacles` (KD-tree), random orientations."""
    free_volume = float(np.prod(box)) - (0 if obstacles is None else 13.5 * len(obstacles))
    spacing = (free_volume / n_waters) ** (1.0 / 3.0)
    for _ in range(40):
        sites = _lattice_sites(box, spacing)
        if obstacles is not None and len(obstacles):
            from scipy.spatial import cKDTree
            tree = cKDTree(obstacles + box / 2, boxsize=box)
            near = tree.query_ball_point((sites + box / 2) % box, r=2.9, return_length=True)
            sites = sites[near == 0]
        if len(sites) >= n_waters:
            break
        spacing *= 0.985
    else:
        raise RuntimeError('could not place %d waters' % n_waters)
    pick = rng.choice(len(sites), size=n_waters, replace=False)
    pick.sort()
    centers = sites[pick] + rng.uniform(-jitter, jitter, size=(n_waters, 3))
    rot = _random_rotations(rng, n_waters)
    return (centers[:, None, :] + np.einsum('nij,kj->nki', rot, _WATER_LOCAL)).reshape(-1, 3)


def _water_topology(n_waters, offset):
    o = offset + 3 * np.arange(n_waters)
    bonds = np.stack([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)], 1).reshape(-1, 2)
    angles = np.stack([o + 1, o, o + 2], 1)
    bond_par = np.tile([450.0 * KCAL, _R_OH], (len(bonds), 1))
    angle_par = np.tile([55.0 * KCAL, _ANG_HOH, 0.0, 0.0], (len(angles), 1))
    return bonds, angles, bond_par, angle_par


def water_box(n_waters=7852, seed=20260001, box=None):
    """Config 2: TIP3P box, 7852 waters = 23556 atoms, cubic L = 61.7 A."""
    rng = np.random.default_rng(seed)
    if box is None:
        box = np.full(3, (n_waters / WATER_DENSITY) ** (1.0 / 3.0))
    box = np.asarray(box, dtype=np.float64)
    pos = _place_waters(rng, box, n_waters)
    bonds, angles, bond_par, angle_par = _water_topology(n_waters, 0)
    types = ['OT', 'HT', 'HT'] * n_waters
    z4 = np.zeros((0, 4), dtype=np.int64)
    return SyntheticSystem(
        box=box, positions=pos.astype(np.float32), types=types,
        masses=np.array([MASS[t] for t in types[:3]] * n_waters, dtype=np.float32),
        charges=np.array([-0.834, 0.417, 0.417] * n_waters, dtype=np.float32),
        bonds=bonds, angles=angles, dihedrals=z4, impropers=z4, bond_par=bond_par, angle_par=angle_par,
        dihedral_par=np.zeros((0, 3)), improper_par=np.zeros((0, 2)), name='tip3p_%d' % (3 * n_waters))


# ---------------------------------------------------------------------------------------
def _place(a, b, c, bond, angle, torsion):
    """NeRF: position d with |cd| = bond, angle(b,c,d) = angle, dihedral(a,b,c,d) = torsion."""
    bc = c - b
    bc /= np.linalg.norm(bc)
    n = np.cross(b - a, bc)
    n /= np.linalg.norm(n)
    m = np.cross(n, bc)
    d2 = np.array([-bond * np.cos(angle), bond * np.sin(angle) * np.cos(torsion), bond * np.sin(angle) * np.sin(torsion)])
    return c + d2[0] * bc + d2[1] * m + d2[2] * n


def _helix_template(n_res=20):
    """One poly-alanine alpha helix: coordinates, types, charges and the bonded topology with
    equilibrium values measured from the built geometry."""
    rad = np.deg2rad
    phi, psi, omega = rad(-57.0), rad(-47.0), rad(180.0)
    xyz, names = [], []
    N = np.array([0.0, 0.0, 0.0]); CA = np.array([1.458, 0.0, 0.0])
    C = CA + 1.525 * np.array([np.cos(np.pi - rad(111.0)), np.sin(np.pi - rad(111.0)), 0.0])
    prevC = None
    for r in range(n_res):
        if r > 0:
            N = _place(pN, pCA, pC, 1.329, rad(116.2), psi)
            CA = _place(pCA, pC, N, 1.458, rad(121.7), omega)
            C = _place(pC, N, CA, 1.525, rad(111.0), phi)
        unit = lambda v: v / np.linalg.norm(v)
        if prevC is not None:   # amide H in the peptide plane, opposite the C(prev) / CA bisector
            HN = N + 1.0 * unit(unit(N - prevC) + unit(N - CA))
        else:
            HN = _place(C, CA, N, 1.0, rad(109.5), rad(180.0))
        u, v = unit(N - CA), unit(C - CA)   # tetrahedral CB / HA on CA
        bis, perp = -unit(u + v), unit(np.cross(u, v))
        th = rad(54.75)
        CB = CA + 1.53 * (np.cos(th) * bis + np.sin(th) * perp)
        HA = CA + 1.08 * (np.cos(th) * bis - np.sin(th) * perp)
        HBs = [_place(N, CA, CB, 1.09, rad(109.5), rad(t)) for t in (60.0, 180.0, -60.0)]
        nextN = _place(N, CA, C, 1.329, rad(116.2), psi)
        O = C + 1.231 * unit(unit(C - CA) + unit(C - nextN))   # carbonyl O in the peptide plane
        xyz += [N, HN, CA, HA, CB] + HBs + [C, O]
        pN, pCA, pC, prevC = N, CA, C, C
    xyz = np.array(xyz)
    per = len(_ALA)
    types = [a[1] for a in _ALA] * n_res
    charges = np.array([a[2] for a in _ALA] * n_res)
    idx = {a[0]: k for k, a in enumerate(_ALA)}
    bonds = []
    for r in range(n_res):
        o = r * per
        for a, b in (('N', 'HN'), ('N', 'CA'), ('CA', 'HA'), ('CA', 'CB'), ('CB', 'HB1'), ('CB', 'HB2'), ('CB', 'HB3'),
                     ('CA', 'C'), ('C', 'O')):
            bonds.append((o + idx[a], o + idx[b]))
        if r + 1 < n_res:
            bonds.append((o + idx['C'], o + per + idx['N']))
    bonds = np.array(bonds)
    n = len(xyz)
    adj = [[] for _ in range(n)]
    for a, b in bonds:
        adj[a].append(b); adj[b].append(a)
    angles = [(a, j, c) for j in range(n) for ia, a in enumerate(adj[j]) for c in adj[j][ia + 1:]]
    dihedrals = []
    for b, c in bonds:
        for a in adj[b]:
            if a == c: continue
            for d in adj[c]:
                if d == b or d == a: continue
                dihedrals.append((a, b, c, d))
    impropers = [(r * per + idx['C'], r * per + idx['CA'], (r + 1) * per + idx['N'], r * per + idx['O'])
                 for r in range(n_res - 1)]
    angles, dihedrals, impropers = np.array(angles), np.array(dihedrals), np.array(impropers).reshape(-1, 4)

    def dist(i, j):
        return np.linalg.norm(xyz[i] - xyz[j], axis=-1)

    def ang(i, j, k):
        u, v = xyz[i] - xyz[j], xyz[k] - xyz[j]
        return np.arccos(np.clip((u * v).sum(-1) / np.linalg.norm(u, axis=-1) / np.linalg.norm(v, axis=-1), -1, 1))

    heavy = np.array([not t.startswith('H') for t in types])
    kb = np.where(heavy[bonds[:, 0]] & heavy[bonds[:, 1]], 300.0, 340.0) * KCAL
    bond_par = np.stack([kb, dist(bonds[:, 0], bonds[:, 1])], 1)
    angle_par = np.stack([np.full(len(angles), 50.0 * KCAL), ang(angles[:, 0], angles[:, 1], angles[:, 2]),
                          np.zeros(len(angles)), np.zeros(len(angles))], 1)
    # generic 3-fold torsions; phase chosen so the built conformation is a minimum of each term
    phi_now = _dihedral_angles(xyz, dihedrals)
    mult = np.full(len(dihedrals), 3.0)
    delta = (mult * phi_now + np.pi) % (2 * np.pi)
    dihedral_par = np.stack([np.full(len(dihedrals), 0.2 * KCAL), mult, delta], 1)
    improper_par = np.stack([np.full(len(impropers), 120.0 * KCAL), _dihedral_angles(xyz, impropers)], 1) \
        if len(impropers) else np.zeros((0, 2))
    return dict(xyz=xyz - xyz.mean(0), types=types, charges=charges, bonds=bonds, angles=angles, dihedrals=dihedrals,
                impropers=impropers, bond_par=bond_par, angle_par=angle_par, dihedral_par=dihedral_par,
                improper_par=improper_par)


def _dihedral_angles(xyz, quads):
    if len(quads) == 0:
        return np.zeros(0)
    p = xyz[quads]
    r1, r2, r3 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 1], p[:, 3] - p[:, 2]
    n1, n2 = np.cross(r1, r2), np.cross(r2, r3)
    return np.arctan2(np.linalg.norm(r2, axis=1) * (r1 * n2).sum(1), (n1 * n2).sum(1))


def solvated_protein_box(n_atoms=92224, box=(108.86, 108.86, 77.76), protein_fraction=0.154, seed=20260002,
                         n_res=20):
    """Configs 3/4: poly-alanine helices on a coarse grid, solvated by lattice TIP3P, exactly
    n_atoms atoms.  Helix count is the nearest to protein_fraction * n_atoms that leaves a
    multiple of 3 atoms for water."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box, dtype=np.float64)
    tpl = _helix_template(n_res)
    per = len(tpl['xyz'])
    h = max(1, int(round(protein_fraction * n_atoms / per)))
    while (n_atoms - h * per) % 3:
        h += 1
    n_waters = (n_atoms - h * per) // 3
    # helix axis is roughly the template's principal axis; lay helices on a grid of cells that hold
    # one randomly spun helix each
    _, _, vt = np.linalg.svd(tpl['xyz'], full_matrices=False)
    local = tpl['xyz'] @ vt.T                      # principal axis along x
    half_len = np.abs(local[:, 0]).max() + 2.5
    radius = np.linalg.norm(local[:, 1:], axis=1).max() + 2.0
    cell = np.array([2 * half_len, 2 * radius, 2 * radius])
    counts = np.maximum(1, np.floor(box / cell).astype(int))
    if counts.prod() < h:
        raise RuntimeError('box too small for %d helices' % h)
    slots = rng.choice(counts.prod(), size=h, replace=False)
    slots.sort()
    ijk = np.stack(np.unravel_index(slots, counts), 1)
    centers = (ijk + 0.5) * (box / counts) - box / 2
    spin = rng.uniform(0, 2 * np.pi, size=h)
    cs, sn = np.cos(spin), np.sin(spin)
    rot = np.zeros((h, 3, 3)); rot[:, 0, 0] = 1
    rot[:, 1, 1], rot[:, 1, 2], rot[:, 2, 1], rot[:, 2, 2] = cs, -sn, sn, cs
    prot = (centers[:, None, :] + np.einsum('nij,kj->nki', rot, local)).reshape(-1, 3)
    water = _place_waters(rng, box, n_waters, obstacles=prot)
    pos = np.concatenate([prot, water])
    off = (np.arange(h) * per)[:, None, None]

    def rep(a):
        return (a[None] + off).reshape(-1, a.shape[1]) if len(a) else a.reshape(0, a.shape[1] if a.ndim == 2 else 4)

    wb, wa, wbp, wap = _water_topology(n_waters, h * per)
    types = tpl['types'] * h + ['OT', 'HT', 'HT'] * n_waters
    masses = np.array([MASS[t] for t in tpl['types']] * h + [MASS['OT'], MASS['HT'], MASS['HT']] * n_waters, dtype=np.float32)
    charges = np.concatenate([np.tile(tpl['charges'], h), np.tile([-0.834, 0.417, 0.417], n_waters)]).astype(np.float32)
    return SyntheticSystem(
        box=box, positions=pos.astype(np.float32), types=types, masses=masses, charges=charges,
        bonds=np.concatenate([rep(tpl['bonds']), wb]), angles=np.concatenate([rep(tpl['angles']), wa]),
        dihedrals=rep(tpl['dihedrals']), impropers=rep(tpl['impropers']),
        bond_par=np.concatenate([np.tile(tpl['bond_par'], (h, 1)), wbp]),
        angle_par=np.concatenate([np.tile(tpl['angle_par'], (h, 1)), wap]),
        dihedral_par=np.tile(tpl['dihedral_par'], (h, 1)), improper_par=np.tile(tpl['improper_par'], (h, 1)),
        name='solvated_%d' % n_atoms)


CONFIGS = {
    'water_23k': lambda: water_box(7852, 20260001),
    'protein_92k': lambda: solvated_protein_box(92224, (108.86, 108.86, 77.76), seed=20260002),
    'protein_1m': lambda: solvated_protein_box(1066628, (216.83, 216.83, 216.83), seed=20260003),
    'water_10m': lambda: water_box(3333334, 20260004),
}
