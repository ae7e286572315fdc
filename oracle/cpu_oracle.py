"""oracle/cpu_oracle.py — numpy/ctypes front end of the CPU oracle.

TEST INFRASTRUCTURE ONLY (see oracle/mdpy_oracle.c).  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by mdpy_b200/.

The host-level logic the reference keeps in Python is restated here in numpy, each
function citing the reference lines it follows; the per-pair loops live in
liboracle.so (mdpy_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'liboracle.so')
_lib = None

CELL_LIST_SKIN = 2.0  # mdpy/core/cell_list.py:17
NEIGHBOR_TEMPLATE = np.array(  # mdpy/constraint/__init__.py:13-19
    [[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], dtype=np.int32)


def build(force=False):
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ('mdpy_oracle.c', 'mdpy_oracle_impl.h')]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(['make', '-C', _HERE, '-B', 'liboracle.so'],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ora_lj_cell_f32.restype = C.c_longlong
        _lib.ora_lj_cell_f64.restype = C.c_longlong
        _lib.ora_pair_set_f32.restype = C.c_longlong
        _lib.ora_pair_set_f32_range.restype = C.c_longlong
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _real(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return np.float32, '_f32', C.c_float
    return np.float64, '_f64', C.c_double


def _pad_rows(a):
    """-1-padded int32 [n,w] table with w >= 1 (reference: topology.py:69-79)."""
    a = np.ascontiguousarray(a, dtype=np.int32)
    if a.ndim != 2 or a.shape[1] == 0:
        a = -np.ones((a.shape[0], 1), dtype=np.int32)
    return a


def _run_slices(piece, i_range, threads):
    """Evaluate piece((i0, i1)) over `threads` interleaved-size slices of i_range on host threads
    (ctypes releases the GIL); the caller adds the partial results."""
    i0, i1 = int(i_range[0]), int(i_range[1])
    threads = max(1, min(int(threads), i1 - i0)) if i1 > i0 else 1
    if threads == 1:
        return [piece((i0, i1))]
    # the all-pairs loop is triangular: equalise work with sqrt-spaced cut points
    frac = 1.0 - np.sqrt(1.0 - np.arange(threads + 1) / threads)
    cuts = np.unique(np.round(i0 + frac * (i1 - i0)).astype(int))
    cuts[0], cuts[-1] = i0, i1
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(piece, list(zip(cuts[:-1], cuts[1:]))))


# ----------------------------------------------------------------------------
# Reference restatement
# ----------------------------------------------------------------------------

def wrap_positions(positions, pbc, pbc_inv):
    """mdpy/utils/pbc.py:28-36.  Returns (wrapped, n_lost, first_lost)."""
    dt, suf, _ = _real(positions.dtype)
    pos = np.ascontiguousarray(positions, dtype=dt)
    out = np.empty_like(pos)
    first = C.c_int(-1)
    lost = getattr(lib(), 'ora_wrap_positions' + suf)(
        pos.shape[0], _p(pos), _p(np.ascontiguousarray(pbc, dtype=dt)),
        _p(np.ascontiguousarray(pbc_inv, dtype=dt)), _p(out), C.byref(first))
    return out, lost, first.value


def cell_attributes(pbc_diag, cutoff_radius, dtype=np.float32):
    """mdpy/core/cell_list.py:56-81 (set_cutoff_radius + _update_attributes).

    Raises ValueError where the reference raises CellListPoorDefinedError (:58-69).
    Returns (num_cells_vec int32[3], cell_matrix[3,3], cell_inv[3,3])."""
    pbc_diag = np.asarray(pbc_diag)
    if cutoff_radius == 0:
        raise ValueError('Cutoff radius is poor defined')
    if (np.floor(pbc_diag / (np.ones(3) * cutoff_radius)) < 2).any():
        raise ValueError('The cutoff_radius is too large to create cell list')
    ncell = np.floor((pbc_diag + CELL_LIST_SKIN) / (np.ones(3) * cutoff_radius)).astype(np.int32)
    ncell[ncell < 3] = 3
    cell_matrix = np.diag((pbc_diag + CELL_LIST_SKIN) / ncell).astype(dtype)
    cell_inv = np.linalg.inv(cell_matrix)
    return ncell, cell_matrix, cell_inv


def cell_list_update(positions, cell_inv, ncell):
    """mdpy/core/cell_list.py:83-116.  Returns (particle_cell_index [n,3], cell_list [nx,ny,nz,P])."""
    dt, suf, _ = _real(positions.dtype)
    pos = np.ascontiguousarray(positions, dtype=dt)
    ncell = np.ascontiguousarray(ncell, dtype=np.int32)
    pci = np.empty((pos.shape[0], 3), dtype=np.int32)
    P = getattr(lib(), 'ora_cell_index' + suf)(
        pos.shape[0], _p(pos), _p(np.ascontiguousarray(cell_inv, dtype=dt)), _p(ncell), _p(pci))
    cl = np.empty((ncell[0], ncell[1], ncell[2], P), dtype=np.int32)
    lib().ora_cell_fill(pos.shape[0], _p(pci), _p(ncell), P, _p(cl))
    return pci, cl


def lj_cell(positions, params, pbc, rc, bonded, scaling, cell_cutoff=None, i_range=None, threads=1):
    """CharmmNonbondedConstraint.update() on the CPU platform
    (charmm_nonbonded_constraint.py:183-195 -> cpu_kernel :64-108) including the cell
    list it reads from State (state.py:56-61; the list's cutoff is the largest constraint
    cutoff, ensemble.py:49-50, passed here as cell_cutoff, default rc... the reference's
    CellList default is 12, cell_list.py:20).  Returns (forces, energy, ordered_visits)."""
    dt, suf, creal = _real(positions.dtype)
    pos = np.ascontiguousarray(positions, dtype=dt)
    pbc = np.ascontiguousarray(pbc, dtype=dt)
    pbc_inv = np.ascontiguousarray(np.linalg.inv(pbc), dtype=dt)
    ncell, _, cell_inv = cell_attributes(pbc.diagonal(), rc if cell_cutoff is None else cell_cutoff, dt)
    pci, cl = cell_list_update(pos, cell_inv, ncell)
    bonded, scaling = _pad_rows(bonded), _pad_rows(scaling)
    params = np.ascontiguousarray(params, dtype=dt)
    fn = getattr(lib(), 'ora_lj_cell' + suf)

    def piece(rng):
        forces = np.empty_like(pos)
        e = C.c_double(0)
        visits = fn(pos.shape[0], _p(pos), _p(params), _p(pbc), _p(pbc_inv), creal(rc), _p(bonded), bonded.shape[1],
                    _p(scaling), scaling.shape[1], _p(pci), _p(cl), _p(ncell), cl.shape[3], int(rng[0]), int(rng[1]),
                    _p(forces), C.byref(e))
        return forces, e.value, visits
    parts = _run_slices(piece, (0, pos.shape[0]) if i_range is None else i_range, threads)
    return sum(p[0] for p in parts), sum(p[1] for p in parts), sum(p[2] for p in parts)


def coulomb_allpairs(positions, charges, pbc, bonded, k, i_range=None, threads=1):
    """ElectrostaticConstraint.update() on the CPU platform
    (electrostatic_constraint.py:137-145 -> cpu_kernel :52-79); k = 4 pi eps0."""
    dt, suf, _ = _real(positions.dtype)
    pos = np.ascontiguousarray(positions, dtype=dt)
    pbc = np.ascontiguousarray(pbc, dtype=dt)
    pbc_inv = np.ascontiguousarray(np.linalg.inv(pbc), dtype=dt)
    bonded = _pad_rows(bonded)
    q = np.ascontiguousarray(charges, dtype=dt).reshape(-1)
    fn = getattr(lib(), 'ora_coulomb_allpairs' + suf)

    def piece(rng):
        forces = np.empty_like(pos)
        e = C.c_double(0)
        fn(pos.shape[0], _p(pos), _p(q), _p(bonded), bonded.shape[1], _p(pbc), _p(pbc_inv), C.c_double(k),
           int(rng[0]), int(rng[1]), _p(forces), C.byref(e))
        return forces, e.value
    parts = _run_slices(piece, (0, pos.shape[0]) if i_range is None else i_range, threads)
    return sum(p[0] for p in parts), sum(p[1] for p in parts)


def verlet(positions, velocities, masses, pbc, dt_fs, num_steps, force_fn, snapshot_steps=None):
    """VerletIntegrator.integrate (verlet_integrator.py:20-50) for a fresh integrator:
    x_prev = x - v dt + a dt^2 (:31-34, sic), x_new = 2x - x_prev + a dt^2 (:40-43), positions
    wrapped into State every step (:44 -> state.py:56-61), reported velocity
    minimg(x_cur - x_prev)/(2 dt) (:47-50, sic — half the true value).
    force_fn(wrapped_positions) -> forces [n,3].  float64 host arithmetic like the reference.
    Returns (state_positions, state_velocities, cur_positions_unwrapped, pre_positions); with snapshot_steps
    also a dict {step: cur_positions_unwrapped after that many steps}."""
    pbc = np.asarray(pbc, dtype=np.float64)
    pbc_inv = np.linalg.inv(pbc)
    masses = np.asarray(masses, dtype=np.float64).reshape(-1, 1)
    state_pos, _, _ = wrap_positions(np.asarray(positions, dtype=np.float64), pbc, pbc_inv)
    acc = force_fn(state_pos) / masses
    cur = state_pos.copy()
    pre = cur - np.asarray(velocities, dtype=np.float64) * dt_fs + acc * dt_fs ** 2
    snaps = {}
    for step in range(num_steps):
        if step != 0:
            acc = force_fn(state_pos) / masses
        cur, pre = 2 * cur - pre + acc * dt_fs ** 2, cur
        state_pos, lost, _ = wrap_positions(cur, pbc, pbc_inv)
        if lost:
            raise RuntimeError('ParticleLossError')
        if snapshot_steps is not None and (step + 1) in snapshot_steps:
            snaps[step + 1] = cur.copy()
    d = cur - pre
    s = d @ pbc_inv
    s -= np.round(s)
    vel = (s @ pbc) / 2 / dt_fs
    if snapshot_steps is not None:
        return state_pos, vel, cur, pre, snaps
    return state_pos, vel, cur, pre


def config1_force_fn(c1, threads=1):
    """Force function of the example system without its dihedral term (the terms of
    tests/golden/config1_verlet_f64.npz): LJ + bare all-pairs Coulomb from the C restatement, bonds / angles /
    impropers from oracle/bonded.py.  c1 = the config1_f64 fixture."""
    from . import bonded
    k_e = 1.0 / (4 * np.pi * float(np.float32(0.5727653)))   # 1 / (4 pi EPSILON0), the reference's float32-rounded constant (Q7)
    box = np.asarray(c1['box'], dtype=np.float64)

    def force(pos):
        t = nonbonded_bruteforce(pos, box, c1['lj_table'], c1['charges'], c1['bonded'], c1['scaling'],
                                 rc_lj=float(c1['rc']), coul_mode=2, k_e=k_e, threads=threads)
        f = t['f_lj'] + t['f_coul']
        f += bonded.bonds(pos, box, c1['CharmmBondConstraint_idx'], c1['CharmmBondConstraint_par'])[0]
        f += bonded.angles(pos, box, c1['CharmmAngleConstraint_idx'], c1['CharmmAngleConstraint_par'])[0]
        f += bonded.impropers(pos, box, c1['CharmmImproperConstraint_idx'], c1['CharmmImproperConstraint_par'])[0]
        return f
    return force


# ----------------------------------------------------------------------------
# Float64 truth
# ----------------------------------------------------------------------------

def nonbonded_bruteforce(positions, box, params, charges, bonded, scaling, rc_lj=12.0, r_on=None,
                         coul_mode=0, k_e=0.0, alpha=0.0, rc_coul=None, i_range=None, threads=1):
    """All-pairs float64 LJ (+switch) and Coulomb (mode 1: erfc + excluded-pair erf
    correction; mode 2: the reference's bare minimum-image sum).  See mdpy_oracle.c.
    Returns dict(f_lj, f_coul, e_lj, e_coul, e_excl, n_lj, n_coul); forces are only
    filled for atoms in i_range (default all).  threads > 1: the rows of i_range are split over host
    threads (rows are independent; partial energies / counts add up)."""
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    n = pos.shape[0]
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(3)
    params = np.ascontiguousarray(params, dtype=np.float64)
    q = np.ascontiguousarray(charges, dtype=np.float64).reshape(-1)
    bonded, scaling = _pad_rows(bonded), _pad_rows(scaling)
    i0, i1 = (0, n) if i_range is None else i_range
    f_lj = np.zeros((n, 3)); f_c = np.zeros((n, 3))

    def piece(rng):
        en = np.zeros(5); cnt = np.zeros(2, dtype=np.int64)
        lib().ora_nonbonded_bruteforce(
            n, _p(pos), _p(box), _p(params), _p(q), _p(bonded), bonded.shape[1], _p(scaling),
            scaling.shape[1], C.c_double(rc_lj), C.c_double(rc_lj if r_on is None else r_on),
            int(coul_mode), C.c_double(k_e), C.c_double(alpha),
            C.c_double(rc_lj if rc_coul is None else rc_coul), int(rng[0]), int(rng[1]), _p(f_lj), _p(f_c),
            _p(en), _p(cnt))     # each slice writes its own rows of the shared force arrays
        return en, cnt
    if threads <= 1 or i1 - i0 < 2 * threads:
        parts = [piece((i0, i1))]
    else:
        cuts = np.linspace(i0, i1, int(threads) + 1).round().astype(int)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(int(threads)) as ex:
            parts = list(ex.map(piece, list(zip(cuts[:-1], cuts[1:]))))
    en = sum(p[0] for p in parts); cnt = sum(p[1] for p in parts)
    return dict(f_lj=f_lj, f_coul=f_c, e_lj=en[0], e_coul=en[1], e_excl=en[2], e_lj_abs=en[3], e_coul_abs=en[4],
                n_lj=int(cnt[0]), n_coul=int(cnt[1]))


def pair_set_f32(positions, box, rc, bonded, threads=1):
    """The canonical fp32 in-cutoff pair set, i<j, lexicographic (mdpy_oracle.c:ora_pair_set_f32).
    threads > 1: rows split over host threads (sqrt-spaced cuts: row i has n-i-1 candidates), results
    concatenated in row order."""
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    n = pos.shape[0]
    box = np.ascontiguousarray(box, dtype=np.float32).reshape(3)
    bonded = _pad_rows(bonded)

    def piece(rng):
        cap = max(1024, int((rng[1] - rng[0]) * 500))
        while True:
            oi = np.empty(cap, dtype=np.int32); oj = np.empty(cap, dtype=np.int32)
            cnt = lib().ora_pair_set_f32_range(n, _p(pos), _p(box), C.c_float(rc), _p(bonded), bonded.shape[1],
                                               int(rng[0]), int(rng[1]), _p(oi), _p(oj), C.c_longlong(cap))
            if cnt <= cap:
                return np.stack([oi[:cnt], oj[:cnt]], axis=1)
            cap = int(cnt)
    parts = _run_slices(piece, (0, n), threads)
    return parts[0] if len(parts) == 1 else np.concatenate(parts)


def ewald_recip(positions, charges, box, alpha, kmax, k_e):
    """Explicit k-space Ewald sum.  Returns (forces, e_rec, e_self, e_background)."""
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    q = np.ascontiguousarray(charges, dtype=np.float64).reshape(-1)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(3)
    kmax = np.ascontiguousarray(np.broadcast_to(kmax, 3), dtype=np.int32)
    f = np.zeros_like(pos); en = np.zeros(3)
    lib().ora_ewald_recip(pos.shape[0], _p(pos), _p(q), _p(box), C.c_double(alpha), _p(kmax),
                          C.c_double(k_e), _p(f), _p(en))
    return f, en[0], en[1], en[2]


def ewald_exact(positions, charges, box, bonded, k_e, tol_exp=36.0):
    """Converged Ewald energy/forces (the PME truth; parity otherwise unpinned):
    alpha chosen so the real-space sum converges inside the minimum image
    (erfc(alpha L/2) ~ e^-tol_exp), k-space summed until exp(-pi^2 m^2/alpha^2) < e^-tol_exp.
    Excluded (bonded) pairs are removed exactly.  Returns (forces, energy)."""
    box = np.asarray(box, dtype=np.float64).reshape(3)
    n = positions.shape[0]
    alpha = 2.0 * np.sqrt(tol_exp) / box.min()
    kmax = np.ceil(np.sqrt(tol_exp) * alpha * box / np.pi).astype(np.int32)
    zero_params = np.zeros((n, 4))
    d = nonbonded_bruteforce(positions, box, zero_params, charges, bonded,
                             -np.ones((n, 1), dtype=np.int32), rc_lj=0.0, coul_mode=1, k_e=k_e,
                             alpha=alpha, rc_coul=0.5 * box.min())
    f_rec, e_rec, e_self, e_bg = ewald_recip(positions, charges, box, alpha, kmax, k_e)
    return d['f_coul'] + f_rec, d['e_coul'] + d['e_excl'] + e_rec + e_self + e_bg


# ----------------------------------------------------------------------------
# Langevin (G-JF) step restated on the host
# ----------------------------------------------------------------------------
# The reference's own LangevinIntegrator is not a usable oracle (force sign, state carry-over);
# what IS taken from it are the coefficients a, b, sigma (mdpy/integrator/langevin_integrator.py:23-31).
# The update is the published G-JF scheme (Gronbech-Jensen & Farago, Mol. Phys. 111, 983 (2013)); the noise
# is Philox4x32-10 (Salmon et al., SC'11: multipliers 0xD2511F53 / 0xCD9E8D57, Weyl keys 0x9E3779B9 /
# 0xBB67AE85) with counter (atom, step_lo, step_hi, 'MDPY') and key = seed, then Box-Muller on 24-bit uniforms.

def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments uint32 arrays / scalars.  Returns 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint64(k0); k1 = np.uint64(k1)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]; p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return [x.astype(np.uint32) for x in c]


def langevin_noise(seed, n_atoms, step):
    """[n,3] standard normal variates of (atom, step): the float32 Box-Muller of the device kernel
    (uniform = (bits >> 8 + 0.5) 2^-24), evaluated with float64 log / sin / cos."""
    atom = np.arange(n_atoms, dtype=np.uint32)
    r = philox4x32_10(atom, np.uint32(step & 0xFFFFFFFF), np.uint32(step >> 32), np.uint32(0x4d445059),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24) for x in r]
    u = [x.astype(np.float64) for x in u]
    m0, m1 = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
    two_pi = float(np.float32(6.283185307179586))
    return np.stack([m0 * np.cos(two_pi * u[1]), m0 * np.sin(two_pi * u[1]), m1 * np.cos(two_pi * u[3])], axis=1)


def gjf_step(x, v, f_old, force_fn, masses, dt, gamma, kT, seed, step):
    """One G-JF step with a = (1 - g dt/2)/(1 + g dt/2), b = 1/(1 + g dt/2) (langevin_integrator.py:23-31):
        x' = x + b dt v + b dt^2/(2m) f + b dt/(2m) beta,      beta = sqrt(2 g m kT dt) xi(atom, step)
        v' = a v + dt/(2m) (a f + f') + b/m beta,              f' = force_fn(x')
    Returns (x', v', f')."""
    m = np.asarray(masses, dtype=np.float64).reshape(-1, 1)
    half = 0.5 * gamma * dt
    a, b = (1.0 - half) / (1.0 + half), 1.0 / (1.0 + half)
    beta = np.sqrt(2.0 * gamma * kT * dt * m) * langevin_noise(seed, x.shape[0], step)
    x_new = x + b * dt * v + 0.5 * b * dt * dt / m * f_old + 0.5 * b * dt / m * beta
    f_new = force_fn(x_new)
    v_new = a * v + 0.5 * dt / m * (a * f_old + f_new) + b / m * beta
    return x_new, v_new, f_new


def steepest_descent(positions, box, force_energy_fn, alpha=0.01, energy_tolerance=0.001, max_iterations=1000):
    """SteepestDescentMinimizer.minimize (mdpy/minimizer/steepest_descent_minimizer.py:30-53) in float64:
    x_i += alpha F_i / |F_i| per atom (:38-41), wrap, re-evaluate; stop when |E - E_prev| / |E_prev| < tolerance (:44,48).
    force_energy_fn(wrapped_positions) -> (forces, potential_energy).  Returns (positions, iterations, energies)."""
    box = np.asarray(box, dtype=np.float64).reshape(3)
    x = np.asarray(positions, dtype=np.float64)
    x = x - box * np.round(x / box)
    f, e = force_energy_fn(x)
    energies = [e]
    it = 0
    while it < max_iterations:
        x = x + alpha * f / np.linalg.norm(f, axis=1).reshape(-1, 1)
        x = x - box * np.round(x / box)
        f, e = force_energy_fn(x)
        it += 1
        energies.append(e)
        if abs((e - energies[-2]) / energies[-2]) < energy_tolerance:
            break
    return x, it, energies


def gjf_rigid_water_call(x, v, f0, force_fn, masses, triplets, d_oh, d_hh, dt, gamma, kT, seed, step):
    """One single-step call of the G-JF integrator with rigid three-site waters on a fresh state, in the order the step
    is split around the force evaluation (advance x with f(x_0), evaluate f(x_1), finish v):
        c   = x + b dt v + b dt^2/(2m) f0 + b dt/(2m) beta          (unconstrained)
        x_1 = SETTLE(x, c);  v += (x_1 - c) / (b dt)                 (the constraint force's share of the velocity update)
        v_1 = RATTLE_v(x_1, a v + dt/(2m) (a f0 + f(x_1)) + b/m beta)
    x, v: [N,3] float64 (whole molecules), triplets [W,3].  Free atoms (not in any triplet) take the plain G-JF step.
    Returns (x_1, v_1, f_1)."""
    from . import settle as st
    m = np.asarray(masses, dtype=np.float64).reshape(-1, 1)
    half = 0.5 * gamma * dt
    a, b = (1.0 - half) / (1.0 + half), 1.0 / (1.0 + half)
    beta = np.sqrt(2.0 * gamma * kT * dt * m) * langevin_noise(seed, x.shape[0], step)
    c = x + b * dt * v + 0.5 * b * dt * dt / m * f0 + 0.5 * b * dt / m * beta
    t = np.asarray(triplets)
    m_o, m_h = float(m[t[0, 0], 0]), float(m[t[0, 1], 0])
    x1 = c.copy()
    x1[t] = st.settle(x[t], c[t], m_o, m_h, d_oh, d_hh)
    v_mid = v + (x1 - c) / (b * dt)
    f1 = force_fn(x1)
    v1 = a * v_mid + 0.5 * dt / m * (a * f0 + f1) + b / m * beta
    v1[t] = st.rattle_velocities(x1[t], v1[t], m_o, m_h)
    return x1, v1, f1
