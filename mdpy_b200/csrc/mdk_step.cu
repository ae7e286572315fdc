// mdk_step.cu — everything around the pair/PME kernels inside one MD step: the
// reference-semantics all-pairs Coulomb, CHARMM bonded terms, Verlet / Langevin updates
// and the term dispatcher behind mdk_compute.
#include "mdk_common.cuh"

namespace mdk {

// ===========================================================================
// Bare Coulomb, all pairs, minimum image, no cutoff — ElectrostaticConstraint.cpu_kernel
// (electrostatic_constraint.py:52-79): E = sum_{i<j, j not bonded to i} q_i q_j/(4 pi eps0 r).
// O(N^2) by definition of the reference; evaluated in float64 (full shell, force on i only,
// j staged through shared memory), then the bonded pairs are taken out again.
constexpr int BARE_T = 128;

__global__ void __launch_bounds__(BARE_T)
k_coulomb_bare(int n, const float4 *__restrict__ xs, const float *__restrict__ q, const int *__restrict__ order,
               double k_e, double Lx, double Ly, double Lz, long long *__restrict__ f_acc,
               long long *__restrict__ e_acc, const double *__restrict__ x_cur, const double *__restrict__ q64) {
    __shared__ double4 sj[BARE_T];
    const int i = blockIdx.x * BARE_T + threadIdx.x;
    double xi = 0, yi = 0, zi = 0, qi = 0;
    // DOUBLE precision: float64 positions / charges of the state instead of the float32 tile-order copy
    auto fetch = [&](int k) {
        const int a = order[k];
        if (x_cur) return make_double4(x_cur[3 * (size_t)a], x_cur[3 * (size_t)a + 1], x_cur[3 * (size_t)a + 2], q64[a]);
        const float4 v = xs[k];
        return make_double4(v.x, v.y, v.z, (double)q[a]);
    };
    if (i < n) { const double4 a = fetch(i); xi = a.x; yi = a.y; zi = a.z; qi = k_e * a.w; }
    double fx = 0, fy = 0, fz = 0, e = 0;
    for (int base = 0; base < n; base += BARE_T) {
        int j = base + threadIdx.x;
        double4 v = make_double4(0, 0, 0, 0);
        if (j < n) v = fetch(j);
        __syncthreads();
        sj[threadIdx.x] = v;
        __syncthreads();
        int lim = min(BARE_T, n - base);
        for (int t = 0; t < lim; ++t) {
            if (base + t == i || i >= n) continue;  // idle lanes sit at the origin: 0 * inf otherwise
            double4 b = sj[t];
            double dx = b.x - xi, dy = b.y - yi, dz = b.z - zi;
            dx -= Lx * rint(dx / Lx); dy -= Ly * rint(dy / Ly); dz -= Lz * rint(dz / Lz);
            double r2 = dx * dx + dy * dy + dz * dz;
            double rinv = rsqrt(r2);
            double qq = qi * b.w;
            double er = qq * rinv;
            e += er;
            double g = -er * rinv * rinv;  // dE/dr / r
            fx += g * dx; fy += g * dy; fz += g * dz;
        }
    }
    if (i < n) {
        atomic_add_fix(&f_acc[3 * (size_t)i + 0], to_fix(fx));
        atomic_add_fix(&f_acc[3 * (size_t)i + 1], to_fix(fy));
        atomic_add_fix(&f_acc[3 * (size_t)i + 2], to_fix(fz));
    }
    e = warp_sum(0.5 * e);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomic_add_fix(&e_acc[MDK_E_COUL_BARE], to_fix(e));
}

__global__ void k_coulomb_bare_excl(int n, int wb, const int *__restrict__ excl_s, const float4 *__restrict__ xs,
                                    const float *__restrict__ q, const int *__restrict__ order, double k_e,
                                    double Lx, double Ly, double Lz, long long *__restrict__ f_acc,
                                    long long *__restrict__ e_acc, const double *__restrict__ x_cur, const double *__restrict__ q64) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (t < n * wb) {
        int k = t / wb;
        int p = excl_s[t];
        if (p > k) {
            double4 a, b;
            if (x_cur) {
                const size_t ia = (size_t)order[k], ib = (size_t)order[p];
                a = make_double4(x_cur[3 * ia], x_cur[3 * ia + 1], x_cur[3 * ia + 2], 0.0);
                b = make_double4(x_cur[3 * ib], x_cur[3 * ib + 1], x_cur[3 * ib + 2], 0.0);
            } else {
                const float4 fa = xs[k], fb = xs[p];
                a = make_double4(fa.x, fa.y, fa.z, 0.0); b = make_double4(fb.x, fb.y, fb.z, 0.0);
            }
            double dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
            dx -= Lx * rint(dx / Lx); dy -= Ly * rint(dy / Ly); dz -= Lz * rint(dz / Lz);
            double r2 = dx * dx + dy * dy + dz * dz;
            double rinv = rsqrt(r2);
            double qq = q64 ? k_e * q64[order[k]] * q64[order[p]] : k_e * (double)q[order[k]] * (double)q[order[p]];
            e = -qq * rinv;
            double g = qq * rinv * rinv * rinv;  // minus the pair's dE/dr / r
            atomic_add_fix(&f_acc[3 * (size_t)k + 0], to_fix(g * dx));
            atomic_add_fix(&f_acc[3 * (size_t)k + 1], to_fix(g * dy));
            atomic_add_fix(&f_acc[3 * (size_t)k + 2], to_fix(g * dz));
            atomic_add_fix(&f_acc[3 * (size_t)p + 0], to_fix(-g * dx));
            atomic_add_fix(&f_acc[3 * (size_t)p + 1], to_fix(-g * dy));
            atomic_add_fix(&f_acc[3 * (size_t)p + 2], to_fix(-g * dz));
        }
    }
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomic_add_fix(&e_acc[MDK_E_COUL_BARE], to_fix(e));
}

int coulomb_bare(mdk_ctx *c) {
    if (!c->have_coul) return fail(c, MDK_ERR_NOT_BOUND, "Coulomb term requested before mdk_set_coulomb");
    PhaseTimer pt(c, PH_BARE);
    long long *e_acc = reinterpret_cast<long long *>(c->e_acc.p);
    const double *xd = (c->dprec && c->have_q64) ? c->x_cur.p : nullptr, *qd = xd ? c->q64.p : nullptr;
    k_coulomb_bare<<<(c->n + BARE_T - 1) / BARE_T, BARE_T, 0, c->stream>>>(
        c->n, c->xs.p, c->q.p, c->order.p, c->k_e, c->box.Ld[0], c->box.Ld[1], c->box.Ld[2], c->f_acc.p, e_acc, xd, qd);
    ++c->n_launches;
    if (c->wb > 0) {
        int total = c->n * c->wb;
        k_coulomb_bare_excl<<<(total + 255) / 256, 256, 0, c->stream>>>(
            c->n, c->wb, c->excl_s.p, c->xs.p, c->q.p, c->order.p, c->k_e, c->box.Ld[0], c->box.Ld[1],
            c->box.Ld[2], c->f_acc.p, e_acc, xd, qd);
        ++c->n_launches;
    }
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

// ===========================================================================
// CHARMM bonded terms.  One thread per term, float64 geometry on the wrapped
// tile-order positions (O(N) terms: the arithmetic is free, and acos / atan2 near their singular points lose
// digits in fp32), minimum image per bond vector, fixed-point force atomics.
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

struct BoxF { double L[3]; };
// where the O(N) terms read positions (and the scaled charge) of a tile slot: the float32 tile-order copy, or — DOUBLE
// precision — the float64 state itself
struct P4 { double x, y, z, w; };
struct PosSrc {
    const float4 *xs;
    const double *x_cur, *q64;
    const int *order;
    double sqrt_ke;
    __device__ __forceinline__ P4 operator()(int slot) const {
        if (x_cur) {
            const size_t a = (size_t)order[slot];
            return P4{x_cur[3 * a], x_cur[3 * a + 1], x_cur[3 * a + 2], q64 ? q64[a] * sqrt_ke : 0.0};
        }
        const float4 v = xs[slot];
        return P4{v.x, v.y, v.z, v.w};
    }
};
__device__ __forceinline__ V3 mi_vec(P4 a, P4 b, const BoxF &bx) {  // b - a, minimum image
    double dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    return {dx - bx.L[0] * rint(dx / bx.L[0]), dy - bx.L[1] * rint(dy / bx.L[1]), dz - bx.L[2] * rint(dz / bx.L[2])};
}
__device__ __forceinline__ void add_force(long long *f_acc, int slot, V3 f) {
    atomic_add_fix(&f_acc[3 * (size_t)slot + 0], to_fix(f.x));
    atomic_add_fix(&f_acc[3 * (size_t)slot + 1], to_fix(f.y));
    atomic_add_fix(&f_acc[3 * (size_t)slot + 2], to_fix(f.z));
}
// per-term conversion to fixed point before the (integer) warp sum: the total does not depend on how
// the terms are grouped into warps, blocks or ranks
__device__ __forceinline__ void block_energy(double e, long long *e_acc, int which) {
    const long long v = warp_sum_ll(to_fix(e));
    if ((threadIdx.x & 31) == 0 && v != 0) atomic_add_fix(&e_acc[which], v);
}

// E = k (r - r0)^2   (charmm_bond_constraint.py:53-73)
// A term belongs to the rank that owns its first atom (tile slot in [own_lo, own_hi)); every rank walks the whole
// term list and skips the others' terms.
__device__ __forceinline__ double term_bond(int t, int own_lo, int own_hi, int nb, const int *__restrict__ idx, const float *__restrict__ par,
        const int *__restrict__ inv_order, const PosSrc &xs, const BoxF &bx, long long *__restrict__ f_acc) {
    double e = 0.0;
    int s1 = t < nb ? inv_order[idx[2 * t]] : -1;
    if (s1 >= own_lo && s1 < own_hi) {
        int s2 = inv_order[idx[2 * t + 1]];
        double k = par[2 * t], r0 = par[2 * t + 1];
        V3 d = mi_vec(xs(s1), xs(s2), bx);
        double r = sqrt(dot(d, d));
        double dr = r - r0;
        e = (k * dr * dr);
        V3 f = (2. * k * dr / r) * d;  // force on atom 1 (toward 2 when stretched)
        add_force(f_acc, s1, f);
        add_force(f_acc, s2, -1. * f);
    }
    return e;
}

// E = k (theta - theta0)^2 + k_ub (r13 - r_ub)^2   (charmm_angle_constraint.py:55-96)
__device__ __forceinline__ double term_angle(int t, int own_lo, int own_hi, int na, const int *__restrict__ idx, const float *__restrict__ par,
        const int *__restrict__ inv_order, const PosSrc &xs, const BoxF &bx, long long *__restrict__ f_acc) {
    double e = 0.0;
    int s1 = t < na ? inv_order[idx[3 * t]] : -1;
    if (s1 >= own_lo && s1 < own_hi) {
        int s2 = inv_order[idx[3 * t + 1]], s3 = inv_order[idx[3 * t + 2]];
        double k = par[4 * t], th0 = par[4 * t + 1], ku = par[4 * t + 2], u0 = par[4 * t + 3];
        P4 p1 = xs(s1), p2 = xs(s2), p3 = xs(s3);
        V3 r21 = mi_vec(p2, p1, bx), r23 = mi_vec(p2, p3, bx);
        double l21 = sqrt(dot(r21, r21)), l23 = sqrt(dot(r23, r23));
        double ct = dot(r21, r23) / (l21 * l23);
        ct = fmin(1., fmax(-1., ct));
        double th = acos(ct);
        double st = sqrt(fmax(1. - ct * ct, 1e-24));
        double dEdth = 2. * k * (th - th0);
        // d theta / d r1 = -(r23/l23 - ct r21/l21) / (l21 st)
        V3 e21 = (1. / l21) * r21, e23 = (1. / l23) * r23;
        V3 f1 = (dEdth / (l21 * st)) * (e23 - ct * e21);
        V3 f3 = (dEdth / (l23 * st)) * (e21 - ct * e23);
        e = (k * (th - th0) * (th - th0));
        add_force(f_acc, s2, -1. * (f1 + f3));
        if (ku != 0.) {  // Urey-Bradley 1-3 spring acts on the end atoms only
            V3 r13 = mi_vec(p1, p3, bx);
            double l13 = sqrt(dot(r13, r13));
            double du = l13 - u0;
            e += (ku * du * du);
            V3 fu = (2. * ku * du / l13) * r13;
            f1 = f1 + fu;
            f3 = f3 - fu;
        }
        add_force(f_acc, s1, f1);
        add_force(f_acc, s3, f3);
    }
    return e;
}

// Torsion geometry shared by dihedrals and impropers: phi by the reference's atan2
// convention (utils/geometry.py:84-96), analytic gradient (Blondel & Karplus form).
__device__ __forceinline__ double torsion(P4 p1, P4 p2, P4 p3, P4 p4, const BoxF &bx, V3 &g1,
                                         V3 &g2, V3 &g3, V3 &g4) {
    V3 r1 = mi_vec(p1, p2, bx), r2 = mi_vec(p2, p3, bx), r3 = mi_vec(p3, p4, bx);
    V3 n1 = cross(r1, r2), n2 = cross(r2, r3);
    double l2 = sqrt(dot(r2, r2));
    double x = l2 * dot(r1, n2), y = dot(n1, n2);
    double phi = atan2(x, y);
    double n1sq = fmax(dot(n1, n1), 1e-30), n2sq = fmax(dot(n2, n2), 1e-30);
    // d phi / d r_a, d phi / d r_d
    g1 = (-l2 / n1sq) * n1;
    g4 = (l2 / n2sq) * n2;
    const double a = -dot(r1, r2) / (l2 * l2), b = -dot(r3, r2) / (l2 * l2);
    g2 = (a - 1.) * g1 + (-b) * g4;
    g3 = (b - 1.) * g4 + (-a) * g1;
    return phi;
}

// E = k (1 + cos(n phi - delta))   (charmm_dihedral_constraint.py:59-95; the force is the
// analytic gradient of this energy — the reference's `-k (1 - n sin(..))` at :80 is not).
__device__ __forceinline__ double term_dihedral(int t, int own_lo, int own_hi, int nd, const int *__restrict__ idx, const float *__restrict__ par,
        const int *__restrict__ inv_order, const PosSrc &xs, const BoxF &bx, long long *__restrict__ f_acc) {
    double e = 0.0;
    int s0 = t < nd ? inv_order[idx[4 * t]] : -1;
    if (s0 >= own_lo && s0 < own_hi) {
        int s[4];
        for (int a = 0; a < 4; ++a) s[a] = inv_order[idx[4 * t + a]];
        double k = par[3 * t], nn = par[3 * t + 1], delta = par[3 * t + 2];
        V3 g1, g2, g3, g4;
        double phi = torsion(xs(s[0]), xs(s[1]), xs(s[2]), xs(s[3]), bx, g1, g2, g3, g4);
        double arg = nn * phi - delta;
        e = (k * (1. + cos(arg)));
        double dEdphi = -k * nn * sin(arg);
        add_force(f_acc, s[0], (-dEdphi) * g1);
        add_force(f_acc, s[1], (-dEdphi) * g2);
        add_force(f_acc, s[2], (-dEdphi) * g3);
        add_force(f_acc, s[3], (-dEdphi) * g4);
    }
    return e;
}

// E = k (psi - psi0)^2   (charmm_improper_constraint.py:57-94)
__device__ __forceinline__ double term_improper(int t, int own_lo, int own_hi, int ni, const int *__restrict__ idx, const float *__restrict__ par,
        const int *__restrict__ inv_order, const PosSrc &xs, const BoxF &bx, long long *__restrict__ f_acc) {
    double e = 0.0;
    int s0 = t < ni ? inv_order[idx[4 * t]] : -1;
    if (s0 >= own_lo && s0 < own_hi) {
        int s[4];
        for (int a = 0; a < 4; ++a) s[a] = inv_order[idx[4 * t + a]];
        double k = par[2 * t], psi0 = par[2 * t + 1];
        V3 g1, g2, g3, g4;
        double psi = torsion(xs(s[0]), xs(s[1]), xs(s[2]), xs(s[3]), bx, g1, g2, g3, g4);
        double d = psi - psi0;
        e = (k * d * d);
        double dE = 2. * k * d;
        add_force(f_acc, s[0], (-dE) * g1);
        add_force(f_acc, s[1], (-dE) * g2);
        add_force(f_acc, s[2], (-dE) * g3);
        add_force(f_acc, s[3], (-dE) * g4);
    }
    return e;
}

// erf for the small arguments of bonded neighbours (alpha r < 1.6): the Maclaurin series, 26 terms to 1e-16
__device__ __forceinline__ double erf_small(double x) {
    if (x > 1.6) return erf(x);
    const double y = x * x;
    double term = 1.0, sum = 1.0;
#pragma unroll
    for (int n = 1; n <= 26; ++n) {
        term *= -y / n;
        sum += term / (2 * n + 1);
    }
    return 1.1283791670955126 * x * sum;
}

// Excluded-pair Ewald correction: the reciprocal sum contains every pair, also the bonded ones the direct sum
// skips; remove -k_e q_i q_j erf(alpha r)/r for each of them.  One thread per excluded pair (a < b, matrix ids; the
// compact list mdk_set_exclusions builds from the -1-padded bonded_particles table); owner = the owner of atom a.
__device__ __forceinline__ double term_excl(int t, int own_lo, int own_hi, int n_pairs, const int2 *__restrict__ pairs,
                                            const int *__restrict__ inv_order, const PosSrc &xs, const BoxF &bx,
                                            double alpha, long long *__restrict__ f_acc) {
    if (t >= n_pairs) return 0.0;
    const int2 pr = pairs[t];
    const int k = inv_order[pr.x];
    if (k < own_lo || k >= own_hi) return 0.0;
    const int p = inv_order[pr.y];
    const P4 a = xs(k), b = xs(p);
    const V3 d = mi_vec(a, b, bx);
    const double r2 = dot(d, d), r = sqrt(r2);
    const double qq = a.w * b.w;
    const double ar = alpha * r, erf_ar = erf_small(ar);
    // dE/dr = -qq (2 alpha/sqrt(pi) exp(-a^2 r^2)/r - erf/r^2);  F_i = dE/dr d/r
    const double g = -qq * (1.1283791670955126 * alpha * exp(-ar * ar) / r - erf_ar / r2) / r;
    add_force(f_acc, k, g * d);
    add_force(f_acc, p, (-g) * d);
    return -qq * erf_ar / r;
}

// All O(N) terms of a force evaluation in ONE launch: the blocks are dealt to the term kinds (bonds, angles, dihedrals,
// impropers, excluded-pair corrections) by ranges, a block handles one kind.  At 23k atoms five separate launches of
// ~5 us each were pure launch latency on the step's critical path.
struct AuxTable {
    int n[5];                // terms of each kind (0 = kind not requested)
    int blk_off[6];          // first block of each kind
    const int *idx[4];
    const float *par[4];
    const int2 *pairs;
    const int *sel[5];       // decomposed runs: the terms this rank owns (indices into the full lists), or null
};
constexpr int AUX_T = 128;

__global__ void __launch_bounds__(AUX_T)
k_aux_terms(AuxTable tb, int own_lo, int own_hi, const int *__restrict__ inv_order, PosSrc xs, BoxF bx,
            double alpha, long long *__restrict__ f_acc, long long *__restrict__ e_acc) {
    int kind = 0;
    while (kind < 4 && (int)blockIdx.x >= tb.blk_off[kind + 1]) ++kind;
    int t = ((int)blockIdx.x - tb.blk_off[kind]) * AUX_T + threadIdx.x;
    int nk = tb.n[kind];
    if (tb.sel[kind]) {           // n[kind] counts the selection: map to the term's index in the full list
        if (t < nk) { t = tb.sel[kind][t]; nk = t + 1; } else { t = 0; nk = 0; }
    }
    tb.n[kind] = nk;
    double e = 0.0;
    int slot = MDK_E_BOND;
    switch (kind) {
        case 0: e = term_bond(t, own_lo, own_hi, tb.n[0], tb.idx[0], tb.par[0], inv_order, xs, bx, f_acc); slot = MDK_E_BOND; break;
        case 1: e = term_angle(t, own_lo, own_hi, tb.n[1], tb.idx[1], tb.par[1], inv_order, xs, bx, f_acc); slot = MDK_E_ANGLE; break;
        case 2: e = term_dihedral(t, own_lo, own_hi, tb.n[2], tb.idx[2], tb.par[2], inv_order, xs, bx, f_acc); slot = MDK_E_DIHEDRAL; break;
        case 3: e = term_improper(t, own_lo, own_hi, tb.n[3], tb.idx[3], tb.par[3], inv_order, xs, bx, f_acc); slot = MDK_E_IMPROPER; break;
        default: e = term_excl(t, own_lo, own_hi, tb.n[4], tb.pairs, inv_order, xs, bx, alpha, f_acc); slot = MDK_E_PME_EXCL; break;
    }
    block_energy(e, e_acc, slot);
}

// bonded terms of `terms` + (with MDK_TERM_PME_RECIP) the excluded-pair Ewald correction, own terms only
int bonded_compute(mdk_ctx *c, unsigned terms) {
    BoxF bx;
    for (int a = 0; a < 3; ++a) bx.L[a] = c->box.Ld[a];
    long long *e_acc = reinterpret_cast<long long *>(c->e_acc.p);
    PhaseTimer pt(c, PH_BONDED);
    static const unsigned bit[4] = {MDK_TERM_BOND, MDK_TERM_ANGLE, MDK_TERM_DIHEDRAL, MDK_TERM_IMPROPER};
    AuxTable tb{};
    int blocks = 0;
    for (int k = 0; k < 4; ++k) {
        const bool sel = c->dd && c->aux_sel_n[k] >= 0;
        tb.n[k] = (terms & bit[k]) ? (sel ? c->aux_sel_n[k] : c->bonded[k].n) : 0;
        tb.sel[k] = sel ? c->aux_sel[k].p : nullptr;
        tb.idx[k] = c->bonded[k].idx.p; tb.par[k] = c->bonded[k].par.p;
        tb.blk_off[k] = blocks;
        blocks += (tb.n[k] + AUX_T - 1) / AUX_T;
    }
    const bool sel4 = c->dd && c->aux_sel_n[4] >= 0;
    tb.n[4] = (terms & MDK_TERM_PME_RECIP) ? (sel4 ? c->aux_sel_n[4] : c->n_excl_pairs) : 0;
    tb.sel[4] = sel4 ? c->aux_sel[4].p : nullptr;
    tb.pairs = c->excl_pairs.p;
    tb.blk_off[4] = blocks;
    blocks += (tb.n[4] + AUX_T - 1) / AUX_T;
    tb.blk_off[5] = blocks;
    if (blocks == 0) return MDK_OK;
    const int lo = own_first(c), hi = c->own_hi < 0 ? c->n_pad : c->own_hi;
    PosSrc ps{c->xs.p, nullptr, nullptr, c->order.p, sqrt(c->k_e)};
    if (c->dprec && c->have_q64) { ps.x_cur = c->x_cur.p; ps.q64 = c->q64.p; }
    k_aux_terms<<<blocks, AUX_T, 0, c->stream>>>(tb, lo, hi, c->inv_order.p, ps, bx, c->alpha, c->f_acc.p, e_acc);
    ++c->n_launches;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

// ===========================================================================
// term dispatcher
// Device work of one force evaluation on a valid tile list: enqueue only, no host synchronisation,
// fixed launch shapes (this is what a CUDA-graph step captures).
// clean_on_entry: the accumulators are already zero (steps of a graph run: k_decide clears the energies,
// the Langevin update clears each force after consuming it), so the two memset nodes are left out.
int forces_enqueue(mdk_ctx *c, unsigned terms, bool clean_on_entry) {
    if (!clean_on_entry) {
        MDK_CUDA(c, cudaMemsetAsync(c->f_acc.p, 0, (size_t)c->n_pad * 3 * sizeof(long long), c->stream));
        MDK_CUDA(c, cudaMemsetAsync(c->e_acc.p, 0, MDK_NUM_ENERGIES * sizeof(long long), c->stream));
    }
    // (single domain; the domain-decomposed step has its own driver, mdk_dd.cu:dd_forces)
    const bool first = true, last = true;
    // Three independent chains, all adding into the same int64 accumulators (integer atomics commute,
    // so concurrency does not change a single bit): k_pair on the main stream, the PME mesh on s_pme,
    // the O(N) kernels on s_aux.  Per-phase profiling (level 2) serialises them on the main stream.
    const bool fork = c->concurrent && c->profiling < 2;
    const bool want_pme = (terms & MDK_TERM_PME_RECIP) && last;
    const unsigned bonded_bits = terms & (MDK_TERM_BOND | MDK_TERM_ANGLE | MDK_TERM_DIHEDRAL | MDK_TERM_IMPROPER);
    const bool want_aux = bonded_bits || (terms & MDK_TERM_PME_RECIP) || (first && (terms & MDK_TERM_COUL_BARE));
    cudaStream_t main_stream = c->stream;
    if (fork && (want_pme || want_aux)) MDK_CUDA(c, cudaEventRecord(c->ev_fork, main_stream));
    if (want_pme) {
        if (fork) { MDK_CUDA(c, cudaStreamWaitEvent(c->s_pme, c->ev_fork, 0)); c->stream = c->s_pme; }
        int rc = pme_compute(c);
        if (fork) { cudaEventRecord(c->ev_pme, c->s_pme); c->stream = main_stream; }
        MDK_TRY(rc);
    }
    if (want_aux) {
        if (fork) { MDK_CUDA(c, cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0)); c->stream = c->s_aux; }
        int rc = MDK_OK;
        if (first && (terms & MDK_TERM_COUL_BARE)) rc = coulomb_bare(c);
        if (rc == MDK_OK && (bonded_bits || (terms & MDK_TERM_PME_RECIP))) rc = bonded_compute(c, terms);
        if (fork) { cudaEventRecord(c->ev_aux, c->s_aux); c->stream = main_stream; }
        MDK_TRY(rc);
    }
    MDK_TRY(pair_compute(c, terms & MDK_TERM_LJ, terms & MDK_TERM_COUL_DIRECT));
    if (fork && want_pme) MDK_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_pme, 0));
    if (fork && want_aux) MDK_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_aux, 0));
    return MDK_OK;
}

int compute_terms(mdk_ctx *c, unsigned terms, bool sync_energies) {
    if (!c->have_box || c->n <= 0 || !c->have_pos)
        return fail(c, MDK_ERR_NOT_BOUND, "mdk_compute before box/atoms/positions were set");
    if (c->dd) return dd_compute_single(c, terms, sync_energies);
    if (!c->xs_current) MDK_TRY(nlist_refresh_sorted(c));
    MDK_TRY(nlist_ensure(c));
    c->xs_current = true;
    MDK_TRY(forces_enqueue(c, terms, false));
    if (sync_energies) {
        MDK_TRY(comm_allreduce_energies(c));
        long long h[MDK_NUM_ENERGIES];
        MDK_CUDA(c, cudaMemcpyAsync(h, c->e_acc.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < MDK_NUM_ENERGIES; ++k) c->last_e[k] = (double)h[k] / FIX_SCALE;
        if (terms & MDK_TERM_PME_RECIP) c->last_e[MDK_E_PME_SELF] = c->e_self_bg;
    }
    return MDK_OK;
}

// ===========================================================================
// Integrators (tile order: thread k handles atom order[k]).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// three N(0,1) variates for (atom, step): Philox4x32-10 bits, Box-Muller in float32 with the MUFU
// log / sin / cos (24 random bits per uniform: tails to 5.9 sigma, variate error ~1e-6 — far below
// what a thermostat can tell; the float64 version spent more time here than in the update itself)
__device__ __forceinline__ void normal3(uint64_t seed, uint32_t atom, uint64_t step, double xi[3]) {
    uint32_t r[4];
    philox4x32_10(atom, (uint32_t)step, (uint32_t)(step >> 32), 0x4d445059u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const float two_m24 = 5.9604644775390625e-08f;
    const float u0 = ((float)(r[0] >> 8) + 0.5f) * two_m24, u1 = ((float)(r[1] >> 8) + 0.5f) * two_m24;
    const float u2 = ((float)(r[2] >> 8) + 0.5f) * two_m24, u3 = ((float)(r[3] >> 8) + 0.5f) * two_m24;
    const float m0 = sqrtf(fmaxf(-2.0f * __logf(u0), 0.f)), m1 = sqrtf(fmaxf(-2.0f * __logf(u2), 0.f));   // lg2.approx may overshoot 0 by an ulp
    float s, cth;
    __sincosf(6.283185307179586f * u1, &s, &cth);
    xi[0] = (double)(m0 * cth); xi[1] = (double)(m0 * s);
    __sincosf(6.283185307179586f * u3, &s, &cth);
    xi[2] = (double)(m1 * cth);
}

struct StepGeom { double L[3]; float Lf[3], invLf[3]; float skin_half2; };

__device__ __forceinline__ void publish_position(int k, const double x[3], const double x_old[3], const StepGeom &g,
                                                 float4 *xs, const float4 *xs_ref, int *flags) {
    float w[3];
    bool lost = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double img = rint(x[d] / g.L[d]);
        w[d] = (float)(x[d] - g.L[d] * img);
        // a diverging trajectory: the reference wraps after every step and raises ParticleLossError when the new
        // coordinate is 2 or more images away, |round(x / L)| >= 2 (utils/pbc.py:29-34), i.e. when one step moved
        // the atom by 1.5 L or more; non-finite coordinates fail the same test
        if (!(fabs(x[d] - x_old[d]) < 1.5 * g.L[d])) lost = true;
    }
    if (lost) flags[0] = 1;
    float4 r = xs_ref[k];
    float dx = min_image(w[0] - r.x, g.Lf[0], g.invLf[0]);
    float dy = min_image(w[1] - r.y, g.Lf[1], g.invLf[1]);
    float dz = min_image(w[2] - r.z, g.Lf[2], g.invLf[2]);
    if (dist2(dx, dy, dz) > g.skin_half2) flags[1] = 1;
    xs[k] = make_float4(w[0], w[1], w[2], xs[k].w);
}

// Verlet.  mode 0: initialise x_prev from (x, v, a) then step; mode 1: step.
// Reference recurrences (verlet_integrator.py:28-43): x_prev = x - v dt + a dt^2 [quirk; textbook
// a dt^2/2], x_new = 2 x - x_prev + a dt^2.
__global__ void k_verlet(int n, int mode, int quirks, double dt, const int *__restrict__ order,
                         const float *__restrict__ mass, const long long *__restrict__ f_acc,
                         double *__restrict__ x_cur, double *__restrict__ x_prev, const double *__restrict__ vel,
                         StepGeom g, float4 *__restrict__ xs, const float4 *__restrict__ xs_ref,
                         int *__restrict__ flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int a = order[k];
    double inv_m = 1.0 / (double)mass[a];
    double xn[3], xo[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double acc = (double)f_acc[3 * (size_t)k + d] * (1.0 / FIX_SCALE) * inv_m;
        double xc = x_cur[3 * a + d];
        double xp = mode == 0 ? xc - vel[3 * a + d] * dt + (quirks ? 1.0 : 0.5) * acc * dt * dt : x_prev[3 * a + d];
        xn[d] = 2.0 * xc - xp + acc * dt * dt;
        xo[d] = xc;
        x_prev[3 * a + d] = xc;
        x_cur[3 * a + d] = xn[d];
    }
    publish_position(k, xn, xo, g, xs, xs_ref, flags);
}

// velocities at the end of VerletIntegrator.integrate (verlet_integrator.py:47-50).
// quirks: minimg(x_cur - x_prev) / (2 dt) (sic).  textbook: (x_cur - x_prev)/dt + a(x_cur) dt / 2.
__global__ void k_verlet_velocity(int n, int quirks, double dt, const int *__restrict__ order,
                                  const float *__restrict__ mass, const long long *__restrict__ f_acc,
                                  const double *__restrict__ x_cur, const double *__restrict__ x_prev,
                                  double *__restrict__ vel, StepGeom g) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int a = order[k];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double dx = x_cur[3 * a + d] - x_prev[3 * a + d];
        if (quirks) {
            dx -= g.L[d] * rint(dx / g.L[d]);
            vel[3 * a + d] = dx / (2.0 * dt);
        } else {
            double acc = (double)f_acc[3 * (size_t)k + d] * (1.0 / FIX_SCALE) / (double)mass[a];
            vel[3 * a + d] = dx / dt + 0.5 * acc * dt;
        }
    }
}

// G-JF Langevin (Gronbech-Jensen & Farago 2013) with the reference's a, b, sigma
// (langevin_integrator.py:23-31):
//   x' = x + b dt v + b dt^2/(2m) f + b dt/(2m) beta',   beta' = sqrt(2 gamma m kT dt) xi
//   v' = a v + dt/(2m) (a f + f') + b/m beta'
// Invariant between calls: x_cur = x_n, vel = v_n, f_prev = f(x_n).
// mode bit 0 (FINISH): f_acc holds f(x_n+1); complete v_n+1 with noise index step-1, set f_prev.
// mode bit 1 (ADVANCE): move x one step with noise index `step` using the newest force.
// mode bit 2 (FROM_PREV): the newest force is f_prev (start of a call on a cached state).
__global__ void k_langevin(int first, int n, int mode, double dt, double ca, double cb, double two_g_kT_dt,
                           uint64_t seed, uint64_t step, const unsigned long long *__restrict__ step_dev,
                           const unsigned long long *__restrict__ mode_dev, const int *__restrict__ order,
                           const float *__restrict__ mass, long long *__restrict__ f_acc,
                           double *__restrict__ x_cur, double *__restrict__ vel, double *__restrict__ f_prev,
                           StepGeom g, float4 *__restrict__ xs, const float4 *__restrict__ xs_ref,
                           int *__restrict__ flags, const unsigned char *__restrict__ rigid) {
    int k = first + blockIdx.x * blockDim.x + threadIdx.x;   // tile slots [first, n): the atoms this rank owns
    if (k >= n) return;
    if (rigid && rigid[order[k]]) return;                    // atoms of rigid waters: k_langevin_water
    // a host-state call that ran ahead of its own change check (mdk_step_langevin_host): the host
    // positions turned out to differ from the device's, the cached force is stale — leave the state alone
    if (flags[4] | flags[0]) return;
    if (step_dev) step = *step_dev;   // graph steps keep the noise counter ...
    if (mode_dev) mode = (int)*mode_dev;   // ... and the mode (3 inside a run, 1 for the last step of a call) on the device
    int a = order[k];
    double m = (double)mass[a], inv_m = 1.0 / m;
    double bs = sqrt(two_g_kT_dt * m);
    double f_new[3], v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        f_new[d] = (mode & 4) ? f_prev[3 * a + d] : (double)f_acc[3 * (size_t)k + d] * (1.0 / FIX_SCALE);
        v[d] = vel[3 * a + d];
    }
    if (mode & 1) {
        double xi[3];
        normal3(seed, (uint32_t)a, step - 1, xi);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            v[d] = ca * v[d] + 0.5 * dt * inv_m * (ca * f_prev[3 * a + d] + f_new[d]) + cb * inv_m * bs * xi[d];
            vel[3 * a + d] = v[d];
        }
    }
    if (!(mode & 4)) {
#pragma unroll
        for (int d = 0; d < 3; ++d) f_prev[3 * a + d] = f_new[d];
    }
    // another step follows (finish + advance): leave the accumulator clean for it; after the last step of
    // a call the forces stay readable (mdk_download_forces)
    if (mode == 3) {
#pragma unroll
        for (int d = 0; d < 3; ++d) f_acc[3 * (size_t)k + d] = 0;
    }
    if (mode & 2) {
        double xi[3], xn[3], xo[3];
        normal3(seed, (uint32_t)a, step, xi);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            xo[d] = x_cur[3 * a + d];
            xn[d] = xo[d] + cb * dt * v[d] + 0.5 * cb * dt * dt * inv_m * f_new[d] +
                    0.5 * cb * dt * inv_m * bs * xi[d];
            x_cur[3 * a + d] = xn[d];
        }
        publish_position(k, xn, xo, g, xs, xs_ref, flags);
    }
}

// ---------------------------------------------------------------------------
// Rigid three-site waters (the reference only carries the flag is_SHAKE,
// forcefield/charmm_forcefield.py:24,32): the same G-JF update as k_langevin for the three atoms of one molecule,
// with the constraints applied analytically — SETTLE (Miyamoto & Kollman 1992) resets the positions after the
// unconstrained move, and the velocities are projected onto the constraint surface (RATTLE's velocity stage).
// The displacement SETTLE applies, dx = x_c - x_unconstrained, is the constraint force's share b dt^2/(2m) g of the
// position update; the same g enters the next velocity update as dt/(2m) a g = a dx / (b dt), which is added to the
// stored velocity right here (the finish stage multiplies it by a), and the projection after the finish stage
// supplies the constraint force at the new positions.
struct WaterGeom { double d_oh, d_hh, m_o, m_h; };

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void unit3(double a[3]) {
    const double s = rsqrt(dot3(a, a));
    a[0] *= s; a[1] *= s; a[2] *= s;
}

// b: constrained positions at the start of the step, c: after the unconstrained update -> c is overwritten
__device__ void settle_positions(const double b[3][3], double c[3][3], const WaterGeom &w) {
    const double wohh = w.m_o + 2.0 * w.m_h, rc = 0.5 * w.d_hh;
    const double t = sqrt(w.d_oh * w.d_oh - rc * rc);
    const double ra = 2.0 * w.m_h * t / wohh, rb = t - ra;
    double com[3], xb0[3], xc0[3], xa1[3], xb1[3], xc1[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        com[d] = (w.m_o * c[0][d] + w.m_h * (c[1][d] + c[2][d])) / wohh;
        xb0[d] = b[1][d] - b[0][d]; xc0[d] = b[2][d] - b[0][d];
        xa1[d] = c[0][d] - com[d]; xb1[d] = c[1][d] - com[d]; xc1[d] = c[2][d] - com[d];
    }
    double xax[3], yax[3], zax[3];
    cross3(xb0, xc0, zax); cross3(xa1, zax, xax); cross3(zax, xax, yax);
    unit3(xax); unit3(yax); unit3(zax);
    const double b0x = dot3(xb0, xax), b0y = dot3(xb0, yax), c0x = dot3(xc0, xax), c0y = dot3(xc0, yax);
    const double a1z = dot3(xa1, zax);
    const double b1x = dot3(xb1, xax), b1y = dot3(xb1, yax), b1z = dot3(xb1, zax);
    const double c1x = dot3(xc1, xax), c1y = dot3(xc1, yax), c1z = dot3(xc1, zax);
    const double sinphi = a1z / ra, cosphi = sqrt(1.0 - sinphi * sinphi);
    const double sinpsi = (b1z - c1z) / (2.0 * rc * cosphi), cospsi = sqrt(1.0 - sinpsi * sinpsi);
    const double a2y = ra * cosphi, b2x = -rc * cospsi, t1 = -rb * cosphi, t2 = rc * sinpsi * sinphi;
    const double b2y = t1 - t2, c2y = t1 + t2;
    const double alpa = b2x * (b0x - c0x) + b0y * b2y + c0y * c2y;
    const double beta = b2x * (c0y - b0y) + b0x * b2y + c0x * c2y;
    const double gama = b0x * b1y - b1x * b0y + c0x * c1y - c1x * c0y;
    const double al2be2 = alpa * alpa + beta * beta;
    const double sinthe = (alpa * gama - beta * sqrt(al2be2 - gama * gama)) / al2be2, costhe = sqrt(1.0 - sinthe * sinthe);
    const double a3[3] = {-a2y * sinthe, a2y * costhe, a1z};
    const double b3[3] = {b2x * costhe - b2y * sinthe, b2x * sinthe + b2y * costhe, b1z};
    const double c3[3] = {-b2x * costhe - c2y * sinthe, -b2x * sinthe + c2y * costhe, c1z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        c[0][d] = com[d] + a3[0] * xax[d] + a3[1] * yax[d] + a3[2] * zax[d];
        c[1][d] = com[d] + b3[0] * xax[d] + b3[1] * yax[d] + b3[2] * zax[d];
        c[2][d] = com[d] + c3[0] * xax[d] + c3[1] * yax[d] + c3[2] * zax[d];
    }
}

// remove the velocity components along the three bonds with impulses along the bonds (3 x 3 system, Cramer)
__device__ void rattle_velocities(const double x[3][3], double v[3][3], const WaterGeom &w) {
    const double im[3] = {1.0 / w.m_o, 1.0 / w.m_h, 1.0 / w.m_h};
    const int pi[3] = {0, 0, 1}, pj[3] = {1, 2, 2};
    double e[3][3], A[3][3], rhs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) e[k][d] = x[pj[k]][d] - x[pi[k]][d];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double dv[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) dv[d] = v[pj[k]][d] - v[pi[k]][d];
        rhs[k] = -dot3(dv, e[k]);
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            // impulse lam_l along e_l: v_p += im_p lam e_l, v_q -= im_q lam e_l (p = pi[l], q = pj[l])
            double coef = 0.0;
            const double ee = dot3(e[l], e[k]);
            if (pi[l] == pj[k]) coef += im[pi[l]] * ee;
            if (pi[l] == pi[k]) coef -= im[pi[l]] * ee;
            if (pj[l] == pj[k]) coef -= im[pj[l]] * ee;
            if (pj[l] == pi[k]) coef += im[pj[l]] * ee;
            A[k][l] = coef;
        }
    }
    const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                       A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    const double l0 = (rhs[0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (rhs[1] * A[2][2] - A[1][2] * rhs[2]) +
                       A[0][2] * (rhs[1] * A[2][1] - A[1][1] * rhs[2])) / det;
    const double l1 = (A[0][0] * (rhs[1] * A[2][2] - A[1][2] * rhs[2]) - rhs[0] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                       A[0][2] * (A[1][0] * rhs[2] - rhs[1] * A[2][0])) / det;
    const double l2 = (A[0][0] * (A[1][1] * rhs[2] - rhs[1] * A[2][1]) - A[0][1] * (A[1][0] * rhs[2] - rhs[1] * A[2][0]) +
                       rhs[0] * (A[1][0] * A[2][1] - A[1][1] * A[2][0])) / det;
    const double lam[3] = {l0, l1, l2};
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int d = 0; d < 3; ++d) { v[pi[l]][d] += im[pi[l]] * lam[l] * e[l][d]; v[pj[l]][d] -= im[pj[l]] * lam[l] * e[l][d]; }
}

// mode as k_langevin; mode bit 3 (8): only project the current positions onto the rigid geometry and the velocities onto
// the constraint surface (called once when the constraints are switched on)
__global__ void k_langevin_water(int nw, const int *__restrict__ trip, int mode, double dt, double ca, double cb, double two_g_kT_dt,
                                 uint64_t seed, uint64_t step, const unsigned long long *__restrict__ step_dev,
                                 const unsigned long long *__restrict__ mode_dev, const int *__restrict__ inv_order,
                                 long long *__restrict__ f_acc, double *__restrict__ x_cur, double *__restrict__ vel,
                                 double *__restrict__ f_prev, StepGeom g, WaterGeom wg, float4 *__restrict__ xs,
                                 const float4 *__restrict__ xs_ref, int *__restrict__ flags) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nw) return;
    if (flags[4] | flags[0]) return;
    if (step_dev) step = *step_dev;
    if (mode_dev) mode = (int)*mode_dev;
    int a[3], s[3];
    double x[3][3], v[3][3], shift[3][3];
    const double m[3] = {wg.m_o, wg.m_h, wg.m_h};
#pragma unroll
    for (int t = 0; t < 3; ++t) { a[t] = trip[3 * w + t]; s[t] = inv_order[a[t]]; }
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            x[t][d] = x_cur[3 * a[t] + d]; v[t][d] = vel[3 * a[t] + d];
            // the molecule whole: hydrogens as the images next to their oxygen (the stored coordinates keep their own image)
            shift[t][d] = t == 0 ? 0.0 : g.L[d] * rint((x[t][d] - x[0][d]) / g.L[d]);
            x[t][d] -= shift[t][d];
        }
    if (mode & 8) {
        double c[3][3];
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int d = 0; d < 3; ++d) c[t][d] = x[t][d];
        settle_positions(x, c, wg);
        rattle_velocities(c, v, wg);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            double xn[3], xo[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) { xn[d] = c[t][d] + shift[t][d]; xo[d] = x[t][d] + shift[t][d]; x_cur[3 * a[t] + d] = xn[d]; vel[3 * a[t] + d] = v[t][d]; }
            if (xs) publish_position(s[t], xn, xo, g, xs, xs_ref, flags);
        }
        return;
    }
    double f_new[3][3];
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int d = 0; d < 3; ++d)
            f_new[t][d] = (mode & 4) ? f_prev[3 * a[t] + d] : (double)f_acc[3 * (size_t)s[t] + d] * (1.0 / FIX_SCALE);
    if (mode & 1) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            double xi[3];
            normal3(seed, (uint32_t)a[t], step - 1, xi);
            const double bs = sqrt(two_g_kT_dt * m[t]), inv_m = 1.0 / m[t];
#pragma unroll
            for (int d = 0; d < 3; ++d)
                v[t][d] = ca * v[t][d] + 0.5 * dt * inv_m * (ca * f_prev[3 * a[t] + d] + f_new[t][d]) + cb * inv_m * bs * xi[d];
        }
        rattle_velocities(x, v, wg);
    }
    if (!(mode & 4)) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int d = 0; d < 3; ++d) f_prev[3 * a[t] + d] = f_new[t][d];
    }
    if (mode == 3) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int d = 0; d < 3; ++d) f_acc[3 * (size_t)s[t] + d] = 0;
    }
    if (mode & 2) {
        double c[3][3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            double xi[3];
            normal3(seed, (uint32_t)a[t], step, xi);
            const double bs = sqrt(two_g_kT_dt * m[t]), inv_m = 1.0 / m[t];
#pragma unroll
            for (int d = 0; d < 3; ++d)
                c[t][d] = x[t][d] + cb * dt * v[t][d] + 0.5 * cb * dt * dt * inv_m * f_new[t][d] + 0.5 * cb * dt * inv_m * bs * xi[d];
        }
        double u[3][3];
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int d = 0; d < 3; ++d) u[t][d] = c[t][d];
        settle_positions(x, c, wg);
        const double inv_bdt = 1.0 / (cb * dt);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            double xn[3], xo[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                v[t][d] += (c[t][d] - u[t][d]) * inv_bdt;
                xn[d] = c[t][d] + shift[t][d]; xo[d] = x[t][d] + shift[t][d];
                x_cur[3 * a[t] + d] = xn[d];
            }
            publish_position(s[t], xn, xo, g, xs, xs_ref, flags);
        }
    }
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int d = 0; d < 3; ++d) vel[3 * a[t] + d] = v[t][d];
}

__global__ void k_kinetic(int n, const float *__restrict__ mass, const double *__restrict__ vel,
                          long long *__restrict__ e_acc) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (a < n) {
        double vx = vel[3 * a], vy = vel[3 * a + 1], vz = vel[3 * a + 2];
        e = 0.5 * (double)mass[a] * (vx * vx + vy * vy + vz * vz);
    }
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomic_add_fix(&e_acc[MDK_E_KINETIC], to_fix(e));
}

__global__ void k_kinetic_own(int first, int end, const int *__restrict__ order, const float *__restrict__ mass,
                              const double *__restrict__ vel, long long *__restrict__ e_acc) {
    int k = first + blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (k < end) {
        int a = order[k];
        double vx = vel[3 * a], vy = vel[3 * a + 1], vz = vel[3 * a + 2];
        e = 0.5 * (double)mass[a] * (vx * vx + vy * vy + vz * vz);
    }
    const long long v = warp_sum_ll(to_fix(e));
    if ((threadIdx.x & 31) == 0 && v != 0) atomic_add_fix(&e_acc[MDK_E_KINETIC], v);
}

StepGeom make_geom(mdk_ctx *c) {
    StepGeom g;
    for (int a = 0; a < 3; ++a) { g.L[a] = c->box.Ld[a]; g.Lf[a] = c->box.L[a]; g.invLf[a] = c->box.invL[a]; }
    g.skin_half2 = 0.25f * c->skin * c->skin;
    return g;
}

// the Langevin update kernels on the context stream: free atoms of tile slots [first, end), then the rigid waters
static void langevin_kernels(mdk_ctx *c, int first, int end, int mode, double dt, double ca, double cb, double tg, uint64_t seed,
                             uint64_t step, const unsigned long long *step_dev, const unsigned long long *mode_dev) {
    StepGeom g = make_geom(c);
    if (end > first)
        k_langevin<<<(end - first + 255) / 256, 256, 0, c->stream>>>(first, end, mode, dt, ca, cb, tg, seed, step, step_dev, mode_dev,
                                                                     c->order.p, c->mass.p, c->f_acc.p, c->x_cur.p, c->vel.p, c->f_prev.p,
                                                                     g, c->xs.p, c->xs_ref.p, c->flags.p, c->n_rigid ? c->rigid_flag.p : nullptr);
    ++c->n_launches;
    if (c->n_rigid > 0) {
        WaterGeom wg{c->rigid_d_oh, c->rigid_d_hh, c->rigid_m_o, c->rigid_m_h};
        k_langevin_water<<<(c->n_rigid + 127) / 128, 128, 0, c->stream>>>(c->n_rigid, c->rigid_trip.p, mode, dt, ca, cb, tg, seed, step, step_dev,
                                                                          mode_dev, c->inv_order.p, c->f_acc.p, c->x_cur.p, c->vel.p,
                                                                          c->f_prev.p, g, wg, c->xs.p, c->xs_ref.p, c->flags.p);
        ++c->n_launches;
    }
}

// one Langevin update of tile slots [first, end) on the context stream (the domain-decomposed step, mdk_dd.cu)
int langevin_launch(mdk_ctx *c, int first, int end, int mode, double dt, double ca, double cb, double tg, uint64_t seed, uint64_t step) {
    PhaseTimer pt(c, PH_INTEGRATE);
    langevin_kernels(c, first, end, mode, dt, ca, cb, tg, seed, step, nullptr, nullptr);
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

// Switching the water constraints on: the current geometry is projected onto the rigid one (positions along their own
// bond directions, velocities onto the constraint surface) so that SETTLE's starting point satisfies the constraints.
int rigid_project(mdk_ctx *c) {
    if (c->n_rigid <= 0 || !c->have_pos) return MDK_OK;
    StepGeom g = make_geom(c);
    WaterGeom wg{c->rigid_d_oh, c->rigid_d_hh, c->rigid_m_o, c->rigid_m_h};
    k_langevin_water<<<(c->n_rigid + 127) / 128, 128, 0, c->stream>>>(c->n_rigid, c->rigid_trip.p, 8, 0.0, 0.0, 0.0, 0.0, 0ull, 0ull, nullptr,
                                                                      nullptr, c->rigid_trip.p /* unused */, nullptr, c->x_cur.p, c->vel.p,
                                                                      nullptr, g, wg, nullptr, nullptr, c->flags.p);
    ++c->n_launches;
    c->xs_current = false;
    c->verlet_cached = false; c->langevin_cached = false;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

int energies_enqueue(mdk_ctx *c) {
    long long *e_acc = reinterpret_cast<long long *>(c->e_acc.p);
    if (c->dd) {     // every rank: its own atoms; the sum is taken with the other energies
        const int f = own_first(c), e = own_end(c);
        if (e > f) k_kinetic_own<<<(e - f + 255) / 256, 256, 0, c->stream>>>(f, e, c->order.p, c->mass.p, c->vel.p, e_acc);
    } else {
        k_kinetic<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->mass.p, c->vel.p, e_acc);
    }
    ++c->n_launches;
    MDK_TRY(comm_allreduce_energies(c));
    // energies, list counters and flags in one 224-byte copy
    MDK_CUDA(c, cudaMemcpyAsync(c->pin_words, c->readback.p, 28 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    return MDK_OK;
}

void energies_finish(mdk_ctx *c, unsigned terms) {
    for (int k = 0; k < MDK_NUM_ENERGIES; ++k) c->last_e[k] = (double)c->pin_words[k] / FIX_SCALE;
    if (terms & MDK_TERM_PME_RECIP) c->last_e[MDK_E_PME_SELF] = c->e_self_bg;
}

// flags[0] raised by an integrator kernel: the state is unusable from here on
int check_lost_flag(mdk_ctx *c) {
    const int *h_flags = reinterpret_cast<const int *>(c->pin_words + 24);
    if (!h_flags[0]) return MDK_OK;
    cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream);
    c->have_pos = false;
    c->verlet_cached = false; c->langevin_cached = false;
    c->nlist_valid = false; c->xs_current = false;
    return fail(c, MDK_ERR_PARTICLE_LOST, "Atom(s) moved beyond 2 PBC image.");
}

static int fetch_energies(mdk_ctx *c, unsigned terms) {
    MDK_TRY(energies_enqueue(c));
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    energies_finish(c, terms);
    return check_lost_flag(c);
}

int integrate_verlet(mdk_ctx *c, double dt, int nsteps, unsigned terms, int quirks) {
    if (nsteps <= 0) return MDK_OK;
    if (c->dd) return fail(c, MDK_ERR_BAD_ARG, "the Verlet integrator is single domain; domain-decomposed runs use the Langevin step");
    if (c->n_rigid > 0) return fail(c, MDK_ERR_BAD_ARG, "rigid waters are integrated by the Langevin step only");
    if (terms != c->cached_terms) { c->verlet_cached = false; c->langevin_cached = false; c->cached_terms = terms; }
    const int n = c->n, T = 256, B = (n + T - 1) / T;
    StepGeom g = make_geom(c);
    MDK_CUDA(c, c->x_prev.reserve((size_t)3 * n));
    for (int s = 0; s < nsteps; ++s) {
        MDK_TRY(compute_terms(c, terms, false));
        PhaseTimer pt(c, PH_INTEGRATE);
        int mode = c->verlet_cached ? 1 : 0;
        k_verlet<<<B, T, 0, c->stream>>>(n, mode, quirks, dt, c->order.p, c->mass.p, c->f_acc.p, c->x_cur.p,
                                         c->x_prev.p, c->vel.p, g, c->xs.p, c->xs_ref.p, c->flags.p);
        ++c->n_launches;
        c->verlet_cached = true;
    }
    if (!quirks) MDK_TRY(compute_terms(c, terms, false));  // a(x_cur) for the velocity
    k_verlet_velocity<<<B, T, 0, c->stream>>>(n, quirks, dt, c->order.p, c->mass.p, c->f_acc.p, c->x_cur.p,
                                              c->x_prev.p, c->vel.p, g);
    ++c->n_launches;
    MDK_CUDA(c, cudaGetLastError());
    return fetch_energies(c, terms);
}

// ===========================================================================
// Trajectory frames (the on-disk side of the loop; reference dumpers: mdpy/dumper/*.py call
// ensemble.state.positions once per dump period from the host).  With a capture stride set, a step call copies the
// wrapped float32 positions of every stride-th step into a page-locked ring on the host: a small kernel on the step
// stream writes them to one of two staging buffers, a copy stream moves them out while the next steps run.
__global__ void k_frame(int n, const double *__restrict__ x_cur, double Lx, double Ly, double Lz, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    const int d = i % 3;
    const double L = d == 0 ? Lx : (d == 1 ? Ly : Lz), x = x_cur[i];
    out[i] = (float)(x - L * rint(x / L));
}

int frame_capture_enqueue(mdk_ctx *c, int step_in_call) {
    if (c->frame_stride <= 0 || step_in_call % c->frame_stride != 0 || c->frame_count >= c->frame_cap) return MDK_OK;
    const int slot = c->frame_count & 1;
    const size_t m = (size_t)3 * c->n;
    MDK_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_frame_done[slot], 0));          // the copy that last used this staging buffer
    k_frame<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(c->n, c->x_cur.p, c->box.Ld[0], c->box.Ld[1], c->box.Ld[2],
                                                               c->frame_dev.p + slot * m);
    ++c->n_launches;
    MDK_CUDA(c, cudaEventRecord(c->ev_frame_ready[slot], c->stream));
    MDK_CUDA(c, cudaStreamWaitEvent(c->s_io, c->ev_frame_ready[slot], 0));
    MDK_CUDA(c, cudaMemcpyAsync(c->frame_host + (size_t)c->frame_count * m, c->frame_dev.p + slot * m, m * sizeof(float),
                                cudaMemcpyDeviceToHost, c->s_io));
    MDK_CUDA(c, cudaEventRecord(c->ev_frame_done[slot], c->s_io));
    ++c->frame_count; ++c->frame_total;
    return MDK_OK;
}

// ===========================================================================
// Steepest descent (mdpy/minimizer/steepest_descent_minimizer.py:30-53): every atom moves a fixed
// length alpha along ITS OWN unit force vector, x_i += alpha F_i / |F_i| (the reference normalises per atom,
// np.linalg.norm(forces, axis=1)); the loop stops when the relative change of the potential energy between two
// iterations drops under the tolerance.  An atom with zero force stays where it is (the reference would divide 0 / 0).
__global__ void k_sd_step(int n, double alpha, const int *__restrict__ order, const long long *__restrict__ f_acc,
                          double *__restrict__ x_cur, StepGeom g, float4 *__restrict__ xs, const float4 *__restrict__ xs_ref,
                          int *__restrict__ flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int a = order[k];
    double f[3], xo[3], xn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { f[d] = (double)f_acc[3 * (size_t)k + d] * (1.0 / FIX_SCALE); xo[d] = x_cur[3 * a + d]; }
    const double norm = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    const double s = norm > 0.0 ? alpha / norm : 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) { xn[d] = xo[d] + s * f[d]; x_cur[3 * a + d] = xn[d]; }
    publish_position(k, xn, xo, g, xs, xs_ref, flags);
}

int minimize_sd(mdk_ctx *c, double alpha, double energy_tolerance, int max_iterations, unsigned terms, int *iterations,
                double *e_first, double *e_prev, double *e_last) {
    if (c->dd) return fail(c, MDK_ERR_BAD_ARG, "the steepest-descent minimizer is single domain");
    const int n = c->n, T = 256, B = (n + T - 1) / T;
    auto potential = [&]() {
        double e = 0;
        for (int k = 0; k < MDK_E_KINETIC; ++k) e += c->last_e[k];
        return e;
    };
    MDK_TRY(compute_terms(c, terms, true));
    double cur = potential(), pre = cur;
    *e_first = cur; *e_prev = cur; *e_last = cur;
    int it = 0;
    c->verlet_cached = false; c->langevin_cached = false;
    while (it < max_iterations) {
        StepGeom g = make_geom(c);
        k_sd_step<<<B, T, 0, c->stream>>>(n, alpha, c->order.p, c->f_acc.p, c->x_cur.p, g, c->xs.p, c->xs_ref.p, c->flags.p);
        ++c->n_launches;
        MDK_TRY(compute_terms(c, terms, true));      // xs was published by the step: no refresh pass; rebuilds when flagged
        cur = potential();
        ++it;
        *e_prev = pre; *e_last = cur;
        if (!(fabs(cur) < 1e300)) return fail(c, MDK_ERR_PARTICLE_LOST, "steepest descent diverged (non-finite energy)");
        if (fabs((cur - pre) / pre) < energy_tolerance) break;
        pre = cur;
    }
    *iterations = it;
    return MDK_OK;
}

// ===========================================================================
// CUDA-graph step.  At 23 k atoms a step is ~25 kernels of 3-60 us: issuing them one by one, with a
// host round trip per step to learn whether the list must be rebuilt, costs more than executing
// them.  The steady-state Langevin step is therefore captured once into a graph whose rebuild
// branch is a conditional (IF) node driven from the device: k_decide reads the skin/2 flag the
// previous step's position update raised and arms the branch; its body is the whole list rebuild.
// ... and does the per-step housekeeping that would otherwise be graph nodes of their own: advance the
// device-side noise counter, clear the energy accumulators for the step that follows.
__global__ void k_decide(cudaGraphConditionalHandle handle, const int *__restrict__ flags,
                         unsigned long long *__restrict__ step_dev, long long *__restrict__ e_acc) {
    cudaGraphSetConditional(handle, flags[1] != 0 ? 1u : 0u);
    *step_dev += 1ull;
#pragma unroll
    for (int k = 0; k < MDK_NUM_ENERGIES; ++k) e_acc[k] = 0;
}

void graph_destroy(mdk_ctx *c) {
    for (int v = 0; v < 2; ++v) {
        if (c->step_exec[v]) cudaGraphExecDestroy(c->step_exec[v]);
        if (c->step_graph[v]) cudaGraphDestroy(c->step_graph[v]);
        c->step_exec[v] = nullptr; c->step_graph[v] = nullptr;
    }
    if (c->upkeep_exec) cudaGraphExecDestroy(c->upkeep_exec);
    if (c->upkeep_graph) cudaGraphDestroy(c->upkeep_graph);
    c->upkeep_exec = nullptr; c->upkeep_graph = nullptr;
}

#define CAP(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess && rc == MDK_OK)                                                            \
            rc = fail(c, MDK_ERR_CUDA, "%s failed while capturing the step graph: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

// Graph 1, "upkeep": k_decide -> IF(rebuild needed) { whole list rebuild }.  Kept separate from the
// force/update graph: a graph that contains a conditional node is executed without branch
// concurrency (measured: the PME / bonded branches stopped overlapping k_pair), so the conditional
// lives in its own two-node graph.
static int graph_build_upkeep(mdk_ctx *c) {
    cudaStream_t s = c->stream;
    int rc = MDK_OK;
    cudaGraph_t graph = nullptr;
    CAP(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    if (rc != MDK_OK) return rc;
    cudaStreamCaptureStatus status;
    unsigned long long id = 0;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    CAP(cudaStreamGetCaptureInfo_v2(s, &status, &id, &graph, &deps, &ndeps));
    cudaGraphConditionalHandle handle;
    CAP(cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault));
    k_decide<<<1, 1, 0, s>>>(handle, c->flags.p, c->step_dev.p, reinterpret_cast<long long *>(c->e_acc.p));
    CAP(cudaStreamGetCaptureInfo_v2(s, &status, &id, &graph, &deps, &ndeps));
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeIf;
    cp.conditional.size = 1;
    cudaGraphNode_t cond = nullptr;
    CAP(cudaGraphAddNode(&cond, graph, deps, ndeps, &cp));
    if (rc == MDK_OK) {
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        // the rebuild, captured into the branch body on a side stream
        CAP(cudaStreamBeginCaptureToGraph(c->s_aux, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
        if (rc == MDK_OK) {
            c->stream = c->s_aux;
            int r2 = nlist_enqueue(c, true);
            c->stream = s;
            cudaGraph_t dummy = nullptr;
            CAP(cudaStreamEndCapture(c->s_aux, &dummy));
            if (r2 != MDK_OK && rc == MDK_OK) rc = r2;
        }
        CAP(cudaStreamUpdateCaptureDependencies(s, &cond, 1, cudaStreamSetCaptureDependencies));
    }
    cudaGraph_t done = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &done);
    if (e != cudaSuccess && rc == MDK_OK) rc = fail(c, MDK_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    if (rc != MDK_OK) { if (done) cudaGraphDestroy(done); cudaGetLastError(); return rc; }
    c->upkeep_graph = done;
    e = cudaGraphInstantiate(&c->upkeep_exec, done, 0);
    if (e != cudaSuccess) return fail(c, MDK_ERR_CUDA, "cudaGraphInstantiate(upkeep): %s", cudaGetErrorString(e));
    return MDK_OK;
}

// Graph 2, "step": forces (three concurrent branches) -> Langevin update (mode read from the device:
// finish + advance inside a run, finish only for the last step of a call) -> step counter.
// variant 1 carries the energy sums in the pair kernel (the step whose energies are read back).
static int graph_build_step(mdk_ctx *c, int variant, double dt, double ca, double cb, double tg, uint64_t seed,
                            unsigned terms) {
    const int n = c->n, T = 256, B = (n + T - 1) / T;
    StepGeom g = make_geom(c);
    cudaStream_t s = c->stream;
    c->in_capture = true;
    c->capture_energy = variant == 1;
    const int64_t launches_before = c->n_launches, pair_before = c->n_pair_launches;
    int rc = MDK_OK;
    cudaGraph_t graph = nullptr;
    CAP(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    if (rc == MDK_OK) {
        rc = forces_enqueue(c, terms, true);
        if (rc == MDK_OK) langevin_kernels(c, 0, n, 3, dt, ca, cb, tg, seed, 0ull, c->step_dev.p, c->step_dev.p + 1);
        cudaError_t e = cudaStreamEndCapture(s, &graph);
        if (e != cudaSuccess && rc == MDK_OK) rc = fail(c, MDK_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    }
    c->in_capture = false;
    c->capture_energy = false;
    c->stream = s;
    c->graph_launches_per_step = (int)(c->n_launches - launches_before) + 2;   // + decide, langevin
    c->n_launches = launches_before; c->n_pair_launches = pair_before;         // capturing is not launching
    if (rc != MDK_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
    c->step_graph[variant] = graph;
    cudaError_t e = cudaGraphInstantiate(&c->step_exec[variant], graph, 0);
    if (e != cudaSuccess) return fail(c, MDK_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    return MDK_OK;
}
#undef CAP

__global__ void k_set_word(unsigned long long *p, unsigned long long v) { *p = v; }

// nsteps force evaluations as graph launches: nsteps - 1 steady-state steps (finish v of the pending
// step, advance x) and a last one that only finishes v and carries the energies.
//
// hosted = true (multi-GPU): only the list upkeep (k_decide -> IF { rebuild }) is a graph; the force
// kernels, the NCCL all-reduce and the Langevin update are launched from the host, still without a
// single host synchronisation inside the run (the rebuild decision stays on the device).
static int graph_run_langevin(mdk_ctx *c, double dt, double ca, double cb, double tg, uint64_t seed, unsigned terms,
                              int nsteps, bool hosted) {
    if (!c->graph_pools) {           // first use: re-plan the pools with graph slack, on the host
        c->graph_pools = true;
        c->nlist_valid = false;
        ++c->graph_epoch;
    }
    // a valid list needs no host round trip here: the upkeep graph reads the skin/2 flag on the device
    if (!c->xs_current) MDK_TRY(nlist_refresh_sorted(c));
    if (!c->nlist_valid) MDK_TRY(nlist_ensure(c));
    c->xs_current = true;
    const bool stale = c->graph_epoch_built != c->graph_epoch || c->graph_key[0] != dt || c->graph_key[1] != tg ||
                       c->graph_seed != seed || c->graph_key[3] != (double)terms || c->graph_key[4] != ca;
    if (stale) {
        graph_destroy(c);
        if (terms & MDK_TERM_PME_RECIP) MDK_TRY(pme_prepare(c));
        MDK_CUDA(c, c->step_dev.reserve(2));
        c->in_capture = true;
        int rc = graph_build_upkeep(c);
        c->in_capture = false;
        if (rc != MDK_OK) { graph_destroy(c); c->graph_epoch_built = -1; return rc; }
        c->graph_key[0] = dt; c->graph_key[1] = tg; c->graph_seed = seed; c->graph_key[3] = (double)terms;
        c->graph_key[4] = ca; c->graph_epoch_built = c->graph_epoch;
    }
    // the steps of a run find the force accumulator clean: zeroed here once, then by every Langevin update
    // that is followed by another step
    MDK_CUDA(c, cudaMemsetAsync(c->f_acc.p, 0, (size_t)c->n_pad * 3 * sizeof(long long), c->stream));
    if (hosted) {
        const int n = c->n, T = 256, B = (n + T - 1) / T;
        StepGeom g = make_geom(c);
        for (int s = 0; s < nsteps; ++s) {
            const bool last = s + 1 == nsteps;
            MDK_CUDA(c, cudaGraphLaunch(c->upkeep_exec, c->stream));
            MDK_TRY(forces_enqueue(c, terms, true));
            langevin_kernels(c, 0, n, last ? 1 : 3, dt, ca, cb, tg, seed, c->langevin_step + (uint64_t)s, nullptr, nullptr);
            c->n_launches += 1;
        }
        c->n_pair_launches += nsteps;
        c->graph_pending = nsteps;
        c->graph_pending_hosted = true;
        return MDK_OK;
    }
    for (int v = 0; v < 2; ++v) {
        const bool need = v == 1 || nsteps > 1;
        if (need && !c->step_exec[v]) {
            int rc = graph_build_step(c, v, dt, ca, cb, tg, seed, terms);
            if (rc != MDK_OK) { graph_destroy(c); c->graph_epoch_built = -1; return rc; }
        }
    }
    unsigned long long h_words[2] = {c->langevin_step - 1ull, nsteps == 1 ? 1ull : 3ull};   // k_decide ticks before each step
    MDK_CUDA(c, cudaMemcpyAsync(c->step_dev.p, h_words, sizeof(h_words), cudaMemcpyHostToDevice, c->stream));
    for (int s = 0; s < nsteps; ++s) {
        const bool last = s + 1 == nsteps;
        MDK_TRY(frame_capture_enqueue(c, s + 1));     // x_cur holds the positions after s + 1 steps of this call
        if (last && nsteps > 1) { k_set_word<<<1, 1, 0, c->stream>>>(c->step_dev.p + 1, 1ull); ++c->n_launches; }
        MDK_CUDA(c, cudaGraphLaunch(c->upkeep_exec, c->stream));
        MDK_CUDA(c, cudaGraphLaunch(c->step_exec[last ? 1 : 0], c->stream));
    }
    c->graph_pending = nsteps;   // counters / flags come back with the energies (energies_enqueue)
    c->graph_pending_hosted = false;
    return MDK_OK;
}

// bookkeeping of a graph run once the stream has been synchronised
int graph_finish(mdk_ctx *c) {
    const int nsteps = c->graph_pending;
    if (nsteps <= 0) return check_lost_flag(c);
    c->graph_pending = 0;
    const int *h_after = reinterpret_cast<const int *>(c->pin_words + 16);
    const int *h_flags = reinterpret_cast<const int *>(c->pin_words + 24);
    c->langevin_step += (uint64_t)(nsteps - 1);
    const int rebuilt = h_after[12] - c->rebuilds_seen;
    c->rebuilds_seen = h_after[12];
    c->n_rebuilds += rebuilt;
    if (rebuilt > 0) { c->stat_units = h_after[13]; c->stat_chunks = h_after[14]; c->stat_masks = h_after[15]; }
    c->n_launches += (int64_t)rebuilt * 10;
    if (!c->graph_pending_hosted) {
        c->n_launches += (int64_t)nsteps * c->graph_launches_per_step;
        c->n_pair_launches += nsteps;
    }
    MDK_TRY(check_lost_flag(c));
    if (h_flags[3] & 1) return fail(c, MDK_ERR_OOM, "tile-list pool overflow inside a graph step (raise the pools: more atoms per box than planned)");
    return MDK_OK;
}

int integrate_langevin(mdk_ctx *c, double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms,
                       int graph_min_steps, bool defer_energies) {
    if (nsteps <= 0) return MDK_OK;
    if (terms != c->cached_terms) { c->verlet_cached = false; c->langevin_cached = false; c->cached_terms = terms; }  // f_prev belongs to another term set
    const int n = c->n, T = 256, B = (n + T - 1) / T;
    StepGeom g = make_geom(c);
    MDK_CUDA(c, c->f_prev.reserve((size_t)3 * n));
    const double ca = (1.0 - 0.5 * gamma * dt) / (1.0 + 0.5 * gamma * dt), cb = 1.0 / (1.0 + 0.5 * gamma * dt);
    const double tg = 2.0 * gamma * kT * dt;
#define LANGEVIN(mode)                                                                                          \
    do {                                                                                                        \
        PhaseTimer pt(c, PH_INTEGRATE);                                                                         \
        langevin_kernels(c, 0, n, (mode), dt, ca, cb, tg, seed, c->langevin_step, nullptr, nullptr);            \
    } while (0)
    // multi-GPU: plain host-launched steps unless one of the two graph flavours is switched on — `hosted`
    // (upkeep graph + host-launched forces / NCCL / update, no host sync inside a run) or `graph_nccl`
    // (NCCL inside the captured step; hung in round 1)
    if (c->dd) return dd_langevin_single(c, dt, kT, gamma, seed, nsteps, terms, defer_energies);
    const bool use_graph = c->use_graph && c->profiling < 2 && nsteps >= graph_min_steps;
    const bool hosted = false;
    if (!c->langevin_cached) {
        MDK_TRY(compute_terms(c, terms, false));  // f(x_0)
        LANGEVIN(2);
        c->langevin_cached = true;
    } else {
        // the cached force is f(x_n): no list needed to advance; the position update publishes the
        // tile-order copy itself when a list exists
        if (!c->nlist_valid || !use_graph) {
            MDK_TRY(nlist_refresh_sorted(c));
            MDK_TRY(nlist_ensure(c));
        }
        LANGEVIN(2 | 4);
        c->xs_current = c->nlist_valid;
    }
    ++c->langevin_step;
    if (use_graph) {
        MDK_TRY(graph_run_langevin(c, dt, ca, cb, tg, seed, terms, nsteps, hosted));
    } else {
        for (int s = 0; s < nsteps; ++s) {
            MDK_TRY(frame_capture_enqueue(c, s + 1));
            MDK_TRY(compute_terms(c, terms, false));  // f(x_n+1)
            const bool more = s + 1 < nsteps;
            LANGEVIN(more ? 3 : 1);
            if (more) ++c->langevin_step;
        }
    }
#undef LANGEVIN
    MDK_CUDA(c, cudaGetLastError());
    MDK_TRY(energies_enqueue(c));
    if (defer_energies) return MDK_OK;     // the caller synchronises once, after queueing its downloads
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    energies_finish(c, terms);
    return graph_finish(c);
}

}  // namespace mdk
