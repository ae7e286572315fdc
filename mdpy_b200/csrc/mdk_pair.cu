// mdk_pair.cu — pair-force kernel over the tile list: CHARMM Lennard-Jones (plain cutoff
// or energy switch) and erfc direct-space Coulomb, half shell, fixed-point accumulation.
//
// Replaces CharmmNonbondedConstraint.cuda_kernel (charmm_nonbonded_constraint.py:110-181:
// one thread per (atom, cell slot), O(B) exclusion scans, 7 global float atomics per pair)
// and supplies the erfc direct-space term the reference does not have.
//
// Mapping.  One warp per work unit (i-block, up to seg chunks of 32 j-atoms).  Lane l owns
// i-atom l for the whole unit (force in registers).  Each chunk's 32 j-atoms are staged in
// shared memory (coalesced float4 gathers); at rotation step k lane l evaluates the pair
// (i = l, j-slot = (l + k) & 31), so the 32 lanes always touch 32 distinct j-slots: the
// j-force accumulators travel with the j-slot through a 3-register shuffle ring and are
// flushed with one int64 atomic per component per j-atom per chunk.  Every pair is
// evaluated once (Newton's third law), against the reference's twice-at-half-weight.
#include "mdk_common.cuh"

namespace mdk {

struct PairParams {
    float L[3], invL[3];
    float rc2_lj, ron2, inv_ab3, rc2_max;  // CHARMM switch: (rc^2 - ron^2)^-3
    float rc2_c, alpha, two_alpha_over_sqrtpi;
    float sw_c0, sw_12inv;      // rc^2 - 3 ron^2, 12 (rc^2 - ron^2)^-3
    float sw_s0, sw_s1;         // S = da^2 (sw_s0 + sw_s1 r^2): the two above times (rc^2 - ron^2)^-3
    float alpha04;              // 0.4 alpha
    float alpha2_log2e;         // alpha^2 log2(e): exp(-alpha^2 r^2) = ex2(-alpha2_log2e r^2)
    // SHIFT kernels: r^2 computed on hoisted images differs from the canonical r^2 by a few ulp of the box
    // length; slots whose r^2 falls inside [lo, hi] around a cutoff are re-decided on the canonical
    // expression (a handful per step), so the pair set is the canonical one bit for bit
    float lj_lo, lj_hi, c_lo, c_hi, max_lo, max_hi;
    int n;
    int max_units;              // work units a warp takes before it retires (0: until the cursor runs out — persistent blocks)
};

// debug instantiation of k_pair: every slot that passes the kernel's own cutoff / exclusion decision is
// written out (mdk_get_pairs with production != 0)
struct EmitOut {
    int *out_i, *out_j;
    const int *order;
    unsigned long long cap;
    unsigned long long *count;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// MUFU.RSQ / MUFU.RCP without the denormal pre-scaling nvcc wraps around rsqrtf / __fdividef: the
// arguments here are r^2 of a pair inside the cutoff and 1 + 0.4 alpha r — never denormal.
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// erfc(x) * exp(x^2) ~= t * P(t), t = 1 / (1 + 0.4 x): degree-8 least-squares fit on
// x in [0, 4.2], max relative error 8e-9 in exact arithmetic, <4e-7 evaluated in fp32
// (fit script: oracle/fit_erfc.py).
// r04a = 0.4 alpha (one FFMA for the argument): x = alpha r
__device__ __forceinline__ float erfcx_poly_r(float r, float r04a) {
    float t = rcp_approx(__fmaf_rn(r04a, r, 1.0f));
    float p = 1.2938003984e-02f;
    p = __fmaf_rn(p, t, 8.5587749120e-03f);
    p = __fmaf_rn(p, t, -2.5232078617e-01f);
    p = __fmaf_rn(p, t, 5.1394868547e-01f);
    p = __fmaf_rn(p, t, -2.5095656442e-01f);
    p = __fmaf_rn(p, t, 3.5446957253e-01f);
    p = __fmaf_rn(p, t, 1.5382895656e-01f);
    p = __fmaf_rn(p, t, 2.3447514612e-01f);
    p = __fmaf_rn(p, t, 2.2505821879e-01f);
    return p * t;
}
__device__ __forceinline__ float erfcx_poly(float x) {
    float t = rcp_approx(__fmaf_rn(0.4f, x, 1.0f));
    float p = 1.2938003984e-02f;
    p = __fmaf_rn(p, t, 8.5587749120e-03f);
    p = __fmaf_rn(p, t, -2.5232078617e-01f);
    p = __fmaf_rn(p, t, 5.1394868547e-01f);
    p = __fmaf_rn(p, t, -2.5095656442e-01f);
    p = __fmaf_rn(p, t, 3.5446957253e-01f);
    p = __fmaf_rn(p, t, 1.5382895656e-01f);
    p = __fmaf_rn(p, t, 2.3447514612e-01f);
    p = __fmaf_rn(p, t, 2.2505821879e-01f);
    return p * t;
}

#ifndef MDK_PAIR_WARPS
#define MDK_PAIR_WARPS 8
#endif
#ifndef MDK_PAIR_UNROLL
#define MDK_PAIR_UNROLL 8
#endif
#define MDK_PRAGMA_(x) _Pragma(#x)
#define MDK_UNROLL(n) MDK_PRAGMA_(unroll n)
constexpr int PAIR_WARPS = MDK_PAIR_WARPS;

// One chunk = 32 j-atoms against the warp's 32 i-atoms, 32 rotation steps.  MASKED chunks carry
// exclusion / 1-4 bits (a few per i-block); the rest skip the bit tests entirely.
// r^2 of one pair by the canonical criterion (oracle/mdpy_oracle.c:ora_pair_set_f32) from the stored
// wrapped coordinates — the rare slow path of the SHIFT kernels
__device__ __noinline__ float canonical_r2(float Lx, float Ly, float Lz, float ix, float iy, float iz,
                                           const float4 *__restrict__ xs, int ia, int ja) {
    const float4 a = xs[ia], b = xs[ja];
    return dist2(min_image(b.x - a.x, Lx, ix), min_image(b.y - a.y, Ly, iy), min_image(b.z - a.z, Lz, iz));
}

template <bool DO_LJ, bool DO_COUL, bool SWITCH, bool SHIFT, bool ONECUT, bool ENERGY, bool MASKED, bool EMIT>
__device__ __forceinline__ void chunk_loop(const PairParams &P, const float4 *__restrict__ sx,
                                           const float4 *__restrict__ slj, const int lane, const float4 xi,
                                           const float4 li, const unsigned excl, const unsigned m14, float &fix,
                                           float &fiy, float &fiz, float &fjx, float &fjy, float &fjz, float &e_lj,
                                           float &e_c, const float4 *__restrict__ xs, const int *__restrict__ cj,
                                           const int ia, const EmitOut &em) {
    // the chunk is staged twice back to back (64 entries), so slot (lane + k) & 31 is entry lane + k: one
    // base address per lane, the rotation step is an immediate offset of the LDS
    sx += lane; slj += lane;
    MDK_UNROLL(MDK_PAIR_UNROLL)
    for (int k = 0; k < 32; ++k) {
        const float4 xj = sx[k];
        float dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
        if (!SHIFT) {
            dx = min_image(dx, P.L[0], P.invL[0]);
            dy = min_image(dy, P.L[1], P.invL[1]);
            dz = min_image(dz, P.L[2], P.invL[2]);
        }
        const float r2 = dist2(dx, dy, dz);
        bool in = r2 <= (SHIFT ? P.max_hi : P.rc2_max);
        if (MASKED) in = in && !((excl >> k) & 1u);
        float r2d = r2;   // the value the cutoff decisions are taken on
        if (SHIFT && in) {
            bool near = r2 > P.max_lo;
            if (!ONECUT) near = near || (r2 > P.lj_lo && r2 <= P.lj_hi) || (r2 > P.c_lo && r2 <= P.c_hi);
            if (near) {
                r2d = canonical_r2(P.L[0], P.L[1], P.L[2], P.invL[0], P.invL[1], P.invL[2], xs, ia, cj[(lane + k) & 31]);
                in = r2d <= P.rc2_max;
            }
        }
        if (in) {
            if (EMIT) {
                const unsigned long long pos = atomicAdd(em.count, 1ull);
                if (pos < em.cap) {
                    const int a = em.order[ia], b = em.order[cj[(lane + k) & 31]];
                    em.out_i[pos] = a < b ? a : b;
                    em.out_j[pos] = a < b ? b : a;
                }
            }
            const float rinv = rsqrt_approx(r2);
            const float r2inv = rinv * rinv;
            float g = 0.f;  // dE/dr / r : F_i = g d, F_j = -g d  (d = x_j - x_i)
            if (DO_LJ) {
                if (ONECUT || r2d <= P.rc2_lj) {
                    const float4 lj = slj[k];
                    float a = li.x * lj.x, s = li.y + lj.y;             // 4 eps_ij, sigma_ij
                    if (MASKED) {
                        if ((m14 >> k) & 1u) { a = li.z * lj.z; s = li.w + lj.w; }
                    }
                    const float s2 = s * s * r2inv;
                    const float s6 = s2 * s2 * s2;
                    const float t = a * s6, w = t * s6;                 // 4 eps s^6, 4 eps s^12
                    float e = w - t;
                    float gl = fmaf(-12.f, w, 6.f * t) * r2inv;
                    if (SWITCH) {
                        if (r2 > P.ron2) {
                            const float da = P.rc2_lj - r2;
                            const float S = da * da * fmaf(P.sw_s1, r2, P.sw_s0);
                            const float dS = P.sw_12inv * da * (P.ron2 - r2);   // (dS/dr)/r
                            gl = fmaf(gl, S, e * dS);
                            e *= S;
                        }
                    }
                    if (ENERGY) e_lj += e;
                    g = gl;
                }
            }
            if (DO_COUL) {
                if (ONECUT || r2d <= P.rc2_c) {
                    // erfc(x) = P(t) exp(-x^2):  E = qq P ex / r,  dE/dr / r = -qq ex (P / r + 2 alpha / sqrt(pi)) / r^2
                    const float u = xi.w * xj.w * ex2_approx(-P.alpha2_log2e * r2);
                    const float v = erfcx_poly_r(r2 * rinv, P.alpha04) * rinv;
                    if (ENERGY) e_c = fmaf(u, v, e_c);
                    g = fmaf(-u, (v + P.two_alpha_over_sqrtpi) * r2inv, g);
                }
            }
            fix = fmaf(g, dx, fix); fiy = fmaf(g, dy, fiy); fiz = fmaf(g, dz, fiz);
            fjx = fmaf(-g, dx, fjx); fjy = fmaf(-g, dy, fjy); fjz = fmaf(-g, dz, fjz);
        }
        // the accumulators follow the j-slot: lane l next serves slot (l + k + 1) & 31,
        // whose running sum sits in lane l + 1
        fjx = __shfl_sync(0xffffffffu, fjx, (lane + 1) & 31);
        fjy = __shfl_sync(0xffffffffu, fjy, (lane + 1) & 31);
        fjz = __shfl_sync(0xffffffffu, fjz, (lane + 1) & 31);
    }
}

// SHIFT: every atom of the unit has a unique periodic image within L/2 of the i-block centre (true when the block's
// bounding box is small against the box: R + h <= L/2 on every axis; the list builder flags the rare "wide" blocks —
// a block that straddles the end of a cell row or a domain boundary — in unit.w and those units take the
// canonical path below).  Positions are reduced to that frame once per atom (i: once per unit, j: once per chunk
// when it is staged) and the 32 x 32 inner loop works on plain differences; slots whose r^2 lands within a few
// ulp(L) of a cutoff are re-decided on the canonical expression (chunk_loop), so the pair set is the canonical
// one bit for bit.  Without SHIFT the loop applies the canonical minimum image to every pair (small boxes).
// ONECUT: LJ and Coulomb share one cutoff.  ENERGY: off for the steps of a graph run whose energies nobody reads.
template <bool DO_LJ, bool DO_COUL, bool SWITCH, bool SHIFT, bool ONECUT, bool ENERGY, bool EMIT>
__device__ __forceinline__ void pair_unit(const PairParams &P, const NlistView &nl, const int4 unit, const float4 *__restrict__ xs,
                                          const float4 *__restrict__ ljs, const float4 *__restrict__ bbc,
                                          long long *__restrict__ f_acc, float4 *__restrict__ s_x, float4 *__restrict__ s_lj,
                                          const int lane, long long &e_lj_tot, long long &e_c_tot, const EmitOut &em) {
    const int ia = unit.x * TILE + lane;
    float4 xi = xs[ia];
    const float4 li = ljs[ia];
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (SHIFT) {
        const float4 c = bbc[unit.x];
        cx = c.x; cy = c.y; cz = c.z;
        xi.x = min_image(xi.x - cx, P.L[0], P.invL[0]);
        xi.y = min_image(xi.y - cy, P.L[1], P.invL[1]);
        xi.z = min_image(xi.z - cz, P.L[2], P.invL[2]);
    }
    float fix = 0.f, fiy = 0.f, fiz = 0.f;
    float e_lj = 0.f, e_c = 0.f;

    for (int cidx = 0; cidx < unit.z; ++cidx) {
        const int chunk = unit.y + cidx;
        const int j = nl.chunk_j[(size_t)chunk * 32 + lane];
        const int mslot = nl.chunk_mask[chunk];
        float4 xj_own = xs[j];
        if (SHIFT) {
            xj_own.x = min_image(xj_own.x - cx, P.L[0], P.invL[0]);
            xj_own.y = min_image(xj_own.y - cy, P.L[1], P.invL[1]);
            xj_own.z = min_image(xj_own.z - cz, P.L[2], P.invL[2]);
        }
        __syncwarp();
        s_x[lane] = xj_own; s_x[lane + 32] = xj_own;
        if (DO_LJ) { const float4 lo = ljs[j]; s_lj[lane] = lo; s_lj[lane + 32] = lo; }
        __syncwarp();
        float fjx = 0.f, fjy = 0.f, fjz = 0.f;
        if (mslot >= 0) {
            const unsigned excl = nl.mask_excl[(size_t)mslot * 32 + lane];
            const unsigned m14 = DO_LJ ? nl.mask_14[(size_t)mslot * 32 + lane] : 0u;
            chunk_loop<DO_LJ, DO_COUL, SWITCH, SHIFT, ONECUT, ENERGY, true, EMIT>(
                P, s_x, s_lj, lane, xi, li, excl, m14, fix, fiy, fiz, fjx, fjy, fjz, e_lj, e_c, xs,
                nl.chunk_j + (size_t)chunk * 32, ia, em);
        } else {
            chunk_loop<DO_LJ, DO_COUL, SWITCH, SHIFT, ONECUT, ENERGY, false, EMIT>(
                P, s_x, s_lj, lane, xi, li, 0u, 0u, fix, fiy, fiz, fjx, fjy, fjz, e_lj, e_c, xs,
                nl.chunk_j + (size_t)chunk * 32, ia, em);
        }
        // after 32 rotations lane l holds the sum for slot l again
        if (fjx != 0.f || fjy != 0.f || fjz != 0.f) {
            atomic_add_fix(&f_acc[3 * (size_t)j + 0], to_fix(fjx));
            atomic_add_fix(&f_acc[3 * (size_t)j + 1], to_fix(fjy));
            atomic_add_fix(&f_acc[3 * (size_t)j + 2], to_fix(fjz));
        }
    }
    if (fix != 0.f || fiy != 0.f || fiz != 0.f) {
        atomic_add_fix(&f_acc[3 * (size_t)ia + 0], to_fix(fix));
        atomic_add_fix(&f_acc[3 * (size_t)ia + 1], to_fix(fiy));
        atomic_add_fix(&f_acc[3 * (size_t)ia + 2], to_fix(fiz));
    }
    if (ENERGY) {
        // per-unit energies are converted to fixed point one by one: the total is then independent of
        // which warp (or which rank) happened to evaluate which unit
        e_lj_tot += to_fix((double)e_lj);
        e_c_tot += to_fix((double)e_c);
    }
}

// ---------------------------------------------------------------------------
// Filter-then-compute variant of the unit body ("v5").  The rotation scheme above evaluates the whole interaction
// for the warp whenever ANY lane's slot is inside the cutoff; with 20-30 % of the slots inside, 40 % of the lanes do
// useful work in the expensive part.  Here a chunk goes through three passes:
//   1. filter   lane l (i-atom l) computes r^2 against all 32 staged j (broadcast LDS, 10 instructions per slot) and
//               keeps a bit mask of the slots inside the cutoff (exclusions removed);
//   2. compute  every lane walks ITS OWN set bits: full interaction for that pair, i-force and energies accumulate in
//               registers, the scalar g = (dE/dr)/r goes to a shared 32 x 32 table;
//   3. scatter  the mask is transposed (5 shuffle stages); lane s (j-atom s) walks the i-atoms that interact with it,
//               reads g from the table, rebuilds d from the staged positions and accumulates the j-force in registers.
// Trip counts are max over lanes of the bit counts (~16 at 30 % density instead of 32), nothing goes through atomics
// or changes summation order from run to run: forces stay bitwise reproducible.  The cutoff decision of the SHIFT
// variant is re-taken on the canonical expression in pass 2 for slots within the band (table entry 0 if it fails).
constexpr int G_STRIDE = 33;     // skewed rows: bank = (l + s) mod 32

__device__ __forceinline__ unsigned transpose32(unsigned x, const int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const unsigned k = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
        const unsigned other = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? (((other >> j) & k) | (x & ~k)) : ((x & k) | ((other & k) << j));
    }
    return x;
}

template <bool DO_LJ, bool DO_COUL, bool SWITCH, bool SHIFT, bool ONECUT, bool ENERGY, bool EMIT>
__device__ __forceinline__ void pair_unit_v5(const PairParams &P, const NlistView &nl, const int4 unit, const float4 *__restrict__ xs,
                                             const float4 *__restrict__ ljs, const float4 *__restrict__ bbc,
                                             long long *__restrict__ f_acc, float4 *__restrict__ s_x, float4 *__restrict__ s_lj,
                                             float4 *__restrict__ s_xi, float *__restrict__ s_g, const int lane,
                                             long long &e_lj_tot, long long &e_c_tot, const EmitOut &em) {
    const int ia = unit.x * TILE + lane;
    float4 xi = xs[ia];
    const float4 li = ljs[ia];
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (SHIFT) {
        const float4 c = bbc[unit.x];
        cx = c.x; cy = c.y; cz = c.z;
        xi.x = min_image(xi.x - cx, P.L[0], P.invL[0]);
        xi.y = min_image(xi.y - cy, P.L[1], P.invL[1]);
        xi.z = min_image(xi.z - cz, P.L[2], P.invL[2]);
    }
    __syncwarp();
    s_xi[lane] = xi;
    float fix = 0.f, fiy = 0.f, fiz = 0.f;
    float e_lj = 0.f, e_c = 0.f;

    for (int cidx = 0; cidx < unit.z; ++cidx) {
        const int chunk = unit.y + cidx;
        const int j = nl.chunk_j[(size_t)chunk * 32 + lane];
        const int mslot = nl.chunk_mask[chunk];
        float4 xj_own = xs[j];
        if (SHIFT) {
            xj_own.x = min_image(xj_own.x - cx, P.L[0], P.invL[0]);
            xj_own.y = min_image(xj_own.y - cy, P.L[1], P.invL[1]);
            xj_own.z = min_image(xj_own.z - cz, P.L[2], P.invL[2]);
        }
        __syncwarp();
        s_x[lane] = xj_own;
        if (DO_LJ) s_lj[lane] = ljs[j];
        unsigned excl = 0u, m14 = 0u;
        if (mslot >= 0) {     // stored rotated for the ring kernel (bit k <-> slot (lane + k) & 31): un-rotate
            const unsigned er = nl.mask_excl[(size_t)mslot * 32 + lane];
            excl = __funnelshift_l(er, er, lane);
            if (DO_LJ) { const unsigned mr = nl.mask_14[(size_t)mslot * 32 + lane]; m14 = __funnelshift_l(mr, mr, lane); }
        }
        __syncwarp();
        // ---- pass 1: filter
        unsigned m = 0u;
        const float lim = SHIFT ? P.max_hi : P.rc2_max;
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) {
            const float4 xj = s_x[sl];
            float dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
            if (!SHIFT) {
                dx = min_image(dx, P.L[0], P.invL[0]);
                dy = min_image(dy, P.L[1], P.invL[1]);
                dz = min_image(dz, P.L[2], P.invL[2]);
            }
            if (dist2(dx, dy, dz) <= lim) m |= 1u << sl;
        }
        m &= ~excl;
        // ---- pass 2: the interaction for the surviving pairs of this lane
        unsigned mm = m;
        while (__any_sync(0xffffffffu, mm != 0u)) {
            if (mm) {
                const int sl = __ffs(mm) - 1;
                mm &= mm - 1u;
                const float4 xj = s_x[sl];
                float dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
                if (!SHIFT) {
                    dx = min_image(dx, P.L[0], P.invL[0]);
                    dy = min_image(dy, P.L[1], P.invL[1]);
                    dz = min_image(dz, P.L[2], P.invL[2]);
                }
                const float r2 = dist2(dx, dy, dz);
                float r2d = r2;
                bool in = true;
                if (SHIFT) {
                    bool near = r2 > P.max_lo;
                    if (!ONECUT) near = near || (r2 > P.lj_lo && r2 <= P.lj_hi) || (r2 > P.c_lo && r2 <= P.c_hi);
                    if (near) {
                        r2d = canonical_r2(P.L[0], P.L[1], P.L[2], P.invL[0], P.invL[1], P.invL[2], xs, ia, nl.chunk_j[(size_t)chunk * 32 + sl]);
                        in = r2d <= P.rc2_max;
                    }
                }
                float g = 0.f;
                if (in) {
                    if (EMIT) {
                        const unsigned long long pos = atomicAdd(em.count, 1ull);
                        if (pos < em.cap) {
                            const int a = em.order[ia], b = em.order[nl.chunk_j[(size_t)chunk * 32 + sl]];
                            em.out_i[pos] = a < b ? a : b;
                            em.out_j[pos] = a < b ? b : a;
                        }
                    }
                    const float rinv = rsqrt_approx(r2);
                    const float r2inv = rinv * rinv;
                    if (DO_LJ) {
                        if (ONECUT || r2d <= P.rc2_lj) {
                            const float4 lj = s_lj[sl];
                            float a = li.x * lj.x, sg = li.y + lj.y;             // 4 eps_ij, sigma_ij
                            if ((m14 >> sl) & 1u) { a = li.z * lj.z; sg = li.w + lj.w; }
                            const float s2 = sg * sg * r2inv;
                            const float s6 = s2 * s2 * s2;
                            const float t = a * s6, w = t * s6;
                            float e = w - t;
                            float gl = fmaf(-12.f, w, 6.f * t) * r2inv;
                            if (SWITCH) {
                                if (r2 > P.ron2) {
                                    const float da = P.rc2_lj - r2;
                                    const float S = da * da * fmaf(2.f, r2, P.sw_c0) * P.inv_ab3;
                                    const float dS = P.sw_12inv * da * (P.ron2 - r2);
                                    gl = fmaf(gl, S, e * dS);
                                    e *= S;
                                }
                            }
                            if (ENERGY) e_lj += e;
                            g = gl;
                        }
                    }
                    if (DO_COUL) {
                        if (ONECUT || r2d <= P.rc2_c) {
                            const float u = xi.w * xj.w * ex2_approx(-P.alpha2_log2e * r2);
                            const float v = erfcx_poly(P.alpha * (r2 * rinv)) * rinv;
                            if (ENERGY) e_c = fmaf(u, v, e_c);
                            g = fmaf(-u, (v + P.two_alpha_over_sqrtpi) * r2inv, g);
                        }
                    }
                    fix = fmaf(g, dx, fix); fiy = fmaf(g, dy, fiy); fiz = fmaf(g, dz, fiz);
                }
                s_g[lane * G_STRIDE + sl] = g;
            }
        }
        __syncwarp();
        // ---- pass 3: the same pairs seen from the j side
        unsigned mt = transpose32(m, lane);
        float fjx = 0.f, fjy = 0.f, fjz = 0.f;
        while (__any_sync(0xffffffffu, mt != 0u)) {
            if (mt) {
                const int il = __ffs(mt) - 1;
                mt &= mt - 1u;
                const float g = s_g[il * G_STRIDE + lane];
                const float4 xo = s_xi[il];
                float dx = xj_own.x - xo.x, dy = xj_own.y - xo.y, dz = xj_own.z - xo.z;
                if (!SHIFT) {
                    dx = min_image(dx, P.L[0], P.invL[0]);
                    dy = min_image(dy, P.L[1], P.invL[1]);
                    dz = min_image(dz, P.L[2], P.invL[2]);
                }
                fjx = fmaf(-g, dx, fjx); fjy = fmaf(-g, dy, fjy); fjz = fmaf(-g, dz, fjz);
            }
        }
        if (fjx != 0.f || fjy != 0.f || fjz != 0.f) {
            atomic_add_fix(&f_acc[3 * (size_t)j + 0], to_fix(fjx));
            atomic_add_fix(&f_acc[3 * (size_t)j + 1], to_fix(fjy));
            atomic_add_fix(&f_acc[3 * (size_t)j + 2], to_fix(fjz));
        }
    }
    if (fix != 0.f || fiy != 0.f || fiz != 0.f) {
        atomic_add_fix(&f_acc[3 * (size_t)ia + 0], to_fix(fix));
        atomic_add_fix(&f_acc[3 * (size_t)ia + 1], to_fix(fiy));
        atomic_add_fix(&f_acc[3 * (size_t)ia + 2], to_fix(fiz));
    }
    if (ENERGY) {
        e_lj_tot += to_fix((double)e_lj);
        e_c_tot += to_fix((double)e_c);
    }
}

template <bool DO_LJ, bool DO_COUL, bool SWITCH, bool SHIFT, bool ONECUT, bool ENERGY, bool EMIT = false>
__global__ void __launch_bounds__(PAIR_WARPS * 32)
k_pair5(PairParams P, NlistView nl, const float4 *__restrict__ xs, const float4 *__restrict__ ljs,
        const float4 *__restrict__ bbc, long long *__restrict__ f_acc, long long *__restrict__ e_acc,
        int *__restrict__ cursor, EmitOut em) {
    __shared__ float4 s_x[PAIR_WARPS][32];
    __shared__ float4 s_lj[PAIR_WARPS][32];
    __shared__ float4 s_xi[PAIR_WARPS][32];
    __shared__ float s_g[PAIR_WARPS][32 * G_STRIDE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_units = *nl.n_units;
    long long e_lj_tot = 0, e_c_tot = 0;
    for (int taken = 0; P.max_units == 0 || taken < P.max_units; ++taken) {
        int u = 0;
        if (lane == 0) u = atomicAdd(cursor, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const int4 unit = nl.units[u];
        if (SHIFT && unit.w)
            pair_unit_v5<DO_LJ, DO_COUL, SWITCH, false, ONECUT, ENERGY, EMIT>(P, nl, unit, xs, ljs, bbc, f_acc, s_x[wid], s_lj[wid], s_xi[wid],
                                                                             s_g[wid], lane, e_lj_tot, e_c_tot, em);
        else
            pair_unit_v5<DO_LJ, DO_COUL, SWITCH, SHIFT, ONECUT, ENERGY, EMIT>(P, nl, unit, xs, ljs, bbc, f_acc, s_x[wid], s_lj[wid], s_xi[wid],
                                                                             s_g[wid], lane, e_lj_tot, e_c_tot, em);
    }
    if (ENERGY && DO_LJ) {
        long long v = warp_sum_ll(e_lj_tot);
        if (lane == 0 && v != 0) atomic_add_fix(&e_acc[MDK_E_LJ], v);
    }
    if (ENERGY && DO_COUL) {
        long long v = warp_sum_ll(e_c_tot);
        if (lane == 0 && v != 0) atomic_add_fix(&e_acc[MDK_E_COUL_DIRECT], v);
    }
}

template <bool DO_LJ, bool DO_COUL, bool SWITCH, bool SHIFT, bool ONECUT, bool ENERGY, bool EMIT = false>
__global__ void __launch_bounds__(PAIR_WARPS * 32)
k_pair(PairParams P, NlistView nl, const float4 *__restrict__ xs, const float4 *__restrict__ ljs,
       const float4 *__restrict__ bbc, long long *__restrict__ f_acc, long long *__restrict__ e_acc,
       int *__restrict__ cursor, EmitOut em) {
    __shared__ float4 s_x[PAIR_WARPS][64];
    __shared__ float4 s_lj[PAIR_WARPS][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_units = *nl.n_units;
    long long e_lj_tot = 0, e_c_tot = 0;

    for (int taken = 0; P.max_units == 0 || taken < P.max_units; ++taken) {
        int u = 0;
        if (lane == 0) u = atomicAdd(cursor, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const int4 unit = nl.units[u];
        if (SHIFT && unit.w)      // a wide i-block: canonical minimum image per pair (warp-uniform branch)
            pair_unit<DO_LJ, DO_COUL, SWITCH, false, ONECUT, ENERGY, EMIT>(P, nl, unit, xs, ljs, bbc, f_acc, s_x[wid], s_lj[wid], lane,
                                                                          e_lj_tot, e_c_tot, em);
        else
            pair_unit<DO_LJ, DO_COUL, SWITCH, SHIFT, ONECUT, ENERGY, EMIT>(P, nl, unit, xs, ljs, bbc, f_acc, s_x[wid], s_lj[wid], lane,
                                                                          e_lj_tot, e_c_tot, em);
    }
    if (ENERGY && DO_LJ) {
        long long v = warp_sum_ll(e_lj_tot);
        if (lane == 0 && v != 0) atomic_add_fix(&e_acc[MDK_E_LJ], v);
    }
    if (ENERGY && DO_COUL) {
        long long v = warp_sum_ll(e_c_tot);
        if (lane == 0 && v != 0) atomic_add_fix(&e_acc[MDK_E_COUL_DIRECT], v);
    }
}

static PairParams make_pair_params(mdk_ctx *c, bool do_lj, bool do_coul) {
    PairParams P{};
    for (int a = 0; a < 3; ++a) { P.L[a] = c->box.L[a]; P.invL[a] = c->box.invL[a]; }
    P.rc2_lj = do_lj ? c->rc_lj * c->rc_lj : -1.f;
    float ron = c->r_switch;
    P.ron2 = ron * ron;
    if (do_lj && ron < c->rc_lj) {
        double a2 = (double)c->rc_lj * c->rc_lj, b2 = (double)ron * ron;
        P.inv_ab3 = (float)(1.0 / ((a2 - b2) * (a2 - b2) * (a2 - b2)));
    }
    P.rc2_c = do_coul ? c->rc_coul * c->rc_coul : -1.f;
    P.rc2_max = fmaxf(P.rc2_lj, P.rc2_c);
    P.alpha = (float)c->alpha;
    P.two_alpha_over_sqrtpi = (float)(2.0 * c->alpha / sqrt(M_PI));
    P.alpha2_log2e = (float)(c->alpha * c->alpha * 1.4426950408889634);
    P.sw_c0 = P.rc2_lj - 3.f * P.ron2;
    P.sw_12inv = 12.f * P.inv_ab3;
    P.sw_s0 = P.sw_c0 * P.inv_ab3; P.sw_s1 = 2.f * P.inv_ab3;
    P.alpha04 = 0.4f * P.alpha;
    P.n = c->n;
    // decision band of the SHIFT kernels.  Hoisted coordinates carry <= 1.5 ulp(L) of rounding per atom and axis,
    // the canonical difference 0.5 ulp(L): |d' - d| <= 2 L 2^-23 per axis; delta takes 4x that.
    // |r'^2 - r^2| <= 2 sqrt(3) r delta (+ the roundings of the sum of squares).
    const float Lmax = fmaxf(c->box.L[0], fmaxf(c->box.L[1], c->box.L[2]));
    const float delta = 8.f * Lmax * 1.1920929e-07f;
    auto band = [&](float rc2, float &lo, float &hi) {
        if (rc2 <= 0.f) { lo = hi = -1.f; return; }
        const float b = 1.5f * 3.4641016f * sqrtf(rc2) * delta + 4e-7f * rc2;
        lo = rc2 - b; hi = rc2 + b;
    };
    band(P.rc2_lj, P.lj_lo, P.lj_hi);
    band(P.rc2_c, P.c_lo, P.c_hi);
    band(P.rc2_max, P.max_lo, P.max_hi);
    return P;
}

// ---------------------------------------------------------------------------
// DOUBLE precision (mdpy/environment.py:23-42 switches the reference's arithmetic type): the same tile list, the same
// warp-per-unit rotation scheme and masks, float64 positions (x_cur, wrapped here), float64 parameters, libm erfc / exp.
// Not tuned — it exists so that env.set_precision('DOUBLE') means float64 arithmetic, as in the reference.
struct PairParams64 {
    double L[3];
    double rc2_lj, ron2, inv_ab3, rc2_c, alpha, sqrt_ke;
    int n;
    bool sw;
};

__global__ void __launch_bounds__(128)
k_pair_f64(PairParams64 P, NlistView nl, const int *__restrict__ order, const double *__restrict__ x_cur,
           const double *__restrict__ q64, const double *__restrict__ lj64, long long *__restrict__ f_acc,
           long long *__restrict__ e_acc, int *__restrict__ cursor) {
    __shared__ double s_x[4][32][4];
    __shared__ double s_l[4][32][4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_units = *nl.n_units;
    auto load = [&](int slot, double x[4], double l[4]) {
        x[0] = x[1] = x[2] = x[3] = 0.0; l[0] = l[1] = l[2] = l[3] = 0.0;
        if (slot < P.n) {
            const int a = order[slot];
#pragma unroll
            for (int d = 0; d < 3; ++d) { const double v = x_cur[3 * (size_t)a + d]; x[d] = v - P.L[d] * rint(v / P.L[d]); }
            x[3] = q64 ? q64[a] * P.sqrt_ke : 0.0;
            if (lj64) { l[0] = 2.0 * sqrt(lj64[4 * (size_t)a]); l[1] = 0.5 * lj64[4 * (size_t)a + 1];
                        l[2] = 2.0 * sqrt(lj64[4 * (size_t)a + 2]); l[3] = 0.5 * lj64[4 * (size_t)a + 3]; }
        }
    };
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(cursor, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const int4 unit = nl.units[u];
        const int ia = unit.x * TILE + lane;
        double xi[4], li[4];
        load(ia, xi, li);
        double fi[3] = {0, 0, 0}, e_lj = 0, e_c = 0;
        for (int cidx = 0; cidx < unit.z; ++cidx) {
            const int chunk = unit.y + cidx;
            const int j = nl.chunk_j[(size_t)chunk * 32 + lane];
            const int mslot = nl.chunk_mask[chunk];
            const unsigned excl = mslot >= 0 ? nl.mask_excl[(size_t)mslot * 32 + lane] : 0u;
            const unsigned m14 = mslot >= 0 ? nl.mask_14[(size_t)mslot * 32 + lane] : 0u;
            double xj[4], lj[4];
            load(j, xj, lj);
            __syncwarp();
#pragma unroll
            for (int d = 0; d < 4; ++d) { s_x[wid][lane][d] = xj[d]; s_l[wid][lane][d] = lj[d]; }
            __syncwarp();
            double fj[3] = {0, 0, 0};
            for (int k = 0; k < 32; ++k) {
                const int s = (lane + k) & 31;
                double d[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) { d[a] = s_x[wid][s][a] - xi[a]; d[a] -= P.L[a] * rint(d[a] / P.L[a]); }
                const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                double g = 0.0;
                if (!((excl >> k) & 1u)) {
                    if (P.rc2_lj > 0 && r2 <= P.rc2_lj) {
                        const bool is14 = (m14 >> k) & 1u;
                        const double a4 = is14 ? li[2] * s_l[wid][s][2] : li[0] * s_l[wid][s][0];
                        const double sg = is14 ? li[3] + s_l[wid][s][3] : li[1] + s_l[wid][s][1];
                        const double s2 = sg * sg / r2, s6 = s2 * s2 * s2;
                        const double t = a4 * s6, w = t * s6;
                        double e = w - t, gl = (-12.0 * w + 6.0 * t) / r2;
                        if (P.sw && r2 > P.ron2) {
                            const double da = P.rc2_lj - r2;
                            const double S = da * da * (P.rc2_lj + 2.0 * r2 - 3.0 * P.ron2) * P.inv_ab3;
                            const double dS = 12.0 * P.inv_ab3 * da * (P.ron2 - r2);
                            gl = gl * S + e * dS;
                            e *= S;
                        }
                        e_lj += e; g += gl;
                    }
                    if (P.rc2_c > 0 && r2 <= P.rc2_c) {
                        const double r = sqrt(r2), ar = P.alpha * r, qq = xi[3] * s_x[wid][s][3];
                        const double ec = erfc(ar);
                        e_c += qq * ec / r;
                        g -= qq * (ec / r + 1.1283791670955126 * P.alpha * exp(-ar * ar)) / r2;
                    }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) { fi[a] += g * d[a]; fj[a] -= g * d[a]; }
#pragma unroll
                for (int a = 0; a < 3; ++a) fj[a] = __shfl_sync(0xffffffffu, fj[a], (lane + 1) & 31);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (fj[a] != 0.0) atomic_add_fix(&f_acc[3 * (size_t)j + a], to_fix(fj[a]));
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (fi[a] != 0.0) atomic_add_fix(&f_acc[3 * (size_t)ia + a], to_fix(fi[a]));
        const long long vl = warp_sum_ll(to_fix(e_lj)), vc = warp_sum_ll(to_fix(e_c));
        if (lane == 0 && vl) atomic_add_fix(&e_acc[MDK_E_LJ], vl);
        if (lane == 0 && vc) atomic_add_fix(&e_acc[MDK_E_COUL_DIRECT], vc);
    }
}

int pair_compute_f64(mdk_ctx *c, bool do_lj, bool do_coul) {
    PhaseTimer pt(c, PH_PAIR);
    PairParams64 P{};
    for (int a = 0; a < 3; ++a) P.L[a] = c->box.Ld[a];
    P.rc2_lj = do_lj ? (double)c->rc_lj * c->rc_lj : -1.0;
    P.ron2 = (double)c->r_switch * c->r_switch;
    P.sw = do_lj && c->r_switch < c->rc_lj;
    if (P.sw) { const double a2 = P.rc2_lj, b2 = P.ron2; P.inv_ab3 = 1.0 / ((a2 - b2) * (a2 - b2) * (a2 - b2)); }
    P.rc2_c = do_coul ? (double)c->rc_coul * c->rc_coul : -1.0;
    P.alpha = c->alpha; P.sqrt_ke = sqrt(c->k_e); P.n = c->n;
    MDK_CUDA(c, cudaMemsetAsync(c->counters.p + 3, 0, sizeof(int), c->stream));
    k_pair_f64<<<c->sm_count * 8, 128, 0, c->stream>>>(P, nlist_view(c), c->order.p, c->x_cur.p, do_coul ? c->q64.p : nullptr,
                                                      do_lj ? c->lj64.p : nullptr, c->f_acc.p, reinterpret_cast<long long *>(c->e_acc.p),
                                                      c->counters.p + 3);
    ++c->n_launches; ++c->n_pair_launches;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

int pair_compute(mdk_ctx *c, bool do_lj, bool do_coul) {
    if (!do_lj && !do_coul) return MDK_OK;
    if (do_lj && !c->have_lj) return fail(c, MDK_ERR_NOT_BOUND, "LJ term requested before mdk_set_lj");
    if (do_coul && !c->have_coul) return fail(c, MDK_ERR_NOT_BOUND, "Coulomb term requested before mdk_set_coulomb");
    if (c->dprec && (!do_lj || c->have_lj64) && (!do_coul || c->have_q64)) return pair_compute_f64(c, do_lj, do_coul);
    PhaseTimer pt(c, PH_PAIR);
    PairParams P = make_pair_params(c, do_lj, do_coul);
    NlistView nl = nlist_view(c);
    MDK_CUDA(c, cudaMemsetAsync(c->counters.p + 3, 0, sizeof(int), c->stream));
    const bool sw = do_lj && c->r_switch < c->rc_lj;
    const bool shift = c->shift_ok && !c->force_canonical;
    const bool onecut = !(do_lj && do_coul) || c->rc_lj == c->rc_coul;
    int grid = c->sm_count * c->pair_blocks_per_sm;
    long long max_blocks = (c->stat_units + PAIR_WARPS - 1) / PAIR_WARPS;
    if (max_blocks < 1) max_blocks = 1;
    if (grid > max_blocks && !c->in_capture) grid = (int)max_blocks;   // a captured launch must fit any later list
    P.max_units = c->pair_units_per_warp;
    if (P.max_units > 0) {
        // short-lived blocks: every warp retires after max_units work units, so SM resources come free all the time and the
        // blocks of the (higher-priority) side streams — PME chain, O(N) terms, NCCL transfers — get in between
        const long long units = c->in_capture ? (long long)c->cap_units : c->stat_units;
        const long long per_block = (long long)PAIR_WARPS * P.max_units;
        grid = (int)((units + per_block - 1) / per_block);
        if (grid < 1) grid = 1;
    }
    dim3 g(grid), b(PAIR_WARPS * 32);
    int *cursor = c->counters.p + 3;
    // inner graph steps do not report energies; the energy-less instantiation is only used where it
    // measured faster (plain-cutoff LJ: -9 % at 23 k atoms; with the CHARMM switch ptxas schedules
    // it 18 % slower than the energy-carrying one, so that combination keeps ENERGY on)
    const bool energy = !c->in_capture || c->graph_energy || c->capture_energy || sw;   // the inner steps of a graph run never report energies
#define LAUNCH(LJ, CO, SW, SH, OC, EN)                                                                          \
    do {                                                                                                          \
        if (c->pair_v5)                                                                                           \
            k_pair5<LJ, CO, SW, SH, OC, EN><<<g, b, 0, c->stream>>>(P, nl, c->xs.p, c->ljs.p, c->bb_center.p, c->f_acc.p, \
                                                                    c->e_acc.p, cursor, EmitOut{});                \
        else                                                                                                      \
            k_pair<LJ, CO, SW, SH, OC, EN><<<g, b, 0, c->stream>>>(P, nl, c->xs.p, c->ljs.p, c->bb_center.p, c->f_acc.p, \
                                                                   c->e_acc.p, cursor, EmitOut{});                 \
    } while (0)
#define PICK_EN(LJ, CO, SW, SH, OC) do { if (energy) LAUNCH(LJ, CO, SW, SH, OC, true); else LAUNCH(LJ, CO, SW, SH, OC, false); } while (0)
#define PICK_OC(LJ, CO, SW, SH) do { if (onecut) PICK_EN(LJ, CO, SW, SH, true); else PICK_EN(LJ, CO, SW, SH, false); } while (0)
#define PICK_SH(LJ, CO, SW) do { if (shift) PICK_OC(LJ, CO, SW, true); else PICK_OC(LJ, CO, SW, false); } while (0)
    if (do_lj && do_coul) { if (sw) PICK_SH(true, true, true); else PICK_SH(true, true, false); }
    else if (do_lj)       { if (sw) PICK_SH(true, false, true); else PICK_SH(true, false, false); }
    else                  PICK_SH(false, true, false);
#undef PICK_SH
#undef PICK_OC
#undef PICK_EN
#undef LAUNCH
    c->n_launches += 1;
    c->n_pair_launches += 1;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

// ---------------------------------------------------------------------------
// Test hook: walk the tile list exactly like k_pair and emit every pair that passes the
// canonical cutoff test and is not masked.
__global__ void k_enumerate(PairParams P, NlistView nl, const float4 *__restrict__ xs,
                            const int *__restrict__ order, int *__restrict__ out_i, int *__restrict__ out_j,
                            unsigned long long cap, unsigned long long *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_units = *nl.n_units;
    for (int u = warp; u < n_units; u += n_warps) {
        const int4 unit = nl.units[u];
        const int ia = unit.x * TILE + lane;
        const float4 xi = xs[ia];
        for (int cidx = 0; cidx < unit.z; ++cidx) {
            const int chunk = unit.y + cidx;
            const int j = nl.chunk_j[(size_t)chunk * 32 + lane];
            const int mslot = nl.chunk_mask[chunk];
            unsigned excl = mslot >= 0 ? nl.mask_excl[(size_t)mslot * 32 + lane] : 0u;
            const float4 xj_own = xs[j];
            for (int k = 0; k < 32; ++k) {
                const int slot = (lane + k) & 31;
                float4 xj;
                xj.x = __shfl_sync(0xffffffffu, xj_own.x, slot);
                xj.y = __shfl_sync(0xffffffffu, xj_own.y, slot);
                xj.z = __shfl_sync(0xffffffffu, xj_own.z, slot);
                const int jj = __shfl_sync(0xffffffffu, j, slot);
                const float dx = min_image(xj.x - xi.x, P.L[0], P.invL[0]);
                const float dy = min_image(xj.y - xi.y, P.L[1], P.invL[1]);
                const float dz = min_image(xj.z - xi.z, P.L[2], P.invL[2]);
                const float r2 = dist2(dx, dy, dz);
                if (r2 <= P.rc2_lj && !((excl >> k) & 1u)) {
                    unsigned long long pos = atomicAdd(count, 1ull);
                    if (pos < cap) {
                        int a = order[ia], b = order[jj];
                        out_i[pos] = a < b ? a : b;
                        out_j[pos] = a < b ? b : a;
                    }
                }
            }
        }
    }
}

// production != 0: the pairs come out of k_pair itself (EMIT instantiation of the variant pair_compute would
// launch: same SHIFT / switch / cutoff decision code), on the tile list and tile-order positions exactly as the
// last force evaluation left them — no refresh, no rebuild — so a list that was rebuilt inside a CUDA graph
// is the one that gets tested.  Forces / energies go to scratch accumulators.
int pair_enumerate(mdk_ctx *c, int32_t *out_i, int32_t *out_j, int64_t cap, int64_t *n_out, int production) {
    if (!c->have_lj) return fail(c, MDK_ERR_NOT_BOUND, "mdk_get_pairs needs mdk_set_lj");
    if (production) {
        if (!c->nlist_valid || !c->xs_current)
            return fail(c, MDK_ERR_NOT_BOUND, "mdk_get_pairs(production): no current tile list (run mdk_compute or a step call first)");
        const bool do_coul = c->have_coul && c->rc_coul > 0.f;
        if (do_coul && c->rc_coul != c->rc_lj)
            return fail(c, MDK_ERR_BAD_ARG, "mdk_get_pairs(production) needs one cutoff for LJ and Coulomb");
    } else {
        MDK_TRY(nlist_refresh_sorted(c));
        MDK_TRY(nlist_ensure(c));
    }
    PairParams P = make_pair_params(c, true, production && c->have_coul && c->rc_coul > 0.f);
    int *d_i = nullptr, *d_j = nullptr;
    unsigned long long *d_cnt = nullptr;
    size_t capn = cap > 0 ? (size_t)cap : 1;
    MDK_CUDA(c, cudaMalloc(&d_i, capn * sizeof(int)));
    MDK_CUDA(c, cudaMalloc(&d_j, capn * sizeof(int)));
    MDK_CUDA(c, cudaMalloc(&d_cnt, sizeof(unsigned long long)));
    MDK_CUDA(c, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c->stream));
    if (!production) {
        k_enumerate<<<c->sm_count * 4, 256, 0, c->stream>>>(P, nlist_view(c), c->xs.p, c->order.p, d_i, d_j,
                                                           (unsigned long long)cap, d_cnt);
    } else {
        long long *scratch = nullptr;   // forces + energies of the emitting launch are thrown away
        const size_t words = (size_t)c->n_pad * 3 + MDK_NUM_ENERGIES;
        cudaError_t e0 = cudaMalloc(&scratch, words * sizeof(long long));
        if (e0 != cudaSuccess) { cudaFree(d_i); cudaFree(d_j); cudaFree(d_cnt); MDK_CUDA(c, e0); }
        cudaMemsetAsync(scratch, 0, words * sizeof(long long), c->stream);
        cudaMemsetAsync(c->counters.p + 3, 0, sizeof(int), c->stream);
        EmitOut em{d_i, d_j, c->order.p, (unsigned long long)cap, d_cnt};
        const bool do_coul = P.rc2_c > 0.f;
        const bool sw = c->r_switch < c->rc_lj, shift = c->shift_ok && !c->force_canonical;
        dim3 g(c->sm_count * c->pair_blocks_per_sm), b(PAIR_WARPS * 32);
        NlistView nl = nlist_view(c);
#define EMIT_LAUNCH(CO, SW, SH)                                                                                    \
        do {                                                                                                       \
            if (c->pair_v5)                                                                                        \
                k_pair5<true, CO, SW, SH, true, true, true><<<g, b, 0, c->stream>>>(P, nl, c->xs.p, c->ljs.p, c->bb_center.p, \
                                                                                scratch, scratch + (size_t)c->n_pad * 3, \
                                                                                c->counters.p + 3, em);                  \
            else                                                                                                   \
                k_pair<true, CO, SW, SH, true, true, true><<<g, b, 0, c->stream>>>(P, nl, c->xs.p, c->ljs.p, c->bb_center.p, \
                                                                               scratch, scratch + (size_t)c->n_pad * 3, \
                                                                               c->counters.p + 3, em);                   \
        } while (0)
#define EMIT_SH(CO, SW) do { if (shift) EMIT_LAUNCH(CO, SW, true); else EMIT_LAUNCH(CO, SW, false); } while (0)
        if (do_coul) { if (sw) EMIT_SH(true, true); else EMIT_SH(true, false); }
        else         { if (sw) EMIT_SH(false, true); else EMIT_SH(false, false); }
#undef EMIT_SH
#undef EMIT_LAUNCH
        cudaStreamSynchronize(c->stream);
        cudaFree(scratch);
    }
    ++c->n_launches;
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    size_t got = h < (unsigned long long)cap ? (size_t)h : (size_t)cap;
    if (e == cudaSuccess && got) {
        e = cudaMemcpy(out_i, d_i, got * sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(out_j, d_j, got * sizeof(int), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_i); cudaFree(d_j); cudaFree(d_cnt);
    MDK_CUDA(c, e);
    *n_out = (int64_t)h;
    return MDK_OK;
}

}  // namespace mdk
