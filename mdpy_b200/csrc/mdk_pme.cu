// mdk_pme.cu — smooth particle-mesh Ewald reciprocal step: B-spline charge spreading,
// cuFFT R2C, influence-function multiply (+ energy), cuFFT C2R, force gather.
//
// [not in the reference tree] — the reference only names PME (mdpy/constraint/__init__.py:22,
// forcefield/charmm_forcefield.py:23-24); the algorithm is Essmann et al. 1995 in mdpy
// units, specified in SURVEY §8c and restated in float64 by oracle/spme.py.
//
//   u = (x/L + 1/2) n,  k0 = floor(u),  w = u - k0
//   Q(k) += q theta_x[jx] theta_y[jy] theta_z[jz],  k_a = (k0_a - P + 1 + j_a) mod n_a,
//   theta[j] = M_P(w + P - 1 - j)
//   E = 1/2 sum_m G(m) |F[Q](m)|^2,  G = exp(-pi^2 m~^2/alpha^2) B(m) / (pi V m~^2)
//   phi = F^-1[G F[Q]] (unnormalised),  F_i = -q_i sum_k grad theta_i(k) phi(k)
// Charges arrive pre-multiplied by sqrt(k_e) (xs.w), so G carries no k_e.
//
// Spreading accumulates in int64 fixed point (scale 2^40) with `red.global.add.u64`, so the
// mesh — and with it the whole force evaluation — is bitwise reproducible run to run; the
// conversion kernel turns the mesh into fp32 for cuFFT and clears it for the next step.
#include <algorithm>

#include "mdk_common.cuh"

namespace mdk {

template <int P>
__device__ __forceinline__ void bspline(float w, float (&th)[P], float (&dth)[P]) {
    th[P - 1] = 0.f;
    th[1] = w;
    th[0] = 1.f - w;
#pragma unroll
    for (int k = 3; k < P; ++k) {
        const float div = 1.f / (k - 1);
        th[k - 1] = div * w * th[k - 2];
#pragma unroll
        for (int l = 1; l < k - 1; ++l)
            th[k - l - 1] = div * ((w + l) * th[k - l - 2] + (k - l - w) * th[k - l - 1]);
        th[0] = div * (1.f - w) * th[0];
    }
    dth[0] = -th[0];
#pragma unroll
    for (int l = 1; l < P; ++l) dth[l] = th[l - 1] - th[l];
    const float div = 1.f / (P - 1);
    th[P - 1] = div * w * th[P - 2];
#pragma unroll
    for (int l = 1; l < P - 1; ++l)
        th[P - l - 1] = div * ((w + l) * th[P - l - 2] + (P - l - w) * th[P - l - 1]);
    th[0] = div * (1.f - w) * th[0];
}

struct PmeParams {
    int first, n;         // own tile slots [first, n)
    int nx, ny, nz, nzc;  // nzc = nz/2 + 1
    float invL[3];
    float scale[3];       // n_a / L_a
};

__device__ __forceinline__ void frac_index(float x, float invL, int n, int &k0, float &w) {
    float u = (x * invL + 0.5f) * (float)n;
    float fl = floorf(u);
    w = u - fl;
    k0 = (int)fl;
}

template <int P>
__global__ void k_spread(PmeParams p, const float4 *__restrict__ xs, long long *__restrict__ grid) {
    int i = p.first + blockIdx.x * blockDim.x + threadIdx.x;   // tile slots [first, n): the atoms this rank owns
    if (i >= p.n) return;
    float4 a = xs[i];
    if (a.w == 0.f) return;
    int k0[3]; float w[3];
    frac_index(a.x, p.invL[0], p.nx, k0[0], w[0]);
    frac_index(a.y, p.invL[1], p.ny, k0[1], w[1]);
    frac_index(a.z, p.invL[2], p.nz, k0[2], w[2]);
    float tx[P], ty[P], tz[P], d[P];
    bspline<P>(w[0], tx, d); bspline<P>(w[1], ty, d); bspline<P>(w[2], tz, d);
    int iz[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { int z = (k0[2] - P + 1 + j) % p.nz; iz[j] = z < 0 ? z + p.nz : z; }
#pragma unroll
    for (int jx = 0; jx < P; ++jx) {
        int x = (k0[0] - P + 1 + jx) % p.nx; if (x < 0) x += p.nx;
        const float qx = a.w * tx[jx];
#pragma unroll
        for (int jy = 0; jy < P; ++jy) {
            int y = (k0[1] - P + 1 + jy) % p.ny; if (y < 0) y += p.ny;
            const float qxy = qx * ty[jy];
            long long *row = grid + ((size_t)x * p.ny + y) * p.nz;
#pragma unroll
            for (int jz = 0; jz < P; ++jz) atomic_add_fix(row + iz[jz], to_fix(qxy * tz[jz]));
        }
    }
}

// Atomics-aware spreading.  Atoms arrive cell sorted, so the 256 consecutive atoms of a thread block sit in a small
// region of the mesh: the block accumulates them in a shared-memory sub-mesh (int32 fixed point, native shared
// atomics) and flushes every touched point with ONE 64-bit global atomic — a mesh point is touched by ~6 atoms at
// water density, most of them in the same block.  Integer sums commute: the mesh stays bitwise reproducible and
// independent of how the atoms are grouped into blocks (each contribution is rounded to 2^-26 on its own).
// A block whose atoms span more mesh than fits (a block that wraps a cell row, a dilute system) falls back to
// direct global atomics.
constexpr int SPREAD_CAP = 15360;            // ints of dynamic shared memory (60 KB)
constexpr float SPREAD_SCALE = 67108864.f;   // 2^26; global mesh = 2^40

template <int P, int SPREAD_T>
__global__ void __launch_bounds__(SPREAD_T)
k_spread_smem(PmeParams p, const float4 *__restrict__ xs, long long *__restrict__ grid) {
    extern __shared__ int s_mesh[];
    __shared__ int s_ref[3], s_lo[3], s_hi[3];
    const int tid = threadIdx.x;
    const int i = p.first + blockIdx.x * SPREAD_T + tid;
    const bool live = i < p.n;
    float4 a = live ? xs[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    int k0[3] = {0, 0, 0}; float w[3] = {0.f, 0.f, 0.f};
    const int mesh[3] = {p.nx, p.ny, p.nz};
    if (live) {
        frac_index(a.x, p.invL[0], p.nx, k0[0], w[0]);
        frac_index(a.y, p.invL[1], p.ny, k0[1], w[1]);
        frac_index(a.z, p.invL[2], p.nz, k0[2], w[2]);
    }
    if (tid == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { s_ref[d] = k0[d]; s_lo[d] = 0; s_hi[d] = 0; }
    }
    __syncthreads();
    const bool use = live && a.w != 0.f;
    int rel[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {          // index relative to the block's first atom, periodic
        int r = k0[d] - s_ref[d];
        const int h = mesh[d] >> 1;
        if (r > h) r -= mesh[d]; else if (r < -h) r += mesh[d];
        rel[d] = r;
        if (use) { atomicMin(&s_lo[d], r); atomicMax(&s_hi[d], r); }
    }
    __syncthreads();
    const int sx = s_hi[0] - s_lo[0] + P, sy = s_hi[1] - s_lo[1] + P, sz = s_hi[2] - s_lo[2] + P;
    const long long pts = (long long)sx * sy * sz;
    const bool fits = pts <= SPREAD_CAP && sx <= p.nx && sy <= p.ny && sz <= p.nz;
    float tx[P], ty[P], tz[P], dd[P];
    if (use) { bspline<P>(w[0], tx, dd); bspline<P>(w[1], ty, dd); bspline<P>(w[2], tz, dd); }
    if (!fits) {                            // direct global atomics (uniform branch)
        if (use) {
#pragma unroll
            for (int jx = 0; jx < P; ++jx) {
                int x = (k0[0] - P + 1 + jx) % p.nx; if (x < 0) x += p.nx;
#pragma unroll
                for (int jy = 0; jy < P; ++jy) {
                    int y = (k0[1] - P + 1 + jy) % p.ny; if (y < 0) y += p.ny;
                    long long *row = grid + ((size_t)x * p.ny + y) * p.nz;
#pragma unroll
                    for (int jz = 0; jz < P; ++jz) {
                        int z = (k0[2] - P + 1 + jz) % p.nz; if (z < 0) z += p.nz;
                        atomic_add_fix(row + z, ((long long)__float2int_rn(a.w * tx[jx] * ty[jy] * tz[jz] * SPREAD_SCALE)) << 14);
                    }
                }
            }
        }
        return;
    }
    for (int k = tid; k < (int)pts; k += SPREAD_T) s_mesh[k] = 0;
    __syncthreads();
    if (use) {
        const int bx = rel[0] - s_lo[0], by = rel[1] - s_lo[1], bz = rel[2] - s_lo[2];   // >= 0: lowest point of the support
#pragma unroll
        for (int jx = 0; jx < P; ++jx) {
            const float qx = a.w * tx[jx];
#pragma unroll
            for (int jy = 0; jy < P; ++jy) {
                const float qxy = qx * ty[jy];
                int *row = s_mesh + ((bx + jx) * sy + (by + jy)) * sz + bz;
#pragma unroll
                for (int jz = 0; jz < P; ++jz) atomicAdd(row + jz, __float2int_rn(qxy * tz[jz] * SPREAD_SCALE));
            }
        }
    }
    __syncthreads();
    // lowest mesh index of the sub-mesh on each axis
    int o[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { int v = (s_ref[d] + s_lo[d] - P + 1) % mesh[d]; o[d] = v < 0 ? v + mesh[d] : v; }
    for (int k = tid; k < (int)pts; k += SPREAD_T) {
        const int v = s_mesh[k];
        if (!v) continue;
        const int iz = k % sz, iy = (k / sz) % sy, ix = k / (sz * sy);
        int x = o[0] + ix; if (x >= p.nx) x -= p.nx;
        int y = o[1] + iy; if (y >= p.ny) y -= p.ny;
        int z = o[2] + iz; if (z >= p.nz) z -= p.nz;
        atomic_add_fix(grid + ((size_t)x * p.ny + y) * p.nz + z, ((long long)v) << 14);
    }
}

__global__ void k_grid_convert(size_t total, long long *__restrict__ fix, float *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long v = fix[i];
    out[i] = (float)((double)v * (1.0 / FIX_SCALE));
    if (v != 0) fix[i] = 0;
}

__global__ void k_convolve(int nx, int ny, int nzc, int nz, float2 *__restrict__ gc,
                           const float *__restrict__ G, long long *__restrict__ e_acc) {
    size_t total = (size_t)nx * ny * nzc;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < total) {
        int kz = (int)(i % nzc);
        float g = G[i];
        float2 c = gc[i];
        float wgt = (kz == 0 || (2 * kz == nz)) ? 1.f : 2.f;
        e = 0.5 * (double)(wgt * g * (c.x * c.x + c.y * c.y));
        gc[i] = make_float2(c.x * g, c.y * g);
    }
    e = warp_sum(e);
    __shared__ double s[8];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s[wid] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s[k];
        if (t != 0.0) atomic_add_fix(&e_acc[MDK_E_PME_RECIP], to_fix(t));
    }
}

template <int P>
__global__ void k_gather(PmeParams p, const float4 *__restrict__ xs, const float *__restrict__ phi,
                         long long *__restrict__ f_acc) {
    int i = p.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float4 a = xs[i];
    if (a.w == 0.f) return;
    int k0[3]; float w[3];
    frac_index(a.x, p.invL[0], p.nx, k0[0], w[0]);
    frac_index(a.y, p.invL[1], p.ny, k0[1], w[1]);
    frac_index(a.z, p.invL[2], p.nz, k0[2], w[2]);
    float tx[P], ty[P], tz[P], dx[P], dy[P], dz[P];
    bspline<P>(w[0], tx, dx); bspline<P>(w[1], ty, dy); bspline<P>(w[2], tz, dz);
    int iz[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { int z = (k0[2] - P + 1 + j) % p.nz; iz[j] = z < 0 ? z + p.nz : z; }
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
    for (int jx = 0; jx < P; ++jx) {
        int x = (k0[0] - P + 1 + jx) % p.nx; if (x < 0) x += p.nx;
#pragma unroll
        for (int jy = 0; jy < P; ++jy) {
            int y = (k0[1] - P + 1 + jy) % p.ny; if (y < 0) y += p.ny;
            const float *row = phi + ((size_t)x * p.ny + y) * p.nz;
            float s0 = 0.f, s1 = 0.f;  // sum_z theta_z phi, sum_z dtheta_z phi
#pragma unroll
            for (int jz = 0; jz < P; ++jz) {
                float v = __ldg(row + iz[jz]);
                s0 = fmaf(tz[jz], v, s0);
                s1 = fmaf(dz[jz], v, s1);
            }
            fx = fmaf(dx[jx] * ty[jy], s0, fx);
            fy = fmaf(tx[jx] * dy[jy], s0, fy);
            fz = fmaf(tx[jx] * ty[jy], s1, fz);
        }
    }
    atomic_add_fix(&f_acc[3 * (size_t)i + 0], to_fix(-a.w * p.scale[0] * fx));
    atomic_add_fix(&f_acc[3 * (size_t)i + 1], to_fix(-a.w * p.scale[1] * fy));
    atomic_add_fix(&f_acc[3 * (size_t)i + 2], to_fix(-a.w * p.scale[2] * fz));
}


// ---------------------------------------------------------------------------
// Small power-of-two meshes (every axis 8..64): the whole mesh chain between spreading and gathering
// in three launches instead of cuFFT's six plus two of ours — at 64^3 those eight kernels are 3-5 us
// each, pure launch latency, and sit on the critical path of the 23k-atom step.
//   k_mesh_fwd_yz   one block per x-plane: int64 mesh -> float (and clear), 2-D FFT over (y, z)
//   k_mesh_x_conv   one block per y: FFT over x, influence function + energy, inverse FFT over x
//   k_mesh_inv_yz   one block per x-plane: inverse 2-D FFT over (y, z), real part -> potential mesh
// Complex-to-complex on the full spectrum (the mesh is 2 MB: redundancy is cheaper than a launch).
// The FFTs run in shared memory without any reordering pass: forward transforms are decimation in
// frequency (natural order in, bit-reversed out), inverse transforms decimation in time (bit-reversed in,
// natural out), so the spectrum simply lives in bit-reversed order between the kernels and only the
// influence-function lookup has to un-reverse its indices.  Two radix-2 stages are fused per pass over
// the tile (four points per thread in registers).  tw[k] = exp(-+ 2 pi i k / 64), k < 32.
constexpr int FFT_NMAX = 64;
constexpr int MESH_T = 1024;

__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// element (b, k) of sequence b lives at s[b * sb + k * sk]; nb = 1 << lognb sequences of length 1 << logn.
// over_batch: consecutive threads walk the batch index (the unit-stride direction of a strided axis).
template <bool INVERSE>
__device__ __forceinline__ void smem_fft(float2 *s, int logn, int lognb, int sb, int sk, const float2 *tw, bool over_batch) {
    const int n = 1 << logn, nb = 1 << lognb;
    if (!INVERSE) {
        int h = n >> 1;   // half size of the next stage
        for (; h >= 2; h >>= 2) {          // fused stages (h, h / 2)
            const int Q = h >> 1, logQ = 31 - __clz(Q);
            for (int t = threadIdx.x; t < (nb << (logn - 2)); t += blockDim.x) {
                int b, q;
                if (over_batch) { b = t & (nb - 1); q = t >> lognb; } else { q = t & ((n >> 2) - 1); b = t >> (logn - 2); }
                const int pos = q & (Q - 1), i0 = ((q >> logQ) << (logQ + 2)) + pos;
                float2 *p = s + b * sb + i0 * sk;
                const int st = Q * sk;
                float2 e0 = p[0], e1 = p[st], e2 = p[2 * st], e3 = p[3 * st];
                const float2 wa = tw[pos * (16 >> logQ)], wb = tw[pos * (16 >> logQ) + 16], wc = tw[pos * (32 >> logQ)];
                const float2 a0 = cadd(e0, e2), a2 = cmul(csub(e0, e2), wa), a1 = cadd(e1, e3), a3 = cmul(csub(e1, e3), wb);
                p[0] = cadd(a0, a1); p[st] = cmul(csub(a0, a1), wc);
                p[2 * st] = cadd(a2, a3); p[3 * st] = cmul(csub(a2, a3), wc);
            }
            __syncthreads();
        }
        if (h == 1) {                      // odd number of stages: last one is a plain butterfly
            for (int t = threadIdx.x; t < (nb << (logn - 1)); t += blockDim.x) {
                int b, j;
                if (over_batch) { b = t & (nb - 1); j = t >> lognb; } else { j = t & ((n >> 1) - 1); b = t >> (logn - 1); }
                float2 *p = s + b * sb + 2 * j * sk;
                const float2 a = p[0], c = p[sk];
                p[0] = cadd(a, c); p[sk] = csub(a, c);
            }
            __syncthreads();
        }
    } else {
        int Q = 1;
        if (logn & 1) {
            for (int t = threadIdx.x; t < (nb << (logn - 1)); t += blockDim.x) {
                int b, j;
                if (over_batch) { b = t & (nb - 1); j = t >> lognb; } else { j = t & ((n >> 1) - 1); b = t >> (logn - 1); }
                float2 *p = s + b * sb + 2 * j * sk;
                const float2 a = p[0], c = p[sk];
                p[0] = cadd(a, c); p[sk] = csub(a, c);
            }
            __syncthreads();
            Q = 2;
        }
        for (; 4 * Q <= n; Q <<= 2) {      // fused stages (Q, 2 Q)
            const int logQ = 31 - __clz(Q);
            for (int t = threadIdx.x; t < (nb << (logn - 2)); t += blockDim.x) {
                int b, q;
                if (over_batch) { b = t & (nb - 1); q = t >> lognb; } else { q = t & ((n >> 2) - 1); b = t >> (logn - 2); }
                const int pos = q & (Q - 1), i0 = ((q >> logQ) << (logQ + 2)) + pos;
                float2 *p = s + b * sb + i0 * sk;
                const int st = Q * sk;
                float2 e0 = p[0], e1 = p[st], e2 = p[2 * st], e3 = p[3 * st];
                const float2 wc = tw[pos * (32 >> logQ)], wa = tw[pos * (16 >> logQ)], wb = tw[pos * (16 >> logQ) + 16];
                const float2 t1 = cmul(e1, wc), t3 = cmul(e3, wc);
                const float2 a0 = cadd(e0, t1), a1 = csub(e0, t1), a2 = cadd(e2, t3), a3 = csub(e2, t3);
                const float2 u2 = cmul(a2, wa), u3 = cmul(a3, wb);
                p[0] = cadd(a0, u2); p[2 * st] = csub(a0, u2);
                p[st] = cadd(a1, u3); p[3 * st] = csub(a1, u3);
            }
            __syncthreads();
        }
    }
}

struct MeshDims { int nx, ny, nz, lx, ly, lz; };

__global__ void __launch_bounds__(MESH_T)
k_mesh_fwd_yz(MeshDims d, long long *__restrict__ fix, float2 *__restrict__ spec, const float2 *__restrict__ tw_g) {
    extern __shared__ float2 sm[];
    float2 *tw = sm, *tile = sm + 32;
    if (threadIdx.x < 32) tw[threadIdx.x] = tw_g[threadIdx.x];
    const int plane = d.ny * d.nz;
    long long *src = fix + (size_t)blockIdx.x * plane;
#pragma unroll 4
    for (int i = threadIdx.x; i < plane; i += MESH_T) {
        const long long v = src[i];
        tile[i] = make_float2((float)((double)v * (1.0 / FIX_SCALE)), 0.f);
        if (v != 0) src[i] = 0;
    }
    __syncthreads();
    smem_fft<false>(tile, d.lz, d.ly, d.nz, 1, tw, false);
    smem_fft<false>(tile, d.ly, d.lz, 1, d.nz, tw, true);
    float2 *dst = spec + (size_t)blockIdx.x * plane;
#pragma unroll 4
    for (int i = threadIdx.x; i < plane; i += MESH_T) dst[i] = tile[i];
}

__global__ void __launch_bounds__(MESH_T)
k_mesh_x_conv(MeshDims d, float2 *__restrict__ spec, const float *__restrict__ G, const float2 *__restrict__ tw_g,
              long long *__restrict__ e_acc) {
    extern __shared__ float2 sm[];
    float2 *twf = sm, *twi = sm + 32, *tile = sm + 64;   // tile[x][z] for this block's y position
    if (threadIdx.x < 64) sm[threadIdx.x] = tw_g[threadIdx.x];
    const int yp = blockIdx.x, nzc = d.nz / 2 + 1, cnt = d.nx * d.nz;
    const int ky = (int)(__brev((unsigned)yp) >> (32 - d.ly));          // the spectrum sits in bit-reversed order
#pragma unroll 4
    for (int i = threadIdx.x; i < cnt; i += MESH_T) {
        const int x = i >> d.lz, z = i & (d.nz - 1);
        tile[i] = spec[(((size_t)x << d.ly) + yp) * d.nz + z];
    }
    __syncthreads();
    smem_fft<false>(tile, d.lx, d.lz, 1, d.nz, twf, true);
    double e = 0.0;
#pragma unroll 4
    for (int i = threadIdx.x; i < cnt; i += MESH_T) {
        const int xp = i >> d.lz, zp = i & (d.nz - 1);
        const int kx = (int)(__brev((unsigned)xp) >> (32 - d.lx)), kz = (int)(__brev((unsigned)zp) >> (32 - d.lz));
        const float g = G[((size_t)kx * d.ny + ky) * nzc + (2 * kz <= d.nz ? kz : d.nz - kz)];
        const float2 c = tile[i];
        e += 0.5 * (double)(g * (c.x * c.x + c.y * c.y));
        tile[i] = make_float2(c.x * g, c.y * g);
    }
    __syncthreads();
    smem_fft<true>(tile, d.lx, d.lz, 1, d.nz, twi, true);
#pragma unroll 4
    for (int i = threadIdx.x; i < cnt; i += MESH_T) {
        const int x = i >> d.lz, z = i & (d.nz - 1);
        spec[(((size_t)x << d.ly) + yp) * d.nz + z] = tile[i];
    }
    // per-thread fixed point, integer sums: the energy does not depend on the reduction order
    long long v = warp_sum_ll(to_fix(e));
    if ((threadIdx.x & 31) == 0 && v != 0) atomic_add_fix(&e_acc[MDK_E_PME_RECIP], v);
}

__global__ void __launch_bounds__(MESH_T)
k_mesh_inv_yz(MeshDims d, const float2 *__restrict__ spec, float *__restrict__ phi, const float2 *__restrict__ tw_g) {
    extern __shared__ float2 sm[];
    float2 *tw = sm, *tile = sm + 32;
    if (threadIdx.x < 32) tw[threadIdx.x] = tw_g[32 + threadIdx.x];
    const int plane = d.ny * d.nz;
    const float2 *src = spec + (size_t)blockIdx.x * plane;
#pragma unroll 4
    for (int i = threadIdx.x; i < plane; i += MESH_T) tile[i] = src[i];
    __syncthreads();
    smem_fft<true>(tile, d.ly, d.lz, 1, d.nz, tw, true);
    smem_fft<true>(tile, d.lz, d.ly, d.nz, 1, tw, false);
    float *dst = phi + (size_t)blockIdx.x * plane;
#pragma unroll 4
    for (int i = threadIdx.x; i < plane; i += MESH_T) dst[i] = tile[i].x;
}

static int ilog2_exact(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return (1 << l) == n ? l : -1;
}
static bool mesh_fast_ok(const mdk_ctx *c) {
    if (c->pme_force_cufft || c->dd) return false;   // the decomposed step adds the other domains' sub-meshes in float: cuFFT chain
    for (int a = 0; a < 3; ++a)
        if (c->pme_n[a] < 8 || c->pme_n[a] > FFT_NMAX || ilog2_exact(c->pme_n[a]) < 0) return false;
    return true;
}

// ---------------------------------------------------------------------------
// host: influence function.  M_P at the integers by the cardinal B-spline recursion.
static void bspline_moduli(int n, int P, std::vector<double> &mod) {
    std::vector<double> M(P + 1, 0.0);  // M_P(k), k = 0..P
    // M_2(u) = 1 - |u - 1| on [0, 2]
    std::vector<double> cur(P + 2, 0.0), nxt(P + 2, 0.0);
    cur[1] = 1.0;  // order 2: M_2(1) = 1
    for (int ord = 3; ord <= P; ++ord) {
        for (int k = 0; k <= ord; ++k) {
            double a = k > 0 || true ? cur[k] : 0.0;
            double b = k >= 1 ? cur[k - 1] : 0.0;
            nxt[k] = ((double)k * a + (double)(ord - k) * b) / (ord - 1);
        }
        cur = nxt;
    }
    for (int k = 0; k <= P; ++k) M[k] = cur[k];
    mod.assign(n, 0.0);
    for (int m = 0; m < n; ++m) {
        double sr = 0.0, si = 0.0;
        for (int k = 0; k <= P - 2; ++k) {
            double arg = 2.0 * M_PI * m * k / n;
            sr += M[k + 1] * cos(arg);
            si += M[k + 1] * sin(arg);
        }
        mod[m] = sr * sr + si * si;
    }
    // standard fix for (near-)zero moduli with odd orders / even n
    for (int m = 0; m < n; ++m)
        if (mod[m] < 1e-7) mod[m] = 0.5 * (mod[(m + n - 1) % n] + mod[(m + 1) % n]);
}

int pme_prepare(mdk_ctx *c) {
    if (!c->have_pme) return fail(c, MDK_ERR_NOT_BOUND, "PME term requested before mdk_set_pme");
    if (!c->have_coul || c->alpha <= 0) return fail(c, MDK_ERR_NOT_BOUND, "PME needs mdk_set_coulomb with alpha > 0");
    if (!c->pme_dirty) return MDK_OK;
    const int nx = c->pme_n[0], ny = c->pme_n[1], nz = c->pme_n[2], nzc = nz / 2 + 1;
    const int P = c->pme_order;
    size_t total = (size_t)nx * ny * nz, totc = (size_t)nx * ny * nzc;
    MDK_CUDA(c, c->grid_fix.reserve(total));
    MDK_CUDA(c, c->grid_r.reserve(total));
    MDK_CUDA(c, c->grid_c.reserve(totc));
    MDK_CUDA(c, c->influence.reserve(totc));
    MDK_CUDA(c, cudaMemsetAsync(c->grid_fix.p, 0, total * sizeof(long long), c->stream));
    std::vector<double> bx, by, bz;
    bspline_moduli(nx, P, bx); bspline_moduli(ny, P, by); bspline_moduli(nz, P, bz);
    std::vector<float> G(totc);
    const double V = c->box.Ld[0] * c->box.Ld[1] * c->box.Ld[2];
    const double pref = 1.0 / (M_PI * V), pa = M_PI * M_PI / (c->alpha * c->alpha);
    for (int ix = 0; ix < nx; ++ix) {
        double mx = (ix <= nx / 2 ? ix : ix - nx) / c->box.Ld[0];
        for (int iy = 0; iy < ny; ++iy) {
            double my = (iy <= ny / 2 ? iy : iy - ny) / c->box.Ld[1];
            for (int iz = 0; iz < nzc; ++iz) {
                double mz = iz / c->box.Ld[2];
                double m2 = mx * mx + my * my + mz * mz;
                double g = 0.0;
                if (m2 > 0) g = pref * exp(-pa * m2) / (m2 * bx[ix] * by[iy] * bz[iz]);
                G[((size_t)ix * ny + iy) * nzc + iz] = (float)g;
            }
        }
    }
    MDK_CUDA(c, cudaMemcpyAsync(c->influence.p, G.data(), totc * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->have_plans) { cufftDestroy(c->plan_r2c); cufftDestroy(c->plan_c2r); c->have_plans = false; }
    c->pme_fast = mesh_fast_ok(c);
    if (c->pme_fast) {
        // full complex spectrum in grid_c, twiddles exp(-+ 2 pi i k / 64)
        MDK_CUDA(c, c->grid_c.reserve(total));
        MDK_CUDA(c, c->fft_tw.reserve(64));
        float2 tw[64];
        for (int k = 0; k < 32; ++k) {
            const double a = 2.0 * M_PI * k / FFT_NMAX;
            tw[k] = make_float2((float)cos(a), (float)-sin(a));
            tw[32 + k] = make_float2((float)cos(a), (float)sin(a));
        }
        MDK_CUDA(c, cudaMemcpyAsync(c->fft_tw.p, tw, sizeof(tw), cudaMemcpyHostToDevice, c->stream));
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
        const int smem = (64 + std::max(ny * nz, nx * nz)) * (int)sizeof(float2);
        MDK_CUDA(c, cudaFuncSetAttribute(k_mesh_fwd_yz, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        MDK_CUDA(c, cudaFuncSetAttribute(k_mesh_x_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        MDK_CUDA(c, cudaFuncSetAttribute(k_mesh_inv_yz, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    } else {
        if (cufftPlan3d(&c->plan_r2c, nx, ny, nz, CUFFT_R2C) != CUFFT_SUCCESS ||
            cufftPlan3d(&c->plan_c2r, nx, ny, nz, CUFFT_C2R) != CUFFT_SUCCESS)
            return fail(c, MDK_ERR_CUDA, "cufftPlan3d(%d,%d,%d) failed", nx, ny, nz);
        c->have_plans = true;
    }
    {
        const int smem = SPREAD_CAP * (int)sizeof(int);
#define SPREAD_ATTR(P) MDK_CUDA(c, cudaFuncSetAttribute(k_spread_smem<P, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
                       MDK_CUDA(c, cudaFuncSetAttribute(k_spread_smem<P, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
        SPREAD_ATTR(4); SPREAD_ATTR(5); SPREAD_ATTR(6); SPREAD_ATTR(8);
#undef SPREAD_ATTR
    }
    c->pme_dirty = false;
    return MDK_OK;
}

static PmeParams make_pme_params(mdk_ctx *c) {
    PmeParams p{};
    p.first = pme_first(c); p.n = pme_end(c);
    p.nx = c->pme_n[0]; p.ny = c->pme_n[1]; p.nz = c->pme_n[2]; p.nzc = p.nz / 2 + 1;
    for (int a = 0; a < 3; ++a) {
        p.invL[a] = c->box.invL[a];
        p.scale[a] = (float)(c->pme_n[a] / c->box.Ld[a]);
    }
    return p;
}

// own atoms -> fixed-point charge mesh
int pme_spread(mdk_ctx *c) {
    PhaseTimer pt(c, PH_SPREAD);
    PmeParams p = make_pme_params(c);
    const int cnt = p.n - p.first;
    if (cnt <= 0) return MDK_OK;
    if (c->spread_smem) {
        // 256 atoms per block merge more contributions per global atomic; small systems take 128 so that the launch
        // still covers the machine (23k atoms: 92 blocks of 256 would leave a third of the SMs idle)
        const bool small = cnt < 256 * c->sm_count * 2;
        const int T = small ? 128 : 256;
        const int B = (cnt + T - 1) / T;
        const size_t smem = SPREAD_CAP * sizeof(int);
#define SPREAD(P) do { if (small) k_spread_smem<P, 128><<<B, 128, smem, c->stream>>>(p, c->xs.p, c->grid_fix.p); \
                       else k_spread_smem<P, 256><<<B, 256, smem, c->stream>>>(p, c->xs.p, c->grid_fix.p); } while (0)
        switch (c->pme_order) {
            case 4: SPREAD(4); break;
            case 5: SPREAD(5); break;
            case 6: SPREAD(6); break;
            case 8: SPREAD(8); break;
            default: return fail(c, MDK_ERR_BAD_ARG, "PME order %d not supported (4, 5, 6, 8)", c->pme_order);
        }
#undef SPREAD
        c->n_launches += 1;
        MDK_CUDA(c, cudaGetLastError());
        return MDK_OK;
    }
    const int B = (cnt + 127) / 128;
    switch (c->pme_order) {
        case 4: k_spread<4><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_fix.p); break;
        case 5: k_spread<5><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_fix.p); break;
        case 6: k_spread<6><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_fix.p); break;
        case 8: k_spread<8><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_fix.p); break;
        default: return fail(c, MDK_ERR_BAD_ARG, "PME order %d not supported (4, 5, 6, 8)", c->pme_order);
    }
    c->n_launches += 1;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

// charge mesh -> potential mesh.  convert: the charge mesh is the fixed-point one (converted to float and cleared
// here); otherwise grid_r already holds it in float (decomposed step: sub-meshes added on the mesh rank).
int pme_mesh(mdk_ctx *c, bool convert) {
    const PmeParams p = make_pme_params(c);
    const size_t total = (size_t)p.nx * p.ny * p.nz, totc = (size_t)p.nx * p.ny * p.nzc;
    if (c->pme_fast) {
        PhaseTimer pt(c, PH_FFT);
        MeshDims d{p.nx, p.ny, p.nz, ilog2_exact(p.nx), ilog2_exact(p.ny), ilog2_exact(p.nz)};
        const size_t smem = (64 + (size_t)std::max(p.ny * p.nz, p.nx * p.nz)) * sizeof(float2);
        k_mesh_fwd_yz<<<p.nx, MESH_T, smem, c->stream>>>(d, c->grid_fix.p, c->grid_c.p, c->fft_tw.p);
        k_mesh_x_conv<<<p.ny, MESH_T, smem, c->stream>>>(d, c->grid_c.p, c->influence.p, c->fft_tw.p,
                                                     reinterpret_cast<long long *>(c->e_acc.p));
        k_mesh_inv_yz<<<p.nx, MESH_T, smem, c->stream>>>(d, c->grid_c.p, c->grid_r.p, c->fft_tw.p);
        c->n_launches += 3;
        MDK_CUDA(c, cudaGetLastError());
        return MDK_OK;
    }
    cufftSetStream(c->plan_r2c, c->stream);
    cufftSetStream(c->plan_c2r, c->stream);
    if (convert) {
        PhaseTimer pt(c, PH_SPREAD);
        k_grid_convert<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(total, c->grid_fix.p, c->grid_r.p);
        c->n_launches += 1;
    }
    PhaseTimer pt(c, PH_FFT);
    if (cufftExecR2C(c->plan_r2c, c->grid_r.p, reinterpret_cast<cufftComplex *>(c->grid_c.p)) != CUFFT_SUCCESS)
        return fail(c, MDK_ERR_CUDA, "cufftExecR2C failed");
    k_convolve<<<(unsigned)((totc + 255) / 256), 256, 0, c->stream>>>(
        p.nx, p.ny, p.nzc, p.nz, c->grid_c.p, c->influence.p, reinterpret_cast<long long *>(c->e_acc.p));
    if (cufftExecC2R(c->plan_c2r, reinterpret_cast<cufftComplex *>(c->grid_c.p), c->grid_r.p) != CUFFT_SUCCESS)
        return fail(c, MDK_ERR_CUDA, "cufftExecC2R failed");
    c->n_launches += 1;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

// potential mesh -> forces on own atoms
int pme_gather(mdk_ctx *c) {
    PhaseTimer pt(c, PH_GATHER);
    PmeParams p = make_pme_params(c);
    const int cnt = p.n - p.first;
    if (cnt <= 0) return MDK_OK;
    const int B = (cnt + 127) / 128;
    switch (c->pme_order) {
        case 4: k_gather<4><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_r.p, c->f_acc.p); break;
        case 5: k_gather<5><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_r.p, c->f_acc.p); break;
        case 6: k_gather<6><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_r.p, c->f_acc.p); break;
        case 8: k_gather<8><<<B, 128, 0, c->stream>>>(p, c->xs.p, c->grid_r.p, c->f_acc.p); break;
        default: return fail(c, MDK_ERR_BAD_ARG, "PME order %d not supported (4, 5, 6, 8)", c->pme_order);
    }
    c->n_launches += 1;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

int pme_compute(mdk_ctx *c) {
    MDK_TRY(pme_prepare(c));
    MDK_TRY(pme_spread(c));
    MDK_TRY(pme_mesh(c, true));
    return pme_gather(c);
}

}  // namespace mdk
