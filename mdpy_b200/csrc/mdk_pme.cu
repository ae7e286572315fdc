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
    int n;
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
    int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    if (cufftPlan3d(&c->plan_r2c, nx, ny, nz, CUFFT_R2C) != CUFFT_SUCCESS ||
        cufftPlan3d(&c->plan_c2r, nx, ny, nz, CUFFT_C2R) != CUFFT_SUCCESS)
        return fail(c, MDK_ERR_CUDA, "cufftPlan3d(%d,%d,%d) failed", nx, ny, nz);
    c->have_plans = true;
    c->pme_dirty = false;
    return MDK_OK;
}

template <int P>
static int pme_run(mdk_ctx *c) {
    PmeParams p{};
    p.n = c->n;
    p.nx = c->pme_n[0]; p.ny = c->pme_n[1]; p.nz = c->pme_n[2]; p.nzc = p.nz / 2 + 1;
    for (int a = 0; a < 3; ++a) {
        p.invL[a] = c->box.invL[a];
        p.scale[a] = (float)(c->pme_n[a] / c->box.Ld[a]);
    }
    size_t total = (size_t)p.nx * p.ny * p.nz, totc = (size_t)p.nx * p.ny * p.nzc;
    cufftSetStream(c->plan_r2c, c->stream);
    cufftSetStream(c->plan_c2r, c->stream);
    {
        PhaseTimer pt(c, PH_SPREAD);
        k_spread<P><<<(c->n + 127) / 128, 128, 0, c->stream>>>(p, c->xs.p, c->grid_fix.p);
        k_grid_convert<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(total, c->grid_fix.p, c->grid_r.p);
        c->n_launches += 2;
    }
    {
        PhaseTimer pt(c, PH_FFT);
        if (cufftExecR2C(c->plan_r2c, c->grid_r.p, reinterpret_cast<cufftComplex *>(c->grid_c.p)) != CUFFT_SUCCESS)
            return fail(c, MDK_ERR_CUDA, "cufftExecR2C failed");
        k_convolve<<<(unsigned)((totc + 255) / 256), 256, 0, c->stream>>>(
            p.nx, p.ny, p.nzc, p.nz, c->grid_c.p, c->influence.p, reinterpret_cast<long long *>(c->e_acc.p));
        if (cufftExecC2R(c->plan_c2r, reinterpret_cast<cufftComplex *>(c->grid_c.p), c->grid_r.p) != CUFFT_SUCCESS)
            return fail(c, MDK_ERR_CUDA, "cufftExecC2R failed");
        c->n_launches += 1;
    }
    {
        PhaseTimer pt(c, PH_GATHER);
        k_gather<P><<<(c->n + 127) / 128, 128, 0, c->stream>>>(p, c->xs.p, c->grid_r.p, c->f_acc.p);
        c->n_launches += 1;
    }
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

int pme_compute(mdk_ctx *c) {
    MDK_TRY(pme_prepare(c));
    switch (c->pme_order) {
        case 4: return pme_run<4>(c);
        case 5: return pme_run<5>(c);
        case 6: return pme_run<6>(c);
        case 8: return pme_run<8>(c);
    }
    return fail(c, MDK_ERR_BAD_ARG, "PME order %d not supported (4, 5, 6, 8)", c->pme_order);
}

}  // namespace mdk
