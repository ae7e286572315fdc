// mdk_api.cu — the extern "C" boundary of libmdpyb200.so (see include/mdpy_b200.h).
#include <stdarg.h>

#include <cmath>

#include "mdk_common.cuh"

namespace mdk {

static thread_local std::string g_create_err;

int fail(mdk_ctx *c, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

// A changed box / cutoff / parameter table / exclusion set: the tile list, the captured graphs and the
// integrators' cached forces (f_prev of the Langevin step, x_prev of the Verlet step) all belong to the old system.
static void invalidate(mdk_ctx *c) {
    c->nlist_valid = false; c->xs_current = false;
    c->verlet_cached = false; c->langevin_cached = false;
    ++c->graph_epoch;
}

__global__ void k_f32_to_f64(size_t n, const float *__restrict__ in, double *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

// forces back to matrix_id order
template <typename T>
__global__ void k_unpermute_forces(int n, const int *__restrict__ order, const long long *__restrict__ f_acc,
                                   T *__restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int a = order[k];
#pragma unroll
    for (int d = 0; d < 3; ++d) out[3 * (size_t)a + d] = (T)((double)f_acc[3 * (size_t)k + d] * (1.0 / FIX_SCALE));
}

__global__ void k_wrapped_positions(int n, const double *__restrict__ x_cur, double Lx, double Ly, double Lz,
                                    float *__restrict__ out) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    double L[3] = {Lx, Ly, Lz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double x = x_cur[3 * (size_t)a + d];
        out[3 * (size_t)a + d] = (float)(x - L[d] * rint(x / L[d]));
    }
}

__global__ void k_f64_to_f32(size_t n, const double *__restrict__ in, float *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// Host state handed to a step call (float32, positions wrapped, matrix_id order) against the device
// state (float64, unwrapped): an atom whose float32 image equals the host value keeps its float64
// coordinate, any other takes the host value.  flags[4]: a position changed (integrator caches are then
// stale), flags[5]: a velocity changed, flags[0]: an atom is 2 or more images away (utils/pbc.py:29-34).
__global__ void k_accept_state(int n, const float *__restrict__ x_in, const float *__restrict__ v_in,
                               double *__restrict__ x_cur, double *__restrict__ vel, double Lx, double Ly, double Lz,
                               int *__restrict__ flags) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const double L[3] = {Lx, Ly, Lz};
    bool changed = false, lost = false, vchanged = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const size_t i = 3 * (size_t)a + d;
        if (x_in) {
            const float h = x_in[i];
            const double x = x_cur[i];
            const float w = (float)(x - L[d] * rint(x / L[d]));
            if (w != h) {
                x_cur[i] = (double)h;
                changed = true;
                if (!(fabs(rint((double)h / L[d])) < 2.0)) lost = true;
            }
        }
        if (v_in) {
            const float h = v_in[i];
            if ((float)vel[i] != h) { vel[i] = (double)h; vchanged = true; }
        }
    }
    if (changed) flags[4] = 1;
    if (vchanged) flags[5] = 1;
    if (lost) flags[0] = 1;
}

// wrapped float32 positions and float32 velocities for the host State, one pass
__global__ void k_export_state(int n, const double *__restrict__ x_cur, const double *__restrict__ vel, double Lx,
                               double Ly, double Lz, float *__restrict__ x_out, float *__restrict__ v_out) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const double L[3] = {Lx, Ly, Lz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const size_t i = 3 * (size_t)a + d;
        if (x_out) { const double x = x_cur[i]; x_out[i] = (float)(x - L[d] * rint(x / L[d])); }
        if (v_out) v_out[i] = (float)vel[i];
    }
}

// self + neutralising-background energy of the Ewald sum (host constant, added to the energies read back)
void prepare_pme_constants(mdk_ctx *c) {
    if (c->have_coul && c->alpha > 0 && c->host_tmp.size() >= 2) {
        double V = c->box.Ld[0] * c->box.Ld[1] * c->box.Ld[2];
        c->e_self_bg = -c->k_e * c->alpha / sqrt(M_PI) * c->host_tmp[1] -
                       c->k_e * M_PI * c->host_tmp[0] * c->host_tmp[0] / (2.0 * V * c->alpha * c->alpha);
    }
}

}  // namespace mdk

using namespace mdk;

// Persistent staging for host transfers: a device scratch buffer plus a pinned host mirror, grown
// on demand — no cudaMalloc / cudaFree (implicit device syncs) on the per-step drop-in path.
static int stage_reserve(mdk_ctx *c, size_t bytes) {
    MDK_CUDA(c, c->io_dev.reserve(bytes));
    if (bytes > c->io_host_cap) {
        if (c->io_host) cudaFreeHost(c->io_host);
        c->io_host = nullptr; c->io_host_cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        MDK_CUDA(c, cudaHostAlloc(&c->io_host, want, cudaHostAllocDefault));
        c->io_host_cap = want;
    }
    return MDK_OK;
}
// device scratch -> caller's buffer through the pinned mirror
static int stage_download(mdk_ctx *c, void *out, size_t bytes) {
    MDK_CUDA(c, cudaMemcpyAsync(c->io_host, c->io_dev.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(out, c->io_host, bytes);
    return MDK_OK;
}

static bool is_pinned(const mdk_ctx *c, const void *p, size_t bytes) {
    const char *q = static_cast<const char *>(p);
    for (const auto &b : c->pinned)
        if (q >= b.first && q + bytes <= b.first + b.second) return true;
    return false;
}
// host -> device on the ctx stream, asynchronous: straight from mdk_host_alloc memory, through the
// pinned mirror (at byte offset `off`, which also is the offset into io_dev) otherwise
static int h2d_async(mdk_ctx *c, size_t off, const void *src, size_t bytes) {
    char *dst = reinterpret_cast<char *>(c->io_dev.p) + off;
    if (!is_pinned(c, src, bytes)) {
        memcpy(static_cast<char *>(c->io_host) + off, src, bytes);
        src = static_cast<char *>(c->io_host) + off;
    }
    MDK_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return MDK_OK;
}

#define NEED_CTX(c) \
    if (!(c)) return MDK_ERR_BAD_ARG

template <typename T>
static int download_forces(mdk_ctx *c, T *out) {
    if (!c->nlist_valid || !out) return fail(c, MDK_ERR_NOT_BOUND, "mdk_download_forces before mdk_compute");
    size_t m = (size_t)3 * c->n;
    MDK_TRY(stage_reserve(c, m * sizeof(T)));
    k_unpermute_forces<T><<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->order.p, c->f_acc.p,
                                                                      reinterpret_cast<T *>(c->io_dev.p));
    ++c->n_launches;
    return stage_download(c, out, m * sizeof(T));
}

extern "C" {

int mdk_create(int device, mdk_ctx **out) {
    if (!out) return MDK_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, MDK_ERR_CUDA, "no CUDA device available (%s); mdpy_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(nullptr, MDK_ERR_BAD_ARG, "device %d out of range [0, %d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, MDK_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, MDK_ERR_CUDA, "device %d is sm_%d%d; libmdpyb200 is built for sm_100a only", device,
                    prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return fail(nullptr, MDK_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    mdk_ctx *c = new mdk_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return fail(nullptr, MDK_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    for (auto &ev : c->ev) cudaEventCreate(&ev);
    // the short side-stream kernels (PME mesh chain, bonded terms) go ahead of queued k_pair blocks
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaStreamCreateWithPriority(&c->s_pme, cudaStreamNonBlocking, prio_hi);
    cudaStreamCreateWithPriority(&c->s_aux, cudaStreamNonBlocking, prio_hi);
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_pme, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming);
    if (c->readback.reserve(32) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void **>(&c->pin_words), 64 * sizeof(long long), cudaHostAllocDefault) != cudaSuccess) {
        mdk_destroy(c);
        return fail(nullptr, MDK_ERR_OOM, "cudaMalloc failed in mdk_create");
    }
    // aliases into the read-back block (cap stays 0: they are not owned)
    c->e_acc.p = c->readback.p;
    c->counters.p = reinterpret_cast<int *>(c->readback.p + 16);
    c->flags.p = reinterpret_cast<int *>(c->readback.p + 24);
    cudaMemset(c->readback.p, 0, 32 * sizeof(long long));
    memset(c->pin_words, 0, 64 * sizeof(long long));
    *out = c;
    return MDK_OK;
}

void mdk_destroy(mdk_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    dd_destroy(c);
    comm_destroy(c);
    graph_destroy(c);
    c->dd_blk.release(); c->dd_mark.release();
    if (c->have_plans) { cufftDestroy(c->plan_r2c); cufftDestroy(c->plan_c2r); }
    c->q.release(); c->mass.release(); c->lj4.release(); c->excl.release(); c->p14.release(); c->excl_pairs.release();
    for (auto &b : c->bonded) { b.idx.release(); b.par.release(); }
    c->rigid_trip.release(); c->rigid_flag.release(); c->q64.release(); c->lj64.release();
    c->x_cur.release(); c->x_prev.release(); c->vel.release(); c->f_prev.release();
    c->order.release(); c->inv_order.release(); c->xs.release(); c->xs_ref.release(); c->ljs.release();
    c->excl_s.release(); c->p14_s.release(); c->f_acc.release(); c->readback.release();
    c->cell_key.release(); c->cell_key_sorted.release(); c->idx_tmp.release(); c->cell_start.release();
    c->sort_tmp.release(); c->sort_buf.release(); c->bb_center.release(); c->bb_half.release();
    c->units.release(); c->chunk_j.release(); c->chunk_mask.release(); c->mask_excl.release(); c->mask_14.release();
    c->grid_fix.release(); c->grid_r.release(); c->grid_c.release(); c->influence.release(); c->fft_tw.release();
    c->io_dev.release();
    if (c->io_host) cudaFreeHost(c->io_host);
    if (c->frame_host) cudaFreeHost(c->frame_host);
    c->frame_dev.release();
    if (c->s_io) cudaStreamDestroy(c->s_io);
    for (int k = 0; k < 2; ++k) { if (c->ev_frame_ready[k]) cudaEventDestroy(c->ev_frame_ready[k]); if (c->ev_frame_done[k]) cudaEventDestroy(c->ev_frame_done[k]); }
    if (c->pin_words) cudaFreeHost(c->pin_words);
    for (auto &b : c->pinned) cudaFreeHost(b.first);
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_pme) cudaEventDestroy(c->ev_pme);
    if (c->ev_aux) cudaEventDestroy(c->ev_aux);
    if (c->s_pme) cudaStreamDestroy(c->s_pme);
    if (c->s_aux) cudaStreamDestroy(c->s_aux);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *mdk_last_error(const mdk_ctx *c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int mdk_set_stream(mdk_ctx *c, void *cuda_stream) {
    NEED_CTX(c);
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return MDK_OK;
}

int mdk_set_box(mdk_ctx *c, const double box[3]) {
    NEED_CTX(c);
    if (!box) return fail(c, MDK_ERR_BAD_ARG, "box is NULL");
    for (int a = 0; a < 3; ++a) {
        if (!(box[a] > 0) || !std::isfinite(box[a])) return fail(c, MDK_ERR_BAD_ARG, "box edge %d = %g is not positive", a, box[a]);
        c->box.Ld[a] = box[a];
        c->box.L[a] = (float)box[a];
        c->box.invL[a] = 1.0f / c->box.L[a];
    }
    c->have_box = true;
    c->pme_dirty = true;
    invalidate(c);
    return MDK_OK;
}

int mdk_set_atoms(mdk_ctx *c, int n, const float *charges, const float *masses) {
    NEED_CTX(c);
    if (n <= 0 || !charges || !masses) return fail(c, MDK_ERR_BAD_ARG, "mdk_set_atoms: n=%d / NULL arrays", n);
    cudaSetDevice(c->device);
    c->have_q64 = false;
    if (n != c->n) {
        c->have_lj64 = false;
        c->have_pos = false; c->have_lj = false; c->wb = c->ws = 0;
        c->verlet_cached = c->langevin_cached = false;
        for (auto &b : c->bonded) b.n = 0;
    }
    c->n = n;
    MDK_CUDA(c, c->q.reserve(n)); MDK_CUDA(c, c->mass.reserve(n));
    MDK_CUDA(c, c->x_cur.reserve((size_t)3 * n)); MDK_CUDA(c, c->vel.reserve((size_t)3 * n));
    MDK_CUDA(c, cudaMemcpyAsync(c->q.p, charges, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    MDK_CUDA(c, cudaMemcpyAsync(c->mass.p, masses, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    MDK_CUDA(c, cudaMemsetAsync(c->vel.p, 0, (size_t)3 * n * sizeof(double), c->stream));
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    double sq = 0, sq2 = 0;
    for (int i = 0; i < n; ++i) { sq += charges[i]; sq2 += (double)charges[i] * charges[i]; }
    c->host_tmp.assign({sq, sq2});
    invalidate(c);
    return MDK_OK;
}

int mdk_set_lj(mdk_ctx *c, const float *eps_sigma, float rc, float r_switch) {
    NEED_CTX(c);
    if (c->n <= 0) return fail(c, MDK_ERR_NOT_BOUND, "mdk_set_lj before mdk_set_atoms");
    if (!eps_sigma) return fail(c, MDK_ERR_BAD_ARG, "eps_sigma is NULL");
    if (!(rc > 0)) return fail(c, MDK_ERR_CUTOFF_TOO_LARGE, "Cutoff radius is poor defined, current value %.3f", rc);
    MDK_CUDA(c, c->lj4.reserve(c->n));
    MDK_CUDA(c, cudaMemcpyAsync(c->lj4.p, eps_sigma, (size_t)c->n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    c->rc_lj = rc;
    c->r_switch = (r_switch > 0 && r_switch < rc) ? r_switch : rc;
    c->have_lj = true;
    invalidate(c);
    return MDK_OK;
}

int mdk_set_precision(mdk_ctx *c, int double_precision) {
    NEED_CTX(c);
    c->dprec = double_precision != 0;
    c->verlet_cached = false; c->langevin_cached = false;
    ++c->graph_epoch;
    return MDK_OK;
}

int mdk_set_params_f64(mdk_ctx *c, const double *charges, const double *eps_sigma) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (c->n <= 0) return fail(c, MDK_ERR_NOT_BOUND, "mdk_set_params_f64 before mdk_set_atoms");
    if (charges) {
        MDK_CUDA(c, c->q64.reserve(c->n));
        MDK_CUDA(c, cudaMemcpy(c->q64.p, charges, (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice));
        c->have_q64 = true;
        double sq = 0, sq2 = 0;
        for (int i = 0; i < c->n; ++i) { sq += charges[i]; sq2 += charges[i] * charges[i]; }
        c->host_tmp.assign({sq, sq2});
    }
    if (eps_sigma) {
        MDK_CUDA(c, c->lj64.reserve((size_t)4 * c->n));
        MDK_CUDA(c, cudaMemcpy(c->lj64.p, eps_sigma, (size_t)4 * c->n * sizeof(double), cudaMemcpyHostToDevice));
        c->have_lj64 = true;
    }
    c->verlet_cached = false; c->langevin_cached = false;
    return MDK_OK;
}

int mdk_set_exclusions(mdk_ctx *c, const int32_t *bonded, int wb, const int32_t *scaling, int ws) {
    NEED_CTX(c);
    if (c->n <= 0) return fail(c, MDK_ERR_NOT_BOUND, "mdk_set_exclusions before mdk_set_atoms");
    if (wb < 0 || ws < 0 || (wb > 0 && !bonded) || (ws > 0 && !scaling))
        return fail(c, MDK_ERR_BAD_ARG, "mdk_set_exclusions: bad widths / NULL tables");
    for (size_t i = 0; i < (size_t)c->n * wb; ++i)
        if (bonded[i] < -1 || bonded[i] >= c->n) return fail(c, MDK_ERR_BAD_ARG, "bonded_particles entry %d out of range", bonded[i]);
    for (size_t i = 0; i < (size_t)c->n * ws; ++i)
        if (scaling[i] < -1 || scaling[i] >= c->n) return fail(c, MDK_ERR_BAD_ARG, "scaling_particles entry %d out of range", scaling[i]);
    c->wb = wb; c->ws = ws;
    {   // compact list of the excluded pairs for the Ewald correction kernel (one thread per pair instead of one per table entry)
        std::vector<int2> pairs;
        for (int i = 0; i < c->n; ++i)
            for (int e = 0; e < wb; ++e) {
                const int p = bonded[(size_t)i * wb + e];
                if (p > i) pairs.push_back(make_int2(i, p));
            }
        c->n_excl_pairs = (int)pairs.size();
        if (!pairs.empty()) {
            MDK_CUDA(c, c->excl_pairs.reserve(pairs.size()));
            MDK_CUDA(c, cudaMemcpy(c->excl_pairs.p, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice));
        }
    }
    if (wb > 0) {
        MDK_CUDA(c, c->excl.reserve((size_t)c->n * wb));
        MDK_CUDA(c, cudaMemcpyAsync(c->excl.p, bonded, (size_t)c->n * wb * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    if (ws > 0) {
        MDK_CUDA(c, c->p14.reserve((size_t)c->n * ws));
        MDK_CUDA(c, cudaMemcpyAsync(c->p14.p, scaling, (size_t)c->n * ws * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    invalidate(c);
    return MDK_OK;
}

int mdk_set_coulomb(mdk_ctx *c, double k_e, double alpha, float rc) {
    NEED_CTX(c);
    if (!(k_e > 0) || alpha < 0 || rc < 0) return fail(c, MDK_ERR_BAD_ARG, "mdk_set_coulomb: k_e=%g alpha=%g rc=%g", k_e, alpha, rc);
    c->k_e = k_e; c->alpha = alpha; c->rc_coul = rc;
    c->have_coul = true;
    c->pme_dirty = true;
    invalidate(c);
    return MDK_OK;
}

int mdk_set_pme(mdk_ctx *c, int nx, int ny, int nz, int order) {
    NEED_CTX(c);
    if (nx < 4 || ny < 4 || nz < 4) return fail(c, MDK_ERR_BAD_ARG, "PME mesh %dx%dx%d too small", nx, ny, nz);
    if (order != 4 && order != 5 && order != 6 && order != 8) return fail(c, MDK_ERR_BAD_ARG, "PME order %d not supported (4, 5, 6, 8)", order);
    if (nx < order || ny < order || nz < order) return fail(c, MDK_ERR_BAD_ARG, "PME mesh smaller than the spline order");
    if (c->have_pme && c->pme_n[0] == nx && c->pme_n[1] == ny && c->pme_n[2] == nz && c->pme_order == order) return MDK_OK;
    c->pme_n[0] = nx; c->pme_n[1] = ny; c->pme_n[2] = nz; c->pme_order = order;
    c->have_pme = true; c->pme_dirty = true;
    c->verlet_cached = false; c->langevin_cached = false;
    ++c->graph_epoch;   // mesh sizes and buffers are baked into captured launches
    return MDK_OK;
}

int mdk_set_nlist(mdk_ctx *c, float skin) {
    NEED_CTX(c);
    if (skin < 0) return fail(c, MDK_ERR_BAD_ARG, "negative skin");
    c->skin = skin;
    invalidate(c);
    return MDK_OK;
}

int mdk_set_bonded(mdk_ctx *c, int kind, int n, const int32_t *idx, const float *par) {
    NEED_CTX(c);
    static const int ni[4] = {2, 3, 4, 4}, np[4] = {2, 4, 3, 2};
    if (kind < 0 || kind > 3 || n < 0 || (n > 0 && (!idx || !par))) return fail(c, MDK_ERR_BAD_ARG, "mdk_set_bonded: bad arguments");
    if (c->n <= 0) return fail(c, MDK_ERR_NOT_BOUND, "mdk_set_bonded before mdk_set_atoms");
    for (size_t i = 0; i < (size_t)n * ni[kind]; ++i)
        if (idx[i] < 0 || idx[i] >= c->n) return fail(c, MDK_ERR_BAD_ARG, "bonded term index %d out of range", idx[i]);
    auto &b = c->bonded[kind];
    b.n = n;
    if (n > 0) {
        MDK_CUDA(c, b.idx.reserve((size_t)n * ni[kind])); MDK_CUDA(c, b.par.reserve((size_t)n * np[kind]));
        MDK_CUDA(c, cudaMemcpyAsync(b.idx.p, idx, (size_t)n * ni[kind] * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        MDK_CUDA(c, cudaMemcpyAsync(b.par.p, par, (size_t)n * np[kind] * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    c->verlet_cached = false; c->langevin_cached = false;
    ++c->graph_epoch;   // term counts and table addresses are baked into captured launches
    return MDK_OK;
}

int mdk_set_rigid_waters(mdk_ctx *c, int n_waters, const int32_t *triplets, double d_oh, double d_hh) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (c->n <= 0) return fail(c, MDK_ERR_NOT_BOUND, "mdk_set_rigid_waters before mdk_set_atoms");
    if (n_waters < 0 || (n_waters > 0 && (!triplets || !(d_oh > 0) || !(d_hh > 0) || d_hh >= 2 * d_oh)))
        return fail(c, MDK_ERR_BAD_ARG, "mdk_set_rigid_waters: n=%d d_oh=%g d_hh=%g", n_waters, d_oh, d_hh);
    c->n_rigid = 0;
    c->verlet_cached = false; c->langevin_cached = false;
    ++c->graph_epoch;              // the water kernel is (or is no longer) part of the captured step
    if (n_waters == 0) return MDK_OK;
    std::vector<float> mass(c->n);
    MDK_CUDA(c, cudaMemcpy(mass.data(), c->mass.p, c->n * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<unsigned char> flag(c->n, 0);
    for (int w = 0; w < n_waters; ++w) {
        for (int t = 0; t < 3; ++t) {
            const int a = triplets[3 * w + t];
            if (a < 0 || a >= c->n || flag[a]) return fail(c, MDK_ERR_BAD_ARG, "mdk_set_rigid_waters: atom %d out of range or in two molecules", a);
            flag[a] = 1;
        }
        // one geometry for all molecules: same oxygen / hydrogen masses everywhere
        const int o = triplets[3 * w], h1 = triplets[3 * w + 1], h2 = triplets[3 * w + 2];
        if (mass[o] != mass[triplets[0]] || mass[h1] != mass[triplets[1]] || mass[h2] != mass[triplets[1]])
            return fail(c, MDK_ERR_BAD_ARG, "mdk_set_rigid_waters: molecule %d has other masses than molecule 0", w);
    }
    MDK_CUDA(c, c->rigid_trip.reserve((size_t)3 * n_waters));
    MDK_CUDA(c, c->rigid_flag.reserve(c->n));
    MDK_CUDA(c, cudaMemcpy(c->rigid_trip.p, triplets, (size_t)3 * n_waters * sizeof(int), cudaMemcpyHostToDevice));
    MDK_CUDA(c, cudaMemcpy(c->rigid_flag.p, flag.data(), c->n, cudaMemcpyHostToDevice));
    c->rigid_d_oh = d_oh; c->rigid_d_hh = d_hh;
    c->rigid_m_o = mass[triplets[0]]; c->rigid_m_h = mass[triplets[1]];
    c->n_rigid = n_waters;
    c->rigid_dirty = true;         // the geometry is projected onto the constraints before the next step call
    return MDK_OK;
}

// ---- state ----
static int check_lost(mdk_ctx *c, const double *x, const float *xf) {
    // utils/pbc.py:29-34: |round(x / L)| >= 2 on any axis
    for (int i = 0; i < c->n; ++i)
        for (int d = 0; d < 3; ++d) {
            double v = xf ? (double)xf[3 * (size_t)i + d] : x[3 * (size_t)i + d];
            if (!(std::fabs(std::nearbyint(v / c->box.Ld[d])) < 2))
                return fail(c, MDK_ERR_PARTICLE_LOST, "Atom(s) with matrix id: [%d] moved beyond 2 PBC image.", i);
        }
    return MDK_OK;
}

int mdk_upload_positions(mdk_ctx *c, const float *xyz) {
    NEED_CTX(c);
    if (c->n <= 0 || !c->have_box) return fail(c, MDK_ERR_NOT_BOUND, "mdk_upload_positions before box/atoms");
    if (!xyz) return fail(c, MDK_ERR_BAD_ARG, "xyz is NULL");
    MDK_TRY(check_lost(c, nullptr, xyz));
    size_t m = (size_t)3 * c->n;
    MDK_TRY(stage_reserve(c, m * sizeof(float)));
    memcpy(c->io_host, xyz, m * sizeof(float));
    float *stage = reinterpret_cast<float *>(c->io_dev.p);
    MDK_CUDA(c, cudaMemcpyAsync(stage, c->io_host, m * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    k_f32_to_f64<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(m, stage, c->x_cur.p);
    ++c->n_launches;
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_pos = true;
    c->xs_current = false;
    c->verlet_cached = false; c->langevin_cached = false;
    c->rigid_dirty = true;
    return MDK_OK;
}

int mdk_upload_positions_f64(mdk_ctx *c, const double *xyz) {
    NEED_CTX(c);
    if (c->n <= 0 || !c->have_box) return fail(c, MDK_ERR_NOT_BOUND, "mdk_upload_positions before box/atoms");
    if (!xyz) return fail(c, MDK_ERR_BAD_ARG, "xyz is NULL");
    MDK_TRY(check_lost(c, xyz, nullptr));
    MDK_CUDA(c, cudaMemcpyAsync(c->x_cur.p, xyz, (size_t)3 * c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_pos = true;
    c->xs_current = false;
    c->verlet_cached = false; c->langevin_cached = false;
    c->rigid_dirty = true;
    return MDK_OK;
}

int mdk_upload_velocities(mdk_ctx *c, const float *v) {
    NEED_CTX(c);
    if (c->n <= 0) return fail(c, MDK_ERR_NOT_BOUND, "mdk_upload_velocities before mdk_set_atoms");
    if (!v) return fail(c, MDK_ERR_BAD_ARG, "v is NULL");
    size_t m = (size_t)3 * c->n;
    MDK_CUDA(c, c->f_prev.reserve(m));
    float *stage = reinterpret_cast<float *>(c->f_prev.p);
    MDK_CUDA(c, cudaMemcpyAsync(stage, v, m * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    k_f32_to_f64<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(m, stage, c->vel.p);
    ++c->n_launches;
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    c->verlet_cached = false; c->langevin_cached = false;
    return MDK_OK;
}

int mdk_download_positions(mdk_ctx *c, float *out) {
    NEED_CTX(c);
    if (!c->have_pos || !out) return fail(c, MDK_ERR_NOT_BOUND, "no positions on the device");
    size_t m = (size_t)3 * c->n;
    MDK_TRY(stage_reserve(c, m * sizeof(float)));
    k_wrapped_positions<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->x_cur.p, c->box.Ld[0], c->box.Ld[1], c->box.Ld[2],
                                                                    reinterpret_cast<float *>(c->io_dev.p));
    ++c->n_launches;
    return stage_download(c, out, m * sizeof(float));
}

int mdk_download_positions_f64(mdk_ctx *c, double *out) {
    NEED_CTX(c);
    if (!c->have_pos || !out) return fail(c, MDK_ERR_NOT_BOUND, "no positions on the device");
    MDK_CUDA(c, cudaMemcpyAsync(out, c->x_cur.p, (size_t)3 * c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    return MDK_OK;
}

int mdk_download_velocities(mdk_ctx *c, float *out) {
    NEED_CTX(c);
    if (c->n <= 0 || !out) return fail(c, MDK_ERR_NOT_BOUND, "no velocities on the device");
    size_t m = (size_t)3 * c->n;
    MDK_TRY(stage_reserve(c, m * sizeof(float)));
    k_f64_to_f32<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(m, c->vel.p, reinterpret_cast<float *>(c->io_dev.p));
    ++c->n_launches;
    return stage_download(c, out, m * sizeof(float));
}

// ---- hot path ----
int mdk_build_nlist(mdk_ctx *c, int64_t *stats) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    MDK_TRY(nlist_rebuild(c));
    if (stats) {
        stats[0] = c->n_blocks; stats[1] = c->stat_units; stats[2] = c->stat_chunks; stats[3] = c->stat_masks;
        stats[4] = c->stat_chunks * 1024;
    }
    return MDK_OK;
}

int mdk_compute(mdk_ctx *c, unsigned terms, double *energies) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    prepare_pme_constants(c);
    if (c->profiling) { for (auto &p : c->phase_ms) p = 0; cudaEventRecord(c->ev[2 * PH_TOTAL], c->stream); }
    MDK_TRY(compute_terms(c, terms, true));
    if (c->profiling) {
        cudaEventRecord(c->ev[2 * PH_TOTAL + 1], c->stream);
        cudaEventSynchronize(c->ev[2 * PH_TOTAL + 1]);
        float ms = 0; cudaEventElapsedTime(&ms, c->ev[2 * PH_TOTAL], c->ev[2 * PH_TOTAL + 1]);
        c->phase_ms[PH_TOTAL] = ms;
    }
    if (energies) memcpy(energies, c->last_e, sizeof(c->last_e));
    return MDK_OK;
}

int mdk_download_forces(mdk_ctx *c, float *out) { NEED_CTX(c); cudaSetDevice(c->device); return download_forces<float>(c, out); }
int mdk_download_forces_f64(mdk_ctx *c, double *out) { NEED_CTX(c); cudaSetDevice(c->device); return download_forces<double>(c, out); }

int mdk_step_verlet(mdk_ctx *c, double dt, int nsteps, unsigned terms, int reference_quirks) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (!(dt > 0)) return fail(c, MDK_ERR_BAD_ARG, "dt must be positive");
    prepare_pme_constants(c);
    if (c->profiling) { for (auto &p : c->phase_ms) p = 0; cudaEventRecord(c->ev[2 * PH_TOTAL], c->stream); }
    MDK_TRY(integrate_verlet(c, dt, nsteps, terms, reference_quirks));
    if (c->profiling) {
        cudaEventRecord(c->ev[2 * PH_TOTAL + 1], c->stream);
        cudaEventSynchronize(c->ev[2 * PH_TOTAL + 1]);
        float ms = 0; cudaEventElapsedTime(&ms, c->ev[2 * PH_TOTAL], c->ev[2 * PH_TOTAL + 1]);
        c->phase_ms[PH_TOTAL] = ms;
    }
    return MDK_OK;
}

// A new integrator object (or erase_cache) starts from scratch: no cached force / previous position, and the
// Langevin noise counter (atom, step) restarts at step 0.
void mdk_verlet_reset(mdk_ctx *c) { if (c) { c->verlet_cached = false; c->langevin_cached = false; c->langevin_step = 0; } }

int mdk_step_langevin(mdk_ctx *c, double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (!(dt > 0) || kT < 0 || gamma < 0) return fail(c, MDK_ERR_BAD_ARG, "mdk_step_langevin: dt=%g kT=%g gamma=%g", dt, kT, gamma);
    prepare_pme_constants(c);
    if (c->rigid_dirty && c->n_rigid > 0 && c->have_pos) { MDK_TRY(rigid_project(c)); c->rigid_dirty = false; }
    if (c->profiling) { for (auto &p : c->phase_ms) p = 0; cudaEventRecord(c->ev[2 * PH_TOTAL], c->stream); }
    MDK_TRY(integrate_langevin(c, dt, kT, gamma, seed, nsteps, terms, 5, false));
    if (c->profiling) {
        cudaEventRecord(c->ev[2 * PH_TOTAL + 1], c->stream);
        cudaEventSynchronize(c->ev[2 * PH_TOTAL + 1]);
        float ms = 0; cudaEventElapsedTime(&ms, c->ev[2 * PH_TOTAL], c->ev[2 * PH_TOTAL + 1]);
        c->phase_ms[PH_TOTAL] = ms;
    }
    return MDK_OK;
}

int mdk_minimize_sd(mdk_ctx *c, double alpha, double energy_tolerance, int max_iterations, unsigned terms, int *iterations,
                    double *energy_first_prev_last, double *energies) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (!(alpha > 0) || energy_tolerance < 0 || max_iterations < 0 || !iterations || !energy_first_prev_last)
        return fail(c, MDK_ERR_BAD_ARG, "mdk_minimize_sd: alpha=%g tolerance=%g max_iterations=%d", alpha, energy_tolerance, max_iterations);
    prepare_pme_constants(c);
    *iterations = 0;
    MDK_TRY(minimize_sd(c, alpha, energy_tolerance, max_iterations, terms, iterations, energy_first_prev_last, energy_first_prev_last + 1,
                        energy_first_prev_last + 2));
    if (energies) memcpy(energies, c->last_e, sizeof(c->last_e));
    return MDK_OK;
}

int mdk_set_frame_capture(mdk_ctx *c, int stride, int max_frames) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (stride < 0 || max_frames < 0 || (stride > 0 && max_frames == 0)) return fail(c, MDK_ERR_BAD_ARG, "mdk_set_frame_capture(%d, %d)", stride, max_frames);
    if (c->s_io) cudaStreamSynchronize(c->s_io);
    c->frame_stride = 0; c->frame_count = 0; c->frame_total = 0;
    if (stride == 0) return MDK_OK;
    if (c->n <= 0) return fail(c, MDK_ERR_NOT_BOUND, "mdk_set_frame_capture before mdk_set_atoms");
    const size_t m = (size_t)3 * c->n;
    if (!c->s_io) {
        MDK_CUDA(c, cudaStreamCreateWithFlags(&c->s_io, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            MDK_CUDA(c, cudaEventCreateWithFlags(&c->ev_frame_ready[k], cudaEventDisableTiming));
            MDK_CUDA(c, cudaEventCreateWithFlags(&c->ev_frame_done[k], cudaEventDisableTiming));
        }
    }
    MDK_CUDA(c, c->frame_dev.reserve(2 * m));
    if (max_frames > c->frame_cap) {
        if (c->frame_host) cudaFreeHost(c->frame_host);
        c->frame_host = nullptr; c->frame_cap = 0;
        MDK_CUDA(c, cudaHostAlloc(reinterpret_cast<void **>(&c->frame_host), (size_t)max_frames * m * sizeof(float), cudaHostAllocDefault));
        c->frame_cap = max_frames;
    }
    c->frame_stride = stride;
    ++c->graph_epoch;
    return MDK_OK;
}

int mdk_get_frames(mdk_ctx *c, float *out, int max_frames, int *n_frames) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (!n_frames || max_frames < 0 || (max_frames > 0 && !out)) return fail(c, MDK_ERR_BAD_ARG, "mdk_get_frames: bad arguments");
    if (c->s_io) MDK_CUDA(c, cudaStreamSynchronize(c->s_io));
    const int k = c->frame_count < max_frames ? c->frame_count : max_frames;
    if (k > 0) memcpy(out, c->frame_host, (size_t)k * 3 * c->n * sizeof(float));
    *n_frames = c->frame_count;
    c->frame_count = 0;            // the ring is handed over: the next call fills it from the start
    return MDK_OK;
}

int mdk_host_alloc(mdk_ctx *c, size_t bytes, void **out) {
    NEED_CTX(c);
    if (!out || bytes == 0) return fail(c, MDK_ERR_BAD_ARG, "mdk_host_alloc: bad arguments");
    cudaSetDevice(c->device);
    void *p = nullptr;
    MDK_CUDA(c, cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    c->pinned.emplace_back(static_cast<char *>(p), bytes);
    *out = p;
    return MDK_OK;
}

int mdk_host_free(mdk_ctx *c, void *p) {
    NEED_CTX(c);
    for (size_t k = 0; k < c->pinned.size(); ++k)
        if (c->pinned[k].first == p) {
            cudaStreamSynchronize(c->stream);
            cudaFreeHost(p);
            c->pinned.erase(c->pinned.begin() + k);
            return MDK_OK;
        }
    return fail(c, MDK_ERR_BAD_ARG, "mdk_host_free: not a block of this context");
}

// Integrator.integrate with the host State as input and output (langevin_integrator.py:38-72 reads
// ensemble.state.positions / velocities and writes them back): one call = host -> device copy of the
// state, nsteps steps, device -> host copy of the new state and the energies.
//
// Steady state (cached force, valid list, graph steps) runs AHEAD of its own change check: the upload,
// the comparison kernel, the steps and the download are queued back to back and the stream is
// synchronised once.  Should the comparison find that the host positions differ from the device's
// (flags[4]) the queued Langevin kernels have left the state alone (k_langevin), and the call is redone
// on the careful path: read the flags first, drop the caches, evaluate f(x_0), then step.
static int host_state_in(mdk_ctx *c, const float *x_in, const float *v_in, size_t m) {
    const size_t mb = m * sizeof(float);
    float *stage = reinterpret_cast<float *>(c->io_dev.p);
    MDK_CUDA(c, cudaMemsetAsync(c->flags.p + 4, 0, 2 * sizeof(int), c->stream));
    if (x_in && !c->have_pos)
        MDK_CUDA(c, cudaMemsetAsync(c->x_cur.p, 0xff, m * sizeof(double), c->stream));  // NaN pattern: everything "differs"
    if (x_in && v_in && v_in == x_in + m && is_pinned(c, x_in, 2 * mb)) {
        MDK_CUDA(c, cudaMemcpyAsync(stage, x_in, 2 * mb, cudaMemcpyHostToDevice, c->stream));   // one block, one copy
    } else {
        if (x_in) MDK_TRY(h2d_async(c, 0, x_in, mb));
        if (v_in) MDK_TRY(h2d_async(c, mb, v_in, mb));
    }
    k_accept_state<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, x_in ? stage : nullptr, v_in ? stage + m : nullptr,
                                                              c->x_cur.p, c->vel.p, c->box.Ld[0], c->box.Ld[1],
                                                              c->box.Ld[2], c->flags.p);
    ++c->n_launches;
    return MDK_OK;
}

static int host_state_out(mdk_ctx *c, float *x_out, float *v_out, size_t m, float **px, float **pv) {
    const size_t mb = m * sizeof(float);
    float *stage = reinterpret_cast<float *>(c->io_dev.p);
    *px = *pv = nullptr;
    if (!x_out && !v_out) return MDK_OK;
    k_export_state<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->x_cur.p, c->vel.p, c->box.Ld[0], c->box.Ld[1],
                                                              c->box.Ld[2], x_out ? stage + 2 * m : nullptr,
                                                              v_out ? stage + 3 * m : nullptr);
    ++c->n_launches;
    if (x_out && v_out && v_out == x_out + m && is_pinned(c, x_out, 2 * mb)) {
        MDK_CUDA(c, cudaMemcpyAsync(x_out, stage + 2 * m, 2 * mb, cudaMemcpyDeviceToHost, c->stream));
        *px = x_out; *pv = v_out;
        return MDK_OK;
    }
    if (x_out) {
        *px = is_pinned(c, x_out, mb) ? x_out : reinterpret_cast<float *>(static_cast<char *>(c->io_host) + 2 * mb);
        MDK_CUDA(c, cudaMemcpyAsync(*px, stage + 2 * m, mb, cudaMemcpyDeviceToHost, c->stream));
    }
    if (v_out) {
        *pv = is_pinned(c, v_out, mb) ? v_out : reinterpret_cast<float *>(static_cast<char *>(c->io_host) + 3 * mb);
        MDK_CUDA(c, cudaMemcpyAsync(*pv, stage + 3 * m, mb, cudaMemcpyDeviceToHost, c->stream));
    }
    return MDK_OK;
}

int mdk_step_langevin_host(mdk_ctx *c, const float *x_in, const float *v_in, float *x_out, float *v_out, double dt,
                           double kT, double gamma, uint64_t seed, int nsteps, unsigned terms, double *energies) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (!(dt > 0) || kT < 0 || gamma < 0 || nsteps < 0) return fail(c, MDK_ERR_BAD_ARG, "mdk_step_langevin_host: dt=%g kT=%g gamma=%g nsteps=%d", dt, kT, gamma, nsteps);
    if (c->n <= 0 || !c->have_box) return fail(c, MDK_ERR_NOT_BOUND, "mdk_step_langevin_host before box/atoms");
    if (!x_in && !c->have_pos) return fail(c, MDK_ERR_NOT_BOUND, "no positions on the device and none passed in");
    prepare_pme_constants(c);
    const size_t m = (size_t)3 * c->n, mb = m * sizeof(float);
    MDK_TRY(stage_reserve(c, 4 * mb));
    if (c->profiling) { for (auto &p : c->phase_ms) p = 0; cudaEventRecord(c->ev[2 * PH_TOTAL], c->stream); }
    const int *h_flags = reinterpret_cast<const int *>(c->pin_words + 24);
    const uint64_t step0 = c->langevin_step;
    bool ahead = (x_in || v_in) && nsteps > 0 && c->have_pos && c->langevin_cached && c->nlist_valid && c->use_graph &&
                 c->profiling < 2 && !c->dd;
    float *px = nullptr, *pv = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        if ((x_in || v_in) && pass == 0) MDK_TRY(host_state_in(c, x_in, v_in, m));
        if ((x_in || v_in) && !ahead) {
            // careful path: learn what changed before queueing anything that depends on it
            MDK_CUDA(c, cudaMemcpyAsync(c->pin_words, c->readback.p, 28 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            MDK_CUDA(c, cudaStreamSynchronize(c->stream));
            const bool lost = h_flags[0] != 0, changed = h_flags[4] != 0;
            MDK_CUDA(c, cudaMemsetAsync(c->flags.p + 4, 0, 2 * sizeof(int), c->stream));
            if (lost) {
                cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream);
                c->have_pos = false;
                c->verlet_cached = false; c->langevin_cached = false;
                return fail(c, MDK_ERR_PARTICLE_LOST, "Atom(s) moved beyond 2 PBC image.");
            }
            if (changed || !c->have_pos) {
                c->have_pos = true;
                c->xs_current = false;
                c->verlet_cached = false; c->langevin_cached = false;
                c->rigid_dirty = true;
            }
        }
        if (c->rigid_dirty && c->n_rigid > 0 && c->have_pos && !ahead) { MDK_TRY(rigid_project(c)); c->rigid_dirty = false; }
        if (nsteps > 0) MDK_TRY(integrate_langevin(c, dt, kT, gamma, seed, nsteps, terms, 1, true));
        MDK_TRY(host_state_out(c, x_out, v_out, m, &px, &pv));
        if (c->profiling) cudaEventRecord(c->ev[2 * PH_TOTAL + 1], c->stream);
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (!ahead || !(h_flags[4] | h_flags[0])) break;
        // ran ahead and lost the bet: nothing was advanced; redo with the flags in hand
        c->graph_pending = 0;
        c->langevin_step = step0;
        ahead = false;
    }
    if (c->profiling) {
        float ms = 0; cudaEventElapsedTime(&ms, c->ev[2 * PH_TOTAL], c->ev[2 * PH_TOTAL + 1]);
        c->phase_ms[PH_TOTAL] = ms;
    }
    if (x_out && px != x_out) memcpy(x_out, px, mb);
    if (v_out && pv != v_out) memcpy(v_out, pv, mb);
    if (nsteps > 0) {
        energies_finish(c, terms);
        MDK_TRY(graph_finish(c));
    }
    if (energies) memcpy(energies, c->last_e, sizeof(c->last_e));
    return MDK_OK;
}

int mdk_last_energies(mdk_ctx *c, double *energies) {
    NEED_CTX(c);
    if (!energies) return fail(c, MDK_ERR_BAD_ARG, "energies is NULL");
    memcpy(energies, c->last_e, sizeof(c->last_e));
    return MDK_OK;
}

int mdk_get_pairs(mdk_ctx *c, int32_t *out_i, int32_t *out_j, int64_t cap, int64_t *n_out) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (!n_out || cap < 0 || (cap > 0 && (!out_i || !out_j))) return fail(c, MDK_ERR_BAD_ARG, "mdk_get_pairs: bad arguments");
    if (!c->have_box || c->n <= 0 || !c->have_pos) return fail(c, MDK_ERR_NOT_BOUND, "mdk_get_pairs before box/atoms/positions");
    return pair_enumerate(c, out_i, out_j, cap, n_out, 0);
}

int mdk_get_pairs_production(mdk_ctx *c, int32_t *out_i, int32_t *out_j, int64_t cap, int64_t *n_out) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    if (!n_out || cap < 0 || (cap > 0 && (!out_i || !out_j))) return fail(c, MDK_ERR_BAD_ARG, "mdk_get_pairs_production: bad arguments");
    if (!c->have_box || c->n <= 0 || !c->have_pos) return fail(c, MDK_ERR_NOT_BOUND, "mdk_get_pairs_production before box/atoms/positions");
    return pair_enumerate(c, out_i, out_j, cap, n_out, 1);
}

int mdk_get_timing(mdk_ctx *c, double *out24) {
    NEED_CTX(c);
    if (!out24) return fail(c, MDK_ERR_BAD_ARG, "out is NULL");
    for (int k = 0; k < 24; ++k) out24[k] = 0;
    for (int k = 0; k <= PH_TOTAL; ++k) out24[k] = c->phase_ms[k];
    out24[9] = c->phase_ms[PH_COMM];
    out24[10] = (double)c->n_launches; out24[11] = (double)c->n_rebuilds; out24[12] = (double)c->n_pair_launches;
    out24[13] = (double)c->stat_units; out24[14] = (double)c->stat_chunks; out24[15] = (double)c->stat_masks;
    out24[16] = (double)c->seg_chunks; out24[17] = (double)c->n_blocks; out24[18] = c->shift_ok ? 1.0 : 0.0;
    return MDK_OK;
}

int mdk_set_profiling(mdk_ctx *c, int level) { NEED_CTX(c); c->profiling = level < 0 ? 0 : level; return MDK_OK; }

int mdk_force_accumulator(mdk_ctx *c, void **dev_ptr, int64_t *n_int64) {
    NEED_CTX(c);
    if (!dev_ptr || !n_int64) return fail(c, MDK_ERR_BAD_ARG, "NULL output");
    *dev_ptr = c->f_acc.p;
    *n_int64 = (int64_t)c->n_pad * 3;
    return MDK_OK;
}

int mdk_set_option(mdk_ctx *c, int key, double value) {
    NEED_CTX(c);
    switch (key) {
        case 0: c->use_graph = value != 0; break;          // CUDA-graph steps in the integrators
        case 1: c->concurrent = value != 0; break;         // PME / bonded on side streams
        case 2: c->force_canonical = value != 0; break;    // per-pair canonical minimum image even in large boxes
        case 3: c->graph_energy = value != 0; break;       // energy sums in every graph step
        case 4: c->graph_nccl = value != 0; break;         // N > 1: NCCL all-reduce inside the captured step (hung in round 1)
        case 5: c->pair_blocks_per_sm = value < 1 ? 1 : (value > 8 ? 8 : (int)value); break;
        case 6: c->pme_force_cufft = value != 0; c->pme_dirty = true; break;   // cuFFT also for small power-of-two meshes
        case 7: c->spread_smem = value != 0; break;        // shared-memory staged charge spreading (default on)
        case 8: c->pair_v5 = value != 0; break;            // filter-then-compute pair kernel (default off: measured slower)
        case 9: c->unit_waves = value < 0.25 ? 0.25 : value; c->nlist_valid = false; break;   // work units per resident warp (list granularity)
        case 10: c->far_split = value != 0; c->nlist_valid = false; break;   // skin-shell j-atoms last in every block's list
        case 11: c->pair_units_per_warp = value < 0 ? 0 : (int)value; break;   // k_pair: work units per warp before it retires (0 = persistent)
        case 12: c->far_flush = value < 32 ? 32 : (value > 992 ? 992 : (int)value); c->nlist_valid = false; break;   // list builder: far-class staging threshold
        case 13: c->dd_late_spread = value != 0; break;   // decomposed step: the round-2a order (spread after the halo exchange)
        case 14: c->dd_early_recv = value != 0; break;    // decomposed step: post the receive of the potential box before the pair kernel
        default: return fail(c, MDK_ERR_BAD_ARG, "mdk_set_option: unknown key %d", key);
    }
    ++c->graph_epoch;
    return MDK_OK;
}

int mdk_flush_l2(mdk_ctx *c) {
    NEED_CTX(c);
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)256 << 20;  // > 126 MB of L2
    MDK_CUDA(c, c->sort_tmp.reserve(bytes));
    MDK_CUDA(c, cudaMemsetAsync(c->sort_tmp.p, 0xA5, bytes, c->stream));
    return MDK_OK;
}

}  // extern "C"
