// mdk_common.cuh — context, buffers and device helpers shared by all translation units
// of libmdpyb200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mdpy_b200.h"

namespace mdk {

constexpr int WARP = 32;
constexpr int TILE = 32;                  // atoms per i-block and per j-chunk
constexpr double FIX_SCALE = 1099511627776.0;   // 2^40: fixed-point forces (int64)
constexpr float FIX_SCALE_F = 1099511627776.0f;
constexpr float RINT_MAGIC = 12582912.0f;  // 1.5 * 2^23: (x + M) - M == rintf(x) for |x| < 2^22

// ---------------------------------------------------------------------------
// growable device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Box {
    float L[3], invL[3];
    double Ld[3];
};

// ---------------------------------------------------------------------------
// Spatial domain decomposition (mdk_dd.cu).  The cell grid is cut by planes into pdim[0] x pdim[1] x pdim[2]
// domains; cells are numbered domain by domain (x fastest inside a domain), so after the cell sort every
// domain's atoms — and with them every rank's i-blocks — are one contiguous range of the tile order.  One
// domain (pdim = 1,1,1) reproduces the plain x-fastest numbering.
constexpr int DD_MAXP = 4;                 // domains per axis
constexpr int DD_MAXR = DD_MAXP * DD_MAXP * DD_MAXP;
// The cuts are a recursive bisection: planes across x, then across y inside every x slab, then across z inside every
// (x, y) column — so the domain volumes can follow per-rank work weights (the rank that also runs the PME mesh gets a
// smaller domain) while every domain stays a box.
struct DDGeom {
    int pdim[3];
    int cut0[DD_MAXP + 1];                       // x: domain dx covers cells [cut0[dx], cut0[dx+1])
    int cut1[DD_MAXP][DD_MAXP + 1];              // y inside x slab dx
    int cut2[DD_MAXP][DD_MAXP][DD_MAXP + 1];     // z inside column (dx, dy)
    int dom_base[DD_MAXR + 1];                   // first cell key of each domain, [ndom] = number of cells
};

// kernel-side view of the tile list
struct NlistView {
    const int4 *units;      // {i-block, first chunk, n chunks, unused}
    const int *n_units;     // device counter
    const int *chunk_j;     // [chunk][32] tile-order atom indices (padding -> 0 + mask bit)
    const int *chunk_mask;  // [chunk] mask slot or -1
    const unsigned *mask_excl;  // [slot][32] rotated exclusion bits per i-lane
    const unsigned *mask_14;    // [slot][32] rotated 1-4 bits per i-lane
};

struct DDState;

struct BondedSet {
    int n = 0;
    DevBuf<int> idx;
    DevBuf<float> par;
};

enum Phase { PH_NLIST = 0, PH_PAIR, PH_SPREAD, PH_FFT, PH_GATHER, PH_BONDED, PH_INTEGRATE, PH_BARE, PH_TOTAL, PH_COMM, PH_COUNT };

}  // namespace mdk

struct mdk_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    // side streams: the PME mesh chain and the O(N) bonded / excluded-pair kernels run beside k_pair
    cudaStream_t s_pme = nullptr, s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_pme = nullptr, ev_aux = nullptr;
    bool concurrent = true;
    bool pair_v5 = false;                     // filter-then-compute pair kernel (option 8).  Off: measured 415 vs 344 us at 92k, 69 vs 54 us
                                              // at 23k — the per-lane bit counts are too uneven
    int pair_blocks_per_sm = 4;               // persistent k_pair blocks per SM (4 fill the register file; fewer leave room for the side-stream kernels)
    std::string err;

    // ---- system (matrix_id order) ----
    int n = 0;
    mdk::Box box{};
    bool have_box = false;
    mdk::DevBuf<float> q, mass;          // [n]
    mdk::DevBuf<float4> lj4;             // [n] eps, sigma, eps14, sigma14
    // DOUBLE precision (env.set_precision('DOUBLE'), mdpy/environment.py:23-42): float64 copies of the per-atom parameters;
    // the pair, bonded and bare-Coulomb arithmetic then runs in float64 on the float64 positions (x_cur)
    bool dprec = false;
    mdk::DevBuf<double> q64, lj64;       // [n], [n,4]
    bool have_q64 = false, have_lj64 = false;
    bool have_lj = false;
    float rc_lj = 0.f, r_switch = 0.f;
    mdk::DevBuf<int> excl, p14;          // [n, wb] / [n, ws], -1 padded, matrix ids
    int wb = 0, ws = 0;
    mdk::DevBuf<int2> excl_pairs;        // the excluded pairs (a < b, matrix ids) of the bonded_particles table, compact
    int n_excl_pairs = 0;
    double k_e = 0.0, alpha = 0.0;
    float rc_coul = 0.f;
    bool have_coul = false;
    float skin = 2.0f;
    mdk::BondedSet bonded[4];
    // rigid three-site waters (SETTLE): triplets (O, H, H) of matrix ids, a flag per atom for the free-atom kernel
    mdk::DevBuf<int> rigid_trip;
    mdk::DevBuf<unsigned char> rigid_flag;
    int n_rigid = 0;
    double rigid_d_oh = 0, rigid_d_hh = 0, rigid_m_o = 0, rigid_m_h = 0;
    bool rigid_dirty = false;                 // positions came from outside since the last projection onto the constraints

    // ---- state (matrix_id order) ----
    mdk::DevBuf<double> x_cur, x_prev, vel;   // [n,3] unwrapped fp64 positions / velocities
    mdk::DevBuf<double> f_prev;               // [n,3] Langevin: forces of the previous step
    bool have_pos = false;
    bool verlet_cached = false, langevin_cached = false;
    uint64_t langevin_step = 0;

    // ---- tile order ----
    int n_blocks = 0;
    int n_pad = 0;                            // n rounded up to 32
    mdk::DevBuf<int> order, inv_order;        // tile slot -> matrix id, matrix id -> slot
    mdk::DevBuf<float4> xs;                   // wrapped xyz + q*sqrt(k_e)
    mdk::DevBuf<float4> xs_ref;               // xs at the last rebuild
    mdk::DevBuf<float4> ljs;                  // 2 sqrt(eps), sigma/2, 2 sqrt(eps14), sigma14/2
    mdk::DevBuf<int> excl_s, p14_s;           // exclusion tables in tile slots
    mdk::DevBuf<long long> f_acc;             // [n_pad,3] fixed-point forces
    mdk::DevBuf<long long> e_acc;             // [MDK_NUM_ENERGIES] fixed point
    mdk::DevBuf<int> flags;                   // [0]=lost atoms [1]=needs rebuild [2]=pool overflow [3]=sticky graph errors
                                              // [4]=host positions differed from the device state [5]=host velocities differed
    // cell grid
    int ncell[3] = {0, 0, 0};
    float cellw[3] = {0, 0, 0};
    mdk::DevBuf<unsigned> cell_key, cell_key_sorted;
    mdk::DevBuf<int> idx_tmp;
    mdk::DevBuf<int> cell_start;              // [ncells + 1]
    mdk::DevBuf<unsigned char> sort_tmp;      // scratch (L2 flush)
    mdk::DevBuf<unsigned char> sort_buf;      // CUB radix sort temporary storage
    size_t sort_tmp_bytes = 0;
    int sort_end_bit = 1;
    long long n_cells = 0;
    int n_parts = 1;
    bool graph_pools = false;                 // size the list pools for rebuilds that run inside a CUDA graph
    mdk::DevBuf<float4> bb_center, bb_half;   // [n_blocks]
    // tile list pools
    mdk::DevBuf<int4> units;
    mdk::DevBuf<int> chunk_j, chunk_mask;
    mdk::DevBuf<unsigned> mask_excl, mask_14;
    mdk::DevBuf<int> counters;                // [0]=units [1]=chunks [2]=mask slots [3]=work cursor [8..10]=max bbox half extents (float bits)
    size_t cap_units = 0, cap_chunks = 0, cap_masks = 0;
    bool nlist_valid = false;
    bool force_canonical = false;             // test hook: always the per-pair canonical minimum image
    bool shift_ok = false;                    // box large enough to hoist the minimum image out of the pair loop
    int seg_chunks = 8;
    int pair_units_per_warp = 0;              // 0: persistent k_pair blocks; > 0: warps retire after this many work units (option 11)
    bool dd_early_recv = false;               // decomposed step: potential boxes return on the side stream, receives posted before k_pair (option 14)
    bool dd_late_spread = false;              // decomposed step: spread after the halo exchange, sub-meshes in an exchange of their own (option 13)
    int far_flush = 992;                      // far-class staging threshold of the list builder (option 12; tests lower it)
    bool far_split = true;                    // list order: skin-shell j-atoms in chunks of their own (option 10)
    double unit_waves = 8.0;                  // work units per resident warp the list planner aims for (option 9): the tail of a pair launch
                                              // is one unit long; 92k box: 347 us at 2 (12 chunks per unit), 305 us at 8 (3 chunks)
    int64_t stat_units = 0, stat_chunks = 0, stat_masks = 0;
    void *nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    mdk::DDState *dd = nullptr;               // spatial domain decomposition (mdk_dd.cu); null = single domain
    mdk::DDGeom dd_geom{};                    // cell numbering (one domain unless dd is set)
    double dd_weight[mdk::DD_MAXR] = {0};     // relative pair-work share of every rank's domain (0 = equal)
    mdk::DevBuf<int> dd_blk;                  // [ndom + 1] first i-block of each domain (device; written by every rebuild)
    mdk::DevBuf<int> dd_mark;                 // [n_pad] 1 = tile slot referenced by this rank's work but owned by another
    mdk::DevBuf<int> aux_sel[5];              // decomposed runs: indices of the bonded terms / excluded pairs this rank owns (per kind)
    int aux_sel_n[5] = {-1, -1, -1, -1, -1};  // -1 = no selection (walk the whole list)
    int own_lo = 0, own_hi = -1;              // tile slots this rank owns (integrates, owns the terms of); -1 = all
    int pme_lo = 0, pme_hi = -1;              // tile slots this rank spreads / gathers on the PME mesh: the atoms whose CELL lies in
                                              // its domain (own slots are whole i-blocks; the block that straddles a domain boundary
                                              // holds a few atoms of the next domain, which may sit anywhere on that domain's border)

    // ---- PME ----
    int pme_n[3] = {0, 0, 0};
    int pme_order = 4;
    bool have_pme = false, pme_dirty = true;
    mdk::DevBuf<long long> grid_fix;          // fixed-point charge mesh
    mdk::DevBuf<float> grid_r;                // real mesh
    mdk::DevBuf<float2> grid_c;               // half spectrum
    mdk::DevBuf<float> influence;             // G(m) on the half spectrum
    mdk::DevBuf<float2> fft_tw;               // twiddles of the small-mesh FFT kernels
    bool pme_fast = false;                    // every mesh axis a power of two in 8..64: own fused FFT kernels instead of cuFFT
    bool pme_force_cufft = false;             // test hook: cuFFT also for small power-of-two meshes
    bool spread_smem = true;                  // shared-memory staged charge spreading (option 7; 0 = one global atomic per spline point)
    cufftHandle plan_r2c = 0, plan_c2r = 0;
    bool have_plans = false;
    double e_self_bg = 0.0;

    // ---- CUDA-graph step ----
    bool use_graph = true, in_capture = false;
    bool capture_energy = false;              // the step graph being captured carries the pair-kernel energy sums
    int graph_pending = 0;                    // graph steps queued whose bookkeeping (graph_finish) is still due
    bool graph_pending_hosted = false;        // ... with host-launched force / update kernels (multi-GPU)
    bool graph_nccl = false;                  // capture the per-step ncclAllReduce into the step graph (N > 1): hung at N = 2 in round 1, off
    bool graph_hosted = false;                // N > 1: upkeep graph + host-launched step kernels (no host sync inside a run).  Off by
                                              // default: its round-1 runs were made with upkeep graphs captured before the shard was set
                                              // (fixed since, mdk_set_shard), and the fixed path has not been re-measured on > 1 GPU yet
    bool graph_energy = false;                // energies in every graph step (the energy-less k_pair variant measured 18 % slower at 92k atoms: ptxas schedules it worse)
    bool xs_current = false;                  // tile-order positions already match x_cur (integrator just published them)
    // two instantiations of the step graph: [0] inner steps (pair kernel without energy sums where that
    // is faster), [1] steps whose energies are read back (the last step of a call)
    cudaGraph_t step_graph[2] = {nullptr, nullptr};
    cudaGraphExec_t step_exec[2] = {nullptr, nullptr};
    cudaGraph_t upkeep_graph = nullptr;       // k_decide -> IF { list rebuild }
    cudaGraphExec_t upkeep_exec = nullptr;
    mdk::DevBuf<unsigned long long> step_dev;  // [0] Langevin noise counter, [1] k_langevin mode of the next graph step
    double graph_key[8] = {0};
    uint64_t graph_seed = 0;                  // Philox key baked into the captured k_langevin (kept as an integer: all 64 bits count)
    unsigned cached_terms = 0;                // term set the integrators' cached forces (f_prev / x_prev) were computed with
    long long graph_epoch = 0, graph_epoch_built = -1;
    int graph_launches_per_step = 0;

    // ---- trajectory frames (mdk_set_frame_capture): wrapped float32 positions every `stride` steps of a step call, copied
    // to page-locked host memory by a copy stream while the steps go on ----
    int frame_stride = 0, frame_cap = 0, frame_count = 0;
    long long frame_total = 0;                // frames captured since the capture was switched on
    mdk::DevBuf<float> frame_dev;             // [2][3 n] double-buffered staging
    float *frame_host = nullptr;              // pinned [frame_cap][3 n]
    cudaStream_t s_io = nullptr;
    cudaEvent_t ev_frame_ready[2] = {nullptr, nullptr}, ev_frame_done[2] = {nullptr, nullptr};

    // ---- host transfer staging ----
    mdk::DevBuf<unsigned char> io_dev;
    void *io_host = nullptr;
    size_t io_host_cap = 0;
    // One device block holds everything the host reads back after a call — e_acc (16 x int64), counters
    // (16 x int), flags (8 x int) alias into it — so a single 256-byte copy into pin_words fetches it all.
    mdk::DevBuf<long long> readback;          // [0..15] energies, [16..23] counters, [24..27] flags
    long long *pin_words = nullptr;           // pinned mirror of `readback` (+ scratch words behind it)
    int rebuilds_seen = 0;                    // counters[12] (rebuilds done inside graphs) at the last read-back
    std::vector<std::pair<char *, size_t>> pinned;   // mdk_host_alloc blocks: copied to / from without staging

    // ---- bookkeeping ----
    double last_e[MDK_NUM_ENERGIES] = {0};
    int profiling = 0;                        // 0 off, 1 whole-call events only, 2 per-phase events (adds syncs)
    cudaEvent_t ev[2 * mdk::PH_COUNT] = {nullptr};
    double phase_ms[mdk::PH_COUNT] = {0};
    int64_t n_launches = 0, n_rebuilds = 0, n_pair_launches = 0;
    std::vector<double> host_tmp;
};

namespace mdk {

int fail(mdk_ctx *c, int code, const char *fmt, ...);
#define MDK_CUDA(c, call)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return mdk::fail((c), e__ == cudaErrorMemoryAllocation ? MDK_ERR_OOM : MDK_ERR_CUDA, \
                             "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define MDK_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != MDK_OK) return rc__; \
    } while (0)

struct PhaseTimer {
    mdk_ctx *c; int ph;
    PhaseTimer(mdk_ctx *c_, int ph_) : c(c_), ph(ph_) {
        if (c->profiling >= 2) cudaEventRecord(c->ev[2 * ph], c->stream);
    }
    ~PhaseTimer() {
        if (c->profiling >= 2) {
            cudaEventRecord(c->ev[2 * ph + 1], c->stream);
            cudaEventSynchronize(c->ev[2 * ph + 1]);
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev[2 * ph], c->ev[2 * ph + 1]);
            c->phase_ms[ph] += ms;
        }
    }
};

// translation-unit entry points
int nlist_refresh_sorted(mdk_ctx *c);          // xs <- wrap(x_cur) in tile order + displacement check
int nlist_rebuild(mdk_ctx *c);
int nlist_enqueue(mdk_ctx *c, bool in_graph);  // device work of a rebuild only (capturable)
int nlist_ensure(mdk_ctx *c);                  // rebuild if flagged / invalid
NlistView nlist_view(mdk_ctx *c);
int pair_compute(mdk_ctx *c, bool do_lj, bool do_coul);
int pair_compute_f64(mdk_ctx *c, bool do_lj, bool do_coul);   // DOUBLE precision: plain float64 kernel over the same tile list
int pair_enumerate(mdk_ctx *c, int32_t *out_i, int32_t *out_j, int64_t cap, int64_t *n_out, int production);
int coulomb_bare(mdk_ctx *c);
int pme_prepare(mdk_ctx *c);
int pme_compute(mdk_ctx *c);
int pme_spread(mdk_ctx *c);                     // own atoms -> fixed-point mesh
int pme_mesh(mdk_ctx *c, bool convert);         // (fixed point -> float,) FFT, influence function + energy, inverse FFT
int pme_gather(mdk_ctx *c);                     // potential mesh -> forces on own atoms
inline int pme_first(const mdk_ctx *c) { return c->pme_hi < 0 ? 0 : c->pme_lo; }
inline int pme_end(const mdk_ctx *c) { return c->pme_hi < 0 ? c->n : c->pme_hi; }
inline int own_first(const mdk_ctx *c) { return c->own_hi < 0 ? 0 : c->own_lo; }
inline int own_end(const mdk_ctx *c) { return c->own_hi < 0 ? c->n : (c->own_hi < c->n ? c->own_hi : c->n); }
int bonded_compute(mdk_ctx *c, unsigned terms);   // bonded terms + (with PME_RECIP) the excluded-pair Ewald correction, one launch
int integrate_verlet(mdk_ctx *c, double dt, int nsteps, unsigned terms, int quirks);
int integrate_langevin(mdk_ctx *c, double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms,
                       int graph_min_steps, bool defer_energies);
int minimize_sd(mdk_ctx *c, double alpha, double energy_tolerance, int max_iterations, unsigned terms, int *iterations,
                double *e_first, double *e_prev, double *e_last);
int energies_enqueue(mdk_ctx *c);                    // kinetic energy + all-reduce + D2H into pin_words (no sync)
void energies_finish(mdk_ctx *c, unsigned terms);    // after the stream was synchronised
int compute_terms(mdk_ctx *c, unsigned terms, bool sync_energies);
int forces_enqueue(mdk_ctx *c, unsigned terms, bool clean_on_entry);
void graph_destroy(mdk_ctx *c);
int graph_finish(mdk_ctx *c);                        // counters / sticky errors of a queued graph run, after a sync
int check_lost_flag(mdk_ctx *c);                     // flags[0] in the last read-back block -> MDK_ERR_PARTICLE_LOST
// sptr / rptr, when set, replace buffer base + offset (one grouped exchange over several buffers)
struct Xfer { int peer; size_t soff, sbytes, roff, rbytes; const void *sptr = nullptr; void *rptr = nullptr; };
int comm_exchange(mdk_ctx *c, const void *sbuf, void *rbuf, const Xfer *x, int nx);   // grouped ncclSend / ncclRecv
int comm_allgather_i32(mdk_ctx *c, const int *mine, int *all, int count);
int comm_allreduce_energies(mdk_ctx *c);
void comm_destroy(mdk_ctx *c);
int dd_compute_single(mdk_ctx *c, unsigned terms, bool sync_energies);
int dd_langevin_single(mdk_ctx *c, double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms, bool defer_energies);
void dd_destroy(mdk_ctx *c);
int langevin_launch(mdk_ctx *c, int first, int end, int mode, double dt, double ca, double cb, double tg, uint64_t seed, uint64_t step);
void prepare_pme_constants(mdk_ctx *c);
int rigid_project(mdk_ctx *c);
int frame_capture_enqueue(mdk_ctx *c, int step_in_call);   // no-op unless this step is a multiple of the capture stride

// ---------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__
// Minimum image along one axis with exactly the rounding sequence of the canonical pair
// criterion (oracle/mdpy_oracle.c:ora_pair_set_f32).
__device__ __forceinline__ float min_image(float d, float L, float invL) {
    float t = __fadd_rn(__fmaf_rn(d, invL, RINT_MAGIC), -RINT_MAGIC);
    return __fmaf_rn(-L, t, d);
}
__device__ __forceinline__ float dist2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
__device__ __forceinline__ long long to_fix(float f) { return __float2ll_rn(f * FIX_SCALE_F); }
__device__ __forceinline__ long long to_fix(double f) { return __double2ll_rn(f * FIX_SCALE); }
__device__ __forceinline__ void atomic_add_fix(long long *p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long *>(p), static_cast<unsigned long long>(v));
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace mdk
