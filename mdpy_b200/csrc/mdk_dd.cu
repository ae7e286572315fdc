// mdk_dd.cu — spatial domain decomposition with halo exchange (the reference has a single State on a
// single device, mdpy/core/state.py:18-28, so all of this is new design).
//
// Decomposition.  The orthorhombic box is cut into px x py x pz domains (2 GPUs: 2 x 1 x 1 along the longest axis,
// 4: 2 x 2 x 1, 8: 2 x 2 x 2), one rank per domain.  Cells are numbered domain by domain (mdk_nlist.cu), so after
// the cell sort each domain's atoms are ONE contiguous range of the tile order; a rank owns the i-blocks of its
// range: it builds and evaluates their work units, owns the bonded / excluded-pair terms whose first atom it owns,
// spreads and gathers its own atoms on the PME mesh and integrates only them.
//
// Memory is replicated (every rank holds arrays for all N atoms: 1.2 GB at a million atoms, of 180 GB), but between
// list rebuilds a rank's copy is CURRENT only for its own atoms and for its halo — the atoms of other domains that
// its work units / terms reference.  Per step:
//   0. PME spread       every rank spreads its own charges (own positions are current after the update) and packs the
//                        sub-mesh it touched for the mesh rank
//   1. halo positions   owner -> user   pack (index list) -> grouped ncclSend / ncclRecv of float4 -> unpack into xs;
//                        a header word per message carries the sender's skin/2 flag, so after this exchange every
//                        rank knows whether ANY atom in the job has moved too far (no separate collective); the
//                        sub-meshes of step 0 travel in the same group (one rendezvous of the ranks per step less)
//   2. forces           pair kernel over own units, PME, bonded terms — all accumulate in int64 fixed point
//   3. halo forces      user -> owner   pack -> grouped ncclSend / ncclRecv of int64 x 3 -> atomic add at the owner
//   4. update           G-JF Langevin of the own atoms only
// PME: every rank spreads its own charges into its copy of the mesh; the part it touched (its domain + spline
// support + drift margin: a box known from the geometry alone) goes to the mesh rank as float, which adds the
// boxes, runs FFT / influence function / inverse FFT on a side stream beside its pair kernel, and returns to every
// rank the potential on that rank's box.
// Rebuild (when the flag of step 1 is up; a few times per 100 steps): all-gather of the float64 state (x, v,
// previous force) from the owners, then every rank redoes the global cell sort (identical on all ranks — this is
// also how atoms migrate between domains) and rebuilds its own lists, halo index lists are derived from what the
// new lists reference and exchanged.
//
// Pair ownership across a domain boundary is balanced by the parity rule of mdk_nlist.cu:k_build_lists.
//
// Backends.  NCCL (one process per GPU; mdk_comm.cu) — or, for tests on a single-GPU box, several contexts of ONE
// process on ONE device ("local" group: the same code path with the transfers done by device-to-device copies
// and the ranks driven in lockstep by one host thread).
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <map>

#include "mdk_common.cuh"

namespace mdk {

struct SubBox { int lo[3], n[3]; size_t pts; };

struct DDState {
    int pdim[3] = {1, 1, 1};
    int pme_rank = 0;
    bool local = false;
    int group_id = -1;
    std::vector<int> blk;                       // [P + 1] first tile slot of every rank (host copy): rank r owns slots [blk[r], blk[r+1])
    DevBuf<int> need, sendl;                    // halo: slots of other ranks this rank needs (ascending = grouped by
                                                // owner) / own slots the peers need (grouped by peer)
    std::vector<int> need_cnt, need_off, send_cnt, send_off;
    int n_need = 0, n_send = 0;
    DevBuf<int> cnt_dev, cnt_all, n_sel, sel_cnt;
    DevBuf<unsigned char> sel_tmp;
    DevBuf<float4> xs_send, xs_recv;            // + one header element per peer at the end
    DevBuf<long long> f_send, f_recv;
    DevBuf<double> st_buf;                      // [9 n] state gather staging
    bool scattered = false;                     // only the own atoms of x_cur / vel / f_prev are current
    bool fresh = false;                         // xs of every atom is current (just rebuilt)
    std::vector<SubBox> box;                    // PME sub-mesh of every rank
    std::vector<size_t> box_off;
    DevBuf<float> m_send, m_recv;
    int *pin = nullptr;                         // pinned: [0..7] flags read-back, [8..] counts / bounds
    // what this member published for the exchange in flight
    const char *sbuf = nullptr;
    char *rbuf = nullptr;
    std::vector<Xfer> xf;
    int64_t stat_exchanges = 0, stat_rebuilds = 0;
    // trace (mdk_dd_trace): with trace on, every phase ends with a stream synchronisation and its wall time is added up
    bool trace = false;
    double t_phase[16] = {0};
    std::chrono::steady_clock::time_point t_mark;
};
enum { TP_HALO_X = 0, TP_REBUILD_GATHER, TP_REBUILD_NLIST, TP_REBUILD_LISTS, TP_AUX_SPREAD, TP_MESH_IN, TP_PAIR, TP_MESH_OUT, TP_GATHER,
       TP_HALO_F, TP_UPDATE, TP_CALL_END, TP_STEPS };


constexpr int PIN_ALL = 8 + 2 * (DD_MAXR + 1);        // pinned words: [0..7] flags, [8..] bounds / own counts, [PIN_ALL..] count matrix
constexpr int PIN_WORDS = PIN_ALL + DD_MAXR * DD_MAXR + 8;
using Group = std::vector<mdk_ctx *>;
static std::map<int, Group> g_groups;

static inline bool own_slot_host(const mdk_ctx *c, int s) { return s >= c->own_lo && s < c->own_hi; }

// ---------------------------------------------------------------------------------------------------------
// kernels
__global__ void k_dd_mark_terms(int nt, int width, const int *__restrict__ idx, const int *__restrict__ inv_order,
                                int own_lo, int own_hi, int *__restrict__ mark) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int s0 = inv_order[idx[(size_t)t * width]];
    if (s0 < own_lo || s0 >= own_hi) return;
    for (int a = 1; a < width; ++a) {
        const int s = inv_order[idx[(size_t)t * width + a]];
        if (s < own_lo || s >= own_hi) mark[s] = 1;
    }
}

__global__ void k_dd_mark_excl(int n_pairs, const int2 *__restrict__ pairs, const int *__restrict__ inv_order, int own_lo, int own_hi,
                               int *__restrict__ mark) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const int k = inv_order[pairs[t].x], p = inv_order[pairs[t].y];
    if (k >= own_lo && k < own_hi && (p < own_lo || p >= own_hi)) mark[p] = 1;     // the pair belongs to the owner of its first atom
}

// predicate of the own-term selection: the term's (pair's) first atom sits in an own tile slot
struct OwnTerm {
    const int *idx; const int *inv_order; int w, lo, hi;
    __device__ bool operator()(int t) const { const int s = inv_order[idx[(size_t)t * w]]; return s >= lo && s < hi; }
};

// cnt[r] = number of needed slots owned by rank r (need is ascending; rank r owns slots [blk[r], blk[r+1]))
__global__ void k_dd_need_counts(const int *__restrict__ need, const int *__restrict__ n_sel, const int *__restrict__ blk,
                                 int P, int *__restrict__ cnt) {
    const int r = threadIdx.x;
    if (r >= P) return;
    const int n = *n_sel;
    auto lb = [&](int v) { int lo = 0, hi = n; while (lo < hi) { int m = (lo + hi) >> 1; if (need[m] < v) lo = m + 1; else hi = m; } return lo; };
    cnt[r] = lb(blk[r + 1]) - lb(blk[r]);
}

__global__ void k_dd_pack_x(int n, const int *__restrict__ list, const float4 *__restrict__ xs, float4 *__restrict__ out,
                            int n_hdr, const int *__restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = xs[list[i]];
    else if (i < n + n_hdr) out[i] = make_float4(flags[1] ? 1.f : 0.f, 0.f, 0.f, 0.f);   // one header per peer
}

__global__ void k_dd_unpack_x(int n, const int *__restrict__ list, const float4 *__restrict__ in, float4 *__restrict__ xs,
                              int n_hdr, int *__restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xs[list[i]] = in[i];
    else if (i < n + n_hdr && in[i].x != 0.f) flags[1] = 1;   // somebody's atom moved skin/2: everybody rebuilds
}

__global__ void k_dd_pack_f(int n, const int *__restrict__ list, long long *__restrict__ f_acc, long long *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t s = 3 * (size_t)list[i];
#pragma unroll
    for (int d = 0; d < 3; ++d) { out[3 * (size_t)i + d] = f_acc[s + d]; f_acc[s + d] = 0; }
}

__global__ void k_dd_unpack_f(int n, const int *__restrict__ list, const long long *__restrict__ in, long long *__restrict__ f_acc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t s = 3 * (size_t)list[i];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const long long v = in[3 * (size_t)i + d];
        if (v) atomic_add_fix(&f_acc[s + d], v);      // several peers may hold partial forces of the same atom
    }
}

// state of tile slots [first, end) <-> staging rows (9 doubles: x, v, previous force), through the tile order
__global__ void k_dd_pack_state(int first, int end, const int *__restrict__ order, const double *__restrict__ x,
                                const double *__restrict__ v, const double *__restrict__ f, double *__restrict__ buf) {
    int k = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= end) return;
    const size_t a = 3 * (size_t)order[k], o = 9 * (size_t)k;
#pragma unroll
    for (int d = 0; d < 3; ++d) { buf[o + d] = x[a + d]; buf[o + 3 + d] = v[a + d]; buf[o + 6 + d] = f ? f[a + d] : 0.0; }
}
__global__ void k_dd_unpack_state(int first, int end, int skip_lo, int skip_hi, const int *__restrict__ order,
                                  const double *__restrict__ buf, double *__restrict__ x, double *__restrict__ v,
                                  double *__restrict__ f) {
    int k = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= end || (k >= skip_lo && k < skip_hi)) return;
    const size_t a = 3 * (size_t)order[k], o = 9 * (size_t)k;
#pragma unroll
    for (int d = 0; d < 3; ++d) { x[a + d] = buf[o + d]; v[a + d] = buf[o + 3 + d]; if (f) f[a + d] = buf[o + 6 + d]; }
}

struct BoxArg { int lo[3], n[3], mesh[3]; };
__device__ __forceinline__ size_t box_index(const BoxArg &b, size_t i) {
    const int iz = (int)(i % b.n[2]);
    const int iy = (int)((i / b.n[2]) % b.n[1]);
    const int ix = (int)(i / ((size_t)b.n[2] * b.n[1]));
    int x = b.lo[0] + ix; if (x >= b.mesh[0]) x -= b.mesh[0];
    int y = b.lo[1] + iy; if (y >= b.mesh[1]) y -= b.mesh[1];
    int z = b.lo[2] + iz; if (z >= b.mesh[2]) z -= b.mesh[2];
    return ((size_t)x * b.mesh[1] + y) * b.mesh[2] + z;
}
__global__ void k_dd_mesh_pack(BoxArg b, size_t pts, long long *__restrict__ fix, float *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pts) return;
    const size_t g = box_index(b, i);
    const long long v = fix[g];
    out[i] = (float)((double)v * (1.0 / FIX_SCALE));
    if (v) fix[g] = 0;
}
__global__ void k_dd_mesh_add(BoxArg b, size_t pts, const float *__restrict__ in, float *__restrict__ grid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pts) grid[box_index(b, i)] += in[i];
}
__global__ void k_dd_mesh_extract(BoxArg b, size_t pts, const float *__restrict__ grid, float *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pts) out[i] = grid[box_index(b, i)];
}
__global__ void k_dd_mesh_put(BoxArg b, size_t pts, const float *__restrict__ in, float *__restrict__ grid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pts) grid[box_index(b, i)] = in[i];
}
__global__ void k_dd_grid_convert(size_t total, long long *__restrict__ fix, float *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long v = fix[i];
    out[i] = (float)((double)v * (1.0 / FIX_SCALE));
    if (v) fix[i] = 0;
}

// ---------------------------------------------------------------------------------------------------------
// phase trace
static void trace_mark(Group &g) {
    for (mdk_ctx *c : g) if (c->dd->trace) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_pme); cudaStreamSynchronize(c->s_aux); c->dd->t_mark = std::chrono::steady_clock::now(); }
}
static void trace_add(Group &g, int phase) {
    for (mdk_ctx *c : g) if (c->dd->trace) {
        cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_pme); cudaStreamSynchronize(c->s_aux);
        const auto now = std::chrono::steady_clock::now();
        c->dd->t_phase[phase] += std::chrono::duration<double, std::milli>(now - c->dd->t_mark).count();
        c->dd->t_mark = now;
    }
}

// ---------------------------------------------------------------------------------------------------------
// group plumbing
static void gsync(Group &g) {
    if (g[0]->dd->local) cudaDeviceSynchronize();     // all members share one device
}
static void each_set_device(mdk_ctx *c) { cudaSetDevice(c->device); }

// every member has filled dd->sbuf / rbuf / xf: run the transfers
static int group_exchange(Group &g) {
    if (!g[0]->dd->local) {
        for (mdk_ctx *c : g) { MDK_TRY(comm_exchange(c, c->dd->sbuf, c->dd->rbuf, c->dd->xf.data(), (int)c->dd->xf.size())); ++c->dd->stat_exchanges; }
        return MDK_OK;
    }
    cudaDeviceSynchronize();
    for (mdk_ctx *c : g) {
        std::vector<int> taken(g.size(), 0);          // transfers from the same peer match in list order
        for (const Xfer &x : c->dd->xf) {
            if (!x.rbytes) continue;
            mdk_ctx *p = g[x.peer];
            int seen = 0; const Xfer *px = nullptr;
            for (const Xfer &y : p->dd->xf) {
                if (y.peer != c->rank || !y.sbytes) continue;
                if (seen++ == taken[x.peer]) { px = &y; break; }
            }
            ++taken[x.peer];
            if (!px || px->sbytes != x.rbytes)
                return fail(c, MDK_ERR_NCCL, "local exchange: rank %d expects %zu bytes from rank %d, which sends %zu", c->rank, x.rbytes,
                            x.peer, px ? px->sbytes : (size_t)0);
            MDK_CUDA(c, cudaMemcpyAsync(x.rptr ? x.rptr : c->dd->rbuf + x.roff, px->sptr ? px->sptr : p->dd->sbuf + px->soff, x.rbytes,
                                        cudaMemcpyDeviceToDevice, c->stream));
        }
        ++c->dd->stat_exchanges;
    }
    cudaDeviceSynchronize();
    return MDK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// state gather: afterwards every rank holds the current x / v / previous force of every atom
static int dd_gather_state(Group &g) {
    bool need = false;
    for (mdk_ctx *c : g) need = need || c->dd->scattered;
    if (!need) return MDK_OK;
    const int P = g[0]->nranks;
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        MDK_CUDA(c, d->st_buf.reserve((size_t)9 * c->n_pad));
        const int lo = own_first(c), hi = own_end(c);
        if (hi > lo)
            k_dd_pack_state<<<(hi - lo + 255) / 256, 256, 0, c->stream>>>(lo, hi, c->order.p, c->x_cur.p, c->vel.p,
                                                                       c->f_prev.p, d->st_buf.p);
        ++c->n_launches;
        d->sbuf = reinterpret_cast<const char *>(d->st_buf.p);
        d->rbuf = reinterpret_cast<char *>(d->st_buf.p);
        d->xf.clear();
        const size_t row = 9 * sizeof(double);
        for (int r = 0; r < P; ++r) {
            if (r == c->rank) continue;
            const int rlo = d->blk[r], rhi = d->blk[r + 1];
            d->xf.push_back(Xfer{r, (size_t)lo * row, (size_t)(hi - lo) * row, (size_t)rlo * row, (size_t)(rhi - rlo) * row});
        }
    }
    MDK_TRY(group_exchange(g));
    for (mdk_ctx *c : g) {
        each_set_device(c);
        k_dd_unpack_state<<<(c->n + 255) / 256, 256, 0, c->stream>>>(0, c->n, own_first(c), own_end(c), c->order.p, c->dd->st_buf.p,
                                                                     c->x_cur.p, c->vel.p, c->f_prev.p);
        ++c->n_launches;
        c->dd->scattered = false;
    }
    gsync(g);
    return MDK_OK;
}

// PME sub-mesh of every rank from the geometry alone (identical on all ranks): the domain's cell range (a rank spreads
// exactly the atoms whose cell lies in its domain) + the drift allowed between rebuilds + spline support
static void dd_mesh_boxes(mdk_ctx *c) {
    DDState *d = c->dd;
    const DDGeom &gm = c->dd_geom;
    const int P = c->nranks;
    d->box.assign(P, SubBox{});
    d->box_off.assign(P + 1, 0);
    for (int r = 0; r < P; ++r) {
        const int dom[3] = {r % gm.pdim[0], (r / gm.pdim[0]) % gm.pdim[1], r / (gm.pdim[0] * gm.pdim[1])};
        const int cell_lo[3] = {gm.cut0[dom[0]], gm.cut1[dom[0]][dom[1]], gm.cut2[dom[0]][dom[1]][dom[2]]};
        const int cell_hi[3] = {gm.cut0[dom[0] + 1], gm.cut1[dom[0]][dom[1] + 1], gm.cut2[dom[0]][dom[1]][dom[2] + 1]};
        SubBox &b = d->box[r];
        b.pts = 1;
        for (int a = 0; a < 3; ++a) {
            const double w = c->cellw[a], L = c->box.Ld[a];
            // (0.75 skin: the step that raises the rebuild flag spreads BEFORE the rebuild, with atoms up to skin/2 + one step's move from
            // where the lists were built)
            const double x0 = cell_lo[a] * w - 0.75 * c->skin - 0.05 * w, x1 = cell_hi[a] * w + 0.75 * c->skin + 0.05 * w;
            const int m = c->pme_n[a];
            int lo = (int)floor(x0 / L * m) - c->pme_order, hi = (int)floor(x1 / L * m) + 2;   // mesh index of x = -L/2 is 0
            int n = hi - lo + 1;
            if (n >= m) { lo = 0; n = m; }
            b.lo[a] = ((lo % m) + m) % m;
            b.n[a] = n;
            b.pts *= (size_t)n;
        }
        d->box_off[r + 1] = d->box_off[r] + b.pts;
    }
}
static BoxArg box_arg(const mdk_ctx *c, const SubBox &b) {
    BoxArg a;
    for (int k = 0; k < 3; ++k) { a.lo[k] = b.lo[k]; a.n[k] = b.n[k]; a.mesh[k] = c->pme_n[k]; }
    return a;
}

// ---------------------------------------------------------------------------------------------------------
// rebuild: global sort on every rank, own lists, halo lists
static int dd_rebuild(Group &g) {
    trace_mark(g);
    MDK_TRY(dd_gather_state(g));
    trace_add(g, TP_REBUILD_GATHER);
    const int P = g[0]->nranks;
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        MDK_TRY(nlist_rebuild(c));                // keys (domain-major) -> sort -> gathers -> bounds -> own lists (+ marks)
        if (d->trace) { Group one{c}; trace_add(one, TP_REBUILD_NLIST); }
        // ownership of this rebuild
        MDK_CUDA(c, cudaMemcpyAsync(d->pin + 8, c->dd_blk.p, (DD_MAXR + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
        d->blk.assign(d->pin + 8, d->pin + 8 + P + 1);
        c->own_lo = d->blk[c->rank];
        c->own_hi = d->blk[c->rank + 1];
        c->pme_lo = c->own_lo; c->pme_hi = c->own_hi;
        // halo: what the lists reference (marked by the builder) + the partners of the own bonded / excluded-pair terms
        static const int width[4] = {2, 3, 4, 4};
        for (int kind = 0; kind < 4; ++kind)
            if (c->bonded[kind].n > 0)
                k_dd_mark_terms<<<(c->bonded[kind].n + 255) / 256, 256, 0, c->stream>>>(c->bonded[kind].n, width[kind], c->bonded[kind].idx.p,
                                                                                     c->inv_order.p, c->own_lo, c->own_hi, c->dd_mark.p);
        if (c->n_excl_pairs > 0)
            k_dd_mark_excl<<<(c->n_excl_pairs + 255) / 256, 256, 0, c->stream>>>(c->n_excl_pairs, c->excl_pairs.p, c->inv_order.p, c->own_lo,
                                                                              c->own_hi, c->dd_mark.p);
        MDK_CUDA(c, d->need.reserve(c->n_pad));
        MDK_CUDA(c, d->n_sel.reserve(1)); MDK_CUDA(c, d->cnt_dev.reserve(DD_MAXR)); MDK_CUDA(c, d->cnt_all.reserve(DD_MAXR * DD_MAXR));
        size_t tmp = 0;
        cub::CountingInputIterator<int> ids(0);
        cub::DeviceSelect::Flagged(nullptr, tmp, ids, c->dd_mark.p, d->need.p, d->n_sel.p, c->n_pad, c->stream);
        MDK_CUDA(c, d->sel_tmp.reserve(tmp));
        MDK_CUDA(c, cub::DeviceSelect::Flagged(d->sel_tmp.p, tmp, ids, c->dd_mark.p, d->need.p, d->n_sel.p, c->n_pad, c->stream));
        k_dd_need_counts<<<1, DD_MAXR, 0, c->stream>>>(d->need.p, d->n_sel.p, c->dd_blk.p, P, d->cnt_dev.p);
        c->n_launches += 6;
        MDK_CUDA(c, cudaMemcpyAsync(d->pin + 8, d->cnt_dev.p, P * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        // the bonded terms / excluded pairs this rank owns, as index lists: the O(N) kernel then walks its share only
        {
            MDK_CUDA(c, d->sel_cnt.reserve(8));
            const int nk[5] = {c->bonded[0].n, c->bonded[1].n, c->bonded[2].n, c->bonded[3].n, c->n_excl_pairs};
            const int *ix[5] = {c->bonded[0].idx.p, c->bonded[1].idx.p, c->bonded[2].idx.p, c->bonded[3].idx.p,
                                reinterpret_cast<const int *>(c->excl_pairs.p)};
            for (int k = 0; k < 5; ++k) {
                c->aux_sel_n[k] = -1;
                if (nk[k] <= 0) continue;
                MDK_CUDA(c, c->aux_sel[k].reserve(nk[k]));
                OwnTerm pred{ix[k], c->inv_order.p, k == 4 ? 2 : width[k], c->own_lo, c->own_hi};
                size_t tmp2 = 0;
                cub::DeviceSelect::If(nullptr, tmp2, ids, c->aux_sel[k].p, d->sel_cnt.p + k, nk[k], pred, c->stream);
                MDK_CUDA(c, d->sel_tmp.reserve(tmp2));
                MDK_CUDA(c, cub::DeviceSelect::If(d->sel_tmp.p, tmp2, ids, c->aux_sel[k].p, d->sel_cnt.p + k, nk[k], pred, c->stream));
                c->aux_sel_n[k] = -2;      // count follows with the read-back below
            }
            MDK_CUDA(c, cudaMemcpyAsync(d->pin + PIN_WORDS - 8, d->sel_cnt.p, 5 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    // counts: need_cnt[o] of every rank -> send_cnt[r] = need_cnt of rank r for me
    if (!g[0]->dd->local) {
        mdk_ctx *c = g[0];
        MDK_TRY(comm_allgather_i32(c, c->dd->cnt_dev.p, c->dd->cnt_all.p, P));
        MDK_CUDA(c, cudaMemcpyAsync(c->dd->pin + PIN_ALL, c->dd->cnt_all.p, (size_t)P * P * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    } else {
        cudaDeviceSynchronize();
        for (mdk_ctx *c : g)
            for (mdk_ctx *p : g)
                for (int o = 0; o < P; ++o) c->dd->pin[PIN_ALL + p->rank * P + o] = p->dd->pin[8 + o];
    }
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        const int *all = d->pin + PIN_ALL;
        d->need_cnt.assign(P, 0); d->need_off.assign(P + 1, 0); d->send_cnt.assign(P, 0); d->send_off.assign(P + 1, 0);
        for (int r = 0; r < P; ++r) {
            d->need_cnt[r] = all[c->rank * P + r];
            d->send_cnt[r] = all[r * P + c->rank];
            if (r == c->rank && (d->need_cnt[r] || d->send_cnt[r])) return fail(c, MDK_ERR_CUDA, "domain decomposition: a rank needs its own atoms as halo");
            d->need_off[r + 1] = d->need_off[r] + d->need_cnt[r];
            d->send_off[r + 1] = d->send_off[r] + d->send_cnt[r];
        }
        d->n_need = d->need_off[P]; d->n_send = d->send_off[P];
        for (int k = 0; k < 5; ++k)
            if (c->aux_sel_n[k] == -2) c->aux_sel_n[k] = d->pin[PIN_WORDS - 8 + k];
        MDK_CUDA(c, d->sendl.reserve(d->n_send + 1));
        MDK_CUDA(c, d->xs_send.reserve(d->n_send + P)); MDK_CUDA(c, d->xs_recv.reserve(d->n_need + P));
        MDK_CUDA(c, d->f_send.reserve(3 * (size_t)d->n_need + 3)); MDK_CUDA(c, d->f_recv.reserve(3 * (size_t)d->n_send + 3));
        // the index lists: my need-list segment of owner o goes to o, which files it as "what rank me needs"
        d->sbuf = reinterpret_cast<const char *>(d->need.p);
        d->rbuf = reinterpret_cast<char *>(d->sendl.p);
        d->xf.clear();
        for (int r = 0; r < P; ++r)
            if (r != c->rank && (d->need_cnt[r] || d->send_cnt[r]))
                d->xf.push_back(Xfer{r, d->need_off[r] * sizeof(int), d->need_cnt[r] * sizeof(int), d->send_off[r] * sizeof(int), d->send_cnt[r] * sizeof(int)});
    }
    MDK_TRY(group_exchange(g));
    for (mdk_ctx *c : g) {
        each_set_device(c);
        if (c->have_pme) dd_mesh_boxes(c);
        c->dd->fresh = true;
        c->xs_current = true;
        ++c->dd->stat_rebuilds;
    }
    gsync(g);
    trace_add(g, TP_REBUILD_LISTS);
    return MDK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// one force evaluation of the whole job
// with_mesh: the sub-meshes (already spread and packed by dd_spread_pack) ride in the same grouped exchange
static int dd_halo_positions(Group &g, bool *rebuild, bool with_mesh) {
    const int P = g[0]->nranks;
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        const int tot = d->n_send + P;
        k_dd_pack_x<<<(tot + 255) / 256, 256, 0, c->stream>>>(d->n_send, d->sendl.p, c->xs.p, d->xs_send.p, P, c->flags.p);
        ++c->n_launches;
        d->sbuf = reinterpret_cast<const char *>(d->xs_send.p);
        d->rbuf = reinterpret_cast<char *>(d->xs_recv.p);
        d->xf.clear();
        for (int r = 0; r < P; ++r) {
            if (r == c->rank) continue;
            if (d->send_cnt[r] || d->need_cnt[r])
                d->xf.push_back(Xfer{r, d->send_off[r] * sizeof(float4), d->send_cnt[r] * sizeof(float4), d->need_off[r] * sizeof(float4),
                                     d->need_cnt[r] * sizeof(float4)});
            d->xf.push_back(Xfer{r, (size_t)(d->n_send + r) * sizeof(float4), sizeof(float4), (size_t)(d->n_need + r) * sizeof(float4), sizeof(float4)});
        }
        if (with_mesh) {
            if (c->rank != d->pme_rank) {
                d->xf.push_back(Xfer{d->pme_rank, 0, d->box[c->rank].pts * sizeof(float), 0, 0, d->m_send.p, nullptr});
            } else {
                for (int r = 0; r < P; ++r)
                    if (r != c->rank) d->xf.push_back(Xfer{r, 0, 0, 0, d->box[r].pts * sizeof(float), nullptr, d->m_recv.p + d->box_off[r]});
            }
        }
    }
    MDK_TRY(group_exchange(g));
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        MDK_CUDA(c, cudaMemsetAsync(d->xs_recv.p + d->n_need + c->rank, 0, sizeof(float4), c->stream));   // own header slot: never received
        const int tot = d->n_need + P;
        k_dd_unpack_x<<<(tot + 255) / 256, 256, 0, c->stream>>>(d->n_need, d->need.p, d->xs_recv.p, c->xs.p, P, c->flags.p);
        ++c->n_launches;
        MDK_CUDA(c, cudaMemcpyAsync(d->pin, c->flags.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    *rebuild = false;
    for (mdk_ctx *c : g) {
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->dd->pin[1]) *rebuild = true;
    }
    return MDK_OK;
}

// clear: zero the force accumulator first (otherwise the caller guarantees clean own / halo rows; a rebuild always clears)
static int dd_forces(Group &g, unsigned terms, bool clear) {
    const int P = g[0]->nranks;
    if (terms & MDK_TERM_COUL_BARE)
        return fail(g[0], MDK_ERR_BAD_ARG, "the all-pairs reference Coulomb sum (MDK_TERM_COUL_BARE) is not domain decomposed");
    bool rebuild = false;
    for (mdk_ctx *c : g) rebuild = rebuild || !c->nlist_valid || !c->xs_current;
    const bool pme = terms & MDK_TERM_PME_RECIP;
    const unsigned bonded_bits = terms & (MDK_TERM_BOND | MDK_TERM_ANGLE | MDK_TERM_DIHEDRAL | MDK_TERM_IMPROPER);
    bool fresh = true;
    for (mdk_ctx *c : g) fresh = fresh && c->dd->fresh;
    // own charges onto the own mesh copy, the touched box packed for the mesh rank (needs the own atoms' positions only)
    auto spread_pack = [&](bool add_xfers) -> int {
        for (mdk_ctx *c : g) {
            each_set_device(c);
            MDK_TRY(pme_prepare(c));
            DDState *d = c->dd;
            if (d->box.empty()) dd_mesh_boxes(c);
            MDK_TRY(pme_spread(c));
            const bool mesh_rank = c->rank == d->pme_rank;
            const size_t total_pts = d->box_off[P];
            MDK_CUDA(c, d->m_send.reserve(mesh_rank ? total_pts : d->box[c->rank].pts));
            MDK_CUDA(c, d->m_recv.reserve(mesh_rank ? total_pts : d->box[c->rank].pts));
            if (add_xfers) {
                d->sbuf = reinterpret_cast<const char *>(d->m_send.p);
                d->rbuf = reinterpret_cast<char *>(d->m_recv.p);
                d->xf.clear();
            }
            if (!mesh_rank) {
                const SubBox &b = d->box[c->rank];
                k_dd_mesh_pack<<<(unsigned)((b.pts + 255) / 256), 256, 0, c->stream>>>(box_arg(c, b), b.pts, c->grid_fix.p, d->m_send.p);
                ++c->n_launches;
                if (add_xfers) d->xf.push_back(Xfer{d->pme_rank, 0, b.pts * sizeof(float), 0, 0});
            } else if (add_xfers) {
                for (int r = 0; r < P; ++r)
                    if (r != c->rank) d->xf.push_back(Xfer{r, 0, 0, d->box_off[r] * sizeof(float), d->box[r].pts * sizeof(float)});
            }
        }
        return MDK_OK;
    };
    // the normal step: spread first, so that the sub-meshes travel with the halo positions in ONE grouped exchange (one
    // rendezvous of the ranks less per step).  A rebuild that this very exchange announces comes after it: the spread then used
    // the previous ownership, which is still a partition of all atoms, and the boxes carry the margin for it (dd_mesh_boxes).
    const bool early = pme && P > 1 && !rebuild && !fresh && !g[0]->dd_late_spread;
    // NCCL backend only: the potential boxes return on the side stream, their receives posted before the pair kernel
    const bool early_recv = pme && P > 1 && !g[0]->dd->local && g[0]->dd_early_recv;
    trace_mark(g);
    if (early) MDK_TRY(spread_pack(false));
    trace_add(g, TP_AUX_SPREAD);
    if (!rebuild && !fresh) MDK_TRY(dd_halo_positions(g, &rebuild, early));
    trace_add(g, TP_HALO_X);
    if (rebuild) MDK_TRY(dd_rebuild(g));
    trace_mark(g);
    if (rebuild || clear)
        for (mdk_ctx *c : g) { each_set_device(c); MDK_CUDA(c, cudaMemsetAsync(c->f_acc.p, 0, (size_t)c->n_pad * 3 * sizeof(long long), c->stream)); }
    for (mdk_ctx *c : g) {
        each_set_device(c);
        c->dd->fresh = false;
        MDK_CUDA(c, cudaMemsetAsync(c->e_acc.p, 0, MDK_NUM_ENERGIES * sizeof(long long), c->stream));
        cudaStream_t main_stream = c->stream;
        // O(N) terms of the own atoms on the side stream
        if (bonded_bits || pme) {
            MDK_CUDA(c, cudaEventRecord(c->ev_fork, main_stream));
            MDK_CUDA(c, cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0));
            c->stream = c->s_aux;
            int rc = bonded_compute(c, terms);
            cudaEventRecord(c->ev_aux, c->s_aux);
            c->stream = main_stream;
            MDK_TRY(rc);
        }
    }
    if (pme && !early) MDK_TRY(spread_pack(true));
    trace_add(g, TP_AUX_SPREAD);
    if (pme && P > 1 && !early) MDK_TRY(group_exchange(g));
    trace_add(g, TP_MESH_IN);
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        cudaStream_t main_stream = c->stream;
        if (pme && c->rank == d->pme_rank) {
            // mesh chain beside the pair kernel: own mesh -> float, + the other domains' boxes, FFTs, boxes out
            MDK_CUDA(c, cudaEventRecord(c->ev_fork, main_stream));
            MDK_CUDA(c, cudaStreamWaitEvent(c->s_pme, c->ev_fork, 0));
            c->stream = c->s_pme;
            const size_t total = (size_t)c->pme_n[0] * c->pme_n[1] * c->pme_n[2];
            k_dd_grid_convert<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(total, c->grid_fix.p, c->grid_r.p);
            for (int r = 0; r < P; ++r)
                if (r != c->rank) {
                    const SubBox &b = d->box[r];
                    k_dd_mesh_add<<<(unsigned)((b.pts + 255) / 256), 256, 0, c->stream>>>(box_arg(c, b), b.pts, d->m_recv.p + d->box_off[r], c->grid_r.p);
                }
            int rc = pme_mesh(c, false);
            for (int r = 0; r < P && rc == MDK_OK; ++r)
                if (r != c->rank) {
                    const SubBox &b = d->box[r];
                    k_dd_mesh_extract<<<(unsigned)((b.pts + 255) / 256), 256, 0, c->stream>>>(box_arg(c, b), b.pts, c->grid_r.p, d->m_send.p + d->box_off[r]);
                }
            c->n_launches += 1 + 2 * (P - 1);
            if (rc == MDK_OK && early_recv) {     // the potential boxes leave from the side stream as soon as the mesh chain is through
                d->xf.clear();
                for (int r = 0; r < P; ++r)
                    if (r != c->rank) d->xf.push_back(Xfer{r, d->box_off[r] * sizeof(float), d->box[r].pts * sizeof(float), 0, 0});
                rc = comm_exchange(c, d->m_send.p, d->m_recv.p, d->xf.data(), (int)d->xf.size());
                ++d->stat_exchanges;
            }
            cudaEventRecord(c->ev_pme, c->s_pme);
            c->stream = main_stream;
            MDK_TRY(rc);
        } else if (early_recv) {
            // the receive of this rank's potential box is posted on the side stream BEFORE the pair kernel is launched: the
            // NCCL kernel holds its few CTAs while the persistent pair kernel fills the rest of the machine, and the box is
            // there when the pair kernel ends (posted after it, the transfer would start only then)
            const SubBox &b = d->box[c->rank];
            MDK_CUDA(c, cudaEventRecord(c->ev_fork, main_stream));
            MDK_CUDA(c, cudaStreamWaitEvent(c->s_pme, c->ev_fork, 0));
            c->stream = c->s_pme;
            d->xf.clear();
            d->xf.push_back(Xfer{d->pme_rank, 0, 0, 0, b.pts * sizeof(float)});
            int rc = comm_exchange(c, d->m_send.p, d->m_recv.p, d->xf.data(), (int)d->xf.size());
            ++d->stat_exchanges;
            if (rc == MDK_OK) {
                k_dd_mesh_put<<<(unsigned)((b.pts + 255) / 256), 256, 0, c->stream>>>(box_arg(c, b), b.pts, d->m_recv.p, c->grid_r.p);
                ++c->n_launches;
            }
            cudaEventRecord(c->ev_pme, c->s_pme);
            c->stream = main_stream;
            MDK_TRY(rc);
        }
        MDK_TRY(pair_compute(c, terms & MDK_TERM_LJ, terms & MDK_TERM_COUL_DIRECT));
        if (pme && early_recv) {
            MDK_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_pme, 0));
        } else if (pme) {
            d->sbuf = reinterpret_cast<const char *>(d->m_send.p);
            d->rbuf = reinterpret_cast<char *>(d->m_recv.p);
            d->xf.clear();
            if (c->rank == d->pme_rank) {
                MDK_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_pme, 0));
                for (int r = 0; r < P; ++r)
                    if (r != c->rank) d->xf.push_back(Xfer{r, d->box_off[r] * sizeof(float), d->box[r].pts * sizeof(float), 0, 0});
            } else {
                d->xf.push_back(Xfer{d->pme_rank, 0, 0, 0, d->box[c->rank].pts * sizeof(float)});
            }
        }
    }
    trace_add(g, TP_PAIR);
    if (pme && P > 1 && !early_recv) MDK_TRY(group_exchange(g));
    trace_add(g, TP_MESH_OUT);
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        if (pme) {
            if (c->rank != d->pme_rank && !early_recv) {
                const SubBox &b = d->box[c->rank];
                k_dd_mesh_put<<<(unsigned)((b.pts + 255) / 256), 256, 0, c->stream>>>(box_arg(c, b), b.pts, d->m_recv.p, c->grid_r.p);
                ++c->n_launches;
            }
            MDK_TRY(pme_gather(c));
        }
        if (bonded_bits || pme) MDK_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_aux, 0));
        // halo forces back to their owners
        if (d->n_need) k_dd_pack_f<<<(d->n_need + 255) / 256, 256, 0, c->stream>>>(d->n_need, d->need.p, c->f_acc.p, d->f_send.p);
        ++c->n_launches;
        d->sbuf = reinterpret_cast<const char *>(d->f_send.p);
        d->rbuf = reinterpret_cast<char *>(d->f_recv.p);
        d->xf.clear();
        const size_t row = 3 * sizeof(long long);
        for (int r = 0; r < P; ++r)
            if (r != c->rank && (d->need_cnt[r] || d->send_cnt[r]))
                d->xf.push_back(Xfer{r, d->need_off[r] * row, d->need_cnt[r] * row, d->send_off[r] * row, d->send_cnt[r] * row});
    }
    trace_add(g, TP_GATHER);
    MDK_TRY(group_exchange(g));
    for (mdk_ctx *c : g) {
        each_set_device(c);
        DDState *d = c->dd;
        if (d->n_send) k_dd_unpack_f<<<(d->n_send + 255) / 256, 256, 0, c->stream>>>(d->n_send, d->sendl.p, d->f_recv.p, c->f_acc.p);
        ++c->n_launches;
        MDK_CUDA(c, cudaGetLastError());
    }
    gsync(g);
    trace_add(g, TP_HALO_F);
    return MDK_OK;
}

// energies of the job: kinetic energy of the own atoms, sum over the ranks, read back
static int dd_energies(Group &g, unsigned terms) {
    for (mdk_ctx *c : g) { each_set_device(c); MDK_TRY(energies_enqueue(c)); }     // NCCL backend: all-reduce inside
    for (mdk_ctx *c : g) MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (g[0]->dd->local) {
        long long sum[MDK_NUM_ENERGIES] = {0};
        for (mdk_ctx *c : g)
            for (int k = 0; k < MDK_NUM_ENERGIES; ++k) sum[k] += c->pin_words[k];
        for (mdk_ctx *c : g)
            for (int k = 0; k < MDK_NUM_ENERGIES; ++k) c->pin_words[k] = sum[k];
    }
    for (mdk_ctx *c : g) { energies_finish(c, terms); MDK_TRY(check_lost_flag(c)); }
    return MDK_OK;
}

// all forces on every rank (Constraint.forces / Ensemble.update read them in matrix order): the own segments of
// the tile-order accumulator are exchanged in place
static int dd_gather_forces(Group &g) {
    const int P = g[0]->nranks;
    const size_t row = 3 * sizeof(long long);
    for (mdk_ctx *c : g) {
        DDState *d = c->dd;
        d->sbuf = reinterpret_cast<const char *>(c->f_acc.p);
        d->rbuf = reinterpret_cast<char *>(c->f_acc.p);
        d->xf.clear();
        const int lo = own_first(c), hi = own_end(c);
        for (int r = 0; r < P; ++r) {
            if (r == c->rank) continue;
            const int rlo = d->blk[r], rhi = d->blk[r + 1];
            d->xf.push_back(Xfer{r, (size_t)lo * row, (size_t)(hi - lo) * row, (size_t)rlo * row, (size_t)(rhi - rlo) * row});
        }
    }
    return group_exchange(g);
}

static int dd_compute_group(Group &g, unsigned terms, bool sync_energies) {
    for (mdk_ctx *c : g) {
        each_set_device(c);
        if (!c->have_box || c->n <= 0 || !c->have_pos) return fail(c, MDK_ERR_NOT_BOUND, "mdk_compute before box/atoms/positions were set");
    }
    MDK_TRY(dd_forces(g, terms, true));
    MDK_TRY(dd_gather_forces(g));
    if (sync_energies) {
        // (kinetic energy rides along; harmless for a pure force evaluation)
        MDK_TRY(dd_energies(g, terms));
    }
    gsync(g);
    return MDK_OK;
}

static int dd_langevin_group(Group &g, double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms, bool defer_energies) {
    if (nsteps <= 0) return MDK_OK;
    const double ca = (1.0 - 0.5 * gamma * dt) / (1.0 + 0.5 * gamma * dt), cb = 1.0 / (1.0 + 0.5 * gamma * dt);
    const double tg = 2.0 * gamma * kT * dt;
    for (mdk_ctx *c : g)
        if (c->n_rigid > 0) return fail(c, MDK_ERR_BAD_ARG, "rigid waters are single domain (a molecule's atoms may belong to two ranks)");
    for (mdk_ctx *c : g) {
        each_set_device(c);
        MDK_CUDA(c, c->f_prev.reserve((size_t)3 * c->n));
        if (terms != c->cached_terms) { c->verlet_cached = false; c->langevin_cached = false; c->cached_terms = terms; }
    }
    bool cached = true;
    for (mdk_ctx *c : g) cached = cached && c->langevin_cached && c->nlist_valid && c->xs_current;
    auto update = [&](int mode) -> int {
        for (mdk_ctx *c : g) {
            each_set_device(c);
            MDK_TRY(langevin_launch(c, own_first(c), own_end(c), mode, dt, ca, cb, tg, seed, c->langevin_step));
            c->dd->scattered = true;
        }
        return MDK_OK;
    };
    auto clear_forces = [&]() -> int {
        for (mdk_ctx *c : g) { each_set_device(c); MDK_CUDA(c, cudaMemsetAsync(c->f_acc.p, 0, (size_t)c->n_pad * 3 * sizeof(long long), c->stream)); }
        return MDK_OK;
    };
    if (!cached) {
        for (mdk_ctx *c : g) c->langevin_cached = false;
        // f(x_0) — also (re)builds the lists and with them the ownership
        MDK_TRY(dd_forces(g, terms, true));
        MDK_TRY(update(2));
    } else {
        MDK_TRY(update(2 | 4));
    }
    for (mdk_ctx *c : g) { c->langevin_cached = true; ++c->langevin_step; }
    MDK_TRY(clear_forces());
    for (int s = 0; s < nsteps; ++s) {
        MDK_TRY(dd_forces(g, terms, false));          // f(x_n+1)
        const bool more = s + 1 < nsteps;
        trace_mark(g);
        MDK_TRY(update(more ? 3 : 1));                // mode 3 leaves the own accumulator rows clean for the next step
        trace_add(g, TP_UPDATE);
        if (more) for (mdk_ctx *c : g) ++c->langevin_step;
        for (mdk_ctx *c : g) c->dd->t_phase[TP_STEPS] += 1.0;
    }
    trace_mark(g);
    // the call hands back a complete state on every rank (positions, velocities, forces of the last evaluation)
    MDK_TRY(dd_gather_state(g));
    MDK_TRY(dd_gather_forces(g));
    for (mdk_ctx *c : g) { each_set_device(c); MDK_TRY(energies_enqueue(c)); }
    trace_add(g, TP_CALL_END);
    if (defer_energies && !g[0]->dd->local) return MDK_OK;     // the caller synchronises once, after queueing its downloads
    for (mdk_ctx *c : g) MDK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (g[0]->dd->local) {
        long long sum[MDK_NUM_ENERGIES] = {0};
        for (mdk_ctx *c : g)
            for (int k = 0; k < MDK_NUM_ENERGIES; ++k) sum[k] += c->pin_words[k];
        for (mdk_ctx *c : g)
            for (int k = 0; k < MDK_NUM_ENERGIES; ++k) c->pin_words[k] = sum[k];
    }
    for (mdk_ctx *c : g) { energies_finish(c, terms); MDK_TRY(check_lost_flag(c)); }
    return MDK_OK;
}

// single-context entry points (NCCL backend: the group is this process's one context)
int dd_compute_single(mdk_ctx *c, unsigned terms, bool sync_energies) {
    if (c->dd->local) return fail(c, MDK_ERR_BAD_ARG, "a context of a local domain group is driven through mdk_dd_compute_group");
    Group g{c};
    return dd_compute_group(g, terms, sync_energies);
}
int dd_langevin_single(mdk_ctx *c, double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms, bool defer_energies) {
    if (c->dd->local) return fail(c, MDK_ERR_BAD_ARG, "a context of a local domain group is driven through mdk_dd_step_langevin_group");
    Group g{c};
    return dd_langevin_group(g, dt, kT, gamma, seed, nsteps, terms, defer_energies);
}

void dd_destroy(mdk_ctx *c) {
    DDState *d = c->dd;
    if (!d) return;
    if (d->local) {
        auto it = g_groups.find(d->group_id);
        if (it != g_groups.end()) {
            for (auto &m : it->second) if (m == c) m = nullptr;
            bool any = false;
            for (auto m : it->second) any = any || m;
            if (!any) g_groups.erase(it);
        }
    }
    d->sel_cnt.release();
    for (int k = 0; k < 5; ++k) { c->aux_sel[k].release(); c->aux_sel_n[k] = -1; }
    d->need.release(); d->sendl.release(); d->cnt_dev.release(); d->cnt_all.release(); d->n_sel.release(); d->sel_tmp.release();
    d->xs_send.release(); d->xs_recv.release(); d->f_send.release(); d->f_recv.release(); d->st_buf.release();
    d->m_send.release(); d->m_recv.release();
    if (d->pin) cudaFreeHost(d->pin);
    delete d;
    c->dd = nullptr;
    c->own_lo = 0; c->own_hi = -1;
    c->pme_lo = 0; c->pme_hi = -1;
}

}  // namespace mdk

using namespace mdk;

extern "C" {

int mdk_dd_init(mdk_ctx *c, int rank, int nranks, int px, int py, int pz, int local_group) {
    if (!c) return MDK_ERR_BAD_ARG;
    cudaSetDevice(c->device);
    if (px < 1 || py < 1 || pz < 1 || px > DD_MAXP || py > DD_MAXP || pz > DD_MAXP || px * py * pz != nranks || rank < 0 || rank >= nranks)
        return fail(c, MDK_ERR_BAD_ARG, "mdk_dd_init(rank %d of %d, grid %d x %d x %d)", rank, nranks, px, py, pz);
    dd_destroy(c);
    if (nranks == 1) { c->nlist_valid = false; ++c->graph_epoch; return MDK_OK; }
    if (local_group < 0 && (!c->nccl_comm || c->rank != rank || c->nranks != nranks))
        return fail(c, MDK_ERR_NCCL, "mdk_dd_init: join the communicator first (mdk_comm_init with the same rank / size)");
    DDState *d = new DDState();
    d->pdim[0] = px; d->pdim[1] = py; d->pdim[2] = pz;
    d->pme_rank = nranks - 1;
    d->local = local_group >= 0;
    d->group_id = local_group;
    if (cudaHostAlloc(reinterpret_cast<void **>(&d->pin), PIN_WORDS * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
        delete d;
        return fail(c, MDK_ERR_OOM, "cudaHostAlloc failed in mdk_dd_init");
    }
    memset(d->pin, 0, PIN_WORDS * sizeof(int));
    d->blk.assign(nranks + 1, 0);
    c->dd = d;
    c->dd_geom.pdim[0] = px; c->dd_geom.pdim[1] = py; c->dd_geom.pdim[2] = pz;
    if (d->local) {
        c->rank = rank; c->nranks = nranks;
        Group &g = g_groups[local_group];
        if ((int)g.size() != nranks) g.assign(nranks, nullptr);
        g[rank] = c;
    }
    c->nlist_valid = false; c->xs_current = false;
    c->verlet_cached = false; c->langevin_cached = false;
    c->pme_dirty = true;          // the decomposed step uses the cuFFT mesh chain (sub-meshes are added in float)
    ++c->graph_epoch;
    return MDK_OK;
}

static int local_group_of(mdk_ctx *const *ctxs, int n, Group &g) {
    if (!ctxs || n < 1 || !ctxs[0] || !ctxs[0]->dd) return MDK_ERR_BAD_ARG;
    g.assign(ctxs, ctxs + n);
    for (int r = 0; r < n; ++r)
        if (!g[r] || !g[r]->dd || g[r]->rank != r || g[r]->nranks != n || g[r]->dd->group_id != g[0]->dd->group_id)
            return fail(g[0], MDK_ERR_BAD_ARG, "mdk_dd_*_group: pass every context of the group, in rank order");
    return MDK_OK;
}

/* One force evaluation of a local group (every context holds the same positions on entry). */
int mdk_dd_compute_group(mdk_ctx *const *ctxs, int n, unsigned terms, double *energies) {
    Group g;
    MDK_TRY(local_group_of(ctxs, n, g));
    for (mdk_ctx *c : g) { cudaSetDevice(c->device); prepare_pme_constants(c); }
    MDK_TRY(dd_compute_group(g, terms, true));
    if (energies) memcpy(energies, g[0]->last_e, sizeof(g[0]->last_e));
    return MDK_OK;
}

int mdk_dd_step_langevin_group(mdk_ctx *const *ctxs, int n, double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms,
                               double *energies) {
    Group g;
    MDK_TRY(local_group_of(ctxs, n, g));
    for (mdk_ctx *c : g) { cudaSetDevice(c->device); prepare_pme_constants(c); }
    MDK_TRY(dd_langevin_group(g, dt, kT, gamma, seed, nsteps, terms, false));
    if (energies) memcpy(energies, g[0]->last_e, sizeof(g[0]->last_e));
    return MDK_OK;
}

/* Phase trace of the decomposed step.  on != 0: every phase of the following calls ends with a stream synchronisation and
 * its wall time is accumulated; out16 (may be NULL) receives the sums in ms — [0] halo positions (+ flag read-back),
 * [1] rebuild: state all-gather, [2] rebuild: sort + own lists, [3] rebuild: halo lists, [4] O(N) terms + spreading,
 * [5] sub-meshes to the mesh rank, [6] pair kernel (+ mesh chain on the mesh rank), [7] potential boxes back,
 * [8] gather, [9] halo forces, [10] update, [11] end of call (state / force gathers, energies), [12] steps counted. */
/* Relative pair-work share of every rank's domain (nranks doubles; NULL = equal): the domain volumes follow the weights, so a
 * rank with extra duties (the PME mesh rank) can be given a smaller domain.  Every rank must pass the same weights. */
int mdk_dd_set_weights(mdk_ctx *c, const double *weights) {
    if (!c || !c->dd) return MDK_ERR_BAD_ARG;
    for (int r = 0; r < DD_MAXR; ++r) c->dd_weight[r] = (weights && r < c->nranks) ? weights[r] : 0.0;
    for (int r = 0; weights && r < c->nranks; ++r)
        if (!(weights[r] > 0)) return fail(c, MDK_ERR_BAD_ARG, "mdk_dd_set_weights: weight %d = %g is not positive", r, weights[r]);
    c->nlist_valid = false;
    return MDK_OK;
}

int mdk_dd_trace(mdk_ctx *c, int on, double *out16) {
    if (!c || !c->dd) return MDK_ERR_BAD_ARG;
    if (out16) memcpy(out16, c->dd->t_phase, sizeof(c->dd->t_phase));
    if (on != (c->dd->trace ? 1 : 0)) { memset(c->dd->t_phase, 0, sizeof(c->dd->t_phase)); c->dd->trace = on != 0; }
    return MDK_OK;
}

/* out[0..7]: own tile slots lo / hi, halo atoms needed, own atoms sent to peers, halo exchanges so far, rebuilds,
 * PME sub-mesh points of this rank, ranks. */
int mdk_dd_stats(mdk_ctx *c, int64_t *out8) {
    if (!c || !out8) return MDK_ERR_BAD_ARG;
    for (int k = 0; k < 8; ++k) out8[k] = 0;
    if (!c->dd) { out8[1] = c->n; out8[7] = 1; return MDK_OK; }
    out8[0] = own_first(c); out8[1] = own_end(c); out8[2] = c->dd->n_need; out8[3] = c->dd->n_send;
    out8[4] = c->dd->stat_exchanges; out8[5] = c->dd->stat_rebuilds;
    out8[6] = c->dd->box.empty() ? 0 : (int64_t)c->dd->box[c->rank].pts; out8[7] = c->nranks;
    return MDK_OK;
}

}  // extern "C"
