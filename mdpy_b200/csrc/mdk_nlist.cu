// mdk_nlist.cu — tile (neighbour) list: cell sort, i-block bounding boxes, per-block
// filtered j-atom chunks with rotated exclusion / 1-4 masks.
//
// Replaces mdpy/core/cell_list.py:83-116 (dense -1-padded [nx,ny,nz,P] cell table rebuilt
// on every State.set_positions, state.py:61) and the 27-cell stencil walk of
// charmm_nonbonded_constraint.py:78-83.  Pair-set definition: every pair
// with minimum-image r <= rc that is not in bonded_particles — the list holds a superset
// (rc + skin against the i-block bounding box) and the pair kernel applies the canonical
// fp32 test.
//
// Layout.  Atoms are sorted by cell (x fastest), 32 consecutive sorted atoms form an
// i-block.  For each i-block the builder emits chunks of 32 j-atoms (tile-order indices)
// whose tile index is larger than the block's (half shell by index) and that lie within
// rc+skin of the block's periodic bounding box; chunk 0 is the block itself.  Chunks are
// grouped into work units of at most seg_chunks chunks taken from a global pool.
// A chunk that contains an excluded or 1-4 pair, padding, or belongs to a ragged last
// block carries 2 x 32 words of masks, already rotated so that lane l at rotation step k
// tests bit k (pair i = l, j-slot = (l + k) & 31).
#include <cub/cub.cuh>

#include <algorithm>

#include "mdk_common.cuh"

namespace mdk {

struct GridParams {
    int n, n_blocks;
    float L[3], invL[3];
    int ncell[3];
    float inv_cw[3];
    float R, R2;
    float Rn2;            // class radius^2 of the list order: j-atoms with an i-atom of the block within it come first ("near"),
                          // the rest (skin shell only) last; < 0: one class
    int far_flush;        // far-class staging: a chunk goes out early when more than this many atoms wait (<= FAR_CAP - 32)
    float skin_half2;
    float wide_lim[3];    // a block whose bounding-box half extent exceeds this on an axis is "wide": canonical minimum image
    DDGeom dd;            // cell numbering: domain by domain
    int dd_rank, dd_nranks;   // nranks > 1: this rank builds the lists of its own domain's i-blocks only
};

// index of the interval of a cut array that holds cell c; lo / hi = its first / one-past-last cell
__host__ __device__ __forceinline__ int dd_find(const int *cut, int p, int c, int &lo, int &hi) {
    int k = 0;
    while (k + 1 < p && c >= cut[k + 1]) ++k;
    lo = cut[k]; hi = cut[k + 1];
    return k;
}
// x slab of cell column cx (the list builder walks rows in pieces that stay inside one slab)
__host__ __device__ __forceinline__ int dd_axis_domain(const DDGeom &d, int a, int c, int &lo, int &hi) {
    return dd_find(d.cut0, d.pdim[0], c, lo, hi);   // only axis 0 is asked for
}
// key of cell (cx, cy, cz): cells of one domain are contiguous, x fastest.  Along an axis that is cut, the walk
// is a serpentine (x runs backwards on every other row, y on every other plane), so that consecutive cells — and
// with them the 32 consecutive sorted atoms of an i-block — stay spatial neighbours at a row / plane end; along an
// uncut axis the periodic wrap already makes the row end and the next row's start neighbours.
__host__ __device__ __forceinline__ int dd_cell_key(const DDGeom &d, int cx, int cy, int cz) {
    int lx, hx, ly, hy, lz, hz;
    const int dx = dd_find(d.cut0, d.pdim[0], cx, lx, hx);
    const int dy = dd_find(d.cut1[dx], d.pdim[1], cy, ly, hy);
    const int dz = dd_find(d.cut2[dx][dy], d.pdim[2], cz, lz, hz);
    const int dom = (dz * d.pdim[1] + dy) * d.pdim[0] + dx;
    const int nx = hx - lx, ny = hy - ly, pz = cz - lz;
    int py = cy - ly, px = cx - lx;
    if (d.pdim[1] > 1 && (pz & 1)) py = ny - 1 - py;
    const int row = pz * ny + py;
    if (d.pdim[0] > 1 && (row & 1)) px = nx - 1 - px;
    return d.dom_base[dom] + row * nx + px;
}

// ---------------------------------------------------------------------------
__global__ void k_cell_keys(int n, const double *__restrict__ x_cur, GridParams g,
                            const double3 Ld, unsigned *__restrict__ keys, int *__restrict__ idx) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    double L[3] = {Ld.x, Ld.y, Ld.z};
    int cidx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double x = x_cur[3 * a + d];
        float w = (float)(x - L[d] * rint(x / L[d]));
        int ci = (int)floorf((w + 0.5f * g.L[d]) * g.inv_cw[d]);
        cidx[d] = min(max(ci, 0), g.ncell[d] - 1);
    }
    keys[a] = (unsigned)dd_cell_key(g.dd, cidx[0], cidx[1], cidx[2]);
    idx[a] = a;
}

__global__ void k_cell_start(int n, int ncells, const unsigned *__restrict__ keys_sorted,
                             int *__restrict__ cell_start) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncells) return;
    // lower_bound(keys_sorted, c)
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (keys_sorted[mid] < (unsigned)c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = lo;
}

// first tile slot of every domain (the atoms whose cell lies in the domain are one contiguous range of the tile order);
// blk[ndom] = n.  A rank owns exactly the atoms of its domain.  An i-block that straddles a domain boundary is listed
// twice, once by each rank for its own lanes of the block (the other lanes are masked like the padding of a ragged
// last block), so every rank's i-blocks stay spatially compact.
__global__ void k_dd_bounds(DDGeom d, int n, const int *__restrict__ cell_start, int *__restrict__ blk) {
    const int ndom = d.pdim[0] * d.pdim[1] * d.pdim[2];
    const int r = threadIdx.x;
    if (r > ndom) return;
    blk[r] = r == ndom ? n : cell_start[d.dom_base[r]];
}

// tile-order gather of positions / charges / LJ parameters; also resets the displacement
// reference.  Slots >= n (ragged last block) are parked at the origin with zero charge/eps.
__global__ void k_gather_sorted(int n, int n_pad, const int *__restrict__ order,
                                const double *__restrict__ x_cur, const float *__restrict__ q,
                                const float4 *__restrict__ lj4, bool have_lj, float sqrt_ke,
                                const double3 Ld, float4 *__restrict__ xs, float4 *__restrict__ xs_ref,
                                float4 *__restrict__ ljs, int *__restrict__ inv_order) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pad) return;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), l = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < n) {
        int a = order[k];
        inv_order[a] = k;
        double L[3] = {Ld.x, Ld.y, Ld.z};
        float w[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double x = x_cur[3 * a + d];
            w[d] = (float)(x - L[d] * rint(x / L[d]));
        }
        p = make_float4(w[0], w[1], w[2], q[a] * sqrt_ke);
        if (have_lj) {
            float4 v = lj4[a];
            l = make_float4(2.f * sqrtf(v.x), 0.5f * v.y, 2.f * sqrtf(v.z), 0.5f * v.w);
        }
    }
    xs[k] = p; xs_ref[k] = p; ljs[k] = l;
}

// exclusion / 1-4 tables translated to tile slots
__global__ void k_tables_sorted(int n, const int *__restrict__ order, const int *__restrict__ inv_order,
                                const int *__restrict__ tab, int w, int *__restrict__ tab_s) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * w) return;
    int k = t / w, e = t - k * w;
    int p = tab[(size_t)order[k] * w + e];
    tab_s[t] = p < 0 ? -1 : inv_order[p];
}

// xs <- wrap(x_cur) in the existing tile order + skin/2 displacement check (flags[1])
__global__ void k_refresh_sorted(int n, const int *__restrict__ order, const double *__restrict__ x_cur,
                                 const double3 Ld, GridParams g, float4 *__restrict__ xs,
                                 const float4 *__restrict__ xs_ref, int *__restrict__ flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int a = order[k];
    double L[3] = {Ld.x, Ld.y, Ld.z};
    float w[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double x = x_cur[3 * a + d];
        w[d] = (float)(x - L[d] * rint(x / L[d]));
    }
    float4 r = xs_ref[k];
    float dx = min_image(w[0] - r.x, g.L[0], g.invL[0]);
    float dy = min_image(w[1] - r.y, g.L[1], g.invL[1]);
    float dz = min_image(w[2] - r.z, g.L[2], g.invL[2]);
    if (dist2(dx, dy, dz) > g.skin_half2) flags[1] = 1;
    float4 p = xs[k];
    xs[k] = make_float4(w[0], w[1], w[2], p.w);
}

// one warp per i-block: periodic bounding box relative to the block's first atom
// (decomposed runs: over the lanes this rank owns — other ranks' blocks are skipped, a straddling block gets the box of
// the own part)
__global__ void k_block_bbox(GridParams g, const float4 *__restrict__ xs, const int *__restrict__ dd_blk, float4 *__restrict__ bbc,
                             float4 *__restrict__ bbh, int *__restrict__ bb_max) {
    int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (b >= g.n_blocks) return;
    int k = b * TILE + lane;
    bool valid = k < g.n;
    if (g.dd_nranks > 1) valid = valid && k >= dd_blk[g.dd_rank] && k < dd_blk[g.dd_rank + 1];
    const unsigned any = __ballot_sync(0xffffffffu, valid);
    if (!any) return;
    const int first = __ffs(any) - 1;
    float4 p = xs[valid ? k : b * TILE + first];
    float rx = __shfl_sync(0xffffffffu, p.x, first), ry = __shfl_sync(0xffffffffu, p.y, first),
          rz = __shfl_sync(0xffffffffu, p.z, first);
    float d[3] = {min_image(p.x - rx, g.L[0], g.invL[0]), min_image(p.y - ry, g.L[1], g.invL[1]),
                  min_image(p.z - rz, g.L[2], g.invL[2])};
    float mn[3], mx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = mx[a] = d[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if (lane == 0) {
        bbc[b] = make_float4(rx + 0.5f * (mn[0] + mx[0]), ry + 0.5f * (mn[1] + mx[1]),
                             rz + 0.5f * (mn[2] + mx[2]), 0.f);
        bbh[b] = make_float4(0.5f * (mx[0] - mn[0]) + 1e-4f, 0.5f * (mx[1] - mn[1]) + 1e-4f,
                             0.5f * (mx[2] - mn[2]) + 1e-4f, 0.f);
#pragma unroll
        for (int a = 0; a < 3; ++a)  // non-negative floats order like their bit patterns
            atomicMax(&bb_max[a], __float_as_int(0.5f * (mx[a] - mn[a]) + 1e-4f));
    }
}

// ---------------------------------------------------------------------------
struct BuildOut {
    int4 *units;
    int *chunk_j, *chunk_mask;
    unsigned *mask_excl, *mask_14;
    int *counters;  // [0]=units [1]=chunks [2]=mask slots
    int *flags;     // [2]=overflow
    int cap_units, cap_chunks, cap_masks;
    int seg;
    int *mark;              // domain decomposition: listed j-atoms of other domains (halo), or null
    int own_lo, own_hi;     // own tile slots
};

struct BlockEmitter {
    int b, lane, n;
    int wide;
    bool i_valid, ragged;
    int seg_base, seg_fill;
    const int *excl_s, *p14_s;
    int wb, ws;
    BuildOut o;

    __device__ int alloc(int *ctr, int count) {
        int v = 0;
        if (lane == 0) v = atomicAdd(ctr, count);
        return __shfl_sync(0xffffffffu, v, 0);
    }
    __device__ void close_unit() {
        if (seg_fill > 0) {
            int u = alloc(&o.counters[0], 1);
            if (u < o.cap_units) {
                if (lane == 0) { o.units[u] = make_int4(b, seg_base, seg_fill, wide); atomicAdd(&o.counters[4], seg_fill); }
            } else if (lane == 0) o.flags[2] = 1;
        }
        seg_fill = 0;
    }
    __device__ void begin() { seg_base = alloc(&o.counters[1], o.seg); seg_fill = 0; }
    // j: tile index of this lane's j atom, or -1 (padding)
    __device__ void emit(int j, bool diagonal) {
        if (seg_fill == o.seg) { close_unit(); seg_base = alloc(&o.counters[1], o.seg); }
        int chunk = seg_base + seg_fill;
        ++seg_fill;
        unsigned jm = 0u, j14 = 0u;
        if (j < 0) {
            jm = 0xffffffffu;
        } else {
            for (int e = 0; e < wb; ++e) {
                int p = excl_s[(size_t)j * wb + e];
                if (p >= 0 && (p >> 5) == b) jm |= 1u << (p & 31);
            }
            for (int e = 0; e < ws; ++e) {
                int p = p14_s[(size_t)j * ws + e];
                if (p >= 0 && (p >> 5) == b) j14 |= 1u << (p & 31);
            }
            j14 &= ~jm;  // excluded wins over 1-4 (charmm_nonbonded_constraint.py:83 before :92)
            if (diagonal) jm |= ~((1u << lane) - 1u);  // keep only i-slot < j-slot
        }
        bool flagged = __ballot_sync(0xffffffffu, (jm | j14) != 0u) != 0u || ragged;
        bool ok = chunk + 1 <= o.cap_chunks;
        if (!ok && lane == 0) o.flags[2] = 1;
        int slot = -1;
        if (flagged) {
            unsigned mi = 0u, m14 = 0u;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                int src = (lane + k) & 31;
                unsigned v = __shfl_sync(0xffffffffu, jm, src);
                unsigned w = __shfl_sync(0xffffffffu, j14, src);
                mi |= ((v >> lane) & 1u) << k;
                m14 |= ((w >> lane) & 1u) << k;
            }
            if (!i_valid) mi = 0xffffffffu;
            slot = alloc(&o.counters[2], 1);
            if (slot < o.cap_masks) {
                o.mask_excl[(size_t)slot * 32 + lane] = mi;
                o.mask_14[(size_t)slot * 32 + lane] = m14;
            } else {
                if (lane == 0) o.flags[2] = 1;
                slot = -1;
            }
        }
        if (ok) {
            o.chunk_j[(size_t)chunk * 32 + lane] = j < 0 ? 0 : j;
            if (lane == 0) o.chunk_mask[chunk] = slot;
            if (o.mark && j >= 0 && (j < o.own_lo || j >= o.own_hi)) o.mark[j] = 1;
        }
    }
};

__device__ __forceinline__ int imod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

constexpr int BUILD_WARPS = 4;
// j-atoms that pass the list filter but have no i-atom of the block within the cutoff itself (they sit in the skin shell:
// ~30 % of a list at rc 12 / skin 2) wait here and are emitted after the block's other atoms, so they end up in chunks of
// their own.  In those chunks k_pair finds (almost) nothing inside the cutoff and every rotation step leaves after the
// distance test — 16 instead of 80 warp instructions; mixed into the other chunks they would cost full steps.
constexpr int FAR_CAP = 1024;

__global__ void __launch_bounds__(BUILD_WARPS * 32)
k_build_lists(GridParams g, int n_parts, const float4 *__restrict__ xs, const float4 *__restrict__ bbc,
              const float4 *__restrict__ bbh, const int *__restrict__ cell_start, const int *__restrict__ dd_blk,
              const int *__restrict__ excl_s, int wb, const int *__restrict__ p14_s, int ws,
              BuildOut o) {
    __shared__ int stage_all[BUILD_WARPS][64];
    __shared__ int far_all[BUILD_WARPS][FAR_CAP];
    __shared__ float4 s_xi[BUILD_WARPS][32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *stage = stage_all[wid], *farb = far_all[wid];
    const bool split = g.Rn2 > 0.f;
    int warp_global = blockIdx.x * BUILD_WARPS + wid, n_warps = gridDim.x * BUILD_WARPS;
    // domain decomposition: this rank lists the i-blocks of its own domain only.  Pairs inside the domain are
    // taken half shell by tile index as on one GPU; a pair of blocks that straddles two domains is taken by the
    // lower block when the sum of the two block indices is even, by the higher one when it is odd — both ranks
    // evaluate the same rule on the same global order, exactly one of them lists the pair, and the cross-domain
    // work is shared evenly whichever side of a boundary a block sits on.
    const bool multi = g.dd_nranks > 1;
    const int own_lo = multi ? dd_blk[g.dd_rank] : 0, own_hi = multi ? dd_blk[g.dd_rank + 1] : g.n_blocks * TILE;
    const int blk_lo = own_lo / TILE, blk_hi = (own_hi + TILE - 1) / TILE;      // the first / last one may be shared with a neighbour rank
    if (multi) { o.own_lo = own_lo; o.own_hi = own_hi; } else { o.mark = nullptr; }
    // work item = (i-block, part): the candidate rows of a block are dealt round-robin to n_parts warps,
    // each emitting its own work units, so small systems still fill the machine during a rebuild
    for (int w = warp_global; w < (blk_hi - blk_lo) * n_parts; w += n_warps) {
        const int b = blk_lo + w / n_parts, part = w % n_parts;
        float4 c = bbc[b], h = bbh[b];
        BlockEmitter em;
        em.b = b; em.lane = lane; em.n = g.n;
        // hoisted minimum image (k_pair<..., SHIFT>): every listed j must have a unique image within L/2 of the
        // block centre.  An interacting j is within R of some i-atom, which is within h of the centre, so
        // R + h <= L/2 on every axis is enough (pairs farther apart than R may then see a non-minimal image, but
        // both distances exceed the cutoff).  Blocks that break the bound are flagged and evaluated canonically.
        em.wide = (h.x > g.wide_lim[0] || h.y > g.wide_lim[1] || h.z > g.wide_lim[2]) ? 1 : 0;
        const int my_slot = b * TILE + lane;
        em.i_valid = my_slot < g.n && my_slot >= own_lo && my_slot < own_hi;
        em.ragged = ((b == g.n_blocks - 1) && (g.n & 31)) || b * TILE < own_lo || (b + 1) * TILE > own_hi;
        em.excl_s = excl_s; em.p14_s = p14_s; em.wb = wb; em.ws = ws; em.o = o;
        // i-atom positions of this block for the exact filter (ragged lanes repeat the first atom)
        __syncwarp();
        {   // i-atom positions for the exact filter, relative to the block centre; lanes this rank does not own repeat an
            // own atom of the block
            const unsigned vm = __ballot_sync(0xffffffffu, em.i_valid);
            float4 q = xs[em.i_valid ? my_slot : b * TILE + (vm ? __ffs(vm) - 1 : 0)];
            q.x = min_image(q.x - c.x, g.L[0], g.invL[0]);
            q.y = min_image(q.y - c.y, g.L[1], g.invL[1]);
            q.z = min_image(q.z - c.z, g.L[2], g.invL[2]);
            s_xi[wid][lane] = q;
        }
        __syncwarp();
        em.begin();
        if (part == 0) {
            int k = b * TILE + lane;
            em.emit(k < g.n ? k : -1, true);
        }
        const int first_j = (b + 1) * TILE;
        // candidate cell ranges: bbox +- R (plus a rounding margin), periodic
        float cc[3] = {c.x, c.y, c.z}, hh[3] = {h.x, h.y, h.z};
        int lo[3], len[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float ext = hh[a] + g.R + 1e-3f + 1e-5f * g.L[a];
            int l0 = (int)floorf((cc[a] - ext + 0.5f * g.L[a]) * g.inv_cw[a]);
            int h0 = (int)floorf((cc[a] + ext + 0.5f * g.L[a]) * g.inv_cw[a]);
            int ln = h0 - l0 + 1;
            if (ln >= g.ncell[a]) { l0 = 0; ln = g.ncell[a]; }
            lo[a] = imod(l0, g.ncell[a]);
            len[a] = ln;
        }
        int nstage = 0, nfar = 0;
        const float R2m = g.R2 * (1.f + 2e-5f);    // the hoisted differences round differently from the pair kernel's
        const float cwy = 1.f / g.inv_cw[1], cwz = 1.f / g.inv_cw[2];
        for (int iz = 0; iz < len[2]; ++iz) {
            int zz = lo[2] + iz; if (zz >= g.ncell[2]) zz -= g.ncell[2];
            for (int iy = 0; iy < len[1]; ++iy) {
                if ((iz * len[1] + iy) % n_parts != part) continue;
                int yy = lo[1] + iy; if (yy >= g.ncell[1]) yy -= g.ncell[1];
                // the x range in pieces that neither wrap around the box nor cross a domain cut: inside a piece
                // the cells are consecutive keys, i.e. one contiguous run of the tile order
                int xg = lo[0], rem = len[0];
                {   // the row's own x range: what is left of R after the row's y / z distance from the bounding box (a sphere
                    // around the box instead of a cube: a third fewer candidates); rows out of reach are skipped
                    const float yc = (yy + 0.5f) * cwy - 0.5f * g.L[1], zc = (zz + 0.5f) * cwz - 0.5f * g.L[2];
                    const float ry = fmaxf(fabsf(min_image(yc - cc[1], g.L[1], g.invL[1])) - hh[1] - 0.5f * cwy - 1e-3f - 1e-5f * g.L[1], 0.f);
                    const float rz = fmaxf(fabsf(min_image(zc - cc[2], g.L[2], g.invL[2])) - hh[2] - 0.5f * cwz - 1e-3f - 1e-5f * g.L[2], 0.f);
                    const float rem2 = R2m - ry * ry - rz * rz;
                    if (rem2 < 0.f) continue;
                    const float ext = hh[0] + sqrtf(rem2) + 1e-3f + 1e-5f * g.L[0];
                    const int l0 = (int)floorf((cc[0] - ext + 0.5f * g.L[0]) * g.inv_cw[0]);
                    const int h0 = (int)floorf((cc[0] + ext + 0.5f * g.L[0]) * g.inv_cw[0]);
                    if (h0 - l0 + 1 < g.ncell[0]) { xg = imod(l0, g.ncell[0]); rem = h0 - l0 + 1; }
                }
#pragma unroll 1
                while (rem > 0) {
                    int dlo, dhi;
                    dd_axis_domain(g.dd, 0, xg, dlo, dhi);
                    const int cn = min(rem, dhi - xg);
                    const int key0 = min(dd_cell_key(g.dd, xg, yy, zz), dd_cell_key(g.dd, xg + cn - 1, yy, zz));   // a row may run backwards
                    int s = cell_start[key0], e = cell_start[key0 + cn];
                    xg += cn; if (xg >= g.ncell[0]) xg = 0;
                    rem -= cn;
                    if (!multi) s = max(s, first_j);
                    // decomposed: inside the own domain only j > b counts (half shell by index), except the few slots
                    // at the start of the domain's range that sit in the previous rank's straddling block
                    int skip_lo = 0, skip_hi = 0;
                    if (multi) { skip_lo = max(s, own_lo); skip_hi = min(e, min(first_j, own_hi)); }
                    for (int base = s; base < e; base += 32) {
                        if (base >= skip_lo && base + 32 <= skip_hi) continue;   // a whole stride of own atoms at or before this block
                        int j = base + lane;
                        bool pass = false, near = !split;
                        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                        bool cand = j < e;
                        if (multi && cand) {
                            const int bj = j >> 5;
                            if (j >= own_lo && j < own_hi) cand = bj > b;
                            else cand = ((b + bj) & 1) ? (bj < b) : (bj > b);
                        }
                        float hx = 0.f, hy = 0.f, hz = 0.f;
                        if (cand) {
                            p = xs[j];
                            hx = min_image(p.x - cc[0], g.L[0], g.invL[0]);
                            hy = min_image(p.y - cc[1], g.L[1], g.invL[1]);
                            hz = min_image(p.z - cc[2], g.L[2], g.invL[2]);
                            const float dx = fmaxf(fabsf(hx) - hh[0], 0.f), dy = fmaxf(fabsf(hy) - hh[1], 0.f),
                                        dz = fmaxf(fabsf(hz) - hh[2], 0.f);
                            cand = dx * dx + dy * dy + dz * dz <= R2m;     // within R of the bounding box
                        }
                        if (!__any_sync(0xffffffffu, cand)) continue;
                        // exact filter (the box test alone keeps ~25 % more atoms than needed): lane = candidate, smallest
                        // distance to the 32 i-atoms, which sit in shared memory relative to the block centre
                        float dmin = 3.0e38f;
                        if (!em.wide) {
#pragma unroll
                            for (int t = 0; t < 32; ++t) {
                                const float4 q = s_xi[wid][t];
                                dmin = fminf(dmin, dist2(hx - q.x, hy - q.y, hz - q.z));
                            }
                        } else {          // the block is too wide for one image per atom: minimum image per pair
#pragma unroll 4
                            for (int t = 0; t < 32; ++t) {
                                const float4 q = s_xi[wid][t];
                                dmin = fminf(dmin, dist2(min_image(hx - q.x, g.L[0], g.invL[0]), min_image(hy - q.y, g.L[1], g.invL[1]),
                                                         min_image(hz - q.z, g.L[2], g.invL[2])));
                            }
                        }
                        pass = cand && dmin <= R2m;
                        if (split) near = dmin <= g.Rn2;
                        const bool p_near = pass && near, p_far = pass && !near;
                        unsigned bal = __ballot_sync(0xffffffffu, p_near);
                        if (p_near) stage[nstage + __popc(bal & ((1u << lane) - 1u))] = j;
                        nstage += __popc(bal);
                        if (split) {
                            bal = __ballot_sync(0xffffffffu, p_far);
                            if (p_far) farb[nfar + __popc(bal & ((1u << lane) - 1u))] = j;
                            nfar += __popc(bal);
                        }
                        __syncwarp();
                        if (nfar > g.far_flush) {       // buffer full: a chunk of far atoms goes out early
                            const int jj = farb[nfar - 32 + lane];
                            nfar -= 32;
                            __syncwarp();
                            em.emit(jj, false);
                        }
                        if (nstage >= 32) {
                            int jj = stage[lane];
                            int rest = stage[32 + lane];
                            __syncwarp();
                            em.emit(jj, false);
                            stage[lane] = rest;
                            nstage -= 32;
                            __syncwarp();
                        }
                    }
                }
            }
        }
        if (nstage > 0) {      // the last near atoms, topped up with far ones
            const int take = min(32 - nstage, nfar);
            const int jj = lane < nstage ? stage[lane] : (lane - nstage < take ? farb[nfar - take + lane - nstage] : -1);
            nfar -= take;
            em.emit(jj, false);
        }
        for (int base = 0; base < nfar; base += 32) em.emit(base + lane < nfar ? farb[base + lane] : -1, false);
        __syncwarp();
        em.close_unit();
    }
}

// ---------------------------------------------------------------------------
static GridParams make_grid_params(mdk_ctx *c) {
    GridParams g{};
    g.n = c->n;
    g.n_blocks = c->n_blocks;
    for (int a = 0; a < 3; ++a) {
        g.L[a] = c->box.L[a]; g.invL[a] = c->box.invL[a];
        g.ncell[a] = c->ncell[a];
        g.inv_cw[a] = 1.0f / c->cellw[a];
    }
    float rc = fmaxf(c->have_lj ? c->rc_lj : 0.f, c->have_coul ? c->rc_coul : 0.f);
    g.R = rc + c->skin;
    g.R2 = g.R * g.R;
    g.Rn2 = c->far_split ? rc * rc : -1.f;
    g.far_flush = c->far_flush;
    for (int a = 0; a < 3; ++a) g.wide_lim[a] = 0.5f * c->box.L[a] - g.R - 0.05f;
    g.skin_half2 = 0.25f * c->skin * c->skin;
    g.dd = c->dd_geom;
    g.dd_rank = c->dd ? c->rank : 0; g.dd_nranks = c->dd ? c->nranks : 1;
    return g;
}

static int check_cutoff(mdk_ctx *c) {
    float rc = fmaxf(c->have_lj ? c->rc_lj : 0.f, c->have_coul ? c->rc_coul : 0.f);
    // cell_list.py:56-69: cutoff 0 or floor(L / rc) < 2 is a poorly defined list
    if (c->have_lj && c->rc_lj <= 0.f)
        return fail(c, MDK_ERR_CUTOFF_TOO_LARGE, "Cutoff radius is poor defined, current value %.3f", c->rc_lj);
    for (int a = 0; a < 3; ++a)
        if (rc > 0.f && floor(c->box.Ld[a] / rc) < 2)
            return fail(c, MDK_ERR_CUTOFF_TOO_LARGE, "The cutoff_radius is too large to create cell list");
    return MDK_OK;
}

int nlist_refresh_sorted(mdk_ctx *c) {
    if (!c->nlist_valid) return MDK_OK;
    GridParams g = make_grid_params(c);
    double3 Ld = make_double3(c->box.Ld[0], c->box.Ld[1], c->box.Ld[2]);
    k_refresh_sorted<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->order.p, c->x_cur.p, Ld, g,
                                                               c->xs.p, c->xs_ref.p, c->flags.p);
    ++c->n_launches;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

NlistView nlist_view(mdk_ctx *c) {
    NlistView v;
    v.units = c->units.p; v.n_units = c->counters.p;
    v.chunk_j = c->chunk_j.p; v.chunk_mask = c->chunk_mask.p;
    v.mask_excl = c->mask_excl.p; v.mask_14 = c->mask_14.p;
    return v;
}

// ---- rebuild = plan (host) + enqueue (device work only) ------------------------------------
// plan: cell geometry, work-unit granularity and pool capacities from the box and the density.
static int nlist_plan(mdk_ctx *c) {
    if (!c->have_box || c->n <= 0 || !c->have_pos)
        return fail(c, MDK_ERR_NOT_BOUND, "neighbour list needs box, atoms and positions");
    MDK_TRY(check_cutoff(c));
    const int n = c->n;
    c->n_blocks = (n + TILE - 1) / TILE;
    c->n_pad = c->n_blocks * TILE;
    // cell geometry: y/z width ~ cbrt(32 / rho) so that 32 consecutive sorted atoms are
    // roughly cubic, x width half of that (finer range ends along the contiguous axis)
    double V = c->box.Ld[0] * c->box.Ld[1] * c->box.Ld[2];
    double rho = n / V;
    float rc = fmaxf(c->have_lj ? c->rc_lj : 0.f, c->have_coul ? c->rc_coul : 0.f);
    double R = rc + c->skin;
    double cyz = cbrt(32.0 / rho);
    if (cyz < 3.0) cyz = 3.0;
    if (cyz > R && R > 0) cyz = R;
    double target[3] = {0.5 * cyz, cyz, cyz};
    long long ncells = 1;
    for (int a = 0; a < 3; ++a) {
        int nc = (int)floor(c->box.Ld[a] / target[a]);
        if (nc < 1) nc = 1;
        if (nc > 1024) nc = 1024;
        const int pd = c->dd ? c->dd_geom.pdim[a] : 1;
        if (pd > 1) { nc -= nc % pd; if (nc < pd) nc = pd; }      // equal cell counts per domain
        c->ncell[a] = nc;
        c->cellw[a] = (float)(c->box.Ld[a] / nc);
        ncells *= nc;
    }
    c->n_cells = ncells;
    // cell numbering: one domain, or the domain grid of mdk_dd_init — recursive bisection of the cell grid by the ranks'
    // work weights (equal weights: equal cell counts)
    {
        DDGeom &d = c->dd_geom;
        const DDGeom before = d;
        int p[3];
        for (int a = 0; a < 3; ++a) {
            p[a] = c->dd ? d.pdim[a] : 1;
            if (p[a] < 1) p[a] = 1;
            if (p[a] > c->ncell[a]) return fail(c, MDK_ERR_BAD_ARG, "domain grid %d along axis %d exceeds the %d cells of the box", p[a], a, c->ncell[a]);
            d.pdim[a] = p[a];
        }
        auto weight = [&](int dx, int dy, int dz) {
            const double w = c->dd ? c->dd_weight[(dz * p[1] + dy) * p[0] + dx] : 1.0;
            return w > 0 ? w : 1.0;
        };
        // cut `cells` cells into n intervals proportional to w[0..n), every interval at least one cell
        auto split = [&](int cells, int n, const double *w, int *cut) {
            double tot = 0; for (int k = 0; k < n; ++k) tot += w[k];
            double acc = 0;
            cut[0] = 0;
            for (int k = 1; k < n; ++k) {
                acc += w[k - 1];
                int v = (int)floor(cells * acc / tot + 0.5);
                if (v < cut[k - 1] + 1) v = cut[k - 1] + 1;
                if (v > cells - (n - k)) v = cells - (n - k);
                cut[k] = v;
            }
            for (int k = n; k <= DD_MAXP; ++k) cut[k] = cells;
        };
        double wx[DD_MAXP] = {0};
        for (int dx = 0; dx < p[0]; ++dx)
            for (int dy = 0; dy < p[1]; ++dy)
                for (int dz = 0; dz < p[2]; ++dz) wx[dx] += weight(dx, dy, dz);
        split(c->ncell[0], p[0], wx, d.cut0);
        for (int dx = 0; dx < DD_MAXP; ++dx) {
            double wy[DD_MAXP] = {0};
            for (int dy = 0; dy < p[1]; ++dy)
                for (int dz = 0; dz < p[2]; ++dz) wy[dy] += dx < p[0] ? weight(dx, dy, dz) : 1.0;
            split(c->ncell[1], p[1], wy, d.cut1[dx]);
            for (int dy = 0; dy < DD_MAXP; ++dy) {
                double wz[DD_MAXP] = {0};
                for (int dz = 0; dz < p[2]; ++dz) wz[dz] = (dx < p[0] && dy < p[1]) ? weight(dx, dy, dz) : 1.0;
                split(c->ncell[2], p[2], wz, d.cut2[dx][dy]);
            }
        }
        int base = 0, dom = 0;
        int size[DD_MAXR] = {0};
        for (int dz = 0; dz < p[2]; ++dz)
            for (int dy = 0; dy < p[1]; ++dy)
                for (int dx = 0; dx < p[0]; ++dx)
                    size[(dz * p[1] + dy) * p[0] + dx] = (d.cut0[dx + 1] - d.cut0[dx]) * (d.cut1[dx][dy + 1] - d.cut1[dx][dy]) *
                                                         (d.cut2[dx][dy][dz + 1] - d.cut2[dx][dy][dz]);
        const int ndom = p[0] * p[1] * p[2];
        for (dom = 0; dom < ndom; ++dom) { d.dom_base[dom] = base; base += size[dom]; }
        for (; dom <= DD_MAXR; ++dom) d.dom_base[dom] = base;
        if (memcmp(&before, &d, sizeof(d)) != 0) ++c->graph_epoch;   // kernel arguments of captured rebuilds
    }
    const int blocks_here = std::max(1, c->dd ? c->n_blocks / c->nranks : c->n_blocks);   // i-blocks this rank lists
    c->n_parts = (int)((4096 + blocks_here - 1) / blocks_here);
    if (c->n_parts < 1) c->n_parts = 1;
    if (c->n_parts > 8) c->n_parts = 8;
    // work-unit granularity: enough units to fill the machine a few times over
    double per_block = rho * (4.0 / 3.0) * M_PI * R * R * R * 0.5 * 2.2 / 32.0 + 1.0;  // chunks (estimate)
    double total_chunks = per_block * c->n_blocks;
    // enough units to fill the machine's resident warps (4 blocks x 8 warps per SM) unit_waves times over: the tail of a
    // launch is one unit long, so more, shorter units balance better (a decomposed rank lists only its share of the chunks)
    double want_units = 32.0 * c->sm_count * c->unit_waves;
    int seg = (int)(total_chunks / (c->dd ? c->nranks : 1) / want_units);
    if (seg < 2) seg = 2;
    if (seg > 16) seg = 16;
    c->seg_chunks = seg;
    const double slack = c->graph_pools ? 2.0 : 1.3;  // rebuilds inside a CUDA graph cannot grow the pools
    size_t est_chunks = (size_t)(total_chunks * slack) + (size_t)c->n_blocks * c->n_parts * (seg + 1) + 1024;
    if (c->cap_chunks < est_chunks) c->cap_chunks = est_chunks;
    size_t est_units = est_chunks / seg + (size_t)c->n_blocks * (c->n_parts + 1) + 1024;
    if (c->cap_units < est_units) c->cap_units = est_units;
    size_t est_masks = (size_t)c->n_blocks * (c->graph_pools ? 24 : 12) + 1024;
    if (c->cap_masks < est_masks) c->cap_masks = est_masks;

    MDK_CUDA(c, c->cell_key.reserve(n)); MDK_CUDA(c, c->cell_key_sorted.reserve(n));
    MDK_CUDA(c, c->idx_tmp.reserve(n)); MDK_CUDA(c, c->order.reserve(n));
    MDK_CUDA(c, c->inv_order.reserve(n)); MDK_CUDA(c, c->cell_start.reserve((size_t)ncells + 1));
    MDK_CUDA(c, c->xs.reserve(c->n_pad)); MDK_CUDA(c, c->xs_ref.reserve(c->n_pad));
    MDK_CUDA(c, c->ljs.reserve(c->n_pad)); MDK_CUDA(c, c->f_acc.reserve((size_t)c->n_pad * 3));
    MDK_CUDA(c, c->bb_center.reserve(c->n_blocks)); MDK_CUDA(c, c->bb_half.reserve(c->n_blocks));
    MDK_CUDA(c, c->dd_blk.reserve(DD_MAXR + 1));
    if (c->dd) MDK_CUDA(c, c->dd_mark.reserve(c->n_pad));
    MDK_CUDA(c, c->excl_s.reserve((size_t)n * (c->wb > 0 ? c->wb : 1)));
    MDK_CUDA(c, c->p14_s.reserve((size_t)n * (c->ws > 0 ? c->ws : 1)));
    c->sort_end_bit = 1;
    while ((1ll << c->sort_end_bit) < ncells) ++c->sort_end_bit;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->cell_key.p, c->cell_key_sorted.p, c->idx_tmp.p,
                                    c->order.p, n, 0, c->sort_end_bit, c->stream);
    c->sort_tmp_bytes = tmp_bytes;
    {
        const void *p0 = c->sort_buf.p, *p1 = c->xs.p, *p2 = c->cell_start.p;
        MDK_CUDA(c, c->sort_buf.reserve(tmp_bytes));
        if (p0 != c->sort_buf.p || p1 != c->xs.p || p2 != c->cell_start.p) ++c->graph_epoch;
    }
    return MDK_OK;
}

static int nlist_reserve_pools(mdk_ctx *c) {
    const void *before[3] = {c->units.p, c->chunk_j.p, c->mask_excl.p};
    MDK_CUDA(c, c->units.reserve(c->cap_units));
    MDK_CUDA(c, c->chunk_j.reserve(c->cap_chunks * 32));
    MDK_CUDA(c, c->chunk_mask.reserve(c->cap_chunks));
    MDK_CUDA(c, c->mask_excl.reserve(c->cap_masks * 32));
    MDK_CUDA(c, c->mask_14.reserve(c->cap_masks * 32));
    if (before[0] != c->units.p || before[1] != c->chunk_j.p || before[2] != c->mask_excl.p)
        ++c->graph_epoch;   // a captured step graph holds the old pool addresses
    return MDK_OK;
}

// Bookkeeping at the end of a rebuild that runs inside a CUDA graph (no host in the loop):
// sticky error bits in flags[3] (1 = pool overflow, 2 = a block outgrew the hoisted-minimum-image
// bound), rebuild counter and list sizes in counters[12..15].
__global__ void k_after_build(int *counters, int *flags) {
    if (flags[2]) flags[3] |= 1;
    counters[12] += 1;
    counters[13] = counters[0]; counters[14] = counters[4]; counters[15] = counters[2];
}

// enqueue: every kernel of a rebuild on c->stream, fixed launch shapes, no host synchronisation
// (usable inside stream capture).
int nlist_enqueue(mdk_ctx *c, bool in_graph) {
    const int n = c->n, T = 256;
    GridParams g = make_grid_params(c);
    double3 Ld = make_double3(c->box.Ld[0], c->box.Ld[1], c->box.Ld[2]);
    const long long ncells = c->n_cells;
    k_cell_keys<<<(n + T - 1) / T, T, 0, c->stream>>>(n, c->x_cur.p, g, Ld, c->cell_key.p, c->idx_tmp.p);
    size_t tmp_bytes = c->sort_tmp_bytes;
    MDK_CUDA(c, cub::DeviceRadixSort::SortPairs(c->sort_buf.p, tmp_bytes, c->cell_key.p, c->cell_key_sorted.p,
                                                c->idx_tmp.p, c->order.p, n, 0, c->sort_end_bit, c->stream));
    k_cell_start<<<(int)((ncells + 1 + T - 1) / T), T, 0, c->stream>>>(n, (int)ncells, c->cell_key_sorted.p,
                                                                       c->cell_start.p);
    k_dd_bounds<<<1, DD_MAXR + 1, 0, c->stream>>>(g.dd, n, c->cell_start.p, c->dd_blk.p);
    if (c->dd) MDK_CUDA(c, cudaMemsetAsync(c->dd_mark.p, 0, (size_t)c->n_pad * sizeof(int), c->stream));
    float sqrt_ke = c->have_coul ? (float)sqrt(c->k_e) : 0.f;
    k_gather_sorted<<<(c->n_pad + T - 1) / T, T, 0, c->stream>>>(n, c->n_pad, c->order.p, c->x_cur.p, c->q.p,
                                                                 c->lj4.p, c->have_lj, sqrt_ke, Ld, c->xs.p,
                                                                 c->xs_ref.p, c->ljs.p, c->inv_order.p);
    if (c->wb > 0)
        k_tables_sorted<<<(n * c->wb + T - 1) / T, T, 0, c->stream>>>(n, c->order.p, c->inv_order.p, c->excl.p,
                                                                      c->wb, c->excl_s.p);
    if (c->ws > 0)
        k_tables_sorted<<<(n * c->ws + T - 1) / T, T, 0, c->stream>>>(n, c->order.p, c->inv_order.p, c->p14.p,
                                                                      c->ws, c->p14_s.p);
    MDK_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 12 * sizeof(int), c->stream));
    MDK_CUDA(c, cudaMemsetAsync(c->flags.p + 1, 0, 2 * sizeof(int), c->stream));
    k_block_bbox<<<(c->n_blocks * 32 + T - 1) / T, T, 0, c->stream>>>(g, c->xs.p, c->dd_blk.p, c->bb_center.p, c->bb_half.p,
                                                                    c->counters.p + 8);
    BuildOut o;
    o.units = c->units.p; o.chunk_j = c->chunk_j.p; o.chunk_mask = c->chunk_mask.p;
    o.mask_excl = c->mask_excl.p; o.mask_14 = c->mask_14.p;
    o.counters = c->counters.p; o.flags = c->flags.p;
    o.cap_units = (int)c->cap_units; o.cap_chunks = (int)c->cap_chunks; o.cap_masks = (int)c->cap_masks;
    o.seg = c->seg_chunks;
    o.mark = c->dd ? c->dd_mark.p : nullptr; o.own_lo = 0; o.own_hi = c->n_pad;
    int blocks = (c->n_blocks * c->n_parts + BUILD_WARPS - 1) / BUILD_WARPS;
    int max_blocks = c->sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    k_build_lists<<<blocks, BUILD_WARPS * 32, 0, c->stream>>>(g, c->n_parts, c->xs.p, c->bb_center.p, c->bb_half.p,
                                                             c->cell_start.p, c->dd_blk.p, c->excl_s.p, c->wb, c->p14_s.p,
                                                             c->ws, o);
    if (in_graph) k_after_build<<<1, 1, 0, c->stream>>>(c->counters.p, c->flags.p);
    c->n_launches += in_graph ? 0 : 10;
    MDK_CUDA(c, cudaGetLastError());
    return MDK_OK;
}

int nlist_rebuild(mdk_ctx *c) {
    MDK_TRY(nlist_plan(c));
    PhaseTimer pt(c, PH_NLIST);
    GridParams g = make_grid_params(c);
    for (int attempt = 0; attempt < 6; ++attempt) {
        MDK_TRY(nlist_reserve_pools(c));
        MDK_TRY(nlist_enqueue(c, false));
        int h_cnt[16], h_flags[4];
        MDK_CUDA(c, cudaMemcpyAsync(h_cnt, c->counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, c->stream));
        MDK_CUDA(c, cudaMemcpyAsync(h_flags, c->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (!h_flags[2]) {
            c->stat_units = h_cnt[0]; c->stat_chunks = h_cnt[4]; c->stat_masks = h_cnt[2];  // [4] = filled chunks
            // hoisted-image kernel where the box leaves room for it (a static property of box and cutoff: blocks
            // that are too wide for it are flagged one by one by the builder)
            const bool shift_before = c->shift_ok;
            c->shift_ok = true;
            for (int a = 0; a < 3; ++a)
                if (g.wide_lim[a] < 1.5f) c->shift_ok = false;
            if (c->shift_ok != shift_before) ++c->graph_epoch;   // template argument of the captured k_pair launch
            c->nlist_valid = true;
            ++c->n_rebuilds;
            return MDK_OK;
        }
        // pool overflow: grow to what the builder asked for (plus slack) and redo
        c->cap_chunks = (size_t)(h_cnt[1] * 1.25) + 1024;
        c->cap_units = (size_t)(h_cnt[0] * 1.25) + 1024;
        c->cap_masks = (size_t)(h_cnt[2] * 1.25) + 1024;
    }
    return fail(c, MDK_ERR_OOM, "tile list pools kept overflowing");
}

int nlist_ensure(mdk_ctx *c) {
    if (c->nlist_valid) {
        int h_flags[4];
        MDK_CUDA(c, cudaMemcpyAsync(h_flags, c->flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
        MDK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (!h_flags[1]) return MDK_OK;
    }
    return nlist_rebuild(c);
}

}  // namespace mdk
