// mdk_comm.cu — multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference has no multi-GPU path at all.  Round-1 scheme:
// positions replicated, i-blocks of the tile list sharded over ranks (each rank builds and
// evaluates only its own blocks' work units), PME on the last rank, bonded / excluded-pair
// terms dealt evenly in contiguous ranges, then ONE ncclAllReduce(sum) of the int64 fixed-point force accumulator per
// force evaluation.  Integer addition commutes, so every rank ends up with bit-identical forces
// and integrates all atoms redundantly; the N-GPU trajectory equals the 1-GPU one.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a torch process that is the
// copy torch already loaded (one NCCL per process), otherwise the system library.
#include <dlfcn.h>

#include "mdk_common.cuh"

namespace mdk {

typedef struct { char internal[128]; } NcclUniqueId;
typedef int (*fn_get_unique_id)(NcclUniqueId *);
typedef int (*fn_comm_init_rank)(void **, int, NcclUniqueId, int);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_comm_destroy)(void *);
typedef const char *(*fn_error_string)(int);

struct NcclApi {
    void *handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_error_string error_string = nullptr;
};
static NcclApi g_nccl;
constexpr int NCCL_INT64 = 4, NCCL_SUM = 0;

static const char *load_nccl() {
    if (g_nccl.handle) return nullptr;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return dlerror();
    g_nccl.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy)
        return "libnccl.so.2 lacks the expected symbols";
    g_nccl.handle = h;
    return nullptr;
}

int comm_allreduce_forces(mdk_ctx *c) {
    if (c->nranks <= 1) return MDK_OK;
    PhaseTimer pt(c, PH_COMM);
    int rc = g_nccl.all_reduce(c->f_acc.p, c->f_acc.p, (size_t)c->n_pad * 3, NCCL_INT64, NCCL_SUM, c->nccl_comm, c->stream);
    if (rc != 0) return fail(c, MDK_ERR_NCCL, "ncclAllReduce(forces): %s", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return MDK_OK;
}

int comm_allreduce_energies(mdk_ctx *c) {
    if (c->nranks <= 1) return MDK_OK;
    int rc = g_nccl.all_reduce(c->e_acc.p, c->e_acc.p, (size_t)MDK_NUM_ENERGIES, NCCL_INT64, NCCL_SUM, c->nccl_comm, c->stream);
    if (rc != 0) return fail(c, MDK_ERR_NCCL, "ncclAllReduce(energies): %s", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return MDK_OK;
}

void comm_destroy(mdk_ctx *c) {
    if (c->nccl_comm && g_nccl.comm_destroy) g_nccl.comm_destroy(c->nccl_comm);
    c->nccl_comm = nullptr;
    c->nranks = 1; c->rank = 0;
    ++c->graph_epoch;
}

}  // namespace mdk

using namespace mdk;

extern "C" {

int mdk_comm_unique_id(void *out128) {
    if (!out128) return MDK_ERR_BAD_ARG;
    const char *err = load_nccl();
    if (err) return fail(nullptr, MDK_ERR_NCCL, "cannot load NCCL: %s", err);
    NcclUniqueId id;
    int rc = g_nccl.get_unique_id(&id);
    if (rc != 0) return fail(nullptr, MDK_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
    memcpy(out128, id.internal, 128);
    return MDK_OK;
}

int mdk_comm_init(mdk_ctx *c, int rank, int nranks, const void *unique_id128) {
    if (!c) return MDK_ERR_BAD_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !unique_id128))
        return fail(c, MDK_ERR_BAD_ARG, "mdk_comm_init(rank=%d, nranks=%d)", rank, nranks);
    cudaSetDevice(c->device);
    comm_destroy(c);
    if (nranks == 1) return MDK_OK;
    const char *err = load_nccl();
    if (err) return fail(c, MDK_ERR_NCCL, "cannot load NCCL: %s", err);
    NcclUniqueId id;
    memcpy(id.internal, unique_id128, 128);
    void *comm = nullptr;
    int rc = g_nccl.comm_init_rank(&comm, nranks, id, rank);
    if (rc != 0) return fail(c, MDK_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    c->nccl_comm = comm;
    c->rank = rank; c->nranks = nranks;
    c->nlist_valid = false;
    ++c->graph_epoch;   // rank / nranks are baked into captured launches (term ranges, PME role)
    return MDK_OK;
}

int mdk_set_shard(mdk_ctx *c, int lo, int hi, int modulus) {
    if (!c) return MDK_ERR_BAD_ARG;
    if (modulus < 1 || lo < 0 || hi > modulus || lo > hi) return fail(c, MDK_ERR_BAD_ARG, "mdk_set_shard(%d, %d, %d)", lo, hi, modulus);
    c->shard_lo = lo; c->shard_hi = hi; c->shard_mod = modulus;
    c->nlist_valid = false;
    // the shard range is a kernel argument of the list builder captured inside the upkeep graph: a graph
    // captured before this call would keep building EVERY block's units on this rank (forces counted
    // nranks times after the first in-graph rebuild)
    ++c->graph_epoch;
    return MDK_OK;
}

}  // extern "C"
