// mdk_comm.cu — multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference has no multi-GPU path at all.  The scheme is spatial domain decomposition with
// halo exchange (mdk_dd.cu); this file holds what it needs from NCCL — grouped point-to-point
// transfers (ncclSend / ncclRecv inside ncclGroupStart / End), a small all-gather and the all-reduce of the
// fixed-point energies — plus an in-process stand-in for the same transfers between several contexts that
// share ONE GPU (the "local" backend: device-to-device copies), which lets the whole decomposition logic be
// tested on a single-GPU box.
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
typedef int (*fn_send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_all_gather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_error_string)(int);

struct NcclApi {
    void *handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_error_string error_string = nullptr;
    fn_send send = nullptr;
    fn_recv recv = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_all_gather all_gather = nullptr;
};
static NcclApi g_nccl;
constexpr int NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_SUM = 0;

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
    g_nccl.send = (fn_send)dlsym(h, "ncclSend");
    g_nccl.recv = (fn_recv)dlsym(h, "ncclRecv");
    g_nccl.group_start = (fn_group)dlsym(h, "ncclGroupStart");
    g_nccl.group_end = (fn_group)dlsym(h, "ncclGroupEnd");
    g_nccl.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy || !g_nccl.send ||
        !g_nccl.recv || !g_nccl.group_start || !g_nccl.group_end || !g_nccl.all_gather)
        return "libnccl.so.2 lacks the expected symbols";
    g_nccl.handle = h;
    return nullptr;
}

// One grouped set of point-to-point transfers on the context stream.  Every transfer names a peer and byte
// ranges of the send / receive buffers; transfers between the same two ranks match in list order.
int comm_exchange(mdk_ctx *c, const void *sbuf, void *rbuf, const Xfer *x, int nx) {
    if (c->nranks <= 1 || nx == 0) return MDK_OK;
    if (!c->nccl_comm) return fail(c, MDK_ERR_NCCL, "comm_exchange without a communicator");
    PhaseTimer pt(c, PH_COMM);
    int rc = g_nccl.group_start();
    for (int k = 0; k < nx && rc == 0; ++k) {
        if (x[k].sbytes) rc = g_nccl.send(x[k].sptr ? static_cast<const char *>(x[k].sptr) : static_cast<const char *>(sbuf) + x[k].soff, x[k].sbytes, NCCL_INT8, x[k].peer, c->nccl_comm, c->stream);
        if (rc == 0 && x[k].rbytes) rc = g_nccl.recv(x[k].rptr ? static_cast<char *>(x[k].rptr) : static_cast<char *>(rbuf) + x[k].roff, x[k].rbytes, NCCL_INT8, x[k].peer, c->nccl_comm, c->stream);
    }
    int rc2 = g_nccl.group_end();
    if (rc == 0) rc = rc2;
    if (rc != 0) return fail(c, MDK_ERR_NCCL, "ncclSend/ncclRecv group: %s", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return MDK_OK;
}

int comm_allgather_i32(mdk_ctx *c, const int *mine, int *all, int count) {
    if (c->nranks <= 1) return MDK_OK;
    int rc = g_nccl.all_gather(mine, all, (size_t)count, NCCL_INT32, c->nccl_comm, c->stream);
    if (rc != 0) return fail(c, MDK_ERR_NCCL, "ncclAllGather: %s", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return MDK_OK;
}

int comm_allreduce_energies(mdk_ctx *c) {
    if (c->nranks <= 1 || !c->nccl_comm) return MDK_OK;   // local-backend groups add their energies on the host (mdk_dd.cu)
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
    ++c->graph_epoch;
    return MDK_OK;
}

}  // extern "C"
