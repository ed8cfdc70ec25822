// Multi-GPU merge of the per-road accumulators: the path shards by tile, a road whose tiles live on several GPUs
// has a row in a boundary table with the same layout on every rank, and ONE grouped NCCL all-reduce(SUM) over NVLink completes
// those rows (uint32 counters: exact, order-independent).  The reference is single-process
// (scripts/statistical_analysis/statistical_analysis.py:187-193 concatenates a road's pixels over all its tiles before the
// groupby); this is what replaces that concatenation when the tiles are spread over devices.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): inside a PyTorch process that resolves to the NCCL torch already
// loaded, in a plain C/C++ host to the system library.  Only the handful of entry points below are used; their C ABI
// (ncclUniqueId = 128 opaque bytes passed by value, ncclComm_t = opaque pointer) is stable across NCCL 2.x.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include "rs_internal.h"

namespace {

struct UniqueId { char internal[RS_COMM_ID_BYTES]; };
typedef void *Comm;
typedef int (*GetUniqueIdFn)(UniqueId *);
typedef int (*CommInitRankFn)(Comm *, int, UniqueId, int);
typedef int (*CommDestroyFn)(Comm);
typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, Comm, cudaStream_t);
typedef int (*GroupFn)(void);

struct Nccl {
    void *handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    GroupFn group_start = nullptr, group_end = nullptr;
    bool ok = false;
};

enum { NCCL_UINT32 = 3, NCCL_SUM = 0 };

Nccl &nccl()
{
    static Nccl n;
    if (n.handle || n.ok) return n;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) break;
    }
    if (!n.handle) return n;
    n.get_unique_id = (GetUniqueIdFn)dlsym(n.handle, "ncclGetUniqueId");
    n.comm_init_rank = (CommInitRankFn)dlsym(n.handle, "ncclCommInitRank");
    n.comm_destroy = (CommDestroyFn)dlsym(n.handle, "ncclCommDestroy");
    n.all_reduce = (AllReduceFn)dlsym(n.handle, "ncclAllReduce");
    n.group_start = (GroupFn)dlsym(n.handle, "ncclGroupStart");
    n.group_end = (GroupFn)dlsym(n.handle, "ncclGroupEnd");
    n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_reduce && n.group_start && n.group_end;
    return n;
}

}  // namespace

extern "C" {

int rs_comm_unique_id(void *id_out)
{
    if (!id_out) return RS_ERR_INVALID_ARG;
    Nccl &n = nccl();
    if (!n.ok) return RS_ERR_NO_NCCL;
    UniqueId id;
    if (n.get_unique_id(&id) != 0) return RS_ERR_NCCL;
    memcpy(id_out, id.internal, RS_COMM_ID_BYTES);
    return RS_OK;
}

int rs_comm_init(rs_ctx *ctx, const void *id, int32_t world, int32_t rank)
{
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return RS_ERR_INVALID_ARG;
    if (ctx->comm) return RS_ERR_INVALID_ARG;                 // one communicator per context
    Nccl &n = nccl();
    if (!n.ok) return RS_ERR_NO_NCCL;
    RS_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    UniqueId uid;
    memcpy(uid.internal, id, RS_COMM_ID_BYTES);
    Comm c = nullptr;
    if (n.comm_init_rank(&c, world, uid, rank) != 0) return RS_ERR_NCCL;
    ctx->comm = c;
    ctx->comm_world = world;
    ctx->comm_rank = rank;
    return RS_OK;
}

int rs_comm_destroy(rs_ctx *ctx)
{
    if (!ctx) return RS_ERR_INVALID_ARG;
    if (!ctx->comm) return RS_OK;
    Nccl &n = nccl();
    if (n.ok) {
        cudaSetDevice(ctx->device);
        n.comm_destroy((Comm)ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->comm_world = 0;
    return RS_OK;
}

int rs_comm_world(rs_ctx *ctx) { return ctx && ctx->comm ? ctx->comm_world : 1; }

int rs_allreduce_accumulators_dev(rs_ctx *ctx, uint32_t *hist_rows, int64_t n_hist, uint32_t *n_allzero_rows, uint32_t *min_zero_rows,
                                  int64_t n_rows, void *stream)
{
    if (!ctx || n_hist < 0 || n_rows < 0) return RS_ERR_INVALID_ARG;
    if ((n_hist > 0 && !hist_rows) || (n_rows > 0 && !n_allzero_rows)) return RS_ERR_INVALID_ARG;
    if (!ctx->comm) return ctx->comm_world <= 1 ? RS_OK : RS_ERR_INVALID_ARG;      // single GPU: nothing to merge
    if (n_hist == 0 && n_rows == 0) return RS_OK;
    Nccl &n = nccl();
    if (!n.ok) return RS_ERR_NO_NCCL;
    RS_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Comm c = (Comm)ctx->comm;
    // one grouped launch: the tables stay where the kernel wrote them (no packing copy), NCCL fuses the operations
    int rc = n.group_start();
    if (rc == 0 && n_hist > 0) rc = n.all_reduce(hist_rows, hist_rows, (size_t)n_hist, NCCL_UINT32, NCCL_SUM, c, st);
    if (rc == 0 && n_rows > 0) rc = n.all_reduce(n_allzero_rows, n_allzero_rows, (size_t)n_rows, NCCL_UINT32, NCCL_SUM, c, st);
    if (rc == 0 && n_rows > 0 && min_zero_rows) rc = n.all_reduce(min_zero_rows, min_zero_rows, (size_t)n_rows, NCCL_UINT32, NCCL_SUM, c, st);
    const int rc_end = n.group_end();
    if (rc != 0 || rc_end != 0) return RS_ERR_NCCL;
    return RS_OK;
}

}  // extern "C"
