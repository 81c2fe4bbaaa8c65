// api_comm.cu -- C-ABI of the multi-GPU exchange (gsb_comm_*): NCCL over NVLink 5 / NVSwitch.
//
// The reference has no distributed code; its `tohnsw` is one process that sketches
// every file and then inserts all signatures into one graph (src/dna/dnasketch.rs:421-435).  Here
// the same job runs as one process per GPU: genomes shard by rank, the finished signatures are
// exchanged with ONE all-gather straight out of the sketcher's output buffer (the rank's slice of
// the replicated signature matrix), and HNSW insertion is sharded by point inside every wave
// (api_index.cu: gsb_index_insert_batch_sharded).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on
// it, a single-GPU consumer never loads it, and inside a torch process the copy torch already
// mapped is the one that is used.  The unique id travels between the processes by whatever the host
// side has (a pipe in gsearch_b200/cli.py, torch.distributed in bench.py, MPI, a file ...).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <new>

#include "api_common.h"

using namespace gsb;

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.lib) return GSB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        set_error("cannot load libnccl.so.2 (%s): multi-GPU entry points need NCCL", dlerror());
        return GSB_ERR_UNSUPPORTED;
    }
    NcclApi a;
    a.lib = lib;
#define GSB_SYM(field, name)                                                 \
    *(void **)(&a.field) = dlsym(lib, name);                                 \
    if (!a.field) {                                                          \
        set_error("libnccl.so.2 lacks %s", name);                            \
        return GSB_ERR_UNSUPPORTED;                                          \
    }
    GSB_SYM(GetUniqueId, "ncclGetUniqueId")
    GSB_SYM(CommInitRank, "ncclCommInitRank")
    GSB_SYM(CommDestroy, "ncclCommDestroy")
    GSB_SYM(AllGather, "ncclAllGather")
    GSB_SYM(Broadcast, "ncclBroadcast")
    GSB_SYM(GroupStart, "ncclGroupStart")
    GSB_SYM(GroupEnd, "ncclGroupEnd")
    GSB_SYM(GetErrorString, "ncclGetErrorString")
    GSB_SYM(GetVersion, "ncclGetVersion")
#undef GSB_SYM
    g_nccl = a;
    return GSB_OK;
}

}  // namespace

struct gsb_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;  // used when the caller passes no stream
};

#define GSB_NCCL_TRY(expr)                                                                          \
    do {                                                                                            \
        ncclResult_t r__ = (expr);                                                                  \
        if (r__ != ncclSuccess) {                                                                   \
            set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
            return GSB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

extern "C" int gsb_comm_unique_id(uint8_t *id_out /* GSB_COMM_ID_BYTES */) {
    if (!id_out) {
        set_error("gsb_comm_unique_id: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == GSB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    GSB_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return GSB_OK;
}

extern "C" int gsb_comm_create(const uint8_t *id, int nranks, int rank, int device, gsb_comm **out) {
    if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("gsb_comm_create: bad argument (nranks %d, rank %d)", nranks, rank);
        return GSB_ERR_INVALID_ARG;
    }
    int rc = check_device(device);
    if (rc) return rc;
    if ((rc = load_nccl())) return rc;
    gsb_comm *c = new (std::nothrow) gsb_comm();
    if (!c) return GSB_ERR_OOM;
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    GSB_CUDA_TRY(cudaSetDevice(device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, uid, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return GSB_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_nccl.CommDestroy(c->comm);
        delete c;
        set_error("cudaStreamCreate failed");
        return GSB_ERR_CUDA;
    }
    *out = c;
    return GSB_OK;
}

extern "C" void gsb_comm_destroy(gsb_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
}

extern "C" int gsb_comm_rank(const gsb_comm *c) { return c ? c->rank : -1; }
extern "C" int gsb_comm_size(const gsb_comm *c) { return c ? c->nranks : 0; }

// every rank contributes `bytes_per_rank` bytes; d_recv (nranks * bytes_per_rank) ends up identical
// on all ranks, rank r's block at offset r * bytes_per_rank.  In place when d_send == d_recv + rank
// * bytes_per_rank (the sketcher wrote straight into its slice of the replicated matrix).
extern "C" int gsb_comm_all_gather(gsb_comm *c, const void *d_send, void *d_recv, uint64_t bytes_per_rank,
                                   void *stream) {
    if (!c || (bytes_per_rank && (!d_send || !d_recv))) {
        set_error("gsb_comm_all_gather: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (bytes_per_rank == 0) return GSB_OK;
    GSB_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    GSB_NCCL_TRY(g_nccl.AllGather(d_send, d_recv, (size_t)bytes_per_rank, ncclUint8, c->comm, st));
    if (!stream) GSB_CUDA_TRY(cudaStreamSynchronize(st));
    return GSB_OK;
}

// all-gather of row-sharded results into GLOBAL unit order: unit i lives on rank i mod nranks as its
// local row i / nranks (how genomes and queries shard); d_out row i = unit i on every rank.
extern "C" int gsb_comm_all_gather_rows(gsb_comm *c, const void *d_local, uint64_t rows_per_rank, uint64_t row_bytes,
                                        uint64_t n_total, void *d_tmp, void *d_out, void *stream) {
    if (!c || !d_local || !d_tmp || !d_out || row_bytes == 0) {
        set_error("gsb_comm_all_gather_rows: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    const uint64_t world = (uint64_t)c->nranks;
    if (rows_per_rank * world < n_total) {
        set_error("gsb_comm_all_gather_rows: %llu rows per rank cannot hold %llu units on %llu ranks",
                  (unsigned long long)rows_per_rank, (unsigned long long)n_total, (unsigned long long)world);
        return GSB_ERR_INVALID_ARG;
    }
    GSB_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    GSB_NCCL_TRY(g_nccl.AllGather(d_local, d_tmp, (size_t)(rows_per_rank * row_bytes), ncclUint8, c->comm, st));
    for (uint64_t r = 0; r < world; r++) {
        const uint64_t mine = n_total > r ? (n_total - r + world - 1) / world : 0;  // units of rank r
        if (!mine) continue;
        GSB_CUDA_TRY(cudaMemcpy2DAsync((uint8_t *)d_out + r * row_bytes, world * row_bytes,
                                       (const uint8_t *)d_tmp + r * rows_per_rank * row_bytes, row_bytes, row_bytes,
                                       mine, cudaMemcpyDeviceToDevice, st));
    }
    if (!stream) GSB_CUDA_TRY(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_comm_broadcast(gsb_comm *c, void *d_buf, uint64_t bytes, int root, void *stream) {
    if (!c || (bytes && !d_buf) || root < 0 || root >= c->nranks) {
        set_error("gsb_comm_broadcast: bad argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (bytes == 0) return GSB_OK;
    GSB_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    GSB_NCCL_TRY(g_nccl.Broadcast(d_buf, d_buf, (size_t)bytes, ncclUint8, root, c->comm, st));
    if (!stream) GSB_CUDA_TRY(cudaStreamSynchronize(st));
    return GSB_OK;
}

namespace gsb {
// used by api_index.cu: three in-place all-gathers of one wave's selections as ONE NCCL group
int comm_all_gather3(gsb_comm *c, void *a, size_t a_bytes, void *b, size_t b_bytes, void *d, size_t d_bytes,
                     cudaStream_t st) {
    GSB_NCCL_TRY(g_nccl.GroupStart());
    ncclResult_t r1 = g_nccl.AllGather((const uint8_t *)a + (size_t)c->rank * a_bytes, a, a_bytes, ncclUint8, c->comm, st);
    ncclResult_t r2 = g_nccl.AllGather((const uint8_t *)b + (size_t)c->rank * b_bytes, b, b_bytes, ncclUint8, c->comm, st);
    ncclResult_t r3 = g_nccl.AllGather((const uint8_t *)d + (size_t)c->rank * d_bytes, d, d_bytes, ncclUint8, c->comm, st);
    GSB_NCCL_TRY(g_nccl.GroupEnd());
    if (r1 != ncclSuccess || r2 != ncclSuccess || r3 != ncclSuccess) {
        set_error("ncclAllGather failed: %s", g_nccl.GetErrorString(r1 != ncclSuccess ? r1 : (r2 != ncclSuccess ? r2 : r3)));
        return GSB_ERR_CUDA;
    }
    return GSB_OK;
}
int comm_rank(const gsb_comm *c) { return c->rank; }
int comm_size(const gsb_comm *c) { return c->nranks; }
int comm_device(const gsb_comm *c) { return c->device; }
}  // namespace gsb
