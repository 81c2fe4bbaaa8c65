// api_index.cu -- C-ABI of the index (gsb_index_*): device-resident HNSW graph + search.
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "api_common.h"
#include "hnsw_search.cuh"

using namespace gsb;

struct gsb_index {
    gsb_index_params p;
    int device = 0;
    uint32_t elem = 4;
    uint64_t n = 0;
    uint32_t entry = 0;
    uint32_t max_list = 0;
    DevBuf d_sigs, d_ids, d_levels, d_ranks, d_list_base, d_nbr_off, d_nbr_idx;
    DevBuf d_queries, d_out, d_counts, d_neval, d_ws, d_counter;
    cudaStream_t stream = nullptr;
};

static uint32_t elem_of(uint32_t sig_type) {
    switch (sig_type) {
    case GSB_SIG_U64: return 8;
    case GSB_SIG_U16: return 2;
    case GSB_SIG_U32:
    case GSB_SIG_F32: return 4;
    default: return 0;
    }
}

extern "C" int gsb_index_create(const gsb_index_params *params, int device, gsb_index **out) {
    if (!params || !out) {
        set_error("gsb_index_create: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (params->max_nb_connection < 1 || params->max_nb_connection > 255) {
        set_error("max_nb_connection %u out of range 1..255 (src/bin/gsearch.rs:266-268)",
                  params->max_nb_connection);
        return GSB_ERR_INVALID_ARG;
    }
    if (!elem_of(params->sig_type) || params->sketch_size == 0) {
        set_error("bad sig_type / sketch_size");
        return GSB_ERR_INVALID_ARG;
    }
    if (params->max_layer < 1 || params->max_layer > 16) {
        set_error("max_layer %u out of range 1..16", params->max_layer);
        return GSB_ERR_INVALID_ARG;
    }
    int rc = check_device(device);
    if (rc) return rc;
    gsb_index *idx = new (std::nothrow) gsb_index();
    if (!idx) return GSB_ERR_OOM;
    idx->p = *params;
    idx->device = device;
    idx->elem = elem_of(params->sig_type);
    cudaSetDevice(device);
    if (cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        delete idx;
        return GSB_ERR_CUDA;
    }
    *out = idx;
    return GSB_OK;
}

extern "C" void gsb_index_destroy(gsb_index *idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&idx->d_sigs, &idx->d_ids, &idx->d_levels, &idx->d_ranks, &idx->d_list_base,
                      &idx->d_nbr_off, &idx->d_nbr_idx, &idx->d_queries, &idx->d_out, &idx->d_counts,
                      &idx->d_neval, &idx->d_ws, &idx->d_counter};
    for (DevBuf *b : bufs) b->release();
    if (idx->stream) cudaStreamDestroy(idx->stream);
    delete idx;
}

extern "C" uint64_t gsb_index_nb_point(const gsb_index *idx) { return idx ? idx->n : 0; }

extern "C" int gsb_index_load_graph(gsb_index *idx, const void *sigs, const uint64_t *ids, uint64_t n,
                                    const uint8_t *levels, const uint32_t *ranks, const uint64_t *nbr_offsets,
                                    const uint32_t *nbr_index, uint64_t entry_point) {
    if (!idx || (n && (!sigs || !ids || !levels || !ranks || !nbr_offsets || !nbr_index))) {
        set_error("gsb_index_load_graph: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (n > idx->p.capacity || n >= 0xFFFFFFFFull) {
        set_error("%llu points exceed the index capacity %llu", (unsigned long long)n,
                  (unsigned long long)idx->p.capacity);
        return GSB_ERR_CAPACITY;
    }
    GSB_CUDA_TRY(cudaSetDevice(idx->device));
    std::vector<uint64_t> list_base(n + 1);
    uint64_t tl = 0;
    for (uint64_t p = 0; p < n; p++) {
        if (levels[p] >= idx->p.max_layer) {
            set_error("point %llu has level %u >= max_layer %u", (unsigned long long)p, levels[p],
                      idx->p.max_layer);
            return GSB_ERR_INVALID_ARG;
        }
        list_base[p] = tl;
        tl += (uint64_t)levels[p] + 1;
    }
    list_base[n] = tl;
    uint32_t max_list = 0;
    for (uint64_t l = 0; l < tl; l++) {
        if (nbr_offsets[l + 1] < nbr_offsets[l]) {
            set_error("nbr_offsets must be non-decreasing");
            return GSB_ERR_INVALID_ARG;
        }
        max_list = std::max<uint32_t>(max_list, (uint32_t)(nbr_offsets[l + 1] - nbr_offsets[l]));
    }
    const uint64_t tn = n ? nbr_offsets[tl] : 0;
    if (max_list > (uint32_t)kMaxList) {
        set_error("neighbour list of %u entries exceeds the kernel limit %d", max_list, kMaxList);
        return GSB_ERR_CAPACITY;
    }
    for (uint64_t i = 0; i < tn; i++)
        if (nbr_index[i] >= n) {
            set_error("neighbour index %u out of range", nbr_index[i]);
            return GSB_ERR_INVALID_ARG;
        }
    if (n && entry_point >= n) {
        set_error("entry point out of range");
        return GSB_ERR_INVALID_ARG;
    }
    const size_t row = (size_t)idx->p.sketch_size * idx->elem;
    int rc;
    if ((rc = idx->d_sigs.ensure(n * row + 64))) return rc;
    if ((rc = idx->d_ids.ensure(n * 8 + 8))) return rc;
    if ((rc = idx->d_levels.ensure(n + 8))) return rc;
    if ((rc = idx->d_ranks.ensure(n * 4 + 8))) return rc;
    if ((rc = idx->d_list_base.ensure((n + 1) * 8))) return rc;
    if ((rc = idx->d_nbr_off.ensure((tl + 1) * 8))) return rc;
    if ((rc = idx->d_nbr_idx.ensure(tn * 4 + 8))) return rc;
    if (n) {
        GSB_CUDA_TRY(cudaMemcpy(idx->d_sigs.p, sigs, n * row, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_ids.p, ids, n * 8, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_levels.p, levels, n, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_ranks.p, ranks, n * 4, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_list_base.p, list_base.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_nbr_off.p, nbr_offsets, (tl + 1) * 8, cudaMemcpyHostToDevice));
        if (tn) GSB_CUDA_TRY(cudaMemcpy(idx->d_nbr_idx.p, nbr_index, tn * 4, cudaMemcpyHostToDevice));
    }
    idx->n = n;
    idx->entry = (uint32_t)entry_point;
    idx->max_list = max_list;
    return GSB_OK;
}

template <int ELEM, bool F32>
static int launch_search(gsb_index *idx, uint32_t nq, uint32_t knbn, uint32_t ef, cudaStream_t st) {
    const size_t row = (size_t)idx->p.sketch_size * ELEM;
    const size_t row128 = (row + 127) & ~(size_t)127;
    const size_t ret_bytes = ((size_t)ef + 2) * sizeof(HItem);
    const size_t smem_max = 220 * 1024;
    if (row128 > smem_max) {
        set_error("signature row of %zu bytes does not fit in shared memory", row);
        return GSB_ERR_UNSUPPORTED;
    }
    const int ret_in_smem = row128 + ret_bytes <= smem_max ? 1 : 0;
    const size_t smem = row128 + (ret_in_smem ? ret_bytes : 0);
    GSB_CUDA_TRY(cudaFuncSetAttribute(k7_hnsw_search<ELEM, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_max));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, idx->device);
    const uint32_t nctas = std::min<uint32_t>(nq, (uint32_t)nsm);
    const size_t n = idx->n;
    size_t stride = (n + 1) * sizeof(HItem) + ((n + 15) & ~(size_t)15) + (ret_in_smem ? 0 : ret_bytes);
    stride = (stride + 255) & ~(size_t)255;
    int rc;
    if ((rc = idx->d_ws.ensure(stride * nctas))) return rc;
    if ((rc = idx->d_counter.ensure(256))) return rc;
    GSB_CUDA_TRY(cudaMemsetAsync(idx->d_counter.p, 0, 256, st));
    GraphView g;
    g.sigs = idx->d_sigs.as<uint8_t>();
    g.ids = idx->d_ids.as<uint64_t>();
    g.levels = idx->d_levels.as<uint8_t>();
    g.ranks = idx->d_ranks.as<uint32_t>();
    g.list_base = idx->d_list_base.as<uint64_t>();
    g.nbr_off = idx->d_nbr_off.as<uint64_t>();
    g.nbr_idx = idx->d_nbr_idx.as<uint32_t>();
    g.n = (uint32_t)idx->n;
    g.entry = idx->entry;
    g.S = idx->p.sketch_size;
    SearchOut so;
    so.out = idx->d_out.as<gsb_neighbour>();
    so.counts = idx->d_counts.as<uint32_t>();
    so.nb_eval = idx->d_neval.as<unsigned long long>();
    k7_hnsw_search<ELEM, F32><<<nctas, kSearchThreads, smem, st>>>(
        g, idx->d_queries.as<uint8_t>(), nq, knbn, ef, ret_in_smem, idx->d_ws.as<uint8_t>(), stride, so,
        idx->d_counter.as<uint32_t>());
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB_OK;
}

extern "C" int gsb_index_search_batch(gsb_index *idx, const void *queries, uint32_t nq, uint32_t knbn,
                                      uint32_t ef, gsb_neighbour *out, uint32_t *counts_out,
                                      uint64_t *nb_eval_out) {
    if (!idx || (nq && (!queries || !out || !counts_out))) {
        set_error("gsb_index_search_batch: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (nq == 0) return GSB_OK;
    if (knbn == 0) {
        set_error("knbn must be > 0");
        return GSB_ERR_INVALID_ARG;
    }
    if (idx->n == 0) {
        for (uint32_t i = 0; i < nq; i++) counts_out[i] = 0;
        if (nb_eval_out) memset(nb_eval_out, 0, (size_t)nq * 8);
        return GSB_OK;
    }
    GSB_CUDA_TRY(cudaSetDevice(idx->device));
    const uint32_t efs = std::max(ef, knbn);  // hnsw_rs: ef = max(ef_arg, knbn)
    const size_t row = (size_t)idx->p.sketch_size * idx->elem;
    int rc;
    if ((rc = idx->d_queries.ensure((size_t)nq * row + 64))) return rc;
    if ((rc = idx->d_out.ensure((size_t)nq * knbn * sizeof(gsb_neighbour)))) return rc;
    if ((rc = idx->d_counts.ensure((size_t)nq * 4))) return rc;
    if ((rc = idx->d_neval.ensure((size_t)nq * 8))) return rc;
    cudaStream_t st = idx->stream;
    GSB_CUDA_TRY(cudaMemsetAsync(idx->d_out.p, 0, (size_t)nq * knbn * sizeof(gsb_neighbour), st));
    GSB_CUDA_TRY(cudaMemcpyAsync(idx->d_queries.p, queries, (size_t)nq * row, cudaMemcpyHostToDevice, st));
    switch (idx->p.sig_type) {
    case GSB_SIG_U64: rc = launch_search<8, false>(idx, nq, knbn, efs, st); break;
    case GSB_SIG_U32: rc = launch_search<4, false>(idx, nq, knbn, efs, st); break;
    case GSB_SIG_F32: rc = launch_search<4, true>(idx, nq, knbn, efs, st); break;
    default: rc = launch_search<2, false>(idx, nq, knbn, efs, st); break;
    }
    if (rc) return rc;
    GSB_CUDA_TRY(cudaMemcpyAsync(out, idx->d_out.p, (size_t)nq * knbn * sizeof(gsb_neighbour),
                                 cudaMemcpyDeviceToHost, st));
    GSB_CUDA_TRY(cudaMemcpyAsync(counts_out, idx->d_counts.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (nb_eval_out)
        GSB_CUDA_TRY(cudaMemcpyAsync(nb_eval_out, idx->d_neval.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
    GSB_CUDA_TRY(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_index_insert_batch(gsb_index *idx, const void *sigs, const uint64_t *ids, uint64_t n) {
    (void)idx;
    (void)sigs;
    (void)ids;
    (void)n;
    set_error("gsb_index_insert_batch: on-device HNSW construction is not built yet (SURVEY 8f f2)");
    return GSB_ERR_UNSUPPORTED;
}

extern "C" int gsb_index_dump(const gsb_index *idx, const char *dir, const char *basename) {
    (void)idx;
    (void)dir;
    (void)basename;
    set_error("gsb_index_dump: hnswio on-disk format is not built yet (SURVEY 8f f1)");
    return GSB_ERR_UNSUPPORTED;
}

extern "C" int gsb_index_load(gsb_index *idx, const char *dir, const char *basename) {
    (void)idx;
    (void)dir;
    (void)basename;
    set_error("gsb_index_load: hnswio on-disk format is not built yet (SURVEY 8f f1)");
    return GSB_ERR_UNSUPPORTED;
}
