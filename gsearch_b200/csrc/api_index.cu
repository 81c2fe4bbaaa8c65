// api_index.cu -- C-ABI of the index (gsb_index_*): device-resident, mutable HNSW graph with
// on-device construction (wave insertion) and search.  Host code here only sizes buffers,
// draws the levels (hnsw_rs LayerGenerator, restated in oracle/hnsw.c) and launches kernels.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "api_common.h"
#include "hnsw_device.cuh"

using namespace gsb;

namespace {

// xoshiro256++ seeded through SplitMix64 (rand_xoshiro seed_from_u64) and the two rand-0.8
// uniform draws the level generator needs -- host twins of common.cuh
struct HostXoshiro {
    uint64_t s[4];
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    void seed(uint64_t z) {
        for (int i = 0; i < 4; i++) {
            z += 0x9e3779b97f4a7c15ULL;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
            x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
            s[i] = x ^ (x >> 31);
        }
    }
    uint64_t next() {
        const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    double uniform_f64() {
        const uint64_t bits = (next() >> 12) | 0x3FF0000000000000ULL;
        double d;
        memcpy(&d, &bits, 8);
        return d - 1.0;
    }
    uint64_t uniform_usize(uint64_t m) {
        const uint64_t zone = UINT64_MAX - ((UINT64_MAX - m + 1) % m);
        for (;;) {
            const uint64_t v = next();
            const unsigned __int128 p = (unsigned __int128)v * (unsigned __int128)m;
            if ((uint64_t)p <= zone) return (uint64_t)(p >> 64);
        }
    }
};

// device buffer that keeps its content when it grows, new space zero-filled
struct GrowBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, cudaStream_t st) {
        if (bytes <= cap) return GSB_OK;
        size_t want = std::max(bytes, cap * 2) + 256;
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return GSB_ERR_OOM;
        }
        // everything on `st` (created non-blocking: the legacy default stream would not order
        // with it), and finished before the old buffer goes
        if (cap) cudaMemcpyAsync(q, p, cap, cudaMemcpyDeviceToDevice, st);
        cudaMemsetAsync((uint8_t *)q + cap, 0, want - cap, st);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("GrowBuf: copy to the grown buffer failed: %s", cudaGetErrorString(e));
            cudaFree(q);
            return GSB_ERR_CUDA;
        }
        if (p) cudaFree(p);
        p = q;
        cap = want;
        return GSB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace

struct gsb_index {
    std::recursive_mutex mu;  // a handle serialises its calls
    gsb_index_params p;
    int device = 0;
    int nsm = 148;
    uint32_t elem = 4;
    uint32_t M = 0;
    uint64_t n = 0;        // points in the graph
    uint64_t nU = 0;       // upper-layer lists in use
    uint32_t entry = 0;
    bool has_dist = true;  // false after load_graph without distances: search only
    uint32_t wave_max = 0;
    HostXoshiro level_rng;
    double level_scale = 1.0;
    std::vector<uint8_t> levels;      // host mirror (entry-point bookkeeping, export)
    std::vector<uint32_t> upper_off;  // host mirror
    uint32_t layer_count[kMaxLayers] = {0};
    GrowBuf d_sigs, d_ids, d_levels, d_ranks, d_nbr0, d_dist0, d_cnt0, d_upper_off, d_nbrU, d_distU, d_cntU,
        d_lock0, d_lockU;
    DevBuf d_queries, d_out, d_counts, d_neval, d_counter, d_sel_n, d_sel_idx, d_sel_d;
    DevBuf d_ws;  // per-CTA workspaces (zeroed when reallocated: the visit stamps live there)
    WsLayout wl{};
    uint32_t ws_ctas = 0, ws_ef = 0;
    uint64_t ws_npts = 0;
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
    if (params->max_nb_connection < 2 || params->max_nb_connection > 255) {
        set_error("max_nb_connection %u out of range 2..255 (src/bin/gsearch.rs:266-268)",
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
    idx->M = params->max_nb_connection;
    idx->level_rng.seed(params->level_seed);
    idx->level_scale = (params->scale_modification > 0 ? params->scale_modification : 1.0) /
                       log((double)params->max_nb_connection);
    cudaSetDevice(device);
    cudaDeviceGetAttribute(&idx->nsm, cudaDevAttrMultiProcessorCount, device);
    idx->wave_max = (uint32_t)idx->nsm * GSB_K8_PER_SM;  // one wave = one point per resident insertion CTA
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
    GrowBuf *gb[] = {&idx->d_sigs, &idx->d_ids, &idx->d_levels, &idx->d_ranks, &idx->d_nbr0, &idx->d_dist0,
                     &idx->d_cnt0, &idx->d_upper_off, &idx->d_nbrU, &idx->d_distU, &idx->d_cntU, &idx->d_lock0,
                     &idx->d_lockU};
    for (GrowBuf *b : gb) b->release();
    DevBuf *bufs[] = {&idx->d_queries, &idx->d_out, &idx->d_counts, &idx->d_neval, &idx->d_counter,
                      &idx->d_sel_n, &idx->d_sel_idx, &idx->d_sel_d, &idx->d_ws};
    for (DevBuf *b : bufs) b->release();
    if (idx->stream) cudaStreamDestroy(idx->stream);
    delete idx;
}

extern "C" uint64_t gsb_index_nb_point(const gsb_index *idx) { return idx ? idx->n : 0; }

extern "C" int gsb_index_set_wave_max(gsb_index *idx, uint32_t wave_max) {
    if (!idx || wave_max < 1 || wave_max > (uint32_t)kMaxWave) {
        set_error("wave_max must be in 1..%d", kMaxWave);
        return GSB_ERR_INVALID_ARG;
    }
    idx->wave_max = wave_max;
    return GSB_OK;
}

namespace {

int ensure_points(gsb_index *idx, uint64_t npts) {
    const size_t row = (size_t)idx->p.sketch_size * idx->elem;
    const size_t M2 = 2 * (size_t)idx->M;
    cudaStream_t st = idx->stream;
    int rc;
    if ((rc = idx->d_sigs.ensure(npts * row + 64, st))) return rc;
    if ((rc = idx->d_ids.ensure(npts * 8 + 8, st))) return rc;
    if ((rc = idx->d_levels.ensure(npts + 8, st))) return rc;
    if ((rc = idx->d_ranks.ensure(npts * 4 + 8, st))) return rc;
    if ((rc = idx->d_upper_off.ensure(npts * 4 + 8, st))) return rc;
    if ((rc = idx->d_nbr0.ensure(npts * M2 * 4 + 8, st))) return rc;
    if ((rc = idx->d_dist0.ensure(npts * M2 * 4 + 8, st))) return rc;
    if ((rc = idx->d_cnt0.ensure(npts * 4 + 8, st))) return rc;
    if ((rc = idx->d_lock0.ensure(npts * 4 + 8, st))) return rc;
    return GSB_OK;
}

int ensure_upper(gsb_index *idx, uint64_t nlists) {
    const size_t M = idx->M;
    cudaStream_t st = idx->stream;
    int rc;
    if ((rc = idx->d_nbrU.ensure(nlists * M * 4 + 8, st))) return rc;
    if ((rc = idx->d_distU.ensure(nlists * M * 4 + 8, st))) return rc;
    if ((rc = idx->d_cntU.ensure(nlists * 4 + 8, st))) return rc;
    if ((rc = idx->d_lockU.ensure(nlists * 4 + 8, st))) return rc;
    return GSB_OK;
}

GraphView graph_view(const gsb_index *idx) {
    GraphView g;
    g.sigs = idx->d_sigs.as<uint8_t>();
    g.ids = idx->d_ids.as<uint64_t>();
    g.levels = idx->d_levels.as<uint8_t>();
    g.ranks = idx->d_ranks.as<uint32_t>();
    g.nbr0 = idx->d_nbr0.as<uint32_t>();
    g.dist0 = idx->d_dist0.as<float>();
    g.cnt0 = idx->d_cnt0.as<uint32_t>();
    g.upper_off = idx->d_upper_off.as<uint32_t>();
    g.nbrU = idx->d_nbrU.as<uint32_t>();
    g.distU = idx->d_distU.as<float>();
    g.cntU = idx->d_cntU.as<uint32_t>();
    g.lock0 = idx->d_lock0.as<uint32_t>();
    g.lockU = idx->d_lockU.as<uint32_t>();
    g.M = idx->M;
    g.n = (uint32_t)idx->n;
    g.entry = idx->entry;
    g.S = idx->p.sketch_size;
    return g;
}

constexpr size_t kSmemMax = 192 * 1024;  // dynamic; the kernels add up to 33 KB of static shared memory

// per-CTA workspace for `nctas` CTAs over `npts` points with a result heap of `ef`; grown
// geometrically (and zeroed: the visit stamps live there) so that a growing index does not
// reallocate at every wave
int ensure_workspace(gsb_index *idx, uint32_t nctas, uint64_t npts, uint32_t ef) {
    if (idx->d_ws.p && npts <= idx->ws_npts && ef <= idx->ws_ef && nctas <= idx->ws_ctas) return GSB_OK;
    const size_t M = idx->M;
    const uint64_t cap_pts = std::max<uint64_t>(std::max<uint64_t>(2 * npts, idx->ws_npts), 1024);
    const uint32_t cap_ef = std::max(ef, idx->ws_ef);
    const uint32_t nc = std::max<uint32_t>(std::max(nctas, idx->ws_ctas), (uint32_t)idx->nsm);
    WsLayout wl;
    size_t off = 0;
    wl.off_ctr = off;
    off += 16;
    wl.off_cand = off;
    off += (std::max<size_t>(cap_pts, (size_t)cap_ef + 4 * M * M) + 4) * sizeof(HItem);
    wl.off_stamp = off;
    off += ((cap_pts + 4) * 4 + 15) & ~(size_t)15;
    wl.off_ret = off;
    off += ((size_t)cap_ef + 2) * sizeof(HItem);
    wl.off_newc = off;
    off += (4 * M * M + 16) * 4;
    wl.stride = (off + 255) & ~(size_t)255;
    cudaStreamSynchronize(idx->stream);
    idx->d_ws.release();
    int rc = idx->d_ws.ensure(wl.stride * nc);
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(idx->d_ws.p, 0, idx->d_ws.cap, idx->stream);
    if (e != cudaSuccess) {
        set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        return GSB_ERR_CUDA;
    }
    idx->wl = wl;
    idx->ws_ctas = nc;
    idx->ws_npts = cap_pts;
    idx->ws_ef = cap_ef;
    return GSB_OK;
}

}  // namespace

extern "C" int gsb_index_load_graph(gsb_index *idx, const void *sigs, const uint64_t *ids, uint64_t n,
                                    const uint8_t *levels, const uint32_t *ranks, const uint64_t *nbr_offsets,
                                    const uint32_t *nbr_index, const float *nbr_dist, uint64_t entry_point) {
    std::unique_lock<std::recursive_mutex> lock_;
    if (idx) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<gsb_index *>(idx)->mu);
    if (!idx || (n && (!sigs || !ids || !levels || !ranks || !nbr_offsets || !nbr_index))) {
        set_error("gsb_index_load_graph: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (n > idx->p.capacity || n >= 0xFFFFFFF0ull) {
        set_error("%llu points exceed the index capacity %llu", (unsigned long long)n,
                  (unsigned long long)idx->p.capacity);
        return GSB_ERR_CAPACITY;
    }
    GSB_CUDA_TRY(cudaSetDevice(idx->device));
    const uint32_t M = idx->M;
    std::vector<uint32_t> upper_off(n, 0);
    uint64_t tl = 0, nU = 0;
    for (uint64_t p = 0; p < n; p++) {
        if (levels[p] >= idx->p.max_layer) {
            set_error("point %llu has level %u >= max_layer %u", (unsigned long long)p, levels[p],
                      idx->p.max_layer);
            return GSB_ERR_INVALID_ARG;
        }
        upper_off[p] = (uint32_t)nU;
        nU += levels[p];
        tl += (uint64_t)levels[p] + 1;
    }
    if (n && entry_point >= n) {
        set_error("entry point out of range");
        return GSB_ERR_INVALID_ARG;
    }
    // CSR image -> fixed-capacity adjacency
    std::vector<uint32_t> nbr0((size_t)n * 2 * M, 0), cnt0(n, 0), nbrU((size_t)nU * M, 0), cntU(nU, 0);
    std::vector<float> dist0((size_t)n * 2 * M, 0.f), distU((size_t)nU * M, 0.f);
    uint64_t li = 0;
    for (uint64_t p = 0; p < n; p++) {
        for (uint32_t l = 0; l <= levels[p]; l++, li++) {
            const uint64_t b = nbr_offsets[li], e = nbr_offsets[li + 1];
            if (e < b || e - b > (l == 0 ? 2u * M : M)) {
                set_error("neighbour list (%llu, layer %u) has %lld entries, limit %u", (unsigned long long)p, l,
                          (long long)(e - b), l == 0 ? 2u * M : M);
                return GSB_ERR_INVALID_ARG;
            }
            uint32_t *di = l == 0 ? &nbr0[(size_t)p * 2 * M] : &nbrU[(size_t)(upper_off[p] + l - 1) * M];
            float *dd = l == 0 ? &dist0[(size_t)p * 2 * M] : &distU[(size_t)(upper_off[p] + l - 1) * M];
            for (uint64_t i = b; i < e; i++) {
                if (nbr_index[i] >= n) {
                    set_error("neighbour index %u out of range", nbr_index[i]);
                    return GSB_ERR_INVALID_ARG;
                }
                di[i - b] = nbr_index[i];
                dd[i - b] = nbr_dist ? nbr_dist[i] : 0.f;
            }
            if (l == 0) cnt0[p] = (uint32_t)(e - b);
            else cntU[upper_off[p] + l - 1] = (uint32_t)(e - b);
        }
    }
    // a fresh graph replaces whatever the index held
    GrowBuf *gb[] = {&idx->d_sigs, &idx->d_ids, &idx->d_levels, &idx->d_ranks, &idx->d_nbr0, &idx->d_dist0,
                     &idx->d_cnt0, &idx->d_upper_off, &idx->d_nbrU, &idx->d_distU, &idx->d_cntU, &idx->d_lock0,
                     &idx->d_lockU};
    cudaStreamSynchronize(idx->stream);
    for (GrowBuf *b : gb) b->release();
    const size_t row = (size_t)idx->p.sketch_size * idx->elem;
    int rc;
    if ((rc = ensure_points(idx, std::max<uint64_t>(n, 1)))) return rc;
    if ((rc = ensure_upper(idx, std::max<uint64_t>(nU, 1)))) return rc;
    if (n) {
        GSB_CUDA_TRY(cudaMemcpy(idx->d_sigs.p, sigs, n * row, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_ids.p, ids, n * 8, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_levels.p, levels, n, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_ranks.p, ranks, n * 4, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_upper_off.p, upper_off.data(), n * 4, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_nbr0.p, nbr0.data(), nbr0.size() * 4, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_dist0.p, dist0.data(), dist0.size() * 4, cudaMemcpyHostToDevice));
        GSB_CUDA_TRY(cudaMemcpy(idx->d_cnt0.p, cnt0.data(), n * 4, cudaMemcpyHostToDevice));
        if (nU) {
            GSB_CUDA_TRY(cudaMemcpy(idx->d_nbrU.p, nbrU.data(), nbrU.size() * 4, cudaMemcpyHostToDevice));
            GSB_CUDA_TRY(cudaMemcpy(idx->d_distU.p, distU.data(), distU.size() * 4, cudaMemcpyHostToDevice));
            GSB_CUDA_TRY(cudaMemcpy(idx->d_cntU.p, cntU.data(), nU * 4, cudaMemcpyHostToDevice));
        }
    }
    idx->n = n;
    idx->nU = nU;
    idx->entry = (uint32_t)entry_point;
    idx->has_dist = nbr_dist != nullptr || n == 0;
    idx->levels.assign(levels, levels + n);
    idx->upper_off = upper_off;
    memset(idx->layer_count, 0, sizeof idx->layer_count);
    for (uint64_t p = 0; p < n; p++) idx->layer_count[levels[p]]++;
    return GSB_OK;
}

struct SearchBufs {
    const uint8_t *queries;
    gsb_neighbour *out;
    uint32_t *counts;
    unsigned long long *nb_eval;
};

template <int ELEM, bool F32>
static int launch_search(gsb_index *idx, const SearchBufs &sb, uint32_t nq, uint32_t knbn, uint32_t ef,
                         cudaStream_t st) {
    const size_t row = (size_t)idx->p.sketch_size * ELEM;
    const size_t ret_bytes = ((size_t)ef + 2) * sizeof(HItem);
    // a row that does not fit in shared memory stays in global memory (S up to 65535 is legal)
    // The query row is NOT staged in shared memory for the search: without it three CTAs fit on an
    // SM, and three independent search chains per SM hide each other's serial phases (heap
    // updates, neighbour gathering) -- measured 12.8 k against 9.6 k queries/s with the row staged
    // and one CTA per SM (GSB_K7_STAGED=1 restores that).  The row is read through L1/L2 instead.
    const int staged = (((row + 127) & ~(size_t)127) <= kSmemMax && getenv("GSB_K7_STAGED")) ? 1 : 0;
    // Default: the candidate rows stream through a TMA ring (hnsw_device.cuh, eval_list_ring); it needs
    // 16-byte rows and pays off from a few KB per row.  The shared memory of a CTA is then budgeted for
    // three CTAs per SM: ring, the top of the result heap (the rest lives in the workspace, like the
    // candidate heap), the visited bitmap if it still fits.
    // The ring pays off when the result heap is too large to leave room for three CTAs' worth of
    // register-load streaming (measured at S = 18000 x u64: ef_search 5000 +15 %, ef_search 1600 -17 %).
    const bool want_ring = getenv("GSB_K7_RING") ? atoi(getenv("GSB_K7_RING")) != 0 : ef + 2 > 2048;
    const int ring = (!staged && want_ring && (row & 15) == 0 && row >= 4096 && (((uintptr_t)sb.queries) & 15) == 0) ? 1 : 0;
    const size_t row128 = staged ? ((row + 127) & ~(size_t)127) : (ring ? (size_t)kRingBytes : 0);
    const size_t bm_bytes = ((idx->n + 31) / 32) * 4;
    uint32_t ret_cs, bm_words;
    size_t smem;
    if (ring) {
        // dynamic bytes of one of three CTAs per SM: (228 KB / 3 - 1 KB reserved) - 19.6 KB static
        constexpr size_t kBudget = 233472 / 3 - 1024 - (20096 - (1024 - kCandSmemK7) * sizeof(HItem)) - 128;
        smem = row128;
        bm_words = (smem + bm_bytes + 256 * sizeof(HItem) <= kBudget && !getenv("GSB_NO_BITMAP")) ? (uint32_t)(bm_bytes / 4) : 0u;
        smem += (size_t)bm_words * 4;
        ret_cs = (uint32_t)std::min<size_t>((size_t)ef + 2, (kBudget - smem) / sizeof(HItem));
        smem += (size_t)ret_cs * sizeof(HItem);
    } else {
        const int ret_in_smem = row128 + ret_bytes <= kSmemMax ? 1 : 0;
        ret_cs = ret_in_smem ? ef + 2 : 0;
        smem = row128 + (ret_in_smem ? ret_bytes : 0);
        // visited bitmap in shared memory when it fits beside the row and the result heap
        bm_words = (smem + bm_bytes <= kSmemMax && !getenv("GSB_NO_BITMAP")) ? (uint32_t)(bm_bytes / 4) : 0u;
        smem += (size_t)bm_words * 4;
    }
    auto kern = ring ? k7_hnsw_search<ELEM, F32, true> : k7_hnsw_search<ELEM, F32, false>;
    GSB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    // (A grid sized so that every CTA gets the same number of queries -- 334 CTAs for 1000 queries --
    // was measured slower than the full 444: the chains are latency bound, more of them in flight win.)
    const uint32_t max_ctas = (uint32_t)idx->nsm * (staged ? 1u : 3u);
    uint32_t nctas = std::min<uint32_t>(max_ctas, nq);
    if (const char *e = getenv("GSB_K7_CTAS")) nctas = std::max(1, std::min<int>(atoi(e), (int)max_ctas));
    int rc;
    if ((rc = ensure_workspace(idx, nctas, idx->n, ef))) return rc;
    if ((rc = idx->d_counter.ensure(256))) return rc;
    GSB_CUDA_TRY(cudaMemsetAsync(idx->d_counter.p, 0, 256, st));
    SearchOut so;
    so.out = sb.out;
    so.counts = sb.counts;
    so.nb_eval = sb.nb_eval;
    kern<<<nctas, kSearchThreads, smem, st>>>(
        graph_view(idx), sb.queries, nq, knbn, ef, ret_cs, bm_words, staged,
        idx->d_ws.as<uint8_t>(), idx->wl, so, idx->d_counter.as<uint32_t>());
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB_OK;
}

// queries / results in host memory (dev = false: copied in and out on the index stream, the call
// returns when the results are in host memory) or in device memory (dev = true: the kernel reads
// and writes the caller's buffers; the call returns after the stream has been synchronised)
static int search_batch_impl(gsb_index *idx, const void *queries, uint32_t nq, uint32_t knbn, uint32_t ef,
                             gsb_neighbour *out, uint32_t *counts_out, uint64_t *nb_eval_out, bool dev) {
    std::unique_lock<std::recursive_mutex> lock_;
    if (idx) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<gsb_index *>(idx)->mu);
    if (!idx || (nq && (!queries || !out || !counts_out))) {
        set_error("gsb_index_search_batch: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (nq == 0) return GSB_OK;
    if (knbn == 0) {
        set_error("knbn must be > 0");
        return GSB_ERR_INVALID_ARG;
    }
    GSB_CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    if (idx->n == 0) {
        if (dev) {
            GSB_CUDA_TRY(cudaMemsetAsync(counts_out, 0, (size_t)nq * 4, st));
            if (nb_eval_out) GSB_CUDA_TRY(cudaMemsetAsync(nb_eval_out, 0, (size_t)nq * 8, st));
            GSB_CUDA_TRY(cudaStreamSynchronize(st));
        } else {
            for (uint32_t i = 0; i < nq; i++) counts_out[i] = 0;
            if (nb_eval_out) memset(nb_eval_out, 0, (size_t)nq * 8);
        }
        return GSB_OK;
    }
    const uint32_t efs = std::max(ef, knbn);  // hnsw_rs: ef = max(ef_arg, knbn)
    const size_t row = (size_t)idx->p.sketch_size * idx->elem;
    int rc;
    if (!dev) {
        if ((rc = idx->d_queries.ensure((size_t)nq * row + 64))) return rc;
        if ((rc = idx->d_out.ensure((size_t)nq * knbn * sizeof(gsb_neighbour)))) return rc;
        if ((rc = idx->d_counts.ensure((size_t)nq * 4))) return rc;
    }
    if (!dev || !nb_eval_out)
        if ((rc = idx->d_neval.ensure((size_t)nq * 8))) return rc;
    SearchBufs sb;
    sb.queries = dev ? (const uint8_t *)queries : idx->d_queries.as<uint8_t>();
    sb.out = dev ? out : idx->d_out.as<gsb_neighbour>();
    sb.counts = dev ? counts_out : idx->d_counts.as<uint32_t>();
    sb.nb_eval = (dev && nb_eval_out) ? (unsigned long long *)nb_eval_out : idx->d_neval.as<unsigned long long>();
    GSB_CUDA_TRY(cudaMemsetAsync(sb.out, 0, (size_t)nq * knbn * sizeof(gsb_neighbour), st));
    if (!dev) GSB_CUDA_TRY(cudaMemcpyAsync(idx->d_queries.p, queries, (size_t)nq * row, cudaMemcpyHostToDevice, st));
    switch (idx->p.sig_type) {
    case GSB_SIG_U64: rc = launch_search<8, false>(idx, sb, nq, knbn, efs, st); break;
    case GSB_SIG_U32: rc = launch_search<4, false>(idx, sb, nq, knbn, efs, st); break;
    case GSB_SIG_F32: rc = launch_search<4, true>(idx, sb, nq, knbn, efs, st); break;
    default: rc = launch_search<2, false>(idx, sb, nq, knbn, efs, st); break;
    }
    if (rc) return rc;
    if (!dev) {
        GSB_CUDA_TRY(cudaMemcpyAsync(out, idx->d_out.p, (size_t)nq * knbn * sizeof(gsb_neighbour),
                                     cudaMemcpyDeviceToHost, st));
        GSB_CUDA_TRY(cudaMemcpyAsync(counts_out, idx->d_counts.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
        if (nb_eval_out)
            GSB_CUDA_TRY(cudaMemcpyAsync(nb_eval_out, idx->d_neval.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
    }
    GSB_CUDA_TRY(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_index_export_signatures(const gsb_index *cidx, void *sigs_out) {
    gsb_index *idx = const_cast<gsb_index *>(cidx);
    if (!idx || (idx->n && !sigs_out)) {
        set_error("gsb_index_export_signatures: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    std::unique_lock<std::recursive_mutex> lock_(idx->mu);
    if (idx->n == 0) return GSB_OK;
    GSB_CUDA_TRY(cudaSetDevice(idx->device));
    const size_t bytes = (size_t)idx->n * idx->p.sketch_size * idx->elem;
    GSB_CUDA_TRY(cudaMemcpyAsync(sigs_out, idx->d_sigs.p, bytes, cudaMemcpyDeviceToHost, idx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    return GSB_OK;
}

extern "C" int gsb_index_search_batch(gsb_index *idx, const void *queries, uint32_t nq, uint32_t knbn,
                                      uint32_t ef, gsb_neighbour *out, uint32_t *counts_out,
                                      uint64_t *nb_eval_out) {
    return search_batch_impl(idx, queries, nq, knbn, ef, out, counts_out, nb_eval_out, false);
}

extern "C" int gsb_index_search_batch_dev(gsb_index *idx, const void *d_queries, uint32_t nq, uint32_t knbn,
                                          uint32_t ef, gsb_neighbour *d_out, uint32_t *d_counts_out,
                                          uint64_t *d_nb_eval_out) {
    return search_batch_impl(idx, d_queries, nq, knbn, ef, d_out, d_counts_out, d_nb_eval_out, true);
}

// ------------------------------------------------------------------------------ construction
namespace {

// hnsw_rs LayerGenerator::generate: floor(-ln(U) * scale), re-drawn uniformly if >= max_layer
uint32_t gen_level(gsb_index *idx) {
    const double xsi = idx->level_rng.uniform_f64();
    const double lv = -log(xsi) * idx->level_scale;
    if (!(lv < (double)idx->p.max_layer)) return (uint32_t)idx->level_rng.uniform_usize(idx->p.max_layer);
    return (uint32_t)floor(lv);
}

uint32_t wave_size(uint64_t nb_point, uint32_t wave_max) {  // == gso_hnsw_wave_size
    uint64_t w = nb_point / 4;
    if (w < 1) w = 1;
    if (w > wave_max) w = wave_max;
    return (uint32_t)w;
}

}  // namespace
namespace gsb {
int comm_all_gather3(gsb_comm *c, void *a, size_t a_bytes, void *b, size_t b_bytes, void *d, size_t d_bytes,
                     cudaStream_t st);
int comm_rank(const gsb_comm *c);
int comm_size(const gsb_comm *c);
int comm_device(const gsb_comm *c);
}  // namespace gsb
namespace {

// One wave.  With a communicator, phase A (K8: search + selection, all the distance evaluations)
// runs only for this rank's slice of the wave; the selections (a few KB per point) are all-gathered
// in place and phase B (K9) is applied by every rank to its own replica: the replicas stay
// bit-identical and the graph is the one a single GPU builds with the same wave size.
template <int ELEM, bool F32>
int launch_wave(gsb_index *idx, uint32_t first, uint32_t W, cudaStream_t st, gsb_comm *comm) {
    const uint32_t ef_c = idx->p.ef_construction;
    const size_t row = (size_t)idx->p.sketch_size * ELEM;
    const size_t ret_bytes = ((size_t)ef_c + 2) * sizeof(HItem);
    // as for the search: rows are read through L1/L2, two CTAs (two insertion chains) per SM
    const int staged = (((row + 127) & ~(size_t)127) <= kSmemMax && getenv("GSB_K8_STAGED")) ? 1 : 0;
    const size_t row128 = staged ? ((row + 127) & ~(size_t)127) : 0;
    const int ret_in_smem = row128 + ret_bytes <= kSmemMax ? 1 : 0;
    size_t smem = row128 + (ret_in_smem ? ret_bytes : 0);
    const size_t bm_bytes = (((size_t)first + W + 31) / 32) * 4;
    const uint32_t bm_words = (smem + bm_bytes <= kSmemMax && !getenv("GSB_NO_BITMAP")) ? (uint32_t)(bm_bytes / 4) : 0u;
    smem += (size_t)bm_words * 4;
    // per device, not per process: set on every launch (cheap) so that an index on another GPU of
    // the same process gets the opt-in too
    GSB_CUDA_TRY(cudaFuncSetAttribute(k8_hnsw_insert_select<ELEM, F32>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    const uint32_t world = comm ? (uint32_t)comm_size(comm) : 1u, rank = comm ? (uint32_t)comm_rank(comm) : 0u;
    const uint32_t per = (W + world - 1) / world;  // points of the wave per rank (the last slices may be short or empty)
    const uint32_t t_begin = std::min(W, rank * per), t_end = std::min(W, t_begin + per);
    const uint32_t nctas = std::max<uint32_t>(1u, std::min<uint32_t>(t_end - t_begin, (uint32_t)idx->nsm * (staged ? 1u : (uint32_t)GSB_K8_PER_SM)));
    int rc;
    if ((rc = ensure_workspace(idx, nctas, (uint64_t)first + W, ef_c))) return rc;
    WaveView wv;
    wv.first = first;
    wv.W = W;
    wv.t_begin = t_begin;
    wv.t_end = t_end;
    wv.entry = idx->entry;
    wv.ef_c = ef_c;
    wv.extend = idx->p.extend_candidates;
    wv.sel_n = idx->d_sel_n.as<uint32_t>();
    wv.sel_idx = idx->d_sel_idx.as<uint32_t>();
    wv.sel_d = idx->d_sel_d.as<float>();
    wv.counter = idx->d_counter.as<uint32_t>();
    GSB_CUDA_TRY(cudaMemsetAsync(idx->d_counter.p, 0, 256, st));
    GraphView g = graph_view(idx);
    g.n = first + W;
    if (t_end > t_begin)
        k8_hnsw_insert_select<ELEM, F32><<<nctas, kInsertThreads, smem, st>>>(g, wv, ret_in_smem, bm_words, staged,
                                                                             idx->d_ws.as<uint8_t>(), idx->wl);
    if (world > 1) {
        const size_t M = idx->M;
        if ((rc = comm_all_gather3(comm, wv.sel_n, (size_t)per * kMaxLayers * 4, wv.sel_idx, (size_t)per * 18 * M * 4,
                                   wv.sel_d, (size_t)per * 18 * M * 4, st)))
            return rc;
    }
    k9_write_own_lists<<<W, 256, 0, st>>>(g, wv);
    k9_reverse_updates<<<W, 256, 0, st>>>(g, wv);
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB_OK;
}

}  // namespace

static int insert_batch_impl(gsb_index *idx, const void *sigs, const uint64_t *ids, uint64_t n, gsb_comm *comm) {
    std::unique_lock<std::recursive_mutex> lock_;
    if (idx) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<gsb_index *>(idx)->mu);
    if (!idx || (n && (!sigs || !ids))) {
        set_error("gsb_index_insert_batch: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (n == 0) return GSB_OK;
    if (idx->n + n > idx->p.capacity || idx->n + n >= 0xFFFFFFF0ull) {
        set_error("%llu points exceed the index capacity %llu", (unsigned long long)(idx->n + n),
                  (unsigned long long)idx->p.capacity);
        return GSB_ERR_CAPACITY;
    }
    if (!idx->has_dist) {
        set_error("this graph was loaded without neighbour distances: it can be searched, not extended");
        return GSB_ERR_UNSUPPORTED;
    }
    if (idx->p.keep_pruned) {
        set_error("keep_pruned = true is not built (the reference calls set_keeping_pruned(false), "
                  "src/dna/dnasketch.rs:160)");
        return GSB_ERR_UNSUPPORTED;
    }
    GSB_CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    const size_t row = (size_t)idx->p.sketch_size * idx->elem;
    const uint64_t n0 = idx->n, n1 = n0 + n;
    // ---- levels, ranks and upper-list slots of the new points (in order: one level draw each)
    std::vector<uint8_t> lv(n);
    std::vector<uint32_t> rk(n), uo(n);
    uint64_t nU = idx->nU;
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t level = gen_level(idx);
        lv[i] = (uint8_t)level;
        rk[i] = idx->layer_count[level]++;
        uo[i] = (uint32_t)nU;
        nU += level;
    }
    int rc;
    if ((rc = ensure_points(idx, n1))) return rc;
    if ((rc = ensure_upper(idx, std::max<uint64_t>(nU, 1)))) return rc;
    // cudaMemcpyDefault: `sigs` may be a host pointer (gsb_index_insert_batch) or a device pointer
    // (gsb_index_insert_batch_dev: signatures straight from the sketcher / the all-gather buffer)
    GSB_CUDA_TRY(cudaMemcpyAsync(idx->d_sigs.as<uint8_t>() + n0 * row, sigs, n * row, cudaMemcpyDefault, st));
    GSB_CUDA_TRY(cudaMemcpyAsync(idx->d_ids.as<uint64_t>() + n0, ids, n * 8, cudaMemcpyHostToDevice, st));
    GSB_CUDA_TRY(cudaMemcpyAsync(idx->d_levels.as<uint8_t>() + n0, lv.data(), n, cudaMemcpyHostToDevice, st));
    GSB_CUDA_TRY(cudaMemcpyAsync(idx->d_ranks.as<uint32_t>() + n0, rk.data(), n * 4, cudaMemcpyHostToDevice, st));
    GSB_CUDA_TRY(cudaMemcpyAsync(idx->d_upper_off.as<uint32_t>() + n0, uo.data(), n * 4, cudaMemcpyHostToDevice, st));
    GSB_CUDA_TRY(cudaStreamSynchronize(st));  // lv/rk/uo are pageable: the copies must be done before they go
    idx->levels.insert(idx->levels.end(), lv.begin(), lv.end());
    idx->upper_off.insert(idx->upper_off.end(), uo.begin(), uo.end());
    idx->nU = nU;
    const uint32_t wmax = idx->wave_max;
    const size_t M = idx->M;
    // room for world * ceil(W / world) points: the all-gather works on equal slices
    const size_t wcap = (size_t)wmax + (comm ? (size_t)comm_size(comm) : 0);
    if ((rc = idx->d_sel_n.ensure(wcap * kMaxLayers * 4))) return rc;
    if ((rc = idx->d_sel_idx.ensure(wcap * 18 * M * 4))) return rc;
    if ((rc = idx->d_sel_d.ensure(wcap * 18 * M * 4))) return rc;
    if ((rc = idx->d_counter.ensure(256))) return rc;
    // ---- waves
    uint64_t i = 0;
    while (i < n) {
        if (idx->n == 0) {  // first point of the index: becomes the entry point
            idx->entry = 0;
            idx->n = 1;
            i++;
            continue;
        }
        uint32_t W = wave_size(idx->n, wmax);
        if (W > n - i) W = (uint32_t)(n - i);
        const uint32_t first = (uint32_t)idx->n;
        switch (idx->p.sig_type) {
        case GSB_SIG_U64: rc = launch_wave<8, false>(idx, first, W, st, comm); break;
        case GSB_SIG_U32: rc = launch_wave<4, false>(idx, first, W, st, comm); break;
        case GSB_SIG_F32: rc = launch_wave<4, true>(idx, first, W, st, comm); break;
        default: rc = launch_wave<2, false>(idx, first, W, st, comm); break;
        }
        if (rc) return rc;
        for (uint32_t t = 0; t < W; t++)
            if (idx->levels[first + t] > idx->levels[idx->entry]) idx->entry = first + t;
        idx->n += W;
        i += W;
    }
    GSB_CUDA_TRY(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_index_insert_batch(gsb_index *idx, const void *sigs, const uint64_t *ids, uint64_t n) {
    return insert_batch_impl(idx, sigs, ids, n, nullptr);
}

extern "C" int gsb_index_insert_batch_dev(gsb_index *idx, const void *d_sigs, const uint64_t *ids, uint64_t n) {
    return insert_batch_impl(idx, d_sigs, ids, n, nullptr);
}

// Sharded construction: EVERY rank of `comm` calls this with the same signatures (host or device
// pointer), ids and index parameters (same level_seed, same wave_max); every rank ends with the same
// graph, the one gsb_index_insert_batch builds on one GPU with that wave_max.
extern "C" int gsb_index_insert_batch_sharded(gsb_index *idx, gsb_comm *comm, const void *sigs, const uint64_t *ids,
                                              uint64_t n) {
    if (!comm) {
        set_error("gsb_index_insert_batch_sharded: NULL communicator");
        return GSB_ERR_INVALID_ARG;
    }
    if (idx && comm_device(comm) != idx->device) {
        set_error("gsb_index_insert_batch_sharded: the communicator lives on device %d, the index on device %d",
                  comm_device(comm), idx->device);
        return GSB_ERR_INVALID_ARG;
    }
    return insert_batch_impl(idx, sigs, ids, n, comm);
}

// ------------------------------------------------------------------------------ export / dump
namespace {
struct HostGraph {
    std::vector<uint8_t> levels;
    std::vector<uint32_t> ranks, cnt0, nbr0, upper_off, cntU, nbrU;
    std::vector<float> dist0, distU;
    std::vector<uint64_t> ids;
};
int fetch_graph(const gsb_index *idx, HostGraph &hg) {
    const uint64_t n = idx->n, nU = idx->nU;
    const size_t M = idx->M;
    cudaSetDevice(idx->device);
    cudaStreamSynchronize(idx->stream);
    hg.levels.resize(n);
    hg.ranks.resize(n);
    hg.ids.resize(n);
    hg.cnt0.resize(n);
    hg.upper_off.resize(n);
    hg.nbr0.resize(n * 2 * M);
    hg.dist0.resize(n * 2 * M);
    hg.cntU.resize(nU);
    hg.nbrU.resize(nU * M);
    hg.distU.resize(nU * M);
    if (!n) return GSB_OK;
    GSB_CUDA_TRY(cudaMemcpy(hg.levels.data(), idx->d_levels.p, n, cudaMemcpyDeviceToHost));
    GSB_CUDA_TRY(cudaMemcpy(hg.ranks.data(), idx->d_ranks.p, n * 4, cudaMemcpyDeviceToHost));
    GSB_CUDA_TRY(cudaMemcpy(hg.ids.data(), idx->d_ids.p, n * 8, cudaMemcpyDeviceToHost));
    GSB_CUDA_TRY(cudaMemcpy(hg.cnt0.data(), idx->d_cnt0.p, n * 4, cudaMemcpyDeviceToHost));
    GSB_CUDA_TRY(cudaMemcpy(hg.upper_off.data(), idx->d_upper_off.p, n * 4, cudaMemcpyDeviceToHost));
    GSB_CUDA_TRY(cudaMemcpy(hg.nbr0.data(), idx->d_nbr0.p, n * 2 * M * 4, cudaMemcpyDeviceToHost));
    GSB_CUDA_TRY(cudaMemcpy(hg.dist0.data(), idx->d_dist0.p, n * 2 * M * 4, cudaMemcpyDeviceToHost));
    if (nU) {
        GSB_CUDA_TRY(cudaMemcpy(hg.cntU.data(), idx->d_cntU.p, nU * 4, cudaMemcpyDeviceToHost));
        GSB_CUDA_TRY(cudaMemcpy(hg.nbrU.data(), idx->d_nbrU.p, nU * M * 4, cudaMemcpyDeviceToHost));
        GSB_CUDA_TRY(cudaMemcpy(hg.distU.data(), idx->d_distU.p, nU * M * 4, cudaMemcpyDeviceToHost));
    }
    return GSB_OK;
}
}  // namespace

extern "C" int gsb_index_graph_sizes(const gsb_index *idx, uint64_t *total_lists, uint64_t *total_nbrs) {
    std::unique_lock<std::recursive_mutex> lock_;
    if (idx) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<gsb_index *>(idx)->mu);
    if (!idx || !total_lists || !total_nbrs) {
        set_error("gsb_index_graph_sizes: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    HostGraph hg;
    int rc = fetch_graph(idx, hg);
    if (rc) return rc;
    uint64_t tl = 0, tn = 0;
    for (uint64_t p = 0; p < idx->n; p++) {
        tl += (uint64_t)hg.levels[p] + 1;
        tn += hg.cnt0[p];
        for (uint32_t l = 1; l <= hg.levels[p]; l++) tn += hg.cntU[hg.upper_off[p] + l - 1];
    }
    *total_lists = tl;
    *total_nbrs = tn;
    return GSB_OK;
}

extern "C" int gsb_index_export_graph(const gsb_index *idx, uint8_t *levels, uint32_t *ranks, uint64_t *ids,
                                      uint64_t *nbr_offsets, uint32_t *nbr_index, float *nbr_dist,
                                      uint64_t *entry_point) {
    std::unique_lock<std::recursive_mutex> lock_;
    if (idx) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<gsb_index *>(idx)->mu);
    if (!idx || !nbr_offsets || !entry_point) {
        set_error("gsb_index_export_graph: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    HostGraph hg;
    int rc = fetch_graph(idx, hg);
    if (rc) return rc;
    const size_t M = idx->M;
    uint64_t li = 0, off = 0;
    for (uint64_t p = 0; p < idx->n; p++) {
        if (levels) levels[p] = hg.levels[p];
        if (ranks) ranks[p] = hg.ranks[p];
        if (ids) ids[p] = hg.ids[p];
        for (uint32_t l = 0; l <= hg.levels[p]; l++) {
            nbr_offsets[li++] = off;
            const uint32_t cnt = l == 0 ? hg.cnt0[p] : hg.cntU[hg.upper_off[p] + l - 1];
            const uint32_t *src = l == 0 ? &hg.nbr0[p * 2 * M] : &hg.nbrU[(size_t)(hg.upper_off[p] + l - 1) * M];
            const float *sd = l == 0 ? &hg.dist0[p * 2 * M] : &hg.distU[(size_t)(hg.upper_off[p] + l - 1) * M];
            for (uint32_t i = 0; i < cnt; i++) {
                if (nbr_index) nbr_index[off] = src[i];
                if (nbr_dist) nbr_dist[off] = sd[i];
                off++;
            }
        }
    }
    nbr_offsets[li] = off;
    *entry_point = idx->n ? idx->entry : UINT64_MAX;
    return GSB_OK;
}

// file_dump / reload.  Two files like hnsw_rs::hnswio (<basename>.hnsw.graph, <basename>.hnsw.data)
// but in THIS library's own layout (little endian, documented in DESIGN.md): the byte layout of
// hnswio is only recalled, not verifiable here, so compatibility is not claimed.
namespace {
constexpr uint64_t kMagicGraph = 0x31485247425347ULL;  // "GSBGRH1"
constexpr uint64_t kMagicData = 0x31544144425347ULL;   // "GSBDAT1"
template <class T>
bool wr(FILE *f, const T *p, size_t n) {
    return n == 0 || fwrite(p, sizeof(T), n, f) == n;
}
template <class T>
bool rd(FILE *f, T *p, size_t n) {
    return n == 0 || fread(p, sizeof(T), n, f) == n;
}
}  // namespace

extern "C" int gsb_index_dump(const gsb_index *idx, const char *dir, const char *basename) {
    std::unique_lock<std::recursive_mutex> lock_;
    if (idx) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<gsb_index *>(idx)->mu);
    if (!idx || !dir || !basename) {
        set_error("gsb_index_dump: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    HostGraph hg;
    int rc = fetch_graph(idx, hg);
    if (rc) return rc;
    const uint64_t n = idx->n;
    uint64_t tl = 0;
    for (uint64_t p = 0; p < n; p++) tl += (uint64_t)hg.levels[p] + 1;
    std::vector<uint64_t> off(tl + 1);
    uint64_t tn = 0;
    {
        uint64_t li = 0;
        for (uint64_t p = 0; p < n; p++)
            for (uint32_t l = 0; l <= hg.levels[p]; l++) {
                off[li++] = tn;
                tn += l == 0 ? hg.cnt0[p] : hg.cntU[hg.upper_off[p] + l - 1];
            }
        off[li] = tn;
    }
    std::vector<uint32_t> nidx(tn);
    std::vector<float> ndist(tn);
    uint64_t entry = 0;
    rc = gsb_index_export_graph(idx, nullptr, nullptr, nullptr, off.data(), nidx.data(), ndist.data(), &entry);
    if (rc) return rc;
    const std::string base = std::string(dir) + "/" + basename;
    FILE *fg = fopen((base + ".hnsw.graph").c_str(), "wb");
    FILE *fd = fopen((base + ".hnsw.data").c_str(), "wb");
    bool ok = fg && fd;
    if (ok) {
        const uint64_t hdr[10] = {kMagicGraph, n, idx->M, idx->p.max_layer, idx->p.ef_construction,
                                  idx->p.sig_type, idx->p.sketch_size, entry, tl, tn};
        ok = wr(fg, hdr, 10) && wr(fg, &idx->p.scale_modification, 1) && wr(fg, hg.levels.data(), n) &&
             wr(fg, hg.ranks.data(), n) && wr(fg, off.data(), tl + 1) && wr(fg, nidx.data(), tn) &&
             wr(fg, ndist.data(), tn);
        const uint64_t dh[4] = {kMagicData, n, idx->p.sketch_size, idx->elem};
        std::vector<uint8_t> sig((size_t)n * idx->p.sketch_size * idx->elem);
        if (n && cudaMemcpy(sig.data(), idx->d_sigs.p, sig.size(), cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
        ok = ok && wr(fd, dh, 4) && wr(fd, hg.ids.data(), n) && wr(fd, sig.data(), sig.size());
    }
    if (fg) ok = (fclose(fg) == 0) && ok;
    if (fd) ok = (fclose(fd) == 0) && ok;
    if (!ok) {
        set_error("gsb_index_dump: cannot write %s.hnsw.{graph,data}", base.c_str());
        return GSB_ERR_IO;
    }
    return GSB_OK;
}

extern "C" int gsb_index_load(gsb_index *idx, const char *dir, const char *basename) {
    std::unique_lock<std::recursive_mutex> lock_;
    if (idx) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<gsb_index *>(idx)->mu);
    if (!idx || !dir || !basename) {
        set_error("gsb_index_load: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    const std::string base = std::string(dir) + "/" + basename;
    FILE *fg = fopen((base + ".hnsw.graph").c_str(), "rb");
    FILE *fd = fopen((base + ".hnsw.data").c_str(), "rb");
    int rc = GSB_ERR_IO;
    do {
        if (!fg || !fd) {
            set_error("gsb_index_load: cannot open %s.hnsw.{graph,data}", base.c_str());
            break;
        }
        uint64_t hdr[10], dh[4];
        double scale;
        if (!rd(fg, hdr, 10) || !rd(fg, &scale, 1) || !rd(fd, dh, 4) || hdr[0] != kMagicGraph || dh[0] != kMagicData) {
            set_error("gsb_index_load: %s is not a gsearch_b200 dump", base.c_str());
            break;
        }
        const uint64_t n = hdr[1], tl = hdr[8], tn = hdr[9];
        if (hdr[2] != idx->M || hdr[5] != idx->p.sig_type || hdr[6] != idx->p.sketch_size || dh[1] != n ||
            dh[2] != idx->p.sketch_size || dh[3] != idx->elem) {
            set_error("gsb_index_load: dump (M=%llu, sig_type=%llu, S=%llu) does not match this index",
                      (unsigned long long)hdr[2], (unsigned long long)hdr[5], (unsigned long long)hdr[6]);
            rc = GSB_ERR_INVALID_ARG;
            break;
        }
        // sizes come from an untrusted header: check them against the file lengths before allocating
        const size_t row = (size_t)idx->p.sketch_size * idx->elem;
        auto file_size = [](FILE *f) -> uint64_t {
            const long at = ftell(f);
            fseek(f, 0, SEEK_END);
            const long sz = ftell(f);
            fseek(f, at, SEEK_SET);
            return sz < 0 ? 0 : (uint64_t)sz;
        };
        const uint64_t gsz = file_size(fg), dsz = file_size(fd);
        const uint64_t lim = 1ull << 40;
        if (n > lim || tl > lim || tn > lim || tl < n || tl > n * (uint64_t)kMaxLayers ||
            gsz != 88 + n * 5 + (tl + 1) * 8 + tn * 8 || dsz != 32 + n * 8 + n * row) {
            set_error("gsb_index_load: header of %s does not match the file sizes (corrupt or truncated dump)",
                      base.c_str());
            break;
        }
        try {
            std::vector<uint8_t> levels(n), sig((size_t)n * row);
            std::vector<uint32_t> ranks(n), nidx(tn);
            std::vector<uint64_t> off(tl + 1), ids(n);
            std::vector<float> ndist(tn);
            if (!rd(fg, levels.data(), n) || !rd(fg, ranks.data(), n) || !rd(fg, off.data(), tl + 1) ||
                !rd(fg, nidx.data(), tn) || !rd(fg, ndist.data(), tn) || !rd(fd, ids.data(), n) ||
                !rd(fd, sig.data(), sig.size())) {
                set_error("gsb_index_load: truncated dump %s", base.c_str());
                break;
            }
            // offsets: monotone, starting at 0, ending at the neighbour count; one list per (point, layer)
            uint64_t want_tl = 0;
            for (uint64_t p = 0; p < n; p++) want_tl += (uint64_t)levels[p] + 1;
            bool ok = want_tl == tl && off[0] == 0 && off[tl] == tn;
            for (uint64_t i = 0; ok && i < tl; i++) ok = off[i] <= off[i + 1];
            if (!ok) {
                set_error("gsb_index_load: inconsistent neighbour offsets in %s", base.c_str());
                rc = GSB_ERR_INVALID_ARG;
                break;
            }
            rc = gsb_index_load_graph(idx, sig.data(), ids.data(), n, levels.data(), ranks.data(), off.data(),
                                      nidx.data(), ndist.data(), hdr[7]);
        } catch (const std::exception &) {
            set_error("gsb_index_load: out of host memory reading %s", base.c_str());
            rc = GSB_ERR_OOM;
        }
    } while (0);
    if (fg) fclose(fg);
    if (fd) fclose(fd);
    return rc;
}
