// hnsw_search.cuh -- K7: on-device HNSW search with DistHamming (sm_100a).
//
// Replaces hnsw_rs::Hnsw::<Sig,DistHamming>::{search, search_layer, parallel_search} [U] as
// called at src/dna/dnarequest.rs:353 (ef_search = 5000, src/bin/gsearch.rs:893).
//
// One CTA owns one query at a time (persistent CTAs pull queries from a counter = the
// reference's rayon par_iter over queries).  The query signature stays in shared memory for
// the whole search (TMA bulk copy), the result heap (ef entries) lives in shared memory too,
// and every neighbour expansion evaluates its <= 2M unvisited candidates with all warps
// streaming candidate signatures from HBM (K6's inner loop).  Heap order, visit order and tie
// behaviour reproduce Rust's std BinaryHeap exactly (see oracle/hnsw.c), so that on the same
// graph the returned ids are identical to the CPU restatement.
#pragma once

#include "common.cuh"
#include "hamming.cuh"

namespace gsb {

constexpr int kSearchThreads = 512;
constexpr int kMaxList = 512;  // >= 2 * max_nb_connection (<= 255)

struct HItem {
    float d;
    uint32_t p;
};

struct DHeap {
    HItem *a;
    uint32_t n;
    __device__ __forceinline__ void sift_up(uint32_t start, uint32_t pos) {
        const HItem e = a[pos];
        while (pos > start) {
            const uint32_t parent = (pos - 1) >> 1;
            if (e.d <= a[parent].d) break;
            a[pos] = a[parent];
            pos = parent;
        }
        a[pos] = e;
    }
    __device__ __forceinline__ void push(float d, uint32_t p) {
        a[n].d = d;
        a[n].p = p;
        n++;
        sift_up(0, n - 1);
    }
    __device__ __forceinline__ void sift_down_to_bottom(uint32_t pos) {
        const uint32_t end = n, start = pos;
        const HItem e = a[pos];
        uint32_t child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            child += (a[child].d <= a[child + 1].d) ? 1u : 0u;
            a[pos] = a[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (end >= 1 && child == end - 1) {
            a[pos] = a[child];
            pos = child;
        }
        a[pos] = e;
        sift_up(start, pos);
    }
    __device__ __forceinline__ HItem pop() {
        HItem item = a[n - 1];
        n--;
        if (n > 0) {
            const HItem t = a[0];
            a[0] = item;
            item = t;
            sift_down_to_bottom(0);
        }
        return item;
    }
    __device__ __forceinline__ void sift_down_range(uint32_t pos, uint32_t end) {
        const HItem e = a[pos];
        uint32_t child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            child += (a[child].d <= a[child + 1].d) ? 1u : 0u;
            if (e.d >= a[child].d) {
                a[pos] = e;
                return;
            }
            a[pos] = a[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (end >= 1 && child == end - 1 && e.d < a[child].d) {
            a[pos] = a[child];
            pos = child;
        }
        a[pos] = e;
    }
    __device__ __forceinline__ void into_sorted() {
        uint32_t end = n;
        while (end > 1) {
            end--;
            const HItem t = a[0];
            a[0] = a[end];
            a[end] = t;
            sift_down_range(0, end);
        }
    }
};

struct GraphView {
    const uint8_t *sigs;        // n x S x elem
    const uint64_t *ids;        // origin ids
    const uint8_t *levels;      // level of each point
    const uint32_t *ranks;      // rank in its layer
    const uint64_t *list_base;  // index of (p, layer 0) in nbr_off; (p, l) = list_base[p] + l
    const uint64_t *nbr_off;    // total_lists + 1
    const uint32_t *nbr_idx;    // neighbour point indices, each list sorted by distance
    uint32_t n;
    uint32_t entry;
    uint32_t S;
};

struct SearchOut {
    gsb_neighbour *out;   // nq x knbn
    uint32_t *counts;     // nq
    unsigned long long *nb_eval;  // nq (may be null)
};

template <int ELEM, bool F32>
__device__ __forceinline__ void eval_list(const uint8_t *smem_q, const GraphView &g, const uint32_t *E,
                                          uint32_t nE, float *D) {
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t row = (size_t)g.S * ELEM;
    const float fS = (float)g.S;
    for (uint32_t i = warp; i < nE; i += nwarps) {
        const uint32_t cnt = warp_row_count<ELEM, F32>(smem_q, g.sigs + (size_t)E[i] * row, g.S);
        if (lane_id() == 0) D[i] = __fdiv_rn((float)cnt, fS);
    }
}

// smem: [query: row bytes rounded to 128][ret heap: (ef+1) items if RET_SMEM]
template <int ELEM, bool F32>
__global__ void __launch_bounds__(kSearchThreads, 1)
k7_hnsw_search(GraphView g, const uint8_t *__restrict__ queries, uint32_t nq, uint32_t knbn, uint32_t ef,
               int ret_in_smem, uint8_t *__restrict__ ws, size_t ws_stride, SearchOut so,
               uint32_t *__restrict__ qcounter) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t E[kMaxList];
    __shared__ float D[kMaxList];
    __shared__ uint32_t wcnt[kSearchThreads / 32];
    __shared__ uint32_t s_nE, s_done, s_q, s_node, s_layer;
    const size_t row = (size_t)g.S * ELEM;
    const size_t row128 = (row + 127) & ~(size_t)127;
    uint8_t *my = ws + (size_t)blockIdx.x * ws_stride;
    HItem *cand_a = reinterpret_cast<HItem *>(my);
    uint8_t *visited = my + ((size_t)g.n + 1) * sizeof(HItem);
    HItem *ret_a = ret_in_smem ? reinterpret_cast<HItem *>(smem + row128)
                               : reinterpret_cast<HItem *>(my + ((size_t)g.n + 1) * sizeof(HItem) +
                                                           (((size_t)g.n + 15) & ~(size_t)15));
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const bool can_tma = (row & 15) == 0;
    for (;;) {
        if (threadIdx.x == 0) s_q = atomicAdd(qcounter, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= nq) break;
        const uint8_t *gq = queries + (size_t)q * row;
        if (can_tma && (((uintptr_t)gq) & 15) == 0) {
            // all generic-proxy reads of the previous query are done (barrier above)
            fence_proxy_async();
            stage_query(smem, gq, (uint32_t)row, &bar, phase);
            phase ^= 1;
        } else {
            for (uint32_t i = threadIdx.x; i < row; i += blockDim.x) smem[i] = gq[i];
        }
        for (uint32_t i = threadIdx.x; i < g.n; i += blockDim.x) visited[i] = 0;
        DHeap cand{cand_a, 0}, ret{ret_a, 0};
        unsigned long long neval = 0;
        uint32_t pivot = g.entry;
        float dist_to_entry = 0.f;
        if (threadIdx.x == 0) {
            E[0] = pivot;
            s_nE = 1;
        }
        __syncthreads();
        eval_list<ELEM, F32>(smem, g, E, 1, D);
        __syncthreads();
        dist_to_entry = D[0];
        neval += 1;
        // ---- one greedy hop per upper layer (hnsw_rs `search`)
        const int top = g.levels[g.entry];
        for (int layer = top; layer >= 1; layer--) {
            const uint64_t li = g.list_base[pivot] + (uint64_t)layer;
            const uint64_t b = g.nbr_off[li], e = g.nbr_off[li + 1];
            const uint32_t len = (uint32_t)(e - b);
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) E[i] = g.nbr_idx[b + i];
            __syncthreads();
            eval_list<ELEM, F32>(smem, g, E, len, D);
            __syncthreads();
            neval += len;
            // every thread scans the same shared arrays: uniform result, no broadcast needed
            uint32_t newp = pivot;
            for (uint32_t i = 0; i < len; i++) {
                if (D[i] < dist_to_entry) {
                    dist_to_entry = D[i];
                    newp = E[i];
                }
            }
            pivot = newp;
        }
        __syncthreads();
        // ---- search_layer(q, pivot, ef, 0)
        if (threadIdx.x == 0) {
            visited[pivot] = 1;
            cand.push(-dist_to_entry, pivot);
            ret.push(dist_to_entry, pivot);
            // the distance to the layer-0 entry point is evaluated again by search_layer
            neval += 1;
        }
        for (;;) {
            if (threadIdx.x == 0) {
                s_done = 0;
                if (cand.n == 0) {
                    s_done = 1;
                } else {
                    const HItem c = cand.pop();
                    if (-c.d > ret.a[0].d) s_done = 1;
                    s_node = c.p;
                }
            }
            __syncthreads();
            if (s_done) break;
            // gather the unvisited neighbours of s_node in list order
            const uint64_t li = g.list_base[s_node];
            const uint64_t b = g.nbr_off[li], e = g.nbr_off[li + 1];
            const uint32_t len = (uint32_t)(e - b);
            uint32_t nb = 0xFFFFFFFFu;
            bool unv = false;
            if (threadIdx.x < len) {
                nb = g.nbr_idx[b + threadIdx.x];
                unv = visited[nb] == 0;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, unv);
            if (lane_id() == 0) wcnt[threadIdx.x >> 5] = __popc(bal);
            __syncthreads();
            uint32_t pre = 0, tot = 0;
            for (uint32_t w = 0; w < kSearchThreads / 32; w++) {
                if (w < (threadIdx.x >> 5)) pre += wcnt[w];
                tot += wcnt[w];
            }
            if (unv) {
                E[pre + __popc(bal & ((1u << lane_id()) - 1))] = nb;
                visited[nb] = 1;
            }
            __syncthreads();
            eval_list<ELEM, F32>(smem, g, E, tot, D);
            __syncthreads();
            if (threadIdx.x == 0) {
                neval += tot;
                for (uint32_t i = 0; i < tot; i++) {
                    const float ed = D[i];
                    if (ed < ret.a[0].d || ret.n < ef) {
                        cand.push(-ed, E[i]);
                        ret.push(ed, E[i]);
                        if (ret.n > ef) (void)ret.pop();
                    }
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            ret.into_sorted();
            uint32_t last = knbn < ef ? knbn : ef;
            if (ret.n < last) last = ret.n;
            for (uint32_t i = 0; i < last; i++) {
                const uint32_t p = ret.a[i].p;
                gsb_neighbour nbq;
                nbq.d_id = g.ids[p];
                nbq.distance = ret.a[i].d;
                nbq.layer = g.levels[p];
                nbq.pad_[0] = nbq.pad_[1] = nbq.pad_[2] = 0;
                nbq.rank = (int32_t)g.ranks[p];
                so.out[(size_t)q * knbn + i] = nbq;
            }
            so.counts[q] = last;
            if (so.nb_eval) so.nb_eval[q] = neval;
        }
        __syncthreads();
    }
}

}  // namespace gsb
