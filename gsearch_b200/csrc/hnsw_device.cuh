// hnsw_device.cuh -- K7/K8: on-device HNSW search and construction with DistHamming (sm_100a).
//
// Replaces hnsw_rs::Hnsw::<Sig,DistHamming>::{search, search_layer, parallel_search,
// parallel_insert} [U] as called at src/dna/dnarequest.rs:353 (ef_search = 5000,
// src/bin/gsearch.rs:893) and src/dna/dnasketch.rs:435 (ef_construction = --ef, extend_candidates
// = true, keep_pruned = false, :159-160).
//
// One CTA owns one query (search) or one new point (insert) at a time; persistent CTAs pull work
// from a counter (the reference's rayon par_iter).  A search is a chain of ~ef expansions with one
// or two new rows each, i.e. serial phases (heap updates by thread 0, neighbour gathering) between
// short streaming phases, so SEVERAL CTAs share an SM (three for search, two for insertion): their
// chains hide each other's serial phases.  To make them fit, the query signature is NOT staged in
// shared memory (it is read through L1/L2 beside every candidate row); the result heap, the head
// of the candidate heap and the visited bitmap are.  Every neighbour expansion evaluates its <= 2M
// unvisited candidates with all warps of the CTA streaming candidate signatures from HBM (K6's
// inner loop).  Heap order, visit order and tie behaviour reproduce Rust's std BinaryHeap exactly
// (see oracle/hnsw.c), so that on the same graph the returned ids are identical to the CPU
// restatement.
// For large ef_search (the reference's 5000) K7 has a second layout, the TMA RING further down:
// candidate rows and the matching query pieces arrive as bulk copies in shared memory, a control
// warp runs the heaps while worker warps already stream the predicted next expansion.
//
// The graph is mutable and device resident: fixed-capacity adjacency (2M entries per point on
// layer 0, M per upper layer, with the distances kept beside the indices because
// reverse_update_neighborhood_simple keeps lists sorted by distance and drops the farthest).
//
// Construction follows the deterministic wave semantics of oracle/hnsw.c
// (gso_hnsw_insert_waves): phase A = every point of a wave searches the graph as it was before
// the wave and also sees the earlier points of its wave; phase B = own lists, then reverse
// updates (order independent: a list keeps its M / 2M smallest by (distance, index)).
#pragma once

#include "common.cuh"
#include "hamming.cuh"

namespace gsb {

constexpr int kSearchThreads = 256;   // K7: 8 warps per CTA, three CTAs per SM (three independent search chains)
#ifndef GSB_K8_PER_SM
#define GSB_K8_PER_SM 2
#endif
constexpr int kInsertThreads = 256;   // K8: 8 warps per CTA, GSB_K8_PER_SM CTAs per SM
constexpr int kMaxList = 512;    // >= 2 * max_nb_connection (<= 255): list scratch of a CTA
constexpr int kMaxWave = 4096;   // points of one insertion wave (wave mates are met kMaxList at a time)
constexpr int kMaxLayers = 17;   // levels 0..16

struct HItem {
    float d;
    uint32_t p;
};

// Binary heap of thread 0, Rust std::collections::BinaryHeap order of operations.  HYB: the
// first `cs` entries live in shared memory and the rest in the CTA's global workspace -- a sift
// walks log2(n) DEPENDENT entries, and a global store invalidates the L1 line, so a heap kept in
// global memory runs at L2 latency (the candidate heap is popped once per expansion).
template <bool HYB>
struct HeapT {
    HItem *a;     // shared-memory part (HYB) or the whole heap
    HItem *g;     // HYB: entry i >= cs is g[i - cs]
    uint32_t cs;
    uint32_t n;
    __device__ __forceinline__ HItem &at(uint32_t i) {
        if (HYB) return i < cs ? a[i] : g[i - cs];
        return a[i];
    }
    __device__ __forceinline__ void sift_up(uint32_t start, uint32_t pos) {
        const HItem e = at(pos);
        while (pos > start) {
            const uint32_t parent = (pos - 1) >> 1;
            const HItem pe = at(parent);
            if (e.d <= pe.d) break;
            at(pos) = pe;
            pos = parent;
        }
        at(pos) = e;
    }
    __device__ __forceinline__ void push(float d, uint32_t p) {
        HItem e;
        e.d = d;
        e.p = p;
        at(n) = e;
        n++;
        sift_up(0, n - 1);
    }
    __device__ __forceinline__ void sift_down_to_bottom(uint32_t pos) {
        const uint32_t end = n, start = pos;
        const HItem e = at(pos);
        uint32_t child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            const HItem c0 = at(child), c1 = at(child + 1);
            const bool right = c0.d <= c1.d;
            child += right ? 1u : 0u;
            at(pos) = right ? c1 : c0;
            pos = child;
            child = 2 * pos + 1;
        }
        if (end >= 1 && child == end - 1) {
            at(pos) = at(child);
            pos = child;
        }
        at(pos) = e;
        sift_up(start, pos);
    }
    __device__ __forceinline__ HItem pop() {
        HItem item = at(n - 1);
        n--;
        if (n > 0) {
            const HItem t = at(0);
            at(0) = item;
            item = t;
            sift_down_to_bottom(0);
        }
        return item;
    }
    __device__ __forceinline__ void sift_down_range(uint32_t pos, uint32_t end) {
        const HItem e = at(pos);
        uint32_t child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            const HItem c0 = at(child), c1 = at(child + 1);
            const bool right = c0.d <= c1.d;
            child += right ? 1u : 0u;
            const HItem c = right ? c1 : c0;
            if (e.d >= c.d) {
                at(pos) = e;
                return;
            }
            at(pos) = c;
            pos = child;
            child = 2 * pos + 1;
        }
        if (end >= 1 && child == end - 1 && e.d < at(child).d) {
            at(pos) = at(child);
            pos = child;
        }
        at(pos) = e;
    }
    __device__ __forceinline__ void into_sorted() {
        uint32_t end = n;
        while (end > 1) {
            end--;
            const HItem t = at(0);
            at(0) = at(end);
            at(end) = t;
            sift_down_range(0, end);
        }
    }
};
constexpr uint32_t kCandSmem = 2048;  // entries of the candidate heap kept in shared memory (K8)
#ifndef GSB_CAND_K7
#define GSB_CAND_K7 512
#endif
constexpr uint32_t kCandSmemK7 = GSB_CAND_K7;  // K7: three CTAs per SM, the row ring takes the room

struct GraphView {
    const uint8_t *sigs;        // n x S x elem
    const uint64_t *ids;        // origin ids
    const uint8_t *levels;      // level of each point
    const uint32_t *ranks;      // rank in its layer
    uint32_t *nbr0;             // [cap x 2M] layer-0 neighbour indices, sorted by (distance, index)
    float *dist0;               // [cap x 2M]
    uint32_t *cnt0;             // [cap]
    const uint32_t *upper_off;  // [cap] list index of (p, layer 1) in the upper pool
    uint32_t *nbrU;             // [capU x M]
    float *distU;               // [capU x M]
    uint32_t *cntU;             // [capU]
    uint32_t *lock0, *lockU;    // per-list locks (phase B)
    uint32_t M, n, entry, S;
};

// Neighbour lists are only written by the phase-B kernels, never while a search or phase-A
// kernel runs, so they are read through L1 (and prefetched there one expansion ahead).
__device__ __forceinline__ const uint32_t *list_of(const GraphView &g, uint32_t p, uint32_t layer, uint32_t &len) {
    if (layer == 0) {
        len = __ldg(&g.cnt0[p]);
        return g.nbr0 + (size_t)p * 2 * g.M;
    }
    const uint32_t li = g.upper_off[p] + layer - 1;
    len = __ldg(&g.cntU[li]);
    return g.nbrU + (size_t)li * g.M;
}

// the distances stored beside a neighbour list (same indexing as list_of)
__device__ __forceinline__ const float *dist_of(const GraphView &g, uint32_t p, uint32_t layer) {
    if (layer == 0) return g.dist0 + (size_t)p * 2 * g.M;
    return g.distU + (size_t)(g.upper_off[p] + layer - 1) * g.M;
}

// warp 0: pull the list of the candidate that will most likely be popped next into L1
__device__ __forceinline__ void prefetch_list(const GraphView &g, uint32_t p, uint32_t layer) {
    const uint32_t *base;
    const uint32_t *cnt;
    uint32_t bytes;
    if (layer == 0) {
        base = g.nbr0 + (size_t)p * 2 * g.M;
        cnt = g.cnt0 + p;
        bytes = 8 * g.M;
    } else {
        const uint32_t li = __ldg(&g.upper_off[p]) + layer - 1;
        base = g.nbrU + (size_t)li * g.M;
        cnt = g.cntU + li;
        bytes = 4 * g.M;
    }
    const uint32_t off = threadIdx.x * 128u;
    if (off < bytes) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const uint8_t *>(base) + off));
    if (threadIdx.x == 31) asm volatile("prefetch.global.L1 [%0];" ::"l"(cnt));
}

struct SearchOut {
    gsb_neighbour *out;   // nq x knbn
    uint32_t *counts;     // nq
    unsigned long long *nb_eval;  // nq (may be null)
};

// scratch of one CTA
struct HnswShared {
    uint32_t E[kMaxList];
    float D[kMaxList];
    uint32_t acc[kMaxList];  // eval_list scratch
    uint32_t wcnt[32];
    uint32_t done, node, flag, work, next;
    float fval;
    HeapT<true> cand;   // owned by thread 0
    HeapT<true> ret;    // first `cs` entries in shared memory, the rest in the CTA's workspace
};

// distances of the staged row to the nE rows E[i] -> D[i].  One warp per row when there are
// enough rows; a short list is split so that every warp streams a segment of a row (the serial
// chain of a graph search is made of short expansions, and a warp alone is latency bound).
// All threads of the CTA must call it (it synchronises when it splits rows).
// visited set of one search_layer call: a bitmap in shared memory when the index fits (the
// gather step of every expansion tests up to 2M neighbours; from global memory that is one more
// dependent L2 round trip per expansion), visit stamps in the CTA's global workspace otherwise
struct Visit {
    uint32_t *bits;    // shared-memory bitmap, or nullptr
    uint32_t nwords;
    uint32_t *stamps;  // global: stamps[p] == stamp marks p visited
    uint32_t stamp;
    __device__ __forceinline__ void begin() {  // all threads of the CTA
        if (bits) {
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) bits[i] = 0;
            __syncthreads();
        } else {
            stamp++;
        }
    }
    __device__ __forceinline__ bool seen(uint32_t x) const {
        return bits ? ((bits[x >> 5] >> (x & 31u)) & 1u) != 0u : stamps[x] == stamp;
    }
    __device__ __forceinline__ void mark(uint32_t x) {
        if (bits) atomicOr(&bits[x >> 5], 1u << (x & 31u));
        else stamps[x] = stamp;
    }
    __device__ __forceinline__ void unmark(uint32_t x) {  // (stamps start at 1: 0 is "never visited")
        if (bits) atomicAnd(&bits[x >> 5], ~(1u << (x & 31u)));
        else stamps[x] = 0;
    }
};

template <int ELEM, bool F32>
__device__ __forceinline__ void eval_list(const uint8_t *smem_q, const GraphView &g, const uint32_t *E,
                                          uint32_t nE, float *D, uint32_t *acc /* [kMaxList] shared scratch */) {
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t row = (size_t)g.S * ELEM;
    const float fS = (float)g.S;
    const bool aligned = (row & 15) == 0;  // rows start 16-byte aligned iff the row size is a multiple of 16
    if (nE == 0) return;
    if (nE * 2 > nwarps || !aligned) {
        for (uint32_t i = warp; i < nE; i += nwarps) {
            const uint8_t *cr = g.sigs + (size_t)E[i] * row;
            uint32_t cnt;
            if (aligned) {
                cnt = warp_row_count<ELEM, F32>(smem_q, cr, g.S);
            } else {  // element-wise (row size not a multiple of 16)
                cnt = 0;
                for (uint32_t e = lane_id(); e < g.S; e += 32) {
                    if (ELEM == 8) cnt += ((const uint64_t *)smem_q)[e] != ((const uint64_t *)cr)[e];
                    else if (ELEM == 4) {
                        if (F32) cnt += ((const float *)smem_q)[e] != ((const float *)cr)[e];
                        else cnt += ((const uint32_t *)smem_q)[e] != ((const uint32_t *)cr)[e];
                    } else cnt += ((const uint16_t *)smem_q)[e] != ((const uint16_t *)cr)[e];
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            }
            if (lane_id() == 0) D[i] = __fdiv_rn((float)cnt, fS);
        }
        return;
    }
    // split: `parts` warps per row
    uint32_t parts = 1;
    while (parts * 2 * nE <= nwarps) parts *= 2;
    if (threadIdx.x < nE) acc[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t r = warp / parts, part = warp % parts;
    if (r < nE) {
        const uint8_t *cr = g.sigs + (size_t)E[r] * row;
        const uint32_t nvec = (uint32_t)(row / 16);
        const uint32_t per = ((nvec + parts - 1) / parts + 31) & ~31u;
        const uint32_t vb = part * per, ve = vb + per < nvec ? vb + per : nvec;
        uint32_t cnt = vb < ve ? warp_vec_count<ELEM, F32>(reinterpret_cast<const uint4 *>(smem_q),
                                                           reinterpret_cast<const uint4 *>(cr), vb, ve)
                               : 0u;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane_id() == 0 && cnt) atomicAdd(&acc[r], cnt);
    }
    __syncthreads();
    if (threadIdx.x < nE) D[threadIdx.x] = __fdiv_rn((float)acc[threadIdx.x], fS);
}

// ---- K7: candidate rows AND the matching query pieces through a TMA ring
// A search chain evaluates one or two rows per expansion (ef_search = 5000: 1.45 on average), so
// its bandwidth is the bytes it keeps in flight, and the query (144 KB at S = 18000 x u64) does not
// fit in shared memory three times per SM: it is re-read from L2 beside every row.  Measured on the
// B200 with this access pattern (scripts/ubench_rowstream.cu, profiles/r2_ubench_rowstream.log):
//   rows alone, TMA or register loads                              7.0 - 7.4 TB/s
//   rows by TMA + query by register loads (LDG, L2 hits)           3.4 TB/s   <- round 1 / first ring
//   rows + query both by TMA, default L2 policy, 444 CTAs          3.8 TB/s   (64 MB of queries thrash L2)
//   rows (evict-first) + query (evict-last) both by TMA, 444 CTAs  7.0 TB/s of rows
// So ONE thread streams, per job, kRingChunk bytes of a row and the same piece of the query as two
// bulk copies (cp.async.bulk + mbarrier: SASS UBLKCP) onto one transaction barrier; the worker threads
// compare the two pieces out of shared memory.  No register loads, no LSU slots, and the L2 policies
// keep the queries resident under the stream of rows.  full[s]: both copies landed; empty[s]: the
// eight warps are done with the slot.
//
// Roles inside the CTA (ring kernel): warp 0 is the CONTROL warp -- its lane 0 owns the two heaps --
// and warps 1..7 are the 224 WORKERS that gather neighbours and compare rows.  While the control
// lane files the results of expansion i into the heaps (a serial chain of dependent sift steps,
// ~6 000 cycles), the workers already gather and stream the expansion the search will most likely
// do next (search_layer_ring below): the heap work left the critical path.
// Ring geometry.  A bulk copy has a fixed cost of a few hundred cycles in the copy engine besides its
// bytes (per job: 824 cycles for 2 x 7 KB, measured with clock64 -- the same for 3 or 10 slots, for one
// or three CTAs per SM, for one or seven issuing threads), so the copies are made as large as the
// shared memory of three CTAs per SM allows and the ring as shallow as a double buffer:
//   3 slots x 7 KB   3 840 queries/s     2 slots x 10.5 KB   4 020     2 slots x 14 KB   4 370
// (same box, ef_search 5000; the last one gives up the shared-memory visited bitmap for room).
#ifndef GSB_RING_SLOTS
#define GSB_RING_SLOTS 2
#define GSB_RING_VEC 4
#endif
constexpr uint32_t kRingSlots = GSB_RING_SLOTS;
constexpr uint32_t kRingWorkers = kSearchThreads - 32;   // 224
constexpr uint32_t kRingVec = GSB_RING_VEC;              // 16-byte vectors of the row per worker and job
constexpr uint32_t kRingChunk = kRingWorkers * 16 * kRingVec;   // 14 336 bytes of a row per job
constexpr uint32_t kRingBytes = kRingSlots * 2 * kRingChunk;
struct RowRing {
    uint8_t *buf;             // kRingSlots x [row piece | query piece], 128-byte aligned
    uint64_t *full, *empty;   // [kRingSlots] each
    uint32_t slot, par;       // next job's slot and the parity of its use (uniform among the workers)
    uint64_t pol_rows, pol_query;
};
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(kRingWorkers) : "memory"); }

// workers only (threadIdx.x >= 32); ends with a workers' barrier: D[0..nE) is complete for them
template <int ELEM, bool F32>
__device__ __noinline__ void eval_list_ring(const uint8_t *q, const GraphView &g, const uint32_t *E, uint32_t nE,
                                            float *D, uint32_t *acc, RowRing &ring) {
    if (nE == 0) return;
    RowRing rr = ring;  // (the caller's copy lives in local memory: this is a real call)
    const uint32_t row = g.S * ELEM;
    const uint32_t nch = (row + kRingChunk - 1) / kRingChunk;
    const uint32_t J = nE * nch;
    const uint32_t t = threadIdx.x - 32;  // worker index
    for (uint32_t i = t; i < nE; i += kRingWorkers) acc[i] = 0;
    bar_workers();
    uint32_t pr = 0, pc = 0, pj = 0, pslot = rr.slot;  // producer (worker 0): next job to issue
    auto issue = [&]() {
        const uint32_t off = pc * kRingChunk;
        const uint32_t bytes = row - off < kRingChunk ? row - off : kRingChunk;
        uint8_t *dst = rr.buf + pslot * (2 * kRingChunk);
        mbar_expect_tx(&rr.full[pslot], 2 * bytes);
        tma_bulk_g2s_hint(dst, g.sigs + (size_t)E[pr] * row + off, bytes, &rr.full[pslot], rr.pol_rows);
        tma_bulk_g2s_hint(dst + kRingChunk, q + off, bytes, &rr.full[pslot], rr.pol_query);
        if (++pslot == kRingSlots) pslot = 0;
        if (++pc == nch) {
            pc = 0;
            pr++;
        }
        pj++;
    };
    if (t == 0)
        while (pj < J && pj < kRingSlots) issue();
    uint32_t r = 0, c = 0, cnt = 0;
    for (uint32_t j = 0; j < J; j++) {
        const uint32_t off = c * kRingChunk;
        const uint32_t nv = (row - off < kRingChunk ? row - off : kRingChunk) / 16;
        mbar_wait(&rr.full[rr.slot], rr.par);  // (one polling lane per warp + __syncwarp: measured 14 % slower)
        const uint4 *s4 = reinterpret_cast<const uint4 *>(rr.buf + rr.slot * (2 * kRingChunk));
        const uint4 *sq = s4 + kRingChunk / 16;
#pragma unroll
        for (uint32_t u = 0; u < kRingVec; u++)
            if (t + u * kRingWorkers < nv) cnt += diff16<ELEM, F32>(sq[t + u * kRingWorkers], s4[t + u * kRingWorkers]);
        if (++c == nch) {  // the row is complete (uniform)
            c = 0;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            if (lane_id() == 0 && cnt) atomicAdd(&acc[r], cnt);
            cnt = 0;
            r++;
        }
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&rr.empty[rr.slot]);
        if (t == 0 && pj < J) {  // refill this slot as soon as the seven warps have released it
            mbar_wait(&rr.empty[rr.slot], rr.par);
            issue();
        }
        if (++rr.slot == kRingSlots) {
            rr.slot = 0;
            rr.par ^= 1u;
        }
    }
    ring.slot = rr.slot;
    ring.par = rr.par;
    bar_workers();
    const float fS = (float)g.S;
    for (uint32_t i = t; i < nE; i += kRingWorkers) D[i] = __fdiv_rn((float)acc[i], fS);
    bar_workers();
}

// stage one signature row in shared memory (TMA bulk copy when alignment allows) and return where
// the row now is; all threads must have finished reading the previous content (caller
// synchronises before).  A row that does not fit in shared memory (S up to 65535 is legal) stays
// in global memory: the compare loop reads it through L1/L2 instead.
__device__ __forceinline__ const uint8_t *stage_row(uint8_t *smem, const uint8_t *grow, size_t row, uint64_t *bar,
                                                    uint32_t &phase, int staged) {
    if (!staged) {
        __syncthreads();
        return grow;
    }
    if ((row & 15) == 0 && (((uintptr_t)grow) & 15) == 0) {
        fence_proxy_async();
        stage_query(smem, grow, (uint32_t)row, bar, phase);
        phase ^= 1;
    } else {
        for (uint32_t i = threadIdx.x; i < row; i += blockDim.x) smem[i] = grow[i];
    }
    __syncthreads();
    return smem;
}

// hnsw_rs search_layer: best-first search on one layer from `ep` (distance d_ep known), result in
// sh.ret (max-heap of at most ef), candidates in sh.cand; `vis` is reset here.
template <int ELEM, bool F32, bool RING = false>
__device__ void search_layer_dev(const GraphView &g, const uint8_t *smem_q, uint32_t ep, float d_ep,
                                 uint32_t ef, uint32_t layer, HnswShared &sh, Visit &vis,
                                 unsigned long long &neval, RowRing *rr) {
    vis.begin();
    __syncthreads();
    // thread 0 owns the heaps: pop the next candidate (or decide to stop) and publish it
    auto pop_next = [&]() {
        sh.done = 0;
        if (sh.cand.n == 0) {
            sh.done = 1;
        } else {
            const HItem c = sh.cand.pop();
            if (-c.d > sh.ret.at(0).d) sh.done = 1;
            sh.node = c.p;
            sh.next = sh.cand.n ? sh.cand.at(0).p : 0xFFFFFFFFu;  // the likely next pop
        }
    };
    if (threadIdx.x == 0) {
        sh.cand.n = 0;
        sh.ret.n = 0;
        vis.mark(ep);
        sh.cand.push(-d_ep, ep);
        sh.ret.push(d_ep, ep);
        neval += 1;  // the reference evaluates the distance to the layer's entry point again
        pop_next();
    }
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t cap = layer == 0 ? 2 * g.M : g.M;  // list capacity: entries beyond the count are stale but readable
#ifdef GSB_K7_PROF
    long long pt_g = 0, pt_e = 0, pt_h = 0, pt_n = 0, pt_rows = 0, pt0;
#endif
    for (;;) {
        __syncthreads();
        if (sh.done) break;
#ifdef GSB_K7_PROF
        pt0 = clock64();
#endif
        if (threadIdx.x < 32 && sh.next != 0xFFFFFFFFu) prefetch_list(g, sh.next, layer);
        // gather the unvisited neighbours of the popped node in list order; the count and the
        // entries are loaded together (one L2 round trip instead of two dependent ones)
        uint32_t len;
        const uint32_t *lst = list_of(g, sh.node, layer, len);
        uint32_t tot = 0;
        for (uint32_t base = 0; base < cap; base += blockDim.x) {  // one pass unless the CTA is narrower than the list
            const uint32_t i = base + threadIdx.x;
            const uint32_t nb = i < cap ? __ldg(&lst[i]) : 0u;
            const bool unv = i < len && !vis.seen(nb);
            const uint32_t bal = __ballot_sync(0xffffffffu, unv);
            if (lane == 0) sh.wcnt[warp] = __popc(bal);
            __syncthreads();
            uint32_t pre = tot;
            for (uint32_t w = 0; w < (blockDim.x >> 5); w++) {
                const uint32_t c = sh.wcnt[w];
                if (w < warp) pre += c;
                tot += c;
            }
            if (unv) {
                sh.E[pre + __popc(bal & ((1u << lane) - 1))] = nb;
                vis.mark(nb);
            }
            __syncthreads();
            if (base + blockDim.x >= len) break;  // nothing valid beyond the count
        }
        __syncthreads();
#ifdef GSB_K7_PROF
        pt_g += clock64() - pt0; pt0 = clock64();
#endif
        eval_list<ELEM, F32>(smem_q, g, sh.E, tot, sh.D, sh.acc);
        __syncthreads();
#ifdef GSB_K7_PROF
        pt_e += clock64() - pt0; pt0 = clock64(); pt_n++; pt_rows += tot;
#endif
        if (threadIdx.x == 0) {
            neval += tot;
            for (uint32_t i = 0; i < tot; i++) {
                const float ed = sh.D[i];
                if (ed < sh.ret.at(0).d || sh.ret.n < ef) {
                    sh.cand.push(-ed, sh.E[i]);
                    sh.ret.push(ed, sh.E[i]);
                    if (sh.ret.n > ef) (void)sh.ret.pop();
                }
            }
            pop_next();
#ifdef GSB_K7_PROF
            pt_h += clock64() - pt0;
#endif
        }
    }
#ifdef GSB_K7_PROF
    uint32_t smid_;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
    if (threadIdx.x == 0 && blockIdx.x % 37 == 0 && ef > 1)
        printf("k7prof cta %u sm %u: expansions %lld rows %lld  cycles/expansion: gather %lld eval %lld heap %lld\n", blockIdx.x, smid_,
               pt_n, pt_rows, pt_g / (pt_n ? pt_n : 1), pt_e / (pt_n ? pt_n : 1), pt_h / (pt_n ? pt_n : 1));
#endif
    __syncthreads();
}

#ifdef GSB_K7_PROF
__device__ __forceinline__ long long gsb_clk() {
    long long c;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
    return c;
}
#endif
// workers: the unvisited neighbours of `node` (layer 0) in list order -> E[0..tot), marked visited
__device__ __forceinline__ uint32_t gather_workers(const GraphView &g, uint32_t node, uint32_t *E, uint32_t *wcnt,
                                                   Visit &vis) {
    const uint32_t t = threadIdx.x - 32, warp = t >> 5, lane = lane_id();
    const uint32_t cap = 2 * g.M;
    uint32_t len;
    const uint32_t *lst = list_of(g, node, 0, len);
    uint32_t tot = 0;
    for (uint32_t base = 0; base < cap; base += kRingWorkers) {
        const uint32_t i = base + t;
        const uint32_t nb = i < cap ? __ldg(&lst[i]) : 0u;
        const bool unv = i < len && !vis.seen(nb);
        const uint32_t bal = __ballot_sync(0xffffffffu, unv);
        if (lane == 0) wcnt[warp] = __popc(bal);
        bar_workers();
        uint32_t pre = tot;
        for (uint32_t w = 0; w < kRingWorkers / 32; w++) {
            const uint32_t c = wcnt[w];
            if (w < warp) pre += c;
            tot += c;
        }
        if (unv) {
            E[pre + __popc(bal & ((1u << lane) - 1))] = nb;
            vis.mark(nb);
        }
        bar_workers();
        if (base + kRingWorkers >= len) break;  // nothing valid beyond the count
    }
    return tot;
}

// search_layer on layer 0 for the ring kernel, same results as search_layer_dev.  Per expansion the
// reference files the evaluated neighbours into the heaps and THEN pops the next candidate; here the
// control lane first PREDICTS that pop -- the current root of the candidate heap unless an evaluated
// neighbour is strictly closer (the first such in list order; an equal one stays below the root in
// Rust's sift-up) -- the workers gather and stream the predicted node while the control lane does
// the heap work, and the prediction is then compared with the real pop.  A wrong prediction (only
// when the heaps' tie handling or the result-heap threshold intervenes) is rolled back: the
// visited marks of the speculative gather are cleared and the real node is expanded.
template <int ELEM, bool F32>
__device__ void search_layer_ring(const GraphView &g, const uint8_t *q, uint32_t ep, float d_ep, uint32_t ef,
                                  HnswShared &sh, uint32_t *E2, float *D2, Visit &vis, unsigned long long &neval,
                                  RowRing &rr) {
    vis.begin();
    __syncthreads();
    const bool worker = threadIdx.x >= 32;
    const bool ctl = threadIdx.x == 0;
    auto pop_next = [&]() {
        sh.done = 0;
        if (sh.cand.n == 0) {
            sh.done = 1;
        } else {
            const HItem c = sh.cand.pop();
            if (-c.d > sh.ret.at(0).d) sh.done = 1;
            sh.node = c.p;
        }
    };
    if (ctl) {
        sh.cand.n = 0;
        sh.ret.n = 0;
        vis.mark(ep);
        sh.cand.push(-d_ep, ep);
        sh.ret.push(d_ep, ep);
        neval += 1;  // the reference evaluates the distance to the layer's entry point again
        pop_next();
    }
    __syncthreads();
    if (sh.done) return;
    uint32_t *Ec = sh.E, *En = E2;
    float *Dc = sh.D, *Dn = D2;
    if (worker) {
        const uint32_t tot = gather_workers(g, sh.node, Ec, sh.wcnt, vis);
        if (threadIdx.x == 32) sh.work = tot;
        eval_list_ring<ELEM, F32>(q, g, Ec, tot, Dc, sh.acc, rr);
    }
#ifdef GSB_K7_PROF
    long long pt_g = 0, pt_e = 0, pt_h = 0, pt_n = 0, pt_w1 = 0, pt_w2 = 0, pt_miss = 0, pt_p = 0, pt0;
#endif
    for (;;) {
#ifdef GSB_K7_PROF
        pt0 = gsb_clk();
#endif
        __syncthreads();  // D of the current expansion is complete, the heaps are at rest
#ifdef GSB_K7_PROF
        pt_w1 += gsb_clk() - pt0;
#endif
        const uint32_t tot = sh.work;
#ifdef GSB_K7_PROF
        const long long ptp = gsb_clk();
#endif
        if (ctl) {  // predict the next pop
            uint32_t pred = sh.cand.n ? sh.cand.at(0).p : 0xFFFFFFFFu;
            float best = sh.cand.n ? -sh.cand.at(0).d : 3.0e38f;
            const bool room = sh.ret.n + tot <= ef;   // every evaluated neighbour will be pushed
            const float top = sh.ret.at(0).d;
            for (uint32_t i = 0; i < tot; i++) {
                const float ed = Dc[i];
                if (ed < best && (room || ed < top)) {
                    best = ed;
                    pred = Ec[i];
                }
            }
            sh.next = pred;
        }
        __syncthreads();
#ifdef GSB_K7_PROF
        pt_p += gsb_clk() - ptp;
#endif
        const uint32_t pred = sh.next;
        uint32_t tot_n = 0;
        if (worker) {
            if (pred != 0xFFFFFFFFu) {
#ifdef GSB_K7_PROF
                pt0 = gsb_clk();
#endif
                tot_n = gather_workers(g, pred, En, sh.wcnt, vis);
#ifdef GSB_K7_PROF
                pt_g += gsb_clk() - pt0; pt0 = gsb_clk();
#endif
                eval_list_ring<ELEM, F32>(q, g, En, tot_n, Dn, sh.acc, rr);
#ifdef GSB_K7_PROF
                pt_e += gsb_clk() - pt0; pt_n++;
#endif
            }
        } else if (ctl) {
#ifdef GSB_K7_PROF
            pt0 = gsb_clk();
#endif
            neval += tot;
            for (uint32_t i = 0; i < tot; i++) {
                const float ed = Dc[i];
                if (ed < sh.ret.at(0).d || sh.ret.n < ef) {
                    sh.cand.push(-ed, Ec[i]);
                    sh.ret.push(ed, Ec[i]);
                    if (sh.ret.n > ef) (void)sh.ret.pop();
                }
            }
            pop_next();
#ifdef GSB_K7_PROF
            pt_h += gsb_clk() - pt0; pt_n++;
#endif
        }
#ifdef GSB_K7_PROF
        pt0 = gsb_clk();
#endif
        __syncthreads();
#ifdef GSB_K7_PROF
        pt_w2 += gsb_clk() - pt0;
        if (pred != sh.node) pt_miss++;
#endif
        if (sh.done) {
            // (a speculative gather left marks behind: the layer search is over, nobody reads them)
            break;
        }
        if (pred != sh.node) {  // rare: undo the speculation, expand the real node
            if (worker) {
                if (pred != 0xFFFFFFFFu)
                    for (uint32_t i = threadIdx.x - 32; i < tot_n; i += kRingWorkers) vis.unmark(En[i]);
                bar_workers();
                tot_n = gather_workers(g, sh.node, En, sh.wcnt, vis);
                eval_list_ring<ELEM, F32>(q, g, En, tot_n, Dn, sh.acc, rr);
            }
        }
        if (threadIdx.x == 32) sh.work = tot_n;
        uint32_t *te = Ec;
        Ec = En;
        En = te;
        float *td = Dc;
        Dc = Dn;
        Dn = td;
    }
#ifdef GSB_K7_PROF
    if ((threadIdx.x == 0 || threadIdx.x == 32) && blockIdx.x % 97 == 0 && pt_n)
        printf("k7prof cta %u thr %u: n %lld miss %lld  cycles/expansion: gather %lld eval %lld heap %lld wait_top %lld wait_end %lld predict+bar %lld\n",
               blockIdx.x, threadIdx.x, pt_n, pt_miss, pt_g / pt_n, pt_e / pt_n, pt_h / pt_n, pt_w1 / pt_n, pt_w2 / pt_n, pt_p / pt_n);
#endif
    __syncthreads();
}

// per-CTA workspace layout in global memory
struct WsLayout {
    size_t stride;     // bytes per CTA
    size_t off_ctr;    // uint32_t: this CTA's visit-stamp counter, kept across launches
    size_t off_cand;   // HItem[cand_cap]
    size_t off_stamp;  // uint32_t[ncap], zeroed when (re)allocated
    size_t off_ret;    // HItem[ef + 2] when the result heap does not fit in shared memory
    size_t off_newc;   // uint32_t[newc_cap] (insert: candidate extension)
};

// smem: [row: row bytes rounded to 128 (staged)] or [row ring (ring)] [ret heap: first ret_cs items]
// [visited bitmap: bm_words]
template <int ELEM, bool F32, bool RING>
__global__ void __launch_bounds__(kSearchThreads, 3)
k7_hnsw_search(GraphView g, const uint8_t *__restrict__ queries, uint32_t nq, uint32_t knbn, uint32_t ef,
               uint32_t ret_cs, uint32_t bm_words, int staged, uint8_t *__restrict__ ws, WsLayout wl,
               SearchOut so, uint32_t *__restrict__ qcounter) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t ring_bar[2 * kRingSlots];
    __shared__ HnswShared sh;
    __shared__ __align__(8) HItem cand_sm[kCandSmemK7];
    __shared__ uint32_t s_q;
    __shared__ uint32_t E2[RING ? kMaxList : 1];   // second expansion buffer (search_layer_ring)
    __shared__ float D2[RING ? kMaxList : 1];
    const size_t row = (size_t)g.S * ELEM;
    const size_t row128 = staged ? ((row + 127) & ~(size_t)127) : (RING ? (size_t)kRingBytes : 0);
    uint8_t *my = ws + (size_t)blockIdx.x * wl.stride;
    uint32_t *stamps = reinterpret_cast<uint32_t *>(my + wl.off_stamp);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        for (uint32_t s = 0; s < kRingSlots; s++) {
            mbar_init(&ring_bar[s], 1);
            mbar_init(&ring_bar[kRingSlots + s], kRingWorkers / 32);
        }
        fence_barrier_init();
        sh.cand.a = cand_sm;
        sh.cand.g = reinterpret_cast<HItem *>(my + wl.off_cand);
        sh.cand.cs = kCandSmemK7;
        sh.ret.a = reinterpret_cast<HItem *>(smem + row128);
        sh.ret.g = reinterpret_cast<HItem *>(my + wl.off_ret);
        sh.ret.cs = ret_cs;
    }
    __syncthreads();
    uint32_t phase = 0;
    RowRing rring;
    rring.buf = smem;
    rring.full = ring_bar;
    rring.empty = ring_bar + kRingSlots;
    rring.slot = 0;
    rring.par = 0;
    rring.pol_rows = l2_evict_first_policy();
    rring.pol_query = l2_evict_last_policy();
    RowRing *rr = &rring;
    Visit vis;
    vis.nwords = bm_words;
    vis.bits = bm_words ? reinterpret_cast<uint32_t *>(smem + row128 + (size_t)ret_cs * sizeof(HItem)) : nullptr;
    vis.stamps = stamps;
    vis.stamp = *reinterpret_cast<uint32_t *>(my + wl.off_ctr);
    for (;;) {
        if (threadIdx.x == 0) s_q = atomicAdd(qcounter, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= nq) break;
#ifdef GSB_K7_PROF
        const long long qt0 = clock64();
#endif
        const uint8_t *cur = stage_row(smem, queries + (size_t)q * row, row, &bar, phase, staged);
        unsigned long long neval = 0;
        uint32_t pivot = g.entry;
        if (threadIdx.x == 0) sh.E[0] = pivot;
        __syncthreads();
        if (RING) {
            if (threadIdx.x >= 32) eval_list_ring<ELEM, F32>(cur, g, sh.E, 1, sh.D, sh.acc, *rr);
        } else {
            eval_list<ELEM, F32>(cur, g, sh.E, 1, sh.D, sh.acc);
        }
        __syncthreads();
        float dist_to_entry = sh.D[0];
        neval += 1;
        // ---- one greedy hop per upper layer (hnsw_rs `search`)
        const int top = g.levels[g.entry];
        for (int layer = top; layer >= 1; layer--) {
            uint32_t len;
            const uint32_t *lst = list_of(g, pivot, (uint32_t)layer, len);
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) sh.E[i] = __ldg(&lst[i]);
            __syncthreads();
            if (RING) {
                if (threadIdx.x >= 32) eval_list_ring<ELEM, F32>(cur, g, sh.E, len, sh.D, sh.acc, *rr);
            } else {
                eval_list<ELEM, F32>(cur, g, sh.E, len, sh.D, sh.acc);
            }
            __syncthreads();
            neval += len;
            // every thread scans the same shared arrays: uniform result, no broadcast needed
            uint32_t newp = pivot;
            for (uint32_t i = 0; i < len; i++) {
                if (sh.D[i] < dist_to_entry) {
                    dist_to_entry = sh.D[i];
                    newp = sh.E[i];
                }
            }
            pivot = newp;
        }
        // ---- search_layer(q, pivot, ef, 0)
        if (RING) search_layer_ring<ELEM, F32>(g, cur, pivot, dist_to_entry, ef, sh, E2, D2, vis, neval, *rr);
        else search_layer_dev<ELEM, F32, false>(g, cur, pivot, dist_to_entry, ef, 0, sh, vis, neval, nullptr);
#ifdef GSB_K7_PROF
        const long long qt1 = clock64();
#endif
        // BinaryHeap::into_sorted_vec is a serial chain of ef sift-downs: run it on a contiguous
        // shared-memory copy of the heap (the row ring is idle now) instead of the hybrid one
        HItem *flat = nullptr;
        const uint32_t nret = sh.ret.n;
        if (RING && (size_t)nret * sizeof(HItem) <= kRingBytes) {
            flat = reinterpret_cast<HItem *>(smem);
            for (uint32_t i = threadIdx.x; i < nret; i += blockDim.x) flat[i] = sh.ret.at(i);
            __syncthreads();
        } else if (sh.ret.cs >= nret) {
            flat = sh.ret.a;
        }
        if (threadIdx.x == 0) {
            HeapT<false> fl;
            fl.a = flat;
            fl.n = nret;
            if (flat) fl.into_sorted();
            else sh.ret.into_sorted();
#ifdef GSB_K7_PROF
            if (blockIdx.x % 37 == 0)
                printf("k7prof cta %u query %u: search %lld cycles, into_sorted %lld cycles\n", blockIdx.x, q, qt1 - qt0,
                       clock64() - qt1);
#endif
            uint32_t last = knbn < ef ? knbn : ef;
            if (nret < last) last = nret;
            for (uint32_t i = 0; i < last; i++) {
                const HItem it = flat ? flat[i] : sh.ret.at(i);
                const uint32_t p = it.p;
                gsb_neighbour nbq;
                nbq.d_id = g.ids[p];
                nbq.distance = it.d;
                nbq.layer = g.levels[p];
                nbq.pad_[0] = nbq.pad_[1] = nbq.pad_[2] = 0;
                nbq.rank = (int32_t)g.ranks[p];
                so.out[(size_t)q * knbn + i] = nbq;
            }
            so.counts[q] = last;
            if (so.nb_eval) so.nb_eval[q] = neval;
        }
        if (RING) fence_proxy_async();  // the ring was written with ordinary stores: order them before the next bulk copies
        __syncthreads();
    }
    if (threadIdx.x == 0) *reinterpret_cast<uint32_t *>(my + wl.off_ctr) = vis.stamp;
}

// =========================================================================== construction
struct WaveView {
    uint32_t first, W;       // the wave = points [first, first + W)
    uint32_t t_begin, t_end; // phase A runs the points [t_begin, t_end) of the wave on this GPU (all of
                             // them on one GPU; a rank's slice when the insertion is sharded)
    uint32_t entry;          // entry point before the wave
    uint32_t ef_c, extend;   // ef_construction, extend_candidates (layer 0 only)
    uint32_t *sel_n;         // [W x kMaxLayers] selected neighbours per layer
    uint32_t *sel_idx;       // [W x 18M] layer 0 at 0 (2M entries), layer l >= 1 at 2M + (l-1) M
    float *sel_d;
    uint32_t *counter;       // work counter of the wave
};

__device__ __forceinline__ size_t sel_off(uint32_t M, uint32_t t, uint32_t l) {
    return (size_t)t * 18 * M + (l == 0 ? 0 : 2 * M + (size_t)(l - 1) * M);
}

// Phase A.  One CTA per new point: greedy descent, search_layer(ef_c) per layer, earlier points
// of the wave merged in, select_neighbours (Malkov heuristic, extension on layer 0), sort.
template <int ELEM, bool F32>
__global__ void __launch_bounds__(kInsertThreads, GSB_K8_PER_SM)
k8_hnsw_insert_select(GraphView g, WaveView wv, int ret_in_smem, uint32_t bm_words, int staged,
                      uint8_t *__restrict__ ws, WsLayout wl) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ HnswShared sh;
    __shared__ __align__(8) HItem cand_sm[kCandSmem];
    __shared__ uint32_t s_t, s_ep, s_nout, s_mode, s_nnew;
    __shared__ float s_dep;
    __shared__ uint32_t outP[kMaxList];
    __shared__ float outD[kMaxList];
    __shared__ float aD[kMaxList];  // selection: own distances of the candidates of the current chunk
    const size_t row = (size_t)g.S * ELEM;
    const size_t row128 = staged ? ((row + 127) & ~(size_t)127) : 0;
    const float fS = (float)g.S;
    uint8_t *my = ws + (size_t)blockIdx.x * wl.stride;
    uint32_t *stamps = reinterpret_cast<uint32_t *>(my + wl.off_stamp);
    uint32_t *newc = reinterpret_cast<uint32_t *>(my + wl.off_newc);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        sh.cand.a = cand_sm;
        sh.cand.g = reinterpret_cast<HItem *>(my + wl.off_cand);
        sh.cand.cs = kCandSmem;
        sh.ret.a = reinterpret_cast<HItem *>(smem + row128);
        sh.ret.g = reinterpret_cast<HItem *>(my + wl.off_ret);
        sh.ret.cs = ret_in_smem ? 0xFFFFFFFFu : 0u;
    }
    __syncthreads();
    uint32_t phase = 0;
    Visit vis;
    vis.nwords = bm_words;
    vis.bits = bm_words ? reinterpret_cast<uint32_t *>(smem + row128 +
                                                       (ret_in_smem ? ((size_t)wv.ef_c + 2) * sizeof(HItem) : 0))
                        : nullptr;
    vis.stamps = stamps;
    vis.stamp = *reinterpret_cast<uint32_t *>(my + wl.off_ctr);
    unsigned long long neval = 0;
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_t = wv.t_begin + atomicAdd(wv.counter, 1u);
        __syncthreads();
        const uint32_t t = s_t;
        if (t >= wv.t_end) break;
        const uint32_t np = wv.first + t;
        const uint32_t level = g.levels[np];
        const uint8_t *qrow = g.sigs + (size_t)np * row;
        if (threadIdx.x < kMaxLayers) wv.sel_n[(size_t)t * kMaxLayers + threadIdx.x] = 0;
        const uint8_t *cur = stage_row(smem, qrow, row, &bar, phase, staged);
        uint32_t ep = wv.entry;
        const uint32_t lmax = g.levels[wv.entry];
        if (threadIdx.x == 0) sh.E[0] = ep;
        __syncthreads();
        eval_list<ELEM, F32>(cur, g, sh.E, 1, sh.D, sh.acc);
        __syncthreads();
        float d_ep = sh.D[0];
        // ---- greedy descent through the layers above the point's level: search_layer(ef = 1)
        for (int l = (int)lmax; l >= (int)level + 1; l--) {
            search_layer_dev<ELEM, F32>(g, cur, ep, d_ep, 1, (uint32_t)l, sh, vis, neval, nullptr);
            if (threadIdx.x == 0) {
                s_ep = ep;
                s_dep = d_ep;
                if (sh.ret.n > 0) {
                    const HItem e = sh.ret.pop();
                    if (e.d < d_ep) {
                        s_ep = e.p;
                        s_dep = e.d;
                    }
                }
            }
            __syncthreads();
            ep = s_ep;
            d_ep = s_dep;
        }
        const int top = (int)(level < lmax ? level : lmax);
        for (int l = top; l >= 0; l--) {
            search_layer_dev<ELEM, F32>(g, cur, ep, d_ep, wv.ef_c, (uint32_t)l, sh, vis, neval, nullptr);
            // ---- earlier points of this wave, in order, as if search_layer had met them last (a wave
            // may be longer than the list scratch: kMaxList mates at a time, the heap sees them in order)
            for (uint32_t base = 0; wv.first + base < np;) {
                uint32_t tot = 0;
                for (; wv.first + base < np && tot + blockDim.x <= (uint32_t)kMaxList; base += blockDim.x) {
                    const uint32_t m = wv.first + base + threadIdx.x;
                    const bool on = m < np && g.levels[m] >= (uint32_t)l;
                    const uint32_t bal = __ballot_sync(0xffffffffu, on);
                    if (lane_id() == 0) sh.wcnt[warp] = __popc(bal);
                    __syncthreads();
                    uint32_t pre = tot;
                    for (uint32_t w = 0; w < (blockDim.x >> 5); w++) {
                        const uint32_t c = sh.wcnt[w];
                        if (w < warp) pre += c;
                        tot += c;
                    }
                    if (on) sh.E[pre + __popc(bal & ((1u << lane_id()) - 1))] = m;
                    __syncthreads();
                }
                eval_list<ELEM, F32>(cur, g, sh.E, tot, sh.D, sh.acc);
                __syncthreads();
                if (threadIdx.x == 0) {
                    for (uint32_t i = 0; i < tot; i++) {
                        const float ed = sh.D[i];
                        if (ed < sh.ret.at(0).d || sh.ret.n < wv.ef_c) {
                            sh.ret.push(ed, sh.E[i]);
                            if (sh.ret.n > wv.ef_c) (void)sh.ret.pop();
                        }
                    }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                // from_positive_binaryheap_to_negative_binary_heap: push in underlying-vec order
                sh.cand.n = 0;
                for (uint32_t i = 0; i < sh.ret.n; i++) sh.cand.push(-sh.ret.at(i).d, sh.ret.at(i).p);
            }
            __syncthreads();
            // ---- select_neighbours
            const uint32_t nb_asked = l == 0 ? 2 * g.M : g.M;
            const bool extend_asked = l == 0 && wv.extend != 0;
            if (threadIdx.x == 0) {
                s_nout = 0;
                s_mode = 1;  // heuristic
                if (sh.cand.n <= nb_asked) s_mode = extend_asked ? 2u : 0u;
                if (s_mode == 0) {  // few candidates, no extension: take them all, nearest first
                    while (sh.cand.n > 0) {
                        const HItem p = sh.cand.pop();
                        outP[s_nout] = p.p;
                        outD[s_nout] = -p.d;
                        s_nout++;
                    }
                }
            }
            __syncthreads();
            if (s_mode == 2) {  // extension: the neighbours of the candidates join them
                vis.begin();
                if (threadIdx.x == 0) {
                    const uint32_t n0 = sh.cand.n;
                    for (uint32_t i = 0; i < n0; i++) vis.mark(sh.cand.at(i).p);
                    uint32_t nnew = 0;
                    for (uint32_t i = 0; i < n0; i++) {
                        uint32_t len;
                        const uint32_t *lst = list_of(g, sh.cand.at(i).p, (uint32_t)l, len);
                        for (uint32_t j = 0; j < len; j++) {
                            const uint32_t e = __ldg(&lst[j]);
                            if (vis.seen(e)) continue;
                            vis.mark(e);
                            newc[nnew++] = e;
                        }
                    }
                    s_nnew = nnew;
                }
                __syncthreads();
            }
            if (s_mode == 2) {
                const uint32_t nnew = s_nnew;
                for (uint32_t c0 = 0; c0 < nnew; c0 += kMaxList) {
                    const uint32_t nc = nnew - c0 < (uint32_t)kMaxList ? nnew - c0 : (uint32_t)kMaxList;
                    __syncthreads();
                    for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) sh.E[i] = newc[c0 + i];
                    __syncthreads();
                    eval_list<ELEM, F32>(cur, g, sh.E, nc, sh.D, sh.acc);
                    __syncthreads();
                    if (threadIdx.x == 0)
                        for (uint32_t i = 0; i < nc; i++) sh.cand.push(-sh.D[i], sh.E[i]);
                }
                __syncthreads();
            }
            if (s_mode != 0) {
                // Malkov heuristic: candidates leave the heap nearest first; one is kept unless an
                // already selected point is at least as close to it as the new point is.  The pop
                // order does not depend on the decisions, so candidates are taken in chunks: the
                // chunk is tested against each selected point in turn (that point's row staged in
                // shared memory, the surviving candidates streamed against it like a search
                // expansion), then scanned in order -- the first survivor is selected and the
                // survivors after it are tested against it.  Same decisions as the one-by-one
                // loop of the reference, but whole batches of rows per staged row.
                for (;;) {
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        uint32_t nc = 0;
                        if (s_nout < nb_asked)
                            while (sh.cand.n > 0 && nc < blockDim.x && nc < (uint32_t)kMaxList) {
                                const HItem e = sh.cand.pop();
                                sh.E[nc] = e.p;
                                aD[nc] = -e.d;
                                nc++;
                            }
                        sh.work = nc;        // alive candidates, in pop order
                        sh.next = s_nout;    // selected points this chunk has not been tested against yet: [0, next)
                        sh.flag = 0;         // index of the first selected point still to test
                    }
                    __syncthreads();
                    if (sh.work == 0) break;
                    for (;;) {  // rounds: one staged selected row each
                        __syncthreads();
                        if (threadIdx.x == 0) {
                            sh.done = 0;
                            if (sh.work == 0) {
                                sh.done = 1;
                            } else if (sh.flag >= s_nout) {
                                // tested against everything selected so far: the first survivor is selected
                                if (sh.work == 0 || s_nout >= nb_asked) {
                                    sh.done = 1;
                                } else {
                                    outP[s_nout] = sh.E[0];
                                    outD[s_nout] = aD[0];
                                    s_nout++;
                                    sh.done = (sh.work == 1 || s_nout >= nb_asked) ? 1u : 2u;  // 2: drop entry 0, test the rest
                                }
                            }
                        }
                        __syncthreads();
                        if (sh.done == 1) break;
                        const uint32_t drop = sh.done == 2 ? 1u : 0u;     // entry 0 was just selected
                        const uint32_t nal = sh.work - drop;
                        const uint32_t si = sh.flag;                      // selected point to test against
                        cur = stage_row(smem, g.sigs + (size_t)outP[si] * row, row, &bar, phase, staged);
                        eval_list<ELEM, F32>(cur, g, sh.E + drop, nal, sh.D, sh.acc);
                        __syncthreads();
                        // ordered compaction of the survivors (distance to the selected point > own distance)
                        uint32_t e = 0;
                        float ed = 0.f;
                        bool keep = false;
                        if (threadIdx.x < nal) {
                            e = sh.E[drop + threadIdx.x];
                            ed = aD[drop + threadIdx.x];
                            keep = !(sh.D[threadIdx.x] <= ed);
                        }
                        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                        if (lane_id() == 0) sh.wcnt[warp] = __popc(bal);
                        __syncthreads();
                        uint32_t pre = 0, tot = 0;
                        for (uint32_t w = 0; w < (blockDim.x >> 5); w++) {
                            if (w < warp) pre += sh.wcnt[w];
                            tot += sh.wcnt[w];
                        }
                        if (keep) {
                            const uint32_t at = pre + __popc(bal & ((1u << lane_id()) - 1));
                            sh.E[at] = e;
                            aD[at] = ed;
                        }
                        if (threadIdx.x == 0) {
                            sh.work = tot;
                            sh.flag = si + 1;
                        }
                    }
                }
            }
            __syncthreads();
            if (l > 0 && s_mode != 0) cur = stage_row(smem, qrow, row, &bar, phase, staged);  // the heuristic replaced q
            // ---- sort by (distance, index) and publish; next entry = nearest selected point
            // that is not of this wave
            const uint32_t nout = s_nout;
            if (threadIdx.x == 0) {
                s_ep = ep;
                s_dep = d_ep;
                sh.work = 0xFFFFFFFFu;  // best rank of an old point
            }
            __syncthreads();
            const size_t so = sel_off(g.M, t, (uint32_t)l);
            for (uint32_t i = threadIdx.x; i < nout; i += blockDim.x) {
                const float d = outD[i];
                const uint32_t p = outP[i];
                uint32_t rank = 0;
                for (uint32_t j = 0; j < nout; j++) {
                    const float dj = outD[j];
                    const uint32_t pj = outP[j];
                    rank += (dj < d || (dj == d && pj < p)) ? 1u : 0u;
                }
                wv.sel_idx[so + rank] = p;
                wv.sel_d[so + rank] = d;
                if (p < wv.first) atomicMin(&sh.work, rank);
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < nout; i += blockDim.x) {
                const float d = outD[i];
                const uint32_t p = outP[i];
                uint32_t rank = 0;
                for (uint32_t j = 0; j < nout; j++) {
                    const float dj = outD[j];
                    const uint32_t pj = outP[j];
                    rank += (dj < d || (dj == d && pj < p)) ? 1u : 0u;
                }
                if (rank == sh.work) {
                    s_ep = p;
                    s_dep = d;
                }
            }
            if (threadIdx.x == 0) wv.sel_n[(size_t)t * kMaxLayers + l] = nout;
            __syncthreads();
            ep = s_ep;
            d_ep = s_dep;
        }
    }
    if (threadIdx.x == 0) *reinterpret_cast<uint32_t *>(my + wl.off_ctr) = vis.stamp;
}

// Phase B, step 1: the selections become the lists of the new points
__global__ void __launch_bounds__(256)
k9_write_own_lists(GraphView g, WaveView wv) {
    const uint32_t t = blockIdx.x;
    if (t >= wv.W) return;
    const uint32_t np = wv.first + t;
    const uint32_t level = g.levels[np];
    for (uint32_t l = 0; l <= level; l++) {
        const uint32_t ns = wv.sel_n[(size_t)t * kMaxLayers + l];
        const size_t so = sel_off(g.M, t, l);
        uint32_t *idx;
        float *dst;
        if (l == 0) {
            idx = g.nbr0 + (size_t)np * 2 * g.M;
            dst = g.dist0 + (size_t)np * 2 * g.M;
            if (threadIdx.x == 0) g.cnt0[np] = ns;
        } else {
            const uint32_t li = g.upper_off[np] + l - 1;
            idx = g.nbrU + (size_t)li * g.M;
            dst = g.distU + (size_t)li * g.M;
            if (threadIdx.x == 0) g.cntU[li] = ns;
        }
        for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) {
            idx[i] = wv.sel_idx[so + i];
            dst[i] = wv.sel_d[so + i];
        }
    }
}

// Phase B, step 2: reverse_update_neighborhood_simple.  One warp per arrival; the target list is
// locked, the new point inserted at its (distance, index) position, the farthest dropped when the
// list is full.  Keeping the M / 2M smallest of a totally ordered set does not depend on the
// order of arrivals, so the result equals the sequential one.
__global__ void __launch_bounds__(256)
k9_reverse_updates(GraphView g, WaveView wv) {
    const uint32_t t = blockIdx.x;
    if (t >= wv.W) return;
    const uint32_t np = wv.first + t;
    const uint32_t level = g.levels[np];
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = lane_id();
    for (uint32_t l = 0; l <= level; l++) {
        const uint32_t ns = wv.sel_n[(size_t)t * kMaxLayers + l];
        const size_t so = sel_off(g.M, t, l);
        const uint32_t thr = l == 0 ? 2 * g.M : g.M;
        for (uint32_t i = warp; i < ns; i += nwarps) {
            const uint32_t qp = wv.sel_idx[so + i];
            const float d = wv.sel_d[so + i];
            if (qp == np || l > g.levels[qp]) continue;
            volatile uint32_t *idx;
            volatile float *dst;
            volatile uint32_t *cnt;
            uint32_t *lock;
            if (l == 0) {
                idx = g.nbr0 + (size_t)qp * 2 * g.M;
                dst = g.dist0 + (size_t)qp * 2 * g.M;
                cnt = g.cnt0 + qp;
                lock = g.lock0 + qp;
            } else {
                const uint32_t li = g.upper_off[qp] + l - 1;
                idx = g.nbrU + (size_t)li * g.M;
                dst = g.distU + (size_t)li * g.M;
                cnt = g.cntU + li;
                lock = g.lockU + li;
            }
            if (lane == 0) {
                while (atomicCAS(lock, 0u, 1u) != 0u) {
                }
                __threadfence();
            }
            __syncwarp();
            const uint32_t n = *cnt;
            // this lane's entries (n <= 510 -> at most 16 per lane), and the insertion position
            uint32_t ri[16];
            float rd[16];
            uint32_t pos = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t j = lane + 32 * k;
                if (j < n) {
                    ri[k] = idx[j];
                    rd[k] = dst[j];
                    pos += (rd[k] < d || (rd[k] == d && ri[k] < np)) ? 1u : 0u;
                }
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, s);
            __syncwarp();
            if (!(n == thr && pos == n)) {  // otherwise the new point would be the one dropped
                const uint32_t nn = n == thr ? n : n + 1;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t j = lane + 32 * k;
                    if (j < n && j >= pos && j + 1 < nn) {
                        idx[j + 1] = ri[k];
                        dst[j + 1] = rd[k];
                    }
                }
                if (lane == 0) {
                    idx[pos] = np;
                    dst[pos] = d;
                    *cnt = nn;
                }
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicExch(lock, 0u);
            }
            __syncwarp();
        }
    }
}

}  // namespace gsb
