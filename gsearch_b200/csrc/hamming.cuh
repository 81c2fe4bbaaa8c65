// hamming.cuh -- K6: batched DistHamming (sm_100a).
//
// Replaces anndists::dist::DistHamming::eval [U] (used through Hnsw::<Sig,DistHamming>::new at
// src/dna/dnasketch.rs:139 and directly at src/bin/bindash.rs:94-95):
//     eval(a, b) = count(a[i] != b[i]) as f32 / len as f32
// Pure streaming compare: the query signature is staged once in shared memory with one TMA
// bulk copy (cp.async.bulk, SASS UBLKCP) and every candidate signature is read exactly once
// from HBM with 128-bit loads.  Algorithmic bytes per (query, candidate) = S * sizeof(Sig).
#pragma once

#include "common.cuh"

namespace gsb {

constexpr int kHamThreads = 512;

// stage `bytes` (multiple of 16, 16-byte aligned source) into shared memory with TMA bulk
// copies of at most 64 KiB each; all threads return after the data has landed
__device__ __forceinline__ void stage_query(uint8_t *smem_q, const uint8_t *gq, uint32_t bytes,
                                            uint64_t *bar, uint32_t phase) {
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        for (uint32_t off = 0; off < bytes; off += 65536u) {
            const uint32_t n = bytes - off < 65536u ? bytes - off : 65536u;
            tma_bulk_g2s(smem_q + off, gq + off, n, bar);
        }
    }
    mbar_wait(bar, phase);
}

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// number of differing elements between two 16-byte vectors, by element width
template <int ELEM, bool IS_F32>
__device__ __forceinline__ uint32_t diff16(const uint4 a, const uint4 b) {
    if (ELEM == 8) {
        return (uint32_t)(((a.x ^ b.x) | (a.y ^ b.y)) != 0u) + (uint32_t)(((a.z ^ b.z) | (a.w ^ b.w)) != 0u);
    } else if (ELEM == 4) {
        if (IS_F32) {
            return (uint32_t)(__uint_as_float(a.x) != __uint_as_float(b.x)) +
                   (uint32_t)(__uint_as_float(a.y) != __uint_as_float(b.y)) +
                   (uint32_t)(__uint_as_float(a.z) != __uint_as_float(b.z)) +
                   (uint32_t)(__uint_as_float(a.w) != __uint_as_float(b.w));
        }
        return (uint32_t)(a.x != b.x) + (uint32_t)(a.y != b.y) + (uint32_t)(a.z != b.z) +
               (uint32_t)(a.w != b.w);
    } else {  // 2-byte elements: per-halfword compare
        uint32_t n = 0;
        const uint32_t x[4] = {a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w};
#pragma unroll
        for (int i = 0; i < 4; i++) n += (uint32_t)((x[i] & 0xFFFFu) != 0u) + (uint32_t)((x[i] >> 16) != 0u);
        return n;
    }
}

// One warp counts the differing elements of the 16-byte vectors [vbeg, vend) of a candidate
// row against the query in shared memory: 8 independent 128-bit loads in flight per lane
// (a warp alone is latency bound: bytes in flight per warp set its bandwidth).
template <int ELEM, bool IS_F32>
__device__ __forceinline__ uint32_t warp_vec_count(const uint4 *qv, const uint4 *__restrict__ cv, uint32_t vbeg,
                                                   uint32_t vend) {
    const uint32_t lane = lane_id();
    uint32_t cnt = 0;
    uint32_t i = vbeg + lane;
    // software pipeline: two groups of four loads; while one group is compared the other is in
    // flight, so the warp always has 4-8 requests outstanding instead of bursts of 8 then none
    if (i + 224 < vend) {
        uint4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) a[u] = ldg_stream(cv + i + 32 * u);
        for (; i + 224 < vend; i += 256) {
#pragma unroll
            for (int u = 0; u < 4; u++) b[u] = ldg_stream(cv + i + 128 + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; u++) cnt += diff16<ELEM, IS_F32>(qv[i + 32 * u], a[u]);
            if (i + 256 + 96 < vend) {
#pragma unroll
                for (int u = 0; u < 4; u++) a[u] = ldg_stream(cv + i + 256 + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) cnt += diff16<ELEM, IS_F32>(qv[i + 128 + 32 * u], b[u]);
        }
        // group `a` of the next iteration may already be loaded: consume it
        if (i + 96 < vend) {
#pragma unroll
            for (int u = 0; u < 4; u++) cnt += diff16<ELEM, IS_F32>(qv[i + 32 * u], a[u]);
            i += 128;
        }
    }
    for (; i + 96 < vend; i += 128) {
        const uint4 c0 = ldg_stream(cv + i), c1 = ldg_stream(cv + i + 32), c2 = ldg_stream(cv + i + 64),
                    c3 = ldg_stream(cv + i + 96);
        cnt += diff16<ELEM, IS_F32>(qv[i], c0) + diff16<ELEM, IS_F32>(qv[i + 32], c1) +
               diff16<ELEM, IS_F32>(qv[i + 64], c2) + diff16<ELEM, IS_F32>(qv[i + 96], c3);
    }
    for (; i < vend; i += 32) cnt += diff16<ELEM, IS_F32>(qv[i], ldg_stream(cv + i));
    return cnt;
}

// scalar tail of a row whose byte length is not a multiple of 16 (per-lane partial count)
template <int ELEM, bool IS_F32>
__device__ __forceinline__ uint32_t warp_tail_count(const uint8_t *smem_q, const uint8_t *__restrict__ c,
                                                    uint32_t S) {
    const uint32_t done = ((S * ELEM) / 16) * (16 / ELEM);
    uint32_t cnt = 0;
    for (uint32_t e = done + lane_id(); e < S; e += 32) {
        if (ELEM == 8) cnt += ((const uint64_t *)smem_q)[e] != ((const uint64_t *)c)[e];
        else if (ELEM == 4) {
            if (IS_F32) cnt += ((const float *)smem_q)[e] != ((const float *)c)[e];
            else cnt += ((const uint32_t *)smem_q)[e] != ((const uint32_t *)c)[e];
        } else cnt += ((const uint16_t *)smem_q)[e] != ((const uint16_t *)c)[e];
    }
    return cnt;
}

// One warp computes count(q != c) for one candidate row; q in shared memory.
// row_bytes is a multiple of 16 in the fast path; the scalar tail handles the rest.
template <int ELEM, bool IS_F32>
__device__ __forceinline__ uint32_t warp_row_count(const uint8_t *smem_q, const uint8_t *__restrict__ c,
                                                   uint32_t S) {
    const uint32_t nvec = (S * ELEM) / 16;
    uint32_t cnt = warp_vec_count<ELEM, IS_F32>(reinterpret_cast<const uint4 *>(smem_q),
                                                reinterpret_cast<const uint4 *>(c), 0, nvec);
    cnt += warp_tail_count<ELEM, IS_F32>(smem_q, c, S);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    return cnt;
}

// grid = (ceil(n / cands_per_cta), nq); one query per CTA row, staged in shared memory
template <int ELEM, bool IS_F32>
__global__ void __launch_bounds__(kHamThreads)
k6_hamming_matrix(const uint8_t *__restrict__ queries, uint32_t nq, const uint8_t *__restrict__ cands,
                  uint32_t n, uint32_t S, uint32_t cands_per_cta, float *__restrict__ out, int staged) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t qi = blockIdx.y;
    const size_t row = (size_t)S * ELEM;
    const uint32_t row16 = (uint32_t)((row + 15) & ~(size_t)15);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const uint8_t *gq = queries + (size_t)qi * row;
    const uint8_t *qrow = smem;  // where the query row is read from
    if (!staged) {
        qrow = gq;  // the row does not fit in shared memory (S up to 65535 is legal): read it through L1/L2
    } else if ((row & 15) == 0 && (((uintptr_t)gq) & 15) == 0) {
        stage_query(smem, gq, (uint32_t)row, &bar, 0);
    } else {
        for (uint32_t i = threadIdx.x; i < row; i += blockDim.x) smem[i] = gq[i];
        __syncthreads();
    }
    (void)row16;
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t c0 = blockIdx.x * cands_per_cta;
    const uint32_t c1 = c0 + cands_per_cta < n ? c0 + cands_per_cta : n;
    const float fS = (float)S;
    for (uint32_t c = c0 + warp; c < c1; c += nwarps) {
        const uint8_t *cr = cands + (size_t)c * row;
        uint32_t cnt;
        if (((((uintptr_t)cr) | ((uintptr_t)qrow)) & 15) == 0) {
            cnt = warp_row_count<ELEM, IS_F32>(qrow, cr, S);
        } else {  // unaligned rows (row size not a multiple of 16): element-wise
            cnt = 0;
            for (uint32_t e = lane_id(); e < S; e += 32) {
                if (ELEM == 8) cnt += ((const uint64_t *)qrow)[e] != ((const uint64_t *)cr)[e];
                else if (ELEM == 4) {
                    if (IS_F32) cnt += ((const float *)qrow)[e] != ((const float *)cr)[e];
                    else cnt += ((const uint32_t *)qrow)[e] != ((const uint32_t *)cr)[e];
                } else cnt += ((const uint16_t *)qrow)[e] != ((const uint16_t *)cr)[e];
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        }
        if (lane_id() == 0) out[(size_t)qi * n + c] = __fdiv_rn((float)cnt, fS);
    }
}

}  // namespace gsb
