// ubench_rowstream.cu -- what bandwidth does K7's access pattern get on this GPU?  `nctas` CTAs of 256
// threads each read random 144 KB rows of a 7.2 GB table, (a) through a TMA ring (cp.async.bulk),
// (b) the same plus a 144 KB "query" re-read from L2 beside every row and compared, (c) register
// loads (8 x 16 B in flight per lane), (d) TMA ring but consecutive rows.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench_rowstream scripts/ubench_rowstream.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_g2s(void *dst, const void *src, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tma_g2s_h(void *dst, const void *src, uint32_t n, uint64_t *b, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

constexpr uint32_t ROW = 144000;

template <int R, int CH, int MODE, bool HINT = false>  // MODE 0: ring only, 1: ring + query compare, 3: consecutive rows
__global__ void __launch_bounds__(256, 3) ring_kernel(const uint8_t *tab, uint32_t nrows, const uint8_t *queries, uint32_t rows_per_cta, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t full[R], empty[R];
    const uint32_t t = threadIdx.x, lane = t & 31;
    uint64_t pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    if (t == 0) {
        for (int s = 0; s < R; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t nch = (ROW + CH - 1) / CH, J = rows_per_cta * nch;
    const uint4 *q4 = reinterpret_cast<const uint4 *>(queries + (size_t)blockIdx.x * ROW);
    auto row_of = [&](uint32_t i) { return MODE == 3 ? (blockIdx.x * rows_per_cta + i) % nrows : mix(blockIdx.x * 7919u + i) % nrows; };
    uint32_t pj = 0, pr = 0, pc = 0, pslot = 0;
    auto issue = [&]() {
        const uint32_t off = pc * CH, bytes = ROW - off < CH ? ROW - off : CH;
        mbar_expect_tx(&full[pslot], MODE == 4 ? 2 * bytes : bytes);
        if (HINT) {
            tma_g2s_h(buf + pslot * CH, tab + (size_t)row_of(pr) * ROW + off, bytes, &full[pslot], pol_first);
            if (MODE == 4) tma_g2s_h(buf + (R + pslot) * CH, queries + (size_t)blockIdx.x * ROW + off, bytes, &full[pslot], pol_last);
        } else {
            tma_g2s(buf + pslot * CH, tab + (size_t)row_of(pr) * ROW + off, bytes, &full[pslot]);
            if (MODE == 4) tma_g2s(buf + (R + pslot) * CH, queries + (size_t)blockIdx.x * ROW + off, bytes, &full[pslot]);
        }
        if (++pslot == R) pslot = 0;
        if (++pc == nch) { pc = 0; pr++; }
        pj++;
    };
    if (t == 0) while (pj < J && pj < R) issue();
    uint32_t slot = 0, par = 0, c = 0, cnt = 0;
    constexpr int PER = CH / 16 / 256;
    for (uint32_t j = 0; j < J; j++) {
        const uint32_t off = c * CH, nv = (ROW - off < CH ? ROW - off : CH) / 16;
        uint4 qv[PER];
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < PER; u++) { const uint32_t v = c * (CH / 16) + t + 256 * u; qv[u] = v < ROW / 16 ? __ldg(q4 + v) : make_uint4(0, 0, 0, 0); }
        }
        mbar_wait(&full[slot], par);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(buf + slot * CH);
        const uint4 *sq = reinterpret_cast<const uint4 *>(buf + (MODE == 4 ? (R + slot) : R) * CH);
#pragma unroll
        for (int u = 0; u < PER; u++) {
            if (t + 256 * u < nv) {
                const uint4 a = s4[t + 256 * u];
                if (MODE == 4 || MODE == 5) { const uint4 b = sq[t + 256 * u]; cnt += (a.x != b.x || a.y != b.y) + (a.z != b.z || a.w != b.w); }
                else if (MODE == 1) cnt += (a.x != qv[u].x || a.y != qv[u].y) + (a.z != qv[u].z || a.w != qv[u].w);
                else cnt += a.x == 12345u;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (t == 0 && pj < J) { mbar_wait(&empty[slot], par); issue(); }
        if (++c == nch) c = 0;
        if (++slot == R) { slot = 0; par ^= 1; }
    }
    if (cnt == 0xFFFFFFFFu) sink[0] = cnt;
}

__global__ void __launch_bounds__(256, 3) ldg_kernel(const uint8_t *tab, uint32_t nrows, uint32_t rows_per_cta, unsigned long long *sink) {
    const uint32_t t = threadIdx.x;
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < rows_per_cta; i++) {
        const uint4 *r4 = reinterpret_cast<const uint4 *>(tab + (size_t)(mix(blockIdx.x * 7919u + i) % nrows) * ROW);
        for (uint32_t v = t; v < ROW / 16; v += 256 * 8) {
            uint4 a[8];
#pragma unroll
            for (int u = 0; u < 8; u++) if (v + 256 * u < ROW / 16) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[u].x), "=r"(a[u].y), "=r"(a[u].z), "=r"(a[u].w) : "l"(r4 + v + 256 * u)); else a[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 8; u++) cnt += a[u].x == 12345u;
        }
    }
    if (cnt == 0xFFFFFFFFu) sink[0] = cnt;
}

template <typename F>
static void timeit(const char *name, uint32_t nctas, uint32_t rows, double extra, F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)nctas * rows * ROW;
    printf("%-44s ctas=%4u  %8.1f GB/s of rows (%.2f of 6451)%s\n", name, nctas, bytes / ms / 1e6, bytes / ms / 1e6 / 6451.5, extra > 0 ? "  + the same again from L2" : "");
    fflush(stdout);
}

int main() {
    const uint32_t nrows = 50000, rows = 400;
    uint8_t *tab, *q;
    unsigned long long *sink;
    CK(cudaMalloc(&tab, (size_t)nrows * ROW));
    CK(cudaMemset(tab, 1, (size_t)nrows * ROW));
    CK(cudaMalloc(&q, (size_t)1024 * ROW));
    CK(cudaMemset(q, 1, (size_t)1024 * ROW));
    CK(cudaMalloc(&sink, 8));
#define RINGH(R, CH, MODE, NAME)                                                                                    \
    for (uint32_t nctas : {296u, 370u, 444u}) {                                                                     \
        const int sm = (MODE == 4 ? 2 * R : (MODE == 5 ? R + 1 : R)) * CH;                                          \
        CK(cudaFuncSetAttribute(ring_kernel<R, CH, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));  \
        timeit(NAME, nctas, rows, 1, [&]() { ring_kernel<R, CH, MODE, true><<<nctas, 256, sm>>>(tab, nrows, q, rows, sink); }); \
    }
#define RING(R, CH, MODE, NAME)                                                                                     \
    for (uint32_t nctas : {148u, 296u, 444u}) {                                                                     \
        const int sm = (MODE == 4 ? 2 * R : (MODE == 5 ? R + 1 : R)) * CH;                                          \
        CK(cudaFuncSetAttribute(ring_kernel<R, CH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));        \
        timeit(NAME, nctas, rows, MODE == 1 || MODE == 4, [&]() { ring_kernel<R, CH, MODE><<<nctas, 256, sm>>>(tab, nrows, q, rows, sink); }); \
    }
    RINGH(3, 8192, 4, "hints; TMA ring 3 x (8+8) KB")
    RINGH(2, 8192, 4, "hints; TMA ring 2 x (8+8) KB")
    RINGH(4, 4096, 4, "hints; TMA ring 4 x (4+4) KB")
    RINGH(6, 4096, 4, "hints; TMA ring 6 x (4+4) KB")
    RINGH(4, 8192, 4, "hints; TMA ring 4 x (8+8) KB")

    return 0;
}
