// Micro-benchmark: random 32-bit atomics / loads into a table, to size the K2 design.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench_atomics.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
// MODE 0: dependent CAS chain (next address waits for the result)   1: B independent CAS per step
// MODE 2: RED (atomicMin, no return)   3: independent loads   4: one atomicOr + one load per step (2 ops)
template <int MODE, int B>
__global__ void k(uint32_t *t, uint32_t cap, int iters, uint32_t *sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    for (int i = 0; i < iters; i += B) {
        uint32_t old[B];
#pragma unroll
        for (int b = 0; b < B; b++) {
            uint32_t h = mix(tid * 0x9e3779b9u + (uint32_t)(i + b) + (MODE == 0 ? (acc & 1) : 0));
            uint32_t idx = __umulhi(h, cap);
            if (MODE <= 1) old[b] = atomicCAS(&t[idx], 0u, h | 1u);
            else if (MODE == 2) { atomicMin(&t[idx], h); old[b] = 0; }
            else if (MODE == 3) old[b] = __ldcg(&t[idx]);
            else {
                old[b] = atomicOr(&t[idx], h | 1u);
                old[b] += __ldcg(&t[__umulhi(mix(h + 77u), cap)]);
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) acc += old[b];
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int MODE, int B>
void run(const char *name, uint32_t *t, size_t cap, uint32_t *sink) {
    const int threads = 256, blocks = 148 * 8, iters = 256;
    cudaMemset(t, 0, cap * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, B><<<blocks, threads>>>(t, (uint32_t)cap, 8, sink);
    cudaMemset(t, 0, cap * 4);
    cudaEventRecord(a);
    k<MODE, B><<<blocks, threads>>>(t, (uint32_t)cap, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * iters * (MODE == 4 ? 2 : 1);
    printf("%-28s table %7.1f MB : %8.2f Gops/s  (%.3f ms)\n", name, cap * 4 / 1e6, ops / ms / 1e6, ms);
}
int main() {
    uint32_t *t, *sink; size_t maxcap = 1u << 28;
    cudaMalloc(&t, maxcap * 4); cudaMalloc(&sink, 4);
    for (size_t cap : {size_t(2) << 20, size_t(8) << 20, size_t(32) << 20, size_t(256) << 20}) {
        run<0, 1>("CAS dependent chain", t, cap, sink);
        run<1, 1>("CAS independent B=1", t, cap, sink);
        run<1, 4>("CAS independent B=4", t, cap, sink);
        run<1, 8>("CAS independent B=8", t, cap, sink);
        run<2, 4>("RED min B=4", t, cap, sink);
        run<3, 4>("LDG.cg B=4", t, cap, sink);
        run<3, 8>("LDG.cg B=8", t, cap, sink);
        run<4, 8>("atomicOr + LDG.cg B=8", t, cap, sink);
    }
    return 0;
}
