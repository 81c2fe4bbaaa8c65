// Micro-benchmark: shared-memory scatter primitives on random addresses, to size the
// partition / count kernels of the prob sketch path (round 2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_smem scripts/ubench_smem.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
constexpr int kTab = 8192;
// MODE 0 ATOMS.ADD returning | 1 ATOMS.ADD result unused | 2 ATOMS.CAS returning | 3 STS + LDS (plain)
// MODE 4 match_any on 10 bits | 5 ATOMS.OR returning | 6 LDS only | 7 STS ; bar ; LDS (claim round, 8 per bar)
// MODE 8 address generation only (the baseline to subtract)
template <int MODE, int B>
__global__ void __launch_bounds__(256) k(int iters, uint32_t *sink) {
    __shared__ uint32_t tab[kTab];
    for (int i = threadIdx.x; i < kTab; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    for (int i = 0; i < iters; i += B) {
        uint32_t idx[B], old[B];
#pragma unroll
        for (int b = 0; b < B; b++) idx[b] = mix(tid * 0x9e3779b9u + (uint32_t)(i + b)) & (kTab - 1);
#pragma unroll
        for (int b = 0; b < B; b++) {
            if (MODE == 0) old[b] = atomicAdd(&tab[idx[b]], 1u);
            else if (MODE == 1) { atomicAdd(&tab[idx[b]], 1u); old[b] = 0; }
            else if (MODE == 2) old[b] = atomicCAS(&tab[idx[b]], 0u, tid | 1u);
            else if (MODE == 3) { tab[idx[b]] = tid; old[b] = 0; }
            else if (MODE == 4) old[b] = __match_any_sync(0xffffffffu, idx[b] & 1023u);
            else if (MODE == 5) old[b] = atomicOr(&tab[idx[b]], 1u << (tid & 31));
            else if (MODE == 6) old[b] = tab[idx[b]];
            else if (MODE == 7) { tab[idx[b]] = tid; old[b] = 0; }
            else old[b] = idx[b];
        }
        if (MODE == 7) __syncthreads();
        if (MODE == 3 || MODE == 7) {
#pragma unroll
            for (int b = 0; b < B; b++) old[b] = tab[idx[b] ^ (MODE == 3 ? 1 : 0)];
        }
        if (MODE == 7) __syncthreads();
#pragma unroll
        for (int b = 0; b < B; b++) acc += old[b];
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int MODE, int B>
void run(const char *name, uint32_t *sink, int ops_per_iter = 1) {
    const int threads = 256, blocks = 148 * 6, iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, B><<<blocks, threads>>>(64, sink);
    cudaEventRecord(a);
    k<MODE, B><<<blocks, threads>>>(iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * iters;
    int dev, mhz; cudaGetDevice(&dev); cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
    printf("%-34s B=%d : %8.2f G steps/s  (%.3f ms)  = %.3f steps/clk/SM at %d MHz  [%d smem ops per step]\n", name, B,
           ops / ms / 1e6, ms, ops / (ms * 1e-3) / 148.0 / (mhz * 1e3), mhz / 1000, ops_per_iter);
}
int main() {
    uint32_t *sink; cudaMalloc(&sink, 4);
    run<8, 8>("address generation only", sink, 0);
    run<0, 8>("ATOMS.ADD returning", sink);
    run<1, 8>("ATOMS.ADD result unused", sink);
    run<2, 8>("ATOMS.CAS returning", sink);
    run<5, 8>("ATOMS.OR returning", sink);
    run<6, 8>("LDS", sink);
    run<3, 8>("STS + LDS", sink, 2);
    run<7, 8>("STS ; bar ; LDS ; bar (8 per bar)", sink, 2);
    run<7, 4>("STS ; bar ; LDS ; bar (4 per bar)", sink, 2);
    run<4, 8>("match_any 10 bits", sink, 0);
    return 0;
}
