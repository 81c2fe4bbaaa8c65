// api_core.cu -- error plumbing, device probing and the distance entry points of the C ABI.
#include <string.h>

#include <new>

#include <math.h>
#include <stdlib.h>

#include "api_common.h"
#include "hamming.cuh"

namespace gsb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); libgsearch_b200 has no CPU path",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return GSB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device ordinal %d out of range (0..%d)", device, n - 1);
        return GSB_ERR_INVALID_ARG;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
        return GSB_ERR_CUDA;
    }
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                  prop.minor);
        return GSB_ERR_NO_DEVICE;
    }
    return GSB_OK;
}

static uint32_t elem_of(uint32_t sig_type) {
    switch (sig_type) {
    case GSB_SIG_U64: return 8;
    case GSB_SIG_U16: return 2;
    case GSB_SIG_U32:
    case GSB_SIG_F32: return 4;
    default: return 0;
    }
}

template <int ELEM, bool F32>
static int launch_hamming(const uint8_t *q, uint32_t nq, const uint8_t *c, uint32_t n, uint32_t S, float *out,
                          cudaStream_t st) {
    const size_t row = (size_t)S * ELEM;
    size_t smem = (row + 127) & ~(size_t)127;
    const int staged = smem <= 200 * 1024 ? 1 : 0;  // larger rows are read from global memory
    if (!staged) smem = 0;
    // the attribute is per device: set it on every launch (a process may use several GPUs)
    GSB_CUDA_TRY(cudaFuncSetAttribute(k6_hamming_matrix<ELEM, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      200 * 1024));
    // enough CTAs per query to fill the machine, but at least one candidate per warp
    uint32_t per = 8;
    while ((uint64_t)nq * ((n + per - 1) / per) > 148ull * 16 && per < 4096) per *= 2;
    dim3 grid((n + per - 1) / per, nq);
    k6_hamming_matrix<ELEM, F32><<<grid, kHamThreads, smem, st>>>(q, nq, c, n, S, per, out, staged);
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB_OK;
}

int hamming_matrix_dev(const void *dq, uint32_t nq, const void *dc, uint32_t n, uint32_t S, uint32_t sig_type,
                       float *dout, cudaStream_t st) {
    if (nq == 0 || n == 0) return GSB_OK;
    if (S == 0) {
        set_error("S must be > 0");
        return GSB_ERR_INVALID_ARG;
    }
    const uint8_t *q = (const uint8_t *)dq, *c = (const uint8_t *)dc;
    switch (sig_type) {
    case GSB_SIG_U64: return launch_hamming<8, false>(q, nq, c, n, S, dout, st);
    case GSB_SIG_U32: return launch_hamming<4, false>(q, nq, c, n, S, dout, st);
    case GSB_SIG_F32: return launch_hamming<4, true>(q, nq, c, n, S, dout, st);
    case GSB_SIG_U16: return launch_hamming<2, false>(q, nq, c, n, S, dout, st);
    default: set_error("unknown sig_type %u", sig_type); return GSB_ERR_INVALID_ARG;
    }
}

}  // namespace gsb

using namespace gsb;

extern "C" const char *gsb_last_error(void) { return g_err; }
extern "C" const char *gsb_version(void) { return "gsearch_b200 0.1.0 (sm_100a)"; }
extern "C" int gsb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int gsb_hamming_matrix_dev(const void *d_queries, uint32_t nq, const void *d_cands, uint32_t n,
                                      uint32_t S, uint32_t sig_type, float *d_out, void *stream) {
    if ((nq && !d_queries) || (n && !d_cands) || (nq && n && !d_out)) {
        set_error("gsb_hamming_matrix_dev: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    return hamming_matrix_dev(d_queries, nq, d_cands, n, S, sig_type, d_out, (cudaStream_t)stream);
}

extern "C" int gsb_hamming_matrix(const void *queries, uint32_t nq, const void *cands, uint32_t n, uint32_t S,
                                  uint32_t sig_type, float *out, int device) {
    if ((nq && !queries) || (n && !cands) || (nq && n && !out)) {
        set_error("gsb_hamming_matrix: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    const uint32_t es = elem_of(sig_type);
    if (!es) {
        set_error("unknown sig_type %u", sig_type);
        return GSB_ERR_INVALID_ARG;
    }
    int rc = check_device(device);
    if (rc) return rc;
    if (nq == 0 || n == 0) return GSB_OK;
    GSB_CUDA_TRY(cudaSetDevice(device));
    const size_t row = (size_t)S * es;
    void *dq = nullptr, *dc = nullptr;
    float *dout = nullptr;
    cudaStream_t st = nullptr;
    GSB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaError_t e1 = cudaMalloc(&dq, row * nq), e2 = cudaMalloc(&dc, row * n),
                e3 = cudaMalloc((void **)&dout, sizeof(float) * (size_t)nq * n);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaFree(dq);
        cudaFree(dc);
        cudaFree(dout);
        cudaStreamDestroy(st);
        set_error("cudaMalloc failed in gsb_hamming_matrix");
        return GSB_ERR_OOM;
    }
    rc = GSB_OK;
    cudaError_t e = cudaMemcpyAsync(dq, queries, row * nq, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dc, cands, row * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = hamming_matrix_dev(dq, nq, dc, n, S, sig_type, dout, st);
    if (e == cudaSuccess && rc == GSB_OK)
        e = cudaMemcpyAsync(out, dout, sizeof(float) * (size_t)nq * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dq);
    cudaFree(dc);
    cudaFree(dout);
    cudaStreamDestroy(st);
    if (e != cudaSuccess) {
        set_error("CUDA failure in gsb_hamming_matrix: %s", cudaGetErrorString(e));
        return GSB_ERR_CUDA;
    }
    return rc;
}

extern "C" int gsb_hamming_batch(const void *q, const void *cands, uint32_t n, uint32_t S, uint32_t sig_type,
                                 float *out, int device) {
    return gsb_hamming_matrix(q, 1, cands, n, S, sig_type, out, device);
}

// ---- scalar Distance::eval-shaped exports (anndists DistCFFI plug point [U]: `extern "C" fn(*const T,
// *const T, len: u64) -> f32`).  A convenience: ONE pair per call, both rows travel to
// device 0 (GSB_DEVICE overrides) and one K6 launch evaluates them -- correct and bit-identical to
// the batched forms, but latency bound; anything hot should use gsb_hamming_matrix / the index.
// Errors cannot be returned through this signature: the result is NaN and gsb_last_error() says why.
static float dist_hamming_scalar(const void *a, const void *b, unsigned long long len, uint32_t sig_type) {
    if (!a || !b || len == 0 || len > 0xFFFFFFFFull) {
        set_error("gsb_dist_hamming: NULL argument or bad length %llu", len);
        return nanf("");
    }
    static int device = -1;
    if (device < 0) {
        const char *e = getenv("GSB_DEVICE");
        device = e ? atoi(e) : 0;
    }
    float out = nanf("");
    if (gsb_hamming_matrix(a, 1, b, 1, (uint32_t)len, sig_type, &out, device) != GSB_OK) return nanf("");
    return out;
}
extern "C" float gsb_dist_hamming_u16(const uint16_t *a, const uint16_t *b, unsigned long long len) {
    return dist_hamming_scalar(a, b, len, GSB_SIG_U16);
}
extern "C" float gsb_dist_hamming_u32(const uint32_t *a, const uint32_t *b, unsigned long long len) {
    return dist_hamming_scalar(a, b, len, GSB_SIG_U32);
}
extern "C" float gsb_dist_hamming_u64(const uint64_t *a, const uint64_t *b, unsigned long long len) {
    return dist_hamming_scalar(a, b, len, GSB_SIG_U64);
}
extern "C" float gsb_dist_hamming_f32(const float *a, const float *b, unsigned long long len) {
    return dist_hamming_scalar(a, b, len, GSB_SIG_F32);
}

// ---- device / pinned-host buffers for hosts that have no CUDA binding of their own (the Rust side
// of INTEGRATION.md, gsearch_b200/cli.py): plain cudaMalloc / cudaMallocHost / cudaMemcpy
extern "C" int gsb_device_malloc(int device, uint64_t bytes, void **out) {
    if (!out) {
        set_error("gsb_device_malloc: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    int rc = check_device(device);
    if (rc) return rc;
    GSB_CUDA_TRY(cudaSetDevice(device));
    GSB_CUDA_TRY(cudaMalloc(out, bytes ? bytes : 1));
    return GSB_OK;
}
extern "C" void gsb_device_free(int device, void *p) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}
extern "C" int gsb_memcpy_h2d(int device, void *d_dst, const void *src, uint64_t bytes) {
    GSB_CUDA_TRY(cudaSetDevice(device));
    GSB_CUDA_TRY(cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice));
    return GSB_OK;
}
extern "C" int gsb_memcpy_d2h(int device, void *dst, const void *d_src, uint64_t bytes) {
    GSB_CUDA_TRY(cudaSetDevice(device));
    GSB_CUDA_TRY(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return GSB_OK;
}
extern "C" int gsb_host_alloc_pinned(uint64_t bytes, void **out) {
    if (!out) {
        set_error("gsb_host_alloc_pinned: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    GSB_CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 1));
    return GSB_OK;
}
extern "C" void gsb_host_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}
