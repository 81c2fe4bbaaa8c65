/*
 * hamming.c -- DistHamming (test infrastructure, see gso.h).
 *
 * Follows anndists::dist::DistHamming::eval [U, high; SURVEY A.9]:
 *   eval(a, b) = (number of i with a[i] != b[i]) as f32 / a.len() as f32
 * used by the reference through Hnsw::<Sig, DistHamming>::new (src/dna/dnasketch.rs:139)
 * and directly at src/bin/bindash.rs:94-95.  For f32 signatures `!=` is the IEEE value
 * comparison.
 */
#include "gso.h"

float gso_hamming(const void *a, const void *b, uint32_t S, uint32_t sig_type) {
    uint32_t cnt = 0;
    switch (sig_type) {
    case GSO_SIG_U32: {
        const uint32_t *x = (const uint32_t *)a, *y = (const uint32_t *)b;
        for (uint32_t i = 0; i < S; i++) cnt += (x[i] != y[i]);
        break;
    }
    case GSO_SIG_U64: {
        const uint64_t *x = (const uint64_t *)a, *y = (const uint64_t *)b;
        for (uint32_t i = 0; i < S; i++) cnt += (x[i] != y[i]);
        break;
    }
    case GSO_SIG_F32: {
        const float *x = (const float *)a, *y = (const float *)b;
        for (uint32_t i = 0; i < S; i++) cnt += (x[i] != y[i]);
        break;
    }
    default: {
        const uint16_t *x = (const uint16_t *)a, *y = (const uint16_t *)b;
        for (uint32_t i = 0; i < S; i++) cnt += (x[i] != y[i]);
        break;
    }
    }
    return (float)cnt / (float)S;
}

static uint32_t esize(uint32_t t) { return t == GSO_SIG_U64 ? 8u : (t == GSO_SIG_U16 ? 2u : 4u); }

void gso_hamming_matrix(const void *q, uint32_t nq, const void *c, uint32_t n, uint32_t S,
                        uint32_t sig_type, float *out, int nthreads) {
    const uint64_t row = (uint64_t)S * esize(sig_type);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t i = 0; i < (int64_t)nq; i++)
        for (uint32_t j = 0; j < n; j++)
            out[(uint64_t)i * n + j] = gso_hamming((const uint8_t *)q + (uint64_t)i * row,
                                                   (const uint8_t *)c + (uint64_t)j * row, S,
                                                   sig_type);
}
