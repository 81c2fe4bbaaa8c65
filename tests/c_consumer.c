/* A plain-C consumer of include/gsearch_b200.h: what a Rust `-sys` crate (cc + bindgen) would link.
 * Built and run by tests/test_c_consumer.py.  Exit code 0 = the whole path ran on a GPU;
 * 2 = the library reported GSB_ERR_NO_DEVICE (CPU box: there is no CPU path). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gsearch_b200.h"

static void make_fasta(char *out, size_t n, unsigned seed, const char *name) {
    size_t o = (size_t)sprintf(out, ">%s\n", name);
    unsigned x = seed;
    for (size_t i = 0; i < n; i++) {
        x = x * 1664525u + 1013904223u;
        out[o++] = "ACGT"[(x >> 24) & 3];
        if ((i + 1) % 70 == 0) out[o++] = '\n';
    }
    out[o++] = '\n';
    out[o] = 0;
}

int main(void) {
    enum { N = 4, L = 20000, S = 256 };
    printf("%s, %d device(s)\n", gsb_version(), gsb_device_count());
    gsb_sketch_params sp = {21, S, GSB_ALGO_PROB3A, GSB_DATA_DNA, 0, 0};
    gsb_sketcher *sk = NULL;
    int rc = gsb_sketcher_create(&sp, 0, &sk);
    if (rc == GSB_ERR_NO_DEVICE) {
        printf("no device: %s\n", gsb_last_error());
        return 2;
    }
    if (rc) {
        printf("create failed %d: %s\n", rc, gsb_last_error());
        return 1;
    }
    char *buf = malloc((size_t)N * (L + L / 70 + 64));
    uint64_t off[N + 1];
    off[0] = 0;
    for (int i = 0; i < N; i++) {
        /* files 0 and 1 are identical sequences under different names, 2 and 3 are unrelated */
        make_fasta(buf + off[i], L, i == 1 ? 7u : 7u + (unsigned)i * 1000u, i == 1 ? "copy" : "genome");
        off[i + 1] = off[i] + strlen(buf + off[i]);
    }
    if (gsb_sketcher_sig_type(sk) != GSB_SIG_U64 || gsb_sketcher_elem_size(sk) != 8) return 1;
    uint64_t *sig = malloc((size_t)N * S * 8), nb[N];
    rc = gsb_sketch_fasta_batch(sk, (const uint8_t *)buf, off, N, sig, nb);
    if (rc) {
        printf("sketch failed %d: %s\n", rc, gsb_last_error());
        return 1;
    }
    for (int i = 0; i < N; i++)
        if (nb[i] != L) return 1;
    float d[N * N];
    if (gsb_hamming_matrix(sig, N, sig, N, S, GSB_SIG_U64, d, 0)) return 1;
    if (d[0 * N + 1] != 0.0f || d[0 * N + 2] < 0.9f) return 1; /* same sequence: distance 0 */
    /* the scalar Distance::eval-shaped export (DistCFnPtr<u64>) gives the same bits, one pair per call */
    float (*eval64)(const uint64_t *, const uint64_t *, unsigned long long) = gsb_dist_hamming_u64;
    if (eval64(sig, sig + 2 * S, S) != d[0 * N + 2] || eval64(sig, sig + S, S) != 0.0f) return 1;
    gsb_index_params ip = {8, 1000, 16, 32, 1.0, GSB_SIG_U64, S, 1, 0, 0x5EED};
    gsb_index *idx = NULL;
    if (gsb_index_create(&ip, 0, &idx)) return 1;
    uint64_t ids[N] = {10, 11, 12, 13};
    if (gsb_index_insert_batch(idx, sig, ids, N) || gsb_index_nb_point(idx) != N) return 1;
    gsb_neighbour out[N * 2];
    uint32_t cnt[N];
    if (gsb_index_search_batch(idx, sig, N, 2, 16, out, cnt, NULL)) return 1;
    if (cnt[2] != 2 || out[2 * 2].d_id != 12 || out[2 * 2].distance != 0.0f) return 1;
    printf("ok: d(0,1)=%g d(0,2)=%g, nearest of genome 2 = id %llu\n", d[1], d[2],
           (unsigned long long)out[2 * 2].d_id);
    gsb_index_destroy(idx);
    gsb_sketcher_destroy(sk);
    free(sig);
    free(buf);
    return 0;
}
