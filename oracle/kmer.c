/*
 * kmer.c -- k-mer generation, the two k-mer hash closures, exact multiplicities
 * (test infrastructure, see gso.h).
 *
 * Follows:
 *   kmerutils KmerSeqIterator<Kmer>::next [U, high]: sliding window,
 *     newest base in the low bits, L-k+1 k-mers per sequence, none if L < k; usage
 *     shape evidenced in-tree at src/bin/hypermash.rs:28-38,146-166
 *   DNA hash closure  src/dna/dnasketch.rs:164-169 (= src/dna/dnarequest.rs:118-122):
 *     canonical = min(kmer, reverse_complement(kmer)); value & (2^(2k) - 1)
 *   AA hash closure   src/aa/aasketch.rs:156-160: value & (2^(5k) - 1), no revcomp
 *   k -> k-mer type -> Val type   src/dna/dnasketch.rs:500-515, src/aa/aasketch.rs:457-466
 *   multiplicity map  kmerutils ProbHash3aSketch::sketch_compressedkmer_seqs [U, high]:
 *     IndexMap<Val, f64>, `*entry(hash).or_insert(0.) += 1.` per k-mer, iterated in
 *     insertion order by ProbMinHash3a::hashset
 */
#include "gso.h"

#include <stdlib.h>
#include <string.h>

int gso_sig_type(const gso_sketch_params *p) {
    if (p->algo == GSO_ALGO_PROB3A || p->algo == GSO_ALGO_SUPER2) { /* SuperHash2Sketch<Kmer, u32|u64, Fx>: dnasketch.rs:575-599 */
        if (p->data_t == GSO_DATA_DNA) {
            if (p->kmer_size <= 14 || p->kmer_size == 16) return GSO_SIG_U32;
            return GSO_SIG_U64;
        }
        return p->kmer_size <= 6 ? GSO_SIG_U32 : GSO_SIG_U64;
    }
    if (p->algo == GSO_ALGO_HLL) return GSO_SIG_U16; /* HyperLogLogSketch<Kmer, u16>: dnasketch.rs:541-573 */
    return GSO_SIG_F32; /* SuperHashSketch<_, f32>, OptDensHashSketch<_, f32> */
}

uint32_t gso_elem_size(const gso_sketch_params *p) {
    int t = gso_sig_type(p);
    return t == GSO_SIG_U64 ? 8u : (t == GSO_SIG_U16 ? 2u : 4u);
}

int gso_kmer_values(const gso_seqs *s, uint32_t data_t, uint32_t k, uint64_t **vals_out,
                    uint64_t *n_out) {
    uint64_t total = 0;
    for (uint64_t q = 0; q < s->nseq; q++) {
        uint64_t L = s->seq_off[q + 1] - s->seq_off[q];
        if (L >= k) total += L - k + 1;
    }
    uint64_t *vals = (uint64_t *)malloc((total ? total : 1) * sizeof(uint64_t));
    if (!vals) return 4;
    uint64_t n = 0;
    if (data_t == GSO_DATA_DNA) {
        const uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
        const int hi_shift = 2 * ((int)k - 1);
        for (uint64_t q = 0; q < s->nseq; q++) {
            const uint8_t *c = s->codes + s->seq_off[q];
            uint64_t L = s->seq_off[q + 1] - s->seq_off[q];
            uint64_t fw = 0, rc = 0;
            for (uint64_t i = 0; i < L; i++) {
                uint64_t b = c[i];
                fw = ((fw << 2) | b) & mask;
                rc = (rc >> 2) | ((3 - b) << hi_shift);
                if (i + 1 >= k) vals[n++] = (fw < rc ? fw : rc) & mask;
            }
        }
    } else {
        const uint64_t mask = (1ULL << (5 * k)) - 1;
        for (uint64_t q = 0; q < s->nseq; q++) {
            const uint8_t *c = s->codes + s->seq_off[q];
            uint64_t L = s->seq_off[q + 1] - s->seq_off[q];
            uint64_t v = 0;
            for (uint64_t i = 0; i < L; i++) {
                v = ((v << 5) | c[i]) & mask;
                if (i + 1 >= k) vals[n++] = v;
            }
        }
    }
    *vals_out = vals;
    *n_out = n;
    return 0;
}

/* insertion-ordered exact multiplicity table (IndexMap semantics) */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

int gso_count_kmers(const uint64_t *vals, uint64_t n, uint64_t **keys_out, double **w_out,
                    uint64_t *ndistinct_out) {
    uint64_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    uint32_t *slot = (uint32_t *)malloc(cap * sizeof(uint32_t)); /* index+1 into keys */
    uint64_t *keys = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    double *w = (double *)malloc((n ? n : 1) * sizeof(double));
    if (!slot || !keys || !w || n >= 0xFFFFFFFFULL) {
        free(slot);
        free(keys);
        free(w);
        return 4;
    }
    memset(slot, 0, cap * sizeof(uint32_t));
    uint64_t nd = 0;
    const uint64_t cm = cap - 1;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t v = vals[i];
        uint64_t h = mix64(v) & cm;
        for (;;) {
            uint32_t e = slot[h];
            if (e == 0) {
                keys[nd] = v;
                w[nd] = 1.0;
                nd++;
                slot[h] = (uint32_t)nd;
                break;
            }
            if (keys[e - 1] == v) {
                w[e - 1] += 1.0;
                break;
            }
            h = (h + 1) & cm;
        }
    }
    free(slot);
    *keys_out = keys;
    *w_out = w;
    *ndistinct_out = nd;
    return 0;
}
