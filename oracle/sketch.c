/*
 * sketch.c -- the three sketchers and the per-file driver (test infrastructure, see gso.h).
 *
 * Follows:
 *   probminhash::probminhasher::ProbMinHash3a::{new, hashset, get_signature} [U,
 *     medium-high; SURVEY A.5] with MaxValueTracker (a max over the m slot values),
 *     ExpRestricted01 (rng.c) and NoHashHasher (seed = big-endian read of the value's
 *     native bytes, i.e. the byte-swapped value on x86; GSO_SPEC_NOHASH_IDENTITY = the
 *     other recalled variant).  Reached from src/dna/dnasketch.rs:336,357 through
 *     kmerutils::ProbHash3aSketch::sketch_compressedkmer[_seqs] [U, high].
 *   probminhash::densminhash::OptDensMinHash::{sketch, end_sketch} [U, medium / low for
 *     densification; SURVEY A.7], hasher = fxhash::FxHasher64 on one word.
 *     Sig = f32 proven in-tree at src/bin/bindash.rs:52; AA dispatch src/aa/aasketch.rs:524-536.
 *   probminhash::superminhasher::SuperMinHash::sketch [U, medium; SURVEY A.8],
 *     dispatch src/dna/dnasketch.rs:520-540.
 *   Per-file driver: one signature per file; seq mode sketches all records of a file into
 *     one signature (sketch_compressedkmer_seqs, src/dna/dnasketch.rs:347-365), block mode
 *     sketches the concatenated file (src/dna/dnasketch.rs:327-345); one file per worker
 *     thread (src/dna/dnasketch.rs:325).
 */
#include "gso.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* ProbMinHash3a                                                              */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t m;
    double *t; /* max-tree: leaves t[m..2m), parent(i) = i/2, root t[1] */
} maxtrack;

static int maxtrack_init(maxtrack *mt, uint32_t m) {
    mt->m = m;
    mt->t = (double *)malloc(2ull * m * sizeof(double));
    if (!mt->t) return 4;
    for (uint32_t i = 0; i < 2 * m; i++) mt->t[i] = 1.7976931348623157e308; /* f64::MAX */
    return 0;
}
static inline double maxtrack_value(const maxtrack *mt, uint32_t k) { return mt->t[mt->m + k]; }
static inline double maxtrack_max(const maxtrack *mt) { return mt->m > 1 ? mt->t[1] : mt->t[1]; }
static void maxtrack_update(maxtrack *mt, uint32_t k, double v) {
    uint32_t i = mt->m + k;
    mt->t[i] = v;
    for (i >>= 1; i >= 1; i >>= 1) {
        double a = mt->t[2 * i], b = mt->t[2 * i + 1];
        double mx = a > b ? a : b;
        if (mt->t[i] == mx) break;
        mt->t[i] = mx;
    }
}

static inline uint64_t nohash_seed(uint64_t v, uint32_t val_bytes, uint32_t spec_flags) {
    if (spec_flags & GSO_SPEC_NOHASH_IDENTITY) return v;
    if (val_bytes == 4) return (uint64_t)__builtin_bswap32((uint32_t)v);
    return __builtin_bswap64(v);
}

typedef struct {
    uint64_t key;
    double winv;
    gso_xoshiro rng;
} pending;

int gso_probminhash3a(const uint64_t *keys, const double *w, uint64_t nd, uint32_t m,
                      uint32_t val_bytes, uint32_t spec_flags, uint64_t *sig_out,
                      double *hmin_out) {
    if (m < 2) return 1;
    maxtrack mt;
    if (maxtrack_init(&mt, m)) return 4;
    gso_exp01 e01;
    gso_exp01_init(&e01, log((double)m / (double)(m - 1)));
    for (uint32_t i = 0; i < m; i++) sig_out[i] = 0; /* initobj = num::zero() */
    pending *buf = NULL;
    uint64_t nbuf = 0, capbuf = 0;

    for (uint64_t it = 0; it < nd; it++) {
        const uint64_t key = keys[it];
        const double winv = 1.0 / w[it];
        gso_xoshiro rng;
        gso_xoshiro_seed_from_u64(&rng, nohash_seed(key, val_bytes, spec_flags));
        double h = winv * gso_exp01_sample(&e01, &rng);
        double qmax = maxtrack_max(&mt);
        if (h < qmax) {
            uint32_t k = (uint32_t)gso_uniform_usize(&rng, m);
            if (h < maxtrack_value(&mt, k)) {
                sig_out[k] = key;
                maxtrack_update(&mt, k, h);
                qmax = maxtrack_max(&mt);
            }
            if (winv < qmax) {
                if (nbuf == capbuf) {
                    capbuf = capbuf ? 2 * capbuf : 1024;
                    pending *nb = (pending *)realloc(buf, capbuf * sizeof(pending));
                    if (!nb) {
                        free(buf);
                        free(mt.t);
                        return 4;
                    }
                    buf = nb;
                }
                buf[nbuf].key = key;
                buf[nbuf].winv = winv;
                buf[nbuf].rng = rng;
                nbuf++;
            }
        }
    }
    uint64_t i = 2;
    while (nbuf > 0) {
        uint64_t insert_pos = 0;
        for (uint64_t j = 0; j < nbuf; j++) {
            pending *p = &buf[j];
            double h = p->winv * (double)(i - 1);
            if (h < maxtrack_max(&mt)) {
                h = h + p->winv * gso_exp01_sample(&e01, &p->rng);
                uint32_t k = (uint32_t)gso_uniform_usize(&p->rng, m);
                if (h < maxtrack_value(&mt, k)) {
                    sig_out[k] = p->key;
                    maxtrack_update(&mt, k, h);
                }
                if (p->winv * (double)i < maxtrack_max(&mt)) {
                    buf[insert_pos] = *p;
                    insert_pos++;
                }
            }
        }
        nbuf = insert_pos;
        i++;
    }
    if (hmin_out)
        for (uint32_t k = 0; k < m; k++) hmin_out[k] = maxtrack_value(&mt, k);
    free(buf);
    free(mt.t);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* OptDensMinHash                                                             */
/* ------------------------------------------------------------------------- */
#define FX_SEED64 0x517cc1b727220a95ULL
#define OPTDENS_LARGE 4294967296.0f /* F::from(u32::MAX) rounded to f32 */

int gso_optdens(const uint64_t *vals, uint64_t n, uint32_t m, uint32_t spec_flags,
                float *sig_out) {
    if (m < 1) return 1;
    for (uint32_t k = 0; k < m; k++) sig_out[k] = OPTDENS_LARGE;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t hval = vals[i] * FX_SEED64; /* FxHasher64, one word */
        gso_xoshiro rng;
        gso_xoshiro_seed_from_u64(&rng, hval);
        float r = (spec_flags & GSO_SPEC_OPTDENS_F64_DRAW) ? (float)gso_uniform_f64(&rng)
                                                           : gso_uniform_f32(&rng);
        uint32_t k = (uint32_t)gso_uniform_usize(&rng, m);
        if (r <= sig_out[k]) sig_out[k] = r;
    }
    /* end_sketch(): densification of empty bins (cold path; [U, low] -- frozen here as:
     * for each empty bin k, rng = seed_from_u64(k), draw j = Uniform[0,m) until bin j was
     * non-empty before densification, copy it). */
    uint32_t nempty = 0;
    for (uint32_t k = 0; k < m; k++) nempty += (sig_out[k] > 1.5f);
    if (nempty == 0 || nempty == m) return 0;
    uint8_t *empty = (uint8_t *)malloc(m);
    if (!empty) return 4;
    for (uint32_t k = 0; k < m; k++) empty[k] = (sig_out[k] > 1.5f);
    for (uint32_t k = 0; k < m; k++) {
        if (!empty[k]) continue;
        gso_xoshiro rng;
        gso_xoshiro_seed_from_u64(&rng, (uint64_t)k);
        for (;;) {
            uint32_t j = (uint32_t)gso_uniform_usize(&rng, m);
            if (!empty[j]) {
                sig_out[k] = sig_out[j];
                break;
            }
        }
    }
    free(empty);
    return 0;
}

/* RevOptDensMinHash (probminhash [U, low]; dispatch src/dna/dnasketch.rs:623-641): the same
 * one-permutation bins as OptDens; only the densification differs -- non-empty bins PUSH their
 * value into empty ones (Mai et al. 2020, "Densified MinHash in O(k log k)").  Frozen here as:
 * rounds a = 0, 1, 2, ...; in a round every bin i that was non-empty BEFORE densification, in
 * increasing i, targets bin j = target(i, a) and fills it if it is still empty; until no bin is
 * empty.  target(i, a) = floor(mix(i, a) * m / 2^64) with the stateless 64-bit mix below, so that
 * the restatement and the CUDA kernel (k3_optdens_finalize, REV mode) need no generator state. */
uint32_t gso_revdens_target(uint32_t i, uint32_t a, uint32_t m) {
    uint64_t z = (uint64_t)i * 0x9e3779b97f4a7c15ULL + (uint64_t)a * 0xd1b54a32d192ed03ULL + 0x2545f4914f6cdd1dULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return (uint32_t)(((__uint128_t)z * (__uint128_t)m) >> 64);
}

int gso_revoptdens(const uint64_t *vals, uint64_t n, uint32_t m, uint32_t spec_flags, float *sig_out) {
    if (m < 1) return 1;
    for (uint32_t k = 0; k < m; k++) sig_out[k] = OPTDENS_LARGE;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t hval = vals[i] * FX_SEED64;
        gso_xoshiro rng;
        gso_xoshiro_seed_from_u64(&rng, hval);
        float r = (spec_flags & GSO_SPEC_OPTDENS_F64_DRAW) ? (float)gso_uniform_f64(&rng) : gso_uniform_f32(&rng);
        uint32_t k = (uint32_t)gso_uniform_usize(&rng, m);
        if (r <= sig_out[k]) sig_out[k] = r;
    }
    uint32_t nempty = 0;
    for (uint32_t k = 0; k < m; k++) nempty += (sig_out[k] > 1.5f);
    if (nempty == 0 || nempty == m) return 0;
    uint8_t *orig = (uint8_t *)malloc(m);
    if (!orig) return 4;
    for (uint32_t k = 0; k < m; k++) orig[k] = !(sig_out[k] > 1.5f);
    for (uint32_t a = 0; nempty > 0; a++) {
        for (uint32_t i = 0; i < m && nempty > 0; i++) {
            if (!orig[i]) continue;
            const uint32_t j = gso_revdens_target(i, a, m);
            if (sig_out[j] > 1.5f) {
                sig_out[j] = sig_out[i];
                nempty--;
            }
        }
    }
    free(orig);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* SuperMinHash                                                               */
/* ------------------------------------------------------------------------- */
int gso_superminhash(const uint64_t *vals, uint64_t n, uint32_t m, float *sig_out) {
    if (m < 1) return 1;
    int64_t *q = (int64_t *)malloc(m * sizeof(int64_t));
    uint32_t *p = (uint32_t *)malloc(m * sizeof(uint32_t));
    int64_t *b = (int64_t *)malloc(m * sizeof(int64_t));
    if (!q || !p || !b) {
        free(q);
        free(p);
        free(b);
        return 4;
    }
    for (uint32_t i = 0; i < m; i++) {
        sig_out[i] = OPTDENS_LARGE;
        q[i] = -1;
        p[i] = 0;
        b[i] = 0;
    }
    b[m - 1] = m;
    uint32_t a_upper = m - 1;
    for (uint64_t it = 0; it < n; it++) {
        const uint64_t hval = vals[it] * FX_SEED64;
        gso_xoshiro rng;
        gso_xoshiro_seed_from_u64(&rng, hval);
        const int64_t irank = (int64_t)it;
        uint32_t j = 0;
        while (j <= a_upper) {
            float r = gso_uniform_f32(&rng);
            /* Uniform::<usize>::new(j, m): low + draw over range m-j */
            uint32_t k = j + (uint32_t)gso_uniform_usize(&rng, (uint64_t)(m - j));
            if (q[j] != irank) {
                q[j] = irank;
                p[j] = j;
            }
            if (q[k] != irank) {
                q[k] = irank;
                p[k] = k;
            }
            uint32_t t = p[j];
            p[j] = p[k];
            p[k] = t;
            float rpj = r + (float)j;
            if (rpj < sig_out[p[j]]) {
                float old = sig_out[p[j]];
                uint32_t j2 = (old >= (float)(m - 1)) ? (m - 1) : (uint32_t)old;
                sig_out[p[j]] = rpj;
                if (j < j2) {
                    b[j2] -= 1;
                    b[j] += 1;
                    while (b[a_upper] == 0) a_upper--;
                }
            }
            j++;
        }
    }
    free(q);
    free(p);
    free(b);
    return 0;
}

/* SuperMinHash2 (probminhash superminhasher2 [U, low-medium]; dispatch src/dna/dnasketch.rs:575-599
 * with the hasher chosen by the CALLER: FxHasher32 for 32-bit k-mers, FxHasher64 for 64-bit ones):
 * Ertl's SuperMinHash whose signature holds, per slot, the HASH of the item that gave the minimum
 * (so signatures compare by equality: DistHamming on u32 / u64).  Frozen here as: per item
 * hval = fx(item) (fx32: (u32)item * 0x9e3779b9 ; fx64: item * 0x517cc1b727220a95), generator seeded
 * by hval; level j draws r = Uniform<f64>[0,1) and k = j + Uniform<usize>[0, m-j), lazy Fisher-Yates
 * swap, value r + j lands in slot p[j]; the slot keeps the smallest value and the hval of its item
 * (identical values: the smaller hval -- probability 2^-52, the reference is order dependent there).
 * A slot no item reached (empty input only: every item visits all m slots at worst) holds the
 * signature type's maximum, as the other order-free minima here do. */
int gso_superminhash2(const uint64_t *vals, uint64_t n, uint32_t m, int kt32, uint64_t *sig_out) {
    if (m < 1) return 1;
    int64_t *q = (int64_t *)malloc(m * sizeof(int64_t));
    uint32_t *p = (uint32_t *)malloc(m * sizeof(uint32_t));
    int64_t *b = (int64_t *)malloc(m * sizeof(int64_t));
    double *h = (double *)malloc(m * sizeof(double));
    if (!q || !p || !b || !h) {
        free(q);
        free(p);
        free(b);
        free(h);
        return 4;
    }
    for (uint32_t i = 0; i < m; i++) {
        sig_out[i] = kt32 ? 0xFFFFFFFFull : ~0ull;
        h[i] = 4294967296.0;
        q[i] = -1;
        p[i] = 0;
        b[i] = 0;
    }
    b[m - 1] = m;
    uint32_t a_upper = m - 1;
    for (uint64_t it = 0; it < n; it++) {
        const uint64_t hval = kt32 ? (uint64_t)((uint32_t)vals[it] * 0x9e3779b9u) : vals[it] * FX_SEED64;
        gso_xoshiro rng;
        gso_xoshiro_seed_from_u64(&rng, hval);
        const int64_t irank = (int64_t)it;
        uint32_t j = 0;
        while (j <= a_upper) {
            const double r = gso_uniform_f64(&rng);
            const uint32_t k = j + (uint32_t)gso_uniform_usize(&rng, (uint64_t)(m - j));
            if (q[j] != irank) {
                q[j] = irank;
                p[j] = j;
            }
            if (q[k] != irank) {
                q[k] = irank;
                p[k] = k;
            }
            const uint32_t t = p[j];
            p[j] = p[k];
            p[k] = t;
            const double rpj = r + (double)j;
            const uint32_t slot = p[j];
            if (rpj < h[slot] || (rpj == h[slot] && hval < sig_out[slot])) {
                const double old = h[slot];
                const uint32_t j2 = (old >= (double)(m - 1)) ? (m - 1) : (uint32_t)old;
                h[slot] = rpj;
                sig_out[slot] = hval;
                if (j < j2) {
                    b[j2] -= 1;
                    b[j] += 1;
                    while (b[a_upper] == 0) a_upper--;
                }
            }
            j++;
        }
    }
    free(q);
    free(p);
    free(b);
    free(h);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* per-file driver                                                            */
/* ------------------------------------------------------------------------- */
/* SetSketch ("--algo hll": HyperLogLogSketch<Kmer, u16> over probminhash::setsketcher::SetSketcher
 * [U, low-medium]; dispatch src/dna/dnasketch.rs:541-573 with SetSketchParams::default() and
 * m = sketch_size: b = 1.001, a = 20, q = 2^16 - 2).  Ertl's SetSketch1: per item, generator seeded
 * by hval = fx64(item); points x_1 < x_2 < ... with x_{j+1} = x_j + (1/a)/(m-j) * E_j (E_j a unit
 * exponential), the j-th point goes to the register drawn WITHOUT replacement (lazy Fisher-Yates,
 * r = j + Uniform<usize>[0, m-j)) with the value k = clamp(floor(1 - log_b x), 0, q+1); a register
 * keeps the maximum.  An item stops as soon as k <= k_low, a lower bound of all registers refreshed
 * every m successful updates -- neutral for the result: later points have smaller k.  Items are a
 * set: repeated k-mers change nothing.
 * Frozen here: E = -ln_spec(1 - Uniform<f64>) (rand_distr::Exp1 is a ziggurat whose tables are not
 * restated), ln_spec in place of libm, draw order (E_j, then the register).                       */
int gso_setsketch(const uint64_t *vals, uint64_t n, uint32_t m, uint16_t *sig_out) {
    if (m < 1) return 1;
    const double b = 1.001, a = 20.0;
    const double lnb = gso_ln_spec(b), inva = 1.0 / a;
    const double qp1 = 65535.0;
    int64_t *q = (int64_t *)malloc(m * sizeof(int64_t));
    uint32_t *p = (uint32_t *)malloc(m * sizeof(uint32_t));
    if (!q || !p) {
        free(q);
        free(p);
        return 4;
    }
    for (uint32_t i = 0; i < m; i++) {
        sig_out[i] = 0;
        q[i] = -1;
        p[i] = 0;
    }
    double k_low = 0.0;
    uint64_t nbmin = 0;
    for (uint64_t it = 0; it < n; it++) {
        const uint64_t hval = vals[it] * FX_SEED64;
        gso_xoshiro rng;
        gso_xoshiro_seed_from_u64(&rng, hval);
        const int64_t irank = (int64_t)it;
        double x = 0.0;
        for (uint32_t j = 0; j < m; j++) {
            const double e = -gso_ln_spec(1.0 - gso_uniform_f64(&rng));
            x = x + (inva / (double)(m - j)) * e;
            double kf = 0.0;
            if (x > 0.0) {
                const double z = 1.0 - gso_ln_spec(x) / lnb;
                kf = floor(z);
                if (kf < 0.0) kf = 0.0;
                if (kf > qp1) kf = qp1;
            } else {
                kf = qp1; /* x == 0: the uniform draw was 0 */
            }
            if (kf <= k_low) break;
            const uint32_t r = j + (uint32_t)gso_uniform_usize(&rng, (uint64_t)(m - j));
            if (q[j] != irank) {
                q[j] = irank;
                p[j] = j;
            }
            if (q[r] != irank) {
                q[r] = irank;
                p[r] = r;
            }
            const uint32_t t = p[j];
            p[j] = p[r];
            p[r] = t;
            const uint32_t reg = p[j];
            if (kf > (double)sig_out[reg]) {
                sig_out[reg] = (uint16_t)kf;
                if (++nbmin % m == 0) {
                    uint16_t mn = 65535;
                    for (uint32_t i = 0; i < m; i++) mn = sig_out[i] < mn ? sig_out[i] : mn;
                    k_low = (double)mn;
                }
            }
        }
    }
    free(q);
    free(p);
    return 0;
}

static int sketch_one(const gso_sketch_params *p, const uint8_t *bytes, uint64_t len, void *sig,
                      uint64_t *nb_bases) {
    gso_seqs s;
    int rc = gso_parse_fasta(bytes, len, p->data_t, p->block_flag, &s);
    if (rc) return rc;
    if (nb_bases) *nb_bases = s.seq_off[s.nseq];
    uint64_t *vals = NULL, n = 0;
    rc = gso_kmer_values(&s, p->data_t, p->kmer_size, &vals, &n);
    gso_seqs_free(&s);
    if (rc) return rc;
    const uint32_t m = p->sketch_size;
    if (p->algo == GSO_ALGO_PROB3A) {
        uint64_t *keys = NULL, nd = 0;
        double *w = NULL;
        rc = gso_count_kmers(vals, n, &keys, &w, &nd);
        if (!rc) {
            uint64_t *sig64 = (uint64_t *)malloc(m * sizeof(uint64_t));
            const uint32_t vb = gso_elem_size(p);
            rc = sig64 ? gso_probminhash3a(keys, w, nd, m, vb, p->spec_flags, sig64, NULL) : 4;
            if (!rc) {
                if (vb == 4)
                    for (uint32_t i = 0; i < m; i++) ((uint32_t *)sig)[i] = (uint32_t)sig64[i];
                else
                    memcpy(sig, sig64, m * sizeof(uint64_t));
            }
            free(sig64);
        }
        free(keys);
        free(w);
    } else if (p->algo == GSO_ALGO_OPTDENS) {
        rc = gso_optdens(vals, n, m, p->spec_flags, (float *)sig);
    } else if (p->algo == GSO_ALGO_SUPER) {
        rc = gso_superminhash(vals, n, m, (float *)sig);
    } else if (p->algo == GSO_ALGO_REVOPTDENS) {
        rc = gso_revoptdens(vals, n, m, p->spec_flags, (float *)sig);
    } else if (p->algo == GSO_ALGO_SUPER2) {
        const int kt32 = gso_sig_type(p) == GSO_SIG_U32;
        uint64_t *tmp = (uint64_t *)malloc((size_t)m * sizeof(uint64_t));
        if (!tmp) {
            free(vals);
            return 4;
        }
        rc = gso_superminhash2(vals, n, m, kt32, tmp);
        if (kt32)
            for (uint32_t i = 0; i < m; i++) ((uint32_t *)sig)[i] = (uint32_t)tmp[i];
        else
            memcpy(sig, tmp, (size_t)m * sizeof(uint64_t));
        free(tmp);
    } else if (p->algo == GSO_ALGO_HLL) {
        rc = gso_setsketch(vals, n, m, (uint16_t *)sig);
    } else {
        rc = 6;
    }
    free(vals);
    return rc;
}

int gso_sketch_fasta_batch(const gso_sketch_params *p, const uint8_t *bytes,
                           const uint64_t *offsets, uint32_t n, void *sig_out,
                           uint64_t *nb_bases_out, int nthreads) {
    if (p->data_t == GSO_DATA_DNA && (p->kmer_size < 1 || p->kmer_size > 31)) return 1;
    if (p->data_t == GSO_DATA_AA && (p->kmer_size < 1 || p->kmer_size > 12)) return 1;
    if (p->sketch_size < 2) return 1;
    const uint64_t row = (uint64_t)p->sketch_size * gso_elem_size(p);
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        int rc = sketch_one(p, bytes + offsets[i], offsets[i + 1] - offsets[i],
                            (uint8_t *)sig_out + (uint64_t)i * row,
                            nb_bases_out ? &nb_bases_out[i] : NULL);
        if (rc) {
#pragma omp critical
            err = rc;
        }
    }
    return err;
}
