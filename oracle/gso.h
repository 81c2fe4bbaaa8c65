/*
 * gso.h -- CPU oracle for the GSearch sketch-and-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's CPU
 * algorithm, used as (i) the parity checker in tests/ and __graft_entry__.smoke()
 * and (ii) the `cpu_baseline` / `--impl reference` arm of bench.py.  Nothing in
 * gsearch_b200/ (the product) may include, link or call it.
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in four upstream crates that
 * are neither vendored in /root/reference nor version-pinned by it:
 *   kmerutils  (git master, no rev)        /root/reference/Cargo.toml:125
 *   probminhash = "0.1"                    /root/reference/Cargo.toml:122
 *   hnsw_rs     = "0.3"                    /root/reference/Cargo.toml:115
 *   anndists    = "0.1"                    /root/reference/Cargo.toml:56
 * and the reference ships no tests, fixtures or golden vectors, and
 * no Rust toolchain exists in this image, so the reference cannot be compiled here.
 * What IS pinned: the published primitive vectors (SplitMix64, xoshiro256++), the
 * in-tree semantics cited per function below, and the published algorithms (Ertl,
 * ProbMinHash, TKDE 2020; Ertl, SuperMinHash, 2017; Shrivastava, Optimal
 * Densification, ICML 2017; Malkov & Yashunin, HNSW).  Choices that could not be
 * verified against upstream source are switchable through `spec_flags`.
 */
#ifndef GSO_H
#define GSO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same numeric values as include/gsearch_b200.h */
enum { GSO_ALGO_PROB3A = 0, GSO_ALGO_SUPER = 1, GSO_ALGO_OPTDENS = 2, GSO_ALGO_REVOPTDENS = 3, GSO_ALGO_SUPER2 = 4, GSO_ALGO_HLL = 5 };
enum { GSO_DATA_DNA = 0, GSO_DATA_AA = 1 };
enum { GSO_SIG_U32 = 0, GSO_SIG_U64 = 1, GSO_SIG_F32 = 2, GSO_SIG_U16 = 3 };
enum { GSO_SPEC_NOHASH_IDENTITY = 1u << 0, GSO_SPEC_OPTDENS_F64_DRAW = 1u << 1 };

typedef struct {
    uint32_t kmer_size;
    uint32_t sketch_size;
    uint32_t algo;
    uint32_t data_t;
    uint32_t block_flag;
    uint32_t spec_flags;
} gso_sketch_params;

/* ---------------- rng.c : primitives ---------------- */
typedef struct { uint64_t s[4]; } gso_xoshiro;

uint64_t gso_splitmix64_next(uint64_t *state);
void gso_xoshiro_seed_from_u64(gso_xoshiro *r, uint64_t seed);
uint64_t gso_xoshiro_next_u64(gso_xoshiro *r);
uint32_t gso_xoshiro_next_u32(gso_xoshiro *r);
double gso_uniform_f64(gso_xoshiro *r);             /* Uniform::<f64>::new(0.,1.) */
float gso_uniform_f32(gso_xoshiro *r);              /* Uniform::<f32>::new(0.,1.) */
uint64_t gso_uniform_usize(gso_xoshiro *r, uint64_t m); /* Uniform::<usize>::new(0,m) */

typedef struct { double lambda, c1, c2, c3; } gso_exp01;
void gso_exp01_init(gso_exp01 *e, double lambda);
double gso_exp01_sample(const gso_exp01 *e, gso_xoshiro *r);
double gso_expm1_spec(double z); /* the frozen expm1 of the sampler's rejection branch, 0 <= z <= ln 2 */
double gso_ln_spec(double x);    /* the frozen natural logarithm of the SetSketch path, normal x > 0 */

/* ---------------- fasta.c : needletail-like parse + encode ------------------ */
/* One encoded sequence per kept record (seq mode) or one per file (block mode).
 * codes: DNA 0..3 (A,C,G,T), AA 1..20 ("ACDEFGHIKLMNPQRSTVWY").                 */
typedef struct {
    uint8_t *codes;     /* concatenated codes of all sequences */
    uint64_t *seq_off;  /* nseq+1 offsets into codes */
    uint64_t nseq;
    uint64_t nb_raw;    /* raw sequence bytes of kept records (nb_bases_file)     */
} gso_seqs;

int gso_parse_fasta(const uint8_t *bytes, uint64_t len, uint32_t data_t, uint32_t block_flag,
                    gso_seqs *out);
void gso_seqs_free(gso_seqs *s);

/* ---------------- kmer.c : k-mer values + hash closures --------------------- */
/* sig type chosen by the reference dispatch tables */
int gso_sig_type(const gso_sketch_params *p);
uint32_t gso_elem_size(const gso_sketch_params *p);
/* all hashed k-mer values of a parsed file, in order of occurrence (malloc'd) */
int gso_kmer_values(const gso_seqs *s, uint32_t data_t, uint32_t k, uint64_t **vals_out,
                    uint64_t *n_out);

/* ---------------- sketchers ------------------------------------------------- */
/* weighted set = distinct values in first-occurrence order + multiplicities */
int gso_count_kmers(const uint64_t *vals, uint64_t n, uint64_t **keys_out, double **w_out,
                    uint64_t *ndistinct_out);
/* ProbMinHash3a::hashset + get_signature : sig[m] of values (as u64) */
int gso_probminhash3a(const uint64_t *keys, const double *w, uint64_t nd, uint32_t m,
                      uint32_t val_bytes, uint32_t spec_flags, uint64_t *sig_out,
                      double *hmin_out /* optional m */);
/* OptDensMinHash::sketch over every occurrence + end_sketch : f32 sig[m] */
int gso_optdens(const uint64_t *vals, uint64_t n, uint32_t m, uint32_t spec_flags,
                float *sig_out);
/* RevOptDensMinHash: OptDens bins, reverse densification (non-empty bins push into empty ones) */
int gso_revoptdens(const uint64_t *vals, uint64_t n, uint32_t m, uint32_t spec_flags, float *sig_out);
uint32_t gso_revdens_target(uint32_t i, uint32_t a, uint32_t m);
/* SuperMinHash2: per slot the fx hash (32- or 64-bit) of the item that gave the minimum */
int gso_superminhash2(const uint64_t *vals, uint64_t n, uint32_t m, int kt32, uint64_t *sig_out);
int gso_setsketch(const uint64_t *vals, uint64_t n, uint32_t m, uint16_t *sig_out);
/* SuperMinHash (f32) over distinct values in first-occurrence order */
int gso_superminhash(const uint64_t *vals, uint64_t n, uint32_t m, float *sig_out);

/* whole path for a batch of files; nthreads = one file per thread (as the reference) */
int gso_sketch_fasta_batch(const gso_sketch_params *p, const uint8_t *bytes,
                           const uint64_t *offsets, uint32_t n, void *sig_out,
                           uint64_t *nb_bases_out, int nthreads);

/* ---------------- hamming.c ------------------------------------------------- */
float gso_hamming(const void *a, const void *b, uint32_t S, uint32_t sig_type);
void gso_hamming_matrix(const void *q, uint32_t nq, const void *c, uint32_t n, uint32_t S,
                        uint32_t sig_type, float *out, int nthreads);

/* ---------------- hnsw.c ---------------------------------------------------- */
typedef struct gso_hnsw gso_hnsw;
typedef struct {
    uint64_t d_id;
    float distance;
    uint8_t layer;
    uint8_t pad_[3];
    int32_t rank;
} gso_neighbour;

gso_hnsw *gso_hnsw_new(uint32_t max_nb_conn, uint64_t capacity, uint32_t max_layer,
                       uint32_t ef_c, double scale_modification, uint32_t sig_type, uint32_t S,
                       uint32_t extend_candidates, uint32_t keep_pruned, uint64_t level_seed);
void gso_hnsw_free(gso_hnsw *h);
/* sequential insertion in the given order; sigs are copied */
int gso_hnsw_insert(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n);
/* deterministic wave insertion (see hnsw.c): what the GPU builder is checked against;
 * wave_max = 1 is identical to gso_hnsw_insert */
int gso_hnsw_insert_waves(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n,
                          uint32_t wave_max);
/* same graph, phase A of every wave on `nthreads` host threads */
int gso_hnsw_insert_waves_mt(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n,
                             uint32_t wave_max, int nthreads);
uint32_t gso_hnsw_wave_size(uint64_t nb_point, uint32_t wave_max);
uint64_t gso_hnsw_nb_point(const gso_hnsw *h);
uint64_t gso_hnsw_nb_eval(const gso_hnsw *h); /* distance evaluations so far */
/* search one query; returns number of neighbours written (<= knbn) */
uint32_t gso_hnsw_search(gso_hnsw *h, const void *q, uint32_t knbn, uint32_t ef,
                         gso_neighbour *out, uint64_t *nb_eval_out);
void gso_hnsw_search_batch(gso_hnsw *h, const void *queries, uint32_t nq, uint32_t knbn,
                           uint32_t ef, gso_neighbour *out, uint32_t *counts,
                           uint64_t *nb_eval_out, int nthreads);
/* graph image import into an empty handle (inverse of gso_hnsw_export) */
int gso_hnsw_import(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n,
                    const uint8_t *levels, const uint32_t *ranks, const uint64_t *nbr_offsets,
                    const uint32_t *nbr_index, const float *nbr_dist, uint64_t entry_point);
/* graph image export: sizes first, then fill */
uint64_t gso_hnsw_total_lists(const gso_hnsw *h); /* sum over points of (level+1) */
uint64_t gso_hnsw_total_nbrs(const gso_hnsw *h);
void gso_hnsw_export(const gso_hnsw *h, uint8_t *levels, uint32_t *ranks, uint64_t *ids,
                     uint64_t *nbr_offsets, uint32_t *nbr_index, float *nbr_dist,
                     uint64_t *entry_point);

#ifdef __cplusplus
}
#endif
#endif
