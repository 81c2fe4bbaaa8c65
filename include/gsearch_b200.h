/*
 * gsearch_b200.h -- C ABI of libgsearch_b200.so
 *
 * The drop-in boundary for the GSearch sketch-and-search hot path on NVIDIA B200
 * (sm_100a).  Every entry point below is what a Rust `-sys` crate (cc + bindgen)
 * for this path would bind; each one names the reference interface it replaces
 * (paths relative to the reference checkout, `[U]` = upstream crate that carries
 * the arithmetic and is not vendored in the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross the ABI;
 *   - every function returns a gsb_status (0 = ok) unless stated otherwise;
 *     the failing call's message is available from gsb_last_error() (thread local);
 *     nothing ever unwinds or aborts across the ABI (reference: panic!/exit(1),
 *     src/dna/dnasketch.rs:228,285,380,463);
 *   - the caller owns all in/out buffers; the library owns opaque handles;
 *   - `*_dev` variants take DEVICE pointers (and, where stated, a CUDA stream passed
 *     as void*, 0 = default stream); the plain variants take HOST pointers and return
 *     when the results are in host memory;
 *   - a handle serialises the calls made on it (internal lock); use one handle per
 *     host thread for concurrency, as the reference clones its sketcher per worker
 *     (src/dna/dnasketch.rs:305);
 *   - there is NO CPU fallback: if no CUDA device is usable every compute call
 *     returns GSB_ERR_NO_DEVICE.
 */
#ifndef GSEARCH_B200_H
#define GSEARCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GSB_API __attribute__((visibility("default")))
#else
#define GSB_API
#endif

/* ------------------------------------------------------------------------- */
/* status codes                                                               */
/* ------------------------------------------------------------------------- */
typedef enum {
    GSB_OK = 0,
    GSB_ERR_INVALID_ARG = 1,   /* bad parameter (k, S, algo, NULL pointer ...)       */
    GSB_ERR_NO_DEVICE = 2,     /* no usable CUDA device: the library has no CPU path */
    GSB_ERR_CUDA = 3,          /* a CUDA runtime call failed                         */
    GSB_ERR_OOM = 4,           /* host or device allocation failed                   */
    GSB_ERR_BAD_INPUT = 5,     /* malformed FASTA / FASTQ (reference: exit(1), dnafiles.rs:54)*/
    GSB_ERR_UNSUPPORTED = 6,   /* valid in the reference, not (yet) built here       */
    GSB_ERR_IO = 7,            /* file dump / reload failed                          */
    GSB_ERR_CAPACITY = 8       /* index capacity or genome-length limit exceeded     */
} gsb_status;

/* thread-local message of the last failing call on this thread ("" if none) */
GSB_API const char *gsb_last_error(void);
/* library version string, e.g. "gsearch_b200 0.1.0 (sm_100a)" */
GSB_API const char *gsb_version(void);
/* number of visible CUDA devices (0 on a CPU box; never fails) */
GSB_API int gsb_device_count(void);

/* ------------------------------------------------------------------------- */
/* sketch parameters                                                          */
/*   mirrors kmerutils::sketcharg::{SeqSketcherParams, SketchAlgo, DataType}  */
/*   [U], re-exported at src/utils/parameters.rs:11 and built at              */
/*   src/bin/gsearch.rs:181-196,258-263                                       */
/* ------------------------------------------------------------------------- */
typedef enum {
    GSB_ALGO_PROB3A = 0,     /* --algo prob     ProbMinHash3a, Sig = k-mer value   */
    GSB_ALGO_SUPER = 1,      /* --algo super    SuperMinHash,  Sig = f32            */
    GSB_ALGO_OPTDENS = 2,    /* --algo optdens  OptDensMinHash, Sig = f32           */
    GSB_ALGO_REVOPTDENS = 3, /* --algo revoptdens  RevOptDensMinHash, Sig = f32      */
    GSB_ALGO_SUPER2 = 4,     /* --algo super2   SuperMinHash2, Sig = item hash u32/u64 */
    GSB_ALGO_HLL = 5         /* --algo hll      SetSketch registers, Sig = u16      */
} gsb_algo;

typedef enum { GSB_DATA_DNA = 0, GSB_DATA_AA = 1 } gsb_data_t;

/* element type of a signature (reference: the `Sig` associated type chosen by the
 * algo x k dispatch tables src/dna/dnasketch.rs:493-644, src/aa/aasketch.rs:449-552;
 * same strings as hnswio `t_name`, src/utils/reloadhnsw.rs:13-37)                 */
typedef enum { GSB_SIG_U32 = 0, GSB_SIG_U64 = 1, GSB_SIG_F32 = 2, GSB_SIG_U16 = 3 } gsb_sig_type;

/* Switches for arithmetic that lives in un-pinned upstream crates (SURVEY 8c:
 * "parity unpinned").  Default (0) is this repository's frozen SPEC; each bit
 * flips one recalled-but-unverifiable upstream choice.                           */
enum {
    GSB_SPEC_NOHASH_IDENTITY = 1u << 0, /* ProbMinHash3a seed = k-mer value (default:
                                           byte-swapped value, NoHashHasher::write) */
    GSB_SPEC_OPTDENS_F64_DRAW = 1u << 1 /* OptDens r = (f32)Uniform<f64> (default:
                                           Uniform<f32> from next_u32)              */
};

typedef struct {
    uint32_t kmer_size;   /* -k ; DNA 1..31 (32 rejected: mask overflow, dnasketch.rs:166), AA 1..12 */
    uint32_t sketch_size; /* -s ; 2..65535 (README.md:676)                                   */
    uint32_t algo;        /* gsb_algo                                                        */
    uint32_t data_t;      /* gsb_data_t (--aa)                                               */
    uint32_t block_flag;  /* --block : concatenate records, k-mers span record boundaries    */
    uint32_t spec_flags;  /* GSB_SPEC_* ; 0 = frozen default                                 */
} gsb_sketch_params;

/* ------------------------------------------------------------------------- */
/* sketcher                                                                   */
/*   replaces X::new(&SeqSketcherParams) (src/dna/dnasketch.rs:502,523,603)   */
/*   and SeqSketcherT::sketch_compressedkmer[_seqs] / SeqSketcherAAT::        */
/*   sketch_compressedkmeraa[_seqs] called at src/dna/dnasketch.rs:336,357,   */
/*   src/dna/dnarequest.rs:272,287, src/aa/aasketch.rs:313,329,               */
/*   src/aa/aarequest.rs:268,283.  The k-mer hash closures                    */
/*   (src/dna/dnasketch.rs:164-169, src/aa/aasketch.rs:156-160) are fixed     */
/*   functions of (data_t, k) and are built in, not callbacks.                */
/* ------------------------------------------------------------------------- */
typedef struct gsb_sketcher gsb_sketcher;

/* device = CUDA ordinal to run on */
GSB_API int gsb_sketcher_create(const gsb_sketch_params *params, int device, gsb_sketcher **out);
GSB_API void gsb_sketcher_destroy(gsb_sketcher *h);
/* gsb_sig_type of this sketcher's signatures, or -1 */
GSB_API int gsb_sketcher_sig_type(const gsb_sketcher *h);
/* bytes per signature element (2, 4 or 8), or 0 */
GSB_API uint32_t gsb_sketcher_elem_size(const gsb_sketcher *h);

/*
 * Sketch n FASTA / FASTQ "files" (already decompressed bytes) -> n signatures.  The first
 * byte of a file picks the format like needletail's parse_fastx_file (src/dna/dnafiles.rs:52):
 * '>' FASTA, '@' four-line FASTQ; anything else is GSB_ERR_BAD_INPUT.
 *   bytes      concatenation of the n files
 *   offsets    n+1 byte offsets into `bytes` (file i = [offsets[i], offsets[i+1]))
 *   sig_out    n * sketch_size * elem_size bytes, row i = signature of file i
 *   nb_bases_out (optional) n values: encoded bases/residues per file
 *              (reference: ItemDict.len, src/dna/dnasketch.rs:341,361)
 * Reproduces src/dna/dnafiles.rs:43-107 (seq mode) / :200-276 (block mode) or
 * src/aa/aafiles.rs, then the sketcher, per file: one signature per file in both
 * modes (seq mode = sketch_compressedkmer_seqs, k-mers never span records).
 */
GSB_API int gsb_sketch_fasta_batch(gsb_sketcher *h, const uint8_t *bytes, const uint64_t *offsets,
                                   uint32_t n, void *sig_out, uint64_t *nb_bases_out);

/* Same with the bytes and the outputs in device memory.  Work is enqueued on `stream`
 * (and on internal streams ordered after it); the call returns after `stream` has been
 * synchronised, because the rare early-stop-bound retries are decided on the host.
 * h_offsets is a HOST array of n+1 values (grid sizing is done on the host).        */
GSB_API int gsb_sketch_fasta_batch_dev(gsb_sketcher *h, const uint8_t *d_bytes,
                                       const uint64_t *h_offsets, uint32_t n, void *d_sig_out,
                                       uint64_t *d_nb_bases_out, void *stream);

/* host FASTA in (as gsb_sketch_fasta_batch), signatures and encoded lengths out to DEVICE memory:
 * e.g. straight into this rank's slice of the signature matrix that is all-gathered next        */
GSB_API int gsb_sketch_fasta_batch_to_dev(gsb_sketcher *h, const uint8_t *bytes,
                                          const uint64_t *offsets, uint32_t n, void *d_sig_out,
                                          uint64_t *d_nb_bases_out /* may be NULL */);
/* number of kernel launches issued by this handle so far (bench bookkeeping) */
GSB_API uint64_t gsb_sketcher_launch_count(const gsb_sketcher *h);
/* number of genomes whose early-stop bound had to be widened and re-run so far */
GSB_API uint64_t gsb_sketcher_retry_count(const gsb_sketcher *h);
/* number of genomes the ProbMinHash partition path handed to the general filter path so far (a
 * bucket array or a counting round overflowed: very large or very repetitive inputs) */
GSB_API uint64_t gsb_sketcher_fallback_count(const gsb_sketcher *h);
/* ProbMinHash multiplicity counting: 0 (default) = hash-partition + shared-memory counting, with the
 * filter path as the fallback for the files it flags; 1 = filter path only (also GSB_PROB_PATH=filter).
 * Both are exact; the switch exists for tests and measurements. */
GSB_API int gsb_sketcher_set_prob_path(gsb_sketcher *h, int path);
/* Optional device timing of the kernel families (CUDA events on the launching stream):
 * index 0 = K1 FASTA pack, 1 = K2 k-mer scan (the dominant kernel family), 2 = K3 slot
 * update + finalize, 3 = per-group reset; 4..6 split K2 into its filter-mark, classify and
 * exact-set kernels, 7 = the tile-summary part of K1.  enable(…, 1) also zeroes the
 * accumulators.                                                                          */
#define GSB_NB_TIMERS 8
GSB_API void gsb_sketcher_enable_timing(gsb_sketcher *h, int on);
GSB_API void gsb_sketcher_kernel_times(const gsb_sketcher *h, double *ms_out /*[GSB_NB_TIMERS]*/,
                                       uint64_t *launches_out /*[GSB_NB_TIMERS]*/);

/* ------------------------------------------------------------------------- */
/* distance                                                                   */
/*   replaces anndists::dist::DistHamming::eval [U] used through              */
/*   Hnsw::<Sig,DistHamming>::new (src/dna/dnasketch.rs:139) and directly at  */
/*   src/bin/bindash.rs:94-95:  count(a[i] != b[i]) as f32 / len as f32.      */
/* ------------------------------------------------------------------------- */

/* Scalar, signature-compatible with anndists' DistCFFI / DistCFnPtr<T> plug point [U]
 * (`extern "C" fn(*const T, *const T, len: u64) -> f32`, SURVEY 8b "distance surface"): what
 * Distance::eval(va, vb) computes for one pair.  One pair per call on device 0 (GSB_DEVICE
 * overrides): a convenience for drop-in wiring, latency bound -- hot paths use the batched forms
 * below or the index.  On error the result is NaN and gsb_last_error() holds the reason.        */
GSB_API float gsb_dist_hamming_u16(const uint16_t *a, const uint16_t *b, unsigned long long len);
GSB_API float gsb_dist_hamming_u32(const uint32_t *a, const uint32_t *b, unsigned long long len);
GSB_API float gsb_dist_hamming_u64(const uint64_t *a, const uint64_t *b, unsigned long long len);
GSB_API float gsb_dist_hamming_f32(const float *a, const float *b, unsigned long long len);

/* Batched: one query signature against n candidate signatures (row-major n x S).
 * elem = gsb_sig_type.  out[i] = hamming(q, cands[i]).  Host pointers.           */
GSB_API int gsb_hamming_batch(const void *q, const void *cands, uint32_t n, uint32_t S,
                              uint32_t sig_type, float *out, int device);
/* nq queries x n candidates -> out[nq*n] (row-major); the all-pairs shape used by
 * src/bin/bindash.rs:93-164.  Host pointers.                                     */
GSB_API int gsb_hamming_matrix(const void *queries, uint32_t nq, const void *cands, uint32_t n,
                               uint32_t S, uint32_t sig_type, float *out, int device);
/* device-pointer variants, enqueued on `stream` */
GSB_API int gsb_hamming_matrix_dev(const void *d_queries, uint32_t nq, const void *d_cands,
                                   uint32_t n, uint32_t S, uint32_t sig_type, float *d_out,
                                   void *stream);

/* ------------------------------------------------------------------------- */
/* index                                                                      */
/*   replaces hnsw_rs::Hnsw<Sig,DistHamming> [U] as used at                   */
/*   src/dna/dnasketch.rs:139-160,435 ; src/dna/dnarequest.rs:353 ;           */
/*   src/utils/dumpload.rs:31 ; src/utils/reloadhnsw.rs:41-51                 */
/* ------------------------------------------------------------------------- */
typedef struct gsb_index gsb_index;

typedef struct {
    uint32_t max_nb_connection; /* M: Hnsw::new arg 1; layer 0 keeps 2M       */
    uint64_t capacity;          /* Hnsw::new arg 2 (1_500_000 at src/bin/gsearch.rs:269)    */
    uint32_t max_layer;         /* Hnsw::new arg 3 (16 at src/dna/dnasketch.rs:139)         */
    uint32_t ef_construction;   /* Hnsw::new arg 4 (--ef)                                   */
    double scale_modification;  /* modify_level_scale (src/dna/dnasketch.rs:141)            */
    uint32_t sig_type;          /* gsb_sig_type                                             */
    uint32_t sketch_size;       /* S                                                        */
    uint32_t extend_candidates; /* set_extend_candidates(true)  src/dna/dnasketch.rs:159    */
    uint32_t keep_pruned;       /* set_keeping_pruned(false)    src/dna/dnasketch.rs:160    */
    uint64_t level_seed;        /* the reference seeds its level RNG from entropy; here it
                                   is explicit so a build is reproducible                   */
} gsb_index_params;

typedef struct {
    uint64_t d_id;    /* Neighbour.d_id   : caller's id of the data point            */
    float distance;   /* Neighbour.distance                                          */
    uint8_t layer;    /* Neighbour.p_id.0 : layer of the point                       */
    uint8_t pad_[3];
    int32_t rank;     /* Neighbour.p_id.1 : rank of the point in its layer           */
} gsb_neighbour;

GSB_API int gsb_index_create(const gsb_index_params *params, int device, gsb_index **out);
GSB_API void gsb_index_destroy(gsb_index *idx);
/* parallel_insert(&[(&Vec<Sig>, usize)]) : n signatures (row-major n x S) with ids */
GSB_API int gsb_index_insert_batch(gsb_index *idx, const void *sigs, const uint64_t *ids,
                                   uint64_t n);
/* same with the signatures already in device memory (the sketcher's output or the all-gather
 * buffer of a multi-GPU tohnsw); ids stay a host array                                     */
GSB_API int gsb_index_insert_batch_dev(gsb_index *idx, const void *d_sigs, const uint64_t *ids,
                                       uint64_t n);
/* parallel_search(&[Vec<Sig>], knbn, ef) -> Vec<Vec<Neighbour>> :
 *   out        nq * knbn entries, row i sorted by ascending distance
 *   counts_out nq values (<= knbn) : valid entries in each row
 *   nb_eval_out (optional) nq values : number of distance evaluations performed   */
GSB_API int gsb_index_search_batch(gsb_index *idx, const void *queries, uint32_t nq,
                                   uint32_t knbn, uint32_t ef, gsb_neighbour *out,
                                   uint32_t *counts_out, uint64_t *nb_eval_out);
/* same with queries and all three outputs in DEVICE memory (the query signatures straight from
 * the sketcher; results for a device-side gather).  Returns after the index stream has been
 * synchronised.                                                                            */
GSB_API int gsb_index_search_batch_dev(gsb_index *idx, const void *d_queries, uint32_t nq,
                                       uint32_t knbn, uint32_t ef, gsb_neighbour *d_out,
                                       uint32_t *d_counts_out, uint64_t *d_nb_eval_out);
/* get_nb_point() */
GSB_API uint64_t gsb_index_nb_point(const gsb_index *idx);
/* Load an externally built graph (CSR-like image) so a graph built elsewhere
 * (the oracle, or a converted hnswdump) can be searched on device.  nbr_dist (the distance of
 * every listed neighbour) may be NULL: the graph can then be searched but not extended.     */
GSB_API int gsb_index_load_graph(gsb_index *idx, const void *sigs, const uint64_t *ids,
                                 uint64_t n, const uint8_t *levels, const uint32_t *ranks,
                                 const uint64_t *nbr_offsets /* sum(level+1) + 1 prefix */,
                                 const uint32_t *nbr_index, const float *nbr_dist,
                                 uint64_t entry_point);
/* The same image back: sizes first (sum over points of level+1, total neighbours), then fill.
 * Any output pointer except nbr_offsets / entry_point may be NULL.                          */
GSB_API int gsb_index_graph_sizes(const gsb_index *idx, uint64_t *total_lists, uint64_t *total_nbrs);
GSB_API int gsb_index_export_graph(const gsb_index *idx, uint8_t *levels, uint32_t *ranks,
                                   uint64_t *ids, uint64_t *nbr_offsets, uint32_t *nbr_index,
                                   float *nbr_dist, uint64_t *entry_point);
/* the signatures of the index, n x sketch_size elements in insertion order, to host memory (with the
 * graph image above this is everything a converter to another dump format needs)                 */
GSB_API int gsb_index_export_signatures(const gsb_index *idx, void *sigs_out);
/* Points inserted together by gsb_index_insert_batch (the reference inserts with one rayon
 * task per point, src/dna/dnasketch.rs:435; here a wave of at most wave_max points searches
 * the graph as it was before the wave).  Default = two per SM; 1 = sequential insertion. */
GSB_API int gsb_index_set_wave_max(gsb_index *idx, uint32_t wave_max);
/* file_dump(dir, basename) / HnswIo::load_hnsw: <basename>.hnsw.graph + <basename>.hnsw.data in
 * this library's own layout; hnswio byte compatibility is not claimed         */
GSB_API int gsb_index_dump(const gsb_index *idx, const char *dir, const char *basename);
GSB_API int gsb_index_load(gsb_index *idx, const char *dir, const char *basename);

/* ------------------------------------------------------------------------- */
/* device / pinned-host buffers for hosts without a CUDA binding of their own  */
/* ------------------------------------------------------------------------- */
GSB_API int gsb_device_malloc(int device, uint64_t bytes, void **out);
GSB_API void gsb_device_free(int device, void *p);
GSB_API int gsb_memcpy_h2d(int device, void *d_dst, const void *src, uint64_t bytes);
GSB_API int gsb_memcpy_d2h(int device, void *dst, const void *d_src, uint64_t bytes);
GSB_API int gsb_host_alloc_pinned(uint64_t bytes, void **out);
GSB_API void gsb_host_free_pinned(void *p);

/* ------------------------------------------------------------------------- */
/* multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch                 */
/*   The reference is one process (src/dna/dnasketch.rs:421-435: sketch every  */
/*   file, then insert all signatures).  Here genomes shard by rank, finished   */
/*   signatures are exchanged with one all-gather, and HNSW insertion is        */
/*   sharded by point inside every wave.  NCCL is loaded at run time (dlopen);  */
/*   single-GPU consumers never need it.                                        */
/* ------------------------------------------------------------------------- */
#define GSB_COMM_ID_BYTES 128
typedef struct gsb_comm gsb_comm;
/* rank 0 draws the id and ships it to the other ranks (pipe, file, MPI, torch.distributed ...) */
GSB_API int gsb_comm_unique_id(uint8_t *id_out /* GSB_COMM_ID_BYTES */);
/* collective: every rank calls it with the same id and nranks */
GSB_API int gsb_comm_create(const uint8_t *id, int nranks, int rank, int device, gsb_comm **out);
GSB_API void gsb_comm_destroy(gsb_comm *c);
GSB_API int gsb_comm_rank(const gsb_comm *c);
GSB_API int gsb_comm_size(const gsb_comm *c);
/* all-gather of `bytes_per_rank` device bytes per rank into d_recv (nranks * bytes_per_rank, rank r at
 * offset r * bytes_per_rank).  In place when d_send == d_recv + rank * bytes_per_rank: the sketcher
 * (gsb_sketch_fasta_batch_dev) can write its signatures straight into its slice of the replicated
 * matrix.  stream NULL: the communicator's own stream, synchronised before returning.         */
GSB_API int gsb_comm_all_gather(gsb_comm *c, const void *d_send, void *d_recv, uint64_t bytes_per_rank,
                                void *stream);
GSB_API int gsb_comm_broadcast(gsb_comm *c, void *d_buf, uint64_t bytes, int root, void *stream);
/* all-gather of row-sharded results into GLOBAL unit order.  Unit i (genome or query) lives on rank
 * i mod nranks as that rank's local row i / nranks; d_local holds rows_per_rank rows of row_bytes,
 * d_tmp nranks * rows_per_rank rows, d_out n_total rows: row i = unit i on every rank.            */
GSB_API int gsb_comm_all_gather_rows(gsb_comm *c, const void *d_local, uint64_t rows_per_rank,
                                     uint64_t row_bytes, uint64_t n_total, void *d_tmp, void *d_out,
                                     void *stream);
/* parallel_insert over the GPUs of `comm`: EVERY rank calls it with the same signatures (host or
 * device pointer; normally the all-gathered matrix), ids and index parameters.  Inside each wave a
 * rank searches / selects for its slice of the points (all the distance evaluations), the selections
 * are all-gathered and every rank applies the same link updates: all replicas end bit-identical to
 * the graph gsb_index_insert_batch builds on one GPU with the same wave_max.                  */
GSB_API int gsb_index_insert_batch_sharded(gsb_index *idx, gsb_comm *comm, const void *sigs,
                                           const uint64_t *ids, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* GSEARCH_B200_H */
