// api_sketch.cu -- C-ABI of the sketcher (gsb_sketcher_*, gsb_sketch_fasta_batch*).
//
// Host-side orchestration only: sizes grids, owns device workspaces, launches the K1/K2/K3
// kernels of fasta_pack.cuh / sketch_kernels.cuh on the caller's stream and handles the
// (rare) early-stop-bound retries.  No arithmetic of the path runs on the CPU.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <new>
#include <vector>

#include "api_common.h"
#include "prob_partition.cuh"
#include "sketch_kernels.cuh"

using namespace gsb;

namespace {

constexpr int kMaxSlots = 16;  // two sets (one per group stream) of up to 8 genomes
// genomes per group on the prob path; two groups are in flight, one per group stream, so that
// the atomic-bound mark kernel of one overlaps the ALU-bound classify kernel of the other (filters
// sized for L2); GSB_PROB_SLOTS overrides (1..8)
// filter slots per input byte; GSB_PROB_LOAD overrides (2 .. 32)
static double prob_load() {
    static double v = 0;
    if (v == 0) {
        const char *e = getenv("GSB_PROB_LOAD");
        v = e ? atof(e) : 8.0;
        if (v < 2.0) v = 2.0;
        if (v > 32.0) v = 32.0;
    }
    return v;
}
// resident CTAs per SM of the persistent scan kernels (GSB_MARK_R / GSB_CLS_R override)
static int env_int(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    int v = e ? atoi(e) : dflt;
    return v < lo ? lo : (v > hi ? hi : v);
}
static int prob_slots() {
    static int v = 0;
    if (!v) {
        const char *e = getenv("GSB_PROB_SLOTS");
        v = e ? atoi(e) : 6;  // measured: 4 -> 8 144, 6 -> 8 457, 8 -> 8 384 genomes/s (5 Mbp, k=21, s=18000)
        if (v < 1) v = 1;
        if (v > kMaxSlots / 2) v = kMaxSlots / 2;
    }
    return v;
}

struct ProbSlot {
    DevBuf bitmap, table, cnt, list, ovf, misc, hmin, sigw;
    DevBuf buckets, cursor, slot2;  // partition path
    size_t cnt_dirty = 0;
};

// GSB_PROB_PATH=filter (or gsb_sketcher_set_prob_path(h, 1)) forces the round-1 filter path (mark /
// classify / exact set) for every file; the default is the partition path with the filter path as
// the fallback for the files it flags
static int prob_path_from_env() {
    const char *e = getenv("GSB_PROB_PATH");
    return (e && !strcmp(e, "filter")) ? 1 : 0;
}

// Small host->device descriptor copies go through a kernel that reads the PINNED host buffer
// directly (unified addressing) instead of cudaMemcpyAsync: a copy-engine transfer would queue
// behind the bulk H2D copy of the next sub-batch and serialise the pipeline of
// gsb_sketch_fasta_batch.
__global__ void k_pull_words(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src_pinned, size_t nwords) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src_pinned[i];
}
static cudaError_t pull_small(void *dst, const void *src_pinned, size_t bytes, cudaStream_t st) {
    const size_t nw = (bytes + 3) / 4;  // buffers are allocated with slack: rounding up is safe
    if (!nw) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<size_t>((nw + 255) / 256, 64);
    k_pull_words<<<grid, 256, 0, st>>>((uint32_t *)dst, (const uint32_t *)src_pinned, nw);
    return cudaGetLastError();
}

// per group: clear the hash sets, reset cursors and slot state, compute the bounds
__global__ void __launch_bounds__(256)
k_prob_reset(const ProbJob *__restrict__ jobs, uint32_t njobs, const FileResult *__restrict__ res,
             SketchConsts sc, ProbBound *__restrict__ bound, uint32_t *__restrict__ overflow,
             uint32_t *__restrict__ retry) {
    const uint32_t j = blockIdx.y;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    // clear the counters left by the previous genome of this slot (entries of kind "repeated")
    {
        const uint32_t np = *job.prev_n;
        for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < np; e += gridDim.x * blockDim.x) {
            const ListEntry le = job.list[e];
            if (le.kind == 1) job.cnt[le.slot] = 0;
        }
    }
    uint4 *t4 = reinterpret_cast<uint4 *>(job.bitmap);
    const size_t n4 = ((size_t)job.nslot1 / 16 + 3) / 4 + 1;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        t4[i] = z;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < sc.m; k += gridDim.x * blockDim.x) {
        job.hmin[k] = 0x7FEFFFFFFFFFFFFFull;  // f64::MAX
        job.sigw[k] = ~0ull;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *job.list_n = 0;
        *job.ovf_n = 0;
        *job.n_coll = 0;
        overflow[j] = 0;
        retry[job.file] = 0;  // k3_prob_finalize accumulates into it
        const uint32_t N = res[job.file].nsym;
        const uint32_t nk = N >= sc.k ? N - sc.k + 1 : 0;
        ProbBound b;
        if (nk == 0) {
            b.T = 0.0;
            b.uT = 0;
        } else {
            b.T = job.tmult * ((double)sc.m / (double)nk) * sc.lnm8;
            b.uT = b.T >= 1.0 ? (1ull << 52) : (uint64_t)(b.T * 4503599627370496.0) + 2;
        }
        bound[j] = b;
    }
}

__global__ void k_dens_reset(uint32_t *bins, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bins[i] = kDensLargeBits;
}

}  // namespace

struct gsb_sketcher {
    std::mutex mu;  // a handle serialises its calls (the reference clones its sketcher per worker)
    gsb_sketch_params p;
    int device = 0;
    int sig_type = 0;
    uint32_t elem = 4;
    bool kt32 = false;
    SketchConsts sc;
    cudaStream_t stream = nullptr;
    DevBuf d_files, d_tile_prefix, d_tc4, d_ttrans, d_tnrec, d_tstate, d_tbase, d_trecbase, d_res,
        d_packed, d_bounds, d_misc, d_retry, d_jobs, d_chunk_prefix, d_bound, d_overflow, d_bins, d_fqflag;
    PinBuf h_files, h_tile_prefix, h_jobs, h_chunk_prefix, h_retry, h_overflow;
    ProbSlot slot[kMaxSlots];
    DevBuf d_bytes, d_sig, d_nb;
    uint64_t launches = 0, retries = 0, fallbacks = 0;
    int prob_path = 0;  // 0 = partition path with filter fallback, 1 = filter path only
    size_t dens_pass_files = 0;  // files of the current run_dens pass (RevOptDens scratch offset)
    std::vector<uint32_t> pending_fallback;  // files flagged by the partition path, waiting for the filter path
    std::vector<double> pending_fallback_t;
    // optional per-kernel-family timing (bench.py's roofline): CUDA events around launches
    bool timing = false;
    struct Span {
        int cat;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> ev_pool;
    double cat_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t cat_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // host-pointer entry point: H2D of the next sub-batch overlaps the kernels of the current one
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_h2d;       // one per H2D chunk of the current host call
    std::vector<uint32_t> h2d_end;         // chunk c holds the files [h2d_end[c-1], h2d_end[c])
    bool h2d_active = false;               // batch_dev must wait for the chunk events
    // host call with a pinned output buffer: the signatures of a genome group start their way back
    // as soon as the group is final (first pass), under the kernels of the later groups
    uint8_t *early_out = nullptr;          // host signatures, or null
    const uint8_t *early_src = nullptr;    // device signatures of the call
    std::vector<cudaEvent_t> ev_grp;
    uint32_t file_base = 0;  // offset added to file indices in messages
    cudaStream_t gstream[2] = {nullptr, nullptr};  // group streams of the prob path
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_k1;  // one per group: its files are packed
    int nsm = 148;
};

namespace {
enum { CAT_K1 = 0, CAT_K2 = 1, CAT_K3 = 2, CAT_RESET = 3, CAT_K2_MARK = 4, CAT_K2_CLASSIFY = 5, CAT_K2_EXACT = 6,
       CAT_K1_SUMMARY = 7, CAT_N = 8 };
struct Timed {
    gsb_sketcher *h;
    cudaStream_t st;
    bool on;
    gsb_sketcher::Span sp;
    Timed(gsb_sketcher *h_, int cat, cudaStream_t st_) : h(h_), st(st_), on(h_->timing) {
        if (!on) return;
        sp.cat = cat;
        for (cudaEvent_t *e : {&sp.a, &sp.b}) {
            if (h->ev_pool.empty()) {
                cudaEventCreate(e);
            } else {
                *e = h->ev_pool.back();
                h->ev_pool.pop_back();
            }
        }
        cudaEventRecord(sp.a, st);
    }
    ~Timed() {
        if (!on) return;
        cudaEventRecord(sp.b, st);
        h->spans.push_back(sp);
    }
};
void collect_spans(gsb_sketcher *h) {  // call after the stream has been synchronised
    for (auto &sp : h->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            h->cat_ms[sp.cat] += ms;
            h->cat_n[sp.cat] += 1;
        }
        h->ev_pool.push_back(sp.a);
        h->ev_pool.push_back(sp.b);
    }
    h->spans.clear();
}
}  // namespace

static int sig_type_of(const gsb_sketch_params &p) {
    if (p.algo == GSB_ALGO_PROB3A || p.algo == GSB_ALGO_SUPER2) {  // SuperHash2Sketch<Kmer, u32 | u64, Fx>: dnasketch.rs:575-599
        if (p.data_t == GSB_DATA_DNA) return (p.kmer_size <= 14 || p.kmer_size == 16) ? GSB_SIG_U32 : GSB_SIG_U64;
        return p.kmer_size <= 6 ? GSB_SIG_U32 : GSB_SIG_U64;
    }
    if (p.algo == GSB_ALGO_HLL) return GSB_SIG_U16;  // HyperLogLogSketch<Kmer, u16>: dnasketch.rs:541-573
    return GSB_SIG_F32;
}

extern "C" int gsb_sketcher_create(const gsb_sketch_params *params, int device, gsb_sketcher **out) {
    if (!params || !out) {
        set_error("gsb_sketcher_create: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    const gsb_sketch_params &p = *params;
    if (p.data_t != GSB_DATA_DNA && p.data_t != GSB_DATA_AA) {
        set_error("data_t must be DNA(0) or AA(1)");
        return GSB_ERR_INVALID_ARG;
    }
    const uint32_t kmax = p.data_t == GSB_DATA_DNA ? 31u : 12u;
    if (p.kmer_size < 1 || p.kmer_size > kmax) {
        set_error("kmer_size %u out of range 1..%u (k=32 overflows the reference's mask, "
                  "src/dna/dnasketch.rs:166)", p.kmer_size, kmax);
        return GSB_ERR_INVALID_ARG;
    }
    if (p.sketch_size < 2 || p.sketch_size > 65535) {
        set_error("sketch_size %u out of range 2..65535", p.sketch_size);
        return GSB_ERR_INVALID_ARG;
    }
    if (p.algo > GSB_ALGO_HLL) {
        set_error("unknown algo %u", p.algo);
        return GSB_ERR_INVALID_ARG;
    }
    int rc = check_device(device);
    if (rc) return rc;
    gsb_sketcher *h = new (std::nothrow) gsb_sketcher();
    if (!h) return GSB_ERR_OOM;
    h->p = p;
    h->device = device;
    h->sig_type = sig_type_of(p);
    h->elem = h->sig_type == GSB_SIG_U64 ? 8 : (h->sig_type == GSB_SIG_U16 ? 2 : 4);
    h->kt32 = p.data_t == GSB_DATA_DNA ? (p.kmer_size <= 14 || p.kmer_size == 16) : (p.kmer_size <= 6);
    SketchConsts &sc = h->sc;
    sc.k = p.kmer_size;
    sc.m = p.sketch_size;
    const uint64_t m = sc.m;
    sc.zone = UINT64_MAX - ((UINT64_MAX - m + 1) % m);
    sc.lnm8 = log((double)m) + 8.0;
    const double lambda = log((double)m / (double)(m - 1));
    sc.e01.lambda = lambda;
    sc.e01.c1 = expm1(lambda) / lambda;
    sc.e01.c2 = log(2.0 / (1.0 + exp(-lambda))) / lambda;
    sc.e01.c3 = (1.0 - exp(-lambda)) / lambda;
    // c1*u >= 1 needs u >= 1/c1; anything from a little below that goes to the exact path
    sc.u_slow = (uint64_t)(4503599627370496.0 / sc.e01.c1) - 8;
    sc.spec_flags = p.spec_flags;
    h->prob_path = prob_path_from_env();
    cudaSetDevice(device);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
        delete h;
        return GSB_ERR_CUDA;
    }
    for (auto &g : h->gstream) cudaStreamCreateWithFlags(&g, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    for (auto &e2 : h->ev_join) cudaEventCreateWithFlags(&e2, cudaEventDisableTiming);
    cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, device);
    *out = h;
    return GSB_OK;
}

extern "C" void gsb_sketcher_destroy(gsb_sketcher *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&h->d_files, &h->d_tile_prefix, &h->d_tc4, &h->d_ttrans, &h->d_tnrec, &h->d_tstate,
                      &h->d_tbase, &h->d_trecbase, &h->d_res, &h->d_packed, &h->d_bounds, &h->d_misc,
                      &h->d_retry, &h->d_jobs, &h->d_chunk_prefix, &h->d_bound, &h->d_overflow, &h->d_bins, &h->d_fqflag,
                      &h->d_bytes, &h->d_sig, &h->d_nb};
    for (DevBuf *b : bufs) b->release();
    for (auto &s : h->slot) {
        s.table.release();
        s.cnt.release();
        s.bitmap.release();
        s.list.release();
        s.ovf.release();
        s.misc.release();
        s.hmin.release();
        s.sigw.release();
        s.buckets.release();
        s.cursor.release();
        s.slot2.release();
    }
    PinBuf *pb[] = {&h->h_files, &h->h_tile_prefix, &h->h_jobs, &h->h_chunk_prefix, &h->h_retry, &h->h_overflow};
    for (PinBuf *b : pb) b->release();
    for (auto &sp : h->spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (auto g : h->gstream)
        if (g) cudaStreamDestroy(g);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto e2 : h->ev_k1) cudaEventDestroy(e2);
    for (auto e2 : h->ev_join)
        if (e2) cudaEventDestroy(e2);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    for (cudaEvent_t e : h->ev_h2d) cudaEventDestroy(e);
    delete h;
}

extern "C" int gsb_sketcher_sig_type(const gsb_sketcher *h) { return h ? h->sig_type : -1; }
extern "C" uint32_t gsb_sketcher_elem_size(const gsb_sketcher *h) { return h ? h->elem : 0; }
extern "C" uint64_t gsb_sketcher_launch_count(const gsb_sketcher *h) { return h ? h->launches : 0; }
extern "C" uint64_t gsb_sketcher_retry_count(const gsb_sketcher *h) { return h ? h->retries : 0; }
extern "C" uint64_t gsb_sketcher_fallback_count(const gsb_sketcher *h) { return h ? h->fallbacks : 0; }
extern "C" int gsb_sketcher_set_prob_path(gsb_sketcher *h, int path) {
    if (!h || path < 0 || path > 1) {
        set_error("gsb_sketcher_set_prob_path: path must be 0 (partition, filter as fallback) or 1 (filter only)");
        return GSB_ERR_INVALID_ARG;
    }
    std::lock_guard<std::mutex> lock(h->mu);
    h->prob_path = path;
    return GSB_OK;
}
extern "C" void gsb_sketcher_enable_timing(gsb_sketcher *h, int on) {
    if (!h) return;
    h->timing = on != 0;
    for (int i = 0; i < CAT_N; i++) {
        h->cat_ms[i] = 0;
        h->cat_n[i] = 0;
    }
}
extern "C" void gsb_sketcher_kernel_times(const gsb_sketcher *h, double *ms_out, uint64_t *n_out) {
    for (int i = 0; i < CAT_N; i++) {
        if (ms_out) ms_out[i] = h ? h->cat_ms[i] : 0.0;
        if (n_out) n_out[i] = h ? h->cat_n[i] : 0;
    }
}

namespace {

// K1 for the files [f0, f0 + nf) of the batch (tiles [tile0, tile0 + ntiles))
template <int DATA_T, bool SEQ_SEP>
void launch_k1(gsb_sketcher *h, const uint8_t *d_bytes, uint64_t total, uint32_t f0, uint32_t nf, uint32_t tile0,
               uint32_t ntiles, bool want_bounds, uint32_t bd_cap, cudaStream_t st) {
    Timed t_(h, CAT_K1, st);
    const FileDesc *files = h->d_files.as<FileDesc>() + f0;
    const uint32_t *tprefix = h->d_tile_prefix.as<uint32_t>() + f0;
    FileResult *res = h->d_res.as<FileResult>() + f0;
    {
        Timed ts_(h, CAT_K1_SUMMARY, st);
        k1a_tile_summary<DATA_T, SEQ_SEP><<<ntiles, kK1Threads, 0, st>>>(
            d_bytes, total, files, tprefix, nf, h->d_tc4.as<uint64_t>(), h->d_ttrans.as<uint8_t>(),
            h->d_tnrec.as<uint16_t>(), tile0, h->d_fqflag.as<uint32_t>() + f0);
    }
    k1b_resolve<<<nf, kK1bThreads, 0, st>>>(
        files, nf, d_bytes, h->d_tc4.as<uint64_t>(), h->d_ttrans.as<uint8_t>(), h->d_tnrec.as<uint16_t>(),
        h->d_tstate.as<uint8_t>(), h->d_tbase.as<uint32_t>(), h->d_trecbase.as<uint32_t>(), res,
        h->d_misc.as<uint32_t>(), bd_cap, want_bounds ? 1 : 0, (DATA_T == 1 && SEQ_SEP) ? 1 : 0,
        h->d_fqflag.as<uint32_t>() + f0);
    k1c_pack<DATA_T, SEQ_SEP><<<ntiles, kK1Threads, 0, st>>>(
        d_bytes, total, files, tprefix, nf, h->d_tstate.as<uint8_t>(), h->d_tbase.as<uint32_t>(),
        h->d_trecbase.as<uint32_t>(), res, DATA_T == 0 ? h->d_packed.as<uint32_t>() : nullptr,
        DATA_T == 1 ? h->d_packed.as<uint8_t>() : nullptr, want_bounds ? h->d_bounds.as<uint32_t>() : nullptr,
        tile0);
    h->launches += 3;
}

// what K1 needs to know about the batch (set by gsb_sketch_fasta_batch_dev before the first pass)
struct K1Plan {
    const uint8_t *d_bytes = nullptr;
    uint64_t total = 0;
    uint32_t bd_cap = 0;
    bool pending = false;  // K1 has not run yet for this batch
};

// host entry point only: the bytes of files <= last arrive on the copy stream, chunk by chunk
cudaError_t wait_files_ready(gsb_sketcher *h, cudaStream_t st, uint32_t last) {
    if (!h->h2d_active) return cudaSuccess;
    for (size_t c = 0; c < h->h2d_end.size(); c++)
        if (last < h->h2d_end[c]) return cudaStreamWaitEvent(st, h->ev_h2d[c], 0);  // copies are in order
    return cudaSuccess;
}

void launch_k1_range(gsb_sketcher *h, const K1Plan &kp, uint32_t f0, uint32_t nf, cudaStream_t st) {
    const uint32_t *htp = h->h_tile_prefix.as<uint32_t>();
    const uint32_t tile0 = htp[f0], ntiles = htp[f0 + nf] - tile0;
    if (!ntiles) return;
    const bool dna = h->p.data_t == GSB_DATA_DNA, seq = !h->p.block_flag;
    if (dna) launch_k1<0, false>(h, kp.d_bytes, kp.total, f0, nf, tile0, ntiles, seq, kp.bd_cap, st);
    else if (seq) launch_k1<1, true>(h, kp.d_bytes, kp.total, f0, nf, tile0, ntiles, false, 0, st);
    else launch_k1<1, false>(h, kp.d_bytes, kp.total, f0, nf, tile0, ntiles, false, 0, st);
}

template <class Src, typename KT>
void launch_prob_group(gsb_sketcher *h, uint32_t joff, uint32_t njobs, uint32_t nchunks, uint32_t cpoff,
                       bool dna, bool want_bounds, void *d_sig, uint64_t *d_nb, cudaStream_t st) {
    const ProbJob *jobs = h->d_jobs.as<ProbJob>() + joff;
    ProbBound *bound = h->d_bound.as<ProbBound>() + joff;
    uint32_t *ovf = h->d_overflow.as<uint32_t>() + joff;
    const FileResult *res = h->d_res.as<FileResult>();
    {
        Timed t_(h, CAT_RESET, st);
        k_prob_reset<<<dim3(296, njobs), 256, 0, st>>>(jobs, njobs, res, h->sc, bound, ovf, h->d_retry.as<uint32_t>());
    }
    if (nchunks) {
        {
            Timed tm_(h, CAT_K2_MARK, st);
            const uint32_t grid = std::min<uint32_t>(nchunks, (uint32_t)(h->nsm * env_int("GSB_MARK_R", 2, 1, 8)));
            k2_prob_mark<Src, KT><<<grid, kK2Threads, 0, st>>>(
                jobs, h->d_chunk_prefix.as<uint32_t>() + cpoff, njobs, h->d_files.as<FileDesc>(), res,
                dna ? h->d_packed.as<uint32_t>() : nullptr, dna ? nullptr : h->d_packed.as<uint8_t>(),
                want_bounds ? h->d_bounds.as<uint32_t>() : nullptr, h->sc, nchunks);
        }
        cudaFuncSetAttribute(k2_prob_classify<Src, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(kStageCap * sizeof(ListEntry)));  // per device: every launch
        {
            Timed tc_(h, CAT_K2_CLASSIFY, st);
            const uint32_t grid = std::min<uint32_t>(nchunks, (uint32_t)(h->nsm * env_int("GSB_CLS_R", 3, 1, 8)));
            k2_prob_classify<Src, KT><<<grid, kK2Threads, kStageCap * sizeof(ListEntry), st>>>(
                jobs, h->d_chunk_prefix.as<uint32_t>() + cpoff, njobs, h->d_files.as<FileDesc>(), res,
                dna ? h->d_packed.as<uint32_t>() : nullptr, dna ? nullptr : h->d_packed.as<uint8_t>(),
                want_bounds ? h->d_bounds.as<uint32_t>() : nullptr, bound, h->sc, ovf, nchunks);
        }
        {
            Timed te_(h, CAT_K2_EXACT, st);
            k2_prob_mid<<<dim3(74, njobs), 256, 0, st>>>(jobs, njobs);
            k2_prob_overflow<Src, KT><<<dim3(148, njobs), 256, 0, st>>>(
                jobs, njobs, h->d_files.as<FileDesc>(), res, dna ? h->d_packed.as<uint32_t>() : nullptr,
                dna ? nullptr : h->d_packed.as<uint8_t>(), h->sc, ovf);
        }
    }
    Timed t3_(h, CAT_K3, st);
    k3_prob_points<KT, 0><<<dim3(148, njobs), 256, 0, st>>>(jobs, njobs, bound, h->sc);
    k3_prob_points<KT, 1><<<dim3(148, njobs), 256, 0, st>>>(jobs, njobs, bound, h->sc);
    if (h->elem == 8)
        k3_prob_finalize<uint64_t><<<dim3(kFinParts, njobs), 256, 0, st>>>(jobs, njobs, bound, res, h->sc, (uint64_t *)d_sig,
                                                          d_nb, h->d_retry.as<uint32_t>());
    else
        k3_prob_finalize<uint32_t><<<dim3(kFinParts, njobs), 256, 0, st>>>(jobs, njobs, bound, res, h->sc, (uint32_t *)d_sig,
                                                          d_nb, h->d_retry.as<uint32_t>());
    h->launches += nchunks ? 8 : 4;
}

static PartConsts part_consts(const gsb_sketcher *h) {
    return make_part_consts(h->p.data_t == GSB_DATA_DNA ? 2 * h->sc.k : 5 * h->sc.k);
}

// partition path: reset, partition, count, then the same exact replay (K3) as the filter path
template <class Src, typename KT, typename KEY, int KBITS = 0>
void launch_part_group(gsb_sketcher *h, uint32_t joff, uint32_t njobs, uint32_t nchunks, uint32_t cpoff, bool dna,
                       bool want_bounds, void *d_sig, uint64_t *d_nb, cudaStream_t st) {
    const ProbJob *jobs = h->d_jobs.as<ProbJob>() + joff;
    ProbBound *bound = h->d_bound.as<ProbBound>() + joff;
    uint32_t *ovf = h->d_overflow.as<uint32_t>() + joff;
    const FileResult *res = h->d_res.as<FileResult>();
    const PartConsts pc = part_consts(h);
    {
        Timed t_(h, CAT_RESET, st);
        k2p_reset<<<dim3(8, njobs), 256, 0, st>>>(jobs, njobs, res, h->sc, bound, ovf, h->d_retry.as<uint32_t>());
    }
    if (nchunks) {
        const int psm = (int)part_smem_bytes<KEY>(), csm = (int)count_smem_bytes<KEY>();
        cudaFuncSetAttribute(k2p_partition<Src, KT, KEY, KBITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm);
        cudaFuncSetAttribute(k2p_count<KT, KEY>, cudaFuncAttributeMaxDynamicSharedMemorySize, csm);
        {
            Timed tm_(h, CAT_K2_MARK, st);
            const int per_sm = psm <= 73 * 1024 ? 3 : (psm <= 110 * 1024 ? 2 : 1);
            const uint32_t grid = std::min<uint32_t>(nchunks, (uint32_t)(h->nsm * env_int("GSB_PART_R", per_sm, 1, 4)));
            k2p_partition<Src, KT, KEY, KBITS><<<grid, kK2Threads, psm, st>>>(
                jobs, h->d_chunk_prefix.as<uint32_t>() + cpoff, njobs, h->d_files.as<FileDesc>(), res,
                dna ? h->d_packed.as<uint32_t>() : nullptr, dna ? nullptr : h->d_packed.as<uint8_t>(),
                want_bounds ? h->d_bounds.as<uint32_t>() : nullptr, bound, h->sc, pc, ovf, nchunks);
        }
        {
            Timed tc_(h, CAT_K2_CLASSIFY, st);
            CountArgs ca;
            const ProbJob *hj = h->h_jobs.as<ProbJob>() + joff;  // host copy of the descriptors
            for (uint32_t i = 0; i < kMaxGroupJobs; i++) {
                const ProbJob &pj = hj[i < njobs ? i : 0];
                ca.buckets[i] = pj.buckets;
                ca.cursor[i] = pj.cursor;
                ca.slot2[i] = pj.slot2;
                ca.cap_g[i] = pj.cap_g;
            }
            k2p_count<KT, KEY><<<dim3(kNB, njobs), kCThreads, csm, st>>>(ca, njobs, bound, h->sc, pc, ovf);
        }
    }
    Timed t3_(h, CAT_K3, st);
    if (h->elem == 8)
        k3p_finalize128<uint64_t><<<dim3(24, njobs), 256, 0, st>>>(jobs, njobs, bound, res, h->sc, (uint64_t *)d_sig, d_nb,
                                                                  h->d_retry.as<uint32_t>());
    else
        k3p_finalize128<uint32_t><<<dim3(24, njobs), 256, 0, st>>>(jobs, njobs, bound, res, h->sc, (uint32_t *)d_sig, d_nb,
                                                                  h->d_retry.as<uint32_t>());
    h->launches += nchunks ? 4 : 2;
}

template <class Src, typename KT>
void launch_part_group_key(gsb_sketcher *h, uint32_t joff, uint32_t njobs, uint32_t nchunks, uint32_t cpoff, bool dna,
                           bool want_bounds, void *d_sig, uint64_t *d_nb, cudaStream_t st) {
    // a key (k-mer bits below the bucket bits) + the light flag must fit the key word
    if constexpr (sizeof(KT) == 4) {
        launch_part_group<Src, KT, uint32_t>(h, joff, njobs, nchunks, cpoff, dna, want_bounds, d_sig, d_nb, st);
    } else {
        if (part_consts(h).keybits <= 31)
            launch_part_group<Src, KT, uint32_t>(h, joff, njobs, nchunks, cpoff, dna, want_bounds, d_sig, d_nb, st);
        else
            launch_part_group<Src, KT, uint64_t>(h, joff, njobs, nchunks, cpoff, dna, want_bounds, d_sig, d_nb, st);
    }
}

// DNA: the k values of the BASELINE configurations get kernels with k fixed at compile time
template <typename KT>
void launch_part_group_dna(gsb_sketcher *h, uint32_t joff, uint32_t njobs, uint32_t nchunks, uint32_t cpoff,
                           bool want_bounds, void *d_sig, uint64_t *d_nb, cudaStream_t st) {
    if constexpr (sizeof(KT) == 4) {
        if (h->sc.k == 16)
            return launch_part_group<SrcDNA<uint32_t, 16>, uint32_t, uint32_t, 32>(h, joff, njobs, nchunks, cpoff, true,
                                                                                   want_bounds, d_sig, d_nb, st);
    } else {
        if (h->sc.k == 21)
            return launch_part_group<SrcDNA<uint64_t, 21>, uint64_t, uint32_t, 42>(h, joff, njobs, nchunks, cpoff, true,
                                                                                   want_bounds, d_sig, d_nb, st);
        if (h->sc.k == 31)
            return launch_part_group<SrcDNA<uint64_t, 31>, uint64_t, uint64_t, 62>(h, joff, njobs, nchunks, cpoff, true,
                                                                                   want_bounds, d_sig, d_nb, st);
    }
    launch_part_group_key<SrcDNA<KT>, KT>(h, joff, njobs, nchunks, cpoff, true, want_bounds, d_sig, d_nb, st);
}

template <class Src, typename KT>
void launch_dens(gsb_sketcher *h, uint32_t joff, uint32_t njobs, uint32_t nchunks, uint32_t cpoff, bool dna,
                 bool want_bounds, void *d_sig, uint64_t *d_nb, cudaStream_t st) {
    const DensJob *jobs = h->d_jobs.as<DensJob>() + joff;
    const FileResult *res = h->d_res.as<FileResult>();
    const bool super2 = h->p.algo == GSB_ALGO_SUPER2;
    const bool hll = h->p.algo == GSB_ALGO_HLL;
    if (nchunks) {
        Timed t_(h, CAT_K2_MARK, st);
        const uint32_t grid = std::min<uint32_t>(nchunks, (uint32_t)(h->nsm * 4));
        if (hll)
            k2_hll<Src, KT><<<grid, kK2Threads, 0, st>>>(
                jobs, h->d_chunk_prefix.as<uint32_t>() + cpoff, njobs, h->d_files.as<FileDesc>(), res,
                dna ? h->d_packed.as<uint32_t>() : nullptr, dna ? nullptr : h->d_packed.as<uint8_t>(),
                want_bounds ? h->d_bounds.as<uint32_t>() : nullptr, h->sc, nchunks);
        else if (super2)
            k2_super2<Src, KT><<<grid, kK2Threads, 0, st>>>(
                jobs, h->d_chunk_prefix.as<uint32_t>() + cpoff, njobs, h->d_files.as<FileDesc>(), res,
                dna ? h->d_packed.as<uint32_t>() : nullptr, dna ? nullptr : h->d_packed.as<uint8_t>(),
                want_bounds ? h->d_bounds.as<uint32_t>() : nullptr, h->sc, nchunks);
        else
            k2_optdens<Src, KT><<<grid, kK2Threads, 0, st>>>(
                jobs, h->d_chunk_prefix.as<uint32_t>() + cpoff, njobs, h->d_files.as<FileDesc>(), res,
                dna ? h->d_packed.as<uint32_t>() : nullptr, dna ? nullptr : h->d_packed.as<uint8_t>(),
                want_bounds ? h->d_bounds.as<uint32_t>() : nullptr, h->sc, nchunks);
    }
    Timed t3_(h, CAT_K3, st);
    if (hll) {
        k3_hll_finalize<<<njobs, 256, 0, st>>>(jobs, njobs, res, h->sc, (uint16_t *)d_sig, d_nb, h->d_retry.as<uint32_t>());
    } else if (super2) {
        if (h->elem == 8)
            k3_super2_finalize<uint64_t><<<njobs, 256, 0, st>>>(jobs, njobs, res, h->sc, (uint64_t *)d_sig, d_nb,
                                                                h->d_retry.as<uint32_t>());
        else
            k3_super2_finalize<uint32_t><<<njobs, 256, 0, st>>>(jobs, njobs, res, h->sc, (uint32_t *)d_sig, d_nb,
                                                                h->d_retry.as<uint32_t>());
    } else {
        // RevOptDens: per-file scratch for the round winners sits behind the bins of the whole pass
        uint32_t *rev_win = h->p.algo == GSB_ALGO_REVOPTDENS ? h->d_bins.as<uint32_t>() + h->dens_pass_files * (size_t)h->sc.m +
                                                                   (size_t)joff * h->sc.m
                                                             : nullptr;
        k3_optdens_finalize<<<njobs, 256, 0, st>>>(jobs, njobs, res, h->sc, (float *)d_sig, d_nb,
                                                   h->d_retry.as<uint32_t>(), h->p.algo == GSB_ALGO_SUPER ? 1 : 0, rev_win);
    }
    h->launches += nchunks ? 2 : 1;
}

// bucket-array capacity (keys) of the partition path for a file of `len` bytes
static size_t part_cap_g(size_t len) { return ((len / kNB) * 5 / 4 + 256 + 3) & ~(size_t)3; }

int run_prob(gsb_sketcher *h, const std::vector<uint32_t> &todo, const std::vector<double> &tmult,
             const uint64_t *h_offsets, void *d_sig, uint64_t *d_nb, cudaStream_t st, K1Plan &kp, bool newpath) {
    const bool dna = h->p.data_t == GSB_DATA_DNA;
    const bool want_bounds = dna && !h->p.block_flag;
    const uint32_t n = (uint32_t)todo.size();
    const int kSlots = prob_slots();
    // slot capacities for this pass
    size_t max_len = 0;
    for (uint32_t f : todo) max_len = std::max<size_t>(max_len, h_offsets[f + 1] - h_offsets[f]);
    if (max_len > kMaxProbSym) {
        set_error("file of %zu bytes exceeds the %u-symbol limit of the prob path's 30-bit position "
                  "field", max_len, kMaxProbSym);
        return GSB_ERR_CAPACITY;
    }
    const size_t cap_max = std::max<size_t>(4096, (size_t)(2.0 * (double)max_len) + 64);
    const size_t filter_bytes_max = (((size_t)(prob_load() * (double)max_len) + 64) / 16 + 8) * 4;
    const size_t light_cap = (size_t)(3.0 * h->sc.m * h->sc.lnm8) + 65536;
    const size_t list_cap_max = std::min<size_t>(max_len + 64, light_cap + max_len / 2) + 64;
    const size_t key_bytes = part_consts(h).keybits <= 31 ? 4 : 8;
    for (int s = 0; s < 2 * kSlots && s < (int)n; s++) {
        ProbSlot &sl = h->slot[s];
        int rc;
        if (newpath) {
            const void *old_misc = sl.misc.p;
            if ((rc = sl.buckets.ensure((size_t)kNB * part_cap_g(max_len) * key_bytes + 64))) return rc;
            if ((rc = sl.cursor.ensure((size_t)kNB * 4))) return rc;
            if ((rc = sl.list.ensure(4096))) return rc;  // (only the filter path fills a candidate list)
            if ((rc = sl.misc.ensure(256))) return rc;
            if (sl.misc.p != old_misc) GSB_CUDA_TRY(cudaMemsetAsync(sl.misc.p, 0, 256, st));
            if ((rc = sl.slot2.ensure((size_t)h->sc.m * 16))) return rc;
            continue;
        }
        const void *old_cnt = sl.cnt.p, *old_list = sl.list.p, *old_misc = sl.misc.p;
        if ((rc = sl.bitmap.ensure(filter_bytes_max + 64))) return rc;
        if ((rc = sl.table.ensure(cap_max * 4 + 64))) return rc;
        if ((rc = sl.cnt.ensure(cap_max * 4 + 64))) return rc;
        if ((rc = sl.list.ensure(list_cap_max * sizeof(ListEntry)))) return rc;
        if ((rc = sl.ovf.ensure((max_len + 64) * sizeof(OvfEntry)))) return rc;
        if ((rc = sl.misc.ensure(256))) return rc;
        if (sl.cnt.p != old_cnt || sl.list.p != old_list || sl.misc.p != old_misc) {
            // a (re)allocated slot starts clean: all counters zero, no pending list
            GSB_CUDA_TRY(cudaMemsetAsync(sl.cnt.p, 0, sl.cnt.cap, st));
            GSB_CUDA_TRY(cudaMemsetAsync(sl.misc.p, 0, 256, st));
        }
        if ((rc = sl.hmin.ensure((size_t)h->sc.m * 8))) return rc;
        if ((rc = sl.sigw.ensure((size_t)h->sc.m * 8))) return rc;
    }
    const uint32_t ngroups = (n + kSlots - 1) / kSlots;
    int rc;
    if ((rc = h->h_jobs.ensure((size_t)n * sizeof(ProbJob)))) return rc;
    if ((rc = h->d_jobs.ensure((size_t)n * sizeof(ProbJob)))) return rc;
    if ((rc = h->h_chunk_prefix.ensure((size_t)ngroups * (kSlots + 1) * 4))) return rc;
    if ((rc = h->d_chunk_prefix.ensure((size_t)ngroups * (kSlots + 1) * 4))) return rc;
    if ((rc = h->d_bound.ensure((size_t)n * sizeof(ProbBound)))) return rc;
    if ((rc = h->d_overflow.ensure((size_t)n * 4))) return rc;
    if ((rc = h->h_overflow.ensure((size_t)n * 4))) return rc;
    ProbJob *hj = h->h_jobs.as<ProbJob>();
    uint32_t *hcp = h->h_chunk_prefix.as<uint32_t>();
    std::vector<uint32_t> group_chunks(ngroups);
    for (uint32_t g = 0; g < ngroups; g++) {
        uint32_t acc = 0;
        for (int s = 0; s <= kSlots; s++) {
            hcp[g * (kSlots + 1) + s] = acc;
            const uint32_t i = g * kSlots + s;
            if (s < kSlots && i < n) {
                const uint32_t f = todo[i];
                const size_t len = h_offsets[f + 1] - h_offsets[f];
                ProbSlot &sl = h->slot[(g & 1) * kSlots + s];
                ProbJob &j = hj[i];
                j.file = f;
                // position field just wide enough for this file (at least 24 bits, so that ordinary
                // genomes keep an 8-bit fingerprint); the rest of the 32-bit entry is fingerprint
                j.posbits = 24;
                while (j.posbits < 30 && ((size_t)1 << j.posbits) < len + 2) j.posbits++;
                j.nslot1 = (uint32_t)(((size_t)(prob_load() * (double)len) + 64) & ~(size_t)15);
                j.bitmap = sl.bitmap.as<uint32_t>();
                j.n_coll = sl.misc.as<uint32_t>() + 3;
                j.cap2_max = (uint32_t)std::max<size_t>(4096, (size_t)(2.0 * (double)len) + 64);
                j.cap2 = sl.misc.as<uint32_t>() + 4;
                j.table = sl.table.as<uint32_t>();
                j.cnt = sl.cnt.as<uint32_t>();
                j.list = sl.list.as<ListEntry>();
                j.list_cap = (uint32_t)(std::min<size_t>(len + 64, light_cap + len / 2) + 64);
                j.list_n = sl.misc.as<uint32_t>();
                j.prev_n = sl.misc.as<uint32_t>() + 1;
                j.ovf = sl.ovf.as<OvfEntry>();
                j.ovf_cap = (uint32_t)(len + 64);
                j.ovf_n = sl.misc.as<uint32_t>() + 2;
                j.hmin = sl.hmin.as<unsigned long long>();
                j.sigw = sl.sigw.as<unsigned long long>();
                j.tmult = tmult[i];
                j.buckets = sl.buckets.p;
                j.cursor = sl.cursor.as<uint32_t>();
                j.cap_g = (uint32_t)part_cap_g(len);
                j.newpath = newpath ? 1u : 0u;
                j.slot2 = sl.slot2.as<ulonglong2>();
                acc += (uint32_t)((len + kChunk - 1) / kChunk);
            }
        }
        group_chunks[g] = acc;
    }
    GSB_CUDA_TRY(pull_small(h->d_jobs.p, hj, (size_t)n * sizeof(ProbJob), st));
    GSB_CUDA_TRY(pull_small(h->d_chunk_prefix.p, hcp, (size_t)ngroups * (kSlots + 1) * 4, st));
    // groups alternate between two streams (and two slot sets); everything before (K1, job
    // descriptors) happened on `st`, everything after waits for both
    GSB_CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    for (auto gs : h->gstream) GSB_CUDA_TRY(cudaStreamWaitEvent(gs, h->ev_fork, 0));
    {
        Timed tp_(h, CAT_K2, st);
        for (uint32_t g = 0; g < ngroups; g++) {
            const uint32_t joff = g * kSlots, nj = std::min<uint32_t>(kSlots, n - joff);
            const uint32_t cpoff = g * (kSlots + 1);
            cudaStream_t gs = h->gstream[g & 1];
            if (kp.pending) {
                // first pass: todo is the whole batch in order, so group g = files [joff, joff + nj).
                // K1 runs ahead on `st`, one group at a time, under the scan kernels of earlier groups.
                GSB_CUDA_TRY(wait_files_ready(h, st, joff + nj - 1));
                launch_k1_range(h, kp, joff, nj, st);
                if (h->ev_k1.size() <= g) {
                    cudaEvent_t e = nullptr;
                    GSB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    h->ev_k1.push_back(e);
                }
                GSB_CUDA_TRY(cudaEventRecord(h->ev_k1[g], st));
                GSB_CUDA_TRY(cudaStreamWaitEvent(gs, h->ev_k1[g], 0));
            }
#define GSB_PROB_LAUNCH(SRC, KT, DNA, WB)                                                                        \
    do {                                                                                                         \
        if (newpath && DNA)                                                                                      \
            launch_part_group_dna<KT>(h, joff, nj, group_chunks[g], cpoff, WB, d_sig, d_nb, gs);                 \
        else if (newpath)                                                                                        \
            launch_part_group_key<SRC<KT>, KT>(h, joff, nj, group_chunks[g], cpoff, DNA, WB, d_sig, d_nb, gs);   \
        else                                                                                                     \
            launch_prob_group<SRC<KT>, KT>(h, joff, nj, group_chunks[g], cpoff, DNA, WB, d_sig, d_nb, gs);       \
    } while (0)
            if (dna) {
                if (h->kt32) GSB_PROB_LAUNCH(SrcDNA, uint32_t, true, want_bounds);
                else GSB_PROB_LAUNCH(SrcDNA, uint64_t, true, want_bounds);
            } else {
                if (h->kt32) GSB_PROB_LAUNCH(SrcAA, uint32_t, false, false);
                else GSB_PROB_LAUNCH(SrcAA, uint64_t, false, false);
            }
#undef GSB_PROB_LAUNCH
            if (kp.pending && h->early_out) {
                if (h->ev_grp.size() <= g) {
                    cudaEvent_t e = nullptr;
                    GSB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    h->ev_grp.push_back(e);
                }
                const size_t rowb = (size_t)h->sc.m * h->elem;
                GSB_CUDA_TRY(cudaEventRecord(h->ev_grp[g], gs));
                GSB_CUDA_TRY(cudaStreamWaitEvent(h->d2h_stream, h->ev_grp[g], 0));
                GSB_CUDA_TRY(cudaMemcpyAsync(h->early_out + (size_t)joff * rowb, h->early_src + (size_t)joff * rowb,
                                             (size_t)nj * rowb, cudaMemcpyDeviceToHost, h->d2h_stream));
            }
        }
        kp.pending = false;
        for (int i = 0; i < 2; i++) {
            GSB_CUDA_TRY(cudaEventRecord(h->ev_join[i], h->gstream[i]));
            GSB_CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join[i], 0));
        }
    }
    GSB_CUDA_TRY(cudaGetLastError());
    GSB_CUDA_TRY(cudaMemcpyAsync(h->h_overflow.p, h->d_overflow.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    return GSB_OK;
}

// SuperMinHash / SuperMinHash2 cold path: files whose k-mers do not reach every slot at the first level
template <class Src, typename KT>
void launch_super_seq(gsb_sketcher *h, uint32_t nlist, bool dna, bool want_bounds, void *d_sig, cudaStream_t st) {
    const uint32_t *packed_dna = dna ? h->d_packed.as<uint32_t>() : nullptr;
    const uint8_t *packed_aa = dna ? nullptr : h->d_packed.as<uint8_t>();
    const uint32_t *bounds = want_bounds ? h->d_bounds.as<uint32_t>() : nullptr;
    if (h->p.algo == GSB_ALGO_HLL) {
        k_hll_sequential<Src, KT><<<(nlist + 31) / 32, 32, 0, st>>>(
            h->d_jobs.as<uint32_t>(), nlist, h->d_files.as<FileDesc>(), h->d_res.as<FileResult>(), packed_dna, packed_aa,
            bounds, h->sc, (uint16_t *)d_sig, h->d_bins.as<uint32_t>());
    } else if (h->p.algo == GSB_ALGO_SUPER2) {
        if (h->elem == 8)
            k_super2_sequential<Src, KT, uint64_t><<<(nlist + 31) / 32, 32, 0, st>>>(
                h->d_jobs.as<uint32_t>(), nlist, h->d_files.as<FileDesc>(), h->d_res.as<FileResult>(), packed_dna,
                packed_aa, bounds, h->sc, (uint64_t *)d_sig, h->d_bins.as<uint8_t>());
        else
            k_super2_sequential<Src, KT, uint32_t><<<(nlist + 31) / 32, 32, 0, st>>>(
                h->d_jobs.as<uint32_t>(), nlist, h->d_files.as<FileDesc>(), h->d_res.as<FileResult>(), packed_dna,
                packed_aa, bounds, h->sc, (uint32_t *)d_sig, h->d_bins.as<uint8_t>());
    } else {
        k_super_sequential<Src, KT><<<(nlist + 31) / 32, 32, 0, st>>>(
            h->d_jobs.as<uint32_t>(), nlist, h->d_files.as<FileDesc>(), h->d_res.as<FileResult>(), packed_dna, packed_aa,
            bounds, h->sc, (float *)d_sig, h->d_bins.as<uint32_t>());
    }
    h->launches += 1;
}

int run_super_sequential(gsb_sketcher *h, const std::vector<uint32_t> &list, void *d_sig, cudaStream_t st) {
    const bool dna = h->p.data_t == GSB_DATA_DNA;
    const bool want_bounds = dna && !h->p.block_flag;
    const uint32_t nl = (uint32_t)list.size();
    int rc;
    if ((rc = h->h_jobs.ensure((size_t)nl * 4))) return rc;
    if ((rc = h->d_jobs.ensure((size_t)nl * 4))) return rc;
    if ((rc = h->d_bins.ensure((size_t)nl * h->sc.m * (h->p.algo == GSB_ALGO_SUPER2 ? 28 : 12) + 64))) return rc;
    memcpy(h->h_jobs.p, list.data(), (size_t)nl * 4);
    GSB_CUDA_TRY(pull_small(h->d_jobs.p, h->h_jobs.p, (size_t)nl * 4, st));
    if (dna) {
        if (h->kt32) launch_super_seq<SrcDNA<uint32_t>, uint32_t>(h, nl, true, want_bounds, d_sig, st);
        else launch_super_seq<SrcDNA<uint64_t>, uint64_t>(h, nl, true, want_bounds, d_sig, st);
    } else {
        if (h->kt32) launch_super_seq<SrcAA<uint32_t>, uint32_t>(h, nl, false, false, d_sig, st);
        else launch_super_seq<SrcAA<uint64_t>, uint64_t>(h, nl, false, false, d_sig, st);
    }
    GSB_CUDA_TRY(cudaGetLastError());
    GSB_CUDA_TRY(cudaStreamSynchronize(st));
    return GSB_OK;
}

int run_dens(gsb_sketcher *h, const std::vector<uint32_t> &todo, const std::vector<double> &tmult,
             const uint64_t *h_offsets, void *d_sig, uint64_t *d_nb, cudaStream_t st, K1Plan &kp) {
    const bool dna = h->p.data_t == GSB_DATA_DNA;
    const bool want_bounds = dna && !h->p.block_flag;
    const uint32_t n = (uint32_t)todo.size();
    // files per group: groups alternate between the two group streams (GSB_DENS_GROUP overrides)
    const uint32_t grp = (uint32_t)env_int("GSB_DENS_GROUP", 16, 1, 256);
    const uint32_t ngroups = (n + grp - 1) / grp;
    int rc;
    if ((rc = h->h_jobs.ensure((size_t)n * sizeof(DensJob)))) return rc;
    if ((rc = h->d_jobs.ensure((size_t)n * sizeof(DensJob)))) return rc;
    if ((rc = h->h_chunk_prefix.ensure((size_t)ngroups * (grp + 1) * 4))) return rc;
    if ((rc = h->d_chunk_prefix.ensure((size_t)ngroups * (grp + 1) * 4))) return rc;
    const bool super2 = h->p.algo == GSB_ALGO_SUPER2;
    // per file: m f32 bins (+ m round winners for RevOptDens), or m 128-bit (r, hash) slots for SuperMinHash2
    const size_t per_file = super2 ? (size_t)h->sc.m * 16 : (size_t)h->sc.m * 4 * (h->p.algo == GSB_ALGO_REVOPTDENS ? 2 : 1);
    if ((rc = h->d_bins.ensure((size_t)n * per_file + 64))) return rc;
    h->dens_pass_files = n;
    DensJob *hj = h->h_jobs.as<DensJob>();
    uint32_t *hcp = h->h_chunk_prefix.as<uint32_t>();
    std::vector<uint32_t> group_chunks(ngroups);
    for (uint32_t g = 0; g < ngroups; g++) {
        uint64_t acc = 0;
        for (uint32_t s = 0; s <= grp; s++) {
            hcp[g * (grp + 1) + s] = (uint32_t)acc;
            const uint32_t i = g * grp + s;
            if (s < grp && i < n) {
                const uint32_t f = todo[i];
                const size_t len = h_offsets[f + 1] - h_offsets[f];
                hj[i].file = f;
                hj[i].bins = super2 ? reinterpret_cast<uint32_t *>(h->d_bins.as<ulonglong2>() + (size_t)i * h->sc.m)
                                    : h->d_bins.as<uint32_t>() + (size_t)i * h->sc.m;
                hj[i].tmult = tmult[i];
                acc += (len + kChunk - 1) / kChunk;
            }
        }
        if (acc > 0x7FFFFFFFull) {
            set_error("batch too large: %llu k-mer chunks", (unsigned long long)acc);
            return GSB_ERR_CAPACITY;
        }
        group_chunks[g] = (uint32_t)acc;
    }
    GSB_CUDA_TRY(pull_small(h->d_jobs.p, hj, (size_t)n * sizeof(DensJob), st));
    GSB_CUDA_TRY(pull_small(h->d_chunk_prefix.p, hcp, (size_t)ngroups * (grp + 1) * 4, st));
    if (h->p.algo == GSB_ALGO_HLL) GSB_CUDA_TRY(cudaMemsetAsync(h->d_bins.p, 0, (size_t)n * h->sc.m * 4, st));  // registers start at 0
    else if (super2) k_super2_reset<<<592, 256, 0, st>>>(h->d_bins.as<ulonglong2>(), (size_t)n * h->sc.m);
    else k_dens_reset<<<592, 256, 0, st>>>(h->d_bins.as<uint32_t>(), (size_t)n * h->sc.m);
    h->launches += 1;
    GSB_CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    for (auto gs : h->gstream) GSB_CUDA_TRY(cudaStreamWaitEvent(gs, h->ev_fork, 0));
    {
        Timed tp_(h, CAT_K2, st);
        for (uint32_t g = 0; g < ngroups; g++) {
            const uint32_t joff = g * grp, nj = std::min<uint32_t>(grp, n - joff);
            const uint32_t cpoff = g * (grp + 1);
            cudaStream_t gs = h->gstream[g & 1];
            if (kp.pending) {  // first pass: group g = files [joff, joff + nj); K1 runs ahead on `st`
                GSB_CUDA_TRY(wait_files_ready(h, st, joff + nj - 1));
                launch_k1_range(h, kp, joff, nj, st);
                if (h->ev_k1.size() <= g) {
                    cudaEvent_t e = nullptr;
                    GSB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    h->ev_k1.push_back(e);
                }
                GSB_CUDA_TRY(cudaEventRecord(h->ev_k1[g], st));
                GSB_CUDA_TRY(cudaStreamWaitEvent(gs, h->ev_k1[g], 0));
            }
            if (dna) {
                if (h->kt32)
                    launch_dens<SrcDNA<uint32_t>, uint32_t>(h, joff, nj, group_chunks[g], cpoff, true, want_bounds,
                                                            d_sig, d_nb, gs);
                else
                    launch_dens<SrcDNA<uint64_t>, uint64_t>(h, joff, nj, group_chunks[g], cpoff, true, want_bounds,
                                                            d_sig, d_nb, gs);
            } else {
                if (h->kt32)
                    launch_dens<SrcAA<uint32_t>, uint32_t>(h, joff, nj, group_chunks[g], cpoff, false, false, d_sig,
                                                           d_nb, gs);
                else
                    launch_dens<SrcAA<uint64_t>, uint64_t>(h, joff, nj, group_chunks[g], cpoff, false, false, d_sig,
                                                           d_nb, gs);
            }
        }
        kp.pending = false;
        for (int i = 0; i < 2; i++) {
            GSB_CUDA_TRY(cudaEventRecord(h->ev_join[i], h->gstream[i]));
            GSB_CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join[i], 0));
        }
    }
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB_OK;
}

}  // namespace

static int sketch_batch_dev_locked(gsb_sketcher *h, const uint8_t *d_bytes, const uint64_t *h_offsets, uint32_t n,
                                   void *d_sig_out, uint64_t *d_nb_bases_out, void *stream);

extern "C" int gsb_sketch_fasta_batch_dev(gsb_sketcher *h, const uint8_t *d_bytes, const uint64_t *h_offsets,
                                          uint32_t n, void *d_sig_out, uint64_t *d_nb_bases_out,
                                          void *stream) {
    if (!h) {
        set_error("gsb_sketch_fasta_batch_dev: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    std::lock_guard<std::mutex> lock(h->mu);
    return sketch_batch_dev_locked(h, d_bytes, h_offsets, n, d_sig_out, d_nb_bases_out, stream);
}

static int sketch_batch_dev_locked(gsb_sketcher *h, const uint8_t *d_bytes, const uint64_t *h_offsets, uint32_t n,
                                   void *d_sig_out, uint64_t *d_nb_bases_out, void *stream) {
    if (!h || !h_offsets || (n && (!d_bytes && h_offsets[n] > 0)) || (n && !d_sig_out)) {
        set_error("gsb_sketch_fasta_batch_dev: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (n == 0) return GSB_OK;
    for (uint32_t i = 0; i < n; i++)
        if (h_offsets[i + 1] < h_offsets[i]) {
            set_error("offsets must be non-decreasing (file %u)", i);
            return GSB_ERR_INVALID_ARG;
        }
    GSB_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool dna = h->p.data_t == GSB_DATA_DNA;
    const bool seq = !h->p.block_flag;
    const bool want_bounds = dna && seq;
    const uint64_t base0 = h_offsets[0];
    const uint64_t total = h_offsets[n];

    // ---- K1 descriptors
    int rc;
    if ((rc = h->h_files.ensure((size_t)n * sizeof(FileDesc)))) return rc;
    if ((rc = h->h_tile_prefix.ensure((size_t)(n + 1) * 4))) return rc;
    FileDesc *hf = h->h_files.as<FileDesc>();
    uint32_t *htp = h->h_tile_prefix.as<uint32_t>();
    uint64_t ntiles = 0, out_off = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t b = h_offsets[i], e = h_offsets[i + 1];
        hf[i].beg = b;
        hf[i].end = e;
        hf[i].out_off = out_off;
        hf[i].tile_first = (uint32_t)ntiles;
        hf[i].ntiles = e > b ? (uint32_t)((e + kTile - 1) / kTile - b / kTile) : 0;
        htp[i] = (uint32_t)ntiles;
        ntiles += hf[i].ntiles;
        const uint64_t len = e - b;
        // DNA: words (16 bases each) rounded to an even count so 8-byte loads stay aligned;
        // AA: bytes rounded to 16
        out_off += dna ? ((len / 16 + 4) & ~1ull) : ((len + 31) & ~15ull);
    }
    htp[n] = (uint32_t)ntiles;
    if (ntiles > 0x7FFFFFFFull) {
        set_error("batch too large: %llu tiles", (unsigned long long)ntiles);
        return GSB_ERR_CAPACITY;
    }
    (void)base0;
    const size_t packed_bytes = (dna ? out_off * 4 : out_off) + (size_t)kChunk + 4096;
    const uint32_t bd_cap = want_bounds ? (uint32_t)std::min<uint64_t>((total - base0) / 32 + n + 1024, 0x7FFFFFFFull) : 0;
    if ((rc = h->d_files.ensure((size_t)n * sizeof(FileDesc)))) return rc;
    if ((rc = h->d_tile_prefix.ensure((size_t)(n + 1) * 4))) return rc;
    if ((rc = h->d_tc4.ensure((size_t)ntiles * 8 + 8))) return rc;
    if ((rc = h->d_ttrans.ensure((size_t)ntiles + 8))) return rc;
    if ((rc = h->d_tnrec.ensure((size_t)ntiles * 2 + 8))) return rc;
    if ((rc = h->d_tstate.ensure((size_t)ntiles + 8))) return rc;
    if ((rc = h->d_tbase.ensure((size_t)ntiles * 4 + 8))) return rc;
    if ((rc = h->d_trecbase.ensure((size_t)ntiles * 4 + 8))) return rc;
    if ((rc = h->d_res.ensure((size_t)n * sizeof(FileResult)))) return rc;
    if ((rc = h->d_packed.ensure(packed_bytes))) return rc;
    if ((rc = h->d_bounds.ensure((size_t)bd_cap * 4 + 8))) return rc;
    if ((rc = h->d_misc.ensure(256))) return rc;
    if ((rc = h->d_retry.ensure((size_t)n * 4))) return rc;
    if ((rc = h->h_retry.ensure((size_t)n * 4))) return rc;
    if ((rc = h->d_fqflag.ensure((size_t)n * 4))) return rc;

    GSB_CUDA_TRY(pull_small(h->d_files.p, hf, (size_t)n * sizeof(FileDesc), st));
    GSB_CUDA_TRY(pull_small(h->d_tile_prefix.p, htp, (size_t)(n + 1) * 4, st));
    GSB_CUDA_TRY(cudaMemsetAsync(h->d_misc.p, 0, 256, st));
    GSB_CUDA_TRY(cudaMemsetAsync(h->d_retry.p, 0, (size_t)n * 4, st));
    GSB_CUDA_TRY(cudaMemsetAsync(h->d_fqflag.p, 0, (size_t)n * 4, st));
    if (dna) GSB_CUDA_TRY(cudaMemsetAsync(h->d_packed.p, 0, packed_bytes, st));
    K1Plan kp;
    kp.d_bytes = d_bytes;
    kp.total = total;
    kp.bd_cap = bd_cap;
    kp.pending = true;
    const bool prob = h->p.algo == GSB_ALGO_PROB3A;
    // files of a group without any tile are not visited by K1: their result is "empty"
    GSB_CUDA_TRY(cudaMemsetAsync(h->d_res.p, 0, (size_t)n * sizeof(FileResult), st));
    if (!ntiles) kp.pending = false;  // otherwise K1 runs group by group inside the first pass

    // ---- K2/K3 with bound retries
    std::vector<uint32_t> todo(n);
    std::vector<double> tmult(n, 1.0);
    for (uint32_t i = 0; i < n; i++) todo[i] = i;
    std::vector<uint32_t> seq_files;  // SuperMinHash: files for the sequential cold path
    // prob: the partition path first; files it flags (a bucket array or a counting round overflowed:
    // very large or very repetitive inputs) are re-run through the general filter path
    bool newpath = prob && h->prob_path == 0;
    if (newpath) {
        size_t max_len = 0;
        for (uint32_t i = 0; i < n; i++) max_len = std::max<size_t>(max_len, h_offsets[i + 1] - h_offsets[i]);
        if (part_cap_g(max_len) > kCapGMax) newpath = false;  // 16-bit duplicate counters
    }
    for (int attempt = 0; attempt < 10 && !todo.empty(); attempt++) {
        rc = prob ? run_prob(h, todo, tmult, h_offsets, d_sig_out, d_nb_bases_out, st, kp, newpath)
                  : run_dens(h, todo, tmult, h_offsets, d_sig_out, d_nb_bases_out, st, kp);
        if (rc) return rc;
        GSB_CUDA_TRY(cudaMemcpyAsync(h->h_retry.p, h->d_retry.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        GSB_CUDA_TRY(cudaStreamSynchronize(st));
        collect_spans(h);
        const uint32_t *hr = h->h_retry.as<uint32_t>();
        const uint32_t *ho = prob ? h->h_overflow.as<uint32_t>() : nullptr;
        std::vector<uint32_t> next, fallback;
        std::vector<double> next_t, fallback_t;
        for (size_t i = 0; i < todo.size(); i++) {
            const uint32_t f = todo[i];
            const uint32_t status = hr[f] >> 8;
            if (status == 5) {
                set_error("file %u starts with neither '>' nor '@' (not FASTA / FASTQ)", h->file_base + f);
                return GSB_ERR_BAD_INPUT;
            }
            if (status == 6) {
                set_error("file %u: malformed FASTQ record (a line that should start with '@' or '+' does not, or "
                          "the last record is truncated)", h->file_base + f);
                return GSB_ERR_BAD_INPUT;
            }
            if (status == 9) {
                set_error("file %u: FASTQ text contains \"capsid\": the reference drops records by id "
                          "(src/dna/dnafiles.rs:67), which the FASTQ device path does not implement", h->file_base + f);
                return GSB_ERR_UNSUPPORTED;
            }
            if (status == 8) {
                set_error("file %u: record-boundary pool exhausted", h->file_base + f);
                return GSB_ERR_CAPACITY;
            }
            if (ho && newpath && (ho[i] & 3u)) {  // does not fit the partition geometry: general path
                fallback.push_back(f);
                fallback_t.push_back(tmult[i]);
                continue;
            }
            if (ho && ho[i]) {
                // candidate list overflow: counters may be dirty; clear and fail loudly
                for (auto &s : h->slot)
                    if (s.cnt.p) {
                        cudaMemsetAsync(s.cnt.p, 0, s.cnt.cap, st);
                        cudaMemsetAsync(s.misc.p, 0, 256, st);
                    }
                set_error("file %u: candidate list overflow (pathological repeat structure)", h->file_base + f);
                return GSB_ERR_CAPACITY;
            }
            if (hr[f] & 1u) {
                next.push_back(f);
                next_t.push_back((prob || h->p.algo == GSB_ALGO_HLL) ? tmult[i] * 8.0 : 1e30);
            }
            if (hr[f] & 2u) seq_files.push_back(f);
        }
        h->retries += next.size() + fallback.size();
        h->fallbacks += fallback.size();
        if (!next.empty()) {  // bound retries of this path first; flagged files wait for their turn
            todo.swap(next);
            tmult.swap(next_t);
            if (!fallback.empty()) {
                h->pending_fallback.insert(h->pending_fallback.end(), fallback.begin(), fallback.end());
                h->pending_fallback_t.insert(h->pending_fallback_t.end(), fallback_t.begin(), fallback_t.end());
            }
        } else {
            fallback.insert(fallback.end(), h->pending_fallback.begin(), h->pending_fallback.end());
            fallback_t.insert(fallback_t.end(), h->pending_fallback_t.begin(), h->pending_fallback_t.end());
            h->pending_fallback.clear();
            h->pending_fallback_t.clear();
            todo.swap(fallback);
            tmult.swap(fallback_t);
            if (!todo.empty()) newpath = false;
        }
    }
    h->pending_fallback.clear();
    h->pending_fallback_t.clear();
    if (!todo.empty()) {
        set_error("early-stop bound did not converge for %zu file(s)", todo.size());
        return GSB_ERR_CUDA;
    }
    if (!seq_files.empty()) {
        h->retries += seq_files.size();
        if ((rc = run_super_sequential(h, seq_files, d_sig_out, st))) return rc;
    }
    return GSB_OK;
}

// host FASTA in; signatures (and encoded lengths) out to host memory or, with out_dev, straight to the
// caller's device buffers (e.g. this rank's slice of the replicated signature matrix)
static int sketch_host_in(gsb_sketcher *h, const uint8_t *bytes, const uint64_t *offsets, uint32_t n, void *sig_out,
                          uint64_t *nb_bases_out, bool out_dev) {
    if (!h || !offsets || (n && !sig_out)) {
        set_error("gsb_sketch_fasta_batch: NULL argument");
        return GSB_ERR_INVALID_ARG;
    }
    if (n == 0) return GSB_OK;
    std::lock_guard<std::mutex> lock(h->mu);
    GSB_CUDA_TRY(cudaSetDevice(h->device));
    const uint64_t lo = offsets[0], hi = offsets[n];
    if (hi < lo || (hi > lo && !bytes)) {
        set_error("gsb_sketch_fasta_batch: bad offsets / NULL bytes");
        return GSB_ERR_INVALID_ARG;
    }
    const size_t sig_row = (size_t)h->sc.m * h->elem;
    int rc;
    if ((rc = h->d_bytes.ensure(hi - lo + 64))) return rc;
    if (!out_dev && (rc = h->d_sig.ensure((size_t)n * sig_row))) return rc;
    if ((!out_dev || !nb_bases_out) && (rc = h->d_nb.ensure((size_t)n * 8))) return rc;
    cudaStream_t st = h->stream;
    if (!h->copy_stream) {
        GSB_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        GSB_CUDA_TRY(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    }
    // The bytes go to the device in chunks of whole files on a copy stream, one event per chunk;
    // the kernels of a genome group only wait for the chunk that holds the group's last file, so
    // the H2D transfer of later files runs under the kernels of earlier ones.  With pageable host
    // memory the copies block the host instead: same result, no overlap.
    h->h2d_end.clear();
    {
        const uint64_t kChunkBytes = 48ull << 20;
        // one chunk per genome group on the prob path: a group starts as soon as its own files are in
        const uint32_t kChunkFiles = h->p.algo == GSB_ALGO_PROB3A ? (uint32_t)prob_slots() : 8u;
        uint32_t b = 0;
        for (uint32_t i = 1; i <= n; i++)
            if (i == n || i - b >= kChunkFiles || offsets[i + 1] - offsets[b] > kChunkBytes) {
                h->h2d_end.push_back(i);
                b = i;
            }
    }
    while (h->ev_h2d.size() < h->h2d_end.size()) {
        cudaEvent_t e = nullptr;
        GSB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_h2d.push_back(e);
    }
    static const bool trace = getenv("GSB_E2E_TRACE") != nullptr;
    static cudaEvent_t tr_c0 = nullptr, tr_c1 = nullptr;
    if (trace && !tr_c0) {
        cudaEventCreate(&tr_c0);
        cudaEventCreate(&tr_c1);
    }
    if (trace) cudaEventRecord(tr_c0, h->copy_stream);
    {
        uint32_t f0 = 0;
        for (size_t c = 0; c < h->h2d_end.size(); c++) {
            const uint64_t b = offsets[f0], e = offsets[h->h2d_end[c]];
            if (e > b)
                GSB_CUDA_TRY(cudaMemcpyAsync(h->d_bytes.as<uint8_t>() + (b - lo), bytes + b, e - b,
                                             cudaMemcpyHostToDevice, h->copy_stream));
            GSB_CUDA_TRY(cudaEventRecord(h->ev_h2d[c], h->copy_stream));
            f0 = h->h2d_end[c];
        }
    }
    std::vector<uint64_t> rel(n + 1);
    for (uint32_t i = 0; i <= n; i++) rel[i] = offsets[i] - lo;
    if (trace) cudaEventRecord(tr_c1, h->copy_stream);
    const auto tr0 = std::chrono::steady_clock::now();
    h->early_out = nullptr;
    if (!out_dev && h->p.algo == GSB_ALGO_PROB3A && !getenv("GSB_NO_EARLY_D2H")) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, sig_out) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
            h->early_out = (uint8_t *)sig_out;  // pinned: an asynchronous copy really is one
            h->early_src = (const uint8_t *)h->d_sig.p;
        } else {
            (void)cudaGetLastError();
        }
    }
    const uint64_t retries_before = h->retries;
    h->h2d_active = true;
    rc = sketch_batch_dev_locked(h, h->d_bytes.as<uint8_t>(), rel.data(), n, out_dev ? sig_out : h->d_sig.p,
                                 (out_dev && nb_bases_out) ? nb_bases_out : h->d_nb.as<uint64_t>(), (void *)st);
    h->h2d_active = false;
    const auto tr1 = std::chrono::steady_clock::now();
    if (rc) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->d2h_stream);  // early copies into the caller's buffer must not outlive the call
        h->early_out = nullptr;
        return rc;
    }
    if (trace) {
        cudaStreamSynchronize(h->copy_stream);
        const auto tr2 = std::chrono::steady_clock::now();
        float h2d_ms = 0;
        cudaEventElapsedTime(&h2d_ms, tr_c0, tr_c1);
        fprintf(stderr, "e2e trace: n=%u compute-done %.3f ms, copy stream idle at %.3f ms, H2D alone took %.3f ms\n", n,
                std::chrono::duration<double, std::milli>(tr1 - tr0).count(),
                std::chrono::duration<double, std::milli>(tr2 - tr0).count(), h2d_ms);
    }
    if (out_dev) return GSB_OK;  // batch_dev returned after synchronising `st`
    // batch_dev returned after synchronising `st`: the results are complete.  Rows that went back
    // early are final unless a file was re-run (bound retry / general path): then everything is copied again
    const bool early_done = h->early_out != nullptr && h->retries == retries_before;
    h->early_out = nullptr;
    if (!early_done)
        GSB_CUDA_TRY(cudaMemcpyAsync(sig_out, h->d_sig.p, (size_t)n * sig_row, cudaMemcpyDeviceToHost, h->d2h_stream));
    if (nb_bases_out)
        GSB_CUDA_TRY(cudaMemcpyAsync(nb_bases_out, h->d_nb.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->d2h_stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
    return GSB_OK;
}

extern "C" int gsb_sketch_fasta_batch(gsb_sketcher *h, const uint8_t *bytes, const uint64_t *offsets,
                                      uint32_t n, void *sig_out, uint64_t *nb_bases_out) {
    return sketch_host_in(h, bytes, offsets, n, sig_out, nb_bases_out, false);
}

extern "C" int gsb_sketch_fasta_batch_to_dev(gsb_sketcher *h, const uint8_t *bytes, const uint64_t *offsets,
                                             uint32_t n, void *d_sig_out, uint64_t *d_nb_bases_out) {
    return sketch_host_in(h, bytes, offsets, n, d_sig_out, d_nb_bases_out, true);
}
