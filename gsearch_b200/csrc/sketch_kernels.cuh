// sketch_kernels.cuh -- K2/K3: encoded sequence -> canonical k-mers -> signature (sm_100a).
//
// Replaces, on device:
//   kmerutils KmerSeqIterator::next [U] and the two hash closures
//     src/dna/dnasketch.rs:164-169 (canonical min(kmer, revcomp) & 2k-bit mask)
//     src/aa/aasketch.rs:156-160  (5k-bit mask, no reverse complement)
//   kmerutils ProbHash3aSketch::sketch_compressedkmer[_seqs] [U] = exact multiplicity map +
//     probminhash ProbMinHash3a::hashset [U]                 (called at dnasketch.rs:336,357)
//   kmerutils OptDensHashSketch + probminhash OptDensMinHash::{sketch,end_sketch} [U]
//                                                            (dispatch aasketch.rs:524-536)
//
// ProbMinHash3a on a GPU (DESIGN.md "ProbMinHash3a, restated for parallel hardware"):
//   sig[k] = argmin over all points (d, i) with slot k_i(d) = k of h_i(d) = (i-1+x_i(d))/w_d,
//   where (x_1,k_1,x_2,k_2,...) is the draw sequence of xoshiro256++ seeded by d.  The result
//   does not depend on processing order, and a point with h >= T can be skipped whenever
//   T >= max_k min-h[k] at the end (verified per genome; widened and re-run otherwise).
//   * K2 streams k-mers twice.  `k2_prob_mark`: every occurrence marks a blocked filter with one
//     atomicOr; a k-mer that occurs more than once always ends with its flag raised.
//     `k2_prob_classify`: a k-mer whose flag is down has weight 1 exactly and is kept only if
//     its first draw can be below T ("light"); flagged occurrences go through an EXACT set
//     (`k2_prob_overflow`: 32-bit entries = fingerprint | position of the first occurrence,
//     equality verified against the packed sequence), which counts them.
//   * K3a/K3b replay the exact f64 arithmetic for those few k-mers: atomicMin on the ordered
//     bit pattern of h, then the owner of the minimum writes the k-mer (ties: smaller k-mer).
#pragma once

#include "common.cuh"
#include "fasta_pack.cuh"

namespace gsb {

constexpr int kK2Threads = 256;
constexpr int kRun = 32;                        // consecutive k-mer positions per thread
constexpr int kChunk = kK2Threads * kRun;       // positions per CTA
constexpr uint32_t kMaxProbSym = (1u << 30) - 2;  // widest position field of an exact-set entry (30 bits)
constexpr int kG = 8;                             // k-mers hashed and probed together per thread
constexpr uint32_t kStageCap = 3072;              // per-CTA candidate stage (entries, 48 KiB)
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;         // list entry of a k-mer that never entered the exact set

struct ListEntry {
    uint64_t kmer;
    uint32_t slot;
    uint32_t kind;  // 0 = light (first occurrence, first draw below T), 1 = repeated,
                    // 2 = from the partition path (prob_partition.cuh): `slot` is the weight itself
};

// per-genome device-side job state for the prob path
struct ProbJob {
    uint32_t file;       // index into FileDesc / FileResult
    uint32_t posbits;    // exact-set entry = fingerprint (32 - posbits bits) | position + 1 (posbits bits)
    uint32_t nslot1;     // filter size in 2-bit units (16 per 32-bit word)
    uint32_t *bitmap;    // blocked filter: per word 24 "seen" bits + 8 "repeated" flags
    uint32_t *n_coll;    // occurrences that found their seen bits already set (statistics)
    uint32_t cap2_max;   // allocated exact-set capacity
    uint32_t *cap2;      // exact-set capacity in use (sized on device from n_coll)
    uint32_t *table;     // exact set: fingerprint | position+1 of the first occurrence
    uint32_t *cnt;       // extra occurrences per slot
    ListEntry *list;     // candidates
    uint32_t list_cap;
    uint32_t *list_n;    // cursor (device)
    uint32_t *prev_n;    // entries of the slot's previous genome whose counters are still set
    struct OvfEntry *ovf;  // flagged occurrences on their way to the exact set
    uint32_t ovf_cap;
    uint32_t *ovf_n;
    unsigned long long *hmin;  // [m] ordered bits of min h per slot
    unsigned long long *sigw;  // [m] winning k-mer per slot
    double tmult;        // early-stop bound multiplier (1 = default, grown on retry)
    // partition path (prob_partition.cuh)
    void *buckets;       // [kNB][cap_g] keys
    uint32_t *cursor;    // [kNB] keys appended to each bucket
    uint32_t cap_g;      // capacity of one bucket array
    uint32_t newpath;    // 1: this job runs the partition path (no extra-occurrence counters to clear)
    ulonglong2 *slot2;   // [m] partition path: (ordered bits of min h, winning k-mer) as ONE 128-bit object
};

struct ProbBound {  // written by k_prob_reset
    double T;       // points with h >= T are skipped
    uint64_t uT;    // 52-bit first draws below this are "light" (conservative superset)
};

struct SketchConsts {
    uint32_t k, m;
    uint64_t zone;     // Uniform<usize>(0,m) rejection zone
    uint64_t u_slow;   // 52-bit draws >= this may take the ExpRestricted01 rejection path
    double lnm8;       // ln(m) + 8
    Exp01 e01;
    uint32_t spec_flags;
};

// ------------------------------------------------------------------ k-mer sources
__device__ __forceinline__ uint64_t revcomp64(uint64_t x, uint32_t nb) {
    x = ~x;
    x = __brevll(x);
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    return nb ? (x >> (64 - 2 * nb)) : 0ull;
}

struct SeqView {
    const uint32_t *dna;  // packed words (DNA)
    const uint8_t *aa;    // symbols (AA)
    const uint32_t *bounds;
    uint32_t nbounds;
    uint32_t N;           // symbols
};

// bases [pos, pos+nb) of a packed DNA sequence as an integer, first base most significant
__device__ __forceinline__ uint64_t dna_extract(const uint32_t *__restrict__ w, uint32_t pos,
                                                uint32_t nb) {
    const uint32_t wi = pos >> 4, sh = (pos & 15u) * 2;
    const uint64_t a = ((uint64_t)__ldg(&w[wi]) << 32) | __ldg(&w[wi + 1]);
    const uint64_t b = ((uint64_t)__ldg(&w[wi + 2]) << 32);
    const uint64_t hi = sh ? ((a << sh) | (b >> (64 - sh))) : a;  // 32 bases from pos
    return nb ? (hi >> (64 - 2 * nb)) : 0ull;
}

// KC > 0 fixes k at compile time (constant shifts and masks in the rolling update); KC = 0 reads it
// from the arguments
template <typename KT, int KC = 0>
struct SrcDNA {
    uint64_t nw;  // the next bases to enter the window, first one in the top two bits
    KT fw, rc, mask;
    uint32_t k, p0, nb, bi;
    const uint32_t *bounds;
    uint32_t nbounds, N;

    __device__ __forceinline__ void init(const SeqView &sv, uint32_t p0_, uint32_t k_) {
        k = KC ? (uint32_t)KC : k_;
        p0 = p0_;
        N = sv.N;
        bounds = sv.bounds;
        nbounds = sv.nbounds;
        const uint32_t wi = p0 >> 4;  // p0 is a multiple of 32 -> wi even
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(sv.dna + wi));
        const uint2 b = __ldg(reinterpret_cast<const uint2 *>(sv.dna + wi + 2));
        const uint64_t W0 = ((uint64_t)a.x << 32) | a.y;
        const uint64_t W1 = ((uint64_t)b.x << 32) | b.y;
        mask = (KT)((k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1));
        const uint64_t f0 = k > 1 ? (W0 >> (64 - 2 * (k - 1))) : 0ull;  // first k-1 bases
        fw = (KT)f0;
        rc = (KT)(revcomp64(f0, k - 1) << 2);
        // bases p0+k-1 .. p0+k+30 : the kRun = 32 bases that complete this thread's k-mers
        const uint32_t sh = 2 * (k - 1);
        nw = sh ? ((W0 << sh) | (W1 >> (64 - sh))) : W0;
        // first record boundary strictly after p0 (bounds is sorted ascending)
        uint32_t lo = 0, hi = nbounds;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&bounds[mid]) > p0) hi = mid; else lo = mid + 1;
        }
        bi = lo;
        nb = bi < nbounds ? __ldg(&bounds[bi]) : N;
    }
    __device__ __forceinline__ void roll(KT &canon) {
        const uint32_t b = (uint32_t)(nw >> 62);
        nw <<= 2;
        const uint32_t kk = KC ? (uint32_t)KC : k;
        const KT m = KC ? (KT)((KC >= 32) ? ~0ull : ((1ull << (2 * (KC ? KC : 1))) - 1)) : mask;
        fw = (KT)(((fw << 2) | b) & m);
        rc = (KT)((rc >> 2) | ((KT)(3u - b) << (2 * (kk - 1))));
        canon = fw < rc ? fw : rc;
    }
    // true if the k-mers starting at p0+i0 .. p0+i0+n-1 are all inside one record
    __device__ __forceinline__ bool all_valid(uint32_t i0, uint32_t n) const {
        return p0 + i0 + n - 1 + (KC ? (uint32_t)KC : k) <= nb;  // nb <= N always
    }
    // advance to the k-mer starting at p0 + i (steps must be taken in order); returns false
    // if it crosses a record boundary or the end of the sequence
    __device__ __forceinline__ bool step(uint32_t i, KT &canon) {
        roll(canon);
        const uint32_t pos = p0 + i;
        while (nb <= pos) {
            bi++;
            nb = bi < nbounds ? __ldg(&bounds[bi]) : N;
            if (bi >= nbounds) break;
        }
        const uint32_t kk = KC ? (uint32_t)KC : k;
        return pos + kk <= nb && pos + kk <= N;
    }
    // same, when all_valid() held for the block containing i
    __device__ __forceinline__ bool step_fast(uint32_t, KT &canon) {
        roll(canon);
        return true;
    }
    static __device__ __forceinline__ KT kmer_at(const SeqView &sv, uint32_t pos, uint32_t k) {
        const uint64_t f = dna_extract(sv.dna, pos, k);
        const uint64_t r = revcomp64(f, k);
        return (KT)(f < r ? f : r);
    }
};

template <typename KT>
struct SrcAA {
    const uint8_t *s;
    KT v, mask;
    uint32_t k, p0, run, N;
    __device__ __forceinline__ void init(const SeqView &sv, uint32_t p0_, uint32_t k_) {
        k = k_;
        p0 = p0_;
        N = sv.N;
        s = sv.aa;
        mask = (KT)((1ull << (5 * k)) - 1);
        v = 0;
        run = 0;
        for (uint32_t j = 0; j + 1 < k; j++) {
            const uint32_t c = (p0 + j < N) ? __ldg(&s[p0 + j]) : 0u;
            v = (KT)(((v << 5) | c) & mask);
            run = c ? run + 1 : 0;
        }
    }
    __device__ __forceinline__ bool step(uint32_t i, KT &val) {
        const uint32_t q = p0 + i + k - 1;
        const uint32_t c = q < N ? __ldg(&s[q]) : 0u;
        v = (KT)(((v << 5) | c) & mask);
        run = c ? run + 1 : 0;
        val = v;
        return run >= k;
    }
    __device__ __forceinline__ bool all_valid(uint32_t, uint32_t) const { return false; }
    __device__ __forceinline__ bool step_fast(uint32_t i, KT &val) { return step(i, val); }
    static __device__ __forceinline__ KT kmer_at(const SeqView &sv, uint32_t pos, uint32_t k) {
        uint64_t x = 0;
        for (uint32_t j = 0; j < k; j++) x = (x << 5) | __ldg(&sv.aa[pos + j]);
        return (KT)x;
    }
};

template <typename KT>
__device__ __forceinline__ uint64_t nohash_seed(KT v, uint32_t spec_flags) {
    if (spec_flags & 1u) return (uint64_t)v;  // GSB_SPEC_NOHASH_IDENTITY
    if (sizeof(KT) == 4) return (uint64_t)__byte_perm((uint32_t)v, 0, 0x0123);
    const uint64_t x = (uint64_t)v;
    return ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) |
           (uint64_t)__byte_perm((uint32_t)(x >> 32), 0, 0x0123);
}

// ------------------------------------------------------------------ prob: K2
struct OvfEntry {   // a k-mer occurrence that must go through the exact set
    uint64_t kmer;
    uint32_t pos1;  // position + 1 of this occurrence (the set entry is fingerprint | pos1)
    uint32_t cand;  // first draw below the bound?
};

template <class Src, typename KT>
__device__ __forceinline__ void set_probe_result(const SeqView &sv, uint32_t k, const ProbJob &job,
                                                 uint32_t cap, uint32_t old, KT kmer, uint32_t entry, bool cand,
                                                 uint32_t &slot, bool &act, bool &put, uint32_t &kind) {
    if (old == 0u) {  // first occurrence of this k-mer
        put = cand;
        kind = 0;
        act = false;
        return;
    }
    bool same = false;
    if ((old >> job.posbits) == (entry >> job.posbits)) {
        const uint32_t pos2 = (old & ((1u << job.posbits) - 1u)) - 1;
        same = Src::kmer_at(sv, pos2, k) == kmer;  // verified against the sequence: exact
    }
    if (same) {
        put = atomicAdd(&job.cnt[slot], 1u) == 0u;  // first repeat reports the k-mer once
        kind = 1;
        act = false;
    } else {
        slot = slot + 1 == cap ? 0 : slot + 1;
    }
}

// warp-level append of up to NI items per lane into a global list
template <int NI, class Item, class Make>
__device__ __forceinline__ void warp_list_append(uint32_t mask_bits /* per-lane item mask */, Item *list,
                                                 uint32_t *cursor, uint32_t cap, uint32_t *overflow_flag,
                                                 Make &&make) {
    const uint32_t lane = lane_id();
    const uint32_t mine = __popc(mask_bits);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
    for (int g = 0; g < NI; g++) {
        if (mask_bits & (1u << g)) {
            if (base < cap) list[base] = make(g);
            else atomicOr(overflow_flag, 1u);
            base++;
        }
    }
}

// common prologue of the chunked scan kernels: which genome / which positions
struct ChunkCtx {
    uint32_t j, p0;
    bool live;
    SeqView sv;
};
__device__ __forceinline__ ChunkCtx chunk_ctx(const uint32_t *__restrict__ chunk_prefix, uint32_t njobs,
                                              uint32_t file_of_job, const FileDesc *__restrict__ files,
                                              const FileResult *__restrict__ res,
                                              const uint32_t *__restrict__ packed_dna,
                                              const uint8_t *__restrict__ packed_aa,
                                              const uint32_t *__restrict__ boundaries, uint32_t j,
                                              uint32_t chunk) {
    ChunkCtx c;
    c.j = j;
    const FileResult fr = res[file_of_job];
    const uint32_t cbase = (chunk - chunk_prefix[j]) * kChunk;
    c.live = fr.status == 0 && cbase < fr.nsym;  // grid is sized from the byte-length upper bound
    const FileDesc fd = files[file_of_job];
    c.sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    c.sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    c.sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
    c.sv.nbounds = boundaries ? fr.nrec : 0;
    c.sv.N = fr.nsym;
    c.p0 = cbase + threadIdx.x * kRun;
    return c;
}

// ---- the first-level filter.  One 32-bit word per k-mer (chosen by a cheap hash of the k-mer):
// bits 0..23 are "seen" bits, of which a k-mer owns two; bits 24..31 are "repeated" flags, of
// which it owns one.  mark: old = atomicOr(word, seen pair); if both were already set the
// occurrence is not the first one mapping there and raises the k-mer's flag.  Atomics on one
// word are totally ordered, so a k-mer that occurs twice ALWAYS ends with its flag raised;
// a k-mer whose flag is down after the pass occurs exactly once (weight 1, exactly).  A false
// flag only sends a unique k-mer through the exact set.
struct FilterPos {
    uint32_t word, seen, flag;
};
template <typename KT>
__device__ __forceinline__ FilterPos filter_pos(KT kmer, uint32_t nwords) {
    uint32_t h = (uint32_t)kmer * 0x9E3779B1u;
    if (sizeof(KT) == 8) h += (uint32_t)((uint64_t)kmer >> 32) * 0x85EBCA77u;
    h ^= h >> 16;
    h *= 0xC2B2AE3Du;
    h ^= h >> 13;
    uint32_t g = h * 0x27D4EB2Fu;
    g ^= g >> 15;
    FilterPos f;
    f.word = __umulhi(h, nwords);
    f.seen = (1u << (((g & 0xFFu) * 24u) >> 8)) | (1u << ((((g >> 8) & 0xFFu) * 24u) >> 8));
    f.flag = 1u << (24u + ((g >> 16) & 7u));
    return f;
}

// ---- pass A: mark.  One returning atomicOr per k-mer on an L2-resident filter, all
// independent (no probing); no SplitMix64 here, only the rolling k-mer and a cheap hash.
template <class Src, typename KT>
__global__ void __launch_bounds__(kK2Threads)
k2_prob_mark(const ProbJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix, uint32_t njobs,
             const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
             const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
             const uint32_t *__restrict__ boundaries, SketchConsts sc, uint32_t nchunks) {
  // persistent CTAs (grid-stride over chunks): a bounded grid leaves room on every SM for the
  // kernels of the other group stream (this kernel waits on L2 atomics, classify on the ALU)
  for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const uint32_t j = find_file(chunk_prefix, njobs, chunk);
    const ProbJob job = jobs[j];
    const ChunkCtx cx = chunk_ctx(chunk_prefix, njobs, job.file, files, res, packed_dna, packed_aa, boundaries, j, chunk);
    if (!cx.live) continue;
    const uint32_t nwords = job.nslot1 >> 4;
    Src src;
    src.init(cx.sv, cx.p0, sc.k);
    uint32_t ncoll = 0;
#pragma unroll 1
    for (uint32_t blk = 0; blk < kRun / kG; blk++) {
        FilterPos fp[kG];
        uint32_t act = 0;
        const bool fast = __all_sync(0xffffffffu, src.all_valid(blk * kG, kG));
#pragma unroll
        for (int g = 0; g < kG; g++) {
            KT kmer;
            const bool valid = fast ? src.step_fast(blk * kG + g, kmer) : src.step(blk * kG + g, kmer);
            fp[g] = filter_pos<KT>(kmer, nwords);
            if (valid) act |= 1u << g;
        }
        uint32_t old[kG];
#pragma unroll
        for (int g = 0; g < kG; g++)
            if (act & (1u << g)) old[g] = atomicOr(&job.bitmap[fp[g].word], fp[g].seen);
#pragma unroll
        for (int g = 0; g < kG; g++) {
            if ((act & (1u << g)) && (old[g] & fp[g].seen) == fp[g].seen) {
                ncoll++;
                if (!(old[g] & fp[g].flag)) atomicOr(&job.bitmap[fp[g].word], fp[g].flag);
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) ncoll += __shfl_xor_sync(0xffffffffu, ncoll, d);
    if (lane_id() == 0 && ncoll) atomicAdd(job.n_coll, ncoll);
  }
}

// ---- pass B: classify.  A k-mer whose flag is down is unique (weight 1, exactly) and is kept
// only if its first draw can be below the bound ("light"); a k-mer whose flag is up goes
// through the exact set.  The scan only builds two 32-bit masks per thread; the few flagged
// positions are re-extracted from the packed sequence afterwards and leave through a CTA
// stage in shared memory (one global atomic per list and CTA).
constexpr uint32_t kStageHalf = kStageCap / 2;

template <class Src, typename KT>
__global__ void __launch_bounds__(kK2Threads, 3)
k2_prob_classify(const ProbJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix, uint32_t njobs,
                 const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
                 const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
                 const uint32_t *__restrict__ boundaries, const ProbBound *__restrict__ bound,
                 SketchConsts sc, uint32_t *__restrict__ overflow, uint32_t nchunks) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  ListEntry *s_lo = reinterpret_cast<ListEntry *>(s_raw);              // [kStageHalf] unique light k-mers
  OvfEntry *s_hi = reinterpret_cast<OvfEntry *>(s_raw) + kStageHalf;   // [kStageHalf] flagged occurrences
  __shared__ uint32_t s_nlo, s_nhi, s_blo, s_bhi;
  for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {  // persistent CTAs
    const uint32_t j = find_file(chunk_prefix, njobs, chunk);
    const ProbJob job = jobs[j];
    const ChunkCtx cx = chunk_ctx(chunk_prefix, njobs, job.file, files, res, packed_dna, packed_aa, boundaries, j, chunk);
    if (!cx.live) continue;  // uniform for the CTA
    const ProbBound pb = bound[j];
    __syncthreads();  // the previous chunk's flush has finished reading the stage
    if (threadIdx.x == 0) {
        s_nlo = 0;
        s_nhi = 0;
    }
    __syncthreads();
    const uint32_t nwords = job.nslot1 >> 4;
    Src src;
    src.init(cx.sv, cx.p0, sc.k);
    uint32_t m_light = 0, m_coll = 0, m_cand = 0;  // bit i = k-mer starting at p0 + i
#pragma unroll 1
    for (uint32_t blk = 0; blk < kRun / kG; blk++) {
        uint32_t word[kG], flag[kG];
        uint32_t act = 0, cand = 0;
        const bool fast = __all_sync(0xffffffffu, src.all_valid(blk * kG, kG));
#pragma unroll
        for (int g = 0; g < kG; g++) {
            KT kmer;
            const bool valid = fast ? src.step_fast(blk * kG + g, kmer) : src.step(blk * kG + g, kmer);
            const FilterPos fp = filter_pos<KT>(kmer, nwords);
            word[g] = fp.word;
            flag[g] = fp.flag;
            uint64_t s0;
            const uint64_t U = first_output(nohash_seed<KT>(kmer, sc.spec_flags), s0) >> 12;
            if ((U < pb.uT) | (U >= sc.u_slow)) cand |= 1u << g;
            if (valid) act |= 1u << g;
        }
        uint32_t w[kG];
#pragma unroll
        for (int g = 0; g < kG; g++) w[g] = (act & (1u << g)) ? __ldcg(&job.bitmap[word[g]]) : 0u;
        uint32_t coll = 0;
#pragma unroll
        for (int g = 0; g < kG; g++)
            if (w[g] & flag[g]) coll |= 1u << g;
        m_coll |= coll << (blk * kG);
        m_cand |= (cand & act) << (blk * kG);
        m_light |= (cand & act & ~coll) << (blk * kG);
    }
    // ---- emission: reserve stage cells (one shared atomic per warp and side)
    const uint32_t lane = lane_id();
    const uint32_t v = __popc(m_light) | (__popc(m_coll) << 16);
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t blo = 0, bhi = 0;
    if (lane == 31) {
        if (tot & 0xFFFFu) blo = atomicAdd(&s_nlo, tot & 0xFFFFu);
        if (tot >> 16) bhi = atomicAdd(&s_nhi, tot >> 16);
    }
    blo = __shfl_sync(0xffffffffu, blo, 31) + ((incl - v) & 0xFFFFu);
    bhi = __shfl_sync(0xffffffffu, bhi, 31) + ((incl - v) >> 16);
    uint32_t todo = m_light | m_coll;
    while (todo) {
        const uint32_t i = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t pos = cx.p0 + i;
        const KT kmer = Src::kmer_at(cx.sv, pos, sc.k);
        if ((m_coll >> i) & 1u) {
            OvfEntry e;
            e.kmer = (uint64_t)kmer;
            e.pos1 = pos + 1;
            e.cand = (m_cand >> i) & 1u;
            if (bhi < kStageHalf) {
                s_hi[bhi] = e;
            } else {  // stage full (very repetitive chunk): straight to the global list
                const uint32_t at = atomicAdd(job.ovf_n, 1u);
                if (at < job.ovf_cap) job.ovf[at] = e;
                else atomicOr(&overflow[j], 1u);
            }
            bhi++;
        } else {
            ListEntry e;
            e.kmer = (uint64_t)kmer;
            e.slot = kNoSlot;
            e.kind = 0;
            if (blo < kStageHalf) {
                s_lo[blo] = e;
            } else {
                const uint32_t at = atomicAdd(job.list_n, 1u);
                if (at < job.list_cap) job.list[at] = e;
                else atomicOr(&overflow[j], 1u);
            }
            blo++;
        }
    }
    __syncthreads();
    const uint32_t nlo = s_nlo < kStageHalf ? s_nlo : kStageHalf;
    const uint32_t nhi = s_nhi < kStageHalf ? s_nhi : kStageHalf;
    if (threadIdx.x == 0) s_blo = nlo ? atomicAdd(job.list_n, nlo) : 0u;
    if (threadIdx.x == 32) s_bhi = nhi ? atomicAdd(job.ovf_n, nhi) : 0u;
    __syncthreads();
    const uint32_t glo = s_blo, ghi = s_bhi;
    for (uint32_t t = threadIdx.x; t < nlo; t += kK2Threads) {
        if (glo + t < job.list_cap) job.list[glo + t] = s_lo[t];
        else atomicOr(&overflow[j], 1u);
    }
    for (uint32_t t = threadIdx.x; t < nhi; t += kK2Threads) {
        if (ghi + t < job.ovf_cap) job.ovf[ghi + t] = s_hi[t];
        else atomicOr(&overflow[j], 1u);
    }
  }
}

// ---- between classify and the exact set: size and clear the set from the number of flagged
// occurrences (every distinct k-mer among them takes one entry; load factor <= 1/2)
__global__ void __launch_bounds__(256)
k2_prob_mid(const ProbJob *__restrict__ jobs, uint32_t njobs) {
    const uint32_t j = blockIdx.y;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    uint32_t novf = *job.ovf_n;
    if (novf > job.ovf_cap) novf = job.ovf_cap;
    const uint64_t want = 2ull * (uint64_t)novf + 1024ull;
    const uint32_t cap2 = (uint32_t)(want < job.cap2_max ? want : job.cap2_max);
    uint4 *t4 = reinterpret_cast<uint4 *>(job.table);
    const size_t n4 = ((size_t)cap2 + 3) / 4;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        t4[i] = z;
    if (blockIdx.x == 0 && threadIdx.x == 0) *job.cap2 = cap2;
}

// exact-set insertions for the occurrences that share a filter slot: one per thread,
// linear probing, equality verified against the packed sequence
template <class Src, typename KT>
__global__ void __launch_bounds__(256)
k2_prob_overflow(const ProbJob *__restrict__ jobs, uint32_t njobs, const FileDesc *__restrict__ files,
                 const FileResult *__restrict__ res, const uint32_t *__restrict__ packed_dna,
                 const uint8_t *__restrict__ packed_aa, SketchConsts sc, uint32_t *__restrict__ overflow) {
    const uint32_t j = blockIdx.y;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    const FileResult fr = res[job.file];
    if (fr.status != 0) return;
    const FileDesc fd = files[job.file];
    SeqView sv;
    sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    sv.bounds = nullptr;
    sv.nbounds = 0;
    sv.N = fr.nsym;
    uint32_t n = *job.ovf_n;
    if (n > job.ovf_cap) n = job.ovf_cap;
    const uint32_t cap2 = *job.cap2;
    const uint32_t stride = gridDim.x * blockDim.x;
    // whole warps iterate together so that the warp-level append below stays convergent
    for (uint32_t e0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; e0 < n; e0 += stride) {
        const uint32_t e = e0 + lane_id();
        bool act = e < n, put = false;
        uint32_t kind = 0, slot = 0, entry = 0;
        KT kmer = 0;
        bool cand = false;
        if (act) {
            const OvfEntry oe = job.ovf[e];
            kmer = (KT)oe.kmer;
            cand = oe.cand != 0;
            const uint64_t s0 = sm64_mix(nohash_seed<KT>(kmer, sc.spec_flags) + kGolden);
            entry = ((uint32_t)s0 << job.posbits) | oe.pos1;
            slot = __umulhi((uint32_t)(s0 >> 32), cap2);
        }
        while (act) {
            const uint32_t old = atomicCAS(&job.table[slot], 0u, entry);
            set_probe_result<Src, KT>(sv, sc.k, job, cap2, old, kmer, entry, cand, slot, act, put, kind);
        }
        warp_list_append<1>(put ? 1u : 0u, job.list, job.list_n, job.list_cap, &overflow[j], [&](int) {
            ListEntry le;
            le.kmer = (uint64_t)kmer;
            le.slot = slot;
            le.kind = kind;
            return le;
        });
    }
}

// ------------------------------------------------------------------ prob: K3
// replay of ProbMinHash3a::hashset for one weighted k-mer against the static bound T
template <typename KT, class F>
__device__ __forceinline__ void pmh_points(KT d, uint32_t w, double T, const SketchConsts &sc,
                                           F &&emit) {
    Xoshiro rng;
    rng.seed(nohash_seed<KT>(d, sc.spec_flags));
    const double winv = __ddiv_rn(1.0, (double)w);
    double x = exp01_sample_from(sc.e01, u01_f64_from_bits(rng.next()), rng);
    double h = __dmul_rn(winv, x);
    if (!(h < T)) return;
    uint32_t k = uniform_usize(rng, sc.m, sc.zone);
    emit(h, k);
    if (!(winv < T)) return;
    for (uint32_t i = 2;; i++) {
        const double h0 = __dmul_rn(winv, (double)(i - 1));
        if (!(h0 < T)) return;
        x = exp01_sample_from(sc.e01, u01_f64_from_bits(rng.next()), rng);
        h = __dadd_rn(h0, __dmul_rn(winv, x));
        k = uniform_usize(rng, sc.m, sc.zone);
        if (h < T) emit(h, k);
        if (!(__dmul_rn(winv, (double)i) < T)) return;
    }
}

// PASS 0: atomicMin of h ; PASS 1: owners of the minimum write the k-mer
template <typename KT, int PASS>
__global__ void __launch_bounds__(256)
k3_prob_points(const ProbJob *__restrict__ jobs, uint32_t njobs,
               const ProbBound *__restrict__ bound, SketchConsts sc) {
    const uint32_t j = blockIdx.y;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    uint32_t n = *job.list_n;
    if (n > job.list_cap) n = job.list_cap;
    const double T = bound[j].T;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const ListEntry le = job.list[e];
        uint32_t w;
        if (le.kind == 2) {
            w = le.slot;
        } else {
            const uint32_t extra = le.slot == kNoSlot ? 0u : job.cnt[le.slot];
            if (le.kind == 0 && extra != 0) continue;  // handled through its "repeated" entry
            w = 1u + extra;
        }
        const KT d = (KT)le.kmer;
        pmh_points<KT>(d, w, T, sc, [&](double h, uint32_t k) {
            const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
            if (PASS == 0) {
                atomicMin(&job.hmin[k], hb);
            } else {
                if (job.hmin[k] == hb) atomicMin(&job.sigw[k], (unsigned long long)d);
            }
        });
    }
}

// per genome: check the bound, convert to the output element type, reset touched counters.
// grid = (kFinParts, njobs): the slots of a genome are split over several CTAs (one CTA per
// genome walked 18 000 slots in 70 dependent rounds); `retry` is zeroed before the pass.
constexpr uint32_t kFinParts = 8;
template <typename SigT>
__global__ void __launch_bounds__(256)
k3_prob_finalize(const ProbJob *__restrict__ jobs, uint32_t njobs,
                 const ProbBound *__restrict__ bound, const FileResult *__restrict__ res,
                 SketchConsts sc, SigT *__restrict__ sig_out, uint64_t *__restrict__ nb_bases_out,
                 uint32_t *__restrict__ retry) {
    const uint32_t j = blockIdx.y;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    const unsigned long long Tb = (unsigned long long)__double_as_longlong(bound[j].T);
    const uint32_t N = res[job.file].nsym;
    const bool has_kmers = N >= sc.k && res[job.file].status == 0;
    bool over = false;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < sc.m; k += gridDim.x * blockDim.x) {
        // exact iff every slot's minimum is strictly below the bound (no skipped point can win
        // or tie); an empty genome is trivially exact
        over |= !(job.hmin[k] < Tb);
        const unsigned long long s = job.sigw[k];
        sig_out[(size_t)job.file * sc.m + k] = (s == ~0ull) ? (SigT)0 : (SigT)s;
    }
    if (__syncthreads_or(over && has_kmers) && threadIdx.x == 0) atomicOr(&retry[job.file], 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (res[job.file].status) atomicOr(&retry[job.file], res[job.file].status << 8);
        if (nb_bases_out) nb_bases_out[job.file] = res[job.file].nbases;
        // the extra-occurrence counters touched by this genome are cleared by the slot's next
        // k_prob_reset (full grid) from the list left here
        const uint32_t n = *job.list_n;
        *job.prev_n = job.newpath ? 0u : (n > job.list_cap ? job.list_cap : n);
    }
}

// ------------------------------------------------------------------ optdens
struct DensJob {
    uint32_t file;
    uint32_t *bins;  // [m] f32 bit patterns, atomicMin
    double tmult;
};

constexpr uint32_t kDensLargeBits = 0x4F800000u;  // 4294967296.0f = F::from(u32::MAX)

// RevOptDens: bin targeted by non-empty bin i in round a (oracle/sketch.c gso_revdens_target)
__device__ __forceinline__ uint32_t revdens_target(uint32_t i, uint32_t a, uint32_t m) {
    uint64_t z = (uint64_t)i * 0x9e3779b97f4a7c15ULL + (uint64_t)a * 0xd1b54a32d192ed03ULL + 0x2545f4914f6cdd1dULL;
    z = sm64_mix(z);
    return (uint32_t)__umul64hi(z, (uint64_t)m);
}

template <class Src, typename KT>
__global__ void __launch_bounds__(kK2Threads)
k2_optdens(const DensJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix,
           uint32_t njobs, const FileDesc *__restrict__ files,
           const FileResult *__restrict__ res, const uint32_t *__restrict__ packed_dna,
           const uint8_t *__restrict__ packed_aa, const uint32_t *__restrict__ boundaries,
           SketchConsts sc, uint32_t nchunks) {
  for (uint32_t gchunk = blockIdx.x; gchunk < nchunks; gchunk += gridDim.x) {  // persistent CTAs
    const uint32_t j = find_file(chunk_prefix, njobs, gchunk);
    const DensJob job = jobs[j];
    const FileResult fr = res[job.file];
    if (fr.status != 0) continue;
    const uint32_t chunk = gchunk - chunk_prefix[j];
    const uint32_t cbase = chunk * kChunk;
    if (cbase >= fr.nsym) continue;
    const FileDesc fd = files[job.file];
    SeqView sv;
    sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
    sv.nbounds = boundaries ? fr.nrec : 0;
    sv.N = fr.nsym;
    // bound on r: bins hold the minimum of ~nk/m draws; skip draws that cannot be a minimum
    const uint32_t nk = fr.nsym >= sc.k ? fr.nsym - sc.k + 1 : 0;
    const double T = nk ? job.tmult * ((double)sc.m / (double)nk) * sc.lnm8 : 2.0;
    const float Tf = T >= 1.5 ? 2.0f : (float)T;
    const bool f64draw = (sc.spec_flags & 2u) != 0;  // GSB_SPEC_OPTDENS_F64_DRAW
    const uint32_t p0 = cbase + threadIdx.x * kRun;
    Src src;
    src.init(sv, p0, sc.k);
    for (uint32_t i = 0; i < kRun; i++) {
        KT val;
        if (!src.step(i, val)) continue;
        const uint64_t seed = (uint64_t)val * kFxSeed64;
        uint64_t s0;
        const uint64_t out1 = first_output(seed, s0);
        const float r = f64draw ? __double2float_rn(u01_f64_from_bits(out1)) : u01_f32_from_bits(out1);
        if (!(r < Tf)) continue;
        Xoshiro rng;
        rng.seed(seed);
        (void)rng.next();
        const uint32_t k = uniform_usize(rng, sc.m, sc.zone);
        atomicMin(&job.bins[k], __float_as_uint(r));
    }
  }
}

// per genome: check the bound, densify empty bins (cold path), write f32 signature
__global__ void __launch_bounds__(256)
k3_optdens_finalize(const DensJob *__restrict__ jobs, uint32_t njobs,
                    const FileResult *__restrict__ res, SketchConsts sc,
                    float *__restrict__ sig_out, uint64_t *__restrict__ nb_bases_out,
                    uint32_t *__restrict__ retry, int super_mode, uint32_t *__restrict__ rev_win /* REVOPTDENS: [njobs][m] */) {
    const uint32_t j = blockIdx.x;
    if (j >= njobs) return;
    const DensJob job = jobs[j];
    const FileResult fr = res[job.file];
    __shared__ uint32_t smax[256];
    __shared__ uint32_t snempty;
    if (threadIdx.x == 0) snempty = 0;
    __syncthreads();
    const uint32_t nk = fr.nsym >= sc.k ? fr.nsym - sc.k + 1 : 0;
    const double T = nk ? job.tmult * ((double)sc.m / (double)nk) * sc.lnm8 : 2.0;
    const bool bounded = T < 1.5;
    const float Tf = bounded ? (float)T : 2.0f;
    uint32_t mx = 0, ne = 0;
    for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
        const uint32_t b = job.bins[k];
        mx = b > mx ? b : mx;
        ne += (b == kDensLargeBits);
    }
    smax[threadIdx.x] = mx;
    if (ne) atomicAdd(&snempty, ne);
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (threadIdx.x < d && smax[threadIdx.x + d] > smax[threadIdx.x])
            smax[threadIdx.x] = smax[threadIdx.x + d];
        __syncthreads();
    }
    const bool ok_status = fr.status == 0;
    // with a bound in force every bin must have ended strictly below it
    const bool need_retry = ok_status && nk > 0 && bounded && !(__uint_as_float(smax[0]) < Tf);
    const uint32_t nempty = snempty;
    // SuperMinHash: the bins ARE the signature when every slot was reached by a first-level
    // value (r + 0 < 1 <= any later level); otherwise the file goes to the sequential kernel
    const bool need_seq = super_mode && ok_status && nk > 0 && !need_retry && nempty != 0;
    if (threadIdx.x == 0) {
        retry[job.file] = (need_retry ? 1u : 0u) | (need_seq ? 2u : 0u) | (fr.status << 8);
        if (nb_bases_out) nb_bases_out[job.file] = fr.nbases;
    }
    float *out = sig_out + (size_t)job.file * sc.m;
    if (rev_win && nempty != 0 && nempty != sc.m && !need_retry) {
        // RevOptDens (SPEC: oracle/sketch.c gso_revoptdens): non-empty bins push their value into empty
        // ones, round after round; in a round the lowest pushing bin wins an empty target
        uint32_t *win = rev_win + (size_t)j * sc.m;
        __shared__ uint32_t s_left;
        for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
            out[k] = __uint_as_float(job.bins[k]);
            win[k] = 0xFFFFFFFFu;
        }
        if (threadIdx.x == 0) s_left = nempty;
        __syncthreads();
        for (uint32_t a = 0; s_left > 0; a++) {
            for (uint32_t i = threadIdx.x; i < sc.m; i += blockDim.x) {
                if (job.bins[i] == kDensLargeBits) continue;
                const uint32_t t = revdens_target(i, a, sc.m);
                if (__float_as_uint(out[t]) == kDensLargeBits) atomicMin(&win[t], i);
            }
            __syncthreads();
            uint32_t filled = 0;
            for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
                const uint32_t w = win[k];
                if (w != 0xFFFFFFFFu) {
                    out[k] = __uint_as_float(job.bins[w]);
                    win[k] = 0xFFFFFFFFu;
                    filled++;
                }
            }
            if (filled) atomicSub(&s_left, filled);
            __syncthreads();
        }
        return;
    }
    for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
        uint32_t b = job.bins[k];
        if (b == kDensLargeBits && nempty != sc.m && !need_retry && !super_mode) {
            // end_sketch(): densification of an empty bin (SPEC: rng seeded by the bin index,
            // draw j until bin j was filled before densification)
            Xoshiro rng;
            rng.seed((uint64_t)k);
            for (;;) {
                const uint32_t jj = uniform_usize(rng, sc.m, sc.zone);
                const uint32_t bj = job.bins[jj];
                if (bj != kDensLargeBits) {
                    b = bj;
                    break;
                }
            }
        }
        out[k] = __uint_as_float(b);
    }
}

// ------------------------------------------------------------------ SuperMinHash, cold path
// probminhash SuperMinHash::sketch [U; oracle/sketch.c gso_superminhash] restated
// sequentially for the files whose k-mers do not reach every slot at the first level (fewer
// k-mers than about m ln m: tiny inputs).  One thread per file; q/p/b live in global scratch.
template <class Src, typename KT>
__global__ void k_super_sequential(const uint32_t *__restrict__ file_list, uint32_t nlist,
                                   const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
                                   const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
                                   const uint32_t *__restrict__ boundaries, SketchConsts sc,
                                   float *__restrict__ sig_out, uint32_t *__restrict__ scratch) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nlist) return;
    const uint32_t f = file_list[li];
    const FileResult fr = res[f];
    const FileDesc fd = files[f];
    const uint32_t m = sc.m;
    float *sig = sig_out + (size_t)f * m;
    uint32_t *q = scratch + (size_t)li * 3 * m, *p = q + m;
    int32_t *b = reinterpret_cast<int32_t *>(p + m);
    for (uint32_t i = 0; i < m; i++) {
        sig[i] = __uint_as_float(kDensLargeBits);
        q[i] = 0xFFFFFFFFu;
        p[i] = 0;
        b[i] = 0;
    }
    b[m - 1] = (int32_t)m;
    uint32_t a_upper = m - 1;
    SeqView sv;
    sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
    sv.nbounds = boundaries ? fr.nrec : 0;
    sv.N = fr.nsym;
    for (uint32_t p0 = 0; p0 < fr.nsym; p0 += kRun) {
        Src src;
        src.init(sv, p0, sc.k);
        for (uint32_t i = 0; i < kRun; i++) {
            KT val;
            if (!src.step(i, val)) continue;
            const uint32_t irank = p0 + i;
            Xoshiro rng;
            rng.seed((uint64_t)val * kFxSeed64);
            uint32_t j = 0;
            while (j <= a_upper) {
                const float r = u01_f32_from_bits(rng.next());
                const uint64_t range = (uint64_t)(m - j);
                const uint32_t k = j + uniform_usize(rng, range, UINT64_MAX - ((UINT64_MAX - range + 1) % range));
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
                const float rpj = __fadd_rn(r, (float)j);
                const uint32_t slot = p[j];
                if (rpj < sig[slot]) {
                    const float old = sig[slot];
                    const uint32_t j2 = (old >= (float)(m - 1)) ? (m - 1) : (uint32_t)old;
                    sig[slot] = rpj;
                    if (j < j2) {
                        b[j2] -= 1;
                        b[j] += 1;
                        while (b[a_upper] == 0) a_upper--;
                    }
                }
                j++;
            }
        }
    }
}

// ------------------------------------------------------------------ SuperMinHash2
// probminhash SuperMinHash2 [U; oracle/sketch.c gso_superminhash2]: per slot the fx hash of the item
// that gave the minimum.  Level-0 values (r < 1) beat every later level, so when each slot is reached
// at level 0 the signature is, per slot, the hash of the item with the smallest first draw: ONE
// 128-bit minimum (bits of r, hash) per slot, merged with compare-and-swap.  Otherwise (fewer k-mers
// than ~ m ln m) the file takes the sequential restatement.
__device__ __forceinline__ uint64_t fx_hash_of(uint64_t v, bool kt32) {
    return kt32 ? (uint64_t)((uint32_t)v * 0x9e3779b9u) : v * kFxSeed64;
}
constexpr unsigned long long kSuper2Large = 0x41F0000000000000ull;  // 4294967296.0

__global__ void k_super2_reset(ulonglong2 *slots, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        slots[i] = make_ulonglong2(kSuper2Large, ~0ull);
}

template <class Src, typename KT>
__global__ void __launch_bounds__(kK2Threads)
k2_super2(const DensJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix, uint32_t njobs,
          const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
          const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
          const uint32_t *__restrict__ boundaries, SketchConsts sc, uint32_t nchunks) {
    for (uint32_t gchunk = blockIdx.x; gchunk < nchunks; gchunk += gridDim.x) {  // persistent CTAs
        const uint32_t j = find_file(chunk_prefix, njobs, gchunk);
        const DensJob job = jobs[j];
        const FileResult fr = res[job.file];
        if (fr.status != 0) continue;
        const uint32_t cbase = (gchunk - chunk_prefix[j]) * kChunk;
        if (cbase >= fr.nsym) continue;
        const FileDesc fd = files[job.file];
        SeqView sv;
        sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
        sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
        sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
        sv.nbounds = boundaries ? fr.nrec : 0;
        sv.N = fr.nsym;
        const uint32_t nk = fr.nsym >= sc.k ? fr.nsym - sc.k + 1 : 0;
        const double T = nk ? job.tmult * ((double)sc.m / (double)nk) * sc.lnm8 : 2.0;
        ulonglong2 *slots = reinterpret_cast<ulonglong2 *>(job.bins);
        Src src;
        src.init(sv, cbase + threadIdx.x * kRun, sc.k);
        for (uint32_t i = 0; i < kRun; i++) {
            KT val;
            if (!src.step(i, val)) continue;
            const uint64_t hv = fx_hash_of((uint64_t)val, sizeof(KT) == 4);
            uint64_t s0;
            const double r = u01_f64_from_bits(first_output(hv, s0));
            if (!(r < T)) continue;
            Xoshiro rng;
            rng.seed(hv);
            (void)rng.next();
            const uint32_t k = uniform_usize(rng, sc.m, sc.zone);
            slot_min128(&slots[k], (unsigned long long)__double_as_longlong(r), hv);
        }
    }
}

template <typename SigT>
__global__ void __launch_bounds__(256)
k3_super2_finalize(const DensJob *__restrict__ jobs, uint32_t njobs, const FileResult *__restrict__ res,
                   SketchConsts sc, SigT *__restrict__ sig_out, uint64_t *__restrict__ nb_bases_out,
                   uint32_t *__restrict__ retry) {
    const uint32_t j = blockIdx.x;
    if (j >= njobs) return;
    const DensJob job = jobs[j];
    const FileResult fr = res[job.file];
    const ulonglong2 *slots = reinterpret_cast<const ulonglong2 *>(job.bins);
    const uint32_t nk = fr.nsym >= sc.k ? fr.nsym - sc.k + 1 : 0;
    const double T = nk ? job.tmult * ((double)sc.m / (double)nk) * sc.lnm8 : 2.0;
    const bool bounded = T < 1.0;
    const unsigned long long Tb = (unsigned long long)__double_as_longlong(bounded ? T : 1.0);
    bool over = false, empty = false;
    for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
        const ulonglong2 s = slots[k];
        empty |= s.y == ~0ull && s.x == kSuper2Large;
        over |= !(s.x < Tb);  // an empty slot is "over" too
        sig_out[(size_t)job.file * sc.m + k] = (SigT)s.y;
    }
    const bool any_over = __syncthreads_or(over), any_empty = __syncthreads_or(empty);
    if (threadIdx.x == 0) {
        const bool ok = fr.status == 0 && nk > 0;
        // with a bound in force a slot at or above it may hide a smaller skipped draw: widen and re-run;
        // without a bound (T >= 1) a slot that level 0 never reached needs the later levels: sequential
        const bool need_retry = ok && bounded && any_over;
        const bool need_seq = ok && !need_retry && any_empty;
        retry[job.file] = (need_retry ? 1u : 0u) | (need_seq ? 2u : 0u) | (fr.status << 8);
        if (nb_bases_out) nb_bases_out[job.file] = fr.nbases;
    }
}

// cold path: the sequential algorithm, one thread per file; q/p/b/h/v live in global scratch
template <class Src, typename KT, typename SigT>
__global__ void k_super2_sequential(const uint32_t *__restrict__ file_list, uint32_t nlist,
                                    const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
                                    const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
                                    const uint32_t *__restrict__ boundaries, SketchConsts sc,
                                    SigT *__restrict__ sig_out, uint8_t *__restrict__ scratch) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nlist) return;
    const uint32_t f = file_list[li];
    const FileResult fr = res[f];
    const FileDesc fd = files[f];
    const uint32_t m = sc.m;
    SigT *sig = sig_out + (size_t)f * m;
    uint8_t *base = scratch + (size_t)li * m * 28;
    double *h = reinterpret_cast<double *>(base);
    uint64_t *v = reinterpret_cast<uint64_t *>(base + (size_t)m * 8);
    uint32_t *q = reinterpret_cast<uint32_t *>(base + (size_t)m * 16), *p = q + m;
    int32_t *b = reinterpret_cast<int32_t *>(p + m);
    for (uint32_t i = 0; i < m; i++) {
        h[i] = 4294967296.0;
        v[i] = 0;
        q[i] = 0xFFFFFFFFu;
        p[i] = 0;
        b[i] = 0;
    }
    b[m - 1] = (int32_t)m;
    uint32_t a_upper = m - 1;
    SeqView sv;
    sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
    sv.nbounds = boundaries ? fr.nrec : 0;
    sv.N = fr.nsym;
    for (uint32_t p0 = 0; p0 < fr.nsym; p0 += kRun) {
        Src src;
        src.init(sv, p0, sc.k);
        for (uint32_t i = 0; i < kRun; i++) {
            KT val;
            if (!src.step(i, val)) continue;
            const uint32_t irank = p0 + i;
            const uint64_t hv = fx_hash_of((uint64_t)val, sizeof(KT) == 4);
            Xoshiro rng;
            rng.seed(hv);
            uint32_t jj = 0;
            while (jj <= a_upper) {
                const double r = u01_f64_from_bits(rng.next());
                const uint64_t range = (uint64_t)(m - jj);
                const uint32_t k = jj + uniform_usize(rng, range, UINT64_MAX - ((UINT64_MAX - range + 1) % range));
                if (q[jj] != irank) {
                    q[jj] = irank;
                    p[jj] = jj;
                }
                if (q[k] != irank) {
                    q[k] = irank;
                    p[k] = k;
                }
                const uint32_t t = p[jj];
                p[jj] = p[k];
                p[k] = t;
                const double rpj = __dadd_rn(r, (double)jj);
                const uint32_t slot = p[jj];
                if (rpj < h[slot] || (rpj == h[slot] && hv < v[slot])) {
                    const double old = h[slot];
                    const uint32_t j2 = (old >= (double)(m - 1)) ? (m - 1) : (uint32_t)old;
                    h[slot] = rpj;
                    v[slot] = hv;
                    if (jj < j2) {
                        b[j2] -= 1;
                        b[jj] += 1;
                        while (b[a_upper] == 0) a_upper--;
                    }
                }
                jj++;
            }
        }
    }
    for (uint32_t i = 0; i < m; i++) sig[i] = (SigT)v[i];
}

// ====================================================================== SetSketch ("--algo hll")
// HyperLogLogSketch<Kmer, u16> = probminhash SetSketcher [U] (dispatch src/dna/dnasketch.rs:541-573);
// SPEC: oracle/sketch.c gso_setsketch.  Order-free definition: register i = max over the k-mers and
// their points of k = clamp(floor(1 - log_b x_j), 0, 65535) for the point the k-mer's lazy
// Fisher-Yates permutation sends to i.  Like the other sketchers the scan uses a STATIC bound:
// points with x > X_hi = (1/a)/m * T cannot raise a register that ends at or above k(X_hi) -- checked
// by the finalize kernel, else the genome is re-run with a larger bound -- so all but ~6 % of the
// k-mers are dismissed by an integer compare on their first draw.
constexpr double kHllB = 1.001, kHllInvA = 1.0 / 20.0;
constexpr uint32_t kHllMaxSteps = 16;         // points of one k-mer below the bound that the scan follows
constexpr uint32_t kHllOverflow = 0xFFFFFFFFu;  // register 0 holds this: the genome needs the sequential path
constexpr double kHllSeqT = 0.25;             // bound (in unit-exponential scale) from which a file goes sequential

__device__ __forceinline__ uint32_t hll_k_of(double x, double lnb) {
    if (!(x > 0.0)) return 65535u;
    const double z = __dadd_rn(1.0, -__ddiv_rn(ln_spec(x), lnb));
    const double kf = floor(z);
    return kf < 0.0 ? 0u : (kf > 65535.0 ? 65535u : (uint32_t)kf);
}

__device__ __forceinline__ double hll_T(const DensJob &job, const FileResult &fr, const SketchConsts &sc) {
    const uint32_t nk = fr.nsym >= sc.k ? fr.nsym - sc.k + 1 : 0;
    return nk ? job.tmult * ((double)sc.m / (double)nk) * sc.lnm8 : 2.0;
}

template <class Src, typename KT>
__global__ void __launch_bounds__(kK2Threads)
k2_hll(const DensJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix, uint32_t njobs,
       const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
       const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
       const uint32_t *__restrict__ boundaries, SketchConsts sc, uint32_t nchunks) {
  const double lnb = ln_spec(kHllB);
  for (uint32_t gchunk = blockIdx.x; gchunk < nchunks; gchunk += gridDim.x) {  // persistent CTAs
    const uint32_t j = find_file(chunk_prefix, njobs, gchunk);
    const DensJob job = jobs[j];
    const FileResult fr = res[job.file];
    if (fr.status != 0) continue;
    const uint32_t chunk = gchunk - chunk_prefix[j];
    const uint32_t cbase = chunk * kChunk;
    if (cbase >= fr.nsym) continue;
    const double T = hll_T(job, fr, sc);
    if (!(T < kHllSeqT)) continue;  // small input: the sequential kernel does the whole file
    const FileDesc fd = files[job.file];
    SeqView sv;
    sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
    sv.nbounds = boundaries ? fr.nrec : 0;
    sv.N = fr.nsym;
    const double x_hi = __dmul_rn(__ddiv_rn(kHllInvA, (double)sc.m), T);
    // first draw E = -ln(1 - U) >= 1.001 T  <=>  U >= 1 - exp(-1.001 T): dismissed without the logarithm
    const double uthr = 1.0 - exp(-1.001 * T);
    const uint64_t thr52 = (uint64_t)(uthr * 4503599627370496.0) + 2;
    const uint32_t p0 = cbase + threadIdx.x * kRun;
    Src src;
    src.init(sv, p0, sc.k);
    for (uint32_t i = 0; i < kRun; i++) {
        KT val;
        if (!src.step(i, val)) continue;
        const uint64_t seed = (uint64_t)val * kFxSeed64;
        uint64_t s0;
        const uint64_t out1 = first_output(seed, s0);
        if ((out1 >> 12) >= thr52) continue;
        Xoshiro rng;
        rng.seed(seed);
        double x = 0.0;
        uint32_t mapk[kHllMaxSteps], mapv[kHllMaxSteps], nmap = 0;
        for (uint32_t jj = 0; jj < sc.m; jj++) {
            const double e = -ln_spec(__dadd_rn(1.0, -u01_f64_from_bits(rng.next())));
            x = __dadd_rn(x, __dmul_rn(__ddiv_rn(kHllInvA, (double)(sc.m - jj)), e));
            if (x > x_hi) break;
            const uint32_t kk = hll_k_of(x, lnb);
            if (kk == 0) break;
            if (jj == kHllMaxSteps) {  // more points below the bound than the scan follows
                atomicMax(&job.bins[0], kHllOverflow);
                break;
            }
            const uint64_t range = (uint64_t)(sc.m - jj);
            const uint32_t r = jj + uniform_usize(rng, range, UINT64_MAX - ((UINT64_MAX - range + 1) % range));
            // lazy Fisher-Yates: position jj is never read again, so only the entries written at r > jj live on
            uint32_t a = jj, b = r;
            for (uint32_t q = 0; q < nmap; q++) {
                if (mapk[q] == jj) a = mapv[q];
                if (mapk[q] == r) b = mapv[q];
            }
            uint32_t reg = a;
            if (r != jj) {
                reg = b;
                uint32_t q = 0;
                while (q < nmap && mapk[q] != r) q++;
                mapk[q] = r;
                mapv[q] = a;
                if (q == nmap) nmap++;
            }
            atomicMax(&job.bins[reg], kk);
        }
    }
  }
}

// per genome: check the bound, write the u16 registers
__global__ void __launch_bounds__(256)
k3_hll_finalize(const DensJob *__restrict__ jobs, uint32_t njobs, const FileResult *__restrict__ res,
                SketchConsts sc, uint16_t *__restrict__ sig_out, uint64_t *__restrict__ nb_bases_out,
                uint32_t *__restrict__ retry) {
    const uint32_t j = blockIdx.x;
    if (j >= njobs) return;
    const DensJob job = jobs[j];
    const FileResult fr = res[job.file];
    __shared__ uint32_t smin[256], smax[256];
    const double T = hll_T(job, fr, sc);
    const bool bounded = T < kHllSeqT;
    const uint32_t nk = fr.nsym >= sc.k ? fr.nsym - sc.k + 1 : 0;
    uint32_t mn = 0xFFFFFFFFu, mx = 0;
    for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
        const uint32_t b = job.bins[k];
        mn = b < mn ? b : mn;
        mx = b > mx ? b : mx;
    }
    smin[threadIdx.x] = mn;
    smax[threadIdx.x] = mx;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (threadIdx.x < d) {
            if (smin[threadIdx.x + d] < smin[threadIdx.x]) smin[threadIdx.x] = smin[threadIdx.x + d];
            if (smax[threadIdx.x + d] > smax[threadIdx.x]) smax[threadIdx.x] = smax[threadIdx.x + d];
        }
        __syncthreads();
    }
    const bool ok_status = fr.status == 0;
    const bool overflow = smax[0] == kHllOverflow;
    // every register must have ended at or above k(X_hi): then no dismissed point could have raised one
    const double x_hi = __dmul_rn(__ddiv_rn(kHllInvA, (double)sc.m), T);
    const uint32_t k_hi = bounded ? hll_k_of(x_hi, ln_spec(kHllB)) : 0u;
    const bool need_seq = ok_status && nk > 0 && (!bounded || overflow);
    const bool need_retry = ok_status && nk > 0 && bounded && !overflow && smin[0] < k_hi;
    if (threadIdx.x == 0) {
        retry[job.file] = (need_retry ? 1u : 0u) | (need_seq ? 2u : 0u) | (fr.status << 8);
        if (nb_bases_out) nb_bases_out[job.file] = fr.nbases;
    }
    uint16_t *out = sig_out + (size_t)job.file * sc.m;
    for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) out[k] = (uint16_t)job.bins[k];
}

// cold path (small inputs): the sequential algorithm of the SPEC, one thread per file; q/p in scratch
template <class Src, typename KT>
__global__ void k_hll_sequential(const uint32_t *__restrict__ file_list, uint32_t nlist,
                                 const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
                                 const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
                                 const uint32_t *__restrict__ boundaries, SketchConsts sc,
                                 uint16_t *__restrict__ sig_out, uint32_t *__restrict__ scratch) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nlist) return;
    const uint32_t f = file_list[li];
    const FileResult fr = res[f];
    const FileDesc fd = files[f];
    const uint32_t m = sc.m;
    uint16_t *sig = sig_out + (size_t)f * m;
    uint32_t *q = scratch + (size_t)li * m * 2, *p = q + m;
    for (uint32_t i = 0; i < m; i++) {
        sig[i] = 0;
        q[i] = 0xFFFFFFFFu;
        p[i] = 0;
    }
    const double lnb = ln_spec(kHllB);
    uint32_t k_low = 0;
    unsigned long long nbmin = 0;
    SeqView sv;
    sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
    sv.nbounds = boundaries ? fr.nrec : 0;
    sv.N = fr.nsym;
    for (uint32_t p0 = 0; p0 < fr.nsym; p0 += kRun) {
        Src src;
        src.init(sv, p0, sc.k);
        for (uint32_t i = 0; i < kRun; i++) {
            KT val;
            if (!src.step(i, val)) continue;
            const uint32_t irank = p0 + i;
            Xoshiro rng;
            rng.seed((uint64_t)val * kFxSeed64);
            double x = 0.0;
            for (uint32_t jj = 0; jj < m; jj++) {
                const double e = -ln_spec(__dadd_rn(1.0, -u01_f64_from_bits(rng.next())));
                x = __dadd_rn(x, __dmul_rn(__ddiv_rn(kHllInvA, (double)(m - jj)), e));
                const uint32_t kk = hll_k_of(x, lnb);
                if (kk <= k_low) break;
                const uint64_t range = (uint64_t)(m - jj);
                const uint32_t r = jj + uniform_usize(rng, range, UINT64_MAX - ((UINT64_MAX - range + 1) % range));
                if (q[jj] != irank) {
                    q[jj] = irank;
                    p[jj] = jj;
                }
                if (q[r] != irank) {
                    q[r] = irank;
                    p[r] = r;
                }
                const uint32_t t = p[jj];
                p[jj] = p[r];
                p[r] = t;
                const uint32_t reg = p[jj];
                if (kk > sig[reg]) {
                    sig[reg] = (uint16_t)kk;
                    if (++nbmin % m == 0) {
                        uint32_t mn = 65535;
                        for (uint32_t t2 = 0; t2 < m; t2++) mn = sig[t2] < mn ? sig[t2] : mn;
                        k_low = mn;
                    }
                }
            }
        }
    }
}

}  // namespace gsb