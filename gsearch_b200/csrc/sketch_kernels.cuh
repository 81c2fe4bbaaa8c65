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
//   * K2 streams k-mers, keeps an EXACT set of the genome's distinct canonical k-mers in a
//     hash set of 32-bit entries (8-bit fingerprint | 24-bit position of first occurrence;
//     equality is verified against the packed sequence, so the set is exact), counts extra
//     occurrences, and emits only k-mers that can matter: first draw below T ("light"), or
//     repeated (weight > 1).
//   * K3a/K3b replay the exact f64 arithmetic for those few k-mers: atomicMin on the ordered
//     bit pattern of h, then the owner of the minimum writes the k-mer (ties: smaller k-mer).
#pragma once

#include "common.cuh"
#include "fasta_pack.cuh"

namespace gsb {

constexpr int kK2Threads = 256;
constexpr int kRun = 32;                        // consecutive k-mer positions per thread
constexpr int kChunk = kK2Threads * kRun;       // positions per CTA
constexpr uint32_t kMaxProbSym = (1u << 24) - 2;  // 24-bit position field of a set entry

struct ListEntry {
    uint64_t kmer;
    uint32_t slot;
    uint32_t kind;  // 0 = light (first occurrence, first draw below T), 1 = repeated
};

// per-genome device-side job state for the prob path
struct ProbJob {
    uint32_t file;       // index into FileDesc / FileResult
    uint32_t cap;        // hash-set capacity (entries)
    uint32_t *table;     // hash set
    uint32_t *cnt;       // extra occurrences per slot
    ListEntry *list;     // candidates
    uint32_t list_cap;
    uint32_t *list_n;    // cursor (device)
    uint32_t *prev_n;    // entries of the slot's previous genome whose counters are still set
    unsigned long long *hmin;  // [m] ordered bits of min h per slot
    unsigned long long *sigw;  // [m] winning k-mer per slot
    double tmult;        // early-stop bound multiplier (1 = default, grown on retry)
};

struct ProbBound {  // written by k_prob_setup
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

template <typename KT>
struct SrcDNA {
    uint64_t W0, W1;
    KT fw, rc, mask;
    uint32_t k, p0, nb, bi;
    const uint32_t *bounds;
    uint32_t nbounds, N;

    __device__ __forceinline__ uint32_t base(uint32_t j) const {
        const uint64_t w = j < 32 ? W0 : W1;
        return (uint32_t)(w >> (62 - 2 * (j & 31u))) & 3u;
    }
    __device__ __forceinline__ void init(const SeqView &sv, uint32_t p0_, uint32_t k_) {
        k = k_;
        p0 = p0_;
        N = sv.N;
        bounds = sv.bounds;
        nbounds = sv.nbounds;
        const uint32_t wi = p0 >> 4;  // p0 is a multiple of 32 -> wi even
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(sv.dna + wi));
        const uint2 b = __ldg(reinterpret_cast<const uint2 *>(sv.dna + wi + 2));
        W0 = ((uint64_t)a.x << 32) | a.y;
        W1 = ((uint64_t)b.x << 32) | b.y;
        mask = (KT)((k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1));
        const uint64_t f0 = k > 1 ? (W0 >> (64 - 2 * (k - 1))) : 0ull;  // first k-1 bases
        fw = (KT)f0;
        rc = (KT)(revcomp64(f0, k - 1) << 2);
        // first record boundary strictly after p0 (bounds is sorted ascending)
        uint32_t lo = 0, hi = nbounds;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&bounds[mid]) > p0) hi = mid; else lo = mid + 1;
        }
        bi = lo;
        nb = bi < nbounds ? __ldg(&bounds[bi]) : N;
    }
    // advance to the k-mer starting at p0 + i; returns false if it crosses a record boundary
    // or the end of the sequence
    __device__ __forceinline__ bool step(uint32_t i, KT &canon) {
        const uint32_t b = base(k - 1 + i);
        fw = (KT)(((fw << 2) | b) & mask);
        rc = (KT)((rc >> 2) | ((KT)(3u - b) << (2 * (k - 1))));
        canon = fw < rc ? fw : rc;
        const uint32_t pos = p0 + i;
        while (nb <= pos) {
            bi++;
            nb = bi < nbounds ? __ldg(&bounds[bi]) : N;
            if (bi >= nbounds) break;
        }
        return pos + k <= nb && pos + k <= N;
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

// ------------------------------------------------------------------ prob: setup
__global__ void k_prob_setup(const ProbJob *__restrict__ jobs, uint32_t njobs,
                             const FileResult *__restrict__ res, SketchConsts sc,
                             ProbBound *__restrict__ bound) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const uint32_t N = res[jobs[j].file].nsym;
    const uint32_t nk = N >= sc.k ? N - sc.k + 1 : 0;
    ProbBound b;
    if (nk == 0) {
        b.T = 0.0;
        b.uT = 0;
    } else {
        b.T = jobs[j].tmult * ((double)sc.m / (double)nk) * sc.lnm8;
        // x1 = c1*u >= u, so u < T is necessary for x1 < T (K3 re-tests exactly)
        b.uT = b.T >= 1.0 ? (1ull << 52) : (uint64_t)(b.T * 4503599627370496.0) + 2;
    }
    bound[j] = b;
}

// ------------------------------------------------------------------ prob: K2
template <class Src, typename KT>
__global__ void __launch_bounds__(kK2Threads)
k2_prob(const ProbJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix,
        uint32_t njobs, const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
        const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
        const uint32_t *__restrict__ boundaries, const ProbBound *__restrict__ bound,
        SketchConsts sc, uint32_t *__restrict__ overflow) {
    const uint32_t j = find_file(chunk_prefix, njobs, blockIdx.x);
    const ProbJob job = jobs[j];
    const FileResult fr = res[job.file];
    if (fr.status != 0) return;
    const uint32_t chunk = blockIdx.x - chunk_prefix[j];
    const uint32_t cbase = chunk * kChunk;
    if (cbase >= fr.nsym) return;  // grid is sized from the byte-length upper bound
    const FileDesc fd = files[job.file];
    SeqView sv;
    sv.dna = packed_dna ? packed_dna + fd.out_off : nullptr;
    sv.aa = packed_aa ? packed_aa + fd.out_off : nullptr;
    sv.bounds = boundaries ? boundaries + fr.bd_off : nullptr;
    sv.nbounds = boundaries ? fr.nrec : 0;
    sv.N = fr.nsym;
    const ProbBound pb = bound[j];
    const uint32_t p0 = cbase + threadIdx.x * kRun;
    Src src;
    src.init(sv, p0, sc.k);
    for (uint32_t i = 0; i < kRun; i++) {
        KT canon;
        const bool valid = src.step(i, canon);
        bool light = false, rep = false;
        uint32_t slot = 0;
        if (valid) {
            uint64_t s0;
            const uint64_t out1 = first_output(nohash_seed<KT>(canon, sc.spec_flags), s0);
            const uint64_t U = out1 >> 12;
            const bool cand = (U < pb.uT) | (U >= sc.u_slow);
            const uint32_t entry = ((uint32_t)(s0 >> 56) << 24) | (p0 + i + 1);
            slot = __umulhi((uint32_t)s0, job.cap);
            for (;;) {
                const uint32_t old = atomicCAS(&job.table[slot], 0u, entry);
                if (old == 0u) {  // first occurrence of this k-mer
                    light = cand;
                    break;
                }
                if ((old >> 24) == (entry >> 24)) {
                    const uint32_t pos2 = (old & 0xFFFFFFu) - 1;
                    if (Src::kmer_at(sv, pos2, sc.k) == canon) {  // exact: a repeat
                        rep = atomicAdd(&job.cnt[slot], 1u) == 0u;
                        break;
                    }
                }
                slot = slot + 1 == job.cap ? 0 : slot + 1;
            }
        }
        const uint32_t idx = warp_append(light | rep, job.list_n);
        if (light | rep) {
            if (idx < job.list_cap) {
                ListEntry e;
                e.kmer = (uint64_t)canon;
                e.slot = slot;
                e.kind = rep ? 1u : 0u;
                job.list[idx] = e;
            } else {
                atomicOr(&overflow[j], 1u);
            }
        }
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
        const uint32_t extra = job.cnt[le.slot];
        if (le.kind == 0 && extra != 0) continue;  // handled through its "repeated" entry
        const KT d = (KT)le.kmer;
        pmh_points<KT>(d, 1u + extra, T, sc, [&](double h, uint32_t k) {
            const unsigned long long hb = (unsigned long long)__double_as_longlong(h);
            if (PASS == 0) {
                atomicMin(&job.hmin[k], hb);
            } else {
                if (job.hmin[k] == hb) atomicMin(&job.sigw[k], (unsigned long long)d);
            }
        });
    }
}

// per genome: check the bound, convert to the output element type, reset touched counters
template <typename SigT>
__global__ void __launch_bounds__(256)
k3_prob_finalize(const ProbJob *__restrict__ jobs, uint32_t njobs,
                 const ProbBound *__restrict__ bound, const FileResult *__restrict__ res,
                 SketchConsts sc, SigT *__restrict__ sig_out, uint64_t *__restrict__ nb_bases_out,
                 uint32_t *__restrict__ retry) {
    const uint32_t j = blockIdx.x;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    __shared__ unsigned long long smax[256];
    const unsigned long long Tb = (unsigned long long)__double_as_longlong(bound[j].T);
    unsigned long long mx = 0;
    for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
        const unsigned long long hb = job.hmin[k];
        mx = hb > mx ? hb : mx;
        const unsigned long long s = job.sigw[k];
        sig_out[(size_t)job.file * sc.m + k] = (s == ~0ull) ? (SigT)0 : (SigT)s;
    }
    smax[threadIdx.x] = mx;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (threadIdx.x < d && smax[threadIdx.x + d] > smax[threadIdx.x])
            smax[threadIdx.x] = smax[threadIdx.x + d];
        __syncthreads();
    }
    const uint32_t N = res[job.file].nsym;
    const bool has_kmers = N >= sc.k && res[job.file].status == 0;
    if (threadIdx.x == 0) {
        // exact iff every slot's minimum is strictly below the bound (no skipped point can
        // win or tie); an empty genome is trivially exact
        uint32_t r = (has_kmers && !(smax[0] < Tb)) ? 1u : 0u;
        retry[job.file] = r | (res[job.file].status << 8);
        if (nb_bases_out) nb_bases_out[job.file] = res[job.file].nbases;
    }
    // the extra-occurrence counters touched by this genome are cleared by the slot's next
    // k_prob_reset (full grid) from the list left here
    if (threadIdx.x == 0) {
        const uint32_t n = *job.list_n;
        *job.prev_n = n > job.list_cap ? job.list_cap : n;
    }
}

// ------------------------------------------------------------------ optdens
struct DensJob {
    uint32_t file;
    uint32_t *bins;  // [m] f32 bit patterns, atomicMin
    double tmult;
};

constexpr uint32_t kDensLargeBits = 0x4F800000u;  // 4294967296.0f = F::from(u32::MAX)

template <class Src, typename KT>
__global__ void __launch_bounds__(kK2Threads)
k2_optdens(const DensJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix,
           uint32_t njobs, const FileDesc *__restrict__ files,
           const FileResult *__restrict__ res, const uint32_t *__restrict__ packed_dna,
           const uint8_t *__restrict__ packed_aa, const uint32_t *__restrict__ boundaries,
           SketchConsts sc) {
    const uint32_t j = find_file(chunk_prefix, njobs, blockIdx.x);
    const DensJob job = jobs[j];
    const FileResult fr = res[job.file];
    if (fr.status != 0) return;
    const uint32_t chunk = blockIdx.x - chunk_prefix[j];
    const uint32_t cbase = chunk * kChunk;
    if (cbase >= fr.nsym) return;
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

// per genome: check the bound, densify empty bins (cold path), write f32 signature
__global__ void __launch_bounds__(256)
k3_optdens_finalize(const DensJob *__restrict__ jobs, uint32_t njobs,
                    const FileResult *__restrict__ res, SketchConsts sc,
                    float *__restrict__ sig_out, uint64_t *__restrict__ nb_bases_out,
                    uint32_t *__restrict__ retry) {
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
    if (threadIdx.x == 0) {
        retry[job.file] = (need_retry ? 1u : 0u) | (fr.status << 8);
        if (nb_bases_out) nb_bases_out[job.file] = fr.nbases;
    }
    float *out = sig_out + (size_t)job.file * sc.m;
    const uint32_t nempty = snempty;
    for (uint32_t k = threadIdx.x; k < sc.m; k += blockDim.x) {
        uint32_t b = job.bins[k];
        if (b == kDensLargeBits && nempty != sc.m && !need_retry) {
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

}  // namespace gsb
