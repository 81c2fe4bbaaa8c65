// prob_partition.cuh -- K2 of the ProbMinHash3a path, round 2: exact k-mer multiplicities by
// hash PARTITION (streamed, CTA-staged writes) + per-bucket counting in SHARED memory.
//
// Replaces the L2-atomic filter of sketch_kernels.cuh (k2_prob_mark / classify / overflow, kept as
// the general fallback) on the hot path of
//   kmerutils ProbHash3aSketch::sketch_compressedkmer[_seqs] [U]   (src/dna/dnasketch.rs:336,357)
// i.e. the multiplicity map `IndexMap<Kmer::Val, f64>` that ProbMinHash3a::hashset consumes.
//
// Why: the filter needs one returning global atomic and one random L2 load per k-mer and sits at
// the L2 request ceiling (127 G atomics/s, profiles/r1_ubench_atomics*.log).  Shared-memory
// atomics on random addresses run at 2 170 G/s chip-wide (ATOMS.ADD returning, 7.5 per clock per
// SM; profiles/r2_ubench_smem.log) -- 17x more -- so the counting moves on chip:
//
//   k2p_partition  one pass over the genome.  Per tile of 8192 k-mer positions a CTA rolls the
//       canonical k-mers in registers, evaluates the first ProbMinHash draw (the "light" test
//       against the static bound T, two SplitMix64 mixes) and a BIJECTIVE mix u of the k-mer value;
//       the top 11 bits of u name one of 2048 buckets, the rest IS the key (31 bits for k = 21, so
//       a key + the light flag is one 32-bit word).  Keys are ranked inside the tile with one
//       shared-memory atomicAdd per k-mer, staged per bucket in shared memory and appended to the
//       bucket's array in global memory with ONE global atomicAdd per (tile, bucket) -- 0.25 global
//       atomics per k-mer instead of 1, and coalesced run writes instead of random L2 traffic.
//   k2p_count      one CTA per (genome, bucket): the ~2 400 keys of a bucket go through a
//       shared-memory hash table (atomicCAS on the key, atomicAdd on a duplicate counter): exact
//       multiplicities with no global atomic per k-mer.  Only keys that are repeated or light are
//       candidates; the CTA replays their exact f64 point sequence on the spot and lowers the
//       genome's 128-bit slot objects (h, k-mer) with compare-and-swap (k3p_finalize128 reads them).
//
// Exactness: u -> (bucket, key) is a bijection of the 2k-bit (5k-bit) k-mer value, so equal keys in
// a bucket are equal k-mers; every occurrence lands in exactly one bucket run (shared-memory region,
// per-tile spill list, or the genome is flagged for the general path); the table counts every
// occurrence once.  Genomes that do not fit the geometry (a bucket array or a table round
// overflowing: very large or very repetitive inputs) are flagged and re-run through the filter path.
#pragma once

#include "sketch_kernels.cuh"

namespace gsb {

constexpr int kPB = 11;                     // log2(buckets per genome)
constexpr uint32_t kNB = 1u << kPB;         // buckets per genome
constexpr uint32_t kRegion = 7;             // staged keys per (tile, bucket); tile mean is 4 (7: three CTAs per SM)
constexpr uint32_t kSpillCap = 1024;        // per-tile spill list (keys beyond a full region)
constexpr uint32_t kCTab = 8192;            // slots of the counting table
constexpr uint32_t kCRound = 3584;          // keys per counting round / stage chunk (14 per thread; 3 CTAs per SM)
constexpr uint32_t kCapGMax = 60000;        // duplicate counters are 16 bits wide
constexpr uint64_t kBMixC1 = 0x9E3779B97F4A7C15ULL, kBMixC2 = 0xD6E8FEB86659FD93ULL;
constexpr uint64_t kBMixC1Inv = 0xF1DE83E19937733DULL, kBMixC2Inv = 0xCFEE444D8B59A89BULL;

struct PartConsts {
    uint64_t mask;      // 2^kbits - 1
    uint32_t kbits;     // 2k (DNA) or 5k (AA)
    uint32_t sh;        // ceil(kbits / 2): x ^= x >> sh is its own inverse
    uint32_t bbits;     // bucket bits = min(kPB, kbits)
    uint32_t keybits;   // kbits - bbits
};

__host__ __device__ constexpr PartConsts make_part_consts(uint32_t kbits) {
    PartConsts pc{};
    pc.kbits = kbits;
    pc.mask = kbits >= 64 ? ~0ull : ((1ull << kbits) - 1);
    pc.sh = (kbits + 1) / 2;
    pc.bbits = kbits < (uint32_t)kPB ? kbits : (uint32_t)kPB;
    pc.keybits = kbits - pc.bbits;
    return pc;
}

// bijection of [0, 2^kbits): odd multiply, xorshift, odd multiply (all mod 2^kbits)
template <typename KT>
__device__ __forceinline__ KT bmix(KT x, const PartConsts &pc) {
    const KT m = (KT)pc.mask;
    x = (KT)(x * (KT)kBMixC1) & m;
    x ^= x >> pc.sh;
    x = (KT)(x * (KT)kBMixC2) & m;
    return x;
}
template <typename KT>
__device__ __forceinline__ KT bunmix(KT x, const PartConsts &pc) {
    const KT m = (KT)pc.mask;
    x = (KT)(x * (KT)kBMixC2Inv) & m;
    x ^= x >> pc.sh;
    x = (KT)(x * (KT)kBMixC1Inv) & m;
    return x;
}

template <typename KEY>
struct KeyTraits {
    static constexpr KEY kFlag = (KEY)1 << (8 * sizeof(KEY) - 1);   // "first draw below the bound"
    static constexpr KEY kEmpty = (KEY)~(KEY)0;                     // table sentinel
};

__device__ __forceinline__ uint32_t smem_cas(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ uint64_t smem_cas(uint64_t *p, uint64_t cmp, uint64_t v) {
    return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long *>(p), (unsigned long long)cmp,
                               (unsigned long long)v);
}

template <typename KEY>
struct SpillEntry {
    KEY key;
    uint32_t br;  // bucket << 16 | rank inside the tile
};

template <typename KEY>
constexpr size_t part_smem_bytes() {
    return (size_t)kNB * kRegion * sizeof(KEY) + kNB * 4 + kSpillCap * sizeof(SpillEntry<KEY>);
}

// per group: reset cursors / slots / candidate cursor, compute the bounds (new path)
__global__ void __launch_bounds__(256)
k2p_reset(const ProbJob *__restrict__ jobs, uint32_t njobs, const FileResult *__restrict__ res, SketchConsts sc,
          ProbBound *__restrict__ bound, uint32_t *__restrict__ overflow, uint32_t *__restrict__ retry) {
    const uint32_t j = blockIdx.y;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    // the slot's previous genome may have gone through the filter path: clear the extra-occurrence
    // counters it left (k3_prob_finalize of THIS job then zeroes prev_n)
    {
        const uint32_t np = *job.prev_n;
        for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < np; e += gridDim.x * blockDim.x) {
            const ListEntry le = job.list[e];
            if (le.kind == 1) job.cnt[le.slot] = 0;
        }
    }
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < kNB; b += gridDim.x * blockDim.x) job.cursor[b] = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < sc.m; k += gridDim.x * blockDim.x)
        job.slot2[k] = make_ulonglong2(0x7FEFFFFFFFFFFFFFull /* f64::MAX */, ~0ull);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *job.list_n = 0;
        overflow[j] = 0;
        retry[job.file] = 0;
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

// ---- pass P: partition.  KBITS > 0 fixes the k-mer width (2k or 5k bits) at compile time: every
// shift and mask of the bijective mix and of the key split is then an immediate
template <class Src, typename KT, typename KEY, int KBITS>
__global__ void __launch_bounds__(kK2Threads, 3)
k2p_partition(const ProbJob *__restrict__ jobs, const uint32_t *__restrict__ chunk_prefix, uint32_t njobs,
              const FileDesc *__restrict__ files, const FileResult *__restrict__ res,
              const uint32_t *__restrict__ packed_dna, const uint8_t *__restrict__ packed_aa,
              const uint32_t *__restrict__ boundaries, const ProbBound *__restrict__ bound, SketchConsts sc,
              PartConsts pc_rt, uint32_t *__restrict__ overflow, uint32_t nchunks) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    KEY *s_tab = reinterpret_cast<KEY *>(s_raw);                                    // [kNB][kRegion]
    // [kNB] keys of this tile per bucket; after the reservation: count | base in the bucket array << 14
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_tab + (size_t)kNB * kRegion);
    SpillEntry<KEY> *s_spill = reinterpret_cast<SpillEntry<KEY> *>(s_cnt + kNB);    // [kSpillCap]
    __shared__ uint32_t s_nspill;
    const PartConsts pc = KBITS ? make_part_consts(KBITS) : pc_rt;
    const KEY keymask = pc.keybits >= 8 * sizeof(KEY) ? (KEY)~(KEY)0 : (KEY)(((KEY)1 << pc.keybits) - 1);
    // light <=> U < uT or U >= u_slow, U = out >> 12  <=>  (out - A) >= (B - A) with A = uT << 12,
    // B = u_slow << 12 (unsigned wrap-around; uT <= u_slow always)
    const uint64_t lightB = sc.u_slow << 12;
    for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {  // persistent CTAs
        const uint32_t j = find_file(chunk_prefix, njobs, chunk);
        const ProbJob job = jobs[j];
        const ChunkCtx cx = chunk_ctx(chunk_prefix, njobs, job.file, files, res, packed_dna, packed_aa, boundaries, j, chunk);
        if (!cx.live) continue;  // uniform for the CTA
        const ProbBound pb = bound[j];
        const uint64_t uT = pb.uT < sc.u_slow ? pb.uT : sc.u_slow;
        const uint64_t lightA = uT << 12, lightSpan = lightB - lightA;
        __syncthreads();  // the previous tile's copy-out has finished reading the stage
        for (uint32_t i = threadIdx.x; i < kNB; i += kK2Threads) s_cnt[i] = 0;
        if (threadIdx.x == 0) s_nspill = 0;
        __syncthreads();
        Src src;
        src.init(cx.sv, cx.p0, sc.k);
#pragma unroll 1
        for (uint32_t blk = 0; blk < kRun / kG; blk++) {
            uint32_t bk[kG];
            KEY key[kG];
            uint32_t act = 0;
            const bool fast = __all_sync(0xffffffffu, src.all_valid(blk * kG, kG));
#pragma unroll
            for (int g = 0; g < kG; g++) {
                KT kmer;
                const bool valid = fast ? src.step_fast(blk * kG + g, kmer) : src.step(blk * kG + g, kmer);
                uint64_t s0;
                const uint64_t out = first_output(nohash_seed<KT>(kmer, sc.spec_flags), s0);
                const bool light = (out - lightA) >= lightSpan;
                const KT u = bmix<KT>(kmer, pc);
                bk[g] = (uint32_t)(u >> pc.keybits);
                key[g] = ((KEY)u & keymask) | (light ? KeyTraits<KEY>::kFlag : (KEY)0);
                if (valid) act |= 1u << g;
            }
            uint32_t rk[kG];
            if (act == (1u << kG) - 1) {  // the common case: no branch around the atomics
#pragma unroll
                for (int g = 0; g < kG; g++) rk[g] = atomicAdd(&s_cnt[bk[g]], 1u);
            } else {
#pragma unroll
                for (int g = 0; g < kG; g++) rk[g] = (act & (1u << g)) ? atomicAdd(&s_cnt[bk[g]], 1u) : 0u;
            }
#pragma unroll
            for (int g = 0; g < kG; g++)
                if (act & (1u << g)) {
                    if (rk[g] < kRegion) {
                        s_tab[bk[g] * kRegion + rk[g]] = key[g];
                    } else {
                        const uint32_t q = atomicAdd(&s_nspill, 1u);
                        if (q < kSpillCap) {
                            s_spill[q].key = key[g];
                            s_spill[q].br = (bk[g] << 16) | rk[g];
                        }
                    }
                }
        }
        __syncthreads();
        // ---- reserve room in the bucket arrays: one global atomic per non-empty (tile, bucket)
        KEY *gb = reinterpret_cast<KEY *>(job.buckets);
        const uint32_t cap_g = job.cap_g;
        bool bad = false;
#pragma unroll 8
        for (uint32_t b = threadIdx.x; b < kNB; b += kK2Threads) {
            const uint32_t c = s_cnt[b];
            if (c) {
                const uint32_t base = atomicAdd(&job.cursor[b], c);
                // a run that does not fit is dropped whole (count 0): the genome goes to the general path
                const bool fits = base + c <= cap_g;
                bad |= !fits;
                s_cnt[b] = fits ? (c | (base << 14)) : 0u;  // cap_g <= kCapGMax < 2^16, c <= 8192
            }
        }
        const uint32_t nspill = s_nspill;
        if (bad || (threadIdx.x == 0 && nspill > kSpillCap)) atomicOr(&overflow[j], 2u);  // -> general path
        __syncthreads();
        // ---- copy-out: a warp moves 4 regions (8 lanes each) per step; runs are contiguous
        {
            const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
            const uint32_t s = lane & 7u;
            uint32_t b = warp * 4 + (lane >> 3);
            KEY *dst = gb + (size_t)b * cap_g + s;
            const size_t dstep = (size_t)(kK2Threads / 8) * cap_g;
            const KEY *src = s_tab + b * kRegion + s;
#pragma unroll 8
            for (uint32_t it = 0; it < kNB / (kK2Threads / 8); it++) {
                const uint32_t cw = s_cnt[b];
                const uint32_t c = cw & 0x3FFFu;
                if (s < c && s < kRegion) dst[cw >> 14] = *src;
                b += kK2Threads / 8;
                dst += dstep;
                src += (kK2Threads / 8) * kRegion;
            }
        }
        const uint32_t ns = nspill < kSpillCap ? nspill : kSpillCap;
        for (uint32_t q = threadIdx.x; q < ns; q += kK2Threads) {
            const SpillEntry<KEY> sp = s_spill[q];
            const uint32_t b = sp.br >> 16, cw = s_cnt[b];
            if (cw & 0x3FFFu) gb[(size_t)b * cap_g + (cw >> 14) + (sp.br & 0xFFFFu)] = sp.key;
        }
    }
}

// ---- pass C: count + slot update.  grid = (kNB, njobs), kCThreads threads
//
// The bucket's run is staged in shared memory by one TMA bulk copy per chunk of kCRound keys
// (cp.async.bulk + mbarrier: SASS UBLKCP) while the table is cleared.  Phase A is straight-line:
// every lane takes its keys four at a time, ONE atomicCAS each; a key that found its slot empty is
// counted, anything else (a second occurrence or a different key in the slot: ~16 %) is parked in a
// pool of the WARP (ballot compaction, no atomic).  Phase B drains the pool, 32 parked keys per trip
// (probe on / bump the occurrence counter).  A key becomes a candidate exactly once -- at its first
// occurrence if it is light, at its second occurrence otherwise -- and only its table slot is
// remembered, again per warp.  After ONE block barrier (the occurrence counters are final) every
// warp turns its candidates into (k-mer, weight), replays the exact f64 arithmetic of
// ProbMinHash3a::hashset against the static bound and lowers the genome's 128-bit slot objects
// (h, k-mer) with compare-and-swap: no candidate list, no second pass, no global cursor.  The two
// global round trips of a compare-and-swap are hidden by the other resident CTAs.
// Job parameters travel in the kernel arguments (constant bank): a CTA's first dependent global
// load is its bucket cursor.  (Tried and measured slower on B200, same 5 Mbp workload: persistent
// CTAs with a double-buffered stage that prefetch the next bucket, 300 us per group of 6 genomes
// against 271; per-lane bit masks instead of warp pools, 413 us; a separate slot-update kernel over a
// candidate list, 226 + 66 us.)
constexpr int kCThreads = 256;
constexpr uint32_t kCWarps = kCThreads / 32;
constexpr uint32_t kCPark = kCRound / kCWarps;   // parked keys per warp and chunk (all of them, at worst)
constexpr uint32_t kCCand = 256;                 // candidate slots per warp and round
constexpr uint32_t kMaxGroupJobs = 8;            // genomes per group (kMaxSlots / 2)

struct CountArgs {
    const void *buckets[kMaxGroupJobs];
    const uint32_t *cursor[kMaxGroupJobs];
    ulonglong2 *slot2[kMaxGroupJobs];    // [m] (ordered bits of min h, winning k-mer) per MinHash slot
    uint32_t cap_g[kMaxGroupJobs];
};

template <typename KEY>
constexpr size_t count_smem_bytes() {
    return (size_t)kCRound * sizeof(KEY) + (size_t)kCTab * sizeof(KEY) + kCTab * 2 + kCWarps * kCPark * 2 +
           kCWarps * kCCand * 2;
}

template <typename KT, typename KEY>
__global__ void __launch_bounds__(kCThreads, 3)
k2p_count(CountArgs args, uint32_t njobs, const ProbBound *__restrict__ bound, SketchConsts sc, PartConsts pc,
          uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    KEY *s_stage = reinterpret_cast<KEY *>(s_raw);                        // [kCRound] keys of the current chunk
    KEY *s_key = s_stage + kCRound;                                       // [kCTab]
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_key + kCTab);        // [kCTab / 2] u16 pairs: extra occurrences
    uint16_t *s_park = reinterpret_cast<uint16_t *>(s_cnt + kCTab / 2);   // [kCWarps][kCPark] parked stage indices
    uint16_t *s_cand = s_park + kCWarps * kCPark;                         // [kCWarps][kCCand] candidate slots
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_special;
    const uint32_t j = blockIdx.y, b = blockIdx.x;
    if (j >= njobs) return;
    uint32_t n = __ldg(args.cursor[j] + b);
    if (n == 0) return;
    const double T = bound[j].T;  // needed after the counting: the load flies meanwhile
    const uint32_t cap_g = args.cap_g[j];
    if (n > cap_g) n = cap_g;  // the genome is flagged already (k2p_partition)
    const KEY *run = reinterpret_cast<const KEY *>(args.buckets[j]) + (size_t)b * cap_g;
    ulonglong2 *slot2 = args.slot2[j];
    const uint32_t rn = (n + kCRound - 1) / kCRound;          // counting rounds (1 for ordinary genomes)
    const uint32_t nchunk = rn;                               // stage loads per round
    // table sized to the round: 2 x keys rounded up to a power of two, 256 .. kCTab slots
    uint32_t lg = 32 - __clz(2 * (rn > 1 ? kCRound : n) - 1);
    lg = lg < 8 ? 8 : (lg > 13 ? 13 : lg);
    static_assert(kCTab == 1u << 13, "table size");
    static_assert(kCRound % kCThreads == 0, "chunk size");
    const uint32_t ts = 1u << lg, tmask = ts - 1;
    const KEY keymask = pc.keybits >= 8 * sizeof(KEY) ? (KEY)~(KEY)0 : (KEY)(((KEY)1 << pc.keybits) - 1);
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint16_t *wpark = s_park + warp * kCPark, *wcand = s_cand + warp * kCCand;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
    }
    uint32_t phase = 0;
    bool full = false;
    for (uint32_t r = 0; r < rn; r++) {
        uint32_t ncand = 0;  // candidates remembered by this warp (uniform)
        for (uint32_t c = 0; c < nchunk; c++) {
            const uint32_t c0 = c * kCRound, cn = n - c0 < kCRound ? n - c0 : kCRound;
            __syncthreads();  // previous chunk drained / previous round flushed (and the barrier is initialised)
            if (threadIdx.x == 0) {  // the bulk copy of the chunk flies while the table is cleared
                const uint32_t bytes = (uint32_t)((cn * sizeof(KEY) + 15) & ~(size_t)15);  // within cap_g (multiple of 4 keys)
                mbar_expect_tx(&s_bar, bytes);
                tma_bulk_g2s(s_stage, run + c0, bytes, &s_bar);
            }
            if (c == 0) {
                const uint4 e4 = make_uint4(~0u, ~0u, ~0u, ~0u), z4 = make_uint4(0, 0, 0, 0);
                uint4 *k4 = reinterpret_cast<uint4 *>(s_key);
                for (uint32_t s = threadIdx.x; s < ts * sizeof(KEY) / 16; s += kCThreads) k4[s] = e4;
                uint4 *c4 = reinterpret_cast<uint4 *>(s_cnt);
                for (uint32_t s = threadIdx.x; s < ts / 8; s += kCThreads) c4[s] = z4;
                if (threadIdx.x == 0) s_special = 0;
            }
            // one thread polls the mbarrier; the others sleep in the hardware barrier (a spin loop in
            // every warp costs more issue slots than the counting itself)
            if (threadIdx.x == 0) mbar_wait(&s_bar, phase);
            phase ^= 1u;
            __syncthreads();
            // ---- phase A
            uint32_t npark = 0;  // uniform in the warp
            constexpr int kU = 4;
            for (uint32_t i0 = threadIdx.x; i0 - lane < cn; i0 += kU * kCThreads) {  // warp-uniform trip count
                KEY w[kU], old[kU];
                uint32_t sl[kU];
                bool on[kU];
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    const uint32_t i = i0 + u * kCThreads;
                    on[u] = i < cn;
                    w[u] = s_stage[on[u] ? i : 0];
                }
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    const uint32_t hw = sizeof(KEY) == 8 ? (uint32_t)(((uint64_t)w[u] * 0x9E3779B97F4A7C15ULL) >> 32)
                                                         : (uint32_t)w[u] * 0x9E3779B1u;
                    sl[u] = hw >> (32 - lg);
                    if (rn > 1) on[u] = on[u] && ((hw >> 4) & 0xFFFFu) % rn == r;  // another round's key
                }
#pragma unroll
                for (int u = 0; u < kU; u++)  // (a no-op for the one word that collides with the sentinel)
                    old[u] = on[u] ? smem_cas(&s_key[sl[u]], KeyTraits<KEY>::kEmpty, w[u]) : (KEY)0;
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    const bool special = w[u] == KeyTraits<KEY>::kEmpty;
                    const bool isnew = on[u] && old[u] == KeyTraits<KEY>::kEmpty && !special;
                    const bool park = on[u] && !isnew;
                    const bool cand = isnew && (w[u] & KeyTraits<KEY>::kFlag);  // light: a candidate from its first occurrence on
                    const uint32_t bp = __ballot_sync(0xffffffffu, park), bc = __ballot_sync(0xffffffffu, cand);
                    if (park) wpark[npark + __popc(bp & lt)] = (uint16_t)(i0 + u * kCThreads);
                    npark += __popc(bp);
                    const uint32_t at = ncand + __popc(bc & lt);
                    if (cand && at < kCCand) wcand[at] = (uint16_t)sl[u];
                    ncand += __popc(bc);
                }
            }
            // ---- phase B: the warp's parked keys probe on (or count one more occurrence), 32 per trip
            for (uint32_t t0 = 0; t0 < npark; t0 += 32) {
                const uint32_t t = t0 + lane;
                const bool act = t < npark;
                const KEY w = s_stage[act ? wpark[t] : 0];
                const bool flagged = (w & KeyTraits<KEY>::kFlag) != 0;
                const uint32_t hw = sizeof(KEY) == 8 ? (uint32_t)(((uint64_t)w * 0x9E3779B97F4A7C15ULL) >> 32)
                                                     : (uint32_t)w * 0x9E3779B1u;
                uint32_t s = hw >> (32 - lg);
                bool cand = false;
                if (act && w == KeyTraits<KEY>::kEmpty) {
                    atomicAdd(&s_special, 1u);
                } else if (act) {
                    for (uint32_t probes = 0;; probes++) {
                        const KEY old = smem_cas(&s_key[s], KeyTraits<KEY>::kEmpty, w);
                        if (old == w) {  // one more occurrence: the first of them makes a heavy key a candidate
                            const uint32_t shv = (s & 1u) * 16u;
                            const uint32_t before = (atomicAdd(&s_cnt[s >> 1], 1u << shv) >> shv) & 0xFFFFu;
                            cand = before == 0 && !flagged;
                            break;
                        }
                        if (old == KeyTraits<KEY>::kEmpty) {  // first occurrence after all
                            cand = flagged;
                            break;
                        }
                        if (probes > ts) {  // table full: more distinct keys than a round holds
                            full = true;
                            break;
                        }
                        s = (s + 1) & tmask;
                    }
                }
                const uint32_t bc = __ballot_sync(0xffffffffu, cand);
                const uint32_t at = ncand + __popc(bc & lt);
                if (cand && at < kCCand) wcand[at] = (uint16_t)s;
                ncand += __popc(bc);
            }
        }
        if (ncand > kCCand) {  // more candidates than a warp remembers: general path
            full = true;
            ncand = kCCand;
        }
        __syncthreads();  // the occurrence counters are final
        // ---- candidates -> (k-mer, weight) -> points -> 128-bit slot minimum
        const uint32_t nspecial = s_special;
        const uint32_t nmine = ncand + ((warp == 0 && nspecial) ? 1u : 0u);
        for (uint32_t t = lane; t < nmine; t += 32) {
            KEY w;
            uint32_t extra;
            if (t < ncand) {
                const uint32_t s = wcand[t];
                w = s_key[s];
                extra = (s_cnt[s >> 1] >> ((s & 1u) * 16u)) & 0xFFFFu;
            } else {
                w = KeyTraits<KEY>::kEmpty;  // the word that collides with the sentinel: light by construction
                extra = nspecial - 1;
            }
            const KT d = bunmix<KT>((KT)(((KT)b << pc.keybits) | (KT)(w & keymask)), pc);
            pmh_points<KT>(d, 1u + extra, T, sc, [&](double h, uint32_t k) {
                slot_min128(&slot2[k], (unsigned long long)__double_as_longlong(h), (unsigned long long)d);
            });
        }
    }
    if (full) atomicOr(&overflow[j], 2u);
}

// ---- finalize of the partition path.  (The slot update itself happens inside k2p_count: every
// candidate replays the exact f64 arithmetic of ProbMinHash3a::hashset against the static bound and
// lowers the 128-bit slot objects (h, k-mer) with a compare-and-swap loop: the lexicographic minimum
// is the reference's result with this repository's tie rule -- identical h: smaller k-mer -- in any
// order of arrival.)
template <typename SigT>
__global__ void __launch_bounds__(256)
k3p_finalize128(const ProbJob *__restrict__ jobs, uint32_t njobs, const ProbBound *__restrict__ bound,
                const FileResult *__restrict__ res, SketchConsts sc, SigT *__restrict__ sig_out,
                uint64_t *__restrict__ nb_bases_out, uint32_t *__restrict__ retry) {
    const uint32_t j = blockIdx.y;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    const unsigned long long Tb = (unsigned long long)__double_as_longlong(bound[j].T);
    const FileResult fr = res[job.file];
    const bool has_kmers = fr.nsym >= sc.k && fr.status == 0;
    bool over = false;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < sc.m; k += gridDim.x * blockDim.x) {
        // exact iff every slot's minimum is strictly below the bound (no skipped point can win or tie)
        const ulonglong2 s = job.slot2[k];
        over |= !(s.x < Tb);
        sig_out[(size_t)job.file * sc.m + k] = (s.y == ~0ull) ? (SigT)0 : (SigT)s.y;
    }
    if (__syncthreads_or(over && has_kmers) && threadIdx.x == 0) atomicOr(&retry[job.file], 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (fr.status) atomicOr(&retry[job.file], fr.status << 8);
        if (nb_bases_out) nb_bases_out[job.file] = fr.nbases;
        *job.prev_n = 0;  // no extra-occurrence counters on this path
    }
}

}  // namespace gsb
