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
//       multiplicities with no global atomic per k-mer.  Only keys that are repeated or light
//       leave the CTA, as (k-mer, weight) candidates for the unchanged exact replay (k3_prob_*).
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
constexpr uint32_t kCRound = 4096;          // keys per counting round (load <= 1/2)
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
template <typename KEY>
constexpr size_t count_smem_bytes() {
    return (size_t)kCRound * sizeof(KEY) + (size_t)kCTab * sizeof(KEY) + kCTab * 2 + 512 * 8 * 2;
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
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < sc.m; k += gridDim.x * blockDim.x) {
        job.hmin[k] = 0x7FEFFFFFFFFFFFFFull;  // f64::MAX
        job.sigw[k] = ~0ull;
    }
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

// ---- pass C: count.  grid = (kNB, njobs), kCThreads threads
//
// The bucket's run is staged in shared memory by one TMA bulk copy per chunk of kCRound keys
// (cp.async.bulk + mbarrier: SASS UBLKCP), then every lane drains its share of the stage through a
// per-lane queue: one probe (atomicCAS) per loop trip for whichever key the lane holds, so lanes
// that need a second probe do not idle the others.  A key becomes a candidate exactly once -- at its
// first occurrence if it is light, at its second occurrence otherwise -- and only its table slot is
// remembered, in a private segment of the thread (no atomic, no vote); after the round the segments
// are compacted by a block-wide prefix sum and all threads turn the slots into (k-mer, weight)
// entries that leave with one global atomicAdd per CTA.
constexpr int kCThreads = 512;
constexpr uint32_t kCSeg = 8;  // candidate slots a thread can remember per round

template <typename KT, typename KEY>
__global__ void __launch_bounds__(kCThreads)
k2p_count(const ProbJob *__restrict__ jobs, uint32_t njobs, const FileResult *__restrict__ res, SketchConsts sc,
          PartConsts pc, uint32_t *__restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    KEY *s_stage = reinterpret_cast<KEY *>(s_raw);                        // [kCRound] keys of the current chunk
    KEY *s_key = s_stage + kCRound;                                       // [kCTab]
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_key + kCTab);        // [kCTab / 2] u16 pairs: extra occurrences
    uint16_t *s_seg = reinterpret_cast<uint16_t *>(s_cnt + kCTab / 2);    // [kCThreads][kCSeg] candidate slots
    uint16_t *s_dense = reinterpret_cast<uint16_t *>(s_stage);            // compacted slots (the stage is free by then)
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_wsum[kCThreads / 32];
    __shared__ uint32_t s_base, s_special;
    const uint32_t j = blockIdx.y, b = blockIdx.x;
    if (j >= njobs) return;
    const ProbJob job = jobs[j];
    if (res[job.file].status != 0) return;
    uint32_t n = job.cursor[b];
    if (n == 0) return;
    if (n > job.cap_g) n = job.cap_g;  // the genome is flagged already (k2p_partition)
    const KEY *run = reinterpret_cast<const KEY *>(job.buckets) + (size_t)b * job.cap_g;
    const uint32_t rn = (n + kCRound - 1) / kCRound;          // counting rounds (1 for ordinary genomes)
    const uint32_t nchunk = rn;                               // stage loads per round
    // table sized to the round: 2 x keys rounded up to a power of two, at least 256 slots
    uint32_t lg = 8;
    {
        const uint32_t per = rn > 1 ? kCRound : n;
        while ((1u << lg) < 2 * per && (1u << lg) < kCTab) lg++;
    }
    const uint32_t ts = 1u << lg, tmask = ts - 1;
    const KEY keymask = pc.keybits >= 8 * sizeof(KEY) ? (KEY)~(KEY)0 : (KEY)(((KEY)1 << pc.keybits) - 1);
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
    }
    uint32_t phase = 0;
    for (uint32_t r = 0; r < rn; r++) {
        __syncthreads();  // previous round's flush is done (and the barrier is initialised)
        {
            const uint4 e4 = make_uint4(~0u, ~0u, ~0u, ~0u), z4 = make_uint4(0, 0, 0, 0);
            uint4 *k4 = reinterpret_cast<uint4 *>(s_key);
            for (uint32_t s = threadIdx.x; s < ts * sizeof(KEY) / 16; s += kCThreads) k4[s] = e4;
            uint4 *c4 = reinterpret_cast<uint4 *>(s_cnt);
            for (uint32_t s = threadIdx.x; s < ts / 8; s += kCThreads) c4[s] = z4;
        }
        if (threadIdx.x == 0) s_special = 0;
        bool full = false;
        uint32_t nloc = 0;  // candidates remembered by this thread
        uint16_t *seg = s_seg + threadIdx.x * kCSeg;
        for (uint32_t c = 0; c < nchunk; c++) {
            const uint32_t c0 = c * kCRound, cn = n - c0 < kCRound ? n - c0 : kCRound;
            __syncthreads();  // table cleared / previous chunk drained: the stage may be overwritten
            if (threadIdx.x == 0) {
                const uint32_t bytes = (uint32_t)((cn * sizeof(KEY) + 15) & ~(size_t)15);  // within cap_g (multiple of 4 keys)
                mbar_expect_tx(&s_bar, bytes);
                tma_bulk_g2s(s_stage, run + c0, bytes, &s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
            // ---- per-lane queue: hold one key, probe once per trip
            uint32_t i = threadIdx.x, s = 0, probes = 0;
            KEY w = 0;
            bool have = false;
            for (;;) {
                if (!have) {
                    if (i >= cn) break;
                    w = s_stage[i];
                    i += kCThreads;
                    const uint32_t hw = sizeof(KEY) == 8 ? (uint32_t)(((uint64_t)w * 0x9E3779B97F4A7C15ULL) >> 32)
                                                         : (uint32_t)w * 0x9E3779B1u;
                    if (rn > 1 && ((hw >> 4) & 0xFFFFu) % rn != r) continue;  // another round's key
                    if (w == KeyTraits<KEY>::kEmpty) {  // the one word that collides with the sentinel
                        atomicAdd(&s_special, 1u);
                        continue;
                    }
                    s = hw >> (32 - lg);
                    probes = 0;
                }
                const KEY old = smem_cas(&s_key[s], KeyTraits<KEY>::kEmpty, w);
                const bool isnew = old == KeyTraits<KEY>::kEmpty, isdup = old == w;
                const bool flagged = (w & KeyTraits<KEY>::kFlag) != 0;
                bool emit = isnew && flagged;  // light: a candidate from its first occurrence on
                if (isdup) {  // one more occurrence: the first of them makes a heavy key a candidate
                    const uint32_t shv = (s & 1u) * 16u;
                    const uint32_t before = (atomicAdd(&s_cnt[s >> 1], 1u << shv) >> shv) & 0xFFFFu;
                    emit = before == 0 && !flagged;
                }
                if (emit) {
                    if (nloc < kCSeg) seg[nloc] = (uint16_t)s;
                    else full = true;  // more candidates than a thread remembers: general path
                    nloc++;
                }
                have = !(isnew || isdup);
                s = (s + 1) & tmask;
                if (have && ++probes > ts) {  // table full: more distinct keys than a round holds
                    full = true;
                    have = false;
                }
            }
        }
        if (full) atomicOr(&overflow[j], 2u);
        if (nloc > kCSeg) nloc = kCSeg;
        // ---- block-wide exclusive prefix of nloc -> dense list of slots in the (now free) stage
        uint32_t incl = nloc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        __syncthreads();  // every thread is done with the stage (and with the table)
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t pre = 0, nl = 0;
#pragma unroll
        for (int w2 = 0; w2 < kCThreads / 32; w2++) {
            const uint32_t v = s_wsum[w2];
            if ((uint32_t)w2 < warp) pre += v;
            nl += v;
        }
        pre += incl - nloc;
        for (uint32_t t = 0; t < nloc; t++) s_dense[pre + t] = seg[t];
        const uint32_t nspecial = s_special;
        const uint32_t nout = nl + (nspecial ? 1u : 0u);
        if (threadIdx.x == 0) s_base = nout ? atomicAdd(job.list_n, nout) : 0u;
        __syncthreads();
        // ---- the remembered slots leave as (k-mer, weight) candidates
        const uint32_t gbase = s_base;
        for (uint32_t t = threadIdx.x; t < nout; t += kCThreads) {
            KEY w;
            uint32_t extra;
            if (t < nl) {
                const uint32_t s = s_dense[t];
                w = s_key[s];
                extra = (s_cnt[s >> 1] >> ((s & 1u) * 16u)) & 0xFFFFu;
            } else {
                w = KeyTraits<KEY>::kEmpty;  // light by construction (all bits set)
                extra = nspecial - 1;
            }
            const KT u = (KT)(((KT)b << pc.keybits) | (KT)(w & keymask));
            ListEntry e;
            e.kmer = (uint64_t)bunmix<KT>(u, pc);
            e.slot = 1u + extra;  // kind 2: the weight itself
            e.kind = 2;
            if (gbase + t < job.list_cap) job.list[gbase + t] = e;
            else atomicOr(&overflow[j], 1u);
        }
    }
}

}  // namespace gsb
