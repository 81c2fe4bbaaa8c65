// fasta_pack.cuh -- K1: raw FASTA bytes in HBM -> encoded sequence (sm_100a).
//
// Replaces, on device, the reference's file tasks
//   src/dna/dnafiles.rs:43-107,115-195 (seq mode), :200-276,283-360 (block mode)
//   src/aa/aafiles.rs:33-99,107-229
// i.e. needletail record splitting [U], the "capsid" record filter (dnafiles.rs:67,145,248,329),
// Sequence::encode_and_add (drop every byte outside ACGT/acgt, 2 bit/base) and
// filter_out_non_aa (aafiles.rs:11-28).
//
// Layout produced (see DESIGN.md "data layout"):
//   DNA : 2 bit/base, 16 bases per uint32 word, FIRST base in the MOST significant bits
//         (so a k-mer value is a funnel-shift of two words), plus a sorted list of record
//         start positions (seq mode only; k-mers must not span records, dnasketch.rs:347-365)
//   AA  : 1 byte/residue, codes 1..20; in seq mode a 0 byte separates records.
//
// Three kernels, no spin-waits:
//   k1a_tile_summary : per 4 KiB tile, the tile's effect on the parser state as a function
//                      of the (unknown) incoming state, and its symbol count for each of the
//                      four possible incoming states;
//   k1b_resolve      : one CTA per file resolves states and prefix-sums counts (two block scans);
//   k1c_pack         : per tile again (bytes now come from L2), compacts and writes.
// FASTQ (first byte '@'; needletail reads four-line records only [U]) runs through the same three
// kernels with another 4-state machine: s = line number mod 4, '\n' is s -> s+1, a byte is emitted
// iff it is in the alphabet and s == 1 (the sequence line), a record starts at the file's first
// byte and at every newline that leads to s == 0.  Lines that must start with '@' / '+' are
// checked (status 6).  A FASTQ file whose text contains "capsid" is refused (status 9): the
// reference would drop such records, this path has no state bit left for it.
// FASTA parser state s = in_header | dropped<<1.  Every byte is a function {0..3}->{0..3}:
//   '>' at a line start : s -> 1          (new record: in header, not dropped)
//   '\n'                : s -> s & 2      (header ends)
//   'c' of "capsid"     : s -> s|2 if s&1 (record id contains "capsid": dropped)
// A byte is emitted iff it is in the alphabet and s == 0.
#pragma once

#include "common.cuh"

namespace gsb {

constexpr int kTile = 4096;       // bytes per tile
constexpr int kK1Threads = 256;   // 16 bytes per thread

struct FileDesc {
    uint64_t beg, end;    // byte range of the file in the batch buffer
    uint64_t out_off;     // DNA: uint32-word offset of its packed output; AA: byte offset
    uint32_t tile_first;  // index of its first tile summary
    uint32_t ntiles;      // tiles covering [beg, end) in absolute 4 KiB coordinates
};

struct FileResult {
    uint32_t nsym;      // symbols written (DNA: bases; AA: residues + separators)
    uint32_t nbases;    // encoded bases / residues (ItemDict.len of the reference)
    uint32_t nrec;      // record starts seen
    uint32_t bd_off;    // DNA seq mode: offset of its boundary list in the boundary pool
    uint32_t status;    // 0 ok, 5 = does not start with '>' (needletail InvalidStart)
    uint32_t pad_;
};

// ---- state-function algebra: 4 fields of 2 bits, field s = image of state s ----
constexpr uint32_t kFnIdent = 0xE4;
__device__ __forceinline__ uint32_t fn_apply(uint32_t f, uint32_t s) { return (f >> (2 * s)) & 3u; }
// apply f first, then g
__device__ __forceinline__ uint32_t fn_compose(uint32_t f, uint32_t g) {
    uint32_t out = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) out |= fn_apply(g, fn_apply(f, s)) << (2 * s);
    return out;
}
// byte lane s of the result = 1 iff f(s) == 0
__device__ __forceinline__ uint32_t fn_zero_lanes(uint32_t f) {
    const uint32_t z = ~(f | (f >> 1)) & 0x55u;
    return (z & 1u) | ((z >> 2) & 1u) << 8 | ((z >> 4) & 1u) << 16 | ((z >> 6) & 1u) << 24;
}

// byte lane s of the result = 1 iff f(s) == 1
__device__ __forceinline__ uint32_t fn_one_lanes(uint32_t f) {
    const uint32_t z = (f & ~(f >> 1)) & 0x55u;
    return (z & 1u) | ((z >> 2) & 1u) << 8 | ((z >> 4) & 1u) << 16 | ((z >> 6) & 1u) << 24;
}
// every image + 1 mod 4 (FASTQ: a newline)
__device__ __forceinline__ uint32_t fn_rot1(uint32_t f) {
    const uint32_t lo = f & 0x55u, hi = (f >> 1) & 0x55u;
    return (lo ^ 0x55u) | ((hi ^ lo) << 1);
}
// residues "ACDEFGHIKLMNPQRSTVWY" -> 1..20 without a table: bit (c - 'A') of the mask says valid,
// the rank of that bit is the code (the alphabet string is in alphabetical order)
constexpr uint32_t kAAMask = 0x16fbdfdu;
__device__ __forceinline__ uint32_t aa_code(uint32_t c) {  // 0 = not a residue
    const uint32_t d = c - 'A';
    if (d > 25u || !((kAAMask >> d) & 1u)) return 0u;
    return (uint32_t)__popc(kAAMask & ((1u << d) - 1u)) + 1u;
}

template <int DATA_T>
__device__ __forceinline__ int sym_code(uint32_t c, const uint8_t *aa_lut) {
    if (DATA_T == 0) {
        const uint32_t d = (c | 0x20u) - 0x61u;  // 'a'..'z' -> 0..25
        constexpr uint32_t ok = 1u | (1u << 2) | (1u << 6) | (1u << 19);  // a c g t
        if (d > 25u || !((ok >> d) & 1u)) return -1;
        const uint32_t x = (c >> 1) & 3u;  // A0 C1 T2 G3
        return (int)(x ^ (x >> 1));        // A0 C1 G2 T3
    } else {
        (void)aa_lut;
        const int v = (int)aa_code(c);
        return v ? v : -1;
    }
}

__device__ __forceinline__ bool is_capsid(const uint8_t *p) {
    return p[0] == 'c' && p[1] == 'a' && p[2] == 'p' && p[3] == 's' && p[4] == 'i' && p[5] == 'd';
}

// stage one tile (+16 B halo on both sides) in shared memory, masking bytes outside the
// file to '\n' (so that the file's first byte is a line start and nothing leaks across files)
__device__ __forceinline__ void load_tile(const uint8_t *__restrict__ bytes, uint64_t total,
                                          uint64_t tbase, uint64_t fbeg, uint64_t fend,
                                          uint8_t *sm /* kTile + 32 */) {
    const int t = threadIdx.x;
    for (int blk = t; blk < kTile / 16 + 2; blk += kK1Threads) {
        const int64_t off = (int64_t)tbase - 16 + (int64_t)blk * 16;
        uint4 v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
        if (off >= 0 && (uint64_t)off + 16 <= total) {
            v = __ldcs(reinterpret_cast<const uint4 *>(bytes + off));  // streaming: do not displace the filters in L2
        } else if (off + 16 > 0 && (uint64_t)(off < 0 ? 0 : off) < total) {
            uint8_t tmp[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int64_t a = off + i;
                tmp[i] = (a >= 0 && (uint64_t)a < total) ? bytes[a] : (uint8_t)'\n';
            }
            v = *reinterpret_cast<uint4 *>(tmp);
        }
        *reinterpret_cast<uint4 *>(sm + blk * 16) = v;
    }
    __syncthreads();
    // mask bytes outside [fbeg, fend): only the first / last tile of a file needs it
    const int64_t lo = (int64_t)tbase - 16, hi = (int64_t)tbase + kTile + 16;
    if ((int64_t)fbeg > lo || (int64_t)fend < hi) {
        for (int i = t; i < kTile + 32; i += kK1Threads) {
            const int64_t a = lo + i;
            if (a < (int64_t)fbeg || a >= (int64_t)fend) sm[i] = (uint8_t)'\n';
        }
        __syncthreads();
    }
}

// ---- SIMD-in-register view of a thread's 16 bytes (the common case: no record start, no
// "capsid" text).  Bit i of a 16-bit mask = byte i.
__device__ __forceinline__ uint32_t bytes_eq(uint32_t w, uint32_t pat4) {  // 0x80 per equal byte
    const uint32_t t = w ^ pat4;
    return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
}
__device__ __forceinline__ uint32_t nib_of(uint32_t m80) {  // 0x80-per-byte mask -> 4 bits
    return (((m80 >> 7) * 0x00204081u) >> 21) & 0xFu;
}

struct Chunk16 {
    uint32_t sym;    // bytes that are alphabet symbols
    uint32_t nl;     // bytes that are '\n'
    uint32_t codes[4];  // DNA: 2-bit code of every byte, one per byte lane
};

// returns false if the chunk needs the byte-serial state machine ('>' or a "capsid" match)
template <int DATA_T>
__device__ __forceinline__ bool chunk16_scan(const uint8_t *p, const uint8_t *aa_lut, Chunk16 &c) {
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t gt = 0, lc = 0;
    c.sym = 0;
    c.nl = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        gt |= bytes_eq(w[j], 0x3e3e3e3eu);
        c.nl |= nib_of(bytes_eq(w[j], 0x0a0a0a0au)) << (4 * j);
        const uint32_t cj = bytes_eq(w[j], 0x63636363u);  // lower-case 'c'
        lc |= nib_of(cj) << (4 * j);
        if (DATA_T == 0) {
            const uint32_t y = w[j] | 0x20202020u;
            const uint32_t m = bytes_eq(y, 0x61616161u) | bytes_eq(y, 0x63636363u) |
                               bytes_eq(y, 0x67676767u) | bytes_eq(y, 0x74747474u);
            c.sym |= nib_of(m) << (4 * j);
            const uint32_t x = (w[j] >> 1) & 0x03030303u;  // A0 C1 T2 G3
            c.codes[j] = x ^ (x >> 1);                     // A0 C1 G2 T3
        }
    }
    if (gt) return false;
    while (lc) {
        const int i = __ffs(lc) - 1;
        lc &= lc - 1;
        if (is_capsid(p + i)) return false;
    }
    if (DATA_T == 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t d = ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu) - 'A';
            c.sym |= ((d <= 25u) ? ((kAAMask >> d) & 1u) : 0u) << i;
        }
    }
    return true;
}

// per-thread fold of its 16 bytes: state function, per-incoming-state counts, record starts
template <int DATA_T, bool SEQ_SEP>
__device__ __forceinline__ void fold16(const uint8_t *p /* own 16 bytes; p[-1], p[16..21] valid */,
                                       const uint8_t *aa_lut, uint32_t &f, uint32_t &cnt4,
                                       uint32_t &nrec, Chunk16 &ck, bool &fast) {
    fast = chunk16_scan<DATA_T>(p, aa_lut, ck);
    if (fast) {
        // no record start and no drop: before the first newline the incoming state decides,
        // after it the header flag is cleared (s -> s & 2)
        nrec = 0;
        if (ck.nl == 0) {
            f = kFnIdent;
            cnt4 = __popc(ck.sym);
        } else {
            const uint32_t first = __ffs(ck.nl) - 1;
            const uint32_t before = __popc(ck.sym & ((1u << first) - 1u));
            const uint32_t after = __popc(ck.sym >> (first + 1));
            f = kFnIdent & 0xAAu;
            cnt4 = (before + after) | (after << 8);
        }
        return;
    }
    f = kFnIdent;
    cnt4 = 0;
    nrec = 0;
    uint32_t inc4 = fn_zero_lanes(f);
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
        const uint32_t c = p[i];
        if (c == '>' && p[i - 1] == '\n') {
            f = 0x55u;
            inc4 = 0;
            nrec++;
            if (DATA_T == 1 && SEQ_SEP) cnt4 += 0x01010101u;  // separator symbol
        } else if (c == '\n') {
            f &= 0xAAu;
            inc4 = fn_zero_lanes(f);
        } else {
            // a "capsid" match only changes states that are inside a header; in a sequence
            // line the 'c' is an ordinary symbol, so fall through to the alphabet test
            if (c == 'c' && is_capsid(p + i)) {
                f |= (f & 0x55u) << 1;
                inc4 = fn_zero_lanes(f);
            }
            if (sym_code<DATA_T>(c, aa_lut) >= 0) cnt4 += inc4;
        }
    }
}

// FASTQ twin of fold16: bytes [lo, hi) of the chunk are inside the file.  `nl` = newlines seen (the
// tile-level record count follows from it and the incoming state), `capsid` = the text occurs.
template <int DATA_T, bool SEQ_SEP>
__device__ __forceinline__ void fold16_fastq(const uint8_t *p, const uint8_t *aa_lut, int lo, int hi, bool file_start,
                                             uint32_t &f, uint32_t &cnt4, uint32_t &nl, bool &capsid) {
    f = kFnIdent;
    cnt4 = 0;
    nl = 0;
    capsid = false;
    uint32_t inc4 = fn_one_lanes(f);
    if (DATA_T == 1 && SEQ_SEP && file_start) cnt4 += 1u;  // separator of the first record (incoming state 0)
#pragma unroll 1
    for (int i = lo; i < hi; i++) {
        const uint32_t c = p[i];
        if (c == '\n') {
            f = fn_rot1(f);
            nl++;
            inc4 = fn_one_lanes(f);
            if (DATA_T == 1 && SEQ_SEP) cnt4 += fn_zero_lanes(f);  // a record starts: separator symbol
        } else {
            if (c == 'c' && is_capsid(p + i)) capsid = true;
            if (sym_code<DATA_T>(c, aa_lut) >= 0) cnt4 += inc4;
        }
    }
}

// block-wide exclusive scan of state functions (composition) -> returns F_excl for this
// thread and the block total in *total
__device__ __forceinline__ uint32_t block_scan_fn(uint32_t f, uint32_t *warp_tot /* 8 */,
                                                  uint32_t *total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    // The common tile has no record start and no "capsid" text: every thread's function is then the
    // identity (no newline in its 16 bytes) or A: s -> s & 2 (a newline ends the header).  A is
    // idempotent and absorbs the identity, so the exclusive prefix is "A iff an earlier thread saw
    // a newline": one ballot per warp instead of a scan of function compositions.
    constexpr uint32_t kFnNl = kFnIdent & 0xAAu;
    if (__syncthreads_and(f == kFnIdent || f == kFnNl)) {
        const uint32_t bal = __ballot_sync(0xffffffffu, f == kFnNl);
        if (lane == 0) warp_tot[warp] = bal;
        __syncthreads();
        uint32_t before = bal & ((1u << lane) - 1u), any = 0;
#pragma unroll
        for (int w = 0; w < kK1Threads / 32; w++) {
            const uint32_t v = warp_tot[w];
            if ((uint32_t)w < warp) before |= v;
            any |= v;
        }
        if (total) *total = any ? kFnNl : kFnIdent;
        __syncthreads();
        return before ? kFnNl : kFnIdent;
    }
    uint32_t incl = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl = fn_compose(up, incl);
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t pre = kFnIdent;
    for (uint32_t w = 0; w < warp; w++) pre = fn_compose(pre, warp_tot[w]);
    uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = kFnIdent;
    excl = fn_compose(pre, excl);
    if (total) {
        uint32_t tot = kFnIdent;
        for (int w = 0; w < kK1Threads / 32; w++) tot = fn_compose(tot, warp_tot[w]);
        *total = tot;
    }
    __syncthreads();
    return excl;
}

// block-wide exclusive sum of a uint32 (two packed 16-bit counters are fine)
__device__ __forceinline__ uint32_t block_scan_add(uint32_t v, uint32_t *warp_tot /* 8 */,
                                                   uint32_t *total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
    for (int w = 0; w < kK1Threads / 32; w++) {
        if ((uint32_t)w < warp) pre += warp_tot[w];
        tot += warp_tot[w];
    }
    if (total) *total = tot;
    __syncthreads();
    return pre + incl - v;
}

// find the file a global tile index belongs to (tile_prefix has n+1 entries)
__device__ __forceinline__ uint32_t find_file(const uint32_t *__restrict__ tile_prefix, uint32_t n,
                                              uint32_t tile) {
    uint32_t lo = 0, hi = n;  // invariant: tile_prefix[lo] <= tile < tile_prefix[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&tile_prefix[mid]) <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void init_aa_lut(uint8_t *lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = 0;
    __syncthreads();
    if (threadIdx.x < 20) lut[(uint8_t)"ACDEFGHIKLMNPQRSTVWY"[threadIdx.x]] = (uint8_t)(threadIdx.x + 1);
    __syncthreads();
}

// ------------------------------------------------------------------ K1a
template <int DATA_T, bool SEQ_SEP>
__global__ void __launch_bounds__(kK1Threads)
k1a_tile_summary(const uint8_t *__restrict__ bytes, uint64_t total,
                 const FileDesc *__restrict__ files, const uint32_t *__restrict__ tile_prefix,
                 uint32_t nfiles, uint64_t *__restrict__ t_counts4, uint8_t *__restrict__ t_trans,
                 uint16_t *__restrict__ t_nrec, uint32_t tile0, uint32_t *__restrict__ fq_flags /* per file of the range */) {
    __shared__ __align__(16) uint8_t sm[kTile + 32];
    __shared__ uint8_t aa_lut[256];
    __shared__ uint32_t wtot[kK1Threads / 32];
    __shared__ unsigned long long wsum[kK1Threads / 32];
    if (DATA_T == 1) init_aa_lut(aa_lut);
    const uint32_t tile = blockIdx.x + tile0;  // files/tile_prefix/res point at the range's first file
    const uint32_t fi = find_file(tile_prefix, nfiles, tile);
    const FileDesc fd = files[fi];
    const uint32_t lt = tile - fd.tile_first;
    const uint64_t tbase = (fd.beg / kTile + lt) * (uint64_t)kTile;
    load_tile(bytes, total, tbase, fd.beg, fd.end, sm);

    const uint8_t *p = sm + 16 + threadIdx.x * 16;
    uint32_t f, cnt4, nrec;
    Chunk16 ck;
    bool fast;
    const bool fq = fd.end > fd.beg && bytes[fd.beg] == '@';  // uniform for the CTA
    if (fq) {
        const int64_t a0 = (int64_t)tbase + threadIdx.x * 16;
        const int lo = (int)max((int64_t)0, (int64_t)fd.beg - a0), hi = (int)min((int64_t)16, (int64_t)fd.end - a0);
        bool capsid;
        fold16_fastq<DATA_T, SEQ_SEP>(p, aa_lut, lo, hi, a0 <= (int64_t)fd.beg && (int64_t)fd.beg < a0 + 16, f, cnt4,
                                      nrec, capsid);
        if (__syncthreads_or(capsid) && threadIdx.x == 0) atomicOr(fq_flags + fi, 1u);
    } else {
        fold16<DATA_T, SEQ_SEP>(p, aa_lut, f, cnt4, nrec, ck, fast);
    }
    uint32_t ftot;
    const uint32_t fex = block_scan_fn(f, wtot, &ftot);
    // this thread's symbol count for each possible tile-incoming state s0
    unsigned long long c4 = 0;
#pragma unroll
    for (int s0 = 0; s0 < 4; s0++) {
        const uint32_t s = fn_apply(fex, s0);
        c4 |= (unsigned long long)((cnt4 >> (8 * s)) & 0xFFu) << (16 * s0);
    }
    // block reductions: 4x16-bit packed counts (<= 4096+256 each) and nrec
    unsigned long long v = c4;
    uint32_t r = nrec;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, d);
        r += __shfl_xor_sync(0xffffffffu, r, d);
    }
    if (lane_id() == 0) {
        wsum[threadIdx.x >> 5] = v;
        wtot[threadIdx.x >> 5] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tv = 0;
        uint32_t tr = 0;
        for (int w = 0; w < kK1Threads / 32; w++) {
            tv += wsum[w];
            tr += wtot[w];
        }
        t_counts4[tile] = tv;
        t_trans[tile] = (uint8_t)ftot;
        t_nrec[tile] = (uint16_t)tr;
    }
}

// ------------------------------------------------------------------ K1b
// one CTA per file: resolve the incoming parser state and the exclusive prefix of symbols / records
// of every tile.  A thread owns a run of consecutive tiles: it composes their state functions, the
// CTA scans the 256 compositions (so every thread knows the state entering its run), the thread
// walks its run again with the states known, and a second scan turns the per-thread totals into
// bases.  (Round 1 walked a file with ONE warp, 32 tiles per trip: 42 us for 5 MB files, as long as
// the tile-summary kernel itself.)
constexpr int kK1bThreads = 256;
__global__ void __launch_bounds__(kK1bThreads)
k1b_resolve(const FileDesc *__restrict__ files, uint32_t nfiles,
                            const uint8_t *__restrict__ bytes,
                            const uint64_t *__restrict__ t_counts4,
                            const uint8_t *__restrict__ t_trans,
                            const uint16_t *__restrict__ t_nrec, uint8_t *__restrict__ t_state,
                            uint32_t *__restrict__ t_base, uint32_t *__restrict__ t_recbase,
                            FileResult *__restrict__ res, uint32_t *__restrict__ bd_cursor,
                            uint32_t bd_capacity, int want_boundaries, int sep_counts_as_symbol,
                            const uint32_t *__restrict__ fq_flags) {
    __shared__ uint32_t s_fn[kK1bThreads];
    __shared__ uint32_t s_cnt[kK1bThreads], s_rec[kK1bThreads];
    const uint32_t fi = blockIdx.x;
    if (fi >= nfiles) return;
    const uint32_t t = threadIdx.x;
    const FileDesc fd = files[fi];
    const uint8_t first = fd.end > fd.beg ? bytes[fd.beg] : (uint8_t)'>';
    const bool fq = first == '@';  // FASTQ: t_nrec holds the tile's newline count
    const uint32_t per = (fd.ntiles + kK1bThreads - 1) / kK1bThreads;
    const uint32_t i0 = t * per, i1 = i0 + per < fd.ntiles ? i0 + per : fd.ntiles;
    // ---- composition of the run's state functions
    uint32_t f = kFnIdent;
    for (uint32_t i = i0; i < i1; i++) f = fn_compose(f, t_trans[fd.tile_first + i]);
    s_fn[t] = f;
    __syncthreads();
    for (int d = 1; d < kK1bThreads; d <<= 1) {  // inclusive scan under composition (earlier first)
        const uint32_t up = t >= (uint32_t)d ? s_fn[t - d] : kFnIdent;
        __syncthreads();
        if (t >= (uint32_t)d) s_fn[t] = fn_compose(up, s_fn[t]);
        __syncthreads();
    }
    uint32_t s = t ? fn_apply(s_fn[t - 1], 0u) : 0u;  // state entering the run (a file starts in state 0)
    // ---- the run with the states known: per-tile state, local prefix of symbols / records
    uint32_t cnt_tot = 0, rec_tot = 0;
    for (uint32_t i = i0; i < i1; i++) {
        const uint32_t T = fd.tile_first + i;
        const unsigned long long c4 = t_counts4[T];
        const uint32_t nr = t_nrec[T];
        const uint32_t cnt = (uint32_t)((c4 >> (16 * s)) & 0xFFFFull);
        // FASTQ: a record starts at the file's first byte and at every newline that leads to line 0
        const uint32_t nrr = fq ? ((s + nr) >> 2) + (i == 0 ? 1u : 0u) : nr;
        t_state[T] = (uint8_t)s;
        t_base[T] = cnt_tot;      // relative to the run; the base of the run is added below
        t_recbase[T] = rec_tot;
        cnt_tot += cnt;
        rec_tot += nrr;
        s = fn_apply(t_trans[T], s);
    }
    s_cnt[t] = cnt_tot;
    s_rec[t] = rec_tot;
    __syncthreads();
    for (int d = 1; d < kK1bThreads; d <<= 1) {
        const uint32_t uc = t >= (uint32_t)d ? s_cnt[t - d] : 0u, ur = t >= (uint32_t)d ? s_rec[t - d] : 0u;
        __syncthreads();
        s_cnt[t] += uc;
        s_rec[t] += ur;
        __syncthreads();
    }
    const uint32_t cbase = s_cnt[t] - cnt_tot, rbase = s_rec[t] - rec_tot;
    for (uint32_t i = i0; i < i1; i++) {
        const uint32_t T = fd.tile_first + i;
        t_base[T] += cbase;
        t_recbase[T] += rbase;
    }
    if (t == kK1bThreads - 1) {
        const uint32_t base = s_cnt[t], recbase = s_rec[t];
        FileResult r;
        r.nsym = base;
        r.nrec = recbase;
        r.nbases = sep_counts_as_symbol ? base - recbase : base;
        r.status = (first != '>' && first != '@') ? 5u : ((fq && fq_flags[fi]) ? 9u : 0u);
        r.bd_off = 0;
        r.pad_ = 0;
        if (want_boundaries) {
            r.bd_off = atomicAdd(bd_cursor, recbase);
            if (r.bd_off + recbase > bd_capacity) r.status = 8u;
        }
        res[fi] = r;
    }
}

// ------------------------------------------------------------------ K1c
template <int DATA_T, bool SEQ_SEP>
__global__ void __launch_bounds__(kK1Threads)
k1c_pack(const uint8_t *__restrict__ bytes, uint64_t total, const FileDesc *__restrict__ files,
         const uint32_t *__restrict__ tile_prefix, uint32_t nfiles,
         const uint8_t *__restrict__ t_state, const uint32_t *__restrict__ t_base,
         const uint32_t *__restrict__ t_recbase, FileResult *__restrict__ res,
         uint32_t *__restrict__ out_dna, uint8_t *__restrict__ out_aa,
         uint32_t *__restrict__ boundaries /* DNA seq mode, else null */, uint32_t tile0) {
    __shared__ __align__(16) uint8_t sm[kTile + 32];
    // AA: one byte per symbol, staged and copied.  DNA: the thread packs its (at most 16) codes into
    // one word, first base in the top bits, and ORs it into a word image of the tile's output at its
    // bit offset -- two shared-memory atomics per thread instead of 16 byte stores and, later, 16
    // byte loads per output word.
    __shared__ __align__(16) uint8_t stage[DATA_T == 1 ? kTile + 256 + 16 : 16];
    __shared__ uint32_t wimg[DATA_T == 0 ? kTile / 16 + 4 : 1];
    __shared__ uint8_t aa_lut[256];
    __shared__ uint32_t wtot[kK1Threads / 32];
    if (DATA_T == 1) init_aa_lut(aa_lut);
    if (DATA_T == 0)
        for (uint32_t w = threadIdx.x; w < kTile / 16 + 4; w += kK1Threads) wimg[w] = 0;  // (load_tile synchronises)
    const uint32_t tile = blockIdx.x + tile0;  // files/tile_prefix/res point at the range's first file
    const uint32_t fi = find_file(tile_prefix, nfiles, tile);
    const FileDesc fd = files[fi];
    const FileResult fr = res[fi];
    if (fr.status != 0) return;
    const uint32_t lt = tile - fd.tile_first;
    const uint64_t tbase = (fd.beg / kTile + lt) * (uint64_t)kTile;
    load_tile(bytes, total, tbase, fd.beg, fd.end, sm);

    const uint8_t *p = sm + 16 + threadIdx.x * 16;
    uint32_t f, cnt4, nrec;
    Chunk16 ck;
    bool fast = false;
    const bool fq = bytes[fd.beg] == '@';  // uniform for the CTA (the file is not empty: it has tiles)
    const int64_t a0 = (int64_t)tbase + threadIdx.x * 16;
    const int lo = (int)max((int64_t)0, (int64_t)fd.beg - a0), hi = (int)min((int64_t)16, (int64_t)fd.end - a0);
    const bool has_start = a0 <= (int64_t)fd.beg && (int64_t)fd.beg < a0 + 16;
    if (fq) {
        bool capsid;
        fold16_fastq<DATA_T, SEQ_SEP>(p, aa_lut, lo, hi, has_start, f, cnt4, nrec, capsid);
    } else {
        fold16<DATA_T, SEQ_SEP>(p, aa_lut, f, cnt4, nrec, ck, fast);
    }
    const uint32_t fex = block_scan_fn(f, wtot, nullptr);
    const uint32_t s0 = t_state[tile];
    uint32_t s = fn_apply(fex, s0);
    if (fq) nrec = ((s + nrec) >> 2) + (has_start ? 1u : 0u);  // record starts of this chunk, now that its state is known
    const uint32_t cnt = (cnt4 >> (8 * s)) & 0xFFu;
    uint32_t tile_tot;
    const uint32_t sc = block_scan_add(cnt | (nrec << 16), wtot, &tile_tot);
    uint32_t o = sc & 0xFFFFu;          // symbols emitted by earlier threads of this tile
    uint32_t ro = sc >> 16;             // record starts in earlier threads of this tile
    const uint32_t ntile = tile_tot & 0xFFFFu;
    const uint32_t pbase = t_base[tile];
    const uint32_t o_first = o;   // DNA: where this thread's codes start; `acc` collects them, top-aligned
    uint32_t acc = 0;
    auto emit_sym = [&](uint32_t code) {
        if (DATA_T == 0) acc |= code << (30u - 2u * (o - o_first));
        else stage[o] = (uint8_t)code;
        o++;
    };
    // concrete pass: emit symbols
    if (fq) {
        uint32_t bad = 0;
#pragma unroll 1
        for (int i = lo; i < hi; i++) {
            const uint32_t c = p[i];
            const int64_t a = a0 + i;
            bool start = a == (int64_t)fd.beg;  // (line state 0 by construction)
            if (c == '\n') {
                s = (s + 1) & 3u;
                const bool more = a + 1 < (int64_t)fd.end;
                const uint32_t nb = more ? p[i + 1] : 0u;
                if (s == 0) {
                    start = true;  // a trailing terminator opens an empty record: harmless (no symbol follows)
                    if (more && nb != '@' && nb != '\n' && nb != '\r') bad = 6;  // InvalidStart inside the file
                } else if (s == 2) {
                    // InvalidSeparator.  (Blank lines after the last record keep the line counter turning:
                    // they emit nothing, so a terminator or the end of the file is let through here; a
                    // record cut before its '+' line is then NOT diagnosed, unlike in the reference.)
                    if (more && nb != '+' && nb != '\n' && nb != '\r') bad = 6;
                }
            } else if (s == 1) {
                const int code = sym_code<DATA_T>(c, aa_lut);
                if (code >= 0) emit_sym((uint32_t)code);
            }
            if (start) {
                if (DATA_T == 1 && SEQ_SEP) emit_sym(0u);
                if (DATA_T == 0 && boundaries) {
                    boundaries[fr.bd_off + t_recbase[tile] + ro] = pbase + o;
                    ro++;
                }
            }
        }
        if (bad) atomicMax(&res[fi].status, bad);
    } else if (fast) {
        // emitted = symbols before the first newline if s == 0, after it if (s & 2) == 0
        uint32_t emit = ck.sym;
        if (ck.nl) {
            const uint32_t first = __ffs(ck.nl) - 1;
            const uint32_t lo = (1u << first) - 1u;
            emit = (s == 0 ? (ck.sym & lo) : 0u) | ((s & 2u) == 0 ? (ck.sym & ~lo) : 0u);
        } else if (s != 0) {
            emit = 0;
        }
        if (DATA_T == 0) {
            if (emit == 0xFFFFu) {
                // four codes per word sit in byte lanes (first byte lowest): one multiply gathers them
                // into a byte, first base highest ((x * 0x40100401) >> 24 for x = b0 + b1<<8 + b2<<16 + b3<<24)
#pragma unroll
                for (int q = 0; q < 4; q++) acc |= (((ck.codes[q] & 0x03030303u) * 0x40100401u) >> 24) << (24 - 8 * q);
                o += 16;
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    if ((emit >> i) & 1u) emit_sym((ck.codes[i >> 2] >> (8 * (i & 3))) & 3u);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++)
                if ((emit >> i) & 1u) emit_sym(aa_code(p[i]));
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < 16; i++) {
            const uint32_t c = p[i];
            if (c == '>' && p[i - 1] == '\n') {
                s = 1;
                if (DATA_T == 1 && SEQ_SEP) emit_sym(0u);
                if (DATA_T == 0 && boundaries) {
                    boundaries[fr.bd_off + t_recbase[tile] + ro] = pbase + o;
                    ro++;
                }
            } else if (c == '\n') {
                s &= 2u;
            } else {
                if (c == 'c' && (s & 1u) && is_capsid(p + i)) s |= 2u;
                if (s == 0) {
                    const int code = sym_code<DATA_T>(c, aa_lut);
                    if (code >= 0) emit_sym((uint32_t)code);
                }
            }
        }
    }
    if (DATA_T == 0 && o != o_first) {
        const uint32_t pos = (pbase & 15u) + o_first, w = pos >> 4, sh = pos & 15u;
        atomicOr(&wimg[w], acc >> (2u * sh));
        if (sh && (o - o_first) + sh > 16u) atomicOr(&wimg[w + 1], acc << (32u - 2u * sh));
    }
    __syncthreads();
    if (DATA_T == 1) {
        uint8_t *dst = out_aa + fd.out_off + pbase;
        for (uint32_t i = threadIdx.x; i < ntile; i += kK1Threads) dst[i] = stage[i];
    } else {
        // 16 bases per word, first base in the top bits; edge words are shared with the
        // neighbouring tiles and go through atomicOr (the output is pre-zeroed)
        if (ntile == 0) return;
        uint32_t *dst = out_dna + fd.out_off;
        const uint32_t w_first = pbase >> 4, w_last = (pbase + ntile - 1) >> 4;
        for (uint32_t w = w_first + threadIdx.x; w <= w_last; w += kK1Threads) {
            const uint32_t val = wimg[w - w_first];
            const bool full = w * 16 >= pbase && w * 16 + 16 <= pbase + ntile;
            if (full) dst[w] = val; else if (val) atomicOr(&dst[w], val);
        }
    }
}

}  // namespace gsb
