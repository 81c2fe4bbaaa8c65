// synth.cpp -- seeded synthetic workloads for the benchmark and the tests.  Host code, no CUDA,
// NOT part of the product: it builds into its own datagen/libgsb_synth.so so that the CPU
// reference arm of bench.py maps no product code.
//
// The reference ships no sample data; the workloads of BASELINE.json are
// synthesised as SURVEY.md 8(d) specifies: genomes come in families of 16 around a random root
// with a per-member substitution rate, 2 % of the length as duplicated segments (so that some
// k-mers have multiplicity > 1), 0.1 % of bases turned into 'N' runs and 1 % lower case, 80-column
// lines.  Proteomes: ~333-residue proteins over the 20-letter alphabet with '*' terminators.
// Generator: xoshiro256++ seeded through SplitMix64 from 0x5EED0000 + index.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define GSB_API __attribute__((visibility("default")))

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t sm(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
        for (auto &v : s) v = sm(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

const double kRates[6] = {0.001, 0.005, 0.01, 0.02, 0.05, 0.10};

void random_bases(Rng &r, std::vector<uint8_t> &seq, size_t L) {
    seq.resize(L);
    size_t i = 0;
    while (i < L) {
        uint64_t v = r.next();
        for (int j = 0; j < 32 && i < L; j++, v >>= 2) seq[i++] = "ACGT"[v & 3];
    }
}

// positions of i.i.d. events with probability p, by geometric skipping
template <class F>
void for_each_event(Rng &r, size_t L, double p, F f) {
    if (p <= 0) return;
    const double lq = log1p(-p);
    double pos = -1;
    for (;;) {
        double u = r.unit();
        if (u <= 0) u = 1e-300;
        pos += 1.0 + floor(log(u) / lq);
        if (pos >= (double)L) return;
        f((size_t)pos);
    }
}

}  // namespace

// Upper bound of the FASTA size of one synthetic genome / proteome
extern "C" GSB_API uint64_t gsb_synth_max_bytes(uint64_t length, uint32_t nrecords) {
    return length + length / 80 + 64ull * (nrecords + 1) + 256;
}

// Synthetic genome `index` of `length` bases split into `ncontigs` records.  Returns the number
// of bytes written (0 if cap is too small).
extern "C" GSB_API uint64_t gsb_synth_dna_genome(uint64_t index, uint64_t length, uint32_t ncontigs,
                                                 uint8_t *out, uint64_t cap) {
    if (ncontigs < 1) ncontigs = 1;
    if (cap < gsb_synth_max_bytes(length, ncontigs)) return 0;
    const uint64_t family = index / 16, member = index % 16;
    std::vector<uint8_t> seq;
    Rng root(0x5EED0000ull + 16 * family + 0xF00D0000ull);
    random_bases(root, seq, length);
    Rng r(0x5EED0000ull + index);
    if (member != 0) {
        const double p = kRates[member % 6];
        for_each_event(r, length, p, [&](size_t i) {
            const char *alt = "ACGT";
            uint8_t c;
            do c = alt[r.next() & 3]; while (c == seq[i]);
            seq[i] = c;
        });
    }
    // 2 % duplicated segments: 20 copies of length/1000 bases
    const size_t seg = std::max<size_t>(1, length / 1000);
    if (length > 4 * seg)
        for (int d = 0; d < 20; d++) {
            const size_t src = r.below(length - seg), dst = r.below(length - seg);
            memmove(&seq[dst], &seq[src], seg);
        }
    // 0.1 % of bases as 'N' runs of 10..100
    size_t n_budget = length / 1000;
    while (n_budget > 0 && length > 200) {
        const size_t run = std::min<size_t>(n_budget, 10 + r.below(91));
        const size_t at = r.below(length - run);
        memset(&seq[at], 'N', run);
        n_budget -= run;
    }
    // 1 % lower case in runs of 50
    size_t lc_budget = length / 100;
    while (lc_budget > 0 && length > 200) {
        const size_t run = std::min<size_t>(lc_budget, 50);
        const size_t at = r.below(length - run);
        for (size_t i = at; i < at + run; i++) seq[i] |= 0x20;
        lc_budget -= run;
    }
    uint64_t o = 0;
    const uint64_t per = (length + ncontigs - 1) / ncontigs;
    for (uint32_t c = 0; c < ncontigs; c++) {
        const uint64_t b = (uint64_t)c * per, e = std::min<uint64_t>(length, b + per);
        if (b >= e && c > 0) break;
        o += (uint64_t)snprintf((char *)out + o, 64, ">syn_%07llu.%u len=%llu\n", (unsigned long long)index, c,
                                (unsigned long long)(e - b));
        for (uint64_t i = b; i < e; i += 80) {
            const uint64_t n = std::min<uint64_t>(80, e - i);
            memcpy(out + o, &seq[i], n);
            o += n;
            out[o++] = '\n';
        }
    }
    return o;
}

// Synthetic proteome `index`: `nprot` proteins of ~`mean_len` residues.
extern "C" GSB_API uint64_t gsb_synth_aa_proteome(uint64_t index, uint32_t nprot, uint32_t mean_len,
                                                  uint8_t *out, uint64_t cap) {
    const uint64_t total = (uint64_t)nprot * (mean_len + mean_len / 2 + 2);
    if (cap < gsb_synth_max_bytes(total, nprot)) return 0;
    static const char aa[] = "ACDEFGHIKLMNPQRSTVWY";
    const uint64_t family = index / 16, member = index % 16;
    Rng root(0x5EED0000ull + 16 * family + 0xAA000000ull);
    Rng r(0x5EED0000ull + index + 0x0A0A0000ull);
    const double p = member ? kRates[member % 6] * 2 : 0.0;
    uint64_t o = 0;
    std::vector<uint8_t> prot;
    for (uint32_t q = 0; q < nprot; q++) {
        const uint32_t L = mean_len / 2 + (uint32_t)root.below(mean_len + 1);
        prot.resize(L);
        for (uint32_t i = 0; i < L; i++) prot[i] = aa[root.below(20)];
        for_each_event(r, L, p, [&](size_t i) { prot[i] = aa[r.below(20)]; });
        o += (uint64_t)snprintf((char *)out + o, 64, ">prot_%07llu_%05u hypothetical protein\n",
                                (unsigned long long)index, q);
        for (uint32_t i = 0; i < L; i += 60) {
            const uint32_t n = std::min<uint32_t>(60, L - i);
            memcpy(out + o, &prot[i], n);
            o += n;
            if (i + 60 >= L) out[o++] = '*';
            out[o++] = '\n';
        }
    }
    return o;
}

// ------------------------------------------------------------------------------------------
// Synthetic signature database for the `request` workload (BASELINE configs[2]; SURVEY 8d:
// "sketches may be generated directly as synthetic signatures with planted family structure").
// A random recursive tree: point 0 is random; point i > 0 keeps each slot of a random earlier
// point par(i) with probability keep(i) in [0.50, 0.95] and draws the other slots fresh, so that
// distances to ancestors and cousins are graded (1 - product of the keeps along the path).
// Every value is a stateless hash of (seed, i, slot): the same database comes out for any
// thread count, and the CPU reference arm and the GPU arm of bench.py see identical bytes.
//   elem_bytes: 8 (u64 values below 2^40), 4 (u32) or 2 (u16)
namespace {
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
inline uint64_t cell(uint64_t seed, uint64_t i, uint64_t s, uint64_t salt) {
    return mix64(mix64(seed + 0x9e3779b97f4a7c15ULL * (i + 1)) ^ (0xD1B54A32D192ED03ULL * (s + 1)) ^ salt);
}
template <class T>
void tree_rows(T *out, uint64_t n, uint32_t S, uint64_t seed, uint64_t vmask, uint64_t first) {
    std::vector<uint64_t> par(n);
    std::vector<uint64_t> keep(n);  // threshold on a 32-bit draw
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t g = first + i;
        par[i] = g ? (uint64_t)(((__uint128_t)cell(seed, g, 0, 0x50415245ull) * g) >> 64) : 0;
        const double k = 0.50 + 0.45 * ((double)(cell(seed, g, 0, 0x4B454550ull) >> 11) * (1.0 / 9007199254740992.0));
        keep[i] = (uint64_t)(k * 4294967296.0);
    }
    // a block of columns is independent of the others: rows in order inside a block
    const uint32_t kCols = 512;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c0 = 0; c0 < (int64_t)S; c0 += kCols) {
        const uint32_t c1 = std::min<uint32_t>(S, (uint32_t)c0 + kCols);
        for (uint64_t i = 0; i < n; i++) {
            const uint64_t g = first + i;
            T *row = out + i * (uint64_t)S;
            const T *prow = (g && par[i] >= first) ? out + (par[i] - first) * (uint64_t)S : nullptr;
            const uint64_t rowh = mix64(seed + 0x9e3779b97f4a7c15ULL * (g + 1));
            for (uint32_t s = (uint32_t)c0; s < c1; s++) {
                const uint64_t h = mix64(rowh ^ (0xD1B54A32D192ED03ULL * (s + 1)));  // == cell(seed, g, s, 0)
                const bool kept = g != 0 && prow != nullptr && (h >> 32) < keep[i];
                row[s] = kept ? prow[s] : (T)(1 + ((h * 0x9E3779B97F4A7C15ULL >> 20) & vmask));
            }
        }
    }
}
}  // namespace

// rows [first, first + n) of the database; with first > 0 a parent outside the range is treated
// as absent (the row is fresh), so shards are self-contained.  Returns 0, or 1 on a bad argument.
extern "C" GSB_API int gsb_synth_signatures(void *out, uint64_t n, uint32_t S, uint32_t elem_bytes, uint64_t seed,
                                            uint64_t first) {
    if (!out || !S) return 1;
    switch (elem_bytes) {
    case 8: tree_rows<uint64_t>((uint64_t *)out, n, S, seed, (1ull << 40) - 2, first); return 0;
    case 4: tree_rows<uint32_t>((uint32_t *)out, n, S, seed, 0xFFFFFFFDull, first); return 0;
    case 2: tree_rows<uint16_t>((uint16_t *)out, n, S, seed, 0xFFFDull, first); return 0;
    default: return 1;
    }
}

// queries: query j copies row pick(j) of `db` (n rows) and redraws each slot with probability
// `noise` -- a mutated member of the database, as a `request` genome is of its family
extern "C" GSB_API int gsb_synth_queries(void *out, uint32_t nq, const void *db, uint64_t n, uint32_t S,
                                         uint32_t elem_bytes, uint64_t seed, double noise, uint64_t *picked) {
    if (!out || !db || !n || !S || (elem_bytes != 8 && elem_bytes != 4 && elem_bytes != 2)) return 1;
    const uint64_t thr = (uint64_t)(noise * 4294967296.0);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)nq; j++) {
        const uint64_t p = (uint64_t)(((__uint128_t)cell(seed, (uint64_t)j, 0, 0x5049434Bull) * n) >> 64);
        if (picked) picked[j] = p;
        for (uint32_t s = 0; s < S; s++) {
            const uint64_t h = cell(seed ^ 0x51554552ull, (uint64_t)j, s, 0);
            const bool fresh = (h >> 32) < thr;
            const uint64_t v = 1 + ((h * 0x9E3779B97F4A7C15ULL >> 20));
            if (elem_bytes == 8)
                ((uint64_t *)out)[(uint64_t)j * S + s] = fresh ? (v & ((1ull << 40) - 2)) + (1ull << 40) : ((const uint64_t *)db)[p * S + s];
            else if (elem_bytes == 4)
                ((uint32_t *)out)[(uint64_t)j * S + s] = fresh ? (uint32_t)(v | 1u) : ((const uint32_t *)db)[p * S + s];
            else
                ((uint16_t *)out)[(uint64_t)j * S + s] = fresh ? (uint16_t)(v | 1u) : ((const uint16_t *)db)[p * S + s];
        }
    }
    return 0;
}
