// datagen.cu -- seeded synthetic FASTA for the benchmark and the tests (host code, no CUDA).
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

#include "../../include/gsearch_b200.h"

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
