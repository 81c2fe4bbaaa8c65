/*
 * rng.c -- random primitives of the reference path (test infrastructure, see gso.h).
 *
 * Follows (all [U] = upstream crate, not vendored):
 *   rand_xoshiro::Xoshiro256PlusPlus::seed_from_u64  -> SplitMix64 x4      (A.3)
 *   rand 0.8 Uniform<f64>/Uniform<f32>/Uniform<usize>                     (A.4)
 *   probminhash::exp01::ExpRestricted01 (Ertl's truncated exponential)    (A.6)
 * Reached from the reference at src/dna/dnasketch.rs:336,357 through
 * kmerutils::ProbHash3aSketch -> probminhash::ProbMinHash3a::hashset.
 * Pinned by the published vectors of SplitMix64 / xoshiro256++ (tests/test_oracle_kat.py).
 */
#include "gso.h"

#include <math.h>
#include <string.h>

uint64_t gso_splitmix64_next(uint64_t *state) {
    uint64_t z = (*state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

void gso_xoshiro_seed_from_u64(gso_xoshiro *r, uint64_t seed) {
    uint64_t st = seed;
    for (int i = 0; i < 4; i++) r->s[i] = gso_splitmix64_next(&st);
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

uint64_t gso_xoshiro_next_u64(gso_xoshiro *r) {
    uint64_t *s = r->s;
    const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

/* rand_xoshiro: next_u32 = (next_u64() >> 32) as u32 */
uint32_t gso_xoshiro_next_u32(gso_xoshiro *r) { return (uint32_t)(gso_xoshiro_next_u64(r) >> 32); }

/* rand 0.8 UniformFloat<f64>::sample with low=0, scale=1:
 * value1_2 = from_bits((u64 >> 12) | exponent(0)); value1_2 - 1.0                  */
double gso_uniform_f64(gso_xoshiro *r) {
    uint64_t bits = (gso_xoshiro_next_u64(r) >> 12) | 0x3FF0000000000000ULL;
    double d;
    memcpy(&d, &bits, 8);
    return d - 1.0;
}

/* rand 0.8 UniformFloat<f32>::sample: (next_u32 >> 9) into [1,2) then - 1.0 */
float gso_uniform_f32(gso_xoshiro *r) {
    uint32_t bits = (gso_xoshiro_next_u32(r) >> 9) | 0x3F800000u;
    float f;
    memcpy(&f, &bits, 4);
    return f - 1.0f;
}

/* rand 0.8 UniformInt<usize>::sample (64-bit target): widening multiply with a
 * rejection zone; range = m, z = (2^64 - m) % m.                                   */
uint64_t gso_uniform_usize(gso_xoshiro *r, uint64_t m) {
    const uint64_t ints_to_reject = (UINT64_MAX - m + 1) % m;
    const uint64_t zone = UINT64_MAX - ints_to_reject;
    for (;;) {
        uint64_t v = gso_xoshiro_next_u64(r);
        __uint128_t p = (__uint128_t)v * (__uint128_t)m;
        uint64_t lo = (uint64_t)p, hi = (uint64_t)(p >> 64);
        if (lo <= zone) return hi;
    }
}

/* ExpRestricted01::new(lambda) */
void gso_exp01_init(gso_exp01 *e, double lambda) {
    e->lambda = lambda;
    e->c1 = expm1(lambda) / lambda;
    e->c2 = log(2.0 / (1.0 + exp(-lambda))) / lambda;
    e->c3 = (1.0 - exp(-lambda)) / lambda;
}

/* expm1(z) for 0 <= z <= ln 2, as this repository FREEZES it: the degree-24 Taylor polynomial in
 * Horner form, every step one correctly rounded IEEE operation (no contraction).  The reference
 * calls the platform libm (Rust f64::exp_m1), whose last bit is not specified; the restatement and
 * the CUDA kernel (common.cuh expm1_spec) must agree with EACH OTHER bit for bit, so both evaluate
 * this.  It is within 2 ulp of libm on the whole range and only decides the last test of the
 * rejection branch below.                                                                    */
double gso_expm1_spec(double z) {
    double r = 1.0;
    for (int k = 24; k >= 2; k--) r = 1.0 + (z / (double)k) * r;
    return z * r;
}

/* ln(x) for normal x > 0, as this repository FREEZES it (SetSketch needs floor(1 - log_b x), and a
 * last-bit difference between the platform libm and the device flips that floor): argument
 * reduction x = 2^e * m with sqrt(1/2) < m <= sqrt(2), s = (m-1)/(m+1), the classic degree-14
 * odd series in s split into even / odd halves (fdlibm's layout and published coefficients), every
 * step one correctly rounded IEEE operation (no contraction).  Within 1 ulp of libm; the
 * restatement and the CUDA kernel (common.cuh ln_spec) agree bit for bit.                        */
double gso_ln_spec(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
    memcpy(&bits, &x, 8);
    int64_t e = (int64_t)((bits >> 52) & 0x7FF) - 1023;
    uint64_t mb = (bits & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;
    double m;
    memcpy(&m, &mb, 8);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)e;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/* ExpRestricted01::sample */
double gso_exp01_sample(const gso_exp01 *e, gso_xoshiro *r) {
    double x = e->c1 * gso_uniform_f64(r);
    if (x < 1.0) return x;
    for (;;) {
        x = gso_uniform_f64(r);
        if (x < e->c2) return x;
        double y = 0.5 * gso_uniform_f64(r);
        if (y > 1.0 - x) {
            x = 1.0 - x;
            y = 1.0 - y;
        }
        if (x <= e->c3 * (1.0 - y)) return x;
        if (e->c1 * y <= 1.0 - x) return x;
        if (y * e->c1 * e->lambda <= gso_expm1_spec(e->lambda * (1.0 - x))) return x;
    }
}
