// common.cuh -- device-side primitives shared by the sm_100a kernels of libgsearch_b200.
//
// Arithmetic follows SURVEY.md Appendix A (the frozen SPEC of this repository):
//   A.3  SplitMix64 / xoshiro256++ (rand_xoshiro::Xoshiro256PlusPlus::seed_from_u64)
//   A.4  rand 0.8 Uniform<f64>, Uniform<f32>, Uniform<usize>
//   A.6  probminhash::exp01::ExpRestricted01
// All f64 steps use explicit round-to-nearest intrinsics so that no FMA contraction can
// change a bit relative to the reference's (Rust, never contracted) arithmetic.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gsb {

constexpr uint64_t kGolden = 0x9e3779b97f4a7c15ULL;
constexpr uint64_t kFxSeed64 = 0x517cc1b727220a95ULL;  // fxhash::FxHasher64, one word

__device__ __forceinline__ uint64_t sm64_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

struct Xoshiro {
    uint64_t s0, s1, s2, s3;
    // seed_from_u64: four successive SplitMix64 outputs
    __device__ __forceinline__ void seed(uint64_t seed) {
        s0 = sm64_mix(seed + kGolden);
        s1 = sm64_mix(seed + 2 * kGolden);
        s2 = sm64_mix(seed + 3 * kGolden);
        s3 = sm64_mix(seed + 4 * kGolden);
    }
    __device__ __forceinline__ uint64_t next() {
        const uint64_t result = rotl64(s0 + s3, 23) + s0;
        const uint64_t t = s1 << 17;
        s2 ^= s0;
        s3 ^= s1;
        s1 ^= s2;
        s0 ^= s3;
        s2 ^= t;
        s3 = rotl64(s3, 45);
        return result;
    }
};

// The first output of a freshly seeded generator only needs s0 and s3: two SplitMix64
// mixes instead of four.  This is what the per-k-mer filter evaluates.
__device__ __forceinline__ uint64_t first_output(uint64_t seed, uint64_t &s0_out) {
    const uint64_t s0 = sm64_mix(seed + kGolden);
    const uint64_t s3 = sm64_mix(seed + 4 * kGolden);
    s0_out = s0;
    return rotl64(s0 + s3, 23) + s0;
}

// Uniform::<f64>::new(0.,1.).sample : 52 random mantissa bits in [1,2) minus 1
__device__ __forceinline__ double u01_f64_from_bits(uint64_t r) {
    return __dadd_rn(__longlong_as_double((long long)((r >> 12) | 0x3FF0000000000000ULL)), -1.0);
}
// Uniform::<f32>::new(0.,1.).sample : next_u32 = high half of next_u64, 23 mantissa bits
__device__ __forceinline__ float u01_f32_from_bits(uint64_t r) {
    return __fadd_rn(__uint_as_float((((uint32_t)(r >> 32)) >> 9) | 0x3F800000u), -1.0f);
}

// Uniform::<usize>::new(0, m).sample : widening multiply + rejection zone
__device__ __forceinline__ uint32_t uniform_usize(Xoshiro &rng, uint64_t m, uint64_t zone) {
    for (;;) {
        const uint64_t v = rng.next();
        const uint64_t lo = v * m;
        if (lo <= zone) return (uint32_t)__umul64hi(v, m);
    }
}

struct Exp01 {
    double lambda, c1, c2, c3;  // computed on the host with libm, exactly as the oracle does
};

// expm1 on [0, ln 2] as the SPEC freezes it (oracle/rng.c gso_expm1_spec): degree-24 Taylor polynomial,
// Horner form, one correctly rounded operation per step -- identical bits on the host and here,
// which libm's expm1 (1-ulp differences between implementations) cannot promise.
__device__ __forceinline__ double expm1_spec(double z) {
    double r = 1.0;
#pragma unroll 1
    for (int k = 24; k >= 2; k--) r = __dadd_rn(1.0, __dmul_rn(__ddiv_rn(z, (double)k), r));
    return __dmul_rn(z, r);
}

// ln(x), normal x > 0, as the SPEC freezes it (oracle/rng.c gso_ln_spec): same reduction, same
// coefficients, same order of correctly rounded operations -- identical bits on the host and here
// (SetSketch takes floor(1 - log_b x): a last-bit difference between two libms flips it).
__device__ __forceinline__ double ln_spec(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7FFull) - 1023;
    double m = __longlong_as_double((long long)((bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    if (m > 1.4142135623730951) {
        m = __dmul_rn(m, 0.5);
        e += 1;
    }
    const double f = __dadd_rn(m, -1.0);
    const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
    const double z = __dmul_rn(s, s);
    const double w = __dmul_rn(z, z);
    const double t1 = __dmul_rn(w, __dadd_rn(Lg2, __dmul_rn(w, __dadd_rn(Lg4, __dmul_rn(w, Lg6)))));
    const double t2 = __dmul_rn(z, __dadd_rn(Lg1, __dmul_rn(w, __dadd_rn(Lg3, __dmul_rn(w, __dadd_rn(Lg5, __dmul_rn(w, Lg7)))))));
    const double R = __dadd_rn(t2, t1);
    const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
    const double dk = (double)e;
    // dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f)
    const double a = __dadd_rn(__dmul_rn(s, __dadd_rn(hfsq, R)), __dmul_rn(dk, ln2_lo));
    return __dadd_rn(__dmul_rn(dk, ln2_hi), -__dadd_rn(__dadd_rn(hfsq, -a), -f));
}

// ExpRestricted01::sample, entered after the first uniform has been drawn (u0)
__device__ __forceinline__ double exp01_sample_from(const Exp01 &e, double u0, Xoshiro &rng) {
    double x = __dmul_rn(e.c1, u0);
    if (x < 1.0) return x;
    for (;;) {
        x = u01_f64_from_bits(rng.next());
        if (x < e.c2) return x;
        double y = __dmul_rn(0.5, u01_f64_from_bits(rng.next()));
        if (y > __dadd_rn(1.0, -x)) {
            x = __dadd_rn(1.0, -x);
            y = __dadd_rn(1.0, -y);
        }
        if (x <= __dmul_rn(e.c3, __dadd_rn(1.0, -y))) return x;
        if (__dmul_rn(e.c1, y) <= __dadd_rn(1.0, -x)) return x;
        if (__dmul_rn(__dmul_rn(y, e.c1), e.lambda) <= expm1_spec(__dmul_rn(e.lambda, __dadd_rn(1.0, -x))))
            return x;
    }
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- 128-bit compare-and-swap in global memory (atom.global.cas.b128, SASS ATOMG.E.CAS.128) and the
// lexicographic minimum it gives: slot = min over (ordered bits of h, k-mer).  One atomic object per
// MinHash slot replaces the two passes "atomicMin of h" / "owners of the minimum write the k-mer".
__device__ __forceinline__ ulonglong2 cas128(ulonglong2 *addr, ulonglong2 cmp, ulonglong2 nw) {
    ulonglong2 prev;
    asm volatile(
        "{\n"
        ".reg .b128 c, n, r;\n"
        "mov.b128 c, {%2, %3};\n"
        "mov.b128 n, {%4, %5};\n"
        "atom.global.cas.b128 r, [%6], c, n;\n"
        "mov.b128 {%0, %1}, r;\n"
        "}\n"
        : "=l"(prev.x), "=l"(prev.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(nw.x), "l"(nw.y), "l"(addr)
        : "memory");
    return prev;
}
__device__ __forceinline__ void slot_min128(ulonglong2 *addr, unsigned long long hb, unsigned long long d) {
    ulonglong2 old = __ldcg(addr);  // a stale (larger) value only costs one more trip
    for (;;) {
        if (!(hb < old.x || (hb == old.x && d < old.y))) return;
        const ulonglong2 prev = cas128(addr, old, make_ulonglong2(hb, d));
        if (prev.x == old.x && prev.y == old.y) return;
        old = prev;
    }
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    while (!mbar_try_wait(bar, phase)) {
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            (uint32_t)__cvta_generic_to_shared(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// L2 policy for data that is read once (streamed rows must not displace what is re-read)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                  uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            (uint32_t)__cvta_generic_to_shared(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}


}  // namespace gsb
