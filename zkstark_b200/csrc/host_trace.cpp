// prover.rs:32-39 on the host for MANY proofs at once: a[0] = 1, a[1] = secret, a[i] = a[i-2]^2 + a[i-1]^2 over F_p.
// The recurrence is one dependent squaring per step, so a single trace cannot go faster than ~7 ns per step on a core --
// but independent traces are SIMD lanes: 8 traces advance together in two interleaved AVX2 vectors of four 64-bit lanes
// (16 with AVX-512), at about the latency of one.  32 traces of 2^20 steps: ~10 ms on 3-4 cores instead of ~80 ms.
// Arithmetic: the same subtractive Montgomery form as field.cuh, a 32-bit residue per 64-bit lane.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

#include "field.cuh"
#include "hostair.hpp"

namespace zk {

#if defined(__x86_64__)
namespace {
struct V4 {  // four traces
    __m256i s, y;   // a[i-2]^2 and a[i-1], Montgomery form
};
__attribute__((target("avx2"), always_inline)) static inline __m256i mont_mul4(__m256i a, __m256i b, __m256i pinv, __m256i p, __m256i lo32)
{
    const __m256i t = _mm256_mul_epu32(a, b);                                  // 64-bit products of the low halves
    const __m256i m = _mm256_and_si256(_mm256_mul_epu32(t, pinv), lo32);       // lo(t) * p^-1 mod 2^32 (mul_epu32 reads lo(t))
    const __m256i u = _mm256_srli_epi64(_mm256_mul_epu32(m, p), 32);           // hi(m p)
    __m256i r = _mm256_sub_epi64(_mm256_srli_epi64(t, 32), u);                 // hi(t) - hi(m p) in (-p, p)
    const __m256i neg = _mm256_cmpgt_epi64(_mm256_setzero_si256(), r);
    return _mm256_add_epi64(r, _mm256_and_si256(neg, p));
}
// one step of all 8 traces (two independent chains: latency overlap); c[h]: the canonical a[i] of half h
template <int NV>
__attribute__((target("avx2"), always_inline)) static inline void step8(V4 (&v)[NV], __m256i (&c)[NV], __m256i p, __m256i pinv, __m256i lo32, __m256i one)
{
    for (int h = 0; h < NV; h++) {
        const __m256i yy = mont_mul4(v[h].y, v[h].y, pinv, p, lo32);        // a[i-1]^2
        __m256i z = _mm256_add_epi64(v[h].s, yy);                           // a[i] = a[i-2]^2 + a[i-1]^2, < 2p < 2^33
        const __m256i ge = _mm256_cmpgt_epi64(z, _mm256_sub_epi64(p, one)); // z >= p
        z = _mm256_sub_epi64(z, _mm256_and_si256(ge, p));
        v[h].s = yy;
        v[h].y = z;
        c[h] = mont_mul4(z, one, pinv, p, lo32);                            // out of Montgomery form (off the chain)
    }
}
// NV = 2: 8 traces as two interleaved vectors (best throughput per core); NV = 1: 4 traces (more groups for more cores)
template <int NV>
__attribute__((target("avx2"))) static void traces_avx2(uint32_t log_n, const uint32_t* secrets, uint32_t* const* rows)
{
    constexpr int NT = 4 * NV;
    const uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    const __m256i p = _mm256_set1_epi64x(P), pinv = _mm256_set1_epi64x(P_INV), lo32 = _mm256_set1_epi64x(0xffffffffll);
    const __m256i one = _mm256_set1_epi64x(1);
    alignas(32) uint64_t sec[NT], tmp[NT];
    for (int j = 0; j < NT; j++) sec[j] = secrets[j] % P;
    V4 v[NV];
    for (int h = 0; h < NV; h++) {
        alignas(32) uint64_t ym[4];
        for (int j = 0; j < 4; j++) ym[j] = to_mont((uint32_t)sec[4 * h + j]);
        v[h].y = _mm256_load_si256((const __m256i*)ym);
        v[h].s = _mm256_set1_epi64x(to_mont(1u));                               // a[0]^2 = 1
    }
    for (int j = 0; j < NT; j++) {
        rows[j][0] = 1;
        if (n1 > 1) rows[j][1] = (uint32_t)sec[j];
    }
#define step(c) step8<NV>(v, (c), p, pinv, lo32, one)
    uint64_t i = 2;
    for (; i < n1 && (i & 3); i++) {                                            // up to a multiple of 4: 16-byte aligned rows
        __m256i c[NV];
        step(c);
        for (int h = 0; h < NV; h++) _mm256_store_si256((__m256i*)(tmp + 4 * h), c[h]);
        for (int j = 0; j < NT; j++) rows[j][i] = (uint32_t)tmp[j];
    }
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);            // the low words of the four 64-bit lanes
    for (; i + 4 <= n1; i += 4) {
        // four steps, then a 4 x 4 transpose per half: every trace gets its four values with one 16-byte store
        __m256i c[4][NV];
        for (int k = 0; k < 4; k++) step(c[k]);
        for (int h = 0; h < NV; h++) {
            const __m256i t0 = _mm256_unpacklo_epi64(c[0][h], c[1][h]), t1 = _mm256_unpackhi_epi64(c[0][h], c[1][h]);
            const __m256i t2 = _mm256_unpacklo_epi64(c[2][h], c[3][h]), t3 = _mm256_unpackhi_epi64(c[2][h], c[3][h]);
            const __m256i r0 = _mm256_permute2x128_si256(t0, t2, 0x20), r1 = _mm256_permute2x128_si256(t1, t3, 0x20);
            const __m256i r2 = _mm256_permute2x128_si256(t0, t2, 0x31), r3 = _mm256_permute2x128_si256(t1, t3, 0x31);
            const __m256i rr[4] = {r0, r1, r2, r3};                             // trace 4h + j: steps i .. i+3 in its 64-bit lanes
            for (int j = 0; j < 4; j++)
                _mm_storeu_si128((__m128i*)(rows[4 * h + j] + i), _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(rr[j], pick)));
        }
    }
    for (; i < n1; i++) {
        __m256i c[NV];
        step(c);
        for (int h = 0; h < NV; h++) _mm256_store_si256((__m256i*)(tmp + 4 * h), c[h]);
        for (int j = 0; j < NT; j++) rows[j][i] = (uint32_t)tmp[j];
    }
}
#undef step

// Other programs of the AIR family (ctx.hpp HostAir): a[i] = c20 a[i-1]^2 + c02 a[i-2]^2 + c11 a[i-2] a[i-1] + c10 a[i-1]
// + c01 a[i-2] + c00, four traces per vector, NV vectors interleaved.  Seven multiplications per step instead of one.
__attribute__((target("avx2"), always_inline)) static inline __m256i addmod4(__m256i a, __m256i b, __m256i p, __m256i one)
{
    __m256i z = _mm256_add_epi64(a, b);
    const __m256i ge = _mm256_cmpgt_epi64(z, _mm256_sub_epi64(p, one));
    return _mm256_sub_epi64(z, _mm256_and_si256(ge, p));
}
template <int NV>
__attribute__((target("avx2"))) static void traces_air_avx2(const HostAir& air, uint32_t log_n, const uint32_t* secrets, uint32_t* const* rows)
{
    constexpr int NT = 4 * NV;
    const uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    const __m256i p = _mm256_set1_epi64x(P), pinv = _mm256_set1_epi64x(P_INV), lo32 = _mm256_set1_epi64x(0xffffffffll);
    const __m256i one = _mm256_set1_epi64x(1);
    const __m256i c20 = _mm256_set1_epi64x(to_mont(air.c20)), c02 = _mm256_set1_epi64x(to_mont(air.c02)),
                  c11 = _mm256_set1_epi64x(to_mont(air.c11)), c10 = _mm256_set1_epi64x(to_mont(air.c10)),
                  c01 = _mm256_set1_epi64x(to_mont(air.c01)), c00 = _mm256_set1_epi64x(to_mont(air.c00));
    alignas(32) uint64_t tmp[NT];
    __m256i x0[NV], x1[NV], q0[NV];   // a[i-2], a[i-1], a[i-2]^2 (Montgomery form)
    for (int h = 0; h < NV; h++) {
        alignas(32) uint64_t ym[4];
        for (int j = 0; j < 4; j++) ym[j] = to_mont(secrets[4 * h + j] % P);
        x1[h] = _mm256_load_si256((const __m256i*)ym);
        x0[h] = _mm256_set1_epi64x(to_mont(air.a0));
        q0[h] = mont_mul4(x0[h], x0[h], pinv, p, lo32);
    }
    for (int j = 0; j < NT; j++) {
        rows[j][0] = air.a0;
        if (n1 > 1) rows[j][1] = secrets[j] % P;
    }
    for (uint64_t i = 2; i < n1; i++) {
        for (int h = 0; h < NV; h++) {
            const __m256i q1 = mont_mul4(x1[h], x1[h], pinv, p, lo32);
            const __m256i xy = mont_mul4(x0[h], x1[h], pinv, p, lo32);
            __m256i z = addmod4(mont_mul4(q1, c20, pinv, p, lo32), mont_mul4(q0[h], c02, pinv, p, lo32), p, one);
            z = addmod4(z, mont_mul4(xy, c11, pinv, p, lo32), p, one);
            z = addmod4(z, addmod4(mont_mul4(x1[h], c10, pinv, p, lo32), mont_mul4(x0[h], c01, pinv, p, lo32), p, one), p, one);
            z = addmod4(z, c00, p, one);
            x0[h] = x1[h];
            q0[h] = q1;
            x1[h] = z;
            _mm256_store_si256((__m256i*)(tmp + 4 * h), mont_mul4(z, one, pinv, p, lo32));
        }
        for (int j = 0; j < NT; j++) rows[j][i] = (uint32_t)tmp[j];
    }
}
}  // namespace
#endif

// scalar reference of the same recurrence (also the tail of a batch that is not a multiple of 8)
static void trace1(uint32_t log_n, uint32_t secret, uint32_t* out)
{
    const uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    uint32_t y = to_mont(secret % P);
    uint32_t xx = mont_mul(to_mont(1), to_mont(1));
    out[0] = 1;
    if (n1 > 1) out[1] = secret % P;
    for (uint64_t i = 2; i < n1; i++) {
        const uint32_t yy = mont_mul(y, y);
        const uint32_t z = fadd(xx, yy);
        out[i] = from_mont(z);
        xx = yy;
        y = z;
    }
}

static void trace1_air(const HostAir& air, uint32_t log_n, uint32_t secret, uint32_t* out)
{
    const uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    const uint32_t c20 = to_mont(air.c20), c02 = to_mont(air.c02), c11 = to_mont(air.c11), c10 = to_mont(air.c10),
                   c01 = to_mont(air.c01), c00 = to_mont(air.c00);
    uint32_t x0 = to_mont(air.a0), x1 = to_mont(secret % P);
    out[0] = air.a0;
    if (n1 > 1) out[1] = secret % P;
    for (uint64_t i = 2; i < n1; i++) {
        uint32_t z = fadd(mont_mul(mont_mul(x1, x1), c20), mont_mul(mont_mul(x0, x0), c02));
        z = fadd(z, mont_mul(mont_mul(x0, x1), c11));
        z = fadd(z, fadd(mont_mul(x1, c10), mont_mul(x0, c01)));
        z = fadd(z, c00);
        out[i] = from_mont(z);
        x0 = x1;
        x1 = z;
    }
}

// traces of `count` proofs: rows[j] receives the 2^log_n - 1 values of secrets[j].  count = 8 or 4 take the SIMD paths.
void host_traces(const HostAir& air, uint32_t log_n, const uint32_t* secrets, uint32_t* const* rows, uint32_t count)
{
    uint32_t j = 0;
    if (air.generic) {
#if defined(__x86_64__)
        if (__builtin_cpu_supports("avx2")) {
            for (; j + 8 <= count; j += 8) traces_air_avx2<2>(air, log_n, secrets + j, rows + j);
            for (; j + 4 <= count; j += 4) traces_air_avx2<1>(air, log_n, secrets + j, rows + j);
        }
#endif
        for (; j < count; j++) trace1_air(air, log_n, secrets[j], rows[j]);
        return;
    }
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) {
        for (; j + 8 <= count; j += 8) traces_avx2<2>(log_n, secrets + j, rows + j);
        for (; j + 4 <= count; j += 4) traces_avx2<1>(log_n, secrets + j, rows + j);
    }
#endif
    for (; j < count; j++) trace1(log_n, secrets[j], rows[j]);
}

}  // namespace zk
