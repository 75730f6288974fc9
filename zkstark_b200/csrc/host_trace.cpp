// prover.rs:32-39 on the host for MANY proofs at once: a[0] = 1, a[1] = secret, a[i] = a[i-2]^2 + a[i-1]^2 over F_p.
// The recurrence is one dependent squaring per step, so a single trace cannot go faster than ~7 ns per step on a core --
// but independent traces are SIMD lanes: 8 traces advance together in two interleaved AVX2 vectors of four 64-bit lanes
// (16 with AVX-512), at about the latency of one.  32 traces of 2^20 steps: ~10 ms on 3-4 cores instead of ~80 ms.
// Arithmetic: the same subtractive Montgomery form as field.cuh, a 32-bit residue per 64-bit lane.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

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

// The same recurrence with AVX-512 IFMA (vpmadd52luq / vpmadd52huq, 4-cycle 52-bit multiply-adds): Montgomery form with
// R = 2^52, eight traces per 512-bit vector, NV vectors interleaved.  The dependent chain of a step shrinks from ~29 to ~17
// cycles (three vpmadd52 + carry + two min-subtracts against three vpmuludq + compare / blend sequences): a group of 8
// traces of 2^20 steps takes 12.8 ms of a core against 18.7 ms with AVX2 (15.9 ms for a group of 4).
//   t = a b < 2^64:  t_lo = lo52(t), t_hi = t >> 52;   m = lo52(t_lo * (-p^-1 mod 2^52));
//   (t + m p) >> 52 = t_hi + hi52(m p) + [t_lo != 0]  in [0, 2p)  -> one conditional subtraction
constexpr uint64_t M52 = (1ull << 52) - 1;
constexpr uint64_t neg_pinv52()
{
    uint64_t x = 1;                                   // Newton: x <- x (2 - p x) doubles the correct low bits
    for (int i = 0; i < 6; i++) x *= 2 - (uint64_t)P * x;
    return (0 - x) & M52;
}
static_assert(((uint64_t)P * ((0 - neg_pinv52()) & M52) & M52) == 1, "p * p^-1 == 1 mod 2^52");
static inline uint64_t to_mont52(uint32_t a) { return (uint64_t)((((unsigned __int128)a) << 52) % P); }
#define ZK_IFMA_TARGET __attribute__((target("avx512f,avx512ifma,avx512dq,avx512vl"), always_inline))
ZK_IFMA_TARGET static inline __m512i mont_mul8(__m512i a, __m512i b, __m512i ninv, __m512i p, __m512i one)
{
    const __m512i zero = _mm512_setzero_si512();
    const __m512i t_lo = _mm512_madd52lo_epu64(zero, a, b);
    const __m512i t_hi = _mm512_madd52hi_epu64(zero, a, b);
    const __m512i m = _mm512_madd52lo_epu64(zero, t_lo, ninv);
    __m512i r = _mm512_madd52hi_epu64(t_hi, m, p);
    r = _mm512_mask_add_epi64(r, _mm512_test_epi64_mask(t_lo, t_lo), r, one);
    return _mm512_min_epu64(r, _mm512_sub_epi64(r, p));
}
template <int NV>
__attribute__((target("avx512f,avx512ifma,avx512dq,avx512vl"))) static void traces_ifma(uint32_t log_n, const uint32_t* secrets,
                                                                                        uint32_t* const* rows)
{
    constexpr int NT = 8 * NV;
    const uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    const __m512i p = _mm512_set1_epi64(P), ninv = _mm512_set1_epi64((long long)neg_pinv52()), one = _mm512_set1_epi64(1);
    alignas(64) uint64_t tmp[NT];
    __m512i y[NV], sq[NV];                            // a[i-1] and a[i-2]^2, Montgomery form (R = 2^52)
    for (int h = 0; h < NV; h++) {
        alignas(64) uint64_t ym[8];
        for (int j = 0; j < 8; j++) ym[j] = to_mont52(secrets[8 * h + j] % P);
        y[h] = _mm512_load_si512(ym);
        sq[h] = _mm512_set1_epi64((long long)to_mont52(1u));
    }
    for (int j = 0; j < NT; j++) {
        rows[j][0] = 1;
        if (n1 > 1) rows[j][1] = secrets[j] % P;
    }
    auto step = [&](__m512i (&c)[NV]) __attribute__((target("avx512f,avx512ifma,avx512dq,avx512vl"), always_inline)) {
        for (int h = 0; h < NV; h++) {
            const __m512i yy = mont_mul8(y[h], y[h], ninv, p, one);
            __m512i z = _mm512_add_epi64(sq[h], yy);
            z = _mm512_min_epu64(z, _mm512_sub_epi64(z, p));
            sq[h] = yy;
            y[h] = z;
            c[h] = mont_mul8(z, one, ninv, p, one);   // canonical a[i], off the chain
        }
    };
    uint64_t i = 2;
    for (; i < n1 && (i & 3); i++) {
        __m512i c[NV];
        step(c);
        for (int h = 0; h < NV; h++) _mm512_store_si512(tmp + 8 * h, c[h]);
        for (int j = 0; j < NT; j++) rows[j][i] = (uint32_t)tmp[j];
    }
    for (; i + 4 <= n1; i += 4) {
        // four steps, then per vector: (c0 | c1 << 32) and (c2 | c3 << 32) interleaved per trace = 16 bytes per trace
        __m512i c[4][NV];
        for (int k = 0; k < 4; k++) step(c[k]);
        for (int h = 0; h < NV; h++) {
            const __m512i u01 = _mm512_or_si512(c[0][h], _mm512_slli_epi64(c[1][h], 32));
            const __m512i u23 = _mm512_or_si512(c[2][h], _mm512_slli_epi64(c[3][h], 32));
            const __m512i ev = _mm512_unpacklo_epi64(u01, u23);    // traces 0, 2, 4, 6 of the vector, 128 bits each
            const __m512i od = _mm512_unpackhi_epi64(u01, u23);    // traces 1, 3, 5, 7
            uint32_t* const* r = rows + 8 * h;
            _mm_storeu_si128((__m128i*)(r[0] + i), _mm512_extracti32x4_epi32(ev, 0));
            _mm_storeu_si128((__m128i*)(r[1] + i), _mm512_extracti32x4_epi32(od, 0));
            _mm_storeu_si128((__m128i*)(r[2] + i), _mm512_extracti32x4_epi32(ev, 1));
            _mm_storeu_si128((__m128i*)(r[3] + i), _mm512_extracti32x4_epi32(od, 1));
            _mm_storeu_si128((__m128i*)(r[4] + i), _mm512_extracti32x4_epi32(ev, 2));
            _mm_storeu_si128((__m128i*)(r[5] + i), _mm512_extracti32x4_epi32(od, 2));
            _mm_storeu_si128((__m128i*)(r[6] + i), _mm512_extracti32x4_epi32(ev, 3));
            _mm_storeu_si128((__m128i*)(r[7] + i), _mm512_extracti32x4_epi32(od, 3));
        }
    }
    for (; i < n1; i++) {
        __m512i c[NV];
        step(c);
        for (int h = 0; h < NV; h++) _mm512_store_si512(tmp + 8 * h, c[h]);
        for (int j = 0; j < NT; j++) rows[j][i] = (uint32_t)tmp[j];
    }
}

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

int host_trace_isa();

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
    const int isa = host_trace_isa();
    if (isa >= 2) {
        for (; j + 8 <= count; j += 8) traces_ifma<1>(log_n, secrets + j, rows + j);
    }
    if (isa >= 1) {
        for (; j + 8 <= count; j += 8) traces_avx2<2>(log_n, secrets + j, rows + j);
        for (; j + 4 <= count; j += 4) traces_avx2<1>(log_n, secrets + j, rows + j);
    }
#endif
    for (; j < count; j++) trace1(log_n, secrets[j], rows[j]);
}

// 0 scalar, 1 AVX2, 2 AVX-512 IFMA: what this CPU has, capped by ZK_HOST_SIMD=scalar|avx2 (tests exercise every path)
int host_trace_isa()
{
    int isa = 0;
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) isa = 1;
    if (isa == 1 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512ifma") && __builtin_cpu_supports("avx512dq") &&
        __builtin_cpu_supports("avx512vl"))
        isa = 2;
#endif
    if (const char* e = getenv("ZK_HOST_SIMD")) {
        if (e[0] == 's') isa = 0;
        else if (e[0] == 'a' && isa > 1) isa = 1;
    }
    return isa;
}

// Traces per group for `n` proofs on `threads` host threads.  A group advances at about the latency of ONE recurrence, so
// the choice minimises (rounds of groups per thread) x (time of a group).  Measured, 2^20 steps on one core: 4 lanes AVX2
// 15.9 ms (1.0), 8 lanes AVX2 18.7 ms (1.18), 8 lanes IFMA 12.8 ms (0.8); two interleaved IFMA vectors (16 lanes) take 31.9 ms
// -- 512-bit multiply-adds issue on one port there, the pair is throughput bound -- and are not used.  Other programs of
// the AIR family have AVX2 code only.
uint32_t host_trace_group(uint32_t n, uint32_t threads, bool generic)
{
    const int isa = host_trace_isa();
    struct Opt { uint32_t g; double t; };
    const Opt avx2[] = {{4, 1.0}, {8, 1.18}}, ifma[] = {{8, 0.8}, {4, 1.0}};
    const Opt* o = (isa >= 2 && !generic) ? ifma : avx2;
    const int no = 2;
    if (threads < 1) threads = 1;
    uint32_t best = 4;
    double best_t = 1e30;
    for (int i = 0; i < no; i++) {
        const uint32_t groups = (n + o[i].g - 1) / o[i].g, rounds = (groups + threads - 1) / threads;
        const double t = rounds * o[i].t;
        if (t < best_t - 1e-9) { best_t = t; best = o[i].g; }
    }
    return best;
}

}  // namespace zk
