// Host SHA-256 (FIPS 180-4) for the sequential tail of the Fiat-Shamir channel
// (channel.rs:19-26 applied to the decommitment records, prover.rs:274-288): a few hundred
// dependent compressions per proof -- glue, not data-parallel work.
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#if defined(__x86_64__) && !defined(ZK_HOST_SHA_PORTABLE)
#include <cpuid.h>
#include <immintrin.h>
#define ZK_HOST_SHANI 1
#endif

namespace zk {
namespace sha {

static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void block(uint32_t s[8], const uint8_t* p)
{
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++)
        w[i] = w[i - 16] + (ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] +
               (ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10));
    uint32_t a = s[0], b = s[1], c = s[2], d = s[3], e = s[4], f = s[5], g = s[6], h = s[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
        uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s[0] += a; s[1] += b; s[2] += c; s[3] += d; s[4] += e; s[5] += f; s[6] += g; s[7] += h;
}

#ifdef ZK_HOST_SHANI
// The same compression with the x86 SHA extensions when the CPU has them (picked once at run time): the host tail of a
// single proof is ~340 dependent compressions, 0.14 ms with the portable code, ~0.02 ms with SHA-NI.
__attribute__((target("sha,sse4.1,ssse3"))) static void block_shani(uint32_t s[8], const uint8_t* p)
{
    const __m128i be = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i lo = _mm_loadu_si128((const __m128i*)&s[0]);    // a b c d
    __m128i hi = _mm_loadu_si128((const __m128i*)&s[4]);    // e f g h
    lo = _mm_shuffle_epi32(lo, 0xB1);
    hi = _mm_shuffle_epi32(hi, 0x1B);
    __m128i abef = _mm_alignr_epi8(lo, hi, 8), cdgh = _mm_blend_epi16(hi, lo, 0xF0);
    const __m128i abef0 = abef, cdgh0 = cdgh;
    __m128i w[4];
    for (int q = 0; q < 16; q++) {                           // four rounds per trip
        if (q < 4) w[q] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * q)), be);
        __m128i kw = _mm_add_epi32(w[q & 3], _mm_loadu_si128((const __m128i*)&K[4 * q]));
        cdgh = _mm_sha256rnds2_epu32(cdgh, abef, kw);
        if (q >= 3 && q < 15) {
            __m128i t = _mm_add_epi32(w[(q + 1) & 3], _mm_alignr_epi8(w[q & 3], w[(q - 1) & 3], 4));
            w[(q + 1) & 3] = _mm_sha256msg2_epu32(t, w[q & 3]);
        }
        abef = _mm_sha256rnds2_epu32(abef, cdgh, _mm_shuffle_epi32(kw, 0x0E));
        if (q >= 1 && q < 13) w[(q - 1) & 3] = _mm_sha256msg1_epu32(w[(q - 1) & 3], w[q & 3]);
    }
    abef = _mm_add_epi32(abef, abef0);
    cdgh = _mm_add_epi32(cdgh, cdgh0);
    lo = _mm_shuffle_epi32(abef, 0x1B);
    hi = _mm_shuffle_epi32(cdgh, 0xB1);
    _mm_storeu_si128((__m128i*)&s[0], _mm_blend_epi16(lo, hi, 0xF0));
    _mm_storeu_si128((__m128i*)&s[4], _mm_alignr_epi8(hi, lo, 8));
}
static bool cpu_has_shani()
{
    unsigned a = 0, b = 0, c = 0, d = 0;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d) || !((b >> 29) & 1)) return false;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
    return ((c >> 9) & 1) && ((c >> 19) & 1);                // SSSE3, SSE4.1
}
static void block_any(uint32_t s[8], const uint8_t* p)
{
    static const bool ni = cpu_has_shani();
    if (ni) block_shani(s, p);
    else block(s, p);
}
#else
static void block_any(uint32_t s[8], const uint8_t* p) { block(s, p); }
#endif

void host_sha256(const uint8_t* data, size_t len, uint8_t out[32])
{
    uint32_t s[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t full = len / 64;
    for (size_t i = 0; i < full; i++) block_any(s, data + 64 * i);
    uint8_t tail[128];
    size_t rem = len - 64 * full;
    memset(tail, 0, sizeof tail);
    if (rem) memcpy(tail, data + 64 * full, rem);
    tail[rem] = 0x80;
    size_t tl = rem < 56 ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    block_any(s, tail);
    if (tl == 128) block_any(s, tail + 64);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(s[i] >> 24); out[4 * i + 1] = (uint8_t)(s[i] >> 16);
        out[4 * i + 2] = (uint8_t)(s[i] >> 8); out[4 * i + 3] = (uint8_t)s[i];
    }
}

}  // namespace sha
}  // namespace zk
