/*
 * oracle/sha256.c -- CPU ORACLE (test infrastructure).  FIPS 180-4 SHA-256, the
 * function `sha2 0.10.6` computes for merkle.rs:31-33,42-45 and channel.rs:21-24.
 * Portable C plus an x86 SHA-NI block function picked at run time (the CPU
 * baseline should use what the host offers; sha2 0.10 does the same via cpufeatures).
 */
#include "oracle.h"
#include <string.h>

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }

static void blocks_portable(uint32_t st[8], const uint8_t *data, size_t nblocks)
{
    uint32_t w[64];
    while (nblocks--) {
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)data[4 * i] << 24) | ((uint32_t)data[4 * i + 1] << 16) |
                   ((uint32_t)data[4 * i + 2] << 8) | data[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = h + S1 + ch + K256[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
        data += 64;
    }
}

#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>

__attribute__((target("sha,sse4.1,ssse3")))
static void blocks_shani(uint32_t st[8], const uint8_t *data, size_t nblocks)
{
    const __m128i mask = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i tmp = _mm_loadu_si128((const __m128i *)&st[0]);
    __m128i s1 = _mm_loadu_si128((const __m128i *)&st[4]);
    tmp = _mm_shuffle_epi32(tmp, 0xB1);
    s1 = _mm_shuffle_epi32(s1, 0x1B);
    __m128i s0 = _mm_alignr_epi8(tmp, s1, 8);
    s1 = _mm_blend_epi16(s1, tmp, 0xF0);
    while (nblocks--) {
        __m128i save0 = s0, save1 = s1, m[4], msg;
        for (int i = 0; i < 16; i++) {
            if (i < 4)
                m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(data + 16 * i)), mask);
            msg = _mm_add_epi32(m[i & 3], _mm_loadu_si128((const __m128i *)&K256[4 * i]));
            s1 = _mm_sha256rnds2_epu32(s1, s0, msg);
            if (i >= 3 && i < 15) {
                tmp = _mm_alignr_epi8(m[i & 3], m[(i - 1) & 3], 4);
                m[(i + 1) & 3] = _mm_add_epi32(m[(i + 1) & 3], tmp);
                m[(i + 1) & 3] = _mm_sha256msg2_epu32(m[(i + 1) & 3], m[i & 3]);
            }
            msg = _mm_shuffle_epi32(msg, 0x0E);
            s0 = _mm_sha256rnds2_epu32(s0, s1, msg);
            if (i >= 1 && i < 13)
                m[(i - 1) & 3] = _mm_sha256msg1_epu32(m[(i - 1) & 3], m[i & 3]);
        }
        s0 = _mm_add_epi32(s0, save0);
        s1 = _mm_add_epi32(s1, save1);
        data += 64;
    }
    tmp = _mm_shuffle_epi32(s0, 0x1B);
    s1 = _mm_shuffle_epi32(s1, 0xB1);
    s0 = _mm_blend_epi16(tmp, s1, 0xF0);
    s1 = _mm_alignr_epi8(s1, tmp, 8);
    _mm_storeu_si128((__m128i *)&st[0], s0);
    _mm_storeu_si128((__m128i *)&st[4], s1);
}

static int detect_shani(void)
{
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return 0;
    int sha = (b >> 29) & 1;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return 0;
    int ssse3 = (c >> 9) & 1, sse41 = (c >> 19) & 1;
    return sha && ssse3 && sse41;
}
#else
static void blocks_shani(uint32_t st[8], const uint8_t *d, size_t n) { blocks_portable(st, d, n); }
static int detect_shani(void) { return 0; }
#endif

static int g_shani = -1;
static int g_force_portable = 0;

int orc_sha256_uses_shani(void)
{
    if (g_shani < 0) g_shani = detect_shani();
    return g_shani && !g_force_portable;
}

void orc_sha256_force_portable(int on) { g_force_portable = on; }

static inline void blocks(uint32_t st[8], const uint8_t *data, size_t n)
{
    if (orc_sha256_uses_shani()) blocks_shani(st, data, n);
    else blocks_portable(st, data, n);
}

/* raw compression over whole 64-byte blocks (used by merkle.c for fixed-size inputs) */
void orc_sha256_blocks(uint32_t st[8], const uint8_t *data, size_t nblocks) { blocks(st, data, nblocks); }

void orc_sha256(const uint8_t *data, size_t len, uint8_t out[32])
{
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                      0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t full = len / 64;
    if (full) blocks(st, data, full);
    uint8_t tail[128];
    size_t rem = len - full * 64;
    memset(tail, 0, sizeof tail);
    if (rem) memcpy(tail, data + full * 64, rem);
    tail[rem] = 0x80;
    size_t tl = (rem < 56) ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    blocks(st, tail, tl / 64);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(st[i] >> 24);
        out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8);
        out[4 * i + 3] = (uint8_t)st[i];
    }
}
