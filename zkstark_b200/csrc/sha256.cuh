// SHA-256 for the Merkle kernels (K5) and the device-side Fiat-Shamir channel.
//
// Replaces sha2::Sha256 as used by merkle.rs:30-34 (leaf = SHA256(4-byte big-endian u32)),
// merkle.rs:40-46 (node = SHA256(left || right)) and channel.rs:21-24.
// Everything is fully unrolled so the round constants become immediates; the two
// fixed-shape messages of the tree get dedicated entry points:
//   leaf : one block  W0 = v, W1 = 0x80000000, W2..14 = 0, W15 = 32   (most of the
//          schedule and the first rounds constant-fold at compile time)
//   node : data block (16 digest words) + the CONSTANT padding block
//          (W0 = 0x80000000, W15 = 512) whose K[i]+W[i] are compile-time constants.
// Digest words are kept as native u32 (H0..H7); byte order is applied on store/load.
#pragma once
#include <stdint.h>

namespace zk {
namespace sha {

struct Consts {
    uint32_t k[64];
    uint32_t kw_pad[64];  // K[i] + W[i] of the node padding block
};

constexpr uint32_t c_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

constexpr Consts make_consts()
{
    Consts c{};
    const uint32_t k[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64] = {};
    w[0] = 0x80000000u;
    w[15] = 512u;
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = c_rotr(w[i - 15], 7) ^ c_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = c_rotr(w[i - 2], 17) ^ c_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    for (int i = 0; i < 64; i++) {
        c.k[i] = k[i];
        c.kw_pad[i] = k[i] + w[i];
    }
    return c;
}

#define ZK_SHA_IV0 0x6a09e667u
#define ZK_SHA_IV1 0xbb67ae85u
#define ZK_SHA_IV2 0x3c6ef372u
#define ZK_SHA_IV3 0xa54ff53au
#define ZK_SHA_IV4 0x510e527fu
#define ZK_SHA_IV5 0x9b05688cu
#define ZK_SHA_IV6 0x1f83d9abu
#define ZK_SHA_IV7 0x5be0cd19u

#if defined(__CUDACC__) && defined(ZK_SHA_DEVICE)

// ---- pipe balancing --------------------------------------------------------------------------
// ncu (profiles/r1a): the hashing kernels saturate the ALU pipe (SHF/LOP3/IADD3, 84 %) while the FMA
// pipe idles (5 %).  Both pipes issue one warp-instruction per 2 cycles per SM sub-partition, so work
// moved to IMAD is free until the pipes balance.  ptxas would turn `x*1 + y` or `x * 2^k` back into
// ALU ops, so the multipliers are read from the constant bank (opaque to the compiler):
//   add       a + b                 ->  IMAD      a * c[ONE] + b
//   shift     x >> k                ->  IMAD.HI   hi(x * c[2^(32-k)])
//   rotate    rotr(x, k) = hi | lo  ->  IMAD.WIDE x * c[2^(32-k)]   (hi, lo disjoint: OR == XOR)
// `V` is a bit mask choosing which groups of operations move (tuned on the GPU):
//   1 round adds   2 schedule adds   4 Sigma1 rotations   8 Sigma0 rotations   16 sigma0/1 shifts   32 K+W adds
//   64 latency form of the round (see ZK_SHA_ROUND): for code that runs one warp per sub-partition
//   512 six-add form of the round (throughput kernels); 4096 its faster statement order
//   128 / 256: give two of the round's sums back to the ALU pipe as 3-input adds (IADD3): IMAD issues at about one warp
//   instruction per 3.6 cycles per sub-partition, not per 2, so a kernel with 7 IMAD per round is FMA-pipe bound
enum ShaConst { SC_ONE = 0, SC_R6, SC_R11, SC_R2, SC_R13, SC_S3, SC_S10, SC_R18, SC_R19, SC_COUNT };
__constant__ uint32_t c_sha[16];  // one translation unit (tree.cu) owns the hashing kernels

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
template <int V, int BIT>
__device__ __forceinline__ uint32_t addv(uint32_t a, uint32_t b)
{
    if (V & BIT) return a * c_sha[SC_ONE] + b;  // IMAD (FMA pipe)
    return a + b;
}
__device__ __forceinline__ void mulwide(uint32_t x, int ci, uint32_t& lo, uint32_t& hi)
{
    uint64_t p = (uint64_t)x * c_sha[ci];  // IMAD.WIDE.U32
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
}
template <int V>
__device__ __forceinline__ uint32_t bsig0(uint32_t x)
{
    if (V & 8) {
        uint32_t l2, h2, l13, h13;
        mulwide(x, SC_R2, l2, h2);
        mulwide(x, SC_R13, l13, h13);
        return (l2 ^ h2 ^ l13) ^ h13 ^ rotr(x, 22);
    }
    return rotr(x, 2) ^ rotr(x, 13) ^ rotr(x, 22);
}
template <int V>
__device__ __forceinline__ uint32_t bsig1(uint32_t x)
{
    if (V & 4) {
        uint32_t l6, h6, l11, h11;
        mulwide(x, SC_R6, l6, h6);
        mulwide(x, SC_R11, l11, h11);
        return (l6 ^ h6 ^ l11) ^ h11 ^ rotr(x, 25);
    }
    return rotr(x, 6) ^ rotr(x, 11) ^ rotr(x, 25);
}
template <int V>
__device__ __forceinline__ uint32_t ssig0(uint32_t x)
{
    if (V & 16) return rotr(x, 7) ^ rotr(x, 18) ^ __umulhi(x, c_sha[SC_S3]);
    return rotr(x, 7) ^ rotr(x, 18) ^ (x >> 3);
}
template <int V>
__device__ __forceinline__ uint32_t ssig1(uint32_t x)
{
    if (V & 16) return rotr(x, 17) ^ rotr(x, 19) ^ __umulhi(x, c_sha[SC_S10]);
    return rotr(x, 17) ^ rotr(x, 19) ^ (x >> 10);
}
// One message-schedule update W[t] = W[t-16] + sigma0(W[t-15]) + W[t-7] + sigma1(W[t-2]); ZK_SCHED_ORDER picks the order
// of the three adds (ptxas schedules them differently; see profiles/README).
#ifndef ZK_SCHED_ORDER
#define ZK_SCHED_ORDER 0
#endif
template <int V>
__device__ __forceinline__ uint32_t sched_word(uint32_t w16, uint32_t w15, uint32_t w7, uint32_t w2)
{
    if (ZK_SCHED_ORDER == 1) return addv<V, 2>(addv<V, 2>(addv<V, 2>(w16, w7), ssig0<V>(w15)), ssig1<V>(w2));
    if (ZK_SCHED_ORDER == 2) return addv<V, 2>(addv<V, 2>(w16, w7), addv<V, 2>(ssig0<V>(w15), ssig1<V>(w2)));
    if (ZK_SCHED_ORDER == 3) return addv<V, 2>(addv<V, 2>(addv<V, 2>(w16, ssig0<V>(w15)), w7), ssig1<V>(w2));
    return addv<V, 2>(addv<V, 2>(addv<V, 2>(w16, ssig1<V>(w2)), w7), ssig0<V>(w15));
}

__device__ __forceinline__ uint32_t ch(uint32_t e, uint32_t f, uint32_t g) { return (e & f) ^ (~e & g); }
__device__ __forceinline__ uint32_t maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) ^ (a & c) ^ (b & c); }

// One round.  The adds are ordered so that the newest value (e, via Sigma1/Ch) enters last: the chain from
// e to the next e is SHF -> LOP3 -> add -> add, everything else (h + K + W, d + h + K + W, Sigma0 + Maj)
// is off the critical path.  Matters for the latency-bound upper levels of a tree (one warp per SM).
#define ZK_SHA_ROUND(V, a, b, c, d, e, f, g, h, kw)                      \
    if ((V) & 64) {                                                      \
        /* latency form (one warp per sub-partition): h + K + W and d + h + K + W on the FMA pipe, off the  \
           chain; the three sums that consume the fresh e / a are single 3-input adds (IADD3) */            \
        uint32_t hk = addv<1, 1>(h, (kw));                               \
        uint32_t dhk = addv<1, 1>(d, hk);                                \
        uint32_t s1 = bsig1<(V) & 4>(e), cc = ch(e, f, g);               \
        uint32_t t2 = hk + bsig0<(V) & 8>(a) + maj(a, b, c);             \
        d = dhk + s1 + cc;                                               \
        h = t2 + s1 + cc;                                                \
    } else if ((V) & 512) {                                              \
        /* throughput form: six adds instead of seven (T1 is formed once and reused); the longer dependency chain  \
           is irrelevant when many warps are resident */                                                          \
        /* The order of the six statements matters to ptxas (2^23-leaf pass: 1306 us as first written, 1290 us with      \
           bit 4096 = Ch + Sigma1 first and d before Sigma0 + Maj; other orders 1308 .. 1342 us; nested expressions 1331) */ \
        if ((V) & 4096) {                                                \
            uint32_t y = addv<V, 1>(ch(e, f, g), bsig1<V>(e));           \
            uint32_t hk = addv<V, 1>(h, (kw));                           \
            uint32_t t1 = addv<V, 1>(hk, y);                             \
            d = addv<V, 1>(d, t1);                                       \
            uint32_t s0m = addv<V, 1>(bsig0<V>(a), maj(a, b, c));        \
            h = addv<V, 1>(t1, s0m);                                     \
        } else {                                                         \
            uint32_t hk = addv<V, 1>(h, (kw));                           \
            uint32_t y = addv<V, 1>(ch(e, f, g), bsig1<V>(e));           \
            uint32_t t1 = addv<V, 1>(hk, y);                             \
            uint32_t s0m = addv<V, 1>(bsig0<V>(a), maj(a, b, c));        \
            d = addv<V, 1>(d, t1);                                       \
            h = addv<V, 1>(t1, s0m);                                     \
        }                                                                \
    } else {                                                             \
        /* bit 128: h + K + W as one 3-input add on the ALU pipe; bit 256: the sum that produces the new a likewise */ \
        uint32_t hk = ((V) & 128) ? h + (kw) : addv<V, 1>(h, (kw));      \
        uint32_t dhk = addv<V, 1>(d, hk);                                \
        uint32_t y = addv<V, 1>(ch(e, f, g), bsig1<V>(e));               \
        uint32_t s0m = addv<V, 1>(bsig0<V>(a), maj(a, b, c));            \
        d = addv<V, 1>(dhk, y);                                          \
        h = ((V) & 256) ? hk + y + s0m : addv<V, 1>(addv<V, 1>(hk, y), s0m); \
    }

// Round constants for the rolled loops (indexed by the loop counter: a uniform constant-bank load).
__constant__ __align__(16) uint32_t c_k[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
__constant__ __align__(16) uint32_t c_kwpad[64];  // K[i] + W[i] of the node padding block (filled by sha_init_constants)

#define ZK_SHA_16ROUNDS(V, KW)                                        \
    ZK_SHA_ROUND(V, a, b, c, d, e, f, g, h, KW(0));                   \
    ZK_SHA_ROUND(V, h, a, b, c, d, e, f, g, KW(1));                   \
    ZK_SHA_ROUND(V, g, h, a, b, c, d, e, f, KW(2));                   \
    ZK_SHA_ROUND(V, f, g, h, a, b, c, d, e, KW(3));                   \
    ZK_SHA_ROUND(V, e, f, g, h, a, b, c, d, KW(4));                   \
    ZK_SHA_ROUND(V, d, e, f, g, h, a, b, c, KW(5));                   \
    ZK_SHA_ROUND(V, c, d, e, f, g, h, a, b, KW(6));                   \
    ZK_SHA_ROUND(V, b, c, d, e, f, g, h, a, KW(7));                   \
    ZK_SHA_ROUND(V, a, b, c, d, e, f, g, h, KW(8));                   \
    ZK_SHA_ROUND(V, h, a, b, c, d, e, f, g, KW(9));                   \
    ZK_SHA_ROUND(V, g, h, a, b, c, d, e, f, KW(10));                  \
    ZK_SHA_ROUND(V, f, g, h, a, b, c, d, e, KW(11));                  \
    ZK_SHA_ROUND(V, e, f, g, h, a, b, c, d, KW(12));                  \
    ZK_SHA_ROUND(V, d, e, f, g, h, a, b, c, KW(13));                  \
    ZK_SHA_ROUND(V, c, d, e, f, g, h, a, b, KW(14));                  \
    ZK_SHA_ROUND(V, b, c, d, e, f, g, h, a, KW(15));

// One compression of a data block.  st: chaining value in/out.  w: 16 message words (clobbered).
// ROLLED: rounds 0-15 are straight-line (round constants as immediates, constant message words fold);
// rounds 16-63 are a 3-trip loop of {16 schedule updates, 16 rounds}.  The fully unrolled form is
// ~1400 instructions (22 KB) per compression and made the hashing kernels instruction-fetch bound
// (ncu r1a: stall_no_instruction was the top stall with 48 warps/SM walking 58 KB of code).
template <int V>
__device__ __forceinline__ void compress(uint32_t st[8], uint32_t w[16])
{
    constexpr Consts C = make_consts();
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#define ZK_KW_FIRST(j) (addv<V, 32>(w[j], C.k[j]))
    ZK_SHA_16ROUNDS(V, ZK_KW_FIRST)
#undef ZK_KW_FIRST
#pragma unroll 1
    for (int i = 16; i < 64; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++)
            w[j] = sched_word<V>(w[j], w[(j + 1) & 15], w[(j + 9) & 15], w[(j + 14) & 15]);
        // fetch the 16 round constants of this trip up front (4 x 128-bit constant loads): loaded one per
        // round right before use, their latency is exposed 48 times per compression when a single warp runs
        uint32_t kk[16];
        {
            const uint4* kp = reinterpret_cast<const uint4*>(c_k + i);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 t = kp[q];
                kk[4 * q] = t.x; kk[4 * q + 1] = t.y; kk[4 * q + 2] = t.z; kk[4 * q + 3] = t.w;
            }
        }
#define ZK_KW_LOOP(j) (addv<V, 32>(w[j], kk[j]))
        ZK_SHA_16ROUNDS(V, ZK_KW_LOOP)
#undef ZK_KW_LOOP
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// The constant padding block that follows a 64-byte message (node second block): no schedule work,
// K[i]+W[i] comes from the constant bank.
template <int V>
__device__ __forceinline__ void compress_pad64(uint32_t st[8])
{
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int i = 0; i < 64; i += 16) {
        uint32_t kk[16];
        {
            const uint4* kp = reinterpret_cast<const uint4*>(c_kwpad + i);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 t = kp[q];
                kk[4 * q] = t.x; kk[4 * q + 1] = t.y; kk[4 * q + 2] = t.z; kk[4 * q + 3] = t.w;
            }
        }
#define ZK_KW_PAD(j) (kk[j])
        ZK_SHA_16ROUNDS(V, ZK_KW_PAD)
#undef ZK_KW_PAD
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// Leaf-or-node compression sharing ONE rolled round loop (the production leaf pass, UNI == 2):
//   leaf (merkle.rs:30-34): W0 = v, W1 = 0x80000000, W2..14 = 0, W15 = 32.  Rounds 0-15 are straight-line with
//     the chaining value = IV and 15 constant words folded at compile time (round 0 collapses to two adds), and
//     the first schedule trip W16..W31 is specialised: 9 of the 32 sigma terms survive.  Then the shared loop
//     runs rounds 16-63.
//   node data block: all 64 rounds go through the shared loop (round constants from the constant bank).
// Instruction footprint is the same as one generic compression + the leaf prologue (~17 KB with the padding
// block), which keeps the kernel out of the instruction-fetch-bound regime of the fully unrolled form.
template <int V>
__device__ __forceinline__ void sched_leaf(uint32_t v, uint32_t w[16])
{
    // constants of the leaf block: c_i = the part of W_i that does not depend on v
    constexpr uint32_t W1 = 0x80000000u, W15 = 32u;
    constexpr uint32_t s0_W1 = c_rotr(W1, 7) ^ c_rotr(W1, 18) ^ (W1 >> 3);
    constexpr uint32_t s1_W15 = c_rotr(W15, 17) ^ c_rotr(W15, 19) ^ (W15 >> 10);
    constexpr uint32_t s0_W15 = c_rotr(W15, 7) ^ c_rotr(W15, 18) ^ (W15 >> 3);
    constexpr uint32_t W17 = s1_W15 + W1;                                        // sigma1(W15) + W10 + sigma0(W2) + W1
    constexpr uint32_t s1_W17 = c_rotr(W17, 17) ^ c_rotr(W17, 19) ^ (W17 >> 10);
    constexpr uint32_t W19 = s1_W17;                                             // + W12 + sigma0(W4) + W3 = 0
    constexpr uint32_t s1_W19 = c_rotr(W19, 17) ^ c_rotr(W19, 19) ^ (W19 >> 10);
    constexpr uint32_t W21 = s1_W19;
    constexpr uint32_t s1_W21 = c_rotr(W21, 17) ^ c_rotr(W21, 19) ^ (W21 >> 10);
    const uint32_t w16 = addv<V, 2>(v, s0_W1);                                   // sigma1(W14) + W9 + sigma0(W1) + W0
    const uint32_t w18 = ssig1<V>(w16);                                          // + W11 + sigma0(W3) + W2 = 0
    const uint32_t w20 = ssig1<V>(w18);
    const uint32_t w22 = addv<V, 2>(ssig1<V>(w20), W15);                         // + sigma0(W7) + W6 = 0
    const uint32_t w23 = addv<V, 2>(w16, s1_W21);                                // sigma1(W21) + W16 + sigma0(W8) + W7
    const uint32_t w24 = addv<V, 2>(ssig1<V>(w22), W17);                         // + sigma0(W9) + W8 = 0
    const uint32_t w25 = addv<V, 2>(ssig1<V>(w23), w18);
    const uint32_t w26 = addv<V, 2>(ssig1<V>(w24), W19);
    const uint32_t w27 = addv<V, 2>(ssig1<V>(w25), w20);
    const uint32_t w28 = addv<V, 2>(ssig1<V>(w26), W21);
    const uint32_t w29 = addv<V, 2>(ssig1<V>(w27), w22);
    const uint32_t w30 = addv<V, 2>(addv<V, 2>(ssig1<V>(w28), w23), s0_W15);     // + W14 = 0
    const uint32_t w31 = addv<V, 2>(addv<V, 2>(addv<V, 2>(ssig1<V>(w29), w24), ssig0<V>(w16)), W15);
    w[0] = w16; w[1] = W17; w[2] = w18; w[3] = W19; w[4] = w20; w[5] = W21; w[6] = w22; w[7] = w23;
    w[8] = w24; w[9] = w25; w[10] = w26; w[11] = w27; w[12] = w28; w[13] = w29; w[14] = w30; w[15] = w31;
}

// st: out only (the chaining value is always the IV).  leaf: w[0] = value, rest ignored; node: w = 16 data words.
template <int V>
__device__ __forceinline__ void compress_leaf_or_node(uint32_t st[8], uint32_t w[16], bool leaf)
{
    constexpr Consts C = make_consts();
    uint32_t a = ZK_SHA_IV0, b = ZK_SHA_IV1, c = ZK_SHA_IV2, d = ZK_SHA_IV3, e = ZK_SHA_IV4, f = ZK_SHA_IV5,
             g = ZK_SHA_IV6, h = ZK_SHA_IV7;
    int i = 0;
    if (leaf) {
        const uint32_t v = w[0];
        w[1] = 0x80000000u;
#pragma unroll
        for (int j = 2; j < 15; j++) w[j] = 0u;
        w[15] = 32u;
#define ZK_KW_LEAF(j) (addv<V, 32>(w[j], C.k[j]))
        ZK_SHA_16ROUNDS(V, ZK_KW_LEAF)
#undef ZK_KW_LEAF
        sched_leaf<V>(v, w);
        i = 16;
    }
#pragma unroll 1
    for (;;) {
        uint32_t kk[16];
        {
            const uint4* kp = reinterpret_cast<const uint4*>(c_k + i);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 t = kp[q];
                kk[4 * q] = t.x; kk[4 * q + 1] = t.y; kk[4 * q + 2] = t.z; kk[4 * q + 3] = t.w;
            }
        }
#define ZK_KW_LOOP(j) (addv<V, 32>(w[j], kk[j]))
        ZK_SHA_16ROUNDS(V, ZK_KW_LOOP)
#undef ZK_KW_LOOP
        i += 16;
        if (i == 64) break;
#pragma unroll
        for (int j = 0; j < 16; j++)
            w[j] = sched_word<V>(w[j], w[(j + 1) & 15], w[(j + 9) & 15], w[(j + 14) & 15]);
    }
    st[0] = a + ZK_SHA_IV0; st[1] = b + ZK_SHA_IV1; st[2] = c + ZK_SHA_IV2; st[3] = d + ZK_SHA_IV3;
    st[4] = e + ZK_SHA_IV4; st[5] = f + ZK_SHA_IV5; st[6] = g + ZK_SHA_IV6; st[7] = h + ZK_SHA_IV7;
}

__device__ __forceinline__ void init_iv(uint32_t st[8])
{
    st[0] = ZK_SHA_IV0; st[1] = ZK_SHA_IV1; st[2] = ZK_SHA_IV2; st[3] = ZK_SHA_IV3;
    st[4] = ZK_SHA_IV4; st[5] = ZK_SHA_IV5; st[6] = ZK_SHA_IV6; st[7] = ZK_SHA_IV7;
}

// merkle.rs:30-34: digest of the 4-byte big-endian encoding of v.  The BE bytes of a u32
// read back as a BE message word are the number itself: W0 = v.
template <int V = 0>
__device__ __forceinline__ void hash_leaf(uint32_t v, uint32_t out[8])
{
    uint32_t w[16];
    w[0] = v; w[1] = 0x80000000u;
#pragma unroll
    for (int i = 2; i < 15; i++) w[i] = 0u;
    w[15] = 32u;
    init_iv(out);
    compress<V>(out, w);
}

// merkle.rs:40-46: digest of left || right (two native-word digests).
template <int V = 0>
__device__ __forceinline__ void hash_node(const uint32_t l[8], const uint32_t r[8], uint32_t out[8])
{
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = l[i]; w[8 + i] = r[i]; }
    init_iv(out);
    compress<V>(out, w);
    compress_pad64<V>(out);
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// Digests live in global memory as the 32 raw bytes sha2 emits (merkle.rs:9): word i
// big-endian.  32-byte aligned, moved as two 128-bit accesses.
__device__ __forceinline__ void store_digest(uint8_t* dst, const uint32_t d[8])
{
    uint4* p = reinterpret_cast<uint4*>(dst);
    p[0] = make_uint4(bswap32(d[0]), bswap32(d[1]), bswap32(d[2]), bswap32(d[3]));
    p[1] = make_uint4(bswap32(d[4]), bswap32(d[5]), bswap32(d[6]), bswap32(d[7]));
}
__device__ __forceinline__ void load_digest(const uint8_t* src, uint32_t d[8])
{
    const uint4* p = reinterpret_cast<const uint4*>(src);
    uint4 a = p[0], b = p[1];
    d[0] = bswap32(a.x); d[1] = bswap32(a.y); d[2] = bswap32(a.z); d[3] = bswap32(a.w);
    d[4] = bswap32(b.x); d[5] = bswap32(b.y); d[6] = bswap32(b.z); d[7] = bswap32(b.w);
}

// same, through L2 (ld.global.cg): for digests written by other CTAs of the running launch
__device__ __forceinline__ void load_digest_cg(const uint8_t* src, uint32_t d[8])
{
    const uint4* p = reinterpret_cast<const uint4*>(src);
    uint4 a = __ldcg(p), b = __ldcg(p + 1);
    d[0] = bswap32(a.x); d[1] = bswap32(a.y); d[2] = bswap32(a.z); d[3] = bswap32(a.w);
    d[4] = bswap32(b.x); d[5] = bswap32(b.y); d[6] = bswap32(b.z); d[7] = bswap32(b.w);
}

#endif  // __CUDACC__ && ZK_SHA_DEVICE

// ---- host SHA-256 (generic length) for the host-side tail of the channel
void host_sha256(const uint8_t* data, size_t len, uint8_t out[32]);

}  // namespace sha
}  // namespace zk
