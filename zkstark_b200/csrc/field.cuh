// K1 -- prime-field arithmetic for p = 3*2^30 + 1 = 0xC0000001 on sm_100a.
//
// Replaces field.rs:96-211 (Gf<P> over num_modular::MontgomeryInt<u32>).  Only canonical
// residues in [0, p) are observable in the reference (field.rs:41-43 `residue`, 20-24
// `From<u32>`), so every value that reaches global memory here is canonical.
//
// p > 2^31, so a+b overflows 32 bits and the usual lazy [0,2p) tricks do not fit; and
// t + m*p overflows 64 bits in textbook REDC.  We use the subtractive Montgomery form
//     mont_mul(a,b) = (a*b - m*p) / 2^32,  m = (a*b mod 2^32) * p^-1 mod 2^32
// whose low words cancel exactly, leaving hi(a*b) - hi(m*p) in (-p, p): one borrow fix.
// Data stays canonical; constants (twiddles, challenges, table entries) are kept in
// Montgomery form xR so that mont_mul(x, cR) = x*c mod p is canonical again.
#pragma once
#include <stdint.h>

namespace zk {

constexpr uint32_t P = 0xC0000001u;      // main.rs:13
constexpr uint32_t P_INV = 0x40000001u;  // p^-1 mod 2^32  (= 2^30 + 1)
constexpr uint32_t R1 = 1073741823u;     // 2^32 mod p   (Montgomery form of 1)
constexpr uint32_t R2 = 1789569709u;     // 2^64 mod p   (to_mont multiplier)
constexpr uint32_t GEN = 5u;             // F::generator(), prover.rs:44-45

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#else
#define ZK_HD inline
#endif

// r = a - b and whether it borrowed.  On the device the subtraction is written with the carry flag
// (sub.cc / subc): ptxas then folds the comparison into the IADD3's carry-out predicate, so
// "subtract, add p back on borrow" is 2 instructions instead of ISETP + IADD + predicated add.
ZK_HD uint32_t sub_borrow(uint32_t a, uint32_t b, uint32_t& borrow)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("{sub.cc.u32 %0, %2, %3; subc.u32 %1, 0, 0;}" : "=r"(r), "=r"(borrow) : "r"(a), "r"(b));
    return r;
#else
    borrow = a < b;
    return a - b;
#endif
}

ZK_HD uint32_t fsub(uint32_t a, uint32_t b)  // field.rs:117-122
{
    uint32_t bw;
    uint32_t r = sub_borrow(a, b, bw);
    return bw ? r + P : r;
}

ZK_HD uint32_t fadd(uint32_t a, uint32_t b)  // field.rs:96-101
{
    // a + b == a - (p - b); borrow <=> a + b < p
    return fsub(a, P - b);
}

ZK_HD uint32_t fneg(uint32_t a) { return a ? P - a : 0u; }  // field.rs:159-164

ZK_HD uint32_t mont_mul(uint32_t a, uint32_t b)  // field.rs:138-143 (one operand in Montgomery form)
{
    uint64_t t = (uint64_t)a * b;
    uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
    uint32_t m = lo * P_INV;
#if defined(__CUDA_ARCH__)
    uint32_t u = __umulhi(m, P);
#else
    uint32_t u = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
    uint32_t bw;
    uint32_t r = sub_borrow(hi, u, bw);
    return bw ? r + P : r;
}

ZK_HD uint32_t to_mont(uint32_t a) { return mont_mul(a, R2); }    // a must be < p
ZK_HD uint32_t from_mont(uint32_t a) { return mont_mul(a, 1u); }
ZK_HD uint32_t reduce_u32(uint32_t a) { return a >= P ? a - P : a; }  // field.rs:20-24 From<u32>

// Montgomery-form exponentiation: base and result in Montgomery form.
ZK_HD uint32_t mont_pow(uint32_t base_m, uint64_t e)  // field.rs:26-38
{
    uint32_t r = R1;
    while (e) {
        if (e & 1) r = mont_mul(r, base_m);
        base_m = mont_mul(base_m, base_m);
        e >>= 1;
    }
    return r;
}
ZK_HD uint32_t mont_inv(uint32_t a_m) { return mont_pow(a_m, (uint64_t)P - 2); }  // field.rs:206-211

// ---- plain host helpers (canonical in, canonical out) used to build constant tables
inline uint32_t h_mul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
inline uint32_t h_pow(uint32_t a, uint64_t e)
{
    uint32_t r = 1;
    a %= P;
    while (e) {
        if (e & 1) r = h_mul(r, a);
        a = h_mul(a, a);
        e >>= 1;
    }
    return r;
}
inline uint32_t h_inv(uint32_t a) { return h_pow(a, (uint64_t)P - 2); }
// generator of the order-2^log subgroup: 5^((p-1)/2^log)  (prover.rs:48-49)
inline uint32_t h_root(uint32_t log) { return h_pow(GEN, ((uint64_t)P - 1) >> log); }
inline uint32_t h_to_mont(uint32_t a) { return (uint32_t)((((uint64_t)a) << 32) % P); }

}  // namespace zk
