/*
 * oracle/field.c -- CPU ORACLE (test infrastructure).
 * Semantics of field.rs: Gf<P> wraps num_modular::MontgomeryInt<u32>; only the
 * canonical residue is observable (field.rs:41-43), so plain arithmetic mod P
 * is an exact restatement.
 */
#include "oracle.h"

uint32_t orc_add(uint32_t a, uint32_t b, uint32_t p) /* field.rs:96-101 */
{
    return (uint32_t)(((uint64_t)a + b) % p);
}

uint32_t orc_sub(uint32_t a, uint32_t b, uint32_t p) /* field.rs:117-122 */
{
    return (uint32_t)(((uint64_t)a + p - (b % p)) % p);
}

uint32_t orc_mul(uint32_t a, uint32_t b, uint32_t p) /* field.rs:138-143 */
{
    return (uint32_t)(((uint64_t)a * b) % p);
}

uint32_t orc_pow(uint32_t a, uint32_t e, uint32_t p) /* field.rs:26-38 */
{
    uint64_t r = 1 % p, b = a % p;
    while (e) {
        if (e & 1) r = r * b % p;
        b = b * b % p;
        e >>= 1;
    }
    return (uint32_t)r;
}

uint32_t orc_inv(uint32_t a, uint32_t p) /* field.rs:206-211 (P prime) */
{
    return orc_pow(a, p - 2, p);
}

uint32_t orc_from_i32(int32_t v, uint32_t p) /* field.rs:10-18 */
{
    if (v < 0) {
        uint32_t m = (uint32_t)(-(int64_t)v) % p;
        return m ? p - m : 0;
    }
    return (uint32_t)v % p;
}

uint32_t orc_order(uint32_t a, uint32_t p) /* field.rs:45-49 (brute force) */
{
    uint64_t x = a % p;
    for (uint32_t it = 1;; it++) {
        if (x == 1) return it;
        x = x * a % p;
    }
}

uint32_t orc_generator(uint32_t p) /* field.rs:52-86 */
{
    uint32_t factors[32], nf = 0, q = p - 1, it = 2;
    while (q != 1) {
        if (q % it == 0) factors[nf++] = it;
        while (q % it == 0) q /= it;
        it++;
    }
    for (uint32_t x = 2; x < p; x++) {
        int ok = 1;
        for (uint32_t i = 0; i < nf; i++) {
            /* field.rs:71: (P-1) * factor^-1 in F_P, which equals the integer (P-1)/factor */
            uint32_t e = orc_mul(p - 1, orc_inv(factors[i], p), p);
            if (orc_pow(x, e, p) == 1) { ok = 0; break; }
        }
        if (ok) return x;
    }
    return 0;
}
