/*
 * oracle/poly.c -- CPU ORACLE (test infrastructure).
 * LITERAL restatement of polynomial.rs: the reference's own schoolbook algorithms
 * (per-term pow in solve, prefix/suffix-product Lagrange, one-term-per-step long
 * division with a full schoolbook product per step, coefficient FRI fold).
 * Coefficient arrays are ascending, as Polynomial<T>.0 is (polynomial.rs:31-41).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static inline uint32_t mulm(uint32_t a, uint32_t b, uint32_t p)
{
    if (p == ORC_P) return (uint32_t)(((uint64_t)a * b) % ORC_P); /* constant-modulus fast path */
    return (uint32_t)(((uint64_t)a * b) % p);
}
static inline uint32_t addm(uint32_t a, uint32_t b, uint32_t p)
{
    uint64_t s = (uint64_t)a + b;
    return (uint32_t)(s >= p ? s - p : s);
}
static inline uint32_t subm(uint32_t a, uint32_t b, uint32_t p)
{
    return a >= b ? a - b : (uint32_t)((uint64_t)a + p - b);
}
static inline uint32_t powm(uint32_t a, uint32_t e, uint32_t p)
{
    uint32_t r = 1 % p;
    while (e) {
        if (e & 1) r = mulm(r, a, p);
        a = mulm(a, a, p);
        e >>= 1;
    }
    return r;
}

/* reduce(): polynomial.rs:15-28 -- strip trailing (high-degree) zeros */
static uint32_t trim(const uint32_t *c, uint32_t len)
{
    while (len && c[len - 1] == 0) len--;
    return len;
}

/* polynomial.rs:49-56: sum_i c_i * t^i with a fresh pow per term */
uint32_t orc_poly_solve(const uint32_t *c, uint32_t len, uint32_t t, uint32_t p)
{
    uint32_t acc = 0;
    for (uint32_t d = 0; d < len; d++) acc = addm(acc, mulm(c[d], powm(t, d, p), p), p);
    return acc;
}

/* polynomial.rs:229-258: schoolbook product, NOT trimmed.  out has la+lb-1 slots. */
static void mul_school(const uint32_t *a, uint32_t la, const uint32_t *b, uint32_t lb,
                       uint32_t *out, uint32_t p)
{
    memset(out, 0, sizeof(uint32_t) * (la + lb - 1));
    for (uint32_t i = 0; i < la; i++)
        for (uint32_t j = 0; j < lb; j++)
            out[i + j] = addm(out[i + j], mulm(a[i], b[j], p), p);
}

/* polynomial.rs:337-383 */
uint32_t orc_lagrange(const uint32_t *xs, const uint32_t *ys, uint32_t n, uint32_t *out, uint32_t p)
{
    /* ll[i] has i+1 coefficients, lr[i] has n-i (polynomial.rs:353-362) */
    uint32_t **ll = malloc(sizeof(*ll) * n), **lr = malloc(sizeof(*lr) * n);
    uint32_t lin[2];
    for (uint32_t i = 0; i < n; i++) {
        ll[i] = malloc(sizeof(uint32_t) * (i + 1));
        lr[i] = malloc(sizeof(uint32_t) * (n - i));
    }
    ll[0][0] = 1 % p;
    for (uint32_t i = 1; i < n; i++) {
        lin[0] = subm(0, xs[i - 1] % p, p); lin[1] = 1 % p;   /* Polynomial::from([1, -x]) */
        mul_school(ll[i - 1], i, lin, 2, ll[i], p);
    }
    lr[n - 1][0] = 1 % p;
    for (uint32_t i = n - 1; i-- > 0;) {
        lin[0] = subm(0, xs[i + 1] % p, p); lin[1] = 1 % p;
        mul_school(lr[i + 1], n - i - 1, lin, 2, lr[i], p);
    }
    uint32_t *basis = malloc(sizeof(uint32_t) * n);
    memset(out, 0, sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; i++) {
        mul_school(ll[i], i + 1, lr[i], n - i, basis, p);          /* :365 */
        uint32_t s = orc_inv(orc_poly_solve(basis, n, xs[i], p), p); /* :369-371 */
        s = mulm(s, ys[i] % p, p);                                  /* :377-379 */
        for (uint32_t k = 0; k < n; k++) out[k] = addm(out[k], mulm(basis[k], s, p), p); /* :381 */
    }
    for (uint32_t i = 0; i < n; i++) { free(ll[i]); free(lr[i]); }
    free(ll); free(lr); free(basis);
    return trim(out, n);
}

/* polynomial.rs:307-334, recursion unrolled into a loop; every step performs the
 * reference's full (diff+1) x rhs_len schoolbook product (:327). */
int orc_poly_div(const uint32_t *lhs_in, uint32_t lhs_len, const uint32_t *rhs, uint32_t rhs_len,
                 uint32_t *q, uint32_t *q_len, uint32_t *r, uint32_t *r_len, uint32_t p)
{
    uint32_t rhs_deg = rhs_len ? rhs_len - 1 : 0;
    uint32_t qcap = lhs_len >= rhs_len && rhs_len ? lhs_len - rhs_len + 1 : 0;
    uint32_t *lhs = malloc(sizeof(uint32_t) * (lhs_len ? lhs_len : 1));
    uint32_t *dv = malloc(sizeof(uint32_t) * (qcap ? qcap : 1));
    uint32_t *prod = malloc(sizeof(uint32_t) * (lhs_len ? lhs_len : 1));
    memcpy(lhs, lhs_in, sizeof(uint32_t) * lhs_len);
    memset(q, 0, sizeof(uint32_t) * qcap);
    uint32_t ql = 0, len = lhs_len;
    int rc = 0;
    for (;;) {
        uint32_t lhs_deg = len ? len - 1 : 0;        /* degree().unwrap_or(0) :309 */
        if (lhs_deg < rhs_deg) break;                /* :313-315 */
        if (len == 0 || rhs_len == 0) { rc = -1; break; } /* lhs[0] would panic (:318) */
        uint32_t lead = mulm(lhs[len - 1], orc_inv(rhs[rhs_len - 1], p), p); /* :324 */
        uint32_t diff = lhs_deg - rhs_deg;
        memset(dv, 0, sizeof(uint32_t) * (diff + 1));
        dv[diff] = lead;
        mul_school(dv, diff + 1, rhs, rhs_len, prod, p);       /* &div * &rhs (:327) */
        for (uint32_t k = 0; k < len; k++) lhs[k] = subm(lhs[k], prod[k], p);
        uint32_t nl = trim(lhs, len);                           /* Sub trims (:225) */
        q[diff] = addm(q[diff], lead, p);                       /* div + q (:333) */
        if (diff + 1 > ql) ql = diff + 1;
        if (nl == len) { rc = -2; break; }                      /* would recurse forever */
        len = nl;
        if (rhs_deg == 0 && len == 0) break;                    /* next step would panic in the reference */
    }
    memcpy(r, lhs, sizeof(uint32_t) * len);
    *r_len = len;
    *q_len = trim(q, ql);
    free(lhs); free(dv); free(prod);
    return rc;
}

/* polynomial.rs:385-400 */
void orc_fri_coeff(const uint32_t *c, uint32_t len, uint32_t beta, uint32_t *out, uint32_t p)
{
    for (uint32_t i = 0; i < len / 2; i++)
        out[i] = addm(c[2 * i], mulm(beta % p, c[2 * i + 1], p), p);
}
