/*
 * oracle/fast.c -- CPU ORACLE (test infrastructure), FAST route.
 * Computes the same values as the reference's schoolbook code, by NTT:
 *   lagrange over the subgroup (prover.rs:60-61 -> polynomial.rs:337)  == iNTT with a phantom point
 *   f_poly.solve over the coset (prover.rs:69-70)                      == coset NTT
 *   the constraint quotients + cp_poly.solve (prover.rs:101-173)       == pointwise rational formula
 *   fri() + solve on the squared half-domain (polynomial.rs:385, prover.rs:204-211)
 *                                                                     == evaluation-form fold (proof.rs:110-112)
 * Equality with the literal route is checked byte-for-byte in tests/ (log_n 4..10).
 * Multithreaded with OpenMP: this is also the CPU baseline bench.py times.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define P ORC_P
#define NT(threads) num_threads((threads) > 0 ? (threads) : 1)

static inline uint32_t mulp(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
static inline uint32_t addp(uint32_t a, uint32_t b)
{
    uint64_t s = (uint64_t)a + b;
    return (uint32_t)(s >= P ? s - P : s);
}
static inline uint32_t subp(uint32_t a, uint32_t b) { return a >= b ? a - b : (uint32_t)((uint64_t)a + P - b); }
static inline uint32_t powp(uint32_t a, uint64_t e)
{
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = mulp(r, a);
        a = mulp(a, a);
        e >>= 1;
    }
    return r;
}
static inline uint32_t invp(uint32_t a) { return powp(a, P - 2); }
/* generator of the order-2^log subgroup: 5^((p-1)/2^log)  (prover.rs:45-49) */
static inline uint32_t root_of_unity(uint32_t log) { return powp(5, ((uint64_t)P - 1) >> log); }

/* prover.rs:32-39 */
void orc_trace_fibsq(uint32_t log_n, uint32_t secret, uint32_t *a)
{
    uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    a[0] = 1;
    if (n1 > 1) a[1] = secret % P;
    for (uint64_t i = 2; i < n1; i++) a[i] = addp(mulp(a[i - 2], a[i - 2]), mulp(a[i - 1], a[i - 1]));
}
/* the recurrence of an orc_air at (x0, x1) = (a[i], a[i+1]) */
uint32_t orc_air_step(const orc_air *air, uint32_t x0, uint32_t x1)
{
    uint32_t r = mulp(air->c20 % P, mulp(x1, x1));
    r = addp(r, mulp(air->c02 % P, mulp(x0, x0)));
    r = addp(r, mulp(air->c11 % P, mulp(x0, x1)));
    r = addp(r, mulp(air->c10 % P, x1));
    r = addp(r, mulp(air->c01 % P, x0));
    return addp(r, air->c00 % P);
}
void orc_trace_air(const orc_air *air, uint32_t log_n, uint32_t secret, uint32_t *a)
{
    uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    a[0] = air->a0 % P;
    if (n1 > 1) a[1] = secret % P;
    for (uint64_t i = 2; i < n1; i++) a[i] = orc_air_step(air, a[i - 2], a[i - 1]);
}

/* in-place radix-2 NTT, natural order in and out: x[k] <- sum_j x[j] w^(jk) */
static void ntt(uint32_t *x, uint32_t log_n, uint32_t w, int threads)
{
    uint64_t n = (uint64_t)1 << log_n;
    if (n == 1) return;
    uint32_t *tw = malloc(sizeof(uint32_t) * (n / 2));
    /* powers of w, built in parallel blocks */
    {
        uint64_t half = n / 2, blk = 4096;
#pragma omp parallel for schedule(static) NT(threads)
        for (int64_t b = 0; b < (int64_t)((half + blk - 1) / blk); b++) {
            uint64_t s = (uint64_t)b * blk, e = s + blk < half ? s + blk : half;
            uint32_t c = powp(w, s);
            for (uint64_t i = s; i < e; i++) { tw[i] = c; c = mulp(c, w); }
        }
    }
    /* bit reversal */
#pragma omp parallel for schedule(static) NT(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint64_t r = 0, v = (uint64_t)i;
        for (uint32_t b = 0; b < log_n; b++) { r = (r << 1) | (v & 1); v >>= 1; }
        if (r > (uint64_t)i) { uint32_t t = x[i]; x[i] = x[r]; x[r] = t; }
    }
    for (uint32_t s = 0; s < log_n; s++) {
        uint64_t m = (uint64_t)1 << s, stride = n >> (s + 1);
#pragma omp parallel for schedule(static) NT(threads)
        for (int64_t j = 0; j < (int64_t)(n / 2); j++) {
            uint64_t grp = (uint64_t)j >> s, pos = (uint64_t)j & (m - 1);
            uint64_t i0 = (grp << (s + 1)) + pos, i1 = i0 + m;
            uint32_t u = x[i0], v = mulp(x[i1], tw[pos * stride]);
            x[i0] = addp(u, v);
            x[i1] = subp(u, v);
        }
    }
    free(tw);
}

/* Unique interpolant of (g^i, y_i).  With n_points == n-1 (prover.rs:60: zip truncates to the
 * 1023 trace values) the degree must be <= n-2: append the phantom value
 *   y_{n-1} = -g * sum_{i<n-1} y_i g^i
 * which makes the x^(n-1) coefficient of the size-n inverse transform vanish. */
int orc_interpolate_subgroup(const uint32_t *y, uint32_t n_points, uint32_t log_n,
                             uint32_t *c, int threads)
{
    uint64_t n = (uint64_t)1 << log_n;
    if (log_n > 30 || (n_points != n && n_points != n - 1)) return -1;
    uint32_t g = root_of_unity(log_n);
    memcpy(c, y, sizeof(uint32_t) * n_points);
    if (n_points == n - 1) {
        uint32_t acc = 0, gi = 1;
        for (uint64_t i = 0; i < n - 1; i++) { acc = addp(acc, mulp(y[i], gi)); gi = mulp(gi, g); }
        c[n - 1] = subp(0, mulp(g, acc));
    }
    ntt(c, log_n, invp(g), threads);
    uint32_t ninv = invp((uint32_t)(n % P));
#pragma omp parallel for schedule(static) NT(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) c[i] = mulp(c[i], ninv);
    return 0;
}

/* prover.rs:69-70: evals[i] = f(shift * h^i) */
int orc_lde_coset(const uint32_t *coeffs, uint32_t log_n, uint32_t log_blowup, uint32_t shift,
                  uint32_t *out, int threads)
{
    uint32_t log_N = log_n + log_blowup;
    if (log_N > 30) return -1;
    uint64_t n = (uint64_t)1 << log_n, N = (uint64_t)1 << log_N, blk = 4096;
    shift %= P;
    memset(out, 0, sizeof(uint32_t) * N);
#pragma omp parallel for schedule(static) NT(threads)
    for (int64_t b = 0; b < (int64_t)((n + blk - 1) / blk); b++) {
        uint64_t s = (uint64_t)b * blk, e = s + blk < n ? s + blk : n;
        uint32_t sk = powp(shift, s);
        for (uint64_t k = s; k < e; k++) { out[k] = mulp(coeffs[k], sk); sk = mulp(sk, shift); }
    }
    ntt(out, log_N, root_of_unity(log_N), threads);
    return 0;
}

/* prover.rs:101-173, pointwise at x = 5 h^i:
 *  cp = a0 (f(x)-a_first)/(x-1) + a1 (f(x)-a_last)/(x-g^(n-2))
 *     + a2 (f(g^2 x) - f(g x)^2 - f(x)^2) (x-g^(n-3))(x-g^(n-2))(x-g^(n-1)) / (x^n-1)   */
int orc_cp_eval_fibsq(const uint32_t *f, uint32_t log_n, uint32_t log_blowup,
                      const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last,
                      uint32_t *cp, int threads)
{
    const orc_air fibsq = ORC_AIR_FIBSQ;
    return orc_cp_eval_air(&fibsq, f, log_n, log_blowup, alpha_raw, a_first, a_last, cp, threads);
}
/* the same with the transition numerator f(g^2 x) - step(f(x), f(g x)) of an orc_air */
int orc_cp_eval_air(const orc_air *air, const uint32_t *f, uint32_t log_n, uint32_t log_blowup,
                    const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last,
                    uint32_t *cp, int threads)
{
    uint32_t log_N = log_n + log_blowup;
    if (log_N > 30 || log_n < 2) return -1;
    uint64_t n = (uint64_t)1 << log_n, N = (uint64_t)1 << log_N, B = (uint64_t)1 << log_blowup;
    uint32_t h = root_of_unity(log_N), g = root_of_unity(log_n);
    uint32_t a0 = alpha_raw[0] % P, a1 = alpha_raw[1] % P, a2 = alpha_raw[2] % P; /* field.rs:20-24 */
    uint32_t g1 = powp(g, n - 1), g2 = powp(g, n - 2), g3 = powp(g, n - 3);
    uint32_t wn = powp(5, n), hn = powp(h, n);
    a_first %= P; a_last %= P;
    const int fibsq = air->c20 == 1 && air->c02 == 1 && !air->c11 && !air->c10 && !air->c01 && !air->c00;
    enum { CH = 1024 };
    uint64_t nch = (N + CH - 1) / CH;
#pragma omp parallel for schedule(static) NT(threads)
    for (int64_t c = 0; c < (int64_t)nch; c++) {
        uint64_t s = (uint64_t)c * CH, e = s + CH < N ? s + CH : N, m = e - s;
        uint32_t xs[CH], den[3 * CH], pre[3 * CH];
        uint32_t x = mulp(5, powp(h, s)), xn = mulp(wn, powp(hn, s));
        for (uint64_t k = 0; k < m; k++) {
            xs[k] = x;
            den[3 * k] = subp(x, 1);
            den[3 * k + 1] = subp(x, g2);
            den[3 * k + 2] = subp(xn, 1);
            x = mulp(x, h); xn = mulp(xn, hn);
        }
        /* batch inversion (no denominator vanishes on the coset) */
        uint32_t acc = 1;
        for (uint64_t k = 0; k < 3 * m; k++) { pre[k] = acc; acc = mulp(acc, den[k]); }
        acc = invp(acc);
        for (uint64_t k = 3 * m; k-- > 0;) { uint32_t d = den[k]; den[k] = mulp(acc, pre[k]); acc = mulp(acc, d); }
        for (uint64_t k = 0; k < m; k++) {
            uint64_t i = s + k;
            uint32_t fx = f[i], fgx = f[(i + B) & (N - 1)], fggx = f[(i + 2 * B) & (N - 1)];
            uint32_t xx = xs[k];
            uint32_t p0 = mulp(subp(fx, a_first), den[3 * k]);
            uint32_t p1 = mulp(subp(fx, a_last), den[3 * k + 1]);
            uint32_t num = fibsq ? subp(subp(fggx, mulp(fgx, fgx)), mulp(fx, fx))       /* prover.rs:134-138 */
                                 : subp(fggx, orc_air_step(air, fx, fgx));
            uint32_t cub = mulp(mulp(subp(xx, g3), subp(xx, g2)), subp(xx, g1));
            uint32_t p2 = mulp(mulp(num, cub), den[3 * k + 2]);
            cp[i] = addp(addp(mulp(a0, p0), mulp(a1, p1)), mulp(a2, p2));
        }
    }
    return 0;
}

/* next[i] = (cur[i]+cur[i+m])/2 + beta (cur[i]-cur[i+m]) / (2 x_i),  x_i = (5 h0^i)^(2^k),
 * m = len/2, h0 of order len<<k.  Same values as fri() then solve (polynomial.rs:402-426). */
int orc_fri_fold(const uint32_t *cur, uint64_t len, uint32_t k, uint32_t beta_raw,
                 uint32_t *next, int threads)
{
    if (len < 2 || (len & (len - 1))) return -1;
    uint32_t log_len = 0;
    while (((uint64_t)1 << log_len) < len) log_len++;
    if (log_len + k > 30) return -1;
    uint64_t m = len / 2, blk = 4096;
    uint32_t beta = beta_raw % P, inv2 = invp(2);
    uint32_t hk = root_of_unity(log_len);            /* h0^(2^k) has order len */
    uint32_t hk_inv = invp(hk);
    uint32_t wk_inv = invp(powp(5, (uint64_t)1 << k));
#pragma omp parallel for schedule(static) NT(threads)
    for (int64_t b = 0; b < (int64_t)((m + blk - 1) / blk); b++) {
        uint64_t s = (uint64_t)b * blk, e = s + blk < m ? s + blk : m;
        uint32_t xinv = mulp(wk_inv, powp(hk_inv, s));
        for (uint64_t i = s; i < e; i++) {
            uint32_t u = cur[i], v = cur[i + m];
            uint32_t ev = mulp(addp(u, v), inv2);
            uint32_t od = mulp(mulp(subp(u, v), inv2), xinv);
            next[i] = addp(ev, mulp(beta, od));
            xinv = mulp(xinv, hk_inv);
        }
    }
    return 0;
}
