/*
 * oracle/prover.c -- CPU ORACLE (test infrastructure).
 * Restates channel.rs (Fiat-Shamir transcript, bincode 1.3 fixint layout) and
 * prover.rs:9-293 (operation AND commit order), generalised from the reference's
 * literals (1023/1024/8192/10 layers/blow-up 8) to n = 2^log_n by SURVEY.md 8(d).
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define P ORC_P
#define LOG_BLOWUP 3u /* 8192/1024, prover.rs:56-57; literal offsets +8/+16 prover.rs:268-271 */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---------------- channel.rs ---------------- */
typedef struct {
    uint8_t state[32]; /* channel.rs:14  [0; 32] */
    uint8_t *data;
    size_t len, cap;
    int overflow;
} chan_t;

/* channel.rs:19-26: state = SHA256(state || v); data.append(v) */
static void chan_commit(chan_t *c, const uint8_t *v, size_t vlen)
{
    uint8_t stackbuf[4096], *buf = stackbuf;
    if (32 + vlen > sizeof stackbuf) buf = malloc(32 + vlen);
    memcpy(buf, c->state, 32);
    memcpy(buf + 32, v, vlen);
    orc_sha256(buf, 32 + vlen, c->state);
    if (buf != stackbuf) free(buf);
    if (c->len + vlen > c->cap) { c->overflow = 1; return; }
    memcpy(c->data + c->len, v, vlen);
    c->len += vlen;
}
static void put_u32le(uint8_t *b, uint32_t v) /* bincode fixint: u32 -> 4 bytes LE */
{
    b[0] = (uint8_t)v; b[1] = (uint8_t)(v >> 8); b[2] = (uint8_t)(v >> 16); b[3] = (uint8_t)(v >> 24);
}
static void put_u64le(uint8_t *b, uint64_t v)
{
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
}
static void chan_commit_u32(chan_t *c, uint32_t v)
{
    uint8_t b[4];
    put_u32le(b, v);
    chan_commit(c, b, 4);
}
/* channel.rs:28-32: f = be32(state[0..4]); commit(f); return f un-reduced */
static uint32_t chan_get_u32(chan_t *c)
{
    uint32_t f = ((uint32_t)c->state[0] << 24) | ((uint32_t)c->state[1] << 16) |
                 ((uint32_t)c->state[2] << 8) | c->state[3];
    chan_commit_u32(c, f);
    return f;
}
/* (u32, Box<[[u8;32]]>): 4 LE | u64 LE length | 32*len raw  (prover.rs:274-277) */
static void chan_commit_val_path(chan_t *c, uint32_t v, const uint8_t *path, uint32_t plen)
{
    uint8_t buf[12 + 32 * 40];
    put_u32le(buf, v);
    put_u64le(buf + 4, plen);
    memcpy(buf + 12, path, 32 * (size_t)plen);
    chan_commit(c, buf, 12 + 32 * (size_t)plen);
}
/* (u32, u32, path, path)  (prover.rs:288) */
static void chan_commit_pair(chan_t *c, uint32_t v0, uint32_t v1, const uint8_t *p0, uint32_t l0,
                             const uint8_t *p1, uint32_t l1)
{
    uint8_t buf[24 + 64 * 40];
    size_t o = 0;
    put_u32le(buf + o, v0); o += 4;
    put_u32le(buf + o, v1); o += 4;
    put_u64le(buf + o, l0); o += 8;
    memcpy(buf + o, p0, 32 * (size_t)l0); o += 32 * (size_t)l0;
    put_u64le(buf + o, l1); o += 8;
    memcpy(buf + o, p1, 32 * (size_t)l1); o += 32 * (size_t)l1;
    chan_commit(c, buf, o);
}

size_t orc_transcript_len(uint32_t k)
{
    /* f_root | 3 alpha | cp_root | k x (beta, root) | free | query | 4 (v, path) | k layer pairs */
    size_t len = 32 + 12 + 32 + (size_t)k * 36 + 4 + 4;
    len += 4 * (12 + 32 * (size_t)(k + LOG_BLOWUP));
    for (uint32_t i = 0; i < k; i++) len += 24 + 64 * (size_t)(k + LOG_BLOWUP - i);
    return len;
}

/* ---------------- literal helpers (polynomial.rs semantics) ---------------- */
static uint32_t trim(const uint32_t *c, uint32_t len)
{
    while (len && c[len - 1] == 0) len--;
    return len;
}
/* &a - &b with trimming (polynomial.rs:163-194); the "0 - rhs = rhs" quirk (:177) is
 * reported through *quirk so tests can assert it never fires on the path. */
static uint32_t poly_sub(const uint32_t *a, uint32_t la, const uint32_t *b, uint32_t lb,
                         uint32_t *out, int *quirk)
{
    if (la == 0 && lb == 0) return 0;
    if (la == 0) { memcpy(out, b, 4 * (size_t)lb); if (quirk) *quirk += 1; return lb; }
    if (lb == 0) { memcpy(out, a, 4 * (size_t)la); return la; }
    uint32_t l = la > lb ? la : lb;
    for (uint32_t i = 0; i < l; i++) {
        uint32_t x = i < la ? a[i] : 0, y = i < lb ? b[i] : 0;
        out[i] = orc_sub(x, y, P);
    }
    return trim(out, l);
}
static void poly_mul(const uint32_t *a, uint32_t la, const uint32_t *b, uint32_t lb, uint32_t *out)
{
    memset(out, 0, 4 * (size_t)(la + lb - 1));
    for (uint32_t i = 0; i < la; i++)
        for (uint32_t j = 0; j < lb; j++)
            out[i + j] = orc_add(out[i + j], orc_mul(a[i], b[j], P), P);
}

int orc_prove_fibsq(uint32_t k, uint32_t secret, int route, int threads, uint8_t *tr, size_t cap,
                    size_t *len_out, uint8_t state_out[32], orc_prove_info *info)
{
    const orc_air fibsq = ORC_AIR_FIBSQ;
    return orc_prove_air(&fibsq, k, secret, route, threads, tr, cap, len_out, state_out, info);
}

/* generate_proof for a program of the orc_air family; at ORC_AIR_FIBSQ this IS prover.rs:9-293 */
int orc_prove_air(const orc_air *air, uint32_t k, uint32_t secret, int route, int threads, uint8_t *tr, size_t cap,
                  size_t *len_out, uint8_t state_out[32], orc_prove_info *info)
{
    const int fib_shape = air->c20 == 1 && air->c02 == 1 && !air->c11 && !air->c10 && !air->c01 && !air->c00;
    const int is_fibsq = fib_shape && air->a0 == 1;
    if (k < 2 || k + LOG_BLOWUP > 30) return -1;
    if (route == ORC_ROUTE_LITERAL && k > 11) return -2; /* O(n^3) */
    const uint32_t n = 1u << k, N = n << LOG_BLOWUP, B = 1u << LOG_BLOWUP, L = k;
    const uint32_t plen0 = k + LOG_BLOWUP;
    int rc = 0, quirk = 0;
    double t0 = now_s(), t;
    orc_prove_info local;
    if (!info) info = &local;
    memset(info, 0, sizeof *info);
    info->log_n = k;

    chan_t ch;
    memset(&ch, 0, sizeof ch);
    ch.data = tr; ch.cap = cap;

    /* ---- Part 1 (prover.rs:31-85) ---- */
    uint32_t *a = malloc(4 * (size_t)n);
    t = now_s();
    if (is_fibsq) orc_trace_fibsq(k, secret, a);
    else orc_trace_air(air, k, secret, a);
    info->t_trace = now_s() - t;
    info->a_last = a[n - 2];

    const uint32_t w = 5; /* F::generator(), prover.rs:45 */
    const uint32_t gen_g = orc_pow(w, (P - 1) / n, P), gen_h = orc_pow(w, (P - 1) / N, P); /* :48-49 */
    uint32_t *f_poly = calloc(n, 4), *f_eval = malloc(4 * (size_t)N);
    uint32_t *g = NULL, *dom = NULL;
    uint32_t f_len = n - 1;

    if (route == ORC_ROUTE_LITERAL) {
        g = malloc(4 * (size_t)n);
        dom = malloc(4 * (size_t)N);
        for (uint32_t i = 0; i < n; i++) g[i] = orc_pow(gen_g, i, P);                     /* :56 */
        for (uint32_t i = 0; i < N; i++) dom[i] = orc_mul(w, orc_pow(gen_h, i, P), P);     /* :57,69 */
        t = now_s();
        f_len = orc_lagrange(g, a, n - 1, f_poly, P);                                      /* :60-61 */
        for (uint32_t i = 0; i < n - 1; i++)                                               /* :64-66 */
            if (orc_poly_solve(f_poly, f_len, g[i], P) != a[i]) { rc = -10; goto done1; }
        info->t_interp = now_s() - t;
        t = now_s();
        for (uint32_t i = 0; i < N; i++) f_eval[i] = orc_poly_solve(f_poly, f_len, dom[i], P); /* :70 */
        info->t_lde = now_s() - t;
    } else {
        t = now_s();
        orc_interpolate_subgroup(a, n - 1, k, f_poly, threads);
        info->t_interp = now_s() - t;
        t = now_s();
        orc_lde_coset(f_poly, k, LOG_BLOWUP, w, f_eval, threads);
        info->t_lde = now_s() - t;
    }

    /* Merkle heaps: f, cp (= FRI layer 0) and FRI layers 1..L */
    uint8_t *f_heap = malloc(32 * (2 * (size_t)N - 1));
    uint8_t **heaps = calloc(L + 1, sizeof *heaps);
    uint32_t **evals = calloc(L + 1, sizeof *evals);
    t = now_s();
    orc_merkle_build(f_eval, N, f_heap, threads);                                          /* :81 */
    info->t_merkle += now_s() - t;
    memcpy(info->f_root, f_heap, 32);
    chan_commit(&ch, f_heap, 32);                                                          /* :85 */

    /* ---- Part 2 (prover.rs:101-180) ---- */
    uint32_t *cp_poly = NULL, cp_len = 0;
    evals[0] = malloc(4 * (size_t)N);
    if (route == ORC_ROUTE_LITERAL) {
        t = now_s();
        uint32_t *num = malloc(4 * (size_t)(2 * n)), *tmpa = malloc(4 * (size_t)(2 * n)),
                 *tmpb = malloc(4 * (size_t)(2 * n)), *rem = malloc(4 * (size_t)(2 * n));
        uint32_t *c0 = calloc(n, 4), *c1 = calloc(n, 4), *c2 = calloc(2 * n, 4), *den = calloc(n + 1, 4);
        uint32_t c0l, c1l, c2l, denl, reml, numl, lin[2], cst[1];
        /* c0 = (f - a[0]) / (x - g[0])   :101-103 */
        cst[0] = a[0];
        numl = poly_sub(f_poly, f_len, cst, trim(cst, 1), num, &quirk);
        lin[0] = orc_sub(0, g[0], P); lin[1] = 1;
        rc |= orc_poly_div(num, numl, lin, 2, c0, &c0l, rem, &reml, P);
        if (reml != 0) rc = -11;                                                           /* :148 */
        /* c1 = (f - a[n-2]) / (x - g[n-2])   :111-113 */
        cst[0] = a[n - 2];
        numl = poly_sub(f_poly, f_len, cst, trim(cst, 1), num, &quirk);
        lin[0] = orc_sub(0, g[n - 2], P);
        rc |= orc_poly_div(num, numl, lin, 2, c1, &c1l, rem, &reml, P);
        if (reml != 0) rc = -12;                                                           /* :149 */
        /* numerator = f(g^2 x) - f(g x)^2 - f(x)^2   :134-138 */
        uint32_t *t0p = malloc(4 * (size_t)n), *t1p = malloc(4 * (size_t)n);
        for (uint32_t d = 0; d < f_len; d++) {
            t0p[d] = orc_mul(f_poly[d], orc_pow(g[2], d, P), P);                           /* apply_const :74-79 */
            t1p[d] = orc_mul(f_poly[d], orc_pow(g[1], d, P), P);
        }
        poly_mul(t1p, f_len, t1p, f_len, tmpa);                                            /* t1*t1 */
        poly_mul(f_poly, f_len, f_poly, f_len, tmpb);                                      /* f*f */
        if (fib_shape) {
            numl = poly_sub(t0p, f_len, tmpa, 2 * f_len - 1, rem, &quirk);
            numl = poly_sub(rem, numl, tmpb, 2 * f_len - 1, num, &quirk);
        } else {
            /* other programs of the family (no reference code): t0 - c20 t1^2 - c02 f^2 - c11 f t1 - c10 t1 - c01 f - c00 */
            const uint32_t q20 = air->c20 % P, q02 = air->c02 % P, q11 = air->c11 % P, q10 = air->c10 % P, q01 = air->c01 % P;
            memset(num, 0, 4 * (size_t)(2 * n));
            for (uint32_t d = 0; d < 2 * f_len - 1; d++)
                num[d] = orc_sub(orc_sub(d < f_len ? t0p[d] : 0, orc_mul(q20, tmpa[d], P), P), orc_mul(q02, tmpb[d], P), P);
            poly_mul(f_poly, f_len, t1p, f_len, tmpa);                                     /* f*t1 */
            for (uint32_t d = 0; d < 2 * f_len - 1; d++) num[d] = orc_sub(num[d], orc_mul(q11, tmpa[d], P), P);
            for (uint32_t d = 0; d < f_len; d++)
                num[d] = orc_sub(orc_sub(num[d], orc_mul(q10, t1p[d], P), P), orc_mul(q01, f_poly[d], P), P);
            num[0] = orc_sub(num[0], air->c00 % P, P);
            numl = trim(num, 2 * f_len - 1);
        }
        /* denominator = (x^n - 1) / ((x-g[n-1])(x-g[n-3])(x-g[n-2]))   :140-144 */
        uint32_t *xn1 = calloc(n + 1, 4), xn1l, cub[4], q2[3], l0[2], l1[2], l2[2];
        {
            uint32_t *xp = calloc(n + 1, 4);
            xp[n] = 1;                                                                     /* x(1, n) */
            cst[0] = 1;                                                                    /* x(1, 0) */
            xn1l = poly_sub(xp, n + 1, cst, 1, xn1, &quirk);
            free(xp);
        }
        l0[0] = orc_sub(0, g[n - 3], P); l0[1] = 1;  /* tp0 */
        l1[0] = orc_sub(0, g[n - 2], P); l1[1] = 1;  /* tp1 */
        l2[0] = orc_sub(0, g[n - 1], P); l2[1] = 1;  /* tp2 */
        poly_mul(l2, 2, l0, 2, q2);                  /* &tp2 * &tp0 */
        poly_mul(q2, 3, l1, 2, cub);                 /* ... * tp1 */
        rc |= orc_poly_div(xn1, xn1l, cub, 4, den, &denl, rem, &reml, P);
        if (reml != 0) rc = -13;                                                           /* :150 */
        rc |= orc_poly_div(num, numl, den, denl, c2, &c2l, rem, &reml, P);                 /* :145 */
        if (reml != 0) rc = -14;                                                           /* :151 */
        if (fib_shape && (c0l != n - 2 || c1l != n - 2 || c2l != n)) rc = -15;             /* :154-156 */
        if (c0l > n - 2 || c1l > n - 2 || c2l > n) rc = -15;
        if (k == 10 && rc == 0 && is_fibsq && secret == 3141592u) {                        /* :157-159 */
            if (orc_poly_solve(c0, c0l, 2718, P) != 2509888982u) rc = -16;
            if (orc_poly_solve(c1, c1l, 5772, P) != 232961446u) rc = -17;
            if (orc_poly_solve(c2, c2l, 31415, P) != 2090051528u) rc = -18;
        }
        info->t_cp = now_s() - t;
        if (rc) { free(num); free(tmpa); free(tmpb); free(rem); free(c0); free(c1); free(c2);
                  free(den); free(t0p); free(t1p); free(xn1); goto done2; }
        /* :163-166 */
        for (int j = 0; j < 3; j++) info->alpha_raw[j] = chan_get_u32(&ch);
        uint32_t al0 = info->alpha_raw[0] % P, al1 = info->alpha_raw[1] % P, al2 = info->alpha_raw[2] % P;
        t = now_s();
        cp_poly = calloc(n, 4);
        for (uint32_t d = 0; d < n; d++) {
            uint32_t v = 0;
            if (d < c0l) v = orc_add(v, orc_mul(c0[d], al0, P), P);
            if (d < c1l) v = orc_add(v, orc_mul(c1[d], al1, P), P);
            if (d < c2l) v = orc_add(v, orc_mul(c2[d], al2, P), P);
            cp_poly[d] = v;
        }
        cp_len = trim(cp_poly, n);
        if (fib_shape && cp_len != n) rc = -19;                                            /* :169 */
        cp_len = n;             /* a lower-degree program keeps its zero top coefficients: the folds halve n exactly */
        for (uint32_t i = 0; i < N; i++) evals[0][i] = orc_poly_solve(cp_poly, cp_len, dom[i], P); /* :173 */
        info->t_cp += now_s() - t;
        free(num); free(tmpa); free(tmpb); free(rem); free(c0); free(c1); free(c2);
        free(den); free(t0p); free(t1p); free(xn1);
        if (rc) goto done2;
    } else {
        for (int j = 0; j < 3; j++) info->alpha_raw[j] = chan_get_u32(&ch);
        t = now_s();
        orc_cp_eval_air(air, f_eval, k, LOG_BLOWUP, info->alpha_raw, a[0], a[n - 2], evals[0], threads);
        info->t_cp = now_s() - t;
    }
    heaps[0] = malloc(32 * (2 * (size_t)N - 1));
    t = now_s();
    orc_merkle_build(evals[0], N, heaps[0], threads);                                      /* :176 */
    info->t_merkle += now_s() - t;
    memcpy(info->cp_root, heaps[0], 32);
    chan_commit(&ch, heaps[0], 32);                                                        /* :180 */

    /* ---- Part 3 (prover.rs:192-254) ---- */
    {
        uint32_t *cur_poly = cp_poly, cur_len = cp_len;
        uint32_t *cur_dom = dom; /* literal only */
        for (uint32_t i = 0; i < L; i++) {
            uint32_t len = N >> i, half = len / 2;
            uint32_t beta_raw = chan_get_u32(&ch);                                         /* :200 */
            info->beta_raw[i] = beta_raw;
            evals[i + 1] = malloc(4 * (size_t)half);
            t = now_s();
            if (route == ORC_ROUTE_LITERAL) {
                uint32_t *np = calloc(cur_len / 2 ? cur_len / 2 : 1, 4);
                orc_fri_coeff(cur_poly, cur_len, beta_raw, np, P);                         /* :201 */
                uint32_t *nd = malloc(4 * (size_t)half);
                for (uint32_t j = 0; j < half; j++) nd[j] = orc_pow(cur_dom[j], 2, P);     /* :204-208 */
                for (uint32_t j = 0; j < half; j++)
                    evals[i + 1][j] = orc_poly_solve(np, cur_len / 2, nd[j], P);           /* :211 */
                if (cur_poly != cp_poly) free(cur_poly);
                if (cur_dom != dom) free(cur_dom);
                cur_poly = np; cur_len = cur_len / 2; cur_dom = nd;
            } else {
                orc_fri_fold(evals[i], len, i, beta_raw, evals[i + 1], threads);
            }
            info->t_fri += now_s() - t;
            heaps[i + 1] = malloc(32 * (2 * (size_t)half - 1));
            t = now_s();
            orc_merkle_build(evals[i + 1], half, heaps[i + 1], threads);                   /* :214 */
            info->t_merkle += now_s() - t;
            memcpy(info->fri_roots[i], heaps[i + 1], 32);
            chan_commit(&ch, heaps[i + 1], 32);                                            /* :224 */
        }
        /* :254 -- cp_polys[L][0]: the constant of the degree-0 polynomial */
        uint32_t free_term;
        if (route == ORC_ROUTE_LITERAL) {
            if (cur_len != 1) rc = -20;                                                    /* :238 */
            free_term = cur_len ? cur_poly[cur_len - 1] : 0;
            if (cur_poly != cp_poly) free(cur_poly);
            if (cur_dom != dom) free(cur_dom);
        } else {
            free_term = evals[L][0];
            for (uint32_t j = 1; j < (N >> L); j++)
                if (evals[L][j] != free_term) rc = -21; /* last layer must be constant (degree 0) */
        }
        info->free_term = free_term;
        chan_commit_u32(&ch, free_term);
    }
    if (rc) goto done2;

    /* ---- Part 4 (prover.rs:263-292) ---- */
    {
        t = now_s();
        uint8_t p0[32 * 40], p1[32 * 40];
        uint32_t q = chan_get_u32(&ch);
        info->query_raw = q;
        uint32_t x = (uint32_t)((size_t)q % (N - 2 * B));                                  /* :263 */
        uint32_t idx[3] = {x, x + B, x + 2 * B};
        for (int j = 0; j < 3; j++) {                                                      /* :266-276 */
            int pl = orc_merkle_trace(f_heap, N, idx[j], p0);
            if ((uint32_t)pl != plen0) rc = -22;
            chan_commit_val_path(&ch, f_eval[idx[j]], p0, (uint32_t)pl);
        }
        {
            int pl = orc_merkle_trace(heaps[0], N, x, p0);                                 /* :272-277 */
            chan_commit_val_path(&ch, evals[0][x], p0, (uint32_t)pl);
        }
        for (uint32_t i = 0; i < L; i++) {                                                 /* :280-289 */
            uint32_t len = N >> i, xi = x % len, nx = (xi + len / 2) % len;
            int l0 = orc_merkle_trace(heaps[i], len, xi, p0);
            int l1 = orc_merkle_trace(heaps[i], len, nx, p1);
            chan_commit_pair(&ch, evals[i][xi], evals[i][nx], p0, (uint32_t)l0, p1, (uint32_t)l1);
        }
        info->t_decommit = now_s() - t;
    }

done2:
    for (uint32_t i = 0; i <= L; i++) { free(heaps[i]); free(evals[i]); }
    free(cp_poly);
    free(f_heap);
    free(heaps); free(evals);
done1:
    free(a); free(f_poly); free(f_eval); free(g); free(dom);
    if (quirk) rc = rc ? rc : -30; /* polynomial.rs:177 "0 - rhs = rhs" must never fire on the path */
    if (ch.overflow) rc = rc ? rc : -3;
    if (len_out) *len_out = ch.len;
    if (state_out) memcpy(state_out, ch.state, 32);
    info->t_total = now_s() - t0;
    return rc;
}
