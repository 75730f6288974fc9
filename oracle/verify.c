/*
 * oracle/verify.c -- CPU ORACLE (test infrastructure).
 * Restates Proof::verify (proof.rs:15-149) generalised to n = 2^log_n; used as the
 * acceptance test for transcripts produced by the CUDA path.  With rederive != 0 it
 * additionally replays the channel (channel.rs:19-32) and checks that every alpha /
 * beta / query in the transcript is the Fiat-Shamir value (the reference reads them
 * from the proof and ignores Proof.state, proof.rs:5-8,22-37).
 */
#include "oracle.h"
#include <string.h>

#define P ORC_P
#define LOG_BLOWUP 3u

static uint32_t rd_u32(const uint8_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24); }
static uint64_t rd_u64(const uint8_t *b)
{
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | b[i];
    return v;
}

typedef struct { uint8_t s[32]; } st_t;
static void replay_commit(st_t *st, const uint8_t *v, size_t len)
{
    uint8_t buf[32 + 24 + 64 * 40];
    memcpy(buf, st->s, 32);
    memcpy(buf + 32, v, len);
    orc_sha256(buf, 32 + len, st->s);
}
static uint32_t be32(const uint8_t *s) { return ((uint32_t)s[0] << 24) | (s[1] << 16) | (s[2] << 8) | s[3]; }

int orc_verify_fibsq(uint32_t k, uint32_t a_last, const uint8_t *tr, size_t len, int rederive,
                     const uint8_t *expect_state)
{
    const orc_air fibsq = ORC_AIR_FIBSQ;
    return orc_verify_air(&fibsq, k, a_last, tr, len, rederive, expect_state);
}
/* the same for an orc_air: first boundary value air->a0, transition numerator f(g^2 x) - step(f(x), f(g x)) */
int orc_verify_air(const orc_air *air, uint32_t k, uint32_t a_last, const uint8_t *tr, size_t len, int rederive,
                   const uint8_t *expect_state)
{
    if (k < 2 || k + LOG_BLOWUP > 30) return -1;
    if (len != orc_transcript_len(k)) return -2;
    const uint32_t n = 1u << k, N = n << LOG_BLOWUP, B = 1u << LOG_BLOWUP, L = k, pl0 = k + LOG_BLOWUP;
    size_t o = 0;
    st_t st;
    memset(&st, 0, sizeof st);

    /* proof.rs:20-46 parse (and, optionally, replay the channel) */
    const uint8_t *f_root = tr + o;
    if (rederive) replay_commit(&st, tr + o, 32);
    o += 32;
    uint32_t alpha[3], beta[33];
    for (int j = 0; j < 3; j++) {
        alpha[j] = rd_u32(tr + o);
        if (rederive) { if (be32(st.s) != alpha[j]) return -40 - j; replay_commit(&st, tr + o, 4); }
        o += 4;
    }
    const uint8_t *roots[33];
    roots[0] = tr + o;
    if (rederive) replay_commit(&st, tr + o, 32);
    o += 32;
    beta[0] = 0; /* proof.rs:27 */
    for (uint32_t i = 1; i <= L; i++) {
        beta[i] = rd_u32(tr + o);
        if (rederive) { if (be32(st.s) != beta[i]) return -50; replay_commit(&st, tr + o, 4); }
        o += 4;
        roots[i] = tr + o;
        if (rederive) replay_commit(&st, tr + o, 32);
        o += 32;
    }
    uint32_t free_term = rd_u32(tr + o);
    if (rederive) replay_commit(&st, tr + o, 4);
    o += 4;
    uint32_t q_raw = rd_u32(tr + o);
    if (rederive) { if (be32(st.s) != q_raw) return -51; replay_commit(&st, tr + o, 4); }
    o += 4;
    uint32_t fv[4];
    const uint8_t *fp[4];
    for (int j = 0; j < 4; j++) {
        size_t s = o;
        fv[j] = rd_u32(tr + o); o += 4;
        if (rd_u64(tr + o) != pl0) return -3;
        o += 8;
        fp[j] = tr + o; o += 32 * (size_t)pl0;
        if (rederive) replay_commit(&st, tr + s, o - s);
    }
    uint32_t lx[33], lnx[33];
    const uint8_t *lpx[33], *lpnx[33];
    for (uint32_t i = 0; i < L; i++) {
        size_t s = o;
        uint32_t pl = pl0 - i;
        lx[i] = rd_u32(tr + o); o += 4;
        lnx[i] = rd_u32(tr + o); o += 4;
        if (rd_u64(tr + o) != pl) return -4;
        o += 8; lpx[i] = tr + o; o += 32 * (size_t)pl;
        if (rd_u64(tr + o) != pl) return -5;
        o += 8; lpnx[i] = tr + o; o += 32 * (size_t)pl;
        if (rederive) replay_commit(&st, tr + s, o - s);
    }
    if (o != len) return -6;
    if (expect_state && memcmp(expect_state, st.s, 32) != 0 && rederive) return -7;

    /* proof.rs:49-60 */
    const uint32_t w = 5;
    const uint32_t gen_g = orc_pow(w, (P - 1) / n, P), gen_h = orc_pow(w, (P - 1) / N, P);
    uint32_t tp = (uint32_t)((size_t)q_raw % (N - 2 * B));
    uint32_t x = orc_mul(w, orc_pow(gen_h, tp, P), P);

    /* proof.rs:63-77 CP identity */
    {
        uint32_t f_x = fv[0] % P, f_gx = fv[1] % P, f_ggx = fv[2] % P;
        uint32_t g1 = orc_pow(gen_g, n - 1, P), g2 = orc_pow(gen_g, n - 2, P), g3 = orc_pow(gen_g, n - 3, P);
        uint32_t p0 = orc_mul(orc_sub(f_x, air->a0 % P, P), orc_inv(orc_sub(x, 1, P), P), P);
        uint32_t p1 = orc_mul(orc_sub(f_x, a_last % P, P), orc_inv(orc_sub(x, g2, P), P), P);
        uint32_t num = orc_sub(f_ggx, orc_air_step(air, f_x, f_gx), P);
        uint32_t cub = orc_mul(orc_mul(orc_sub(x, g3, P), orc_sub(x, g2, P), P), orc_sub(x, g1, P), P);
        uint32_t z = orc_sub(orc_pow(x, n, P), 1, P);
        uint32_t den = orc_mul(z, orc_inv(cub, P), P);
        uint32_t p2 = orc_mul(num, orc_inv(den, P), P);
        uint32_t cp0 = orc_add(orc_add(orc_mul(alpha[0] % P, p0, P), orc_mul(alpha[1] % P, p1, P), P),
                               orc_mul(alpha[2] % P, p2, P), P);
        if (cp0 != fv[3]) return -10;
    }
    /* proof.rs:80-95 */
    uint8_t r[32];
    orc_root_from_path(fv[0], tp, fp[0], pl0, r);          if (memcmp(r, f_root, 32)) return -11;
    orc_root_from_path(fv[1], tp + B, fp[1], pl0, r);      if (memcmp(r, f_root, 32)) return -12;
    orc_root_from_path(fv[2], tp + 2 * B, fp[2], pl0, r);  if (memcmp(r, f_root, 32)) return -13;
    orc_root_from_path(fv[3], tp, fp[3], pl0, r);          if (memcmp(r, roots[0], 32)) return -14;
    if (lx[0] != fv[3]) return -15; /* layer-0 value is cp[x] (implied by the shared tree) */

    /* proof.rs:101-126 fold identities */
    for (uint32_t i = 0; i < L; i++) {
        uint32_t next = (i + 1 < L) ? lx[i + 1] : free_term;
        uint32_t xi = orc_pow(x, 1u << i, P);
        uint32_t g_xx = orc_mul(orc_add(lx[i] % P, lnx[i] % P, P), orc_inv(2, P), P);
        uint32_t h_xx = orc_mul(orc_sub(lx[i] % P, lnx[i] % P, P), orc_inv(orc_mul(xi, 2, P), P), P);
        uint32_t calc = orc_add(g_xx, orc_mul(beta[i + 1] % P, h_xx, P), P);
        if (calc != next) return -20 - (int)i;
    }
    /* proof.rs:129-148 */
    for (uint32_t i = 0; i < L; i++) {
        uint32_t size = N >> i, pl = pl0 - i;
        orc_root_from_path(lx[i], tp % size, lpx[i], pl, r);
        if (memcmp(r, roots[i], 32)) return -100 - (int)i;
        orc_root_from_path(lnx[i], (tp + size / 2) % size, lpnx[i], pl, r);
        if (memcmp(r, roots[i], 32)) return -200 - (int)i;
    }
    return 0;
}
