// Proof::verify (proof.rs:15-149) as a product feature, generalised to n = 2^log_n.
// Host code: one query point, ~24 authentication paths -- not data-parallel work.  Beyond the reference it can
// (a) replay the Fiat-Shamir channel (channel.rs:19-32) and check that every alpha / beta / query in the transcript
// is the value the channel would have drawn (the reference reads them from the proof, proof.rs:22-37) and
// (b) check Proof.state, which the reference ignores (proof.rs:5-8).
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/zkstark_b200.h"
#include "field.cuh"
#include "hostair.hpp"

namespace zk {
namespace sha {
void host_sha256(const uint8_t* data, size_t len, uint8_t out[32]);
}
}  // namespace zk

using namespace zk;

namespace {

struct Reader {  // bincode 1.3 fixint little-endian (channel.rs:20)
    const uint8_t* p;
    size_t len, off = 0;
    bool ok = true;
    uint32_t u32()
    {
        if (off + 4 > len) { ok = false; return 0; }
        uint32_t v = p[off] | (p[off + 1] << 8) | (p[off + 2] << 16) | ((uint32_t)p[off + 3] << 24);
        off += 4;
        return v;
    }
    uint64_t u64()
    {
        uint64_t lo = u32(), hi = u32();
        return lo | (hi << 32);
    }
    const uint8_t* bytes(size_t n)
    {
        if (off + n > len) { ok = false; return p; }
        const uint8_t* r = p + off;
        off += n;
        return r;
    }
};

struct Replay {  // channel.rs:11-32
    uint8_t state[32];
    bool on;
    explicit Replay(bool enable) : on(enable) { memset(state, 0, 32); }
    void commit(const uint8_t* v, size_t n)
    {
        if (!on) return;
        uint8_t buf[32 + 24 + 64 * 40];
        memcpy(buf, state, 32);
        memcpy(buf + 32, v, n);
        sha::host_sha256(buf, 32 + n, state);
    }
    uint32_t draw() const { return ((uint32_t)state[0] << 24) | (state[1] << 16) | (state[2] << 8) | state[3]; }
};

uint32_t f_add(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + b) % P); }
uint32_t f_sub(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + P - b) % P); }
uint32_t f_div(uint32_t a, uint32_t b) { return h_mul(a, h_inv(b)); }

// merkle.rs:82-110 compute_root_from_path
void root_from_path(uint32_t element, uint64_t index, const uint8_t* path, uint32_t plen, uint8_t out[32])
{
    uint8_t cur[32], buf[64];
    index += ((uint64_t)1 << plen) - 1;
    uint8_t be[4] = {(uint8_t)(element >> 24), (uint8_t)(element >> 16), (uint8_t)(element >> 8), (uint8_t)element};
    sha::host_sha256(be, 4, cur);
    for (uint32_t i = 0; i < plen; i++) {
        const uint8_t* h = path + 32 * (size_t)i;
        if (index % 2 == 0) { memcpy(buf, h, 32); memcpy(buf + 32, cur, 32); index -= 2; }
        else { memcpy(buf, cur, 32); memcpy(buf + 32, h, 32); index -= 1; }
        sha::host_sha256(buf, 64, cur);
        index >>= 1;
    }
    memcpy(out, cur, 32);
}

// One query: the 4 + L decommitment records of a test point (prover.rs:266-289) parsed, replayed into the channel and
// checked (proof.rs:63-148).  Returns 0 or the proof.rs line of the assert that fails.
struct ProofHead {
    uint32_t k, L, log_N;
    uint64_t n, N, B;
    const uint8_t* f_root;
    uint32_t alpha[3], beta[33];
    const uint8_t* roots[33];
    uint32_t free_term, a_last;
    zk::HostAir air;    // the program (default: the reference's FibonacciSq)
};
int check_query(const ProofHead& H, Reader& rd, Replay& ch, const uint8_t* transcript, uint32_t q_raw)
{
    const uint32_t L = H.L, log_N = H.log_N, k = H.k;
    const uint64_t n = H.n, N = H.N, B = H.B;
    size_t at;
    uint32_t fv[4];
    const uint8_t* fp[4];
    for (int j = 0; j < 4; j++) {
        at = rd.off;
        fv[j] = rd.u32();
        if (rd.u64() != log_N) return 38 + j;
        fp[j] = rd.bytes(32 * (size_t)log_N);
        if (!rd.ok) return 38 + j;
        ch.commit(transcript + at, rd.off - at);
    }
    uint32_t lx[33], lnx[33];
    const uint8_t *px[33], *pnx[33];
    for (uint32_t i = 0; i < L; i++) {
        at = rd.off;
        const uint32_t pl = log_N - i;
        lx[i] = rd.u32();
        lnx[i] = rd.u32();
        if (rd.u64() != pl) return 44;
        px[i] = rd.bytes(32 * (size_t)pl);
        if (rd.u64() != pl) return 44;
        pnx[i] = rd.bytes(32 * (size_t)pl);
        if (!rd.ok) return 44;
        ch.commit(transcript + at, rd.off - at);
    }
    // proof.rs:49-60 -- protocol constants and the test point
    const uint32_t w = GEN, g = h_root(k), h = h_root(log_N);
    const uint64_t tp = (uint64_t)q_raw % (N - 2 * B);
    const uint32_t x = h_mul(w, h_pow(h, tp));
    // proof.rs:63-77 -- the composition polynomial at the test point
    {
        const uint32_t f_x = fv[0] % P, f_gx = fv[1] % P, f_ggx = fv[2] % P;
        const uint32_t g1 = h_pow(g, n - 1), g2 = h_pow(g, n - 2), g3 = h_pow(g, n - 3);
        const uint32_t p0 = f_div(f_sub(f_x, H.air.a0), f_sub(x, 1));     // proof.rs:66: the first trace value (1 upstream)
        const uint32_t p1 = f_div(f_sub(f_x, H.a_last % P), f_sub(x, g2));
        // proof.rs:70: f(g^2 x) - f(g x)^2 - f(x)^2 upstream; in general f(g^2 x) - step(f(x), f(g x)) of the program
        uint32_t step = f_add(h_mul(H.air.c20, h_mul(f_gx, f_gx)), h_mul(H.air.c02, h_mul(f_x, f_x)));
        step = f_add(step, h_mul(H.air.c11, h_mul(f_x, f_gx)));
        step = f_add(step, f_add(h_mul(H.air.c10, f_gx), h_mul(H.air.c01, f_x)));
        step = f_add(step, H.air.c00);
        const uint32_t num = f_sub(f_ggx, step);
        const uint32_t den = f_div(f_sub(h_pow(x, n), 1), h_mul(h_mul(f_sub(x, g3), f_sub(x, g2)), f_sub(x, g1)));
        const uint32_t p2 = f_div(num, den);
        const uint32_t cp0 = f_add(f_add(h_mul(H.alpha[0] % P, p0), h_mul(H.alpha[1] % P, p1)), h_mul(H.alpha[2] % P, p2));
        if (cp0 != fv[3]) return 76;
    }
    // proof.rs:80-95 -- trace authentication paths
    uint8_t r[32];
    root_from_path(fv[0], tp, fp[0], log_N, r);          if (memcmp(r, H.f_root, 32)) return 80;
    root_from_path(fv[1], tp + B, fp[1], log_N, r);      if (memcmp(r, H.f_root, 32)) return 84;
    root_from_path(fv[2], tp + 2 * B, fp[2], log_N, r);  if (memcmp(r, H.f_root, 32)) return 88;
    root_from_path(fv[3], tp, fp[3], log_N, r);          if (memcmp(r, H.roots[0], 32)) return 92;
    // proof.rs:101-126 -- fold relations cp_{i+1}(x^2) = g(x^2) + beta h(x^2)
    const uint32_t inv2 = h_inv(2);
    for (uint32_t i = 0; i < L; i++) {
        const uint32_t next = (i + 1 < L) ? lx[i + 1] : H.free_term;
        const uint32_t xi = h_pow(x, (uint64_t)1 << i);
        const uint32_t g_xx = h_mul(f_add(lx[i] % P, lnx[i] % P), inv2);
        const uint32_t h_xx = f_div(f_sub(lx[i] % P, lnx[i] % P), h_mul(xi, 2));
        if (f_add(g_xx, h_mul(H.beta[i + 1] % P, h_xx)) != next) return i + 1 < L ? 113 : 125;
    }
    // proof.rs:129-148 -- FRI authentication paths
    for (uint32_t i = 0; i < L; i++) {
        const uint64_t size = N >> i;
        root_from_path(lx[i], tp % size, px[i], log_N - i, r);
        if (memcmp(r, H.roots[i], 32)) return 136;
        root_from_path(lnx[i], (tp + size / 2) % size, pnx[i], log_N - i, r);
        if (memcmp(r, H.roots[i], 32)) return 140;
    }
    return 0;
}

}  // namespace

extern "C" size_t zk_transcript_len_ex(uint32_t log_n, uint32_t n_extra_queries)
{
    return zk_transcript_len(log_n) + (size_t)n_extra_queries * (4 + zk_decommit_len(log_n));
}

// proof.rs:15-149 for a proof made by zk_prove_fibsq_ex: the channel starts from init_state (null: zeros, channel.rs:12-17)
// and n_extra_queries further test points, each drawn from the channel after the records of the previous one, follow the
// reference's own.  With init_state null and n_extra_queries 0 this is zk_verify_fibsq.
extern "C" int zk_verify_fibsq_ex(uint32_t log_n, uint32_t a_last, const uint8_t* transcript, size_t len, const uint8_t* state,
                                  const uint8_t* init_state, uint32_t n_extra_queries, int rederive, int* failed_check)
{
    return zk_verify_air(nullptr, log_n, a_last, transcript, len, state, init_state, n_extra_queries, rederive, failed_check);
}

// The same for a proof of another program of the zk_air family (air null: the reference's FibonacciSq): the
// first boundary value is air->a0 and the transition numerator is f(g^2 x) - step(f(x), f(g x)).
extern "C" int zk_verify_air(const zk_air* air, uint32_t log_n, uint32_t a_last, const uint8_t* transcript, size_t len,
                             const uint8_t* state, const uint8_t* init_state, uint32_t n_extra_queries, int rederive,
                             int* failed_check)
{
    int dummy = 0;
    if (!failed_check) failed_check = &dummy;
    *failed_check = 0;
    if (!transcript || log_n < ZK_MIN_LOG_N || log_n > ZK_MAX_LOG_N || n_extra_queries > 65535) return ZK_ERR_INVALID;
    if (len != zk_transcript_len_ex(log_n, n_extra_queries)) { *failed_check = 20; return ZK_ERR_VERIFY; }  // deserialize .unwrap() (proof.rs:20-46)
    ProofHead H{};
    H.k = log_n; H.L = log_n; H.log_N = log_n + ZK_LOG_BLOWUP;
    H.n = (uint64_t)1 << log_n; H.N = H.n << ZK_LOG_BLOWUP; H.B = (uint64_t)1 << ZK_LOG_BLOWUP;
    H.a_last = a_last;
    if (air) {
        if (air->flags & ~ZK_AIR_NO_SPECIALISE) return ZK_ERR_INVALID;
        H.air.a0 = air->a0 % P; H.air.c20 = air->c20 % P; H.air.c02 = air->c02 % P; H.air.c11 = air->c11 % P;
        H.air.c10 = air->c10 % P; H.air.c01 = air->c01 % P; H.air.c00 = air->c00 % P;
    }
    const uint32_t L = H.L;
    Reader rd{transcript, len};
    Replay ch(rederive != 0);
    if (init_state) memcpy(ch.state, init_state, 32);
#define FAIL(line) do { *failed_check = (line); return ZK_ERR_VERIFY; } while (0)

    // proof.rs:20-46 -- pull elements out of the proof (and replay the channel beside it)
    H.f_root = rd.bytes(32);
    ch.commit(H.f_root, 32);
    for (int j = 0; j < 3; j++) {
        size_t at = rd.off;
        H.alpha[j] = rd.u32();
        if (ch.on && ch.draw() != H.alpha[j]) FAIL(22 + j);
        ch.commit(transcript + at, 4);
    }
    H.roots[0] = rd.bytes(32);
    ch.commit(H.roots[0], 32);
    for (uint32_t i = 1; i <= L; i++) {
        size_t at = rd.off;
        H.beta[i] = rd.u32();
        if (ch.on && ch.draw() != H.beta[i]) FAIL(30);
        ch.commit(transcript + at, 4);
        H.roots[i] = rd.bytes(32);
        ch.commit(H.roots[i], 32);
    }
    size_t at = rd.off;
    H.free_term = rd.u32();
    ch.commit(transcript + at, 4);
    for (uint32_t q = 0; q <= n_extra_queries; q++) {
        at = rd.off;
        const uint32_t q_raw = rd.u32();
        if (!rd.ok) FAIL(37);
        if (ch.on && ch.draw() != q_raw) FAIL(37);           // prover.rs:263: the test point comes from the channel
        ch.commit(transcript + at, 4);
        const int line = check_query(H, rd, ch, transcript, q_raw);
        if (line) FAIL(line);
    }
    if (!rd.ok || rd.off != len) FAIL(46);
    if (ch.on && state && memcmp(state, ch.state, 32) != 0) FAIL(6);  // Proof.state (proof.rs:6), unchecked upstream
#undef FAIL
    return ZK_OK;
}

extern "C" int zk_verify_fibsq(uint32_t log_n, uint32_t a_last, const uint8_t* transcript, size_t len,
                               const uint8_t* state, int rederive, int* failed_check)
{
    return zk_verify_fibsq_ex(log_n, a_last, transcript, len, state, nullptr, 0, rederive, failed_check);
}

// ---------------------------------------------------------------- wire format
// `Proof { state: [u8; 32], data: Box<[u8]> }` (proof.rs:5-8) as bincode 1.x serialises the struct: the 32 state bytes
// (serde tuple: no length), the u64 little-endian length of `data`, the bytes of `data`.  `data` is the channel
// transcript (channel.rs:8,34-36), already bincode, which the untouched `Proof::new(state, data).verify()` reads as is.
extern "C" size_t zk_proof_serialized_len(size_t transcript_len) { return 32 + 8 + transcript_len; }

extern "C" int zk_proof_serialize(const uint8_t state[32], const uint8_t* transcript, size_t transcript_len, uint8_t* out, size_t cap,
                                  size_t* len_out)
{
    if (!state || (!transcript && transcript_len) || !out) return ZK_ERR_INVALID;
    const size_t need = 40 + transcript_len;
    if (cap < need) return ZK_ERR_BUFFER;
    memcpy(out, state, 32);
    for (int i = 0; i < 8; i++) out[32 + i] = (uint8_t)((uint64_t)transcript_len >> (8 * i));
    if (transcript_len) memcpy(out + 40, transcript, transcript_len);
    if (len_out) *len_out = need;
    return ZK_OK;
}

extern "C" int zk_proof_deserialize(const uint8_t* bytes, size_t len, uint8_t state_out[32], const uint8_t** transcript_out,
                                    size_t* transcript_len_out)
{
    if (!bytes || !state_out || !transcript_out || !transcript_len_out) return ZK_ERR_INVALID;
    if (len < 40) return ZK_ERR_BUFFER;                       // truncated
    uint64_t n = 0;
    for (int i = 0; i < 8; i++) n |= (uint64_t)bytes[32 + i] << (8 * i);
    if (n != (uint64_t)(len - 40)) return ZK_ERR_INVALID;     // the length prefix does not match the payload
    memcpy(state_out, bytes, 32);
    *transcript_out = bytes + 40;
    *transcript_len_out = (size_t)n;
    return ZK_OK;
}

extern "C" void zk_sha256(const uint8_t* data, size_t len, uint8_t out[32]) { sha::host_sha256(data, len, out); }
