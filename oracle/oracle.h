/*
 * oracle/oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the algorithm of the reference STARK-101 prover
 * (Crocodoctopus/zkstark, /root/reference/src/ *.rs).  Only tests/, the smoke()
 * entry and bench.py's cpu_baseline / --impl reference legs may link or call
 * this.  The shipped CUDA path (zkstark_b200/csrc) never does.
 *
 * The reference is a Rust crate; there is no cargo/rustc in this image, so
 * the reference itself cannot be compiled (no oracle/_ref).  Parity is pinned
 * on the known-answer values the reference embeds in its own source:
 *   merkle.rs:117-181   (7 digests + 4 auth paths of the tree over [1,2,3,4])
 *   polynomial.rs:407-425 (coefficient FRI fold over GF(101))
 *   polynomial.rs:428-490 (lagrange / div tests)
 *   field.rs:213-226    (generator test over GF(4391))
 *   prover.rs:42,73-78,157-159 (trace end value, six LDE values, c0/c1/c2 KATs)
 * The transcript byte layout (bincode 1.3 fixint little-endian; channel.rs:19-26)
 * is NOT pinned by any reference test: "parity unpinned" for that layer only --
 * the only check available is that the restated verifier (proof.rs:15-149)
 * deserialises what the restated channel serialised.
 *
 * Third-party algorithms restated here because their source is not under
 * /root/reference: num-modular 0.5.1 (MontgomeryInt<u32>: plain arithmetic mod
 * P, canonical residues), sha2 0.10.6 (FIPS 180-4 SHA-256), bincode 1.3.3 +
 * serde 1.0.147 (fixint LE, u64 length prefix for Box<[T]>).
 *
 * Two routes produce byte-identical transcripts:
 *   ORC_ROUTE_LITERAL  the reference's own algorithms (O(n^3) Lagrange,
 *                      long division, per-point evaluation), log_n <= 10;
 *   ORC_ROUTE_FAST     NTT-based, multithreaded (the CPU baseline).
 */
#ifndef ZK_ORACLE_H
#define ZK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_P 3221225473u /* main.rs:13  type F = Gf<3221225473> */
#define ORC_ROUTE_LITERAL 0
#define ORC_ROUTE_FAST 1

/* ---- field (field.rs:96-211, generic modulus for the reference's unit tests) */
uint32_t orc_add(uint32_t a, uint32_t b, uint32_t p);
uint32_t orc_sub(uint32_t a, uint32_t b, uint32_t p);
uint32_t orc_mul(uint32_t a, uint32_t b, uint32_t p);
uint32_t orc_pow(uint32_t a, uint32_t e, uint32_t p);
uint32_t orc_inv(uint32_t a, uint32_t p);
uint32_t orc_from_i32(int32_t v, uint32_t p);   /* field.rs:10-18 */
uint32_t orc_generator(uint32_t p);             /* field.rs:52-86 */
uint32_t orc_order(uint32_t a, uint32_t p);     /* field.rs:45-49 */

/* ---- SHA-256 (sha2 0.10.6 = FIPS 180-4) */
void orc_sha256(const uint8_t *data, size_t len, uint8_t out[32]);
int orc_sha256_uses_shani(void);
void orc_sha256_force_portable(int on);

/* ---- merkle.rs */
/* merkle.rs:14-51  heap of 2*size-1 digests, root at 0, leaves at [size-1, 2size-1) */
int orc_merkle_build(const uint32_t *leaves, uint64_t size, uint8_t *heap_out, int threads);
/* merkle.rs:54-71  sibling digests leaf -> root; returns path length */
int orc_merkle_trace(const uint8_t *heap, uint64_t size, uint64_t idx, uint8_t *path_out);
/* merkle.rs:82-110 */
void orc_root_from_path(uint32_t element, uint64_t index, const uint8_t *path,
                        uint32_t path_len, uint8_t out[32]);

/* ---- polynomial.rs (literal algorithms, ascending coefficient arrays) */
/* polynomial.rs:49-56 */
uint32_t orc_poly_solve(const uint32_t *c, uint32_t len, uint32_t t, uint32_t p);
/* polynomial.rs:337-383; out has npts coefficients (ascending), returns trimmed length */
uint32_t orc_lagrange(const uint32_t *xs, const uint32_t *ys, uint32_t npts,
                      uint32_t *out, uint32_t p);
/* polynomial.rs:307-334; q has max(lhs_len-rhs_len+1,0) slots, r has lhs_len slots.
 * Lengths returned through *q_len, *r_len (trimmed as Add/Sub do). */
int orc_poly_div(const uint32_t *lhs, uint32_t lhs_len, const uint32_t *rhs, uint32_t rhs_len,
                 uint32_t *q, uint32_t *q_len, uint32_t *r, uint32_t *r_len, uint32_t p);
/* polynomial.rs:385-400; out has len/2 slots */
void orc_fri_coeff(const uint32_t *c, uint32_t len, uint32_t beta, uint32_t *out, uint32_t p);

/* ---- stage functions, fast route (what the CUDA kernels are compared with) */
/* the program being proven as data -- a one-column AIR over a three-row window,
 *   a[0] = a0, a[1] = secret, a[i+2] = c20 a[i+1]^2 + c02 a[i]^2 + c11 a[i] a[i+1] + c10 a[i+1] + c01 a[i] + c00,
 * with the reference's two boundary constraints (first value, value n-2).  The reference's own program
 * (prover.rs:32-39, :134-138) is ORC_AIR_FIBSQ; every *_fibsq function below is the *_air one at that point. */
typedef struct { uint32_t a0, c20, c02, c11, c10, c01, c00; } orc_air;
#define ORC_AIR_FIBSQ {1u, 1u, 1u, 0u, 0u, 0u, 0u}
uint32_t orc_air_step(const orc_air *air, uint32_t x0, uint32_t x1);   /* the right-hand side at (a[i], a[i+1]) */
void orc_trace_air(const orc_air *air, uint32_t log_n, uint32_t secret, uint32_t *out);
/* prover.rs:32-39 generalised: a0=1, a1=secret, a_i=a_{i-2}^2+a_{i-1}^2; n-1 values */
void orc_trace_fibsq(uint32_t log_n, uint32_t secret, uint32_t *out);
/* polynomial.rs:337 for x_i = g^i: n_points = n-1 (phantom point) or n; n coeffs ascending */
int orc_interpolate_subgroup(const uint32_t *y, uint32_t n_points, uint32_t log_n,
                             uint32_t *coeffs_out, int threads);
/* prover.rs:69-70: evals[i] = f(shift * h^i), i < n << log_blowup, natural order */
int orc_lde_coset(const uint32_t *coeffs, uint32_t log_n, uint32_t log_blowup, uint32_t shift,
                  uint32_t *evals_out, int threads);
/* prover.rs:101-173 pointwise */
int orc_cp_eval_fibsq(const uint32_t *f_eval, uint32_t log_n, uint32_t log_blowup,
                      const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last,
                      uint32_t *cp_out, int threads);
int orc_cp_eval_air(const orc_air *air, const uint32_t *f_eval, uint32_t log_n, uint32_t log_blowup,
                    const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last,
                    uint32_t *cp_out, int threads);
/* polynomial.rs:385 + prover.rs:204-211 in evaluation form; cur has len values on the
 * domain (5 h^i)^(2^layer_k), h of order domain_size0 = len << layer_k */
int orc_fri_fold(const uint32_t *cur, uint64_t len, uint32_t layer_k, uint32_t beta_raw,
                 uint32_t *next_out, int threads);

/* ---- prover.rs:9-293 generalised by SURVEY 8(d).  Returns 0 or <0. */
typedef struct {
    uint32_t log_n;
    uint32_t a_last;        /* a[n-2] */
    uint32_t alpha_raw[3];
    uint32_t beta_raw[32];
    uint32_t free_term;
    uint32_t query_raw;
    uint8_t f_root[32];
    uint8_t cp_root[32];
    uint8_t fri_roots[32][32];
    double t_trace, t_interp, t_lde, t_cp, t_fri, t_merkle, t_decommit, t_total;
} orc_prove_info;

int orc_prove_fibsq(uint32_t log_n, uint32_t secret, int route, int threads,
                    uint8_t *transcript_out, size_t cap, size_t *len_out,
                    uint8_t state_out[32], orc_prove_info *info /* may be NULL */);
int orc_prove_air(const orc_air *air, uint32_t log_n, uint32_t secret, int route, int threads,
                  uint8_t *transcript_out, size_t cap, size_t *len_out,
                  uint8_t state_out[32], orc_prove_info *info /* may be NULL */);
size_t orc_transcript_len(uint32_t log_n);

/* ---- proof.rs:15-149 generalised; rederive != 0 additionally re-derives alpha/beta/query
 * from the transcript through the channel (the reference reads them from the proof). */
int orc_verify_fibsq(uint32_t log_n, uint32_t a_last, const uint8_t *transcript, size_t len,
                     int rederive, const uint8_t *expect_state /* may be NULL */);
int orc_verify_air(const orc_air *air, uint32_t log_n, uint32_t a_last, const uint8_t *transcript, size_t len,
                   int rederive, const uint8_t *expect_state /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif
