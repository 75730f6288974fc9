/*
 * zkstark_b200.h -- C ABI of libzkstark_b200.so: the B200 (sm_100a) hot path of the STARK-101
 * FibonacciSq prover in Crocodoctopus/zkstark, as a drop-in for the Rust crate's
 * field / polynomial / merkle / prover entry points (file:line below are the reference
 * interfaces each export replaces; INTEGRATION.md shows the Rust `extern "C"` side).
 *
 * Conventions
 *  - field elements cross as canonical u32 residues mod p = 3221225473 (field.rs:41-43), never
 *    Montgomery form; raw Fiat-Shamir words (alpha/beta) cross un-reduced (channel.rs:28-32)
 *    and are reduced inside, as F::from(u32) does (field.rs:20-24);
 *  - digests cross as the 32 raw bytes sha2 emits (merkle.rs:9);
 *  - the caller owns every buffer; `zk_ctx` owns device memory, tables and one CUDA stream;
 *  - every call returns 0 on success or a negative zk_status; nothing unwinds across the
 *    boundary (the reference panics: the Rust shim turns non-zero into panic!);
 *  - a ctx is not thread-safe: one ctx per (thread, GPU).  Host-buffer calls are synchronous on
 *    return; `_dev` calls take device pointers, enqueue on the ctx stream and do not synchronize;
 *  - there is no CPU fallback: without a CUDA device zk_ctx_create fails with ZK_ERR_CUDA.
 */
#ifndef ZKSTARK_B200_H
#define ZKSTARK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZK_MODULUS 3221225473u /* main.rs:13 */
#define ZK_LOG_BLOWUP 3u       /* prover.rs:56-57 (8192 / 1024) */
#define ZK_MIN_LOG_N 3u
#define ZK_MAX_LOG_N 27u

typedef struct zk_ctx zk_ctx;

typedef enum zk_status {
    ZK_OK = 0,
    ZK_ERR_INVALID = -1, /* bad argument (size not a power of two, log_n out of range, null) */
    ZK_ERR_CUDA = -2,    /* CUDA runtime error, see zk_last_error */
    ZK_ERR_NOMEM = -3,   /* device or host allocation failed */
    ZK_ERR_BUFFER = -4,  /* caller buffer too small */
    ZK_ERR_ASSERT = -5,  /* an inline assert of the reference failed (e.g. prover.rs:42) */
    ZK_ERR_VERIFY = -6,  /* zk_verify_fibsq: the proof was rejected */
    ZK_ERR_COMM = -7     /* NCCL is unavailable or a collective failed (sharded proofs only) */
} zk_status;

/* ---- context ---------------------------------------------------------------------------- */
int zk_ctx_create(int device, zk_ctx **out);
/* enqueue on an existing CUDA stream (cudaStream_t) instead of a private one, e.g. torch's */
int zk_ctx_create_on_stream(int device, void *cuda_stream, zk_ctx **out);
void zk_ctx_destroy(zk_ctx *ctx);
const char *zk_last_error(const zk_ctx *ctx);
const char *zk_version(void);
int zk_ctx_synchronize(zk_ctx *ctx);
/* the cudaStream_t every call of this context enqueues on (its own, or the one given to zk_ctx_create_on_stream): callers
 * that mix their own device work with `_dev` calls must issue it on this stream or order it with events */
void *zk_ctx_stream(const zk_ctx *ctx);
/* number of kernels launched by this ctx since creation (bench.py's gpu_launches) */
uint64_t zk_ctx_launch_count(const zk_ctx *ctx);
/* device-side time of the last zk_prove_* call in ms, by stage:
 * [0] interpolate [1] lde [2] f tree [3] cp+tree [4] fri folds+trees [5] decommit [6] total */
int zk_ctx_last_stage_ms(zk_ctx *ctx, float out_ms[7]);
/* enable per-stage CUDA-event timing inside zk_prove_* (off by default: events cost launches) */
int zk_ctx_set_stage_timing(zk_ctx *ctx, int on);

/* ---- polynomial.rs ---------------------------------------------------------------------- */
/* lagrange::<F>(&points) for points (g^i, y_i), g = 5^((p-1)/2^log_n)   (polynomial.rs:337,
 * called at prover.rs:60-61).  n_points is 2^log_n - 1 (the reference: zip truncates to the
 * 1023 trace values) or 2^log_n.  coeffs_out: 2^log_n ascending coefficients (Polynomial.0). */
int zk_interpolate_subgroup(zk_ctx *ctx, const uint32_t *y, uint32_t n_points, uint32_t log_n,
                            uint32_t *coeffs_out);
/* f_poly.solve(shift * h^i) for i < 2^(log_n+log_blowup), h = 5^((p-1)/N), natural order
 * (polynomial.rs:49-56 mapped over f_domain, prover.rs:69-70).  log_blowup must be 3. */
int zk_lde_coset(zk_ctx *ctx, const uint32_t *coeffs, uint32_t log_n, uint32_t log_blowup,
                 uint32_t shift, uint32_t *evals_out);
/* fri(poly, beta) followed by evaluation on the squared half domain (polynomial.rs:385-400,
 * prover.rs:200-211), in evaluation form.  cur: `len` values of layer `layer_k` on the domain
 * (5 h^i)^(2^layer_k), h of order len << layer_k.  next_out: len/2 values. */
int zk_fri_fold(zk_ctx *ctx, const uint32_t *cur, uint64_t len, uint32_t layer_k, uint32_t beta_raw,
                uint32_t *next_out);

/* ---- prover.rs:101-173 ------------------------------------------------------------------ */
/* The three FibonacciSq constraints combined with alpha0..2, evaluated on the LDE domain.
 * a_first = a[0], a_last = a[n-2] (the public inputs, prover.rs:33,42). */
int zk_cp_eval_fibsq(zk_ctx *ctx, const uint32_t *f_eval, uint32_t log_n, uint32_t log_blowup,
                     const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last, uint32_t *cp_out);

/* ---- merkle.rs -------------------------------------------------------------------------- */
/* Merkle::new(size, leaves)  (merkle.rs:14-51).  heap_out: 32*(2*size-1) bytes, root first,
 * children of i at 2i+1 / 2i+2, leaves at [size-1, 2*size-1).  size: power of two, 1..2^30. */
int zk_merkle_build(zk_ctx *ctx, const uint32_t *leaves, uint64_t size, uint8_t *heap_out);
/* only the root (merkle[0], prover.rs:82): skips the device->host copy of the heap */
int zk_merkle_root(zk_ctx *ctx, const uint32_t *leaves, uint64_t size, uint8_t root_out[32]);
/* Merkle::trace(i) on a heap held by the caller (merkle.rs:54-71); host-side, returns path length */
int zk_merkle_trace(const uint8_t *heap, uint64_t size, uint64_t index, uint8_t *path_out);

/* ---- prover.rs:9-293 -------------------------------------------------------------------- */
size_t zk_transcript_len(uint32_t log_n);
/* generate_proof(Channel::new()) generalised to n = 2^log_n (the reference is log_n = 10,
 * secret = 3141592).  transcript_out receives Proof.data (channel.rs:8,34-36), state_out
 * Proof.state.  The trace is generated on the host (prover.rs:32-39 is sequential). */
int zk_prove_fibsq(zk_ctx *ctx, uint32_t log_n, uint32_t secret, uint8_t *transcript_out, size_t cap,
                   size_t *len_out, uint8_t state_out[32]);
/* The reference's protocol with the two extensions its readme.md:1 and SURVEY.md 8(f) name: a caller-chosen ("random") initial
 * channel state instead of Channel::new()'s zeros (init_state: 32 bytes, or NULL), and n_extra_queries further test points,
 * each drawn from the channel like the first (prover.rs:263) once the records of the previous one are committed: the
 * transcript is the reference's, followed n_extra_queries times by the raw query word (u32 LE) and the 4 + log_n records
 * of x = word % (8n - 16) -- zk_transcript_len_ex(log_n, n_extra_queries) bytes.  With NULL and 0 this is zk_prove_fibsq;
 * zk_verify_fibsq_ex checks such a proof (every test point against proof.rs:63-148, every challenge and test point
 * against the replayed channel when rederive != 0). */
size_t zk_transcript_len_ex(uint32_t log_n, uint32_t n_extra_queries);
int zk_prove_fibsq_ex(zk_ctx *ctx, uint32_t log_n, uint32_t secret, const uint8_t *init_state, uint32_t n_extra_queries,
                      uint8_t *transcript_out, size_t cap, size_t *len_out, uint8_t state_out[32]);
int zk_verify_fibsq_ex(uint32_t log_n, uint32_t a_last, const uint8_t *transcript, size_t len, const uint8_t *state,
                       const uint8_t *init_state, uint32_t n_extra_queries, int rederive, int *failed_check);
/* same, from a caller-provided host trace of 2^log_n - 1 values (H2D copy is part of the call) */
int zk_prove_fibsq_from_trace(zk_ctx *ctx, uint32_t log_n, const uint32_t *trace, uint8_t *transcript_out,
                              size_t cap, size_t *len_out, uint8_t state_out[32]);
/* n_proofs independent proofs of the same size; traces are n_proofs rows of 2^log_n - 1 values,
 * transcripts_out n_proofs rows of zk_transcript_len(log_n) bytes, states_out rows of 32 bytes.
 * Internally processed in device-resident sub-batches sized to the ctx memory budget. */
int zk_prove_fibsq_batch(zk_ctx *ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t *traces,
                         uint8_t *transcripts_out, uint8_t *states_out);
/* same as zk_prove_fibsq_batch but from the secrets: the traces (prover.rs:32-39) are generated inside the call, one
 * sub-batch ahead of the proving (see zk_prove_stream_submit) */
int zk_prove_fibsq_batch_secrets(zk_ctx *ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t *secrets,
                                 uint8_t *transcripts_out, uint8_t *states_out);
/* The same, pipelined: zk_prove_stream_submit starts the trace generation of a batch (host threads, one trace each, rows
 * uploaded from pinned memory on a side stream) and returns at once; zk_prove_stream_collect proves the OLDEST submitted
 * batch (at most two in flight) and fills its transcripts / states.  Submitting batch i + 1 before collecting batch i hides
 * the sequential recurrence (~7.5 ms of a host core per 2^20 trace, whatever the batch size) and the upload behind the
 * hashing of batch i.  This is what bench.py times as `e2e` (secrets in, proofs out).  zk_prove_fibsq_batch_secrets runs
 * its sub-batches through the same pipeline.  ZK_HOST_THREADS caps the worker threads (default: all hardware threads). */
int zk_prove_stream_submit(zk_ctx *ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t *secrets);
int zk_prove_stream_collect(zk_ctx *ctx, uint8_t *transcripts_out, uint8_t *states_out);
/* Proof::verify (proof.rs:15-149) generalised to n = 2^log_n; host code, no GPU needed.  a_last = a[n-2] is the
 * public input (2338775057 in the reference, proof.rs:70).  Returns ZK_OK or ZK_ERR_VERIFY with *failed_check set
 * to the proof.rs line of the assert that would have fired.  rederive != 0 additionally replays the channel and
 * checks alpha / beta / query (and `state`, when given) -- the reference trusts the values in the proof. */
int zk_verify_fibsq(uint32_t log_n, uint32_t a_last, const uint8_t *transcript, size_t len, const uint8_t *state,
                    int rederive, int *failed_check);
/* SHA-256 of a host buffer (FIPS 180-4; SHA-NI when the CPU has it): what the reference takes from sha2 0.10.6 for its
 * host-side Channel (channel.rs:21-24) and compute_root_from_path (merkle.rs:82-110), for hosts without a hash crate. */
void zk_sha256(const uint8_t *data, size_t len, uint8_t out[32]);
/* Wire format of `Proof { state, data }` (proof.rs:5-8) as bincode 1.x writes the struct: 32 state bytes, u64 LE length,
 * transcript.  Host code.  zk_proof_deserialize checks the length prefix and returns a pointer INTO `bytes`. */
size_t zk_proof_serialized_len(size_t transcript_len);
int zk_proof_serialize(const uint8_t state[32], const uint8_t *transcript, size_t transcript_len, uint8_t *out, size_t cap,
                       size_t *len_out);
int zk_proof_deserialize(const uint8_t *bytes, size_t len, uint8_t state_out[32], const uint8_t **transcript_out,
                         size_t *transcript_len_out);
/* prover.rs:32-39: a[0] = 1, a[1] = secret, a[i] = a[i-2]^2 + a[i-1]^2; 2^log_n - 1 values (host) */
int zk_trace_fibsq(uint32_t log_n, uint32_t secret, uint32_t *trace_out);
/* releases the proving work space, the scratch buffers and the per-size composition-polynomial tables (tables_too != 0:
 * also the power tables); everything is re-created on demand.  Synchronises the context's streams. */
int zk_ctx_trim(zk_ctx *ctx, int tables_too);
/* the same for n_proofs secrets at once (host): groups of 8 traces advance together as SIMD lanes (AVX2), the groups are
 * spread over `threads` host threads (0: all).  out: n_proofs rows of 2^log_n - 1 values.  This is what
 * zk_prove_stream_submit and zk_prove_fibsq_batch_secrets run one batch ahead of the proving. */
int zk_trace_fibsq_batch(uint32_t log_n, const uint32_t *secrets, uint32_t n_proofs, uint32_t threads, uint32_t *out);

/* ---- the program being proven as data.  The reference hard-codes FibonacciSq (trace loop prover.rs:32-39,
 * transition numerator prover.rs:134-138, verifier proof.rs:66-70); a zk_air describes a one-column program over a
 * three-row window with the reference's two boundary constraints (first value, value n - 2):
 *   a[0] = a0, a[1] = secret, a[i+2] = c20 a[i+1]^2 + c02 a[i]^2 + c11 a[i] a[i+1] + c10 a[i+1] + c01 a[i] + c00.
 * The coefficients are u32 reduced mod p like every value that enters the field (field.rs:20-24).  The transition degree
 * is at most 2, so the composition polynomial keeps degree < n and every other step of the protocol (domains, folds,
 * transcript layout and length) is the reference's.  {1, 1,1,0,0,0,0} is FibonacciSq and runs the specialised code unless
 * ZK_AIR_NO_SPECIALISE is set (testing: the general code path on the reference's program must give the reference's
 * transcript).  zk_ctx_set_air (null: FibonacciSq) applies to every prove entry point of the context -- from secrets, from
 * traces, batched, pipelined, sharded -- and to zk_cp_eval_fibsq[_dev]; traces handed in must satisfy the program. */
typedef struct zk_air { uint32_t a0, c20, c02, c11, c10, c01, c00, flags; } zk_air;
#define ZK_AIR_NO_SPECIALISE 1u
int zk_ctx_set_air(zk_ctx *ctx, const zk_air *air);
int zk_ctx_get_air(const zk_ctx *ctx, zk_air *out);
/* the trace of the program (host; SIMD lanes over proofs like zk_trace_fibsq_batch): rows of 2^log_n - 1 values */
int zk_trace_air(const zk_air *air, uint32_t log_n, uint32_t secret, uint32_t *trace_out);
int zk_trace_air_batch(const zk_air *air, uint32_t log_n, const uint32_t *secrets, uint32_t n_proofs, uint32_t threads,
                       uint32_t *out);
/* zk_verify_fibsq_ex for a proof of the program `air` (null: FibonacciSq) */
int zk_verify_air(const zk_air *air, uint32_t log_n, uint32_t a_last, const uint8_t *transcript, size_t len,
                  const uint8_t *state, const uint8_t *init_state, uint32_t n_extra_queries, int rederive, int *failed_check);
/* cap on device bytes a ctx may hold for proving work space (default: 70% of free memory) */
int zk_ctx_set_memory_budget(zk_ctx *ctx, size_t bytes);

/* ---- device-pointer variants (no host copies, no synchronize; pointers are CUDA device memory,
 * 16-byte aligned).  Used by bench.py for HBM-resident timing and by multi-GPU drivers. ------ */
int zk_interpolate_subgroup_dev(zk_ctx *ctx, uint32_t *y_inout_n, uint32_t n_points, uint32_t log_n,
                                uint32_t *coeffs_out);
int zk_lde_coset_dev(zk_ctx *ctx, const uint32_t *coeffs, uint32_t log_n, uint32_t log_blowup,
                     uint32_t shift, uint32_t *evals_out);
int zk_fri_fold_dev(zk_ctx *ctx, const uint32_t *cur, uint64_t len, uint32_t layer_k, uint32_t beta_raw,
                    uint32_t *next_out);
int zk_cp_eval_fibsq_dev(zk_ctx *ctx, const uint32_t *f_eval, uint32_t log_n, uint32_t log_blowup,
                         const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last, uint32_t *cp_out);
int zk_merkle_build_dev(zk_ctx *ctx, const uint32_t *leaves, uint64_t size, uint8_t *heap_out);
/* whole prove() with the traces already resident: traces_dev is n_proofs rows of 2^log_n words
 * (word n-1 of each row is scratch); transcripts/states are HOST buffers filled after one
 * synchronize.  This is the region bench.py times for `value`. */
int zk_prove_fibsq_resident(zk_ctx *ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t *traces_dev,
                            uint8_t *transcripts_out, uint8_t *states_out);

/* ---- further query openings (the reference opens ONE position, prover.rs:263-289; soundness wants many).
 * The arrays and heaps of the last prove call on this context are still resident: for every position xs[q] < 8n - 16
 * of its proof `proof_index`, writes the same 4 + log_n records the prover commits for its own query
 * (f at x, x+8, x+16 and cp at x with their paths; every FRI layer at x_i and x_i + len/2 with theirs), in the same
 * bincode layout, zk_decommit_len(log_n) bytes per position (out: n_queries * zk_decommit_len bytes, host memory;
 * log_n must be the resident proof's size and out_cap the capacity of `out`: ZK_ERR_INVALID / ZK_ERR_BUFFER otherwise).
 * Spliced behind the head of the transcript with the query word replaced, each opening passes the reference's own
 * checks (proof.rs:63-148). */
size_t zk_decommit_len(uint32_t log_n);
int zk_open_queries(zk_ctx *ctx, uint32_t log_n, uint32_t proof_index, const uint32_t *xs, uint32_t n_queries, uint8_t *out,
                    size_t out_cap);

/* ---- one proof sharded over several GPUs (BASELINE configs[4]; driver: zkstark_b200/sharded.py) ------------
 * Ownership of every evaluation array is blocked-cyclic: global index i lives on rank (i >> log_block) % n_ranks
 * at local index ((i >> log_block) / n_ranks) << log_block | (i & (2^log_block - 1)).  FRI folds (partner
 * i + len/2) and the bottom log_block Merkle levels are then rank-local; what crosses NVLink is one all-to-all
 * after the LDE (cosets -> blocks), the block roots of every tree and a 16-value halo per block for the CP.
 * All pointers are device memory; calls enqueue on the ctx stream.  state_dev: 256-byte per-proof Fiat-Shamir
 * state kept on the device and replicated on every rank; transcript_dev: zk_transcript_len bytes (+64). */
int zk_shard_state_init_dev(zk_ctx *ctx, void *state_dev, uint32_t a_first, uint32_t a_last);
/* out[j] = f(5 h^(8 j + coset)), j < 2^log_n: one of the 8 coset transforms of the LDE; work: 2^log_n scratch */
int zk_shard_lde_coset_dev(zk_ctx *ctx, const uint32_t *coeffs, uint32_t log_n, uint32_t coset, uint32_t *work,
                           uint32_t *out);
/* Merkle levels of the rank-local array up to its 2^log_block-leaf block roots, in heap order of the local
 * array: the block roots are the local_len >> log_block digests of heap level log2(local_len) - log_block */
int zk_shard_merkle_blocks_dev(zk_ctx *ctx, const uint32_t *vals_local, uint64_t local_len, uint32_t log_block,
                               uint8_t *heap_local);
/* upper part of a tree: upper_heap holds 2^log_count digests at heap level log_count (all block roots, global
 * order); builds the levels above them, then runs the channel step post_op (0 none, 1 f root + alphas,
 * 2 cp root + beta_0, 3 FRI layer `fri_layer` root + next beta) on state_dev / transcript_dev */
int zk_shard_upper_tree_dev(zk_ctx *ctx, uint8_t *upper_heap, uint32_t log_count, int post_op, uint32_t fri_layer,
                            uint32_t n_layers, void *state_dev, uint8_t *transcript_dev);
/* composition polynomial on the rank-local points; halo: 16 values per local block = head of the next block */
int zk_shard_cp_dev(zk_ctx *ctx, const uint32_t *f_local, const uint32_t *halo, uint64_t local_len, uint32_t log_n,
                    uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void *state_dev, uint32_t *cp_local);
/* FRI fold of the rank-local part of layer layer_k (beta from state_dev); log_N0 = log2 of the LDE size */
int zk_shard_fold_dev(zk_ctx *ctx, const uint32_t *cur_local, uint64_t local_len, uint32_t layer_k, uint32_t log_N0,
                      uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void *state_dev, uint32_t *next_local);
/* zk_shard_cp_dev followed by zk_shard_merkle_blocks_dev of the composition layer, fused in one pass */
int zk_shard_cp_blocks_dev(zk_ctx *ctx, const uint32_t *f_local, const uint32_t *halo, uint64_t local_len,
                           uint32_t log_n, uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void *state_dev,
                           uint32_t *cp_local, uint8_t *heap_local);
/* zk_shard_fold_dev followed by zk_shard_merkle_blocks_dev of the folded layer, fused in one pass (the fold feeds
 * the leaf hashing): next_local gets local_len/2 values, heap_local their block subtrees */
int zk_shard_fold_blocks_dev(zk_ctx *ctx, const uint32_t *cur_local, uint64_t local_len, uint32_t layer_k,
                             uint32_t log_N0, uint32_t n_ranks, uint32_t rank, uint32_t log_block,
                             const void *state_dev, uint32_t *next_local, uint8_t *heap_local);
/* replicated tail once a layer is too small to shard: layer first_layer is complete at layers_arena[0..);
 * layers first_layer+1..L (values, heaps at 64 bytes per value of the preceding layers) are appended */
int zk_shard_fri_tail_dev(zk_ctx *ctx, uint32_t log_n, uint32_t first_layer, uint32_t *layers_arena,
                          uint8_t *heaps_arena, void *state_dev, uint8_t *transcript_dev);

/* prover.rs:263-289 (query decommitment) for the sharded proof.  tree_ptrs: HOST array of 3 * (first_tail + 2)
 * device pointers {rank-local values, rank-local heap, replicated upper heap} of the f tree and of FRI layers
 * 0..first_tail; tail_*: the arenas of zk_shard_fri_tail_dev.  Each rank writes only the values / sibling digests
 * it owns into transcript_dev (rank 0 also the length prefixes and the replicated tail layers); the decommitment
 * region must be zero on entry, and the byte-wise maximum over ranks (one all-reduce) completes the transcript. */
int zk_shard_decommit_dev(zk_ctx *ctx, uint32_t log_n, uint32_t n_ranks, uint32_t rank, uint32_t log_block,
                          uint32_t first_tail, const void *const *tree_ptrs, const uint32_t *tail_vals,
                          const uint8_t *tail_heaps, const void *state_dev, uint8_t *transcript_dev);

/* ---- the sharded prover as ONE call (csrc/sharded.cu): the host side of the scheme above -- ownership, the all-to-all,
 * the root all-gathers, the halo ring, the replicated tail, the decommitment all-reduce and the channel tail -- in C++
 * over NCCL, everything enqueued on the ctx stream with one synchronisation at the end.  One process per GPU:
 *   rank 0: zk_shard_unique_id(id), hand `id` to the other processes by any means (file, socket, MPI, torch.distributed)
 *   every rank: zk_shard_comm_init(ctx, id, n_ranks, rank, &comm)            -- ncclCommInitRank; collective
 *   every rank: zk_prove_fibsq_sharded(comm, log_n, trace, ...)              -- collective; every rank gets the proof
 * NCCL is bound at run time (libnccl.so.2; ZK_NCCL_LIB overrides the path): single-GPU users do not need it.
 * Ranks that share a node exchange the LDE through peer memory: every rank reads the evaluations it owns straight out of
 * the other ranks' buffers over NVLink (cudaIpc peer mappings; ZK_SHARD_NO_P2P=1 or a failed mapping falls back to an NCCL
 * all-to-all on all ranks).
 * zk_shard_comm_init_local runs G VIRTUAL ranks on the context's own device (collectives become device copies): the
 * same driver and kernels on one GPU, used by the single-GPU test tier.  n_ranks: 1, 2, 4 or 8 (the LDE has 8 cosets).
 * A comm borrows its ctx: destroy the comm first.  With real ranks zk_shard_comm_destroy is collective as well (like
 * ncclCommDestroy): buffers a peer has mapped are freed only after every rank has closed its mappings.  Generalises generate_proof (prover.rs:9) like zk_prove_fibsq. */
#define ZK_SHARD_ID_BYTES 128
typedef struct zk_shard_comm zk_shard_comm;
int zk_shard_unique_id(uint8_t id_out[ZK_SHARD_ID_BYTES]);
int zk_shard_comm_init(zk_ctx *ctx, const uint8_t id[ZK_SHARD_ID_BYTES], uint32_t n_ranks, uint32_t rank, zk_shard_comm **out);
int zk_shard_comm_init_local(zk_ctx *ctx, uint32_t n_virtual_ranks, zk_shard_comm **out);
void zk_shard_comm_destroy(zk_shard_comm *comm);
/* tuning: log_block (0 keeps it; default: 13, or half the rank's share of the evaluations when that is smaller), log2 of the largest layer that finishes in the replicated tail
 * (0xffffffff keeps it; default 18; 0 shards every layer that can be sharded) */
int zk_shard_comm_set(zk_shard_comm *comm, uint32_t log_block, uint32_t log_tail_len);
/* trace: HOST memory, 2^log_n - 1 values, identical on every rank (each rank uploads only its 1/n_ranks slice).
 * transcript_out / state_out: as zk_prove_fibsq, on every rank. */
int zk_prove_fibsq_sharded(zk_shard_comm *comm, uint32_t log_n, const uint32_t *trace, uint8_t *transcript_out, size_t cap,
                           size_t *len_out, uint8_t state_out[32]);
/* device time of the last sharded proof by phase, ms: [0] upload + trace all-gather + interpolation, [1] this rank's LDE
 * cosets, [2] all-to-all, [3] unpack, [4] f tree, [5] halo (ring exchange; none in peer mode), [6] cp + tree, [7] sharded FRI layers, [8] gather +
 * replicated FRI tail, [9] decommit + all-reduce, [10] download, [11] total */
int zk_shard_last_phase_ms(const zk_shard_comm *comm, float out_ms[12]);
/* bytes this rank received from other ranks during the last sharded proof */
uint64_t zk_shard_last_comm_bytes(const zk_shard_comm *comm);
/* index of the first FRI layer that ran in the replicated tail */
uint32_t zk_shard_last_first_tail(const zk_shard_comm *comm);

/* ---- K1 on its own: the device field arithmetic (csrc/field.cuh, replacing Gf<P> over num-modular, field.rs:96-211)
 * applied element-wise to host arrays a, b of n raw u32 words, each reduced first as F::from(u32) does (field.rs:20-24).
 * out: 6 rows of n canonical residues -- a + b, a - b, a * b, -a, a^-1 (0 for a = 0), a^b (exponent: the raw word b,
 * Pow<u32>, field.rs:26-38).  For parity tests against the scalar Rust type; not on the proving path. */
int zk_field_ops(zk_ctx *ctx, const uint32_t *a, const uint32_t *b, size_t n, uint32_t *out);

/* ---- measurement support (bench.py): runs ONE hot-path kernel (or kernel group) `iters` times on
 * the ctx stream, each launch bracketed by CUDA events (optionally after an L2 flush that is not
 * timed), on synthetic canonical residues it generates on the device.  ms_out[0] = mean ms per
 * launch, ms_out[1] = min ms.  which:
 *   0 Merkle leaf pass only (tree_pass_kernel<values>, 2^log_size leaves, 5 levels per thread)
 *   1 whole Merkle tree over 2^log_size leaves (all passes + top)
 *   2 coset LDE, all passes (log_size = log_n)          3 interpolation, all passes (log_n)
 *   4 FRI fold of 2^log_size values (layer 0)            5 CP evaluation (log_size = log_n)
 *   6 integer-pipe microbenchmark: ms_out[2] = achieved Tiop/s; log_size selects the mix:
 *     0 = ALU pipe (SHF/LOP3/IADD3), 1 = FMA pipe (IMAD), 2 = 1:1 ALU/IMAD
 *   8 butterfly microbenchmark (the instruction roofline of the transforms): radix-16 register rounds of
 *     (u, v) -> (u + v, (u - v) w) without memory traffic; ms_out[2] = achieved tera-butterflies/s; log_size selects
 *     0 = whole butterflies, 1 = only the Montgomery products, 2 = only the additions / subtractions
 *   7 node-hash latency probe: one CTA of `log_size` warps hashes a chain of `iters` nodes;
 *     ms_out[2] = SM cycles per node hash                                                        */
/* tuning knobs of zk_bench_kernel(0): SHA-256 pipe-balancing variant (-1 = production) and levels per thread */
int zk_bench_set(zk_ctx *ctx, int sha_variant, uint32_t leaf_pass_k);
/* trees of up to 2^log leaves (x batch) are built leaves -> root by the tile kernel in one launch (process-wide) */
int zk_bench_set_tile_direct_log(zk_ctx *ctx, uint32_t log_leaves);
int zk_bench_kernel(zk_ctx *ctx, int which, uint32_t log_size, uint32_t iters, int flush_l2, float ms_out[3]);

#ifdef __cplusplus
}
#endif
#endif
