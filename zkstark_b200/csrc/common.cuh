// Shared declarations for the hot-path kernels: per-proof device state, power tables,
// kernel parameter blocks and host-side launcher prototypes.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "field.cuh"

namespace zk {

constexpr uint32_t LOG_BLOWUP = 3;  // 8192/1024 (prover.rs:56-57); literal +8/+16 at prover.rs:268-271
constexpr uint32_t MAX_LOG_N = 27;  // N = 2^30 is the largest power-of-two subgroup of F_p^*
constexpr uint32_t TILE_LOG = 11;   // the tile kernel reduces up to 2^TILE_LOG inputs per CTA, two per thread
constexpr uint32_t TILE_PAIRS = 1u << (TILE_LOG - 1);
constexpr uint32_t TILE_MIN_LOG = 7;
#ifndef ZK_TILE_SHA_V
#define ZK_TILE_SHA_V 515  // SHA-256 variant of the wide (one warp per 32 nodes) levels of the tile kernel
#endif
#ifndef ZK_TILE_DIRECT_LOG
#define ZK_TILE_DIRECT_LOG 17  // trees up to this many leaves (x batch) go leaves -> root in one tile-kernel launch
#endif
#ifndef ZK_TILE_SHIFT
#define ZK_TILE_SHIFT 7         // the tile kernel splits a tree into 2^7 = 128 tiles: one CTA per SM (148 SMs)
#endif
#ifndef ZK_PASS_LOG_THREADS
#define ZK_PASS_LOG_THREADS 18  // the leaf pass runs about this many threads (log2), each hashing a 2^k-leaf subtree
#endif
#ifndef ZK_COOP_V
#define ZK_COOP_V 96                    // SHA-256 variant of the main warp of a cooperative pair: latency form of the round, K+W on the FMA pipe
#endif
#ifndef ZK_COOP_HELPER_V
#define ZK_COOP_HELPER_V ZK_TILE_SHA_V  // ... and of its helper warp
#endif
#ifndef ZK_COOP_CNT
#define ZK_COOP_CNT 128         // tile kernel: levels with <= this many nodes hash each node with a warp pair (0: off)
#endif
#ifndef ZK_PASS_THREADS
#define ZK_PASS_THREADS 256
#endif
constexpr uint32_t PASS_THREADS = ZK_PASS_THREADS;
#ifndef ZK_PASS_MINBLOCKS
#define ZK_PASS_MINBLOCKS 5       // 256 x 5: <= 51 registers, no spills in any source variant; 1354 -> 1333 us per 2^23 leaves (swept 32..1024 x 1..16)
#endif
constexpr uint32_t MAX_PASS_K = 6;
#ifndef ZK_EPILOGUE_LEVELS
#define ZK_EPILOGUE_LEVELS 8
#endif
#ifndef ZK_PASS_MAX_K
#define ZK_PASS_MAX_K 6
#endif
#ifndef ZK_SHA_UNI
#define ZK_SHA_UNI 2  // 1: leaf and node hashes share one copy of the data-block compression; 2: + leaf-specialised prologue and schedule
#endif
#ifndef ZK_SHA_V
#define ZK_SHA_V 4611  // SHA-256 variant of the throughput-bound kernels (sha256.cuh): adds on the FMA pipe, six-add round
#endif

// Two-level table of powers of h (generator of the order-N LDE subgroup, prover.rs:49):
//   h^e = lo[e & (2^lo_bits-1)] * hi[e >> lo_bits],   both halves in Montgomery form.
// Inverse powers are read at the negated exponent (h^-e = h^(N-e)).
struct PowTab {
    const uint32_t* lo;
    const uint32_t* hi;
    uint32_t lo_bits;
    uint32_t log_N;
};

// Per-proof Fiat-Shamir state kept on the device between kernels (channel.rs:6-9) plus
// the challenges already drawn and the public inputs of the FibonacciSq statement.
struct ProofState {
    uint32_t chan[8];       // channel.rs:7 `state`, as native words H0..H7
    uint32_t alpha_raw[3];  // prover.rs:163-165 (un-reduced, channel.rs:28-32)
    uint32_t free_term;     // prover.rs:254
    uint32_t beta_raw[32];  // prover.rs:200
    uint32_t query_raw;     // prover.rs:263
    uint32_t a_first;       // a[0]   (prover.rs:33)
    uint32_t a_last;        // a[n-2] (prover.rs:42)
    uint32_t chan_at_query[8];  // channel state right after the query draw (host finishes the tail)
    uint32_t pad[9];
};
static_assert(sizeof(ProofState) == 64 * 4, "ProofState layout");

// What the tree-top kernel does with the root it just computed (prover.rs commit order).
enum PostOp : int {
    POST_NONE = 0,
    POST_F = 1,    // prover.rs:85 commit f root; :163-165 draw alpha0..2
    POST_CP = 2,   // prover.rs:180 commit cp root; :200 draw beta_0
    POST_FRI = 3,  // prover.rs:224 commit layer root; next beta, or (:254,:263) free term + query
};

enum SrcKind : int { SRC_VALUES = 0, SRC_DIGESTS = 1, SRC_FOLD = 2, SRC_CP = 3, SRC_CP_TAB = 4 /* SRC_CP with CpArgs::tab */ };

// Blocked-cyclic ownership of an evaluation array across ranks (single proof sharded over GPUs):
// global index i lives on rank (i >> log_block) % n_ranks at local index
//   ((i >> log_block) / n_ranks) << log_block | (i & (2^log_block - 1)).
// Keeps every FRI fold (partner i + len/2) and the bottom log_block Merkle levels rank-local.
struct ShardMap {
    uint32_t n_ranks;    // 0 or 1: not sharded
    uint32_t rank;
    uint32_t log_block;
    uint32_t pad;
};
__host__ __device__ inline uint32_t shard_global(const ShardMap& m, uint32_t local)
{
    if (m.n_ranks <= 1) return local;
    return ((((local >> m.log_block) * m.n_ranks) + m.rank) << m.log_block) | (local & ((1u << m.log_block) - 1u));
}

// Store layout of the last pass of a ONE-coset LDE transform when the evaluations go straight into the all-to-all send
// buffer of the sharded prover (csrc/sharded.cu): point j of the transform (j < n) belongs to global block j / jb,
// which lives on rank d = (j / jb) % G as its local block bl = j / jb / G; the send buffer is [d][slot][bl][tt].
struct NttPack {
    uint32_t log_jb;     // log2 of the points of ONE coset per block (= log_block - 3); 0: plain natural-order store
    uint32_t log_G;      // log2 of the number of ranks
    uint32_t log_nbl;    // log2 of the blocks per rank
    uint32_t slot;       // which of this rank's n_slots coset transforms this is
    uint32_t n_slots;    // 8 / G
};
__host__ __device__ inline size_t pack_index(const NttPack& p, uint32_t j)
{
    const uint32_t tt = j & ((1u << p.log_jb) - 1u), b = j >> p.log_jb;
    const uint32_t d = b & ((1u << p.log_G) - 1u), bl = b >> p.log_G;
    return ((((size_t)d * p.n_slots + p.slot) << p.log_nbl) + bl) << p.log_jb | tt;
}
// Fused leaf sources.  `vals` is where leaf values live (or are produced into).
struct FoldArgs {           // polynomial.rs:385-400 + prover.rs:204-211 in evaluation form
    const uint32_t* cur;    // layer k, len = 2*half (batch stride cur_stride)
    uint64_t cur_stride;
    uint32_t layer_k;       // index of `cur`
    uint32_t log_half;      // log2(len/2)
    uint32_t winv2k_m;      // Montgomery form of (5^(2^k))^-1
    uint32_t inv2_m;        // Montgomery form of 2^-1
    int beta_from_state;    // read beta_raw[layer_k] from ProofState, else use beta_raw below
    uint32_t beta_raw;
    ShardMap shard;         // local -> global index (domain point) when the layer is sharded
};
// The transition constraint of the program being proven (hostair.hpp): the numerator is
//   f(g^2 x) - [c20 f(gx)^2 + c02 f(x)^2 + c11 f(x) f(gx) + c10 f(gx) + c01 f(x) + c00].
// generic == 0: the reference's FibonacciSq (prover.rs:134-138), the specialised expression; the fields are unused.
struct AirCoef {
    uint32_t generic;
    uint32_t c20m, c02m, c11m;   // c R (Montgomery form): mont_mul(mont_mul(u, v), cm) = c u v / R
    uint32_t c10, c01;           // canonical:             mont_mul(u, c)               = c u / R
    uint32_t c00r;               // c00 / R
};

struct CpArgs {             // prover.rs:101-173 pointwise
    const uint32_t* f;      // f_eval, N values (batch stride f_stride)
    uint64_t f_stride;
    uint32_t log_n;
    uint32_t g1, g2, g3;    // g^(n-1), g^(n-2), g^(n-3) canonical
    uint32_t zinv_r6[8];    // 1/(x^n - 1) for i mod 8, times R^6 (see cp_chunk: the product with alpha_2 lands at R^5)
    int alpha_from_state;
    uint32_t alpha_raw[3];
    uint32_t a_first, a_last;  // used when !alpha_from_state (stand-alone API)
    ShardMap shard;            // sharded: f is the rank-local array, x comes from the global index
    const uint32_t* halo;      // sharded: first 16 values of the block that follows each local block
    uint32_t halo_roll;        // halo row of local block bl is (bl + halo_roll) % n_blocks: the heads arrive from rank + 1 as
    uint32_t halo_mask;        //   ITS local blocks; on the last rank the following block is rank 0's NEXT one (halo_mask = n_blocks - 1)
    uint32_t log_len;          // log2 of the number of local points (== log_N when not sharded)
    // Optional per-size constant tables (built once per context and size by launch_cp_tables; they depend on the
    // domain only, like twiddles): tab[i] = 1/(x_i - 1) R, tab[N + i] = 1/(x_i - g^(n-2)) R,
    // tab[2N + i] = (x_i - g^(n-3))(x_i - g^(n-2))(x_i - g^(n-1)) / (x_i^n - 1) R^2, x_i = 5 h^i.  Null: computed on the fly.
    const uint32_t* tab;
    AirCoef air;               // the transition constraint (zk_ctx_set_air); zero = the reference's program
};

// Launch-shape knobs of the tree kernels.  One copy per context (zk_ctx::tune; zk_bench_set* edit it), so contexts of
// different threads / GPUs in one process never see each other's experiments.
struct TreeTune {
    uint32_t tile_direct_log = ZK_TILE_DIRECT_LOG;   // trees up to 2^this leaves (x batch) skip the per-thread-subtree leaf pass
    uint32_t tile_log_override = 0;                  // forced tile size of the tile kernel (0 = adaptive)
    uint32_t tile_shift = ZK_TILE_SHIFT;             // the tile kernel runs 2^this tiles per tree
    uint32_t pass_log_threads = ZK_PASS_LOG_THREADS; // the leaf pass runs ~2^this threads (summed over the batch)
    uint32_t coop_cnt = ZK_COOP_CNT;                 // tile kernel: widest level hashed by warp pairs
    uint32_t pass_epilogue = ZK_EPILOGUE_LEVELS;     // CTA epilogue of the pass kernel: at most this many levels (0 = off)
    uint32_t pass_max_k = ZK_PASS_MAX_K;             // deepest per-thread subtree of the leaf pass
};
TreeTune default_tree_tune();   // compile-time defaults overridden by ZK_EPILOGUE_LEVELS / ZK_PASS_MAX_K in the environment

struct TreeArgs {
    const TreeTune* tune;   // host side only (null: defaults)
    uint32_t* vals;         // leaf values of this tree (read, or written by a fused source)
    uint64_t vals_stride;   // elements between consecutive proofs of the batch
    uint8_t* heap;          // 32*(2*size-1) bytes per proof, root first (merkle.rs:14-51)
    uint64_t heap_stride;   // bytes
    uint32_t log_size;      // tree has 2^log_size leaves
    uint32_t in_level;      // level the kernel consumes (== log_size for leaf passes)
    uint32_t k;             // levels produced by one thread of a pass kernel
    uint32_t ep_levels;     // pass kernel: further levels reduced per CTA through shared memory (0..log2 PASS_THREADS; set by the launcher)
    uint32_t log_tile;      // tile kernel: inputs per CTA (set by the launcher)
    uint32_t coop_cnt;      // tile kernel: levels of at most this many nodes use two warps per node (set by the launcher)
    ProofState* state;      // per proof (may be null for POST_NONE)
    uint32_t* counters;     // per proof tile-arrival counter (zero between launches); needed above 2^TILE_LOG inputs
    uint8_t* transcript;    // per proof
    uint64_t transcript_stride;
    int post;               // PostOp
    uint32_t fri_layer;     // POST_FRI: index i (1..L) of the layer whose tree this is
    uint32_t n_layers;      // L
    PowTab pow;
    FoldArgs fold;
    CpArgs cp;
};

// ---- programmatic dependent launch.  Every kernel of the prove pipeline starts with pdl_wait() (griddepcontrol.wait:
// until the previous kernel of the stream has completed and its memory is visible) and is launched through zk_launch()
// with the programmatic-serialization attribute, so its CTAs may become resident while the previous kernel is still
// running and the ~2-3 us launch/drain gap between dependent kernels shrinks to the wake-up of resident CTAs.
// pdl_trigger() (griddepcontrol.launch_dependents) is what allows the NEXT kernel in: the small NTT / element-wise
// kernels call it first thing; the tree kernels only in the closing CTA just before the channel step, because CTAs
// that become resident early are packed onto whatever SMs are free at that moment and the latency-bound tile kernel
// wants its 128 CTAs on 128 different SMs (measured: early trigger there costs +11 us per tree).
// Without the launch attribute (ZK_NO_PDL=1) both instructions are no-ops and the stream serialises as usual.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter()
{
    pdl_trigger();
    pdl_wait();
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t zk_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- launchers (tree.cu)
// Builds the whole tree over 2^log_size leaves taken from `src_kind`, per proof of the batch.
cudaError_t launch_merkle_tree(cudaStream_t s, int src_kind, TreeArgs a, uint32_t batch, int* n_launches);
// FRI layers a.fri_layer .. last_layer in one launch (a describes the first of them; every one is a single tile)
cudaError_t launch_fri_tail(cudaStream_t s, TreeArgs a, uint32_t batch, uint32_t last_layer);
// levels log_size .. stop_level only (no top kernel, no channel step): used for the rank-local block subtrees
cudaError_t launch_merkle_partial(cudaStream_t s, int src_kind, TreeArgs a, uint32_t batch, uint32_t stop_level, int* n_launches);
// upper part of a tree whose level `in_level` digests are already in the heap: digest passes + top (+ channel)
cudaError_t launch_merkle_upper(cudaStream_t s, TreeArgs a, uint32_t batch, int* n_launches);
// the leaf pass alone (bench support): a.in_level / a.k preset by the caller
cudaError_t launch_tree_leaf_pass(cudaStream_t s, TreeArgs a, uint32_t batch, int variant);
void set_tile_direct_log(TreeTune& t, uint32_t v);
cudaError_t sha_init_constants();
cudaError_t launch_node_latency(cudaStream_t s, uint32_t* out, uint32_t reps, uint32_t warps, int variant, long long* cycles);
cudaError_t launch_decommit(cudaStream_t s, uint32_t log_n, const uint32_t* f_eval, uint64_t f_stride,
                            const uint8_t* f_heap, uint64_t f_heap_stride,
                            const uint32_t* layers, uint64_t layers_stride,
                            const uint8_t* layer_heaps, uint64_t layer_heaps_stride,
                            ProofState* state, uint8_t* transcript, uint64_t transcript_stride,
                            uint32_t batch);
// Decommitment of a proof sharded blocked-cyclically (zkstark_b200/sharded.py): tree 0 is f, tree 1 + i is FRI layer i
// (i <= first_tail, committed distributed); layers above first_tail live replicated in the tail arenas.
constexpr uint32_t SHARD_MAX_TREES = 30;
struct ShardDecommitArgs {
    uint32_t log_n, n_ranks, rank, log_block, first_tail;
    const uint32_t* vals[SHARD_MAX_TREES];   // rank-local values
    const uint8_t* heap[SHARD_MAX_TREES];    // rank-local heap (levels up to the block roots)
    const uint8_t* upper[SHARD_MAX_TREES];   // replicated heap over the block roots
    const uint32_t* tail_vals;
    const uint8_t* tail_heaps;
    const ProofState* state;
    uint8_t* out;                            // transcript buffer; only the records this rank owns are written
    uint32_t flag_off;                       // byte offset in `out` (inside its 64-byte slack behind the transcript) of a u32 that
                                             // is set non-zero when the last FRI layer is not constant (prover.rs:238)
};
cudaError_t launch_shard_decommit(cudaStream_t s, const ShardDecommitArgs& a);
cudaError_t launch_open_queries(cudaStream_t s, uint32_t log_n, const uint32_t* f_eval, const uint8_t* f_heap,
                                const uint32_t* layers, const uint8_t* layer_heaps, ProofState* state,
                                const uint32_t* xs_dev, uint32_t n_q, uint8_t* out_dev, uint64_t out_stride);
struct ChanSeed { uint32_t w[8]; };   // initial channel state as native words H0..H7 (zeros: Channel::new(), channel.rs:12-17)
cudaError_t launch_state_init(cudaStream_t s, ProofState* st, const uint32_t* traces, uint64_t trace_stride,
                              uint32_t log_n, uint32_t batch, ChanSeed seed);
// stand-alone element-wise stages (stages in isolation, for the per-function C-ABI)
cudaError_t launch_fold(cudaStream_t s, const uint32_t* cur, uint32_t* next, uint32_t log_half,
                        FoldArgs f, PowTab pow, uint32_t batch, uint64_t next_stride, const ProofState* st = nullptr);
cudaError_t launch_cp(cudaStream_t s, uint32_t* cp_out, CpArgs c, PowTab pow, uint32_t batch,
                      uint64_t out_stride, const ProofState* st = nullptr);
cudaError_t launch_cp_tables(cudaStream_t s, uint32_t* tab, CpArgs c, PowTab pow);
cudaError_t launch_chan_tail(cudaStream_t s, ProofState* states, const uint8_t* transcripts, uint64_t stride,
                             uint32_t log_n, uint32_t batch);
cudaError_t launch_state_set(cudaStream_t s, ProofState* st, uint32_t a_first, uint32_t a_last);

// offsets inside the bincode transcript (channel.rs:19-26 appends in commit order)
inline size_t tr_off_alpha() { return 32; }
inline size_t tr_off_cp_root() { return 44; }
inline size_t tr_off_beta(uint32_t i) { return 76 + 36 * (size_t)i; }
inline size_t tr_off_fri_root(uint32_t i /*1..L*/) { return 80 + 36 * (size_t)(i - 1); }
inline size_t tr_off_free(uint32_t L) { return 76 + 36 * (size_t)L; }
inline size_t tr_off_query(uint32_t L) { return 80 + 36 * (size_t)L; }
inline size_t tr_off_decommit(uint32_t L) { return 84 + 36 * (size_t)L; }
inline size_t transcript_len(uint32_t k)
{
    size_t len = tr_off_decommit(k) + 4 * (12 + 32 * (size_t)(k + LOG_BLOWUP));
    for (uint32_t i = 0; i < k; i++) len += 24 + 64 * (size_t)(k + LOG_BLOWUP - i);
    return len;
}
// byte offset of layer i's heap inside the per-proof layer-heap arena; layer i has N>>i leaves
inline size_t layer_heap_off(uint32_t log_N, uint32_t i)
{
    size_t off = 0;
    for (uint32_t j = 0; j < i; j++) off += 32 * (((size_t)2 << (log_N - j)) - 1) + 32;  // +32 keeps 64B alignment
    return off;
}
inline size_t layer_vals_off(uint32_t log_N, uint32_t i)
{
    size_t off = 0;
    for (uint32_t j = 0; j < i; j++) off += (size_t)1 << (log_N - j);
    return off;
}

}  // namespace zk
