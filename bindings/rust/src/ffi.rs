//! `extern "C"` view of include/zkstark_b200.h: every export, generated from the header by tools/gen_rust_ffi.py
//! (there is no bindgen binary in the build image).  Safe wrappers live in gpu.rs.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct zk_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct zk_shard_comm {
    _private: [u8; 0],
}
/// The program being proven as data (include/zkstark_b200.h `zk_air`): a[0] = a0, a[1] = secret,
/// a[i+2] = c20 a[i+1]^2 + c02 a[i]^2 + c11 a[i] a[i+1] + c10 a[i+1] + c01 a[i] + c00.  FIBSQ is the reference's program.
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct zk_air {
    pub a0: u32,
    pub c20: u32,
    pub c02: u32,
    pub c11: u32,
    pub c10: u32,
    pub c01: u32,
    pub c00: u32,
    pub flags: u32,
}
pub const ZK_AIR_FIBSQ: zk_air = zk_air { a0: 1, c20: 1, c02: 1, c11: 0, c10: 0, c01: 0, c00: 0, flags: 0 };
pub const ZK_AIR_NO_SPECIALISE: u32 = 1;

pub const ZK_OK: c_int = 0;
pub const ZK_ERR_VERIFY: c_int = -6;
pub const ZK_SHARD_ID_BYTES: usize = 128;

extern "C" {
    // BEGIN GENERATED (tools/gen_rust_ffi.py from include/zkstark_b200.h)
    pub fn zk_ctx_create(device: c_int, out: *mut *mut zk_ctx) -> c_int;
    pub fn zk_ctx_create_on_stream(device: c_int, cuda_stream: *mut c_void, out: *mut *mut zk_ctx) -> c_int;
    pub fn zk_ctx_destroy(ctx: *mut zk_ctx);
    pub fn zk_last_error(ctx: *const zk_ctx) -> *const c_char;
    pub fn zk_version() -> *const c_char;
    pub fn zk_ctx_synchronize(ctx: *mut zk_ctx) -> c_int;
    pub fn zk_ctx_stream(ctx: *const zk_ctx) -> *mut c_void;
    pub fn zk_ctx_launch_count(ctx: *const zk_ctx) -> u64;
    pub fn zk_ctx_last_stage_ms(ctx: *mut zk_ctx, out_ms: *mut f32) -> c_int;
    pub fn zk_ctx_set_stage_timing(ctx: *mut zk_ctx, on: c_int) -> c_int;
    pub fn zk_interpolate_subgroup(ctx: *mut zk_ctx, y: *const u32, n_points: u32, log_n: u32, coeffs_out: *mut u32) -> c_int;
    pub fn zk_lde_coset(ctx: *mut zk_ctx, coeffs: *const u32, log_n: u32, log_blowup: u32, shift: u32, evals_out: *mut u32) -> c_int;
    pub fn zk_fri_fold(ctx: *mut zk_ctx, cur: *const u32, len: u64, layer_k: u32, beta_raw: u32, next_out: *mut u32) -> c_int;
    pub fn zk_cp_eval_fibsq(ctx: *mut zk_ctx, f_eval: *const u32, log_n: u32, log_blowup: u32, alpha_raw: *const u32, a_first: u32, a_last: u32, cp_out: *mut u32) -> c_int;
    pub fn zk_merkle_build(ctx: *mut zk_ctx, leaves: *const u32, size: u64, heap_out: *mut u8) -> c_int;
    pub fn zk_merkle_root(ctx: *mut zk_ctx, leaves: *const u32, size: u64, root_out: *mut u8) -> c_int;
    pub fn zk_merkle_trace(heap: *const u8, size: u64, index: u64, path_out: *mut u8) -> c_int;
    pub fn zk_transcript_len(log_n: u32) -> usize;
    pub fn zk_prove_fibsq(ctx: *mut zk_ctx, log_n: u32, secret: u32, transcript_out: *mut u8, cap: usize, len_out: *mut usize, state_out: *mut u8) -> c_int;
    pub fn zk_transcript_len_ex(log_n: u32, n_extra_queries: u32) -> usize;
    pub fn zk_prove_fibsq_ex(ctx: *mut zk_ctx, log_n: u32, secret: u32, init_state: *const u8, n_extra_queries: u32, transcript_out: *mut u8, cap: usize, len_out: *mut usize, state_out: *mut u8) -> c_int;
    pub fn zk_verify_fibsq_ex(log_n: u32, a_last: u32, transcript: *const u8, len: usize, state: *const u8, init_state: *const u8, n_extra_queries: u32, rederive: c_int, failed_check: *mut c_int) -> c_int;
    pub fn zk_prove_fibsq_from_trace(ctx: *mut zk_ctx, log_n: u32, trace: *const u32, transcript_out: *mut u8, cap: usize, len_out: *mut usize, state_out: *mut u8) -> c_int;
    pub fn zk_prove_fibsq_batch(ctx: *mut zk_ctx, log_n: u32, n_proofs: u32, traces: *const u32, transcripts_out: *mut u8, states_out: *mut u8) -> c_int;
    pub fn zk_prove_fibsq_batch_secrets(ctx: *mut zk_ctx, log_n: u32, n_proofs: u32, secrets: *const u32, transcripts_out: *mut u8, states_out: *mut u8) -> c_int;
    pub fn zk_prove_stream_submit(ctx: *mut zk_ctx, log_n: u32, n_proofs: u32, secrets: *const u32) -> c_int;
    pub fn zk_prove_stream_collect(ctx: *mut zk_ctx, transcripts_out: *mut u8, states_out: *mut u8) -> c_int;
    pub fn zk_verify_fibsq(log_n: u32, a_last: u32, transcript: *const u8, len: usize, state: *const u8, rederive: c_int, failed_check: *mut c_int) -> c_int;
    pub fn zk_sha256(data: *const u8, len: usize, out: *mut u8);
    pub fn zk_proof_serialized_len(transcript_len: usize) -> usize;
    pub fn zk_proof_serialize(state: *const u8, transcript: *const u8, transcript_len: usize, out: *mut u8, cap: usize, len_out: *mut usize) -> c_int;
    pub fn zk_proof_deserialize(bytes: *const u8, len: usize, state_out: *mut u8, transcript_out: *mut *const u8, transcript_len_out: *mut usize) -> c_int;
    pub fn zk_trace_fibsq(log_n: u32, secret: u32, trace_out: *mut u32) -> c_int;
    pub fn zk_ctx_trim(ctx: *mut zk_ctx, tables_too: c_int) -> c_int;
    pub fn zk_trace_fibsq_batch(log_n: u32, secrets: *const u32, n_proofs: u32, threads: u32, out: *mut u32) -> c_int;
    pub fn zk_ctx_set_air(ctx: *mut zk_ctx, air: *const zk_air) -> c_int;
    pub fn zk_ctx_get_air(ctx: *const zk_ctx, out: *mut zk_air) -> c_int;
    pub fn zk_trace_air(air: *const zk_air, log_n: u32, secret: u32, trace_out: *mut u32) -> c_int;
    pub fn zk_trace_air_batch(air: *const zk_air, log_n: u32, secrets: *const u32, n_proofs: u32, threads: u32, out: *mut u32) -> c_int;
    pub fn zk_verify_air(air: *const zk_air, log_n: u32, a_last: u32, transcript: *const u8, len: usize, state: *const u8, init_state: *const u8, n_extra_queries: u32, rederive: c_int, failed_check: *mut c_int) -> c_int;
    pub fn zk_ctx_set_memory_budget(ctx: *mut zk_ctx, bytes: usize) -> c_int;
    pub fn zk_interpolate_subgroup_dev(ctx: *mut zk_ctx, y_inout_n: *mut u32, n_points: u32, log_n: u32, coeffs_out: *mut u32) -> c_int;
    pub fn zk_lde_coset_dev(ctx: *mut zk_ctx, coeffs: *const u32, log_n: u32, log_blowup: u32, shift: u32, evals_out: *mut u32) -> c_int;
    pub fn zk_fri_fold_dev(ctx: *mut zk_ctx, cur: *const u32, len: u64, layer_k: u32, beta_raw: u32, next_out: *mut u32) -> c_int;
    pub fn zk_cp_eval_fibsq_dev(ctx: *mut zk_ctx, f_eval: *const u32, log_n: u32, log_blowup: u32, alpha_raw: *const u32, a_first: u32, a_last: u32, cp_out: *mut u32) -> c_int;
    pub fn zk_merkle_build_dev(ctx: *mut zk_ctx, leaves: *const u32, size: u64, heap_out: *mut u8) -> c_int;
    pub fn zk_prove_fibsq_resident(ctx: *mut zk_ctx, log_n: u32, n_proofs: u32, traces_dev: *const u32, transcripts_out: *mut u8, states_out: *mut u8) -> c_int;
    pub fn zk_decommit_len(log_n: u32) -> usize;
    pub fn zk_open_queries(ctx: *mut zk_ctx, log_n: u32, proof_index: u32, xs: *const u32, n_queries: u32, out: *mut u8, out_cap: usize) -> c_int;
    pub fn zk_shard_state_init_dev(ctx: *mut zk_ctx, state_dev: *mut c_void, a_first: u32, a_last: u32) -> c_int;
    pub fn zk_shard_lde_coset_dev(ctx: *mut zk_ctx, coeffs: *const u32, log_n: u32, coset: u32, work: *mut u32, out: *mut u32) -> c_int;
    pub fn zk_shard_merkle_blocks_dev(ctx: *mut zk_ctx, vals_local: *const u32, local_len: u64, log_block: u32, heap_local: *mut u8) -> c_int;
    pub fn zk_shard_upper_tree_dev(ctx: *mut zk_ctx, upper_heap: *mut u8, log_count: u32, post_op: c_int, fri_layer: u32, n_layers: u32, state_dev: *mut c_void, transcript_dev: *mut u8) -> c_int;
    pub fn zk_shard_cp_dev(ctx: *mut zk_ctx, f_local: *const u32, halo: *const u32, local_len: u64, log_n: u32, n_ranks: u32, rank: u32, log_block: u32, state_dev: *const c_void, cp_local: *mut u32) -> c_int;
    pub fn zk_shard_fold_dev(ctx: *mut zk_ctx, cur_local: *const u32, local_len: u64, layer_k: u32, log_N0: u32, n_ranks: u32, rank: u32, log_block: u32, state_dev: *const c_void, next_local: *mut u32) -> c_int;
    pub fn zk_shard_cp_blocks_dev(ctx: *mut zk_ctx, f_local: *const u32, halo: *const u32, local_len: u64, log_n: u32, n_ranks: u32, rank: u32, log_block: u32, state_dev: *const c_void, cp_local: *mut u32, heap_local: *mut u8) -> c_int;
    pub fn zk_shard_fold_blocks_dev(ctx: *mut zk_ctx, cur_local: *const u32, local_len: u64, layer_k: u32, log_N0: u32, n_ranks: u32, rank: u32, log_block: u32, state_dev: *const c_void, next_local: *mut u32, heap_local: *mut u8) -> c_int;
    pub fn zk_shard_fri_tail_dev(ctx: *mut zk_ctx, log_n: u32, first_layer: u32, layers_arena: *mut u32, heaps_arena: *mut u8, state_dev: *mut c_void, transcript_dev: *mut u8) -> c_int;
    pub fn zk_shard_decommit_dev(ctx: *mut zk_ctx, log_n: u32, n_ranks: u32, rank: u32, log_block: u32, first_tail: u32, tree_ptrs: *const *const c_void, tail_vals: *const u32, tail_heaps: *const u8, state_dev: *const c_void, transcript_dev: *mut u8) -> c_int;
    pub fn zk_shard_unique_id(id_out: *mut u8) -> c_int;
    pub fn zk_shard_comm_init(ctx: *mut zk_ctx, id: *const u8, n_ranks: u32, rank: u32, out: *mut *mut zk_shard_comm) -> c_int;
    pub fn zk_shard_comm_init_local(ctx: *mut zk_ctx, n_virtual_ranks: u32, out: *mut *mut zk_shard_comm) -> c_int;
    pub fn zk_shard_comm_destroy(comm: *mut zk_shard_comm);
    pub fn zk_shard_comm_set(comm: *mut zk_shard_comm, log_block: u32, log_tail_len: u32) -> c_int;
    pub fn zk_prove_fibsq_sharded(comm: *mut zk_shard_comm, log_n: u32, trace: *const u32, transcript_out: *mut u8, cap: usize, len_out: *mut usize, state_out: *mut u8) -> c_int;
    pub fn zk_shard_last_phase_ms(comm: *const zk_shard_comm, out_ms: *mut f32) -> c_int;
    pub fn zk_shard_last_comm_bytes(comm: *const zk_shard_comm) -> u64;
    pub fn zk_shard_last_first_tail(comm: *const zk_shard_comm) -> u32;
    pub fn zk_field_ops(ctx: *mut zk_ctx, a: *const u32, b: *const u32, n: usize, out: *mut u32) -> c_int;
    pub fn zk_bench_set(ctx: *mut zk_ctx, sha_variant: c_int, leaf_pass_k: u32) -> c_int;
    pub fn zk_bench_set_tile_direct_log(ctx: *mut zk_ctx, log_leaves: u32) -> c_int;
    pub fn zk_bench_kernel(ctx: *mut zk_ctx, which: c_int, log_size: u32, iters: u32, flush_l2: c_int, ms_out: *mut f32) -> c_int;
    // END GENERATED
}
