//! `extern "C"` view of include/zkstark_b200.h.  One declaration per export the shim uses.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct zk_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn zk_ctx_create(device: c_int, out: *mut *mut zk_ctx) -> c_int;
    pub fn zk_ctx_destroy(ctx: *mut zk_ctx);
    pub fn zk_last_error(ctx: *const zk_ctx) -> *const c_char;
    pub fn zk_interpolate_subgroup(ctx: *mut zk_ctx, y: *const u32, n_points: u32, log_n: u32, coeffs_out: *mut u32) -> c_int;
    pub fn zk_lde_coset(ctx: *mut zk_ctx, coeffs: *const u32, log_n: u32, log_blowup: u32, shift: u32, evals_out: *mut u32) -> c_int;
    pub fn zk_fri_fold(ctx: *mut zk_ctx, cur: *const u32, len: u64, layer_k: u32, beta_raw: u32, next_out: *mut u32) -> c_int;
    pub fn zk_cp_eval_fibsq(ctx: *mut zk_ctx, f_eval: *const u32, log_n: u32, log_blowup: u32, alpha_raw: *const u32,
                            a_first: u32, a_last: u32, cp_out: *mut u32) -> c_int;
    pub fn zk_merkle_build(ctx: *mut zk_ctx, leaves: *const u32, size: u64, heap_out: *mut u8) -> c_int;
    pub fn zk_transcript_len(log_n: u32) -> usize;
    pub fn zk_prove_fibsq(ctx: *mut zk_ctx, log_n: u32, secret: u32, transcript_out: *mut u8, cap: usize,
                          len_out: *mut usize, state_out: *mut u8) -> c_int;
    pub fn zk_prove_fibsq_batch(ctx: *mut zk_ctx, log_n: u32, n_proofs: u32, traces: *const u32,
                                transcripts_out: *mut u8, states_out: *mut u8) -> c_int;
    pub fn zk_prove_fibsq_batch_secrets(ctx: *mut zk_ctx, log_n: u32, n_proofs: u32, secrets: *const u32,
                                        transcripts_out: *mut u8, states_out: *mut u8) -> c_int;
    pub fn zk_verify_fibsq(log_n: u32, a_last: u32, transcript: *const u8, len: usize, state: *const u8,
                           rederive_challenges: c_int, failed_check: *mut c_int) -> c_int;
    pub fn zk_decommit_len(log_n: u32) -> usize;
    pub fn zk_open_queries(ctx: *mut zk_ctx, proof_index: u32, xs: *const u32, n_queries: u32, out: *mut u8) -> c_int;
}

/// One context per thread (zk_ctx is not thread-safe), created on first use.
pub fn with_ctx<R>(f: impl FnOnce(*mut zk_ctx) -> R) -> R {
    thread_local! {
        static CTX: *mut zk_ctx = unsafe {
            let mut p: *mut zk_ctx = std::ptr::null_mut();
            let rc = zk_ctx_create(0, &mut p);
            assert_eq!(rc, 0, "zk_ctx_create failed: no CUDA device (there is no CPU fallback)");
            p
        };
    }
    CTX.with(|c| f(*c))
}

/// The reference panics on every error (assert_eq!/unwrap); keep that behaviour at the boundary.
pub fn check(ctx: *mut zk_ctx, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(zk_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("zkstark_b200 error {rc}: {msg}");
    }
}

#[allow(dead_code)]
pub type Opaque = c_void;
