//! Safe wrappers over the C ABI (ffi.rs) in the vocabulary of the reference crate.  Values cross as canonical
//! `u32` residues (`F::residue()` / `F::from(u32)`), digests as `[u8; 32]`; every error becomes a `panic!`, which
//! is what the reference does on every failure (assert_eq! / unwrap).  There is no CPU fallback: without a CUDA
//! device `zk_ctx_create` fails and the first call panics.
#![allow(dead_code)]
use crate::ffi;
use crate::merkle::Hash;
use crate::F;
use std::any::{Any, TypeId};
use std::os::raw::c_int;

/// prover.rs:56-57: the evaluation domain is 8 times the trace domain.
pub const LOG_BLOWUP: u32 = 3;
/// prover.rs:44-45,69: `F::generator()` is 5, also the coset shift.
pub const SHIFT: u32 = 5;
/// main.rs:13
pub const MODULUS: u32 = 3221225473;

/// One context per thread (zk_ctx is not thread-safe), created on first use, destroyed with the thread.
struct Ctx(*mut ffi::zk_ctx);
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { ffi::zk_ctx_destroy(self.0) }
    }
}
pub fn with_ctx<R>(f: impl FnOnce(*mut ffi::zk_ctx) -> R) -> R {
    thread_local! {
        static CTX: Ctx = unsafe {
            let mut p: *mut ffi::zk_ctx = std::ptr::null_mut();
            let rc = ffi::zk_ctx_create(0, &mut p);
            assert_eq!(rc, 0, "zk_ctx_create failed: no CUDA device (there is no CPU fallback)");
            Ctx(p)
        };
    }
    CTX.with(|c| f(c.0))
}

/// The reference panics on every error; keep that behaviour at the boundary.
pub fn check(ctx: *mut ffi::zk_ctx, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::zk_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("zkstark_b200 error {rc}: {msg}");
    }
}

pub fn log2_exact(n: usize) -> u32 {
    assert!(n.is_power_of_two(), "size must be a power of two");
    n.trailing_zeros()
}

fn pow_mod(mut b: u64, mut e: u64) -> u32 {
    let m = MODULUS as u64;
    let mut r = 1u64;
    b %= m;
    while e != 0 {
        if e & 1 == 1 {
            r = r * b % m;
        }
        b = b * b % m;
        e >>= 1;
    }
    r as u32
}

/// generator of the order-2^log subgroup: 5^((p-1)/2^log)   (prover.rs:48-49)
pub fn subgroup_generator(log: u32) -> u32 {
    pow_mod(SHIFT as u64, ((MODULUS as u64) - 1) >> log)
}

/// polynomial.rs:337 for points (g^i, y_i): ascending coefficients (`Polynomial.0`), 2^log_n of them.
/// `ys.len()` is 2^log_n - 1 (prover.rs:60: zip truncates to the 1023 trace values) or 2^log_n.
pub fn interpolate_subgroup(ys: &[u32], log_n: u32) -> Vec<u32> {
    let mut c = vec![0u32; 1usize << log_n];
    with_ctx(|ctx| unsafe {
        check(ctx, ffi::zk_interpolate_subgroup(ctx, ys.as_ptr(), ys.len() as u32, log_n, c.as_mut_ptr()));
    });
    c
}

/// `f_domain.iter().map(|&n| poly.solve(n))` (prover.rs:69-70): all 8n evaluations on 5 h^i, natural order.
pub fn lde_coset(coeffs: &[u32], log_n: u32) -> Vec<u32> {
    assert_eq!(coeffs.len(), 1usize << log_n);
    let mut out = vec![0u32; 1usize << (log_n + LOG_BLOWUP)];
    with_ctx(|ctx| unsafe {
        check(ctx, ffi::zk_lde_coset(ctx, coeffs.as_ptr(), log_n, LOG_BLOWUP, SHIFT, out.as_mut_ptr()));
    });
    out
}

/// `fri(poly, beta)` followed by `solve` over the squared half domain (polynomial.rs:385, prover.rs:200-211), in
/// evaluation form: `cur` are the values of layer `layer_k`; beta is the raw channel word (reduced inside).
pub fn fri_fold(cur: &[u32], layer_k: u32, beta_raw: u32) -> Vec<u32> {
    let mut out = vec![0u32; cur.len() / 2];
    with_ctx(|ctx| unsafe {
        check(ctx, ffi::zk_fri_fold(ctx, cur.as_ptr(), cur.len() as u64, layer_k, beta_raw, out.as_mut_ptr()));
    });
    out
}

/// prover.rs:101-173: the three constraint quotients mixed with alpha_0..2, evaluated on the LDE domain.
pub fn cp_eval_fibsq(f_eval: &[u32], log_n: u32, alpha_raw: [u32; 3], a_first: u32, a_last: u32) -> Vec<u32> {
    assert_eq!(f_eval.len(), 1usize << (log_n + LOG_BLOWUP));
    let mut out = vec![0u32; f_eval.len()];
    with_ctx(|ctx| unsafe {
        check(
            ctx,
            ffi::zk_cp_eval_fibsq(ctx, f_eval.as_ptr(), log_n, LOG_BLOWUP, alpha_raw.as_ptr(), a_first, a_last, out.as_mut_ptr()),
        );
    });
    out
}

/// merkle.rs:14-51: the whole heap, root first.
pub fn merkle_build(leaves: &[u32]) -> Vec<Hash> {
    let size = leaves.len();
    assert!(size.is_power_of_two()); // merkle.rs:17-21
    let mut heap = vec![Hash::default(); 2 * size - 1];
    with_ctx(|ctx| unsafe {
        check(ctx, ffi::zk_merkle_build(ctx, leaves.as_ptr(), size as u64, heap.as_mut_ptr() as *mut u8));
    });
    heap
}

/// prover.rs:9-293 in one call: (`Proof.state`, `Proof.data`).
pub fn prove_fibsq(log_n: u32, secret: u32) -> (Hash, Vec<u8>) {
    let cap = unsafe { ffi::zk_transcript_len(log_n) };
    let mut data = vec![0u8; cap];
    let mut state: Hash = [0u8; 32];
    let mut len = 0usize;
    with_ctx(|ctx| unsafe {
        check(ctx, ffi::zk_prove_fibsq(ctx, log_n, secret, data.as_mut_ptr(), cap, &mut len, state.as_mut_ptr()));
    });
    data.truncate(len);
    (state, data)
}

/// n independent proofs, secrets in, (state, data) out.
pub fn prove_fibsq_batch(log_n: u32, secrets: &[u32]) -> Vec<(Hash, Vec<u8>)> {
    let tl = unsafe { ffi::zk_transcript_len(log_n) };
    let mut data = vec![0u8; tl * secrets.len()];
    let mut states = vec![0u8; 32 * secrets.len()];
    with_ctx(|ctx| unsafe {
        check(
            ctx,
            ffi::zk_prove_fibsq_batch_secrets(ctx, log_n, secrets.len() as u32, secrets.as_ptr(), data.as_mut_ptr(), states.as_mut_ptr()),
        );
    });
    (0..secrets.len())
        .map(|j| {
            let mut st: Hash = [0u8; 32];
            st.copy_from_slice(&states[32 * j..32 * j + 32]);
            (st, data[tl * j..tl * (j + 1)].to_vec())
        })
        .collect()
}

/// The program as data (include/zkstark_b200.h `zk_air`): prove `a[i+2] = c20 a[i+1]^2 + c02 a[i]^2 + c11 a[i] a[i+1] +
/// c10 a[i+1] + c01 a[i] + c00` from `a[0] = a0, a[1] = secret` instead of the reference's hard-coded FibonacciSq
/// (prover.rs:32-39, :134-138).  The thread's context proves `air` for the duration of the call only.
pub fn prove_air(air: &ffi::zk_air, log_n: u32, secret: u32) -> (Hash, Vec<u8>) {
    with_ctx(|ctx| unsafe { check(ctx, ffi::zk_ctx_set_air(ctx, air)) });
    struct Reset;
    impl Drop for Reset {
        fn drop(&mut self) {
            with_ctx(|ctx| unsafe {
                ffi::zk_ctx_set_air(ctx, std::ptr::null());
            });
        }
    }
    let _reset = Reset;
    prove_fibsq(log_n, secret)
}

/// The trace of `air` (host): 2^log_n - 1 values.
pub fn trace_air(air: &ffi::zk_air, log_n: u32, secret: u32) -> Vec<u32> {
    let mut out = vec![0u32; (1usize << log_n) - 1];
    let rc = unsafe { ffi::zk_trace_air(air, log_n, secret, out.as_mut_ptr()) };
    assert_eq!(rc, 0, "zk_trace_air: bad log_n or zk_air");
    out
}

/// proof.rs:15-149 for a proof of `air`, the channel replayed as well; panics like `Proof::verify` with the line of the
/// failing assert.
pub fn verify_air(air: &ffi::zk_air, log_n: u32, a_last: u32, state: &Hash, data: &[u8]) {
    let mut line: c_int = 0;
    let rc = unsafe {
        ffi::zk_verify_air(air, log_n, a_last, data.as_ptr(), data.len(), state.as_ptr(), std::ptr::null(), 0, 1, &mut line)
    };
    assert_eq!(rc, 0, "proof rejected: the assert at proof.rs:{line} fails");
}

/// Wire format of `Proof { state, data }` as bincode 1.x writes the struct: state || u64-LE length || data.
pub fn proof_to_bytes(state: &Hash, data: &[u8]) -> Vec<u8> {
    let cap = unsafe { ffi::zk_proof_serialized_len(data.len()) };
    let mut out = vec![0u8; cap];
    let mut len = 0usize;
    let rc = unsafe { ffi::zk_proof_serialize(state.as_ptr(), data.as_ptr(), data.len(), out.as_mut_ptr(), cap, &mut len) };
    assert_eq!(rc, 0);
    out.truncate(len);
    out
}

/// Hook for `lagrange::<T>` (polynomial.rs:337).  When `T` is the crate's field `F` the interpolation runs on the
/// GPU and the ascending coefficients are returned; for any other `T` (the unit tests over GF(7) / GF(101)) the
/// reference body runs as before (`None`).  For `T == F` the abscissae must be g^i, g = 5^((p-1)/n), n the next
/// power of two >= points.len() -- exactly what prover.rs:56-61 builds; anything else panics (no CPU fallback).
pub fn lagrange_hook<T: Copy + 'static>(points: &Vec<(T, T)>) -> Option<Vec<T>> {
    if TypeId::of::<T>() != TypeId::of::<F>() {
        return None;
    }
    let as_f = |t: &T| -> F { *(t as &dyn Any).downcast_ref::<F>().unwrap() };
    let n = points.len().next_power_of_two();
    assert!(points.len() == n || points.len() == n - 1, "lagrange on the GPU needs n or n - 1 points of the subgroup");
    let log_n = log2_exact(n);
    let g = subgroup_generator(log_n) as u64;
    let mut x = 1u64;
    let mut ys = Vec::with_capacity(points.len());
    for (px, py) in points.iter() {
        assert_eq!(as_f(px).residue() as u64, x, "lagrange on the GPU needs abscissae g^i (prover.rs:56-61)");
        ys.push(as_f(py).residue());
        x = x * g % (MODULUS as u64);
    }
    let coeffs = interpolate_subgroup(&ys, log_n);
    // Polynomial.0 holds the ascending coefficients WITHOUT trailing zeros (polynomial.rs:15-28 `reduce`)
    let mut out: Vec<F> = coeffs.into_iter().map(F::from).collect();
    while out.last().map_or(false, |c| c.residue() == 0) {
        out.pop();
    }
    Some(out.into_iter().map(|f| *(&f as &dyn Any).downcast_ref::<T>().unwrap()).collect())
}
