//! Hot-path entry points of the reference's src/polynomial.rs, forwarded to the GPU when the call has the
//! shape prover.rs uses; every other generic Polynomial<T> method stays as in the reference (not repeated here).
use crate::ffi;
use crate::F;

/// polynomial.rs:337 `lagrange(&points)` for points (g^i, y_i) on the 2^log_n subgroup (prover.rs:60-61).
pub fn lagrange_subgroup(ys: &[F], log_n: u32) -> Vec<F> {
    let y: Vec<u32> = ys.iter().map(|v| v.residue()).collect();
    let mut c = vec![0u32; 1usize << log_n];
    ffi::with_ctx(|ctx| unsafe {
        let rc = ffi::zk_interpolate_subgroup(ctx, y.as_ptr(), y.len() as u32, log_n, c.as_mut_ptr());
        ffi::check(ctx, rc);
    });
    c.into_iter().map(F::from).collect()
}

/// `f_domain.iter().map(|&n| f_poly.solve(n))` (prover.rs:69-70): all 8n evaluations in one call.
pub fn solve_on_coset(coeffs: &[F], log_n: u32) -> Vec<F> {
    let c: Vec<u32> = coeffs.iter().map(|v| v.residue()).collect();
    let mut out = vec![0u32; 8usize << log_n];
    ffi::with_ctx(|ctx| unsafe {
        let rc = ffi::zk_lde_coset(ctx, c.as_ptr(), log_n, 3, 5, out.as_mut_ptr());
        ffi::check(ctx, rc);
    });
    out.into_iter().map(F::from).collect()
}

/// `fri(poly, beta)` + evaluation on the squared half domain (polynomial.rs:385, prover.rs:200-211).
pub fn fri_fold_evals(cur: &[F], layer_k: u32, beta_raw: u32) -> Vec<F> {
    let c: Vec<u32> = cur.iter().map(|v| v.residue()).collect();
    let mut out = vec![0u32; c.len() / 2];
    ffi::with_ctx(|ctx| unsafe {
        let rc = ffi::zk_fri_fold(ctx, c.as_ptr(), c.len() as u64, layer_k, beta_raw, out.as_mut_ptr());
        ffi::check(ctx, rc);
    });
    out.into_iter().map(F::from).collect()
}
