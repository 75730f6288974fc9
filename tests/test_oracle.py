"""CPU tests: the oracle (oracle/*.c) against every known-answer value the reference holds
for the hot path (tests/golden/reference_kats.json: copied from /root/reference/src, file:line
inside) and against the derived transcript goldens (tests/golden/transcripts.json)."""
import hashlib

import numpy as np
import pytest

P = 3221225473


def hexs(a):
    return bytes(a).hex()


def test_sha256_both_block_functions(orc):
    L = orc.lib()
    for force in (0, 1):
        L.orc_sha256_force_portable(force)
        for n in [0, 1, 3, 4, 55, 56, 63, 64, 65, 119, 120, 127, 128, 1000]:
            d = bytes((i * 7 + 3) & 255 for i in range(n))
            assert orc.sha256(d) == hashlib.sha256(d).digest(), (force, n)
    L.orc_sha256_force_portable(0)


def test_merkle_kat(orc, kats):  # merkle.rs:112-182
    k = kats["merkle_test"]
    heap = orc.merkle_build(k["leaves"])
    assert [hexs(h) for h in heap] == k["heap"]
    for leaf, sibs in k["traces"].items():
        tr = orc.merkle_trace(heap, int(leaf))
        assert [hexs(t) for t in tr] == [k["heap"][s] for s in sibs]
    assert hexs(orc.root_from_path(1, 0, orc.merkle_trace(heap, 0))) == k["heap"][0]  # merkle.rs:181


def test_generator_kat(orc, kats):  # field.rs:213-226
    L = orc.lib()
    k = kats["generator_test"]
    p = k["modulus"]
    g = L.orc_generator(p)
    assert L.orc_order(g, p) == k["order"]
    seen = {L.orc_pow(g, e, p) for e in range(1, p)}
    assert seen == set(range(1, p))
    assert L.orc_generator(P) == kats["prover"]["primitive_root"]  # prover.rs:44


def test_fri_kat(orc, kats):  # polynomial.rs:402-426
    L = orc.lib()
    k = kats["fri_test"]
    p = k["modulus"]
    asc = np.array(k["p0_descending"][::-1], dtype=np.uint32)
    out = np.zeros(len(asc) // 2, dtype=np.uint32)
    L.orc_fri_coeff(asc.ctypes.data, len(asc), k["beta"], out.ctypes.data, p)
    assert list(out[::-1]) == k["p1_descending"]
    x = k["property_x"]
    nx = p - x
    sol = lambda c, t: L.orc_poly_solve(c.ctypes.data, len(c), t, p)
    inv = lambda a: L.orc_inv(a, p)
    g_xx = (sol(asc, x) + sol(asc, nx)) * inv(2) % p
    h_xx = (sol(asc, x) - sol(asc, nx)) * inv(2 * x % p) % p
    assert (g_xx + k["beta"] * h_xx) % p == sol(out, x * x % p)


def test_lagrange_kat(orc, kats):  # polynomial.rs:428-454
    L = orc.lib()
    k = kats["lagrange_test"]
    p = k["modulus"]

    def lag(pts):
        xs = np.array([a for a, _ in pts], dtype=np.uint32)
        ys = np.array([b for _, b in pts], dtype=np.uint32)
        out = np.zeros(len(pts), dtype=np.uint32)
        n = L.orc_lagrange(xs.ctypes.data, ys.ctypes.data, len(pts), out.ctypes.data, p)
        return list(out[:n])

    p0, p1, p2, p3 = [tuple(t) for t in k["points"]]
    poly = lag([p0, p1, p2, p3])
    c = np.array(poly, dtype=np.uint32)
    ev = lambda t: (t, L.orc_poly_solve(c.ctypes.data, len(c), t, p))
    p4, p5, p6 = ev(4), ev(5), ev(6)
    for pts in ([p0, p3, p5, p6], [p1, p6, p3, p2], [p3, p2, p1, p0], [p6, p5, p4, p3]):
        assert lag(pts) == poly


def test_div_kat(orc, kats):  # polynomial.rs:456-490 (i32 coefficients mapped into F_p)
    L = orc.lib()
    to_f = lambda v: v % P
    for case in kats["div_test"]["cases"]:
        lhs = np.array([to_f(v) for v in case["lhs_desc"][::-1]], dtype=np.uint32)
        rhs = np.array([to_f(v) for v in case["rhs_desc"][::-1]], dtype=np.uint32)
        q = np.zeros(len(lhs), dtype=np.uint32)
        r = np.zeros(len(lhs), dtype=np.uint32)
        import ctypes as C
        ql, rl = C.c_uint32(0), C.c_uint32(0)
        rc = L.orc_poly_div(lhs.ctypes.data, len(lhs), rhs.ctypes.data, len(rhs), q.ctypes.data,
                            C.addressof(ql), r.ctypes.data, C.addressof(rl), P)
        assert rc == 0
        assert list(q[: ql.value][::-1]) == [to_f(v) for v in case["q_desc"]]
        assert list(r[: rl.value][::-1]) == [to_f(v) for v in case["r_desc"]]


def test_prover_inline_kats(orc, kats):  # prover.rs:42,48-53,73-78
    L = orc.lib()
    k = kats["prover"]
    a = orc.trace_fibsq(10, k["secret"])
    assert len(a) == 1023 and int(a[1022]) == k["a_1022"]
    g = L.orc_pow(5, k["exp_g"], P)
    h = L.orc_pow(5, k["exp_h"], P)
    assert L.orc_order(g, P) == k["order_g"] and L.orc_order(h, P) == k["order_h"]
    coeffs = orc.interpolate_subgroup(a, 10)
    for i in (0, 1, 500, 1022):  # prover.rs:64-66
        assert L.orc_poly_solve(coeffs.ctypes.data, 1024, L.orc_pow(g, i, P), P) == int(a[i])
    f_eval = orc.lde_coset(coeffs, 10)
    for idx, v in k["f_eval"].items():
        assert int(f_eval[int(idx)]) == v


@pytest.mark.parametrize("log_n", [4, 5, 6, 8, 10, 12, 16])
def test_fast_route_matches_transcript_goldens(orc, transcripts, log_n):
    g = transcripts[str(log_n)]
    tr, st, info = orc.prove_fibsq(log_n, route=orc.ROUTE_FAST)
    assert len(tr) == g["len"] == orc.lib().orc_transcript_len(log_n)
    assert hashlib.sha256(tr).hexdigest() == g["sha256"]
    assert st.hex() == g["state"]
    assert info.a_last == g["a_last"] and info.free_term == g["free"] and info.query_raw == g["query_raw"]
    assert hexs(info.f_root) == g["f_root"] and hexs(info.cp_root) == g["cp_root"]
    if "alpha_raw" in g:
        assert list(info.alpha_raw) == g["alpha_raw"]
        assert list(info.beta_raw)[:log_n] == g["beta_raw"]
        assert [hexs(info.fri_roots[i]) for i in range(log_n)] == g["fri_roots"]
    assert orc.verify_fibsq(log_n, info.a_last, tr, True, st) == 0


@pytest.mark.parametrize("log_n", [4, 5, 6, 7, 8])
def test_literal_route_equals_fast_route(orc, log_n):
    """The reference's own algorithms (Lagrange, long division, per-point solve, coefficient FRI)
    give the same bytes as the NTT route; includes the c0/c1/c2 and degree asserts of prover.rs."""
    a = orc.prove_fibsq(log_n, route=orc.ROUTE_LITERAL, threads=1)
    b = orc.prove_fibsq(log_n, route=orc.ROUTE_FAST)
    assert a[0] == b[0] and a[1] == b[1]


def test_literal_route_canonical_config(orc, transcripts):
    """log_n = 10, the reference's only configuration: runs prover.rs:157-159's KATs inside."""
    tr, st, info = orc.prove_fibsq(10, route=orc.ROUTE_LITERAL, threads=1)
    g = transcripts["10"]
    assert hashlib.sha256(tr).hexdigest() == g["sha256"] and st.hex() == g["state"]
    assert len(tr) + 48 == 7884  # Proof::size(), proof.rs:151-154


def test_verifier_rejects_tampering(orc):
    tr, st, info = orc.prove_fibsq(6)
    assert orc.verify_fibsq(6, info.a_last, tr, True, st) == 0
    for pos in (0, 40, 100, len(tr) // 2, len(tr) - 1):
        bad = bytearray(tr)
        bad[pos] ^= 1
        assert orc.verify_fibsq(6, info.a_last, bytes(bad), True, None) != 0
    assert orc.verify_fibsq(6, (info.a_last + 1) % P, tr, True, None) != 0


def test_stage_functions_consistent(orc):
    """fold(evals) == evaluations of the coefficient fold (polynomial.rs:419-425, as arrays)."""
    L = orc.lib()
    log_n = 6
    rng = np.random.default_rng(1)
    c = rng.integers(0, P, size=1 << log_n, dtype=np.uint64).astype(np.uint32)
    ev = orc.lde_coset(c, log_n)
    beta = 0xF1234567
    folded = orc.fri_fold(ev, 0, beta)
    c2 = np.zeros(len(c) // 2, dtype=np.uint32)
    L.orc_fri_coeff(c.ctypes.data, len(c), beta, c2.ctypes.data, P)
    h = L.orc_pow(5, (P - 1) >> (log_n + 3), P)
    for i in (0, 1, 17, len(folded) - 1):
        x = 5 * L.orc_pow(h, i, P) % P
        assert int(folded[i]) == L.orc_poly_solve(c2.ctypes.data, len(c2), x * x % P, P)
