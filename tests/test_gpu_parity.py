"""GPU parity tests (-m gpu): every C-ABI entry point against the oracle on the same inputs,
bit-exact, plus the reference's known-answer values through the CUDA path."""
import hashlib

import numpy as np
import pytest

import zkstark_b200 as zk

pytestmark = pytest.mark.gpu
P = 3221225473


def rnd(n, seed):
    return np.random.default_rng(seed).integers(0, P, size=n, dtype=np.uint64).astype(np.uint32)


def edge_values(n, seed):
    v = rnd(n, seed)
    edges = [0, 1, 2, P - 1, P - 2, 1 << 31, (1 << 31) - 1, (1 << 31) + 1, P // 2, 5]
    v[: min(n, len(edges))] = edges[: min(n, len(edges))]
    return v


# ---------------------------------------------------------------- K5 Merkle
def test_merkle_reference_kat(ctx, kats):  # merkle.rs:112-182 through the CUDA kernels
    k = kats["merkle_test"]
    m = zk.Merkle.new(4, np.array(k["leaves"], dtype=np.uint32), ctx=ctx)
    assert [m[i].hex() for i in range(7)] == k["heap"]
    for leaf, sibs in k["traces"].items():
        assert [t.hex() for t in m.trace(int(leaf))] == [k["heap"][s] for s in sibs]


@pytest.mark.parametrize("log_size", list(range(0, 20)))   # every tile / warp-pair / leaf-pass regime
def test_merkle_random_vs_oracle(ctx, orc, log_size):
    leaves = edge_values(1 << log_size, 100 + log_size)
    got = ctx.merkle_build(leaves)
    want = orc.merkle_build(leaves, threads=orc.default_threads())
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert ctx.merkle_root(leaves) == want[0].tobytes()


def test_merkle_rejects_bad_size(ctx):
    with pytest.raises(zk.ZkError):
        ctx.merkle_build(np.zeros(3, dtype=np.uint32))  # merkle.rs:17-21 assert
    with pytest.raises(zk.ZkError):
        zk.Merkle.new(8, np.zeros(4, dtype=np.uint32), ctx=ctx)


def test_merkle_full_size_properties(ctx, orc):
    """2^23 leaves (the 2^20-trace LDE): root equals the oracle's; every auth path of a few leaves
    recomputes to the root (merkle.rs:82-110)."""
    n = 1 << 23
    leaves = rnd(n, 7)
    heap = ctx.merkle_build(leaves)
    want_root = orc.merkle_build(leaves, threads=orc.default_threads())[0].tobytes()
    assert heap[0].tobytes() == want_root
    for i in (0, 1, n // 3, n - 1):
        path = zk.merkle_trace_host(heap, i)
        assert len(path) == 23
        assert orc.root_from_path(int(leaves[i]), i, path).tobytes() == want_root


# ---------------------------------------------------------------- K2 interpolation + LDE
@pytest.mark.parametrize("log_n", [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22])
def test_interpolate_phantom_vs_oracle(ctx, orc, log_n):
    y = edge_values((1 << log_n) - 1, 200 + log_n)
    got = ctx.interpolate_subgroup(y, log_n)
    want = orc.interpolate_subgroup(y, log_n, threads=orc.default_threads())
    assert np.array_equal(got, want)
    assert got[-1] == 0  # degree <= n-2 (prover.rs:154 implies it)


@pytest.mark.parametrize("log_n", [3, 6, 10, 11, 15])
def test_interpolate_full_vs_oracle(ctx, orc, log_n):
    y = edge_values(1 << log_n, 300 + log_n)
    assert np.array_equal(ctx.interpolate_subgroup(y, log_n), orc.interpolate_subgroup(y, log_n))


@pytest.mark.parametrize("log_n", [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21])
def test_lde_vs_oracle(ctx, orc, log_n):
    c = edge_values(1 << log_n, 400 + log_n)
    got = ctx.lde_coset(c, log_n)
    want = orc.lde_coset(c, log_n, threads=orc.default_threads())
    assert np.array_equal(got, want)


def test_lde_other_shift(ctx, orc):
    c = rnd(1 << 9, 11)
    assert np.array_equal(ctx.lde_coset(c, 9, shift=7), orc.lde_coset(c, 9, shift=7))
    assert np.array_equal(ctx.lde_coset(c, 9, shift=1), orc.lde_coset(c, 9, shift=1))


def test_lde_reference_kat(ctx, kats):  # prover.rs:42,73-78 through the CUDA kernels
    k = kats["prover"]
    a = zk.trace_fibsq(10, k["secret"])
    f = ctx.lde_coset(ctx.interpolate_subgroup(a, 10), 10)
    for idx, v in k["f_eval"].items():
        assert int(f[int(idx)]) == v


def test_lde_linearity_full_size(ctx):
    """size-independent property at 2^20: LDE(a + b) = LDE(a) + LDE(b) mod p."""
    n = 1 << 20
    a, b = rnd(n, 1), rnd(n, 2)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    la, lb, ls = ctx.lde_coset(a, 20), ctx.lde_coset(b, 20), ctx.lde_coset(s, 20)
    assert np.array_equal(((la.astype(np.uint64) + lb) % P).astype(np.uint32), ls)


def test_interpolate_then_lde_hits_trace(ctx):
    """f(g^i) = a[i] (prover.rs:64-66): LDE with shift 1 passes through the subgroup at stride 8."""
    log_n = 12
    a = zk.trace_fibsq(log_n, 99)
    c = ctx.interpolate_subgroup(a, log_n)
    ev = ctx.lde_coset(c, log_n, shift=1)
    assert np.array_equal(ev[::8][:-1], a)


# ---------------------------------------------------------------- K3 composition polynomial
@pytest.mark.parametrize("log_n", [3, 4, 8, 10, 13, 16])
def test_cp_vs_oracle(ctx, orc, log_n):
    f = edge_values(8 << log_n, 500 + log_n)
    alpha = [361545003, 3235878091, 0xFFFFFFFF]  # raw challenges >= p occur
    got = ctx.cp_eval_fibsq(f, log_n, alpha, 1, 2338775057)
    want = orc.cp_eval_fibsq(f, log_n, alpha, 1, 2338775057, threads=orc.default_threads())
    assert np.array_equal(got, want)


def test_cp_reference_values(ctx, transcripts):  # derived goldens for the canonical run
    g = transcripts["10"]
    a = zk.trace_fibsq(10, 3141592)
    f = ctx.lde_coset(ctx.interpolate_subgroup(a, 10), 10)
    cp = ctx.cp_eval_fibsq(f, 10, g["alpha_raw"], 1, int(a[1022]))
    for idx, v in g["cp_eval"].items():
        assert int(cp[int(idx)]) == v


# ---------------------------------------------------------------- K4 FRI fold
@pytest.mark.parametrize("log_len,layer_k", [(1, 0), (2, 3), (3, 9), (4, 0), (10, 2), (13, 0), (13, 7), (20, 1)])
def test_fold_vs_oracle(ctx, orc, log_len, layer_k):
    cur = edge_values(1 << log_len, 600 + log_len)
    for beta in (3, 4195595581):
        got = ctx.fri_fold(cur, layer_k, beta)
        want = orc.fri_fold(cur, layer_k, beta, threads=orc.default_threads())
        assert np.array_equal(got, want)


def test_fold_is_coefficient_fold(ctx, orc):
    """fri_test's identity (polynomial.rs:419-425) on arrays: fold of evaluations ==
    evaluations of fri(poly, beta) on the squared half domain."""
    log_n, beta = 7, 123456789
    c = rnd(1 << log_n, 5)
    ev = ctx.lde_coset(c, log_n)
    c2 = np.zeros(len(c) // 2, dtype=np.uint32)
    orc.lib().orc_fri_coeff(c.ctypes.data, len(c), beta, c2.ctypes.data, P)
    folded = ctx.fri_fold(ev, 0, beta)
    # evaluations of c2 (degree < n/2) on {(5 h^i)^2}: an LDE of c2 with shift 25 over the order-N/2 group
    want = orc.lde_coset(np.concatenate([c2]), log_n - 1, shift=25)
    assert np.array_equal(folded, want)


# ---------------------------------------------------------------- prove(): bit-exact transcripts
@pytest.mark.parametrize("log_n", [4, 5, 6, 8, 10, 12, 16])
def test_prove_matches_goldens(ctx, transcripts, log_n):
    g = transcripts[str(log_n)]
    tr, st = ctx.prove_fibsq(log_n, 3141592)
    assert len(tr) == g["len"]
    assert tr[:32].hex() == g["f_root"]
    assert tr[44:76].hex() == g["cp_root"]
    assert hashlib.sha256(tr).hexdigest() == g["sha256"]
    assert st.hex() == g["state"]


@pytest.mark.parametrize("log_n,secret", [(3, 3141592), (7, 1), (9, 0), (11, 42), (13, 3221225473 + 7), (14, 2), (18, 3141592)])
def test_prove_vs_oracle(ctx, orc, log_n, secret):
    tr, st = ctx.prove_fibsq(log_n, secret)
    otr, ost, info = orc.prove_fibsq(log_n, secret)
    assert tr == otr and st == ost
    assert orc.verify_fibsq(log_n, info.a_last, tr, True, st) == 0


def test_generate_proof_mirror(ctx, transcripts):
    proof = zk.generate_proof(zk.Channel.new(), ctx=ctx)  # the reference's only config
    assert proof.size() == 7884  # proof.rs:151-154
    assert proof.state.hex() == transcripts["10"]["state"]
    ch = zk.Channel.new()
    ch.state = b"\x01" * 32
    with pytest.raises(zk.ZkError):
        zk.generate_proof(ch, ctx=ctx)


def test_prove_full_size(ctx, orc):
    """BASELINE config 3 (2^20 trace): bit-exact with the oracle, and accepted by the restated verifier."""
    tr, st = ctx.prove_fibsq(20, 3141592)
    otr, ost, info = orc.prove_fibsq(20, 3141592)
    assert hashlib.sha256(tr).hexdigest() == hashlib.sha256(otr).hexdigest()
    assert st == ost
    assert orc.verify_fibsq(20, info.a_last, tr, True, st) == 0


@pytest.mark.parametrize("log_n", [5, 12, 16])
def test_prove_cp_forms_and_plain_hashing_agree(ctx, orc, log_n):
    """The composition polynomial from the per-size constant tables and its per-point form, and the tile kernel with
    and without two-warp node hashing, all give the oracle's bytes."""
    otr, ost, _ = orc.prove_fibsq(log_n, 77)
    try:
        for cp_form, coop in [(-101, 128), (-100, 128), (-101, 0), (-101, 32)]:
            ctx.bench_set(cp_form, 0)
            ctx.bench_set(-1, (coop + 1) << 8)
            tr, st = ctx.prove_fibsq(log_n, 77)
            assert tr == otr and st == ost, (cp_form, coop)
    finally:
        ctx.bench_set(-101, 0)
        ctx.bench_set(-1, (128 + 1) << 8)


@pytest.mark.parametrize("log_n", [18, 22, 24])
def test_prove_large_sizes_match_golden_digests(ctx, log_n):
    """BASELINE configs up to the 2^24 trace (2^27-point LDE, 28 GB of work space): transcript and final channel state
    against tests/golden/large_digests.json (made by tests/golden/make_large_digests.py from the oracle; 2^24 takes the
    oracle 35 s, too slow to redo per run)."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "large_digests.json")))[str(log_n)]
    tr, st = ctx.prove_fibsq(log_n, 3141592)
    assert len(tr) == g["len"] == zk.transcript_len(log_n)
    assert hashlib.sha256(tr).hexdigest() == g["sha256"]
    assert st.hex() == g["state"]
    ok, line = zk.verify_fibsq(log_n, g["a_last"], tr, st)
    assert ok, line


def test_prove_batch(ctx, orc):
    """BASELINE config 4 in miniature: independent proofs, secret_j = 3141592 + j."""
    log_n, nb = 10, 5
    traces = np.stack([zk.trace_fibsq(log_n, 3141592 + j) for j in range(nb)])
    trs, sts = ctx.prove_fibsq_batch(log_n, traces)
    for j in range(nb):
        otr, ost, _ = orc.prove_fibsq(log_n, 3141592 + j)
        assert trs[j].tobytes() == otr and sts[j].tobytes() == ost


@pytest.mark.parametrize("log_n", [3, 6, 14])
def test_prove_batch_device_channel_tail(ctx, orc, log_n):
    """Batches of >= 8 proofs finish the channel over the decommitment records on the device (chan_tail_kernel)."""
    nb = 9
    traces = np.stack([zk.trace_fibsq(log_n, 500 + j) for j in range(nb)])
    trs, sts = ctx.prove_fibsq_batch(log_n, traces)
    for j in range(nb):
        otr, ost, _ = orc.prove_fibsq(log_n, 500 + j)
        assert trs[j].tobytes() == otr and sts[j].tobytes() == ost


def test_prove_is_repeatable_under_concurrency(ctx, orc):
    """Races in the cooperative hashing, the last-arriver tile reduction, the two-lane batches or the programmatic
    launches would show up as run-to-run differences: 40 repetitions, one answer."""
    log_n, nb = 11, 12
    traces = np.stack([zk.trace_fibsq(log_n, 900 + j) for j in range(nb)])
    want = [orc.prove_fibsq(log_n, 900 + j)[:2] for j in range(nb)]
    for rep in range(20):
        trs, sts = ctx.prove_fibsq_batch(log_n, traces)
        for j in range(nb):
            assert trs[j].tobytes() == want[j][0] and sts[j].tobytes() == want[j][1], (rep, j)
    otr, ost, _ = orc.prove_fibsq(15, 4242)
    for rep in range(20):
        assert ctx.prove_fibsq(15, 4242) == (otr, ost), rep


def test_prove_batch_sub_batching(ctx, orc):
    log_n, nb = 8, 7
    ctx.set_memory_budget(3 * 1672 * (1 << log_n) + 4096)  # forces sub-batches of <= 3 proofs
    try:
        traces = np.stack([zk.trace_fibsq(log_n, 10 + j) for j in range(nb)])
        trs, sts = ctx.prove_fibsq_batch(log_n, traces)
    finally:
        ctx.set_memory_budget(0)
    for j in range(nb):
        otr, ost, _ = orc.prove_fibsq(log_n, 10 + j)
        assert trs[j].tobytes() == otr and sts[j].tobytes() == ost


def test_prove_rejects_bad_args(ctx):
    with pytest.raises(zk.ZkError):
        ctx.prove_fibsq(2)
    with pytest.raises(zk.ZkError):
        ctx.prove_fibsq(28)
    with pytest.raises(zk.ZkError):
        ctx.lde_coset(np.zeros(16, dtype=np.uint32), 4, log_blowup=2)
    with pytest.raises(zk.ZkError):
        ctx.fri_fold(np.zeros(12, dtype=np.uint32), 0, 1)


def test_launch_counter_counts_own_kernels(ctx):
    before = ctx.launch_count()
    ctx.prove_fibsq(8)
    assert ctx.launch_count() - before >= 12


# ---------------------------------------------------------------- 8(f) rows: device trace generation, verifier
def test_prove_batch_from_secrets(ctx, orc):
    """zk_prove_fibsq_batch_secrets: traces generated inside the call (host SIMD lanes x threads, one sub-batch ahead of the
    proving; 70 proofs: cut into four sub-batches), prover.rs:32-39."""
    log_n = 9
    secrets = np.array([3141592, 0, 1, P - 1, P + 5, 0xFFFFFFFF] + list(range(100, 164)), dtype=np.uint32)
    trs, sts = ctx.prove_fibsq_batch_secrets(log_n, secrets)
    for j in (0, 1, 2, 3, 4, 5, 17, 18, 35, 36, 53, 69):
        otr, ost, info = orc.prove_fibsq(log_n, int(secrets[j]))
        assert trs[j].tobytes() == otr and sts[j].tobytes() == ost
        assert zk.verify_fibsq(log_n, info.a_last, otr, ost) == (True, 0)


@pytest.mark.parametrize("log_n", [5, 10, 14])
def test_open_more_queries(ctx, orc, log_n):
    """openings of further positions of the resident proof.  The proof's own position reproduces the
    transcript's records byte for byte; any other position, spliced in place of the original query, passes the
    reference's own checks (proof.rs:63-148: every Merkle path, the CP constraint, every fold step)."""
    import struct
    tr, st = ctx.prove_fibsq(log_n, 4242)
    otr, ost, info = orc.prove_fibsq(log_n, 4242)
    assert tr == otr
    N = 8 << log_n
    off_q = 80 + 36 * log_n
    own = struct.unpack_from("<I", tr, off_q)[0] % (N - 16)
    rnd = np.random.default_rng(log_n)
    xs = [own, 0, 1, N - 17, N // 2 - 8, N // 2] + [int(v) for v in rnd.integers(0, N - 16, 10)]
    recs = ctx.open_queries(log_n, xs)
    assert recs[0] == tr[off_q + 4:]
    for x, rec in zip(xs, recs):
        spliced = tr[:off_q] + struct.pack("<I", x) + rec
        assert len(spliced) == len(tr)
        ok, line = zk.verify_fibsq(log_n, info.a_last, spliced, st, rederive=False)
        assert ok, (x, line)
    bad = bytearray(recs[3]); bad[40] ^= 1
    ok, _ = zk.verify_fibsq(log_n, info.a_last, tr[:off_q] + struct.pack("<I", xs[3]) + bytes(bad), st, rederive=False)
    assert not ok
    with pytest.raises(zk.ZkError):
        ctx.open_queries(log_n, [N - 16])


def test_gpu_proof_passes_product_verifier(ctx, kats):
    proof = zk.generate_proof(zk.Channel.new(), ctx=ctx)
    proof.verify()  # proof.rs:15-149 with a[1022] = 2338775057 (proof.rs:70)
    tampered = zk.Proof(proof.state, proof.data[:100] + bytes([proof.data[100] ^ 1]) + proof.data[101:])
    with pytest.raises(zk.ZkError):
        tampered.verify()


def test_prove_stream_pipeline(ctx, orc):
    """zk_prove_stream_submit / _collect: secrets in, proofs out, two batches in flight, FIFO order."""
    k = 12
    sa = np.array([3141592 + j for j in range(5)], dtype=np.uint32)
    sb = np.array([17 + j for j in range(3)], dtype=np.uint32)
    with pytest.raises(zk.ZkError):
        ctx.prove_stream_collect(k, 1)                  # nothing submitted
    ctx.prove_stream_submit(k, sa)
    ctx.prove_stream_submit(k, sb)
    with pytest.raises(zk.ZkError):
        ctx.prove_stream_submit(k, sb)                  # two batches already in flight
    with pytest.raises(zk.ZkError):
        ctx.trim()                                      # batches in flight
    tra, sta = ctx.prove_stream_collect(k, len(sa))
    ctx.prove_stream_submit(k, sa[:2])                  # refill while one is pending
    trb, stb = ctx.prove_stream_collect(k, len(sb))
    trc, stc = ctx.prove_stream_collect(k, 2)
    for secrets, tr, st in ((sa, tra, sta), (sb, trb, stb), (sa[:2], trc, stc)):
        for j, s in enumerate(secrets):
            otr, ost, _ = orc.prove_fibsq(k, int(s))
            assert tr[j].tobytes() == otr and st[j].tobytes() == ost


def test_trim_then_prove_again(ctx, orc):
    """zk_ctx_trim releases the work space and the per-size tables; the next call rebuilds them."""
    a = ctx.prove_fibsq(11, 77)
    ctx.trim()
    b = ctx.prove_fibsq(11, 77)
    ctx.trim(tables_too=True)
    c = ctx.prove_fibsq(11, 77)
    otr, ost, _ = orc.prove_fibsq(11, 77)
    assert a == b == c == (otr, ost)
    with pytest.raises(zk.ZkError):
        ctx.open_queries(12, [5])                       # log_n differs from the resident proof's
    assert len(ctx.open_queries(11, [5, 9])) == 2


def test_lde_other_coset_shift(ctx, orc):
    """zk_lde_coset with shift != 5 (tables of the shift are cached per context; no stream synchronisation)."""
    k = 9
    rng = np.random.default_rng(5)
    coeffs = rng.integers(0, zk.P, size=1 << k, dtype=np.uint64).astype(np.uint32)
    for shift in (7, 5, 7, 3221225472):
        got = ctx.lde_coset(coeffs, k, shift=shift)
        h = pow(5, (zk.P - 1) >> (k + 3), zk.P)
        idx = [0, 1, 2, 77, (8 << k) - 1]
        for i in idx:
            x = shift % zk.P * pow(h, i, zk.P) % zk.P
            want = 0
            for c in reversed(coeffs.tolist()):
                want = (want * x + c) % zk.P
            assert int(got[i]) == want


def test_field_ops_direct(ctx, orc):
    """K1 directly: device fadd / fsub / mont_mul / fneg / mont_inv / mont_pow (csrc/field.cuh) against oracle/field.c
    (field.rs:96-211) on the full edge grid and on 2^20 random pairs."""
    P = zk.P
    L = orc.lib()
    edge = [0, 1, 2, P - 2, P - 1, P, P + 1, 1 << 31, (1 << 31) + 1, 0xFFFFFFFF]   # raw words: >= p are reduced first (field.rs:20-24)
    ga = np.array([a for a in edge for _ in edge], dtype=np.uint32)
    gb = np.array([b for _ in edge for b in edge], dtype=np.uint32)
    out = ctx.field_ops(ga, gb)
    for i, (a, b) in enumerate(zip(ga.tolist(), gb.tolist())):
        x, y = a % P, b % P
        assert int(out[0, i]) == L.orc_add(x, y, P)
        assert int(out[1, i]) == L.orc_sub(x, y, P)
        assert int(out[2, i]) == L.orc_mul(x, y, P)
        assert int(out[3, i]) == (P - x) % P
        assert int(out[4, i]) == (L.orc_inv(x, P) if x else 0)
        assert int(out[5, i]) == L.orc_pow(x, b, P)
    rng = np.random.default_rng(2026)
    n = 1 << 20
    a = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    out = ctx.field_ops(a, b)
    x, y = a.astype(np.uint64) % P, b.astype(np.uint64) % P
    assert np.array_equal(out[0], ((x + y) % P).astype(np.uint32))
    assert np.array_equal(out[1], ((x + P - y) % P).astype(np.uint32))
    assert np.array_equal(out[2], ((x * y) % P).astype(np.uint32))             # < 2^64: exact in uint64
    assert np.array_equal(out[3], ((P - x) % P).astype(np.uint32))
    assert np.array_equal((out[4].astype(np.uint64) * x) % P, (x != 0).astype(np.uint64))   # a * a^-1 = 1
    for i in range(0, n, n // 2048):                                                # oracle on a sample of inv / pow
        xi = int(x[i])
        assert int(out[4, i]) == (L.orc_inv(xi, P) if xi else 0)
        assert int(out[5, i]) == L.orc_pow(xi, int(b[i]), P) == pow(xi, int(b[i]), P)


@pytest.mark.parametrize("log_n", [6, 10, 13])
def test_prove_ex_seeded_channel_and_channel_drawn_queries(ctx, log_n):
    """SURVEY 8(f) rank 2 (readme.md:1, prover.rs:263): a caller-chosen initial channel state and further test points drawn
    from the channel.  With neither it is the reference's proof; with them the product verifier replays the channel from
    the seed and checks every test point against proof.rs:63-148."""
    k = log_n
    base_tr, base_st = ctx.prove_fibsq(k)
    assert ctx.prove_fibsq_ex(k) == (base_tr, base_st)
    a_last = int(zk.trace_fibsq(k)[-1])
    # further test points on the reference's own channel: the reference transcript is a prefix
    tr2, st2 = ctx.prove_fibsq_ex(k, 3141592, None, 2)
    dl = zk.lib().zk_decommit_len(k)
    assert len(tr2) == len(base_tr) + 2 * (4 + dl) == zk.lib().zk_transcript_len_ex(k, 2)
    assert tr2[: len(base_tr)] == base_tr and st2 != base_st
    assert zk.verify_fibsq_ex(k, a_last, tr2, st2, None, 2) == (True, 0)
    # each further opening is what zk_open_queries returns for the channel-drawn position
    off = len(base_tr)
    xs = []
    for _ in range(2):
        xs.append(int.from_bytes(tr2[off: off + 4], "little") % ((8 << k) - 16))
        off += 4 + dl
    ctx.prove_fibsq(k)
    opened = ctx.open_queries(k, xs)
    off = len(base_tr)
    for o in opened:
        assert tr2[off + 4: off + 4 + dl] == o
        off += 4 + dl
    # seeded channel
    seed = bytes((7 * i + 1) & 0xFF for i in range(32))
    tr, st = ctx.prove_fibsq_ex(k, 3141592, seed, 3)
    assert tr[:32] == base_tr[:32]                       # the f root does not depend on the channel ...
    assert tr[32:44] != base_tr[32:44]                   # ... the challenges do
    assert zk.verify_fibsq_ex(k, a_last, tr, st, seed, 3) == (True, 0)
    assert zk.verify_fibsq_ex(k, a_last, tr, None, seed, 3, rederive=False) == (True, 0)   # the reference's checks alone
    ok, line = zk.verify_fibsq_ex(k, a_last, tr, st, None, 3)                               # replayed from the wrong state
    assert not ok and line == 22
    ok, line = zk.verify_fibsq_ex(k, a_last, tr, st, seed, 2)                               # wrong number of test points
    assert not ok and line == 20
    bad = bytearray(tr)
    bad[-40] ^= 1                                                                           # a digest of the last opening
    ok, line = zk.verify_fibsq_ex(k, a_last, bytes(bad), st, seed, 3)
    assert not ok and line in (136, 140, 6)
    bad = bytearray(tr)
    bad[len(base_tr) + 1] ^= 4                                                              # a channel-drawn query word
    ok, line = zk.verify_fibsq_ex(k, a_last, bytes(bad), st, seed, 3)
    assert not ok and line == 37
