"""SURVEY 8(f4) on the GPU: zk_ctx_set_air through every prove entry point, against the oracle (tests/test_air.py pins the
oracle's general code: literal == pointwise route, FibonacciSq point == the reference's golden)."""
import hashlib
import json
import os

import numpy as np
import pytest

import zkstark_b200 as zk

pytestmark = pytest.mark.gpu
P = 3221225473
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AIRS = [
    (1, 0, 0, 0, 1, 1, 0),
    (7, 3, 5, 11, 13, 17, 19),
    (2, 0, 1, 1, 0, 0, 5),
    (3, 2**32 - 1, P + 1, 7, 0, 9, P),
]


@pytest.fixture()
def air_ctx(ctx):
    yield ctx
    ctx.set_air(None)                       # the session context goes back to the reference's program
    ctx.bench_set(-101, 0)


def test_general_code_on_the_reference_program_gives_the_golden_transcripts(air_ctx):
    """ZK_AIR_NO_SPECIALISE: FibonacciSq through the general trace generator and the general transition numerator."""
    ctx = air_ctx
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "transcripts.json")))
    ctx.set_air(zk.AIR_FIBSQ, no_specialise=True)
    assert ctx.get_air() == (zk.AIR_FIBSQ, 1)
    for form in (-101, -100):               # constant tables / per-point form of the composition polynomial
        ctx.bench_set(form, 0)
        for log_n in (10, 16):
            tr, st = ctx.prove_fibsq(log_n, 3141592)
            assert hashlib.sha256(tr).hexdigest() == gold[str(log_n)]["sha256"]
            assert st.hex() == gold[str(log_n)]["state"]
    ctx.bench_set(-101, 0)
    big = json.load(open(os.path.join(ROOT, "tests", "golden", "large_digests.json")))["20"]
    tr, st = ctx.prove_fibsq(20, 3141592)
    assert hashlib.sha256(tr).hexdigest() == big["sha256"] and st.hex() == big["state"]


@pytest.mark.parametrize("air", AIRS)
@pytest.mark.parametrize("log_n", [4, 9, 13, 16])
def test_prove_matches_oracle(air_ctx, orc, air, log_n):
    ctx = air_ctx
    ctx.set_air(air)
    want_tr, want_st, info = orc.prove_air(air, log_n, secret=2024 + log_n)
    for form in (-101, -100):
        ctx.bench_set(form, 0)
        tr, st = ctx.prove_fibsq(log_n, 2024 + log_n)
        assert tr == want_tr and st == want_st
    assert zk.verify_air(air, log_n, info.a_last, tr, st) == (True, 0)
    assert zk.verify_air(None, log_n, info.a_last, tr, st)[0] is False
    # from a trace handed in
    tr2, st2 = ctx.prove_fibsq_from_trace(log_n, zk.trace_air(air, log_n, 2024 + log_n))
    assert tr2 == want_tr and st2 == want_st


@pytest.mark.parametrize("air", AIRS[:2])
def test_cp_eval_per_function(air_ctx, orc, air):
    ctx = air_ctx
    ctx.set_air(air)
    log_n = 12
    f = np.random.default_rng(5).integers(0, P, size=8 << log_n, dtype=np.uint64).astype(np.uint32)
    al = [361545003, 3235878091, 2708123352]
    want = orc.cp_eval_air(air, f, log_n, al, 7, 99, threads=orc.default_threads())
    for form in (-101, -100):
        ctx.bench_set(form, 0)
        assert np.array_equal(ctx.cp_eval_fibsq(f, log_n, al, 7, 99), want)


def test_batches_and_pipeline(air_ctx, orc):
    """batch from traces (device channel tail at >= 8 proofs), batch from secrets, submit/collect pipeline"""
    ctx = air_ctx
    air = AIRS[1]
    ctx.set_air(air)
    log_n, n = 11, 19
    secrets = (np.arange(n, dtype=np.uint32) * 7919 + 3).astype(np.uint32)
    want = [orc.prove_air(air, log_n, secret=int(s))[:2] for s in secrets]
    trs, sts = ctx.prove_fibsq_batch(log_n, zk.trace_air_batch(air, log_n, secrets))
    for j in range(n):
        assert bytes(trs[j]) == want[j][0] and bytes(sts[j]) == want[j][1]
    trs, sts = ctx.prove_fibsq_batch_secrets(log_n, secrets)
    for j in range(n):
        assert bytes(trs[j]) == want[j][0] and bytes(sts[j]) == want[j][1]
    ctx.prove_stream_submit(log_n, secrets[:8])
    ctx.prove_stream_submit(log_n, secrets[8:])
    with pytest.raises(zk.ZkError):
        ctx.set_air(None)                   # not while a submitted batch is uncollected
    a = ctx.prove_stream_collect(log_n, 8)
    b = ctx.prove_stream_collect(log_n, n - 8)
    got = [bytes(x) for x in a[0]] + [bytes(x) for x in b[0]]
    assert got == [w[0] for w in want]


@pytest.mark.parametrize("G,lb", [(2, 6), (8, 5)])
def test_sharded_driver(air_ctx, orc, G, lb):
    ctx = air_ctx
    air = AIRS[3]
    ctx.set_air(air)
    log_n = 12
    comm = zk.ShardComm.local(ctx, G)
    try:
        comm.set(log_block=lb, log_tail_len=0)
        tr, st = comm.prove(log_n, zk.trace_air(air, log_n, 31337))
    finally:
        comm.close()
    want_tr, want_st, info = orc.prove_air(air, log_n, secret=31337)
    assert tr == want_tr and st == want_st


def test_seeded_channel_and_further_queries(air_ctx, orc):
    """zk_prove_fibsq_ex under another program: the reference-shaped transcript is a prefix, the whole verifies under the
    program (seed, 5 further channel-drawn test points) and under no other"""
    ctx = air_ctx
    air = AIRS[2]
    ctx.set_air(air)
    log_n = 10
    seed = hashlib.sha256(b"air seed").digest()
    base_tr, base_st = ctx.prove_fibsq(log_n, 55)
    tr, st = ctx.prove_fibsq_ex(log_n, 55, None, 5)
    assert tr[: len(base_tr)] == base_tr
    a_last = int(zk.trace_air(air, log_n, 55)[-1])
    assert zk.verify_air(air, log_n, a_last, tr, st, None, 5) == (True, 0)
    tr, st = ctx.prove_fibsq_ex(log_n, 55, seed, 3)
    assert zk.verify_air(air, log_n, a_last, tr, st, seed, 3) == (True, 0)
    assert not zk.verify_air(air, log_n, a_last, tr, st, None, 3)[0]
    assert not zk.verify_air(AIRS[0], log_n, a_last, tr, st, seed, 3)[0]


def test_full_size_program(air_ctx, orc):
    """BASELINE configs[2] size (2^20 trace, 2^23-point LDE) under an every-term program, against the oracle"""
    ctx = air_ctx
    air = AIRS[1]
    ctx.set_air(air)
    want_tr, want_st, info = orc.prove_air(air, 20, secret=3141592)
    tr, st = ctx.prove_fibsq(20, 3141592)
    assert hashlib.sha256(tr).hexdigest() == hashlib.sha256(want_tr).hexdigest() and st == want_st
    assert zk.verify_air(air, 20, info.a_last, tr, st) == (True, 0)


def test_device_trace_generator(orc, monkeypatch):
    """ZK_TRACE_DEVICE=1 (read at context creation): the traces of a submitted batch come from trace_batch_kernel instead
    of the host threads -- the reference's program and another one, 37 proofs (a full warp of proofs + a ragged one)."""
    monkeypatch.setenv("ZK_TRACE_DEVICE", "1")
    ctx = zk.Context(0)
    try:
        log_n, n = 10, 37
        secrets = (np.arange(n, dtype=np.uint32) * 104729 + 11).astype(np.uint32)
        for air in (None, AIRS[1]):
            ctx.set_air(air)
            ctx.prove_stream_submit(log_n, secrets)
            trs, sts = ctx.prove_stream_collect(log_n, n)
            for j in (0, 1, 31, 32, 36):
                want = orc.prove_air(air, log_n, secret=int(secrets[j])) if air else orc.prove_fibsq(log_n, int(secrets[j]))
                assert bytes(trs[j]) == want[0] and bytes(sts[j]) == want[1]
    finally:
        ctx.close()


def test_api_mirror(air_ctx, orc):
    from zkstark_b200 import api
    air = AIRS[0]
    proof = api.generate_proof(api.Channel.new(), 9, 17, ctx=air_ctx, air=air)
    a_last = int(zk.trace_air(air, 9, 17)[-1])
    proof.verify(9, a_last)
    assert air_ctx.get_air() == (zk.AIR_FIBSQ, 0)
    assert (proof.data, proof.state) == orc.prove_air(air, 9, secret=17)[:2]
    with pytest.raises(zk.ZkError):
        api.Proof(proof.state, proof.data).verify(9, a_last)          # not a FibonacciSq proof


def test_bad_air(ctx):
    import ctypes as C
    bad = (C.c_uint32 * 8)(1, 1, 1, 0, 0, 0, 0, 2)           # unknown flag
    with pytest.raises(zk.ZkError):
        ctx._ck(zk.lib().zk_ctx_set_air(ctx._h, bad))
    assert ctx.get_air() == (zk.AIR_FIBSQ, 0)
