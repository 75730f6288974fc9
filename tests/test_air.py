"""the program being proven as data (zk_air / orc_air) -- CPU side.

The reference proves one hard-coded program (FibonacciSq: prover.rs:32-39, :134-138).  The AIR family keeps every other
step of its protocol and replaces the recurrence; there is no reference code for the other programs, so parity is pinned
three ways: (1) at the FibonacciSq point the general code must give the reference's golden transcript, (2) the oracle's
LITERAL route (polynomial algebra: apply_const / Mul / div, polynomial.rs) and its pointwise route agree byte for byte,
(3) the verifier (proof.rs:15-149 with the program's constraint) accepts, and rejects under any other program."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle as orc
import zkstark_b200 as zk

P = 3221225473
AIRS = [
    (1, 0, 0, 0, 1, 1, 0),                                  # Fibonacci (degree 1)
    (7, 3, 5, 11, 13, 17, 19),                              # every term present
    (2, 0, 1, 1, 0, 0, 5),                                  # a[i+2] = a[i]^2 + a[i] a[i+1] + 5
    (0, 0, 0, 1, 0, 0, 1),                                  # a[i+2] = a[i] a[i+1] + 1 from a[0] = 0
    (3, 2**32 - 1, P + 1, 7, 0, 9, P),                      # coefficients >= p are reduced (field.rs:20-24)
]
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "transcripts.json")))


def test_fibsq_point_is_the_reference_program():
    tr, st, info = orc.prove_air(orc.AIR_FIBSQ, 10)
    assert (tr, st) == orc.prove_fibsq(10)[:2]
    assert np.array_equal(orc.trace_air(orc.AIR_FIBSQ, 10), orc.trace_fibsq(10))
    assert orc.trace_air(orc.AIR_FIBSQ, 10)[1022] == 2338775057          # prover.rs:42


@pytest.mark.parametrize("air", AIRS)
@pytest.mark.parametrize("log_n", [3, 5, 7])
def test_oracle_literal_route_equals_pointwise_route(air, log_n):
    fast = orc.prove_air(air, log_n, secret=12345)
    lit = orc.prove_air(air, log_n, secret=12345, route=orc.ROUTE_LITERAL)
    assert fast[0] == lit[0] and fast[1] == lit[1]
    assert orc.verify_air(air, log_n, fast[2].a_last, fast[0], True, fast[1]) == 0
    assert orc.verify_fibsq(log_n, fast[2].a_last, fast[0], True, fast[1]) != 0     # not a FibonacciSq proof


@pytest.mark.parametrize("air", AIRS + [orc.AIR_FIBSQ])
def test_host_traces_match_oracle(air):
    for log_n in (1, 2, 3, 6, 11):
        assert np.array_equal(zk.trace_air(air, log_n, 4242), orc.trace_air(air, log_n, 4242))
    secrets = (np.arange(21, dtype=np.uint64) * 2654435761 % (1 << 32)).astype(np.uint32)     # 8 + 8 + 4 + 1: every SIMD group size
    for ns in (False, True):
        rows = zk.trace_air_batch(air, 9, secrets, threads=3, no_specialise=ns)
        for j, s in enumerate(secrets):
            assert np.array_equal(rows[j], orc.trace_air(air, 9, int(s)))
    assert zk.trace_air(air, 4, 9)[0] == air[0] % P


def test_trace_obeys_the_recurrence():
    air = AIRS[1]
    a = [int(v) for v in zk.trace_air(air, 8, 99)]
    a0, c20, c02, c11, c10, c01, c00 = air
    for i in range(len(a) - 2):
        assert a[i + 2] == (c20 * a[i + 1] ** 2 + c02 * a[i] ** 2 + c11 * a[i] * a[i + 1] + c10 * a[i + 1] + c01 * a[i] + c00) % P


@pytest.mark.parametrize("air", AIRS)
def test_product_verifier_on_oracle_proofs(air):
    tr, st, info = orc.prove_air(air, 8, secret=777)
    assert zk.verify_air(air, 8, info.a_last, tr, st) == (True, 0)
    for other in AIRS + [orc.AIR_FIBSQ]:
        if tuple(v % P for v in other) != tuple(v % P for v in air):
            assert zk.verify_air(other, 8, info.a_last, tr, st) == (False, 76)      # proof.rs:76, the CP identity
    assert zk.verify_air(air, 8, (info.a_last + 1) % P, tr, st) == (False, 76)
    bad = bytearray(tr); bad[100] ^= 1
    assert not zk.verify_air(air, 8, info.a_last, bytes(bad), st)[0]


def test_null_air_is_fibsq():
    tr, st, info = orc.prove_fibsq(10)
    assert hashlib.sha256(tr).hexdigest() == GOLD["10"]["sha256"]
    assert zk.verify_air(None, 10, info.a_last, tr, st) == (True, 0)
    assert zk.verify_air(orc.AIR_FIBSQ, 10, info.a_last, tr, st) == (True, 0)
    with pytest.raises(ValueError):
        zk.trace_air((1, 2, 3), 4)
