"""The product verifier (zk_verify_fibsq, host C++; proof.rs:15-149) against the oracle's restatement: same
accept/reject decisions on honest and tampered transcripts.  CPU only."""
import numpy as np
import pytest

import zkstark_b200 as zk

P = 3221225473


@pytest.mark.parametrize("log_n", [3, 4, 7, 10, 12])
def test_accepts_honest_proofs(orc, log_n):
    tr, st, info = orc.prove_fibsq(log_n, 3141592 + log_n)
    assert zk.verify_fibsq(log_n, info.a_last, tr, st, rederive=True) == (True, 0)
    assert zk.verify_fibsq(log_n, info.a_last, tr, None, rederive=False) == (True, 0)  # the reference's own checks only


def test_canonical_proof_size_and_verify(orc, transcripts):
    tr, st, info = orc.prove_fibsq(10)
    proof = zk.Proof(st, tr)
    proof.verify()                      # main.rs:28
    assert proof.size() == 7884         # main.rs:35
    assert st.hex() == transcripts["10"]["state"]


def test_rejects_like_the_oracle(orc):
    log_n = 6
    tr, st, info = orc.prove_fibsq(log_n)
    rng = np.random.default_rng(0)
    positions = [0, 31, 32, 40, 44, 75, 76, 80, 120, len(tr) - 1] + list(rng.integers(0, len(tr), 40))
    for pos in positions:
        bad = bytearray(tr)
        bad[int(pos)] ^= 1 << int(rng.integers(0, 8))
        ok, line = zk.verify_fibsq(log_n, info.a_last, bytes(bad), None, rederive=True)
        want = orc.verify_fibsq(log_n, info.a_last, bytes(bad), True, None) == 0
        assert ok == want and not ok, pos
        assert line > 0
    # wrong public input, wrong state, truncated transcript
    assert zk.verify_fibsq(log_n, (info.a_last + 1) % P, tr, st)[0] is False
    assert zk.verify_fibsq(log_n, info.a_last, tr, bytes(32)) == (False, 6)
    assert zk.verify_fibsq(log_n, info.a_last, tr[:-1], st) == (False, 20)


def test_reference_checks_alone_miss_forged_challenges(orc):
    """The reference reads alpha/beta from the proof (proof.rs:22-37): a transcript whose FRI root of the LAST
    layer is replaced passes nothing here, but flipping an unused high bit pattern in a challenge that reduces
    to the same field element is only caught when the channel is replayed."""
    log_n = 5
    tr, st, info = orc.prove_fibsq(log_n)
    import struct
    for j in range(3):
        a = struct.unpack_from("<I", tr, 32 + 4 * j)[0]
        if a + P < 2 ** 32:                      # same residue, different raw word
            forged = bytearray(tr)
            struct.pack_into("<I", forged, 32 + 4 * j, a + P)
            assert zk.verify_fibsq(log_n, info.a_last, bytes(forged), None, rederive=False)[0] is True
            assert zk.verify_fibsq(log_n, info.a_last, bytes(forged), None, rederive=True)[0] is False
            return
    pytest.skip("no alpha below 2^32 - p in this transcript")


def test_proof_wire_format_round_trip(orc, transcripts):
    """state || u64 length || transcript: what bincode 1.x writes for `struct Proof`."""
    import zkstark_b200 as zk
    tr, st, info = orc.prove_fibsq(10, 3141592)
    blob = zk.Proof(st, tr).to_bytes()
    assert len(blob) == 32 + 8 + 7836 and blob[32:40] == (7836).to_bytes(8, "little")
    back = zk.Proof.from_bytes(blob)
    assert back.state == st and back.data == tr
    back.verify()                                      # product verifier (host C++), replays the channel
    with pytest.raises(zk.ZkError):
        zk.Proof.from_bytes(blob[:-1])
    with pytest.raises(zk.ZkError):
        zk.Proof.from_bytes(blob[:20])


def test_every_byte_of_the_transcript_is_checked(orc):
    """Exhaustive: one flipped bit in ANY byte of a small proof is rejected once the channel is replayed (the verifier
    leaves no byte of the transcript unbound), by the product verifier and the oracle's alike."""
    log_n = 3
    tr, st, info = orc.prove_fibsq(log_n, 77)
    rng = np.random.default_rng(1)
    for pos in range(len(tr)):
        bad = bytearray(tr)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        ok, line = zk.verify_fibsq(log_n, info.a_last, bytes(bad), st, rederive=True)
        assert not ok and line > 0, pos
        assert orc.verify_fibsq(log_n, info.a_last, bytes(bad), True, st) != 0, pos
