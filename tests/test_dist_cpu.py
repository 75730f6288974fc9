"""CPU tests (gloo, world_size 2) of the N>1 host logic: work split, max-over-ranks timing, result gathering.
Proofs themselves are produced by the oracle here (no GPU in this container); the GPU path plugs into the
same helpers in bench.py."""
import hashlib
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from zkstark_b200 import dist as zd


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 1025):
        for ws in (1, 2, 3, 4, 8):
            spans = [zd.shard_range(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        zd.shard_range(4, 2, 2)


def test_secrets_follow_survey_rule():
    assert zd.proof_secret(0) == 3141592 and zd.proof_secret(1023) == 3141592 + 1023


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_proofs, log_n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        import oracle
        lo, hi = zd.shard_range(n_proofs, ws, rank)
        trs = [oracle.prove_fibsq(log_n, zd.proof_secret(j), threads=1)[0] for j in range(lo, hi)]
        t = zd.max_over_ranks(10.0 + rank)            # slowest rank defines the time
        total = zd.sum_over_ranks(hi - lo)
        digests = zd.gather_transcript_digests(trs)
        q.put((rank, t, total, digests))
    finally:
        dist.destroy_process_group()


def test_two_rank_batch_gloo(orc):
    ws, n_proofs, log_n = 2, 5, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_proofs, log_n, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [hashlib.sha256(orc.prove_fibsq(log_n, zd.proof_secret(j), threads=1)[0]).hexdigest() for j in range(n_proofs)]
    for rank, t, total, digests in res:
        assert t == 11.0            # max over ranks
        assert total == n_proofs    # every proof done exactly once
        assert digests == want      # global order, identical on every rank
