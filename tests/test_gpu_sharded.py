"""GPU tests of the sharded single-proof path (BASELINE configs[4]) with G virtual ranks on ONE device:
same kernels, same ownership and exchanges as the multi-process NCCL run (tools/sharded_run.py), the
collectives being tensor shuffles.  The transcript must equal the oracle's byte for byte."""
import hashlib

import numpy as np
import pytest

import zkstark_b200 as zk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tctx():
    import torch
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        c = zk.Context(0, stream=s.cuda_stream)
        yield c, s
        c.close()


@pytest.mark.parametrize("log_n,G,lb", [(8, 1, 5), (8, 2, 6), (9, 4, 5), (10, 8, 6), (12, 2, 10), (13, 8, 7), (14, 4, 10), (16, 8, 10)])
def test_sharded_virtual_ranks_match_oracle(tctx, orc, log_n, G, lb):
    import torch
    from zkstark_b200.sharded import LocalComm, ShardedProver
    ctx, s = tctx
    with torch.cuda.stream(s):
        trace = zk.trace_fibsq(log_n, 3141592 + G)
        prover = ShardedProver(ctx, log_n, LocalComm(G), log_block=lb, tail_len=0)   # shard every layer that can be sharded
        res = prover.prove(trace)
    res = [res] if G == 1 else res
    otr, ost, info = orc.prove_fibsq(log_n, 3141592 + G)
    for tr, st in res:   # every rank assembles the same proof
        assert hashlib.sha256(tr).hexdigest() == hashlib.sha256(otr).hexdigest()
        assert st == ost
    assert orc.verify_fibsq(log_n, info.a_last, res[0][0], True, res[0][1]) == 0


@pytest.mark.parametrize("G", [2, 8])
def test_sharded_full_size_matches_single_gpu(tctx, G):
    """BASELINE config 3 size (2^20 trace, 2^23-point LDE) through the sharded driver with G virtual ranks: every
    tile / leaf-pass regime of the block subtrees, 13 sharded FRI layers, the replicated tail and the decommit gather;
    the transcript must equal the monolithic pipeline's (itself checked against the oracle in test_gpu_parity)."""
    import torch
    from zkstark_b200.sharded import LocalComm, ShardedProver
    ctx, s = tctx
    log_n = 20
    with torch.cuda.stream(s):
        trace = zk.trace_fibsq(log_n, 3141592)
        want = ctx.prove_fibsq_from_trace(log_n, trace)
        res = ShardedProver(ctx, log_n, LocalComm(G)).prove(trace)
    for tr, st in res:
        assert hashlib.sha256(tr).hexdigest() == hashlib.sha256(want[0]).hexdigest()
        assert st == want[1]


def test_sharded_rejects_too_small(tctx):
    from zkstark_b200.sharded import LocalComm, ShardedProver
    ctx, _ = tctx
    with pytest.raises(ValueError):
        ShardedProver(ctx, 6, LocalComm(8), log_block=10)
    with pytest.raises(ValueError):
        ShardedProver(ctx, 10, LocalComm(3), log_block=5)
