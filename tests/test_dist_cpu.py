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


# ---------------------------------------------------------------- sharded single proof: exchange logic on CPU
def test_coset_to_block_exchange_is_blocked_cyclic():
    """LDE cosets -> one all-to-all -> blocked-cyclic ownership (zkstark_b200/sharded.py), with virtual ranks."""
    import torch
    from zkstark_b200.sharded import LocalComm, local_to_global, pack_cosets, unpack_blocks
    for G, log_n, lb in [(2, 8, 5), (4, 9, 6), (8, 10, 7), (1, 6, 4)]:
        n, N = 1 << log_n, 8 << log_n
        f_eval = torch.arange(N, dtype=torch.int32) * 7 + 3
        send = []
        for r in range(G):
            E = torch.stack([f_eval[(r + c * G)::8] for c in range(8 // G)])      # coset r + cG: f_eval[8j + coset]
            send.append(pack_cosets(E, G, lb))
        recv = LocalComm(G).all_to_all(send)
        for r in range(G):
            local = unpack_blocks(recv[r], G, lb)
            assert local.numel() == N // G
            idx = torch.tensor([local_to_global(i, G, r, lb) for i in range(N // G)])
            assert torch.equal(local, f_eval[idx])
            # fold partners stay on the rank: i and i + N/2 map to local indices loc and loc + N/(2G)
            half = N // G // 2
            assert torch.equal(idx[:half] + N // 2, idx[half:])


def _comm_worker(rank, ws, port, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from zkstark_b200.sharded import TorchComm
        comm = TorchComm()
        x = torch.arange(8, dtype=torch.int32).view(ws, -1) + 100 * rank
        a2a = comm.all_to_all([x])[0]
        ag = comm.all_gather([torch.tensor([rank, rank + 10], dtype=torch.int32)])[0]
        buf = torch.zeros(6, dtype=torch.uint8)
        buf[rank * 3: rank * 3 + 3] = torch.tensor([1 + rank, 200, 7], dtype=torch.uint8)   # disjoint owners
        mx = comm.all_reduce_max([buf])[0]
        q.put((rank, a2a.tolist(), ag.tolist(), mx.tolist()))
    finally:
        dist.destroy_process_group()


def test_torch_comm_matches_local_comm_gloo():
    """The real collectives (here gloo, on the GPU box NCCL) and the virtual-rank shuffles agree."""
    import torch
    from zkstark_b200.sharded import LocalComm
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_comm_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lc = LocalComm(ws)
    xs = [torch.arange(8, dtype=torch.int32).view(ws, -1) + 100 * r for r in range(ws)]
    want_a2a = lc.all_to_all(xs)
    want_ag = lc.all_gather([torch.tensor([r, r + 10], dtype=torch.int32) for r in range(ws)])
    bufs = []
    for r in range(ws):
        b = torch.zeros(6, dtype=torch.uint8)
        b[r * 3: r * 3 + 3] = torch.tensor([1 + r, 200, 7], dtype=torch.uint8)
        bufs.append(b)
    want_mx = lc.all_reduce_max(bufs)
    for rank, a2a, ag, mx in res:
        assert a2a == want_a2a[rank].view(ws, -1).tolist() or a2a == want_a2a[rank].tolist()
        assert ag == want_ag[rank].tolist()
        assert mx == want_mx[rank].tolist()
