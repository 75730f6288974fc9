"""GPU tests of the native sharded prover (csrc/sharded.cu, zk_prove_fibsq_sharded; BASELINE configs[4]).

* G virtual ranks on ONE device (zk_shard_comm_init_local): the same C++ driver, kernels, ownership and buffer
  layouts as the multi-process run, the collectives being device copies -- how the single-GPU tier covers G = 2, 4, 8;
* two real ranks over NCCL, one process per GPU (skipped unless the box has >= 2 GPUs): tools/shard_nccl_worker.py
  is launched twice, the unique id travels through a file.
The transcript must equal the oracle's (and the golden digests) byte for byte on every rank."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import zkstark_b200 as zk

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("log_n,G,lb", [(8, 1, 5), (8, 2, 6), (9, 4, 5), (10, 8, 6), (12, 2, 10), (13, 8, 7), (14, 4, 10),
                                         (16, 8, 10), (16, 1, 10)])
def test_native_virtual_ranks_match_oracle(ctx, orc, log_n, G, lb):
    comm = zk.ShardComm.local(ctx, G)
    try:
        comm.set(log_block=lb, log_tail_len=0)          # shard every layer that can be sharded
        trace = zk.trace_fibsq(log_n, 3141592 + G)
        tr, st = comm.prove(log_n, trace)               # the driver itself checks that all virtual ranks agree
        otr, ost, info = orc.prove_fibsq(log_n, 3141592 + G)
        assert hashlib.sha256(tr).hexdigest() == hashlib.sha256(otr).hexdigest()
        assert st == ost
        assert orc.verify_fibsq(log_n, info.a_last, tr, True, st) == 0
        if G > 1:
            assert comm.last_comm_bytes() > 4 * (8 << log_n) * (G - 1) // G // G      # per rank: at least its share of the all-to-all
        # a second proof on the same communicator reuses every buffer
        tr2, st2 = comm.prove(log_n, trace)
        assert tr2 == tr and st2 == st
    finally:
        comm.close()


@pytest.mark.parametrize("G", [1, 2, 8])
def test_native_full_size_matches_single_gpu_and_golden(ctx, G):
    """BASELINE configs[2] size through the sharded driver (default block / tail settings): every leaf-pass regime of
    the block subtrees, the sharded FRI layers, the replicated tail, the decommit gather."""
    log_n = 20
    trace = zk.trace_fibsq(log_n, 3141592)
    want = ctx.prove_fibsq_from_trace(log_n, trace)
    comm = zk.ShardComm.local(ctx, G)
    try:
        tr, st = comm.prove(log_n, trace)
        ph = comm.last_phase_ms()
    finally:
        comm.close()
    assert tr == want[0] and st == want[1]
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "large_digests.json")))
    assert hashlib.sha256(tr).hexdigest() == gold["20"]["sha256"]
    assert ph["total"] > 0


@pytest.mark.parametrize("log_n,G,lb", [(9, 4, 5), (12, 2, 10), (13, 8, 7)])
def test_native_all_to_all_path_matches_peer_path(ctx, orc, log_n, G, lb, monkeypatch):
    """ZK_SHARD_NO_P2P=1: the coset transforms fill a local send buffer and an all-to-all re-owns the evaluations, instead
    of storing straight into the owners' receive buffers (the default, fused path).  Same transcript either way."""
    trace = zk.trace_fibsq(log_n, 99)
    otr, ost, _ = orc.prove_fibsq(log_n, 99)
    for no_p2p in ("1", "0"):
        monkeypatch.setenv("ZK_SHARD_NO_P2P", no_p2p)
        comm = zk.ShardComm.local(ctx, G)
        try:
            comm.set(log_block=lb, log_tail_len=0)
            assert comm.prove(log_n, trace) == (otr, ost)
        finally:
            comm.close()


def test_native_single_real_rank(ctx, orc):
    """zk_shard_comm_init with n_ranks = 1: the real-rank code path without NCCL (no peers, no collectives)."""
    comm = zk.ShardComm(ctx, 1, 0, bytes(128))
    try:
        trace = zk.trace_fibsq(13, 5)
        assert comm.prove(13, trace) == orc.prove_fibsq(13, 5)[:2]
        assert comm.last_comm_bytes() == 0
    finally:
        comm.close()


def test_native_rejects_bad_arguments(ctx, orc):
    with pytest.raises(zk.ZkError):
        zk.ShardComm.local(ctx, 3)
    comm = zk.ShardComm.local(ctx, 8)
    try:
        with pytest.raises(zk.ZkError):
            comm.prove(5, zk.trace_fibsq(5))         # 256 evaluations: not even two blocks of 32 per rank
        # the smallest proof 8 ranks can share with the default (automatic) block size: two blocks of 32 points each
        tr6, st6 = comm.prove(6, zk.trace_fibsq(6, 99))
        assert (tr6, st6) == orc.prove_fibsq(6, 99)[:2]
        comm.set(log_block=10, log_tail_len=0xFFFFFFFF)
        with pytest.raises(zk.ZkError):
            comm.prove(6, zk.trace_fibsq(6))         # too small for 8 ranks x 2 blocks of 1024
        tr = np.empty(16, dtype=np.uint8)
        st = np.empty(32, dtype=np.uint8)
        t = zk.trace_fibsq(16)
        with pytest.raises(zk.ZkError):
            comm.prove_ptr(16, t.ctypes.data, tr.ctypes.data, 16, st.ctypes.data)   # transcript buffer too small
    finally:
        comm.close()


@pytest.mark.parametrize("no_p2p", [False, True])
def test_native_nccl_two_processes(tmp_path, orc, no_p2p):
    """Real ranks: 2 processes, 2 GPUs; exchanges through peer memory (flag barriers) or, with ZK_SHARD_NO_P2P, NCCL
    all-to-all / all-gathers / ring / all-reduce.  The communicator first proves smaller sizes of another program (SURVEY
    8(f4)) in growing order -- the peer buffers are re-allocated and re-mapped between them -- then 2^18 of the reference's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    log_n = 18
    air = (7, 3, 5, 11, 13, 17, 19)
    idfile = str(tmp_path / "nccl_id")
    worker = os.path.join(ROOT, "tools", "shard_nccl_worker.py")
    extra = ["--no-p2p"] if no_p2p else []
    procs = [subprocess.Popen([sys.executable, worker, "--rank", str(r), "--world", "2", "--log-n", str(log_n), "--id-file", idfile,
                               "--out", str(tmp_path / f"out{r}.json"), "--more", "12,16,14", "--air", ",".join(map(str, air))] + extra)
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "large_digests.json")))
    want_more = {k: orc.prove_air(air, k, secret=3141592) for k in (12, 16, 14)}
    for r in range(2):
        o = json.load(open(tmp_path / f"out{r}.json"))
        assert o["transcript_sha256"] == gold[str(log_n)]["sha256"]
        assert o["state"] == gold[str(log_n)]["state"]
        assert [m["log_n"] for m in o["more"]] == [12, 16, 14]
        for m in o["more"]:
            w = want_more[m["log_n"]]
            assert m["transcript_sha256"] == hashlib.sha256(w[0]).hexdigest() and m["state"] == w[1].hex()
