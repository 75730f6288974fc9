#!/usr/bin/env python
"""One real rank of the native sharded prover (zk_prove_fibsq_sharded over NCCL), one process per GPU.
Launched N times by tests/test_gpu_sharded_native.py; the ncclUniqueId travels through a file (no torch.distributed).

  python tools/shard_nccl_worker.py --rank R --world N --log-n K --id-file PATH --out PATH.json
"""
import argparse
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--log-n", type=int, default=18)
    ap.add_argument("--id-file", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--more", default="", help="further sizes proven on the same communicator before --log-n, e.g. 14,16 "
                    "(growing sizes re-allocate and re-map the peer buffers)")
    ap.add_argument("--air", default="", help="a0,c20,c02,c11,c10,c01,c00: another program for the --more proofs")
    ap.add_argument("--no-p2p", action="store_true", help="NCCL collectives only (ZK_SHARD_NO_P2P=1)")
    a = ap.parse_args()
    if a.no_p2p:
        os.environ["ZK_SHARD_NO_P2P"] = "1"
    import zkstark_b200 as zk
    ctx = zk.Context(a.rank)
    if a.rank == 0:
        uid = zk.shard_unique_id()
        tmp = a.id_file + ".tmp"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, a.id_file)
    else:
        t0 = time.time()
        while not os.path.exists(a.id_file):
            if time.time() - t0 > 120:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
        uid = open(a.id_file, "rb").read()
    comm = zk.ShardComm(ctx, a.world, a.rank, uid)
    more = []
    air = tuple(int(v) for v in a.air.split(",")) if a.air else None
    for k in [int(v) for v in a.more.split(",") if v]:
        ctx.set_air(air)
        t = zk.trace_air(air, k, 3141592) if air else zk.trace_fibsq(k, 3141592)
        tr, st = comm.prove(k, t)
        more.append({"log_n": k, "transcript_sha256": hashlib.sha256(tr).hexdigest(), "state": st.hex()})
    ctx.set_air(None)
    trace = zk.trace_fibsq(a.log_n, 3141592)
    for _ in range(a.reps):
        tr, st = comm.prove(a.log_n, trace)
    out = {"rank": a.rank, "transcript_sha256": hashlib.sha256(tr).hexdigest(), "state": st.hex(),
           "phase_ms": comm.last_phase_ms(), "comm_bytes": comm.last_comm_bytes(), "first_tail": comm.last_first_tail(), "more": more}
    with open(a.out, "w") as f:
        json.dump(out, f)
    comm.close()
    ctx.close()


if __name__ == "__main__":
    main()
