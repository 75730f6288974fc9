"""Multi-process run of one sharded proof (torchrun --nproc-per-node G tools/sharded_run.py LOG_N [REPS]).
Rank 0 prints one JSON line: time per proof (max over ranks, CUDA events), LDE elements/s, bytes over NVLink,
and whether the transcript equals the single-GPU one (log_n <= 22) / the oracle's (log_n <= 20)."""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import zkstark_b200 as zk
from zkstark_b200.sharded import LocalComm, ShardedProver, TorchComm

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ctx = zk.Context(local, stream=s.cuda_stream)
    comm = TorchComm() if world > 1 else LocalComm(1)
    prover = ShardedProver(ctx, k, comm)
    trace_np = zk.trace_fibsq(k, 3141592)
    trace = torch.from_numpy(trace_np.view(np.int32)).pin_memory()
    times = []
    for it in range(reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr, st = prover.prove(trace)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it:
            times.append(dt)
    best = min(times)
    prof = None
    if os.environ.get("ZK_SHARD_PROFILE"):
        prover.profile = {}
        prover.prove(trace)
        prof = {k_: round(1e3 * v, 3) for k_, v in prover.profile.items()}
        prover.profile = None
    tmax = torch.tensor([best], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ok = None
    if rank == 0:
        if k <= 24:
            c1 = zk.Context(local)
            tr1, st1 = c1.prove_fibsq_from_trace(k, trace_np)
            ok = bool(tr1 == tr and st1 == st)
            c1.close()
        print(json.dumps({"sharded_prove": True, "log_n": k, "n_gpus": world, "ms_per_proof": 1e3 * float(tmax.item()),
                          "proofs_per_s": 1.0 / float(tmax.item()), "nvlink_bytes_per_proof": prover.comm_bytes,
                          "first_replicated_layer": prover.first_tail_layer, "transcript_sha256": hashlib.sha256(tr).hexdigest(),
                          "matches_single_gpu": ok, "phase_ms": prof}), flush=True)
    ctx.close()
if world > 1:
    dist.destroy_process_group()
