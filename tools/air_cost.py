"""Cost of the general AIR code path: prove() of one 2^log_n trace and a 32-proof batch with the reference's
program through the specialised code, through the general code (ZK_AIR_NO_SPECIALISE), and with an every-term program;
host trace generation of 32 traces beside it."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zkstark_b200 as zk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = zk.Context(0)
secrets = (np.arange(32, dtype=np.uint32) + 3141592).astype(np.uint32)


def timed(f, reps=5):
    f()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


for name, air, ns in (("fibsq specialised", None, False), ("fibsq general code", zk.AIR_FIBSQ, True), ("every-term program", (7, 3, 5, 11, 13, 17, 19), False)):
    ctx.set_air(air, no_specialise=ns) if air else ctx.set_air(None)
    t0 = time.perf_counter()
    traces = zk.trace_air_batch(air, k, secrets, threads=0, no_specialise=ns)
    t_tr = (time.perf_counter() - t0) * 1e3
    one = timed(lambda: ctx.prove_fibsq_from_trace(k, traces[0]))
    batch = timed(lambda: ctx.prove_fibsq_batch(k, traces), reps=3)
    print(f"{name:22s} log_n {k}: prove {one:.3f} ms, 32-proof batch from host traces {batch:.1f} ms, 32 host traces {t_tr:.1f} ms", flush=True)
ctx.set_air(None)
