"""Time of the device trace generation alone and of secrets -> proofs with device / host traces: python tools/trace_time.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import zkstark_b200 as zk

k, nb = 20, 32
ctx = zk.Context(0)
sec = np.arange(nb, dtype=np.uint32) + 3141592
for _ in range(2):
    ctx.prove_fibsq_batch_secrets(k, sec)
t = []
for _ in range(3):
    t0 = time.perf_counter()
    ctx.prove_fibsq_batch_secrets(k, sec)
    t.append(1e3 * (time.perf_counter() - t0))
print("single call, secrets in, ms:", [round(x, 1) for x in t], "ZK_TRACE_DEVICE =", os.environ.get("ZK_TRACE_DEVICE"))
# pipelined
ctx.prove_stream_submit(k, sec)
t0 = time.perf_counter()
for i in range(6):
    ctx.prove_stream_submit(k, sec)
    ctx.prove_stream_collect(k, nb)
dt = 1e3 * (time.perf_counter() - t0) / 6
ctx.prove_stream_collect(k, nb)
print("pipelined ms per step:", round(dt, 1))
