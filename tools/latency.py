"""Single-proof latency (device stage times, best of N) at the BASELINE sizes: python tools/latency.py [log_n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk

ctx = zk.Context(0)
ctx.set_stage_timing(True)
for k in [int(a) for a in sys.argv[1:]] or [10, 16, 20]:
    tr = zk.trace_fibsq(k)
    best, wall = None, None
    for _ in range(12):
        t0 = time.perf_counter()
        ctx.prove_fibsq_from_trace(k, tr)
        w = 1e3 * (time.perf_counter() - t0)
        st = ctx.last_stage_ms()
        if best is None or st["total"] < best["total"]:
            best = st
        wall = w if wall is None else min(wall, w)
    print(f"2^{k}:", {a: round(b, 4) for a, b in best.items()}, "call wall ms", round(wall, 4), flush=True)
