"""Instruction roofline of the transforms: butterflies/s without memory traffic (zk_bench_kernel(8))."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk

ctx = zk.Context(0)
for mode, name in ((0, "butterfly (add, sub, montgomery product)"), (1, "montgomery product only"), (2, "add + sub only")):
    ms, mn, tb = ctx.bench_kernel(8, mode, iters=5, flush_l2=False)
    print(f"{name}: {tb * 1e3:.1f} G/s  ({mn * 1e3:.1f} us per launch)")
for mode, name in ((0, "ALU pipe"), (1, "FMA pipe (IMAD)"), (2, "1:1 mix")):
    ms, mn, t = ctx.bench_kernel(6, mode, iters=5, flush_l2=False)
    print(f"int pipe {name}: {t:.2f} Tiop/s")
