"""LDE + interpolation of ONE size a few times (driver for ncu): python tools/ntt_one.py [log_n] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
it = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = zk.Context(0)
_, lde, _ = ctx.bench_kernel(2, k, iters=it, flush_l2=False)
_, itp, _ = ctx.bench_kernel(3, k, iters=it, flush_l2=False)
print(f"2^{k}: lde {lde * 1e3:.1f} us interp {itp * 1e3:.1f} us")
