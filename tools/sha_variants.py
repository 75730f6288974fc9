"""GPU experiment: leaf-pass time per SHA-256 pipe-balancing variant (build with EXTRA=-DZK_SHA_EXPERIMENTS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
_, _, peak = ctx.bench_kernel(6, 0, iters=5, flush_l2=False)
print("int pipe peak Tiop/s", peak)
for k in (5,):
    for v in (0, 1, 17, 3, 19, 21, 100, 101, 117, 103, 119, 121):
        ctx.bench_set(v, k)
        ms, mn, _ = ctx.bench_kernel(0, 23, iters=8, flush_l2=True)
        print(f"k={k} V={v:2d} mean {ms*1e3:8.1f} us  min {mn*1e3:8.1f} us   {3*(1<<23)/ms/1e6:6.2f} Gcompress/s")
