"""GPU experiment: leaf-pass time per SHA-256 pipe-balancing variant (build with EXTRA=-DZK_SHA_EXPERIMENTS).
usage: sha_variants.py [variant ...]   (100+v = shared leaf/node data-block code)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
_, _, peak = ctx.bench_kernel(6, 0, iters=5, flush_l2=False)
print("int pipe peak Tiop/s", peak)
variants = [int(v) for v in sys.argv[1:]] or [3, 103]
for k in (5,):
    for v in variants:
        ctx.bench_set(v, k)
        ms, mn, _ = ctx.bench_kernel(0, 23, iters=8, flush_l2=True)
        print(f"k={k} V={v:3d} mean {ms*1e3:8.1f} us  min {mn*1e3:8.1f} us   {3*(1<<23)/ms/1e6:6.2f} Gcompress/s", flush=True)
