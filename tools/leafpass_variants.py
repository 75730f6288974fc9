"""GPU experiment: leaf-pass time per SHA variant (ZK_LIB must be a -DZK_SHA_EXPERIMENTS build).
variant ids: v (UNI 0), 100+v (UNI 1), 200+v (UNI 2, production structure)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
for v in [int(x) for x in sys.argv[1:]] or [-1, 203, 235, 233, 219]:
    ctx.bench_set(v, 0)
    ms, mn, _ = ctx.bench_kernel(0, 23, iters=8, flush_l2=True)
    print(f"variant {v:4d}: leaf pass 2^23 mean {ms*1e3:8.1f} us min {mn*1e3:8.1f}", flush=True)
