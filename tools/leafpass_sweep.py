"""GPU experiment: leaf-pass time by levels per thread (k) for the library named by ZK_LIB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
print("lib", os.path.basename(zk.LIB_PATH))
for ls in (23, 21):
    for k in (4, 5, 6):
        ctx.bench_set(-1, k)
        ms, mn, _ = ctx.bench_kernel(0, ls, iters=6, flush_l2=True)
        print(f"leaves 2^{ls} k={k} mean {ms*1e3:8.1f} us min {mn*1e3:8.1f}", flush=True)
