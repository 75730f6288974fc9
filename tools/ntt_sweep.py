"""GPU experiment: LDE / interpolation time by size for the library named by ZK_LIB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
row = [os.path.basename(zk.LIB_PATH)]
for k in ([int(a) for a in sys.argv[1:]] or [16, 18, 20, 22, 24]):
    _, lde, _ = ctx.bench_kernel(2, k, iters=6, flush_l2=(k >= 22))
    _, itp, _ = ctx.bench_kernel(3, k, iters=6, flush_l2=(k >= 22))
    row.append(f"2^{k}: lde {lde*1e3:8.1f} interp {itp*1e3:7.1f} us")
print("  ".join(row), flush=True)
