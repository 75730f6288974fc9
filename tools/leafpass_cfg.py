import sys, os
sys.path.insert(0, "/root/repo")
import zkstark_b200 as zk
ctx = zk.Context(0)
ms, mn, _ = ctx.bench_kernel(0, 23, iters=8, flush_l2=True)
print(os.environ.get("ZK_LIB","default"), f"leaf pass 2^23 mean {ms*1e3:8.1f} us min {mn*1e3:8.1f}", flush=True)
