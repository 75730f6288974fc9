"""Minimal driver for ncu: prove() of one 2^log_n trace through the C ABI, `reps` times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = zk.Context(0)
a = zk.trace_fibsq(k)
for _ in range(reps):
    tr, st = ctx.prove_fibsq_from_trace(k, a)
print("ok", len(tr), ctx.launch_count())
