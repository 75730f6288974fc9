"""GPU experiment: latency of one dependent node hash (2 compressions) per SHA variant and warps per CTA
(build with EXTRA=-DZK_SHA_EXPERIMENTS for variants other than the production one)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
variants = [int(v) for v in sys.argv[1:]] or [-1]
for v in variants:
    ctx.bench_set(v, 0)
    row = []
    for warps in (1, 4, 8, 16, 32):
        _, _, cyc = ctx.bench_kernel(7, warps, iters=64, flush_l2=False)
        _, _, cyc = ctx.bench_kernel(7, warps, iters=64, flush_l2=False)
        row.append(f"w{warps}:{cyc:7.0f}")
    print(f"V={v:3d} cycles/node  " + "  ".join(row), flush=True)
