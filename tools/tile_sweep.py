"""GPU experiment: whole-tree time by size for each leaf-pass / tile-kernel switch point, and prove() times.
usage: tile_sweep.py [direct_log ...]   (ZK_LIB=<experiment .so> selects a build)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk, hashlib
ctx = zk.Context(0)
directs = [int(v) for v in sys.argv[1:]] or [10, 16, 17, 18]
print("lib", os.path.basename(zk.LIB_PATH))
for direct in directs:
    ctx.bench_set_tile_direct_log(direct)
    row = []
    for ls in (10, 11, 13, 15, 16, 17, 18, 20, 23):
        ms, mn, _ = ctx.bench_kernel(1, ls, iters=6, flush_l2=False)
        row.append(f"2^{ls}:{mn*1e3:7.1f}")
    print(f"direct<=2^{direct:2d} tree us  " + "  ".join(row), flush=True)
ctx.set_stage_timing(True)
for direct in directs:
    ctx.bench_set_tile_direct_log(direct)
    for k in (10, 16, 20):
        a = zk.trace_fibsq(k)
        best = None
        for it in range(5):
            tr, st = ctx.prove_fibsq_from_trace(k, a)
            s = ctx.last_stage_ms()
            if best is None or s["total"] < best["total"]: best = s
        print("direct", direct, "log_n", k, hashlib.sha256(tr).hexdigest()[:16], {n: round(v, 3) for n, v in best.items()}, flush=True)
