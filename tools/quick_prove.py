import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk, hashlib
ctx = zk.Context(0)
ctx.set_stage_timing(True)
for k in (10, 16, 20):
    a = zk.trace_fibsq(k)
    best = None
    for it in range(5):
        t = time.perf_counter(); tr, st = ctx.prove_fibsq_from_trace(k, a); dt = time.perf_counter() - t
        s = ctx.last_stage_ms()
        if best is None or s["total"] < best["total"]: best = s
    print(k, hashlib.sha256(tr).hexdigest()[:16], "wall %.3f ms" % (dt * 1e3), {n: round(v, 3) for n, v in best.items()})
for which, ls, name in [(0, 23, "leaf pass 2^23"), (1, 23, "tree 2^23"), (1, 16, "tree 2^16"), (1, 10, "tree 2^10"), (2, 20, "lde 2^20"), (3, 20, "interp 2^20"), (4, 23, "fold 2^23"), (5, 20, "cp 2^20")]:
    ms, mn, _ = ctx.bench_kernel(which, ls, iters=8, flush_l2=True)
    print(f"{name:16s} mean {ms*1e3:9.1f} us  min {mn*1e3:9.1f} us")
