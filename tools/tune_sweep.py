"""GPU experiment: prove() time per (tile shift, leaf-pass log-threads) pair.
usage: tune_sweep.py [shift:logthreads:coop ...]   (0 = library default; coop = widest cooperative level, -1 = default)"""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
combos = [tuple(int(x) for x in v.split(":")) for v in sys.argv[1:]] or [(0, 0, -1), (0, 0, 0), (0, 0, 32), (0, 0, 64), (0, 0, 128)]
ctx.set_stage_timing(True)
for shift, lthr, coop in combos:
    ctx.bench_set_tile_direct_log(17 | (shift << 16) | (lthr << 24))
    if coop >= 0:
        ctx.bench_set(-1, (coop + 1) << 8)
    for k in (20, 16, 10):
        a = zk.trace_fibsq(k)
        best = None
        for it in range(6):
            tr, st = ctx.prove_fibsq_from_trace(k, a)
            s = ctx.last_stage_ms()
            if best is None or s["total"] < best["total"]: best = s
        print(f"shift {shift} log_threads {lthr} coop {coop} log_n {k}", hashlib.sha256(tr).hexdigest()[:16],
              {n: round(v, 3) for n, v in best.items()}, flush=True)
