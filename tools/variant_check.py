"""Leaf-pass time and golden parity of the library named by ZK_LIB (experiment builds, tools/build_variant.sh)."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "transcripts.json")))
ok = all(hashlib.sha256(ctx.prove_fibsq(k, 3141592)[0]).hexdigest() == gold[str(k)]["sha256"] for k in (10, 16))
ms, mn, _ = ctx.bench_kernel(0, 23, iters=10, flush_l2=True)
tr = zk.trace_fibsq(20)
import time
ctx.prove_fibsq_from_trace(20, tr)
ctx.set_stage_timing(True)
ctx.prove_fibsq_from_trace(20, tr)
st = ctx.last_stage_ms()
print(f"{os.environ.get('ZK_LIB', 'production'):45s} parity {ok}  leaf pass 2^23 mean {ms * 1e3:7.1f} us min {mn * 1e3:7.1f}  prove 2^20 {st['total']:.3f} ms", flush=True)
