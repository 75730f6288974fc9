"""Host trace generation per ISA cap (ZK_HOST_SIMD): python tools/trace_isa_time.py"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkstark_b200 as zk
L = zk.lib()
k = 20
def run(n, th, isa):
    if isa: os.environ["ZK_HOST_SIMD"] = isa
    else: os.environ.pop("ZK_HOST_SIMD", None)
    secs = (np.arange(n, dtype=np.uint32) + 5).astype(np.uint32)
    out = np.zeros((n, (1 << k) - 1), dtype=np.uint32)   # touched
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        rc = L.zk_trace_fibsq_batch(k, secs.ctypes.data_as(C.c_void_p), n, th, out.ctypes.data_as(C.c_void_p))
        best = min(best, time.perf_counter() - t0)
    return best * 1e3
for n, th in ((4, 1), (8, 1), (16, 1), (32, 2), (32, 3), (32, 4), (32, 8)):
    print(f"n={n:2d} threads={th}: auto {run(n, th, ''):7.1f} ms   avx2 {run(n, th, 'avx2'):7.1f} ms   scalar {run(n, th, 'scalar'):7.1f} ms", flush=True)
