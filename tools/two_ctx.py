"""GPU experiment: aggregate proofs/s of T host threads, each with its own context (own stream) and batch B."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkstark_b200 as zk
k = 20
for T, B in [(1, 32), (2, 16), (4, 8), (1, 64), (2, 32)]:
    ctxs, bufs = [], []
    for t in range(T):
        s = torch.cuda.Stream()
        c = zk.Context(0, stream=s.cuda_stream)
        tr = torch.zeros((B, 1 << k), dtype=torch.int32, device="cuda")
        tr[:, : (1 << k) - 1] = torch.from_numpy(zk.trace_fibsq(k, 3141592 + t).view(np.int32)).cuda()
        out = torch.empty((B, zk.transcript_len(k)), dtype=torch.uint8).pin_memory()
        st = torch.empty((B, 32), dtype=torch.uint8).pin_memory()
        ctxs.append(c); bufs.append((tr, out, st, s))
    def work(i, reps):
        c = ctxs[i]; tr, out, st, s = bufs[i]
        for _ in range(reps):
            c.prove_fibsq_resident_ptr(k, B, tr.data_ptr(), out.data_ptr(), st.data_ptr())
    for i in range(T): work(i, 1)
    torch.cuda.synchronize()
    reps = 6
    th = [threading.Thread(target=work, args=(i, reps)) for i in range(T)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"threads {T} batch {B}: {T*B*reps/dt:8.1f} proofs/s", flush=True)
    for c in ctxs: c.close()
    del bufs, ctxs
    torch.cuda.empty_cache()
