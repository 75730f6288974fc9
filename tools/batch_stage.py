"""Where a resident batch spends its time: `python tools/batch_stage.py <log_n> <n_proofs>` -- device time of
zk_prove_fibsq_resident for the batch, and the stage split of one sub-batch (zk_ctx_set_stage_timing; stage timing runs
the two lanes serially, so its total is an upper bound)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zkstark_b200 as zk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = zk.Context(0)
n = 1 << k
secrets = (np.arange(nb, dtype=np.uint32) + 3141592).astype(np.uint32)
rows = zk.trace_fibsq_batch(k, secrets)
tr_dev = torch.zeros((nb, n), dtype=torch.int32, device="cuda")
tr_dev[:, : n - 1] = torch.from_numpy(rows.view(np.int32)).cuda()
tl = zk.transcript_len(k)
out = np.empty((nb, tl), dtype=np.uint8)
st = np.empty((nb, 32), dtype=np.uint8)
run = lambda: ctx.prove_fibsq_resident_ptr(k, nb, tr_dev.data_ptr(), out.ctypes.data, st.ctypes.data)
for _ in range(2):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"2^{k} x {nb}: {ms:.2f} ms per batch, {ms / nb * 1e3:.1f} us per proof, {nb / ms * 1e3:.0f} proofs/s", flush=True)
ctx.set_stage_timing(True)
run()
print({k_: round(v, 3) for k_, v in ctx.last_stage_ms().items()})
ctx.set_stage_timing(False)
