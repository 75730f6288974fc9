"""Out-of-bounds evidence without compute-sanitizer (it is closed on this GPU pool: profiles/r2_sanitizer_unavailable.txt).

Every device-pointer entry point writes its output into the MIDDLE of a sentinel-filled allocation: the red zones on
both sides must come back untouched, read-only inputs must come back unchanged, and the payload must equal the host-buffer
call (itself checked against the oracle in test_gpu_parity).  Sizes cover single-pass, two-pass and three-pass transforms
and every tree regime.  tools/sanitize_workload.py (the workload prepared for the sanitizers) runs as a plain test."""
import os
import subprocess
import sys

import numpy as np
import pytest

import zkstark_b200 as zk

pytestmark = pytest.mark.gpu
P = 3221225473
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RED = 4096          # sentinel words on each side
SENT = -1412567295  # 0xABCDEF01 as int32


@pytest.fixture(scope="module")
def tctx():
    import torch
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        c = zk.Context(0, stream=s.cuda_stream)
        yield c, s
        c.close()


class Guarded:
    """`words` int32 words of device memory between two red zones."""

    def __init__(self, torch, words, fill=None):
        self.t = torch.full((words + 2 * RED,), SENT, dtype=torch.int32, device="cuda")
        self.words = words
        if fill is not None:
            self.t[RED: RED + words] = torch.from_numpy(np.ascontiguousarray(fill).view(np.int32)).cuda()

    @property
    def ptr(self):
        return self.t.data_ptr() + 4 * RED

    def payload(self):
        return self.t[RED: RED + self.words].cpu().numpy().view(np.uint32)

    def intact(self):
        return bool((self.t[:RED] == SENT).all().item() and (self.t[RED + self.words:] == SENT).all().item())


def rnd(n, seed):
    return np.random.default_rng(seed).integers(0, P, size=n, dtype=np.uint64).astype(np.uint32)


@pytest.mark.parametrize("log_n", [3, 5, 9, 10, 11, 12, 14, 16, 20])
def test_transform_redzones(tctx, log_n):
    import torch
    ctx, s = tctx
    n = 1 << log_n
    with torch.cuda.stream(s):
        y = rnd(n, log_n)
        y[n - 1] = 0
        gy = Guarded(torch, n, y)
        gc = Guarded(torch, n)
        ctx.interpolate_subgroup_dev(gy.ptr, n - 1, log_n, gc.ptr)
        ctx.synchronize()
        assert gy.intact() and gc.intact()
        coeffs = gc.payload().copy()
        assert np.array_equal(coeffs, ctx.interpolate_subgroup(y[: n - 1], log_n))
        ge = Guarded(torch, 8 * n)
        gci = Guarded(torch, n, coeffs)
        ctx.lde_coset_dev(gci.ptr, log_n, ge.ptr)
        ctx.synchronize()
        assert ge.intact() and gci.intact()
        assert np.array_equal(gci.payload(), coeffs)                 # read-only input
        evals = ge.payload().copy()
        assert np.array_equal(evals, ctx.lde_coset(coeffs, log_n))
        gcp = Guarded(torch, 8 * n)
        al = [361545003, 3235878091, 2708123352]
        ctx.cp_eval_fibsq_dev(ge.ptr, log_n, al, 1, 12345, gcp.ptr)
        ctx.synchronize()
        assert gcp.intact() and ge.intact() and np.array_equal(ge.payload(), evals)
        assert np.array_equal(gcp.payload(), ctx.cp_eval_fibsq(evals, log_n, al, 1, 12345))
        gf = Guarded(torch, 4 * n)
        ctx.fri_fold_dev(gcp.ptr, 8 * n, 0, 0xDEADBEEF, gf.ptr)
        ctx.synchronize()
        assert gf.intact() and gcp.intact()
        assert np.array_equal(gf.payload(), ctx.fri_fold(gcp.payload(), 0, 0xDEADBEEF))


@pytest.mark.parametrize("log_size", [0, 1, 5, 8, 11, 12, 15, 17, 18, 21])
def test_merkle_redzones(tctx, log_size):
    import torch
    ctx, s = tctx
    size = 1 << log_size
    with torch.cuda.stream(s):
        leaves = rnd(size, 50 + log_size)
        gl = Guarded(torch, size, leaves)
        gh = Guarded(torch, 8 * (2 * size - 1))
        ctx.merkle_build_dev(gl.ptr, size, gh.ptr)
        ctx.synchronize()
        assert gh.intact() and gl.intact() and np.array_equal(gl.payload(), leaves)
        heap = gh.payload().view(np.uint8).reshape(2 * size - 1, 32)
        if log_size <= 18:
            assert np.array_equal(heap, ctx.merkle_build(leaves))
        else:
            assert heap[0].tobytes() == ctx.merkle_root(leaves)


def test_sanitizer_workload_runs_clean():
    """the workload prepared for compute-sanitizer (every kernel family at small sizes, each checked against the oracle)"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_workload.py")], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "sanitize workload ok" in out.stdout
