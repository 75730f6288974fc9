"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/zkstark_b200.h
declares, and its host-only entry points agree with the oracle.  No compute call needs a GPU here."""
import ctypes as C
import os

import numpy as np
import pytest

import zkstark_b200 as zk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_built_in_tree():
    assert os.path.exists(zk.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(zk.LIB_PATH).endswith("zkstark_b200")


def test_header_symbols_all_exported():
    declared = zk.header_symbols()
    exported = set(zk.exported_symbols())
    assert len(declared) >= 25
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    L = zk.lib()
    for s in declared:
        assert hasattr(L, s)


def test_version_string():
    assert b"sm_100a" in zk.lib().zk_version()


def test_transcript_len_matches_oracle(orc):
    assert zk.transcript_len(10) == 7836  # Proof::size() 7884 - 48 (proof.rs:151-154)
    for k in range(3, 28):
        assert zk.transcript_len(k) == orc.lib().orc_transcript_len(k)


def test_trace_matches_oracle_and_kat(orc, kats):
    a = zk.trace_fibsq(10, 3141592)
    assert int(a[1022]) == kats["prover"]["a_1022"]  # prover.rs:42
    for k, secret in [(3, 1), (4, 0), (7, 3141592), (12, 0xFFFFFFFF), (14, 3221225473 + 5)]:
        assert np.array_equal(zk.trace_fibsq(k, secret), orc.trace_fibsq(k, secret))


def test_merkle_trace_host_matches_reference_kat(orc, kats):
    k = kats["merkle_test"]
    heap = np.array([list(bytes.fromhex(h)) for h in k["heap"]], dtype=np.uint8)
    for leaf, sibs in k["traces"].items():
        got = zk.merkle_trace_host(heap, int(leaf))
        assert [bytes(g).hex() for g in got] == [k["heap"][s] for s in sibs]
    big = orc.merkle_build(np.arange(64, dtype=np.uint32))
    for i in (0, 1, 31, 63):
        assert np.array_equal(zk.merkle_trace_host(big, i), orc.merkle_trace(big, i))


def test_no_cpu_fallback_without_device():
    """On a box without a GPU, creating a context must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zk.ZkError):
        zk.Context(0)


def test_invalid_arguments_do_not_crash():
    L = zk.lib()
    assert L.zk_trace_fibsq(0, 1, None) < 0
    assert L.zk_merkle_trace(None, 4, 0, None) < 0
    assert L.zk_ctx_synchronize(None) < 0
    assert L.zk_ctx_launch_count(None) == 0
    L.zk_ctx_destroy(None)


def test_product_does_not_import_oracle():
    """The shipped package must never route through oracle/ (tier rule 3)."""
    root = os.path.dirname(zk.LIB_PATH)
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "orc_" not in txt, f


def test_host_sha256_matches_hashlib():
    """The host SHA-256 of the channel tail / verifier (SHA-NI when the CPU has it, portable otherwise) against hashlib."""
    import ctypes
    import hashlib
    import random
    import subprocess
    sym = [l.split()[-1] for l in subprocess.check_output(["nm", "-D", "--defined-only", zk.LIB_PATH], text=True).splitlines()
           if "host_sha256" in l]
    assert sym, "host_sha256 not found in the library"
    f = getattr(zk.lib(), sym[0])
    f.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    f.restype = None
    rnd = random.Random(7)
    for n in list(range(0, 140)) + [255, 256, 257, 780, 1000, 4097]:
        d = bytes(rnd.getrandbits(8) for _ in range(n))
        out = ctypes.create_string_buffer(32)
        f(d, n, out)
        assert out.raw == hashlib.sha256(d).digest(), n


def test_rust_ffi_matches_header():
    """bindings/rust/src/ffi.rs declares every export of the header, and is what tools/gen_rust_ffi.py generates."""
    import re
    import subprocess
    import sys
    ffi = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (zk_[a-z0-9_]+)\(", ffi))
    assert declared == set(zk.header_symbols())
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]) == 0
    gpu = open(os.path.join(ROOT, "bindings", "rust", "src", "gpu.rs")).read()
    for name in set(re.findall(r"ffi::(zk_[a-z0-9_]+)\(", gpu)):
        assert name in declared, name


def test_rust_patches_apply(tmp_path):
    """The level-2 patches apply cleanly to the reference tree (only where the reference is present: not on the GPU box)."""
    import shutil
    import subprocess
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "src")) or shutil.which("patch") is None:
        pytest.skip("no reference checkout / patch(1) here")
    work = tmp_path / "ref"
    shutil.copytree(ref, work, ignore=shutil.ignore_patterns(".git"))
    for p in ("01_hooks.patch", "02_prover_eval_loops.patch"):
        r = subprocess.run(["patch", "-p1", "-d", str(work), "-i", os.path.join(ROOT, "bindings", "rust", "patches", p)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    src = (work / "src" / "polynomial.rs").read_text()
    assert "crate::gpu::lagrange_hook(points)" in src and "pub fn lagrange<T>(points: &Vec<(T, T)>) -> Polynomial<T>" in src
    assert "pub fn new(mut size: usize, data: impl Iterator<Item = u32>) -> Self" in (work / "src" / "merkle.rs").read_text()


@pytest.mark.parametrize("simd", ["", "avx2", "scalar"])
def test_trace_batch_simd_matches_scalar(orc, simd, monkeypatch):
    """zk_trace_fibsq_batch (4 or 8 traces per SIMD group -- AVX-512 IFMA, AVX2 or scalar code, ZK_HOST_SIMD caps what the CPU
    offers -- over several threads) == zk_trace_fibsq == the oracle's trace, incl. a batch that is not a multiple of 8 and
    secrets 0, 1, p - 1, >= p."""
    if simd:
        monkeypatch.setenv("ZK_HOST_SIMD", simd)
    else:
        monkeypatch.delenv("ZK_HOST_SIMD", raising=False)
    P = 3221225473
    secrets = np.array([3141592, 0, 1, P - 1, P, P + 7, 0xFFFFFFFF, 12345, 99, 2**31, 2**31 + 1, 7, 8, 9, 10, 11, 12, 13, 14], dtype=np.uint32)
    for log_n in (1, 2, 3, 10, 13):
        for threads in (1, 3):
            got = zk.trace_fibsq_batch(log_n, secrets, threads)
            for j, s in enumerate(secrets.tolist()):
                assert np.array_equal(got[j], zk.trace_fibsq(log_n, s)), (log_n, j)
                assert np.array_equal(got[j], orc.trace_fibsq(log_n, s)), (log_n, j)
    assert int(zk.trace_fibsq_batch(10, secrets[:1])[0][1022]) == 2338775057      # prover.rs:42
    import time
    t0 = time.perf_counter()
    zk.trace_fibsq_batch(18, np.arange(8, dtype=np.uint32) + 5, 1)
    t8 = time.perf_counter() - t0
    t0 = time.perf_counter()
    zk.trace_fibsq(18, 5)
    t1 = time.perf_counter() - t0
    print(f"8 traces SIMD {t8 * 1e3:.2f} ms vs 1 scalar {t1 * 1e3:.2f} ms")
