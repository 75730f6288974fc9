"""CPU ORACLE bindings (TEST INFRASTRUCTURE -- never imported by zkstark_b200/).

ctypes view of oracle/liboracle.so, the plain-C restatement of the reference prover
(/root/reference/src/*.rs; see oracle/oracle.h for the file:line map).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "liboracle.so")

P = 3221225473
ROUTE_LITERAL, ROUTE_FAST = 0, 1


def build(force=False):
    """Compile oracle/*.c -> oracle/liboracle.so with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".c", ".h"))]
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class ProveInfo(C.Structure):
    _fields_ = [
        ("log_n", C.c_uint32), ("a_last", C.c_uint32), ("alpha_raw", C.c_uint32 * 3),
        ("beta_raw", C.c_uint32 * 32), ("free_term", C.c_uint32), ("query_raw", C.c_uint32),
        ("f_root", C.c_uint8 * 32), ("cp_root", C.c_uint8 * 32), ("fri_roots", (C.c_uint8 * 32) * 32),
        ("t_trace", C.c_double), ("t_interp", C.c_double), ("t_lde", C.c_double), ("t_cp", C.c_double),
        ("t_fri", C.c_double), ("t_merkle", C.c_double), ("t_decommit", C.c_double), ("t_total", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    u32, u64, i32, vp = C.c_uint32, C.c_uint64, C.c_int, C.c_void_p
    for name in ("orc_add", "orc_sub", "orc_mul", "orc_pow"):
        getattr(L, name).argtypes = [u32, u32, u32]
        getattr(L, name).restype = u32
    L.orc_inv.argtypes = [u32, u32]; L.orc_inv.restype = u32
    L.orc_from_i32.argtypes = [C.c_int32, u32]; L.orc_from_i32.restype = u32
    L.orc_generator.argtypes = [u32]; L.orc_generator.restype = u32
    L.orc_order.argtypes = [u32, u32]; L.orc_order.restype = u32
    L.orc_sha256.argtypes = [vp, C.c_size_t, vp]; L.orc_sha256.restype = None
    L.orc_sha256_uses_shani.restype = i32
    L.orc_sha256_force_portable.argtypes = [i32]
    L.orc_merkle_build.argtypes = [vp, u64, vp, i32]; L.orc_merkle_build.restype = i32
    L.orc_merkle_trace.argtypes = [vp, u64, u64, vp]; L.orc_merkle_trace.restype = i32
    L.orc_root_from_path.argtypes = [u32, u64, vp, u32, vp]; L.orc_root_from_path.restype = None
    L.orc_poly_solve.argtypes = [vp, u32, u32, u32]; L.orc_poly_solve.restype = u32
    L.orc_lagrange.argtypes = [vp, vp, u32, vp, u32]; L.orc_lagrange.restype = u32
    L.orc_poly_div.argtypes = [vp, u32, vp, u32, vp, vp, vp, vp, u32]; L.orc_poly_div.restype = i32
    L.orc_fri_coeff.argtypes = [vp, u32, u32, vp, u32]; L.orc_fri_coeff.restype = None
    L.orc_trace_fibsq.argtypes = [u32, u32, vp]; L.orc_trace_fibsq.restype = None
    L.orc_interpolate_subgroup.argtypes = [vp, u32, u32, vp, i32]; L.orc_interpolate_subgroup.restype = i32
    L.orc_lde_coset.argtypes = [vp, u32, u32, u32, vp, i32]; L.orc_lde_coset.restype = i32
    L.orc_cp_eval_fibsq.argtypes = [vp, u32, u32, vp, u32, u32, vp, i32]; L.orc_cp_eval_fibsq.restype = i32
    L.orc_fri_fold.argtypes = [vp, u64, u32, u32, vp, i32]; L.orc_fri_fold.restype = i32
    L.orc_prove_fibsq.argtypes = [u32, u32, i32, i32, vp, C.c_size_t, vp, vp, vp]
    L.orc_prove_fibsq.restype = i32
    L.orc_transcript_len.argtypes = [u32]; L.orc_transcript_len.restype = C.c_size_t
    L.orc_verify_fibsq.argtypes = [u32, u32, vp, C.c_size_t, i32, vp]; L.orc_verify_fibsq.restype = i32
    L.orc_air_step.argtypes = [vp, u32, u32]; L.orc_air_step.restype = u32
    L.orc_trace_air.argtypes = [vp, u32, u32, vp]; L.orc_trace_air.restype = None
    L.orc_cp_eval_air.argtypes = [vp, vp, u32, u32, vp, u32, u32, vp, i32]; L.orc_cp_eval_air.restype = i32
    L.orc_prove_air.argtypes = [vp, u32, u32, i32, i32, vp, C.c_size_t, vp, vp, vp]; L.orc_prove_air.restype = i32
    L.orc_verify_air.argtypes = [vp, u32, u32, vp, C.c_size_t, i32, vp]; L.orc_verify_air.restype = i32
    _lib = L
    return L


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def default_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def sha256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_sha256(data, len(data), out)
    return out.raw


def trace_fibsq(log_n, secret=3141592):
    out = np.empty((1 << log_n) - 1, dtype=np.uint32)
    lib().orc_trace_fibsq(log_n, secret, _ptr(out))
    return out


def interpolate_subgroup(y, log_n, threads=1):
    y = _u32(y)
    out = np.empty(1 << log_n, dtype=np.uint32)
    rc = lib().orc_interpolate_subgroup(_ptr(y), len(y), log_n, _ptr(out), threads)
    assert rc == 0, rc
    return out


def lde_coset(coeffs, log_n, log_blowup=3, shift=5, threads=1):
    coeffs = _u32(coeffs)
    out = np.empty(1 << (log_n + log_blowup), dtype=np.uint32)
    rc = lib().orc_lde_coset(_ptr(coeffs), log_n, log_blowup, shift, _ptr(out), threads)
    assert rc == 0, rc
    return out


def cp_eval_fibsq(f_eval, log_n, alpha_raw, a_first, a_last, log_blowup=3, threads=1):
    f_eval = _u32(f_eval)
    al = _u32(alpha_raw)
    out = np.empty_like(f_eval)
    rc = lib().orc_cp_eval_fibsq(_ptr(f_eval), log_n, log_blowup, _ptr(al), a_first, a_last, _ptr(out), threads)
    assert rc == 0, rc
    return out


def fri_fold(cur, layer_k, beta_raw, threads=1):
    cur = _u32(cur)
    out = np.empty(len(cur) // 2, dtype=np.uint32)
    rc = lib().orc_fri_fold(_ptr(cur), len(cur), layer_k, beta_raw, _ptr(out), threads)
    assert rc == 0, rc
    return out


def merkle_build(leaves, threads=1):
    leaves = _u32(leaves)
    heap = np.empty((2 * len(leaves) - 1, 32), dtype=np.uint8)
    rc = lib().orc_merkle_build(_ptr(leaves), len(leaves), _ptr(heap), threads)
    assert rc == 0, rc
    return heap


def merkle_trace(heap, idx):
    size = (heap.shape[0] + 1) // 2
    out = np.empty((64, 32), dtype=np.uint8)
    n = lib().orc_merkle_trace(_ptr(heap), size, idx, _ptr(out))
    return out[:n].copy()


def root_from_path(element, index, path):
    path = np.ascontiguousarray(path, dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    lib().orc_root_from_path(element, index, _ptr(path), path.shape[0], _ptr(out))
    return out


# ---- the AIR family of SURVEY 8(f4): air = (a0, c20, c02, c11, c10, c01, c00), see oracle.h
AIR_FIBSQ = (1, 1, 1, 0, 0, 0, 0)


def _air(air):
    a = (C.c_uint32 * 7)(*[int(v) for v in air])
    return a


def trace_air(air, log_n, secret=3141592):
    out = np.empty((1 << log_n) - 1, dtype=np.uint32)
    lib().orc_trace_air(_air(air), log_n, secret, _ptr(out))
    return out


def cp_eval_air(air, f_eval, log_n, alpha_raw, a_first, a_last, log_blowup=3, threads=1):
    f_eval = _u32(f_eval)
    al = _u32(alpha_raw)
    out = np.empty_like(f_eval)
    rc = lib().orc_cp_eval_air(_air(air), _ptr(f_eval), log_n, log_blowup, _ptr(al), a_first, a_last, _ptr(out), threads)
    assert rc == 0, rc
    return out


def prove_air(air, log_n, secret=3141592, route=ROUTE_FAST, threads=None):
    """-> (transcript bytes, state bytes, ProveInfo)"""
    threads = threads or default_threads()
    cap = lib().orc_transcript_len(log_n)
    buf = C.create_string_buffer(cap)
    state = C.create_string_buffer(32)
    n = C.c_size_t(0)
    info = ProveInfo()
    rc = lib().orc_prove_air(_air(air), log_n, secret, route, threads, buf, cap, C.byref(n), state, C.byref(info))
    if rc != 0:
        raise RuntimeError(f"orc_prove_air failed rc={rc}")
    return buf.raw[: n.value], state.raw, info


def verify_air(air, log_n, a_last, transcript: bytes, rederive=True, expect_state: bytes = None):
    return lib().orc_verify_air(_air(air), log_n, a_last, transcript, len(transcript), int(rederive), expect_state)


def prove_fibsq(log_n, secret=3141592, route=ROUTE_FAST, threads=None):
    """-> (transcript bytes, state bytes, ProveInfo)"""
    threads = threads or default_threads()
    cap = lib().orc_transcript_len(log_n)
    buf = C.create_string_buffer(cap)
    state = C.create_string_buffer(32)
    n = C.c_size_t(0)
    info = ProveInfo()
    rc = lib().orc_prove_fibsq(log_n, secret, route, threads, buf, cap, C.byref(n), state, C.byref(info))
    if rc != 0:
        raise RuntimeError(f"orc_prove_fibsq failed rc={rc}")
    return buf.raw[: n.value], state.raw, info


def verify_fibsq(log_n, a_last, transcript: bytes, rederive=True, expect_state: bytes = None):
    return lib().orc_verify_fibsq(log_n, a_last, transcript, len(transcript), int(rederive), expect_state)
