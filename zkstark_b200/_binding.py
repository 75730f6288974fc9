"""ctypes binding of include/zkstark_b200.h (the C ABI the Rust shim binds)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# ZK_LIB: an experiment build of the same library (tools/ only); the product path is the in-tree .so
LIB_PATH = os.environ.get("ZK_LIB") or os.path.join(_PKG, "libzkstark_b200.so")
HEADER = os.path.join(_ROOT, "include", "zkstark_b200.h")

ZK_OK = 0


class ZkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"zkstark_b200 error {code}: {msg}")
        self.code = code


def build_library(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> zkstark_b200/libzkstark_b200.so"""
    csrc = os.path.join(_PKG, "csrc")
    args = ["make", "-C", csrc, "-j4"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZkError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is the only implementation)")
    L = C.CDLL(LIB_PATH)
    u32, u64, i32, vp, sz = C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t
    sig = {
        "zk_ctx_create": ([i32, C.POINTER(vp)], i32),
        "zk_ctx_create_on_stream": ([i32, vp, C.POINTER(vp)], i32),
        "zk_ctx_destroy": ([vp], None),
        "zk_last_error": ([vp], C.c_char_p),
        "zk_version": ([], C.c_char_p),
        "zk_ctx_synchronize": ([vp], i32),
        "zk_ctx_stream": ([vp], vp),
        "zk_ctx_launch_count": ([vp], u64),
        "zk_ctx_last_stage_ms": ([vp, vp], i32),
        "zk_ctx_set_stage_timing": ([vp, i32], i32),
        "zk_ctx_set_memory_budget": ([vp, sz], i32),
        "zk_ctx_trim": ([vp, i32], i32),
        "zk_interpolate_subgroup": ([vp, vp, u32, u32, vp], i32),
        "zk_lde_coset": ([vp, vp, u32, u32, u32, vp], i32),
        "zk_fri_fold": ([vp, vp, u64, u32, u32, vp], i32),
        "zk_cp_eval_fibsq": ([vp, vp, u32, u32, vp, u32, u32, vp], i32),
        "zk_merkle_build": ([vp, vp, u64, vp], i32),
        "zk_merkle_root": ([vp, vp, u64, vp], i32),
        "zk_merkle_trace": ([vp, u64, u64, vp], i32),
        "zk_transcript_len": ([u32], sz),
        "zk_prove_fibsq": ([vp, u32, u32, vp, sz, C.POINTER(sz), vp], i32),
        "zk_prove_fibsq_from_trace": ([vp, u32, vp, vp, sz, C.POINTER(sz), vp], i32),
        "zk_prove_fibsq_batch": ([vp, u32, u32, vp, vp, vp], i32),
        "zk_trace_fibsq": ([u32, u32, vp], i32),
        "zk_trace_fibsq_batch": ([u32, vp, u32, u32, vp], i32),
        "zk_prove_fibsq_batch_secrets": ([vp, u32, u32, vp, vp, vp], i32),
        "zk_transcript_len_ex": ([u32, u32], sz),
        "zk_prove_fibsq_ex": ([vp, u32, u32, vp, u32, vp, sz, C.POINTER(sz), vp], i32),
        "zk_verify_fibsq_ex": ([u32, u32, vp, sz, vp, vp, u32, i32, C.POINTER(i32)], i32),
        "zk_ctx_set_air": ([vp, vp], i32),
        "zk_ctx_get_air": ([vp, vp], i32),
        "zk_trace_air": ([vp, u32, u32, vp], i32),
        "zk_trace_air_batch": ([vp, u32, vp, u32, u32, vp], i32),
        "zk_verify_air": ([vp, u32, u32, vp, sz, vp, vp, u32, i32, C.POINTER(i32)], i32),
        "zk_prove_stream_submit": ([vp, u32, u32, vp], i32),
        "zk_prove_stream_collect": ([vp, vp, vp], i32),
        "zk_sha256": ([vp, sz, vp], None),
        "zk_field_ops": ([vp, vp, vp, sz, vp], i32),
        "zk_proof_serialized_len": ([sz], sz),
        "zk_proof_serialize": ([vp, vp, sz, vp, sz, C.POINTER(sz)], i32),
        "zk_proof_deserialize": ([vp, sz, vp, C.POINTER(vp), C.POINTER(sz)], i32),
        "zk_verify_fibsq": ([u32, u32, vp, sz, vp, i32, C.POINTER(i32)], i32),
        "zk_interpolate_subgroup_dev": ([vp, vp, u32, u32, vp], i32),
        "zk_lde_coset_dev": ([vp, vp, u32, u32, u32, vp], i32),
        "zk_fri_fold_dev": ([vp, vp, u64, u32, u32, vp], i32),
        "zk_cp_eval_fibsq_dev": ([vp, vp, u32, u32, vp, u32, u32, vp], i32),
        "zk_merkle_build_dev": ([vp, vp, u64, vp], i32),
        "zk_prove_fibsq_resident": ([vp, u32, u32, vp, vp, vp], i32),
        "zk_bench_kernel": ([vp, i32, u32, u32, i32, vp], i32),
        "zk_shard_state_init_dev": ([vp, vp, u32, u32], i32),
        "zk_shard_lde_coset_dev": ([vp, vp, u32, u32, vp, vp], i32),
        "zk_shard_merkle_blocks_dev": ([vp, vp, u64, u32, vp], i32),
        "zk_shard_upper_tree_dev": ([vp, vp, u32, i32, u32, u32, vp, vp], i32),
        "zk_shard_cp_dev": ([vp, vp, vp, u64, u32, u32, u32, u32, vp, vp], i32),
        "zk_shard_fold_dev": ([vp, vp, u64, u32, u32, u32, u32, u32, vp, vp], i32),
        "zk_shard_fri_tail_dev": ([vp, u32, u32, vp, vp, vp, vp], i32),
        "zk_decommit_len": ([u32], C.c_size_t),
        "zk_open_queries": ([vp, u32, u32, vp, u32, vp, sz], i32),
        "zk_shard_fold_blocks_dev": ([vp, vp, u64, u32, u32, u32, u32, u32, vp, vp, vp], i32),
        "zk_shard_cp_blocks_dev": ([vp, vp, vp, u64, u32, u32, u32, u32, vp, vp, vp], i32),
        "zk_shard_decommit_dev": ([vp, u32, u32, u32, u32, u32, vp, vp, vp, vp, vp], i32),
        "zk_shard_unique_id": ([vp], i32),
        "zk_shard_comm_init": ([vp, vp, u32, u32, C.POINTER(vp)], i32),
        "zk_shard_comm_init_local": ([vp, u32, C.POINTER(vp)], i32),
        "zk_shard_comm_destroy": ([vp], None),
        "zk_shard_comm_set": ([vp, u32, u32], i32),
        "zk_prove_fibsq_sharded": ([vp, u32, vp, vp, sz, C.POINTER(sz), vp], i32),
        "zk_shard_last_phase_ms": ([vp, vp], i32),
        "zk_shard_last_comm_bytes": ([vp], u64),
        "zk_shard_last_first_tail": ([vp], u32),
        "zk_bench_set": ([vp, i32, u32], i32),
        "zk_bench_set_tile_direct_log": ([vp, u32], i32),
    }
    for name, (argt, rest) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argt
        fn.restype = rest
    _lib = L
    return L


def header_symbols():
    """Every function include/zkstark_b200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", txt)))


def exported_symbols():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH], text=True)
    return sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("zk_"))


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def transcript_len(log_n):
    return lib().zk_transcript_len(log_n)


def trace_fibsq(log_n, secret=3141592):
    """prover.rs:32-39 (host, sequential)."""
    out = np.empty((1 << log_n) - 1, dtype=np.uint32)
    rc = lib().zk_trace_fibsq(log_n, secret, _p(out))
    if rc:
        raise ZkError(rc, "zk_trace_fibsq")
    return out


def trace_fibsq_batch(log_n, secrets, threads=0):
    """prover.rs:32-39 for many secrets at once (host, SIMD lanes x threads): rows of 2^log_n - 1 values."""
    secrets = _u32(secrets)
    out = np.empty((len(secrets), (1 << log_n) - 1), dtype=np.uint32)
    rc = lib().zk_trace_fibsq_batch(log_n, _p(secrets), len(secrets), threads, _p(out))
    if rc:
        raise ZkError(rc, "zk_trace_fibsq_batch")
    return out


# ---- the program as data (include/zkstark_b200.h zk_air)
AIR_FIBSQ = (1, 1, 1, 0, 0, 0, 0)


def _air(air, no_specialise=False):
    """(a0, c20, c02, c11, c10, c01, c00) -> zk_air (None: null pointer = the reference's FibonacciSq)"""
    if air is None:
        return None
    air = tuple(int(v) for v in air)
    if len(air) != 7:
        raise ValueError("an AIR is (a0, c20, c02, c11, c10, c01, c00)")
    return (C.c_uint32 * 8)(*air, 1 if no_specialise else 0)


def trace_air(air, log_n, secret=3141592):
    """The trace of the program `air` (host): a[0] = a0, a[1] = secret, a[i+2] = step(a[i], a[i+1])."""
    out = np.empty((1 << log_n) - 1, dtype=np.uint32)
    rc = lib().zk_trace_air(_air(air), log_n, secret, _p(out))
    if rc:
        raise ZkError(rc, "zk_trace_air")
    return out


def trace_air_batch(air, log_n, secrets, threads=0, no_specialise=False):
    secrets = _u32(secrets)
    out = np.empty((len(secrets), (1 << log_n) - 1), dtype=np.uint32)
    rc = lib().zk_trace_air_batch(_air(air, no_specialise), log_n, _p(secrets), len(secrets), threads, _p(out))
    if rc:
        raise ZkError(rc, "zk_trace_air_batch")
    return out


def verify_air(air, log_n, a_last, transcript, state=None, init_state=None, n_extra_queries=0, rederive=True):
    """zk_verify_air: proof.rs:15-149 for a proof of the program `air`.  -> (ok, failed_check)"""
    fc = C.c_int(0)
    rc = lib().zk_verify_air(_air(air), log_n, a_last, transcript, len(transcript), state, init_state, n_extra_queries,
                             int(rederive), C.byref(fc))
    if rc not in (0, -6):
        raise ZkError(rc, "zk_verify_air: invalid argument")
    return rc == 0, fc.value


def verify_fibsq(log_n, a_last, transcript, state=None, rederive=True):
    """proof.rs:15-149 (host).  Returns (ok, failed_check) where failed_check is the proof.rs line of the assert."""
    fc = C.c_int(0)
    rc = lib().zk_verify_fibsq(log_n, a_last, transcript, len(transcript), state, int(rederive), C.byref(fc))
    if rc not in (0, -6):
        raise ZkError(rc, "zk_verify_fibsq: invalid argument")
    return rc == 0, fc.value


def proof_serialize(state, transcript):
    """`Proof { state, data }` -> bytes (bincode 1.x struct layout; zk_proof_serialize)."""
    cap = lib().zk_proof_serialized_len(len(transcript))
    out = (C.c_uint8 * cap)()
    n = C.c_size_t(0)
    rc = lib().zk_proof_serialize(bytes(state), bytes(transcript), len(transcript), out, cap, C.byref(n))
    if rc:
        raise ZkError(rc, "zk_proof_serialize")
    return bytes(out[: n.value])


def proof_deserialize(blob):
    """bytes -> (state, transcript); raises ZkError on a truncated blob or a wrong length prefix."""
    blob = bytes(blob)
    buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob or b"\0")
    st = (C.c_uint8 * 32)()
    ptr = C.c_void_p()
    n = C.c_size_t(0)
    rc = lib().zk_proof_deserialize(buf, len(blob), st, C.byref(ptr), C.byref(n))
    if rc:
        raise ZkError(rc, "truncated proof" if rc == -4 else "proof length prefix does not match the payload")
    off = ptr.value - C.addressof(buf)
    return bytes(st), blob[off: off + n.value]


def verify_fibsq_ex(log_n, a_last, transcript, state=None, init_state=None, n_extra_queries=0, rederive=True):
    """zk_verify_fibsq_ex: a proof with a seeded channel and channel-drawn further test points.  -> (ok, failed_check)"""
    fc = C.c_int(0)
    rc = lib().zk_verify_fibsq_ex(log_n, a_last, transcript, len(transcript), state, init_state, n_extra_queries, int(rederive),
                                  C.byref(fc))
    if rc not in (0, -6):
        raise ZkError(rc, "zk_verify_fibsq_ex: invalid argument")
    return rc == 0, fc.value


def merkle_trace_host(heap, index):
    """merkle.rs:54-71 on a host heap array of shape (2*size-1, 32)."""
    heap = np.ascontiguousarray(heap, dtype=np.uint8)
    size = (heap.shape[0] + 1) // 2
    out = np.empty((64, 32), dtype=np.uint8)
    n = lib().zk_merkle_trace(_p(heap), size, index, _p(out))
    if n < 0:
        raise ZkError(n, "zk_merkle_trace")
    return out[:n].copy()


SHARD_ID_BYTES = 128
SHARD_PHASES = ["upload+gather+interpolate", "lde_cosets", "all_to_all", "unpack", "f_tree", "halo", "cp_tree", "fri_sharded",
                "fri_tail", "decommit+allreduce", "download", "total"]


def shard_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls this and hands the 128 bytes to the other ranks)."""
    buf = (C.c_uint8 * SHARD_ID_BYTES)()
    rc = lib().zk_shard_unique_id(buf)
    if rc:
        raise ZkError(rc, "zk_shard_unique_id: NCCL (libnccl.so.2) is not available")
    return bytes(buf)


class ShardComm:
    """zk_shard_comm: the group of ranks that prove ONE FibonacciSq trace together (BASELINE configs[4]).
    ShardComm(ctx, n_ranks, rank, unique_id) -- one process per GPU over NCCL;
    ShardComm.local(ctx, n_virtual_ranks)    -- the same driver with G virtual ranks on the context's device."""

    def __init__(self, ctx, n_ranks, rank, unique_id):
        self._ctx = ctx
        self._h = C.c_void_p()
        idb = (C.c_uint8 * SHARD_ID_BYTES)(*unique_id)
        ctx._ck(lib().zk_shard_comm_init(ctx._h, idb, n_ranks, rank, C.byref(self._h)))
        self.n_ranks, self.rank = n_ranks, rank

    @classmethod
    def local(cls, ctx, n_virtual_ranks):
        self = cls.__new__(cls)
        self._ctx = ctx
        self._h = C.c_void_p()
        ctx._ck(lib().zk_shard_comm_init_local(ctx._h, n_virtual_ranks, C.byref(self._h)))
        self.n_ranks, self.rank = n_virtual_ranks, 0
        return self

    def close(self):
        if self._h:
            lib().zk_shard_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, log_block=0, log_tail_len=0xFFFFFFFF):
        self._ctx._ck(lib().zk_shard_comm_set(self._h, log_block, log_tail_len))

    def prove_ptr(self, log_n, trace_host_ptr, transcript_ptr, cap, state_ptr):
        n = C.c_size_t(0)
        self._ctx._ck(lib().zk_prove_fibsq_sharded(self._h, log_n, C.c_void_p(trace_host_ptr), C.c_void_p(transcript_ptr), cap,
                                                   C.byref(n), C.c_void_p(state_ptr)))
        return n.value

    def prove(self, log_n, trace):
        """trace: 2^log_n - 1 host values, identical on every rank -> (transcript bytes, state bytes)."""
        trace = _u32(trace)
        assert len(trace) == (1 << log_n) - 1
        cap = transcript_len(log_n)
        tr = np.empty(cap, dtype=np.uint8)
        st = np.empty(32, dtype=np.uint8)
        n = self.prove_ptr(log_n, trace.ctypes.data, tr.ctypes.data, cap, st.ctypes.data)
        return tr[:n].tobytes(), st.tobytes()

    def last_phase_ms(self):
        arr = (C.c_float * 12)()
        self._ctx._ck(lib().zk_shard_last_phase_ms(self._h, arr))
        return dict(zip(SHARD_PHASES, [float(x) for x in arr]))

    def last_comm_bytes(self):
        return int(lib().zk_shard_last_comm_bytes(self._h))

    def last_first_tail(self):
        return int(lib().zk_shard_last_first_tail(self._h))


class Context:
    """zk_ctx: one per (thread, GPU).  `stream` is a raw cudaStream_t (e.g. torch's current stream)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        L = lib()
        rc = (L.zk_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(self._h)) if stream
              else L.zk_ctx_create(device, C.byref(self._h)))
        if rc != ZK_OK:
            self._h = C.c_void_p()
            raise ZkError(rc, "zk_ctx_create failed: no usable CUDA device (there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().zk_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != ZK_OK:
            raise ZkError(rc, lib().zk_last_error(self._h).decode())

    # ---- host-buffer entry points
    def interpolate_subgroup(self, y, log_n):
        y = _u32(y)
        out = np.empty(1 << log_n, dtype=np.uint32)
        self._ck(lib().zk_interpolate_subgroup(self._h, _p(y), len(y), log_n, _p(out)))
        return out

    def lde_coset(self, coeffs, log_n, log_blowup=3, shift=5):
        coeffs = _u32(coeffs)
        assert len(coeffs) == 1 << log_n
        out = np.empty(1 << (log_n + log_blowup), dtype=np.uint32)
        self._ck(lib().zk_lde_coset(self._h, _p(coeffs), log_n, log_blowup, shift, _p(out)))
        return out

    def fri_fold(self, cur, layer_k, beta_raw):
        cur = _u32(cur)
        out = np.empty(len(cur) // 2, dtype=np.uint32)
        self._ck(lib().zk_fri_fold(self._h, _p(cur), len(cur), layer_k, beta_raw, _p(out)))
        return out

    def cp_eval_fibsq(self, f_eval, log_n, alpha_raw, a_first, a_last, log_blowup=3):
        f_eval = _u32(f_eval)
        al = _u32(alpha_raw)
        out = np.empty_like(f_eval)
        self._ck(lib().zk_cp_eval_fibsq(self._h, _p(f_eval), log_n, log_blowup, _p(al), a_first, a_last, _p(out)))
        return out

    def field_ops(self, a, b):
        """K1 on the device: rows a+b, a-b, a*b, -a, a^-1, a^b of canonical residues (zk_field_ops)."""
        a, b = _u32(a), _u32(b)
        assert a.shape == b.shape and a.ndim == 1
        out = np.empty((6, len(a)), dtype=np.uint32)
        self._ck(lib().zk_field_ops(self._h, _p(a), _p(b), len(a), _p(out)))
        return out

    def merkle_build(self, leaves):
        leaves = _u32(leaves)
        heap = np.empty((2 * len(leaves) - 1, 32), dtype=np.uint8)
        self._ck(lib().zk_merkle_build(self._h, _p(leaves), len(leaves), _p(heap)))
        return heap

    def merkle_root(self, leaves):
        leaves = _u32(leaves)
        root = np.empty(32, dtype=np.uint8)
        self._ck(lib().zk_merkle_root(self._h, _p(leaves), len(leaves), _p(root)))
        return root.tobytes()

    def prove_fibsq(self, log_n, secret=3141592):
        cap = transcript_len(log_n)
        tr = np.empty(cap, dtype=np.uint8)
        st = np.empty(32, dtype=np.uint8)
        n = C.c_size_t(0)
        self._ck(lib().zk_prove_fibsq(self._h, log_n, secret, _p(tr), cap, C.byref(n), _p(st)))
        return tr[: n.value].tobytes(), st.tobytes()

    def prove_fibsq_ex(self, log_n, secret=3141592, init_state=None, n_extra_queries=0):
        """zk_prove_fibsq_ex: seeded initial channel state (32 bytes or None) and channel-drawn further test points."""
        cap = lib().zk_transcript_len_ex(log_n, n_extra_queries)
        tr = np.empty(cap, dtype=np.uint8)
        st = np.empty(32, dtype=np.uint8)
        n = C.c_size_t(0)
        self._ck(lib().zk_prove_fibsq_ex(self._h, log_n, secret, init_state, n_extra_queries, _p(tr), cap, C.byref(n), _p(st)))
        return tr[: n.value].tobytes(), st.tobytes()

    def prove_fibsq_from_trace(self, log_n, trace):
        trace = _u32(trace)
        assert len(trace) == (1 << log_n) - 1
        cap = transcript_len(log_n)
        tr = np.empty(cap, dtype=np.uint8)
        st = np.empty(32, dtype=np.uint8)
        n = C.c_size_t(0)
        self._ck(lib().zk_prove_fibsq_from_trace(self._h, log_n, _p(trace), _p(tr), cap, C.byref(n), _p(st)))
        return tr[: n.value].tobytes(), st.tobytes()

    def prove_fibsq_batch(self, log_n, traces):
        traces = _u32(traces)
        assert traces.ndim == 2 and traces.shape[1] == (1 << log_n) - 1
        nb = traces.shape[0]
        tl = transcript_len(log_n)
        tr = np.empty((nb, tl), dtype=np.uint8)
        st = np.empty((nb, 32), dtype=np.uint8)
        self._ck(lib().zk_prove_fibsq_batch(self._h, log_n, nb, _p(traces), _p(tr), _p(st)))
        return tr, st

    def prove_fibsq_batch_secrets(self, log_n, secrets):
        secrets = _u32(secrets)
        nb = len(secrets)
        tl = transcript_len(log_n)
        tr = np.empty((nb, tl), dtype=np.uint8)
        st = np.empty((nb, 32), dtype=np.uint8)
        self._ck(lib().zk_prove_fibsq_batch_secrets(self._h, log_n, nb, _p(secrets), _p(tr), _p(st)))
        return tr, st

    def prove_stream_submit(self, log_n, secrets):
        secrets = _u32(secrets)
        self._ck(lib().zk_prove_stream_submit(self._h, log_n, len(secrets), _p(secrets)))

    def prove_stream_collect_ptr(self, transcripts_ptr, states_ptr):
        self._ck(lib().zk_prove_stream_collect(self._h, C.c_void_p(transcripts_ptr), C.c_void_p(states_ptr)))

    def prove_stream_collect(self, log_n, n_proofs):
        tl = transcript_len(log_n)
        tr = np.empty((n_proofs, tl), dtype=np.uint8)
        st = np.empty((n_proofs, 32), dtype=np.uint8)
        self.prove_stream_collect_ptr(tr.ctypes.data, st.ctypes.data)
        return tr, st

    # ---- raw-pointer entry points (device memory owned by the caller, e.g. torch tensors)
    def prove_fibsq_batch_ptr(self, log_n, n_proofs, traces_host_ptr, transcripts_ptr, states_ptr):
        self._ck(lib().zk_prove_fibsq_batch(self._h, log_n, n_proofs, C.c_void_p(traces_host_ptr),
                                            C.c_void_p(transcripts_ptr), C.c_void_p(states_ptr)))

    def prove_fibsq_batch_secrets_ptr(self, log_n, n_proofs, secrets_host_ptr, transcripts_ptr, states_ptr):
        self._ck(lib().zk_prove_fibsq_batch_secrets(self._h, log_n, n_proofs, C.c_void_p(secrets_host_ptr),
                                                    C.c_void_p(transcripts_ptr), C.c_void_p(states_ptr)))

    def prove_fibsq_from_trace_ptr(self, log_n, trace_host_ptr, transcript_ptr, cap, state_ptr):
        n = C.c_size_t(0)
        self._ck(lib().zk_prove_fibsq_from_trace(self._h, log_n, C.c_void_p(trace_host_ptr), C.c_void_p(transcript_ptr), cap,
                                                 C.byref(n), C.c_void_p(state_ptr)))
        return n.value

    def prove_fibsq_resident_ptr(self, log_n, n_proofs, traces_dev_ptr, transcripts_ptr, states_ptr):
        self._ck(lib().zk_prove_fibsq_resident(self._h, log_n, n_proofs, C.c_void_p(traces_dev_ptr),
                                               C.c_void_p(transcripts_ptr), C.c_void_p(states_ptr)))

    def interpolate_subgroup_dev(self, y_ptr, n_points, log_n, coeffs_ptr):
        self._ck(lib().zk_interpolate_subgroup_dev(self._h, C.c_void_p(y_ptr), n_points, log_n, C.c_void_p(coeffs_ptr)))

    def lde_coset_dev(self, coeffs_ptr, log_n, evals_ptr, log_blowup=3, shift=5):
        self._ck(lib().zk_lde_coset_dev(self._h, C.c_void_p(coeffs_ptr), log_n, log_blowup, shift, C.c_void_p(evals_ptr)))

    def fri_fold_dev(self, cur_ptr, length, layer_k, beta_raw, next_ptr):
        self._ck(lib().zk_fri_fold_dev(self._h, C.c_void_p(cur_ptr), length, layer_k, beta_raw, C.c_void_p(next_ptr)))

    def cp_eval_fibsq_dev(self, f_ptr, log_n, alpha_raw, a_first, a_last, cp_ptr, log_blowup=3):
        al = _u32(alpha_raw)
        self._ck(lib().zk_cp_eval_fibsq_dev(self._h, C.c_void_p(f_ptr), log_n, log_blowup, _p(al), a_first, a_last,
                                            C.c_void_p(cp_ptr)))

    def merkle_build_dev(self, leaves_ptr, size, heap_ptr):
        self._ck(lib().zk_merkle_build_dev(self._h, C.c_void_p(leaves_ptr), size, C.c_void_p(heap_ptr)))

    def open_queries(self, log_n, xs, proof_index=0):
        """Further openings of a proof of the LAST prove call (zk_open_queries): one bytes object per position, each the
        4 + log_n records of prover.rs:266-289 for that position."""
        xs = np.ascontiguousarray(xs, dtype=np.uint32)
        dl = lib().zk_decommit_len(log_n)
        out = np.zeros(dl * len(xs), dtype=np.uint8)
        self._ck(lib().zk_open_queries(self._h, log_n, proof_index, xs.ctypes.data_as(C.c_void_p), len(xs),
                                       out.ctypes.data_as(C.c_void_p), out.nbytes))
        return [out[dl * i: dl * (i + 1)].tobytes() for i in range(len(xs))]

    # ---- building blocks of a proof sharded over GPUs (zkstark_b200/sharded.py)
    def shard_state_init(self, state_ptr, a_first, a_last):
        self._ck(lib().zk_shard_state_init_dev(self._h, C.c_void_p(state_ptr), a_first, a_last))

    def shard_lde_coset(self, coeffs_ptr, log_n, coset, work_ptr, out_ptr):
        self._ck(lib().zk_shard_lde_coset_dev(self._h, C.c_void_p(coeffs_ptr), log_n, coset, C.c_void_p(work_ptr), C.c_void_p(out_ptr)))

    def shard_merkle_blocks(self, vals_ptr, local_len, log_block, heap_ptr):
        self._ck(lib().zk_shard_merkle_blocks_dev(self._h, C.c_void_p(vals_ptr), local_len, log_block, C.c_void_p(heap_ptr)))

    def shard_upper_tree(self, upper_heap_ptr, log_count, post_op, fri_layer, n_layers, state_ptr, transcript_ptr):
        self._ck(lib().zk_shard_upper_tree_dev(self._h, C.c_void_p(upper_heap_ptr), log_count, post_op, fri_layer, n_layers,
                                               C.c_void_p(state_ptr), C.c_void_p(transcript_ptr)))

    def shard_cp(self, f_ptr, halo_ptr, local_len, log_n, n_ranks, rank, log_block, state_ptr, cp_ptr):
        self._ck(lib().zk_shard_cp_dev(self._h, C.c_void_p(f_ptr), C.c_void_p(halo_ptr), local_len, log_n, n_ranks, rank, log_block,
                                       C.c_void_p(state_ptr), C.c_void_p(cp_ptr)))

    def shard_fold(self, cur_ptr, local_len, layer_k, log_N0, n_ranks, rank, log_block, state_ptr, next_ptr):
        self._ck(lib().zk_shard_fold_dev(self._h, C.c_void_p(cur_ptr), local_len, layer_k, log_N0, n_ranks, rank, log_block,
                                         C.c_void_p(state_ptr), C.c_void_p(next_ptr)))

    def shard_decommit(self, log_n, n_ranks, rank, log_block, first_tail, tree_ptrs, tail_vals_ptr, tail_heaps_ptr,
                       state_ptr, transcript_ptr):
        """tree_ptrs: flat list of 3 * (first_tail + 2) device addresses (values, local heap, upper heap per tree)"""
        arr = (C.c_void_p * len(tree_ptrs))(*tree_ptrs)
        self._ck(lib().zk_shard_decommit_dev(self._h, log_n, n_ranks, rank, log_block, first_tail, arr,
                                             C.c_void_p(tail_vals_ptr), C.c_void_p(tail_heaps_ptr),
                                             C.c_void_p(state_ptr), C.c_void_p(transcript_ptr)))

    def shard_cp_blocks(self, f_ptr, halo_ptr, local_len, log_n, n_ranks, rank, log_block, state_ptr, cp_ptr, heap_ptr):
        self._ck(lib().zk_shard_cp_blocks_dev(self._h, C.c_void_p(f_ptr), C.c_void_p(halo_ptr), local_len, log_n, n_ranks, rank,
                                              log_block, C.c_void_p(state_ptr), C.c_void_p(cp_ptr), C.c_void_p(heap_ptr)))

    def shard_fold_blocks(self, cur_ptr, local_len, layer_k, log_N0, n_ranks, rank, log_block, state_ptr, next_ptr, heap_ptr):
        self._ck(lib().zk_shard_fold_blocks_dev(self._h, C.c_void_p(cur_ptr), local_len, layer_k, log_N0, n_ranks, rank, log_block,
                                                C.c_void_p(state_ptr), C.c_void_p(next_ptr), C.c_void_p(heap_ptr)))

    def shard_fri_tail(self, log_n, first_layer, layers_ptr, heaps_ptr, state_ptr, transcript_ptr):
        self._ck(lib().zk_shard_fri_tail_dev(self._h, log_n, first_layer, C.c_void_p(layers_ptr), C.c_void_p(heaps_ptr),
                                             C.c_void_p(state_ptr), C.c_void_p(transcript_ptr)))

    def stream(self):
        """raw cudaStream_t of the context (0 for the legacy default stream)"""
        return lib().zk_ctx_stream(self._h) or 0

    def synchronize(self):
        self._ck(lib().zk_ctx_synchronize(self._h))

    def launch_count(self):
        return lib().zk_ctx_launch_count(self._h)

    def set_stage_timing(self, on=True):
        self._ck(lib().zk_ctx_set_stage_timing(self._h, int(on)))

    def last_stage_ms(self):
        arr = (C.c_float * 7)()
        self._ck(lib().zk_ctx_last_stage_ms(self._h, arr))
        names = ["interpolate", "lde", "f_tree", "cp_tree", "fri", "decommit", "total"]
        return dict(zip(names, [float(x) for x in arr]))

    def bench_kernel(self, which, log_size, iters=10, flush_l2=True):
        """-> (mean ms, min ms, Tiop/s for the int-pipe microbenchmark)"""
        arr = (C.c_float * 3)()
        self._ck(lib().zk_bench_kernel(self._h, which, log_size, iters, int(flush_l2), arr))
        return float(arr[0]), float(arr[1]), float(arr[2])

    def bench_set_tile_direct_log(self, log_leaves):
        self._ck(lib().zk_bench_set_tile_direct_log(self._h, log_leaves))

    def bench_set(self, sha_variant=-1, leaf_pass_k=0):
        self._ck(lib().zk_bench_set(self._h, sha_variant, leaf_pass_k))

    def set_air(self, air, no_specialise=False):
        """The program every prove entry point of this context proves (None: the reference's FibonacciSq)."""
        self._ck(lib().zk_ctx_set_air(self._h, _air(air, no_specialise)))

    def get_air(self):
        out = (C.c_uint32 * 8)()
        self._ck(lib().zk_ctx_get_air(self._h, out))
        return tuple(out[:7]), int(out[7])

    def trim(self, tables_too=False):
        self._ck(lib().zk_ctx_trim(self._h, int(tables_too)))

    def set_memory_budget(self, nbytes):
        self._ck(lib().zk_ctx_set_memory_budget(self._h, nbytes))
