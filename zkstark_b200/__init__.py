"""zkstark_b200 -- B200 (sm_100a) hot path of the STARK-101 FibonacciSq prover of
Crocodoctopus/zkstark behind the reference's own entry points.

The product is `libzkstark_b200.so` (hand-written CUDA + a C ABI, include/zkstark_b200.h).
This package is the thin host-side mirror used by tests/ and bench.py: a ctypes binding plus
classes named after the reference's (`Merkle::new`/`trace`, `lagrange`, `fri`, `generate_proof`,
`Proof`).  There is NO CPU fallback: importing works without a GPU (so the ABI can be checked),
but creating a `Context` without a CUDA device raises.
"""
from ._binding import (  # noqa: F401
    LIB_PATH,
    Context,
    ShardComm,
    shard_unique_id,
    ZkError,
    build_library,
    exported_symbols,
    header_symbols,
    lib,
    transcript_len,
    trace_fibsq,
    trace_fibsq_batch,
    trace_air,
    trace_air_batch,
    verify_air,
    AIR_FIBSQ,
    merkle_trace_host,
    proof_serialize,
    proof_deserialize,
    verify_fibsq,
    verify_fibsq_ex,
)
from .api import (  # noqa: F401
    Channel,
    Merkle,
    Proof,
    cp_eval_fibsq,
    fri_fold,
    generate_proof,
    lagrange_subgroup,
    lde_coset,
    default_context,
)

P = 3221225473
LOG_BLOWUP = 3
