"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the Rust crate, so the parity tests read
like the reference's own tests:

    reference (Rust)                                   here
    Merkle::new(size, iter) / merkle[i] / .trace(i)    Merkle.new(size, data) / m[i] / m.trace(i)      merkle.rs:14,54,74
    lagrange::<F>(&points)   (x_i = g^i)               lagrange_subgroup(ys, log_n)                    polynomial.rs:337
    f_poly.solve over f_domain                         lde_coset(coeffs, log_n)                        prover.rs:69-70
    fri::<F>(&poly, beta) + solve on squared domain    fri_fold(cur_evals, layer_k, beta_raw)          polynomial.rs:385
    constraints + cp_poly.solve                        cp_eval_fibsq(f_eval, log_n, alphas, a0, a_last) prover.rs:101-173
    generate_proof(Channel::new()) -> Proof            generate_proof(Channel.new(), log_n) -> Proof   prover.rs:9

Errors: the reference panics (assert_eq!/unwrap); here every failure raises ZkError.
"""
import numpy as np

from ._binding import Context, ZkError

_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class Merkle:
    """merkle.rs:5-79.  `heap[i]` is node i of the implicit heap: root 0, children 2i+1 / 2i+2."""

    def __init__(self, heap):
        self.heap = heap  # (2*size-1, 32) uint8

    @staticmethod
    def new(size, data, ctx=None):
        data = np.fromiter(data, dtype=np.uint32) if not isinstance(data, np.ndarray) else data
        if size <= 0 or size & (size - 1):
            raise ZkError(-1, "size must be a power of two (merkle.rs:17-21)")
        if len(data) != size:
            raise ZkError(-1, "iterator length must equal size (merkle.rs:28-35 trusts it)")
        return Merkle((ctx or default_context()).merkle_build(data))

    def __getitem__(self, i):  # merkle.rs:74-79
        return self.heap[i].tobytes()

    def trace(self, i):  # merkle.rs:54-71
        from ._binding import merkle_trace_host
        return [row.tobytes() for row in merkle_trace_host(self.heap, i)]


def lagrange_subgroup(ys, log_n, ctx=None):
    """Ascending coefficients of the interpolant through (g^i, ys[i]); len(ys) is n-1 or n."""
    return (ctx or default_context()).interpolate_subgroup(ys, log_n)


def lde_coset(coeffs, log_n, log_blowup=3, shift=5, ctx=None):
    return (ctx or default_context()).lde_coset(coeffs, log_n, log_blowup, shift)


def fri_fold(cur, layer_k, beta_raw, ctx=None):
    return (ctx or default_context()).fri_fold(cur, layer_k, beta_raw)


def cp_eval_fibsq(f_eval, log_n, alpha_raw, a_first, a_last, ctx=None):
    return (ctx or default_context()).cp_eval_fibsq(f_eval, log_n, alpha_raw, a_first, a_last)


class Channel:
    """channel.rs:6-17.  Only the reference's constructor state ([0; 32], empty transcript) exists."""

    def __init__(self):
        self.state = bytes(32)
        self.data = b""

    @staticmethod
    def new():
        return Channel()


class Proof:
    """proof.rs:5-13: the final channel state and the bincode transcript."""

    def __init__(self, state, data, air=None):
        self.state = state
        self.data = data
        self.air = air          # the program the proof is about; None: the reference's FibonacciSq

    def verify(self, log_n=10, a_last=2338775057):
        """proof.rs:15-149: raises (the reference panics) if any check fails; also replays the channel."""
        from ._binding import verify_air
        ok, line = verify_air(self.air, log_n, a_last, self.data, self.state, rederive=True)
        if not ok:
            raise ZkError(-6, f"proof rejected: the assert at proof.rs:{line} fails")

    # ---- wire format.  The reference never serialises a Proof itself (proof.rs:5-13 has no
    # Serialize derive; `Proof::new(state, data)` takes the two parts), so the stable container is what bincode 1.x
    # would write for the struct: the 32 raw state bytes, the u64 little-endian length of `data`, then `data` --
    # `data` being the channel transcript, already bincode (channel.rs:19-26), which the untouched Rust
    # `Proof::new(state, data.into_boxed_slice()).verify()` consumes as is.
    def to_bytes(self):
        from ._binding import proof_serialize
        return proof_serialize(self.state, self.data)          # zk_proof_serialize (C ABI)

    @staticmethod
    def from_bytes(blob):
        from ._binding import proof_deserialize
        return Proof(*proof_deserialize(blob))                  # zk_proof_deserialize (C ABI)

    def size(self):  # proof.rs:151-154: size_of::<Proof>() == 48 (32-byte array + Box<[u8]> fat pointer)
        return 48 + len(self.data)


def generate_proof(channel=None, log_n=10, secret=3141592, ctx=None, air=None):
    """prover.rs:9 generalised to n = 2^log_n (the reference hard-codes log_n = 10) and, with `air` =
    (a0, c20, c02, c11, c10, c01, c00), to the other programs of the AIR family (zk_ctx_set_air; the context returns to
    the reference's program afterwards)."""
    channel = channel or Channel.new()
    if channel.state != bytes(32) or channel.data:
        raise ZkError(-1, "generate_proof expects a fresh Channel::new() (channel.rs:12-17)")
    ctx = ctx or default_context()
    if air is None:
        data, state = ctx.prove_fibsq(log_n, secret)
        return Proof(state, data)
    ctx.set_air(air)
    try:
        data, state = ctx.prove_fibsq(log_n, secret)
    finally:
        ctx.set_air(None)
    return Proof(state, data, air)
