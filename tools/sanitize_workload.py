#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): small instances of every kernel family
-- the shared-memory staged NTT passes, the named-barrier warp pairs and the last-arriver tile reduction of the tree
kernels, the two-lane batch, the sharded driver with 8 virtual ranks, the extra query openings -- each checked against
the oracle so that a sanitizer-clean run is also a correct one.
    compute-sanitizer --tool racecheck python tools/sanitize_workload.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
import zkstark_b200 as zk

ctx = zk.Context(0)
# one proof: every tree regime up to the pass kernel + tile kernel with several tiles (2^15 leaves), device channel
for k in (4, 10, 12):
    tr, st = ctx.prove_fibsq(k, 3141592)
    otr, ost, _ = oracle.prove_fibsq(k, 3141592)
    assert (tr, st) == (otr, ost), k
# extra openings of the resident proof
assert len(ctx.open_queries(12, [0, 77, (8 << 12) - 17])) == 3
# batch of 12 x 2^11: two lanes on two streams, device channel tail (>= 8 proofs)
secrets = np.arange(12, dtype=np.uint32) + 5
trs, sts = ctx.prove_fibsq_batch_secrets(11, secrets)
for j, s in enumerate(secrets):
    otr, ost, _ = oracle.prove_fibsq(11, int(s))
    assert trs[j].tobytes() == otr and sts[j].tobytes() == ost, j
# per-function exports on their own
rng = np.random.default_rng(1)
leaves = rng.integers(0, zk.P, size=1 << 13, dtype=np.uint64).astype(np.uint32)
assert np.array_equal(ctx.merkle_build(leaves), oracle.merkle_build(leaves))
y = rng.integers(0, zk.P, size=(1 << 11) - 1, dtype=np.uint64).astype(np.uint32)
c = ctx.interpolate_subgroup(y, 11)
e = ctx.lde_coset(c, 11)
f = ctx.fri_fold(e, 0, 12345)
ctx.field_ops(leaves[:4096], leaves[4096:8192])
# the sharded driver, 8 virtual ranks, every layer that can be sharded is sharded
comm = zk.ShardComm.local(ctx, 8)
comm.set(log_block=6, log_tail_len=0)
tr, st = comm.prove(10, zk.trace_fibsq(10))
otr, ost, _ = oracle.prove_fibsq(10, 3141592)
assert (tr, st) == (otr, ost)
comm.close()
ctx.close()
print("sanitize workload ok")
