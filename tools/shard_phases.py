#!/usr/bin/env python
"""Phase breakdown of the native sharded prover with G virtual ranks on one GPU, beside the monolithic pipeline.
  python tools/shard_phases.py [log_n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk

ctx = zk.Context(0)
for k in [int(a) for a in sys.argv[1:]] or [20, 24]:
    tr = zk.trace_fibsq(k)
    for G in (1, 8):
        c = zk.ShardComm.local(ctx, G)
        for _ in range(3):
            c.prove(k, tr)
        print(k, G, {a: round(b, 3) for a, b in c.last_phase_ms().items()}, "first_tail", c.last_first_tail(), flush=True)
        c.close()
    ctx.set_stage_timing(True)
    for _ in range(3):
        ctx.prove_fibsq_from_trace(k, tr)
    print(k, "mono", {a: round(b, 3) for a, b in ctx.last_stage_ms().items()}, flush=True)
    ctx.set_stage_timing(False)
