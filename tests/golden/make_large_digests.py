"""Generates tests/golden/large_digests.json: sha256 of the transcript and the final channel state of the oracle's
proofs (NTT route, secret 3141592) at the sizes too slow to recompute in every test run.  Run from the repo root:
    python tests/golden/make_large_digests.py
(2^24 takes ~35 s on 8 cores.)"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle

out = {"_provenance": "oracle.prove_fibsq(log_n, 3141592, ROUTE_FAST); the oracle itself is pinned on the reference KATs and "
                      "the Appendix-A transcripts (tests/test_oracle.py)"}
for k in (18, 20, 22, 24):
    tr, st, info = oracle.prove_fibsq(k, 3141592, oracle.ROUTE_FAST, oracle.default_threads())
    out[str(k)] = {"len": len(tr), "sha256": hashlib.sha256(tr).hexdigest(), "state": st.hex(), "a_last": int(info.a_last)}
    print(k, out[str(k)], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "large_digests.json"), "w"), indent=1)
