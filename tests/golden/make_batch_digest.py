"""Generates tests/golden/batch_2p16_digest.json: BASELINE configs[3] -- 1024 independent proofs of 2^16 traces, secret_j =
3141592 + j -- as the oracle proves them (NTT route): the sha256 of every transcript, folded into one
checksum of checksums sha256(hex digests concatenated, in proof order), per prefix length so that smaller batches can be
checked too.  Run from the repo root (about a minute on 8 cores):
    python tests/golden/make_batch_digest.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle

K, COUNT = 16, 1024
digests, states = [], []
for j in range(COUNT):
    tr, st, _ = oracle.prove_fibsq(K, 3141592 + j, oracle.ROUTE_FAST, oracle.default_threads())
    digests.append(hashlib.sha256(tr).hexdigest())
    states.append(st.hex())
    if j % 128 == 0:
        print(j, digests[-1], flush=True)
out = {"_provenance": "oracle.prove_fibsq(16, 3141592 + j, ROUTE_FAST), j < 1024 (tests/golden/make_batch_digest.py)",
       "log_n": K, "count": COUNT,
       "checksum_of_checksums": {str(c): hashlib.sha256("".join(digests[:c]).encode()).hexdigest() for c in (1, 8, 64, 128, 256, 512, 1024)},
       "states_checksum": hashlib.sha256("".join(states).encode()).hexdigest(),
       "first": digests[0], "last": digests[-1]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "batch_2p16_digest.json"), "w"), indent=1)
print(out["checksum_of_checksums"])
