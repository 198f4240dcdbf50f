#!/usr/bin/env python3
"""G2 point pool for benchmarks / large parity tests: Q_j = [b_j]G2 with the scalars b_j recorded
(generated with the Python oracle; the product has no G2 scalar-mul kernel yet -- SURVEY 8f-3)."""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.params import CURVE_IDS
from oracle.curve import Curve
from oracle import codec

out = {}
for cid in (1, 3, 4):
    P, _ = CURVE_IDS[cid]
    C = Curve(P)
    rnd = random.Random(9000 + cid)
    entries = []
    for j in range(128):
        b = rnd.randrange(1, P.r)
        entries.append({"b": "%064x" % b, "g2": codec.g2_to_bytes(P, C.g2_mul(C.g2, b)).hex()})
    out[str(cid)] = entries
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "g2_pool.json"), "w") as f:
    json.dump(out, f)
print("ok")
