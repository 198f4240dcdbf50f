#!/usr/bin/env python3
"""Generate tests/golden/vectors_<curve>.json from the Python oracle (seeded, reproducible).

The reference itself cannot run here (no Go toolchain; its arithmetic lives in un-vendored
modules), so these vectors are oracle outputs, not gnark/kilic outputs.  They pin the C++ CPU
oracle, the host-emulated kernels and the CUDA kernels to one another and to the textbook
definition.  Run from the repo root:  python tests/golden/gen_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.params import CURVE_IDS  # noqa: E402
from oracle.pairing import Pairing  # noqa: E402
from oracle import codec  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def gen(curve_id, seed):
    P, sem = CURVE_IDS[curve_id]
    pr = Pairing(P)
    C, T = pr.C, pr.T
    rnd = random.Random(seed)
    h = lambda b: b.hex()
    g1b = lambda pt: h(codec.g1_to_bytes(P, pt))
    g2b = lambda pt: h(codec.g2_to_bytes(P, pt))
    gtb = lambda f: h(codec.gt_to_bytes(P, f))
    rs = lambda: rnd.randrange(1, P.r)
    out = {"curve_id": curve_id, "name": P.name, "semantics": sem, "seed": seed}

    # ---- pairing: Pairing(G2,G1) raw (driver semantics), and FExp of it
    cases = []
    a, b = rs(), rs()
    pts = [(C.g1_mul(C.g1, a), C.g2_mul(C.g2, b)),
           (C.g1, C.g2),
           (None, C.g2_mul(C.g2, rs())),
           (C.g1_mul(C.g1, rs()), None),
           (C.g1_mul(C.g1, rs()), C.g2_mul(C.g2, rs()))]
    for (p1, p2) in pts:
        raw = pr.pairing(p2, p1, sem)
        cases.append({"g1": g1b(p1), "g2": g2b(p2), "pairing": gtb(raw), "fexp": gtb(pr.fexp(raw, sem)),
                      "canonical": gtb(pr.final_exp(pr.miller_textbook([(p1, p2)])))})
    out["pairing"] = cases

    # ---- pairing2
    cases = []
    x = rs()
    sig_like = (C.g1_mul(C.g1, x), C.g2, C.g1_neg(C.g1), C.g2_mul(C.g2, x))      # e(xG1,G2)*e(-G1,xG2) = 1
    quads = [(C.g1_mul(C.g1, rs()), C.g2_mul(C.g2, rs()), C.g1_mul(C.g1, rs()), C.g2_mul(C.g2, rs())),
             sig_like,
             (None, C.g2_mul(C.g2, rs()), C.g1_mul(C.g1, rs()), C.g2_mul(C.g2, rs())),
             (C.g1_mul(C.g1, rs()), C.g2_mul(C.g2, rs()), C.g1_mul(C.g1, rs()), None)]
    for (p1a, p2a, p1b, p2b) in quads:
        raw = pr.pairing2(p2a, p2b, p1a, p1b, sem)
        fe = pr.fexp(raw, sem)
        cases.append({"g1a": g1b(p1a), "g2a": g2b(p2a), "g1b": g1b(p1b), "g2b": g2b(p2b),
                      "pairing2": gtb(raw), "fexp": gtb(fe), "unity": fe == T.f12_one})
    out["pairing2"] = cases

    # ---- G1.Mul
    cases = []
    base = C.g1_mul(C.g1, rs())
    for k in [0, 1, 2, P.r - 1, P.r, P.r + 5, rs(), rs(), (1 << 256) - 1]:
        cases.append({"p": g1b(base), "k": h(int(k % (1 << 256)).to_bytes(32, "big")), "out": g1b(C.g1_mul(base, k))})
    cases.append({"p": g1b(None), "k": h(int(rs()).to_bytes(32, "big")), "out": g1b(None)})
    out["g1_mul"] = cases

    # ---- G1.Mul2
    cases = []
    Q = C.g1_mul(C.g1, rs())
    for (e, f) in [(rs(), rs()), (0, rs()), (rs(), 0), (1, P.r - 1), ((1 << 63) - 1, (1 << 63) - 1), (5, 5)]:
        cases.append({"p": g1b(base), "e": h(int(e).to_bytes(32, "big")), "q": g1b(Q), "f": h(int(f).to_bytes(32, "big")),
                      "out": g1b(C.g1_mul2(base, e, Q, f))})
    # same base twice, opposite scalars -> infinity ; P and -P
    cases.append({"p": g1b(base), "e": h(int(7).to_bytes(32, "big")), "q": g1b(base), "f": h(int(P.r - 7).to_bytes(32, "big")),
                  "out": g1b(None)})
    cases.append({"p": g1b(base), "e": h(int(9).to_bytes(32, "big")), "q": g1b(C.g1_neg(base)), "f": h(int(4).to_bytes(32, "big")),
                  "out": g1b(C.g1_mul(base, 5))})
    out["g1_mul2"] = cases

    # ---- MSM
    cases = []
    for n in [0, 1, 2, 10, 33]:
        pts_ = [C.g1_mul(C.g1, rs()) for _ in range(n)]
        ks = [rs() for _ in range(n)]
        if n >= 10:
            pts_[3] = None                  # infinity point
            ks[4] = 0                       # zero scalar
            pts_[6] = pts_[5]               # repeated point
            ks[6] = ks[5]                   # ... in the same bucket
            pts_[8] = C.g1_neg(pts_[7])     # P and -P with the same scalar cancel
            ks[8] = ks[7]
            ks[9] = P.r - 1
        cases.append({"n": n, "points": "".join(g1b(p) for p in pts_),
                      "scalars": "".join(h(int(k).to_bytes(32, "big")) for k in ks),
                      "out": g1b(C.g1_msm(pts_, ks))})
    out["msm"] = cases

    # ---- G2.Mul / G2.Add (appended after the hot-path cases so their seeded values are unchanged)
    cases = []
    base2 = C.g2_mul(C.g2, rs())
    for k in [0, 1, 2, P.r - 1, P.r, rs(), rs(), (1 << 256) - 1]:
        cases.append({"p": g2b(base2), "k": h(int(k % (1 << 256)).to_bytes(32, "big")), "out": g2b(C.g2_mul(base2, k))})
    cases.append({"p": g2b(None), "k": h(int(rs()).to_bytes(32, "big")), "out": g2b(None)})
    out["g2_mul"] = cases
    Q2 = C.g2_mul(C.g2, rs())
    out["g2_add"] = [{"p": g2b(a_), "q": g2b(b_), "out": g2b(C.g2_add(a_, b_))}
                     for (a_, b_) in [(base2, Q2), (base2, base2), (base2, C.g2_neg(base2)), (None, Q2), (Q2, None)]]

    # ---- Gt.Mul / Gt.Inverse / Gt.Exp on a pairing value (order r) and on a raw Miller-loop value (not cyclotomic)
    p1, p2 = C.g1_mul(C.g1, rs()), C.g2_mul(C.g2, rs())
    e_can = pr.final_exp(pr.miller_textbook([(p1, p2)]))
    e_raw = pr.miller_projective([(p1, p2)])
    cases = []
    for el in (e_can, e_raw):
        for k in [0, 1, 2, 3, P.r - 1, P.r, rs(), (1 << 256) - 1]:
            cases.append({"a": gtb(el), "k": h(int(k).to_bytes(32, "big")), "out": gtb(T.f12_pow(el, k))})
    out["gt_exp"] = cases
    out["gt_mul"] = [{"a": gtb(x_), "b": gtb(y_), "out": gtb(T.f12_mul(x_, y_))}
                     for (x_, y_) in [(e_can, e_raw), (e_raw, e_raw), (e_can, T.f12_one), (e_can, T.f12_inv(e_can))]]
    out["gt_inv"] = [{"a": gtb(x_), "out": gtb(T.f12_inv(x_))} for x_ in (e_can, e_raw, T.f12_one)]
    return out


def main():
    for cid, seed in [(1, 101), (3, 303), (4, 404), (5, 505)]:
        v = gen(cid, seed)
        path = os.path.join(HERE, "vectors_%d.json" % cid)
        with open(path, "w") as f:
            json.dump(v, f, indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main()
