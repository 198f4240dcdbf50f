"""MSM latency at small n (development aid): host buffers, end to end through b200_g1_msm."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mathlib_b200 as m
lib = m.load(); c = m.Curves[5]
rnd = random.Random(1)
for n in (1, 4, 10, 33, 100, 1000, 10000):
    ks = b"".join(rnd.randrange(c.order).to_bytes(32, "big") for _ in range(n))
    pts = b"".join(p.Bytes() for p in c.G1MulBatch(c.GenG1.Bytes() * n, ks, n))
    c.MsmBatch(pts, ks, n)
    t0 = time.perf_counter()
    for _ in range(5):
        c.MsmBatch(pts, ks, n)
    print(n, "msm ms", round((time.perf_counter() - t0) / 5 * 1e3, 3), flush=True)
t0 = time.perf_counter()
for _ in range(5):
    c.GenG1.Mul(c.NewZrFromInt(rnd.randrange(c.order)))
print("single G1.Mul ms", round((time.perf_counter() - t0) / 5 * 1e3, 3))
t0 = time.perf_counter()
for _ in range(3):
    c.FExp(c.Pairing(c.GenG2, c.GenG1))
print("single Pairing+FExp ms", round((time.perf_counter() - t0) / 3 * 1e3, 3))
