"""one host-buffer MSM of N points (development aid for per-kernel timing under ncu)"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mathlib_b200 as m
lib = m.load(); c = m.Curves[5]
n = int(os.environ.get("N", "10000")); rnd = random.Random(1)
ks = b"".join(rnd.randrange(c.order).to_bytes(32, "big") for _ in range(n))
pts = b"".join(p.Bytes() for p in c.G1MulBatch(c.GenG1.Bytes() * n, ks, n))
for _ in range(2):
    print(c.MsmBatch(pts, ks, n).hex()[:16])
