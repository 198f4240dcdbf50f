"""G2.Mul / G2 MSM timing probe (development aid): 16,384 BLS12-381 points, device resident."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mathlib_b200 as m
import bench
lib = m.load(); dev = torch.device("cuda:0")
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
cid = int(os.environ.get("CID", "5")); c = m.Curves[cid]
n = int(os.environ.get("N", "16384"))
rng = np.random.default_rng(9)
kk = torch.from_numpy(bench.scalars_mod_r(rng, n, cid).reshape(-1)).to(dev)
g2 = torch.frombuffer(bytearray(c.GenG2.Bytes() * n), dtype=torch.uint8).to(dev)
o2 = torch.empty(n * c.G2ByteSize, dtype=torch.uint8, device=dev)
o3 = torch.empty(c.G2ByteSize, dtype=torch.uint8, device=dev)
def t(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
res = {"n": n}
res["g2_mul_ms"] = t(lambda: m.check(lib.b200_g2_mul_batch(cid, n, g2.data_ptr(), kk.data_ptr(), o2.data_ptr(), m.DEVICE_PTRS)))
res["g2_msm_ms"] = t(lambda: m.check(lib.b200_g2_msm(cid, n, o2.data_ptr(), kk.data_ptr(), o3.data_ptr(), m.DEVICE_PTRS)))
print(json.dumps(res))
