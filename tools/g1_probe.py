import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mathlib_b200 as m
lib = m.load(); dev = torch.device("cuda:0")
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
cid = int(os.environ.get("CID", "5")); c = m.Curves[cid]; n = int(os.environ.get("N", "65536"))
rng = np.random.default_rng(1)
ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); ks[:, 0] &= 0x3F
d_k = torch.from_numpy(ks.reshape(-1)).to(dev)
gen = torch.frombuffer(bytearray(c.GenG1.Bytes()), dtype=torch.uint8).to(dev).repeat(n)
pts = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
out = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
def t(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
m.check(lib.b200_g1_mul_batch(cid, n, gen.data_ptr(), d_k.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS))
r = {"g1_mul_ms": t(lambda: m.check(lib.b200_g1_mul_batch(cid, n, pts.data_ptr(), d_k.data_ptr(), out.data_ptr(), m.DEVICE_PTRS))),
     "g1_mul2_ms": t(lambda: m.check(lib.b200_g1_mul2_batch(cid, n, pts.data_ptr(), d_k.data_ptr(), gen.data_ptr(), d_k.data_ptr(), out.data_ptr(), m.DEVICE_PTRS)))}
r["g1_mul_per_s"] = n / r["g1_mul_ms"] * 1e3; r["g1_mul2_per_s"] = n / r["g1_mul2_ms"] * 1e3
print(json.dumps(r))
