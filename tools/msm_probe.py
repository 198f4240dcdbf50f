"""MSM 2^20 probe with uniform scalars (development aid): prints latency; run under ncu for the per-kernel split."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mathlib_b200 as m
lib = m.load(); dev = torch.device("cuda:0")
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
cid = int(os.environ.get("CID", "5")); c = m.Curves[cid]
lg = int(os.environ.get("LG", "20")); n = 1 << lg
rng = np.random.default_rng(5)
import bench
def rand_scalars():
    return torch.from_numpy(bench.scalars_mod_r(rng, n, cid).reshape(-1)).to(dev)        # uniform in [0, r), as in bench.py
d_k = rand_scalars()
gen = torch.frombuffer(bytearray(c.GenG1.Bytes()), dtype=torch.uint8).to(dev).repeat(n)
pts = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
m.check(lib.b200_g1_mul_batch(cid, n, gen.data_ptr(), d_k.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS | m.OUT_MONT))
d_k2 = rand_scalars()
o = torch.empty(c.G1ByteSize, dtype=torch.uint8, device=dev)
def run(): m.check(lib.b200_g1_msm(cid, n, pts.data_ptr(), d_k2.data_ptr(), o.data_ptr(), m.DEVICE_PTRS | m.IN_MONT))
run(); torch.cuda.synchronize()
for _ in range(int(os.environ.get("REPS", "3"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"n": n, "ms": e0.elapsed_time(e1)}), flush=True)
print(o.cpu().numpy().tobytes().hex())
