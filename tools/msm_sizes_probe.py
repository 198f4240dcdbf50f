"""MSM latency across sizes (development aid), device-resident Montgomery points, uniform scalars."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mathlib_b200 as m
lib = m.load(); dev = torch.device("cuda:0")
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
cid = int(os.environ.get("CID", "5")); c = m.Curves[cid]
nmax = 1 << 20
rng = np.random.default_rng(5)
import bench
def rand_scalars(n):
    return torch.from_numpy(bench.scalars_mod_r(rng, n, cid).reshape(-1)).to(dev)      # uniform in [0, r)
d_k = rand_scalars(nmax)
gen = torch.frombuffer(bytearray(c.GenG1.Bytes()), dtype=torch.uint8).to(dev).repeat(nmax)
pts = torch.empty(nmax * c.G1ByteSize, dtype=torch.uint8, device=dev)
m.check(lib.b200_g1_mul_batch(cid, nmax, gen.data_ptr(), d_k.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS | m.OUT_MONT))
d_k2 = rand_scalars(nmax)
o = torch.empty(c.G1ByteSize, dtype=torch.uint8, device=dev)
for lg in (4, 7, 10, 12, 13, 14, 15, 16, 17, 18, 19, 20):
    n = 1 << lg
    def run(): m.check(lib.b200_g1_msm(cid, n, pts.data_ptr(), d_k2.data_ptr(), o.data_ptr(), m.DEVICE_PTRS | m.IN_MONT))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"lg_n": lg, "ms": round(e0.elapsed_time(e1), 3)}), flush=True)
