"""MSM latency with all scalars equal (development aid): worst-case bucket skew."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mathlib_b200 as m
lib = m.load(); dev = torch.device("cuda:0"); c = m.Curves[5]
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
nmax = 1 << 20
rng = np.random.default_rng(5)
ks = rng.integers(0, 256, size=(nmax, 32), dtype=np.uint8); ks[:, 0] &= 0x0F
d_k = torch.from_numpy(ks.reshape(-1)).to(dev)
gen = torch.frombuffer(bytearray(c.GenG1.Bytes()), dtype=torch.uint8).to(dev).repeat(nmax)
pts = torch.empty(nmax * c.G1ByteSize, dtype=torch.uint8, device=dev)
m.check(lib.b200_g1_mul_batch(5, nmax, gen.data_ptr(), d_k.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS | m.OUT_MONT))
one = torch.from_numpy(np.tile(ks[:1], (nmax, 1)).reshape(-1)).to(dev)
o = torch.empty(c.G1ByteSize, dtype=torch.uint8, device=dev)
for lg in (16, 20):
    n = 1 << lg
    def run(): m.check(lib.b200_g1_msm(5, n, pts.data_ptr(), one.data_ptr(), o.data_ptr(), m.DEVICE_PTRS | m.IN_MONT))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"lg_n": lg, "all_equal_scalars_ms": round(e0.elapsed_time(e1), 2)}), flush=True)
