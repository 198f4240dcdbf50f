"""Quick device-resident throughput probe (development aid; bench.py is the contract)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mathlib_b200 as m

lib = m.load()
dev = torch.device("cuda:0")
res = {}
def ev_time(fn, reps=3):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
for cid in (5, 1, 4):
    c = m.Curves[cid]
    n = int(os.environ.get("QB_N", "16384"))
    # inputs: tile the generators' multiples (made on the GPU)
    ks = b"".join(int(i + 2).to_bytes(32, "big") for i in range(n))
    g1 = torch.frombuffer(bytearray(c.GenG1.Bytes() * n), dtype=torch.uint8).to(dev)
    kk = torch.frombuffer(bytearray(ks), dtype=torch.uint8).to(dev)
    pts = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
    t = ev_time(lambda: m.check(lib.b200_g1_mul_batch(cid, n, g1.data_ptr(), kk.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS)))
    res["g1_mul_%d" % cid] = {"n": n, "ms": t, "per_s": n / t * 1e3}
    g2 = torch.frombuffer(bytearray(c.GenG2.Bytes() * n), dtype=torch.uint8).to(dev)
    out = torch.empty(n * c.GtByteSize, dtype=torch.uint8, device=dev)
    t = ev_time(lambda: m.check(lib.b200_pairing_batch(cid, n, pts.data_ptr(), g2.data_ptr(), out.data_ptr(), m.DEVICE_PTRS | m.FEXP)))
    res["pairing_fexp_%d" % cid] = {"n": n, "ms": t, "per_s": n / t * 1e3}
    t = ev_time(lambda: m.check(lib.b200_pairing_batch(cid, n, pts.data_ptr(), g2.data_ptr(), out.data_ptr(), m.DEVICE_PTRS)))
    res["miller_%d" % cid] = {"n": n, "ms": t, "per_s": n / t * 1e3}
    t = ev_time(lambda: m.check(lib.b200_pairing2_batch(cid, n, pts.data_ptr(), g2.data_ptr(), pts.data_ptr(), g2.data_ptr(), out.data_ptr(), m.DEVICE_PTRS | m.FEXP)))
    res["pairing2_fexp_%d" % cid] = {"n": n, "ms": t, "per_s": n / t * 1e3}
    o1 = torch.empty(c.G1ByteSize, dtype=torch.uint8, device=dev)
    for lg in (16, 20):
        nn = 1 << lg
        reps = nn // n
        big_pts = pts.repeat(reps)
        big_k = kk.repeat(reps)
        t = ev_time(lambda: m.check(lib.b200_g1_msm(cid, nn, big_pts.data_ptr(), big_k.data_ptr(), o1.data_ptr(), m.DEVICE_PTRS)))
        res["msm_2^%d_%d" % (lg, cid)] = {"ms": t}
    print(json.dumps(res), flush=True)
print(json.dumps(res))
