"""Pairing kernel timing probe (development aid): Miller only vs Miller+FExp, Pairing vs Pairing2."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mathlib_b200 as m
import bench
lib = m.load(); dev = torch.device("cuda:0")
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
cid = int(os.environ.get("CID", "5")); c = m.Curves[cid]
n = int(os.environ.get("N", "65536"))
ins = bench.make_inputs(m, cid, n, seed=3)
d = [torch.frombuffer(bytearray(x), dtype=torch.uint8).to(dev) for x in ins[:4]]
o = torch.empty(n * c.GtByteSize, dtype=torch.uint8, device=dev)
def t(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
res = {}
res["miller2_ms"] = t(lambda: m.check(lib.b200_pairing2_batch(cid, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), o.data_ptr(), m.DEVICE_PTRS)))
res["miller2_fexp_ms"] = t(lambda: m.check(lib.b200_pairing2_batch(cid, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), o.data_ptr(), m.DEVICE_PTRS | m.FEXP)))
res["miller1_ms"] = t(lambda: m.check(lib.b200_pairing_batch(cid, n, d[0].data_ptr(), d[1].data_ptr(), o.data_ptr(), m.DEVICE_PTRS)))
res["miller1_fexp_ms"] = t(lambda: m.check(lib.b200_pairing_batch(cid, n, d[0].data_ptr(), d[1].data_ptr(), o.data_ptr(), m.DEVICE_PTRS | m.FEXP)))
print(json.dumps(res))
