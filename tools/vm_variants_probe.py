"""Development aid: time Pairing2+FExp (BLS12-381, kilic semantics) on the library named by B200_LIB and check the
verdicts, at 65,536 and 1,024 checks.  One JSON line."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mathlib_b200 as m
import bench

lib = m.load()
dev = torch.device("cuda:0")
m.check(lib.b200_set_device(0))
lib.b200_set_stream(torch.cuda.current_stream().cuda_stream)
c = m.Curves[3]
res = {"lib": os.path.basename(m.LIB_PATH)}
for n in (65536, 1024):
    ins = bench.make_inputs(m, 3, n, seed=3)
    d = [torch.frombuffer(bytearray(x), dtype=torch.uint8).to(dev) for x in ins[:4]]
    v = torch.empty(n, dtype=torch.uint8, device=dev)
    o = torch.empty(n * c.GtByteSize, dtype=torch.uint8, device=dev)

    def run_v():
        m.check(lib.b200_pairing2_batch(3, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                        v.data_ptr(), m.DEVICE_PTRS | m.FEXP | m.OUT_UNITY_ONLY))

    def run_o():
        m.check(lib.b200_pairing2_batch(3, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                        o.data_ptr(), m.DEVICE_PTRS | m.FEXP))
    run_v()
    torch.cuda.synchronize()
    res["ok_%d" % n] = bool((v.cpu().numpy() == ins[4]).all())
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_o()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res["ms_%d" % n] = best
print(json.dumps(res))
