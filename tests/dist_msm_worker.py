"""torchrun worker of tests/test_gpu_large.py::test_torchrun_two_ranks_sharded_msm: one process per GPU, NCCL."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import mathlib_b200 as m
    from mathlib_b200 import shard
    import bench
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = m.load()
    m.check(lib.b200_init(0))
    m.check(lib.b200_set_device(local))
    m.check(lib.b200_set_stream(torch.cuda.current_stream().cuda_stream))
    for cid, n in ((5, 1 << 17), (4, 40001), (1, 5)):
        c = m.Curves[cid]
        nd = min(n, 4096)
        kp = bench.scalars_mod_r(np.random.default_rng(7), nd, cid)
        base = b"".join(p.Bytes() for p in c.G1MulBatch(c.GenG1.Bytes() * nd, kp.tobytes(), nd))
        pts = (base * (n // nd + 1))[:n * c.G1ByteSize]
        ks = bench.scalars_mod_r(np.random.default_rng(8), n, cid).tobytes()
        lo, hi = shard.shard_range(n, rank, world)
        d_pts = torch.frombuffer(bytearray(pts[lo * c.G1ByteSize:hi * c.G1ByteSize]), dtype=torch.uint8).to(dev)
        d_ks = torch.frombuffer(bytearray(ks[lo * 32:hi * 32]), dtype=torch.uint8).to(dev)
        got = shard.msm_sharded_device(lib, cid, d_pts, d_ks, hi - lo, dist, dev).cpu().numpy().tobytes()
        # host-buffer flavour of the same split (bytes in, bytes out)
        got2 = shard.msm_sharded(c, pts, ks, n, dist, device=dev)
        if rank == 0:
            from oracle import cpu_binding as orc
            want = orc.g1_msm(cid, n, pts, ks)
            assert got == want, "device-resident sharded MSM differs from the oracle (curve %d)" % cid
            assert got2 == want, "host-buffer sharded MSM differs from the oracle (curve %d)" % cid
    dist.barrier()
    if rank == 0:
        print("SHARDED_MSM_OK ranks=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
