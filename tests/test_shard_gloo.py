"""N>1 host logic on CPU: world_size-2 gloo processes exercise the MSM range split + partial-sum combine plumbing
(mathlib_b200.shard) with the oracle standing in for the per-rank GPU calls."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mathlib_b200 import shard
    from mathlib_b200.driver import Curve
    from oracle import cpu_binding as orc
    from oracle.params import CURVE_IDS
    from oracle.curve import Curve as OC
    from oracle import codec
    with open(os.path.join(ROOT, "tests", "golden", "vectors_5.json")) as f:
        case = json.load(f)["msm"][4]
    c = Curve(5)
    P, _ = CURVE_IDS[5]
    oc = OC(P)

    def sum_fn(parts):
        acc = None
        for x in parts:
            acc = oc.g1_add(acc, codec.g1_from_bytes(P, x))
        return codec.g1_to_bytes(P, acc)

    out = shard.msm_sharded(c, bytes.fromhex(case["points"]), bytes.fromhex(case["scalars"]), case["n"], dist,
                            msm_fn=lambda p, k, m: orc.g1_msm(5, m, p, k, nthreads=1), sum_fn=sum_fn)
    q.put((rank, out.hex() == case["out"], shard.shard_range(case["n"], rank, world)))
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from mathlib_b200.shard import shard_range
    for n in (0, 1, 7, 1024, 65537):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))


def test_msm_sharded_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sorted(r for _, _, r in res) == [(0, 16), (16, 33)]
