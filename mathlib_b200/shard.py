"""Multi-GPU partitioning of the hot path, one process per GPU under torch.distributed.

* Pairing / FExp / Mul / Mul2 batches are independent units: contiguous index split, NO collective.
* MultiScalarMul: split the point range, each rank reduces its range to one affine partial sum on its GPU,
  partials (2*FpBytes each) are all-gathered (NCCL over NVLink, or gloo in the CPU tests) and every rank adds the
  `world` partials with b200_g1_sum.  Bandwidth is irrelevant (<= 8 x 96 B); the step is latency only.
"""


def shard_range(n, rank, world):
    """[lo, hi) of the contiguous index split used everywhere (same formula as abi.cu run_split)."""
    return n * rank // world, n * (rank + 1) // world


def gather_bytes(payload, dist, device=None):
    """all_gather equal-length byte strings; returns the list ordered by rank."""
    import torch
    world = dist.get_world_size()
    t = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def msm_sharded(curve, pts, scalars, n, dist, device=None, msm_fn=None, sum_fn=None):
    """sum_i [k_i]P_i with the point range split over the ranks of `dist`.

    pts / scalars hold ALL n elements in the reference BYTES encoding on every rank (a production caller
    would keep only its own shard resident).  msm_fn / sum_fn default to the GPU entry points; the CPU tests
    inject oracle functions to exercise the plumbing without a GPU.
    """
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    g1sz = curve.G1ByteSize
    msm_fn = msm_fn or (lambda p, k, m: curve.MsmBatch(p, k, m))
    sum_fn = sum_fn or (lambda parts: curve.G1Sum([curve.NewG1FromBytes(x) for x in parts]).Bytes())
    partial = msm_fn(pts[lo * g1sz:hi * g1sz], scalars[lo * 32:hi * 32], hi - lo)
    parts = gather_bytes(partial, dist, device)
    return sum_fn(parts)


def msm_sharded_device(lib, cid, d_pts, d_scalars, n_local, dist, device, in_flags=0, out_flags=0, scratch=None):
    """Device-resident flavour: this rank's point range (torch uint8 tensor on `device`, BYTES or MONT per in_flags) and
    scalars stay in HBM; local MSM -> 2*FpBytes affine partial (Montgomery limbs) -> NCCL all-gather -> b200_g1_sum.
    Everything is enqueued on the current stream; returns the device tensor holding the combined point (every rank has
    it).  `scratch` = (partial, gathered, out) tensors to reuse between calls."""
    import torch
    from ._lib import DEVICE_PTRS, IN_MONT, OUT_MONT, check
    world = dist.get_world_size() if dist is not None else 1
    g1sz = 2 * lib.b200_fp_bytes(cid)
    if scratch is None:
        scratch = (torch.empty(g1sz, dtype=torch.uint8, device=device),
                   torch.empty(world * g1sz, dtype=torch.uint8, device=device),
                   torch.empty(g1sz, dtype=torch.uint8, device=device))
    part, gath, out = scratch
    check(lib.b200_g1_msm(cid, n_local, d_pts.data_ptr() if n_local else None, d_scalars.data_ptr() if n_local else None,
                          part.data_ptr(), DEVICE_PTRS | in_flags | OUT_MONT))
    if world > 1:
        dist.all_gather_into_tensor(gath, part)
        src = gath
    else:
        src = part
    check(lib.b200_g1_sum(cid, world, src.data_ptr(), out.data_ptr(), DEVICE_PTRS | IN_MONT | out_flags))
    return out
