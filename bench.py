#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 backend for IBM/mathlib's pairing / G1 hot path.

Contract (see task statement):  python bench.py --gpus N --steps K --warmup W [--impl reference]
prints ONE JSON line on rank 0.

Workload (BASELINE.json metric "BLS12-381 pairings/s & G1 MSM 2^20 latency at 1/2/4/8 B200 vs CPU driver"):
  headline   one step = one batch of 65,536 independent BLS12-381 (kilic semantics) Pairing2+FExp checks per GPU (the
             op of configs[0] / perf_test.go:541-560 at the batch size of configs[1]: 1,024 checks cannot fill 148 SMs).
             Half of the batch are valid BLS-style checks (product == 1), half random.  `value` counts pairings (2 per
             check) with inputs resident in HBM; `e2e` is the same metric through the C ABI with HOST buffers (H2D +
             kernel + D2H inside the timed region).
  extra      the BASELINE configs, at EVERY N (torchrun): configs[0] verbatim (1,024 checks, with its own e2e),
             configs[1] BN254, configs[2] MSM 2^20 STRONG-scaled over the N ranks (range split, NCCL all-gather of the
             96-byte partial sums, b200_g1_sum) with latency, e2e from host buffers, roofline and CPU baseline,
             configs[3] 2^21 BLS12-377 points per rank + combine, configs[4] 12,500 BBS-style verifications per rank.

--impl reference times the reference's CPU algorithm for the same check.  The reference's Go drivers cannot be
built here (no Go toolchain), so this arm runs oracle/cpu (the C++ restatement, kind "port") on all host threads, on a
bounded sample (the first 1,024 checks) of the very same seeded 65,536-check batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CID = 3                      # BLS12_381, kilic semantics: Pairing2 includes the final exponentiation
BATCH = 65536
REF_SAMPLE = 1024            # checks per step of the CPU arm (bounded sample of the same batch)
# Algorithmic work per op in Fp Montgomery products `m`, counted by instrumenting the optimised CPU oracle path
# (oracle/cpu, orc_mul_count) -- see DESIGN.md "work units".  FERMAT_M: the Fp inversion of the easy part, which the
# oracle does as a^(p-2) (~570 m) and the GPU does with a binary Euclid on the ALU pipe: the roofline counts the
# multiplier work only, so it is subtracted.
M_PER_OP = {"bls381_pairing2_fexp": 19825, "bls381_pairing_fexp": 15198, "bn254_pairing_fexp": 17496}
FERMAT_M = {12: 570, 8: 380}
MAC_PER_M = {12: 300, 8: 136}            # 2n^2 + n 32x32->64 multiply-accumulates per Montgomery product
# IMAD.WIDE.U32 issue peak of one B200, measured with tools/imad_peak.cu (profiles/peaks_r1.json):
# fmaheavy pipe, 4 cycles per warp instruction -> 32 lanes/clk/SM * 148 SM * 1.965 GHz = 9.31e12/s nominal.
IMAD_WIDE_PEAK = 8.98e12
ORDERS = {
    1: 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
    3: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    4: 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
}
ORDERS[5] = ORDERS[6] = ORDERS[7] = ORDERS[3]


def headline_config(n, world):
    """config of the headline line -- shared by both arms so that the driver sees the same workload description."""
    return {"workload": "BLS12_381 (kilic semantics, mathlib CurveID 3) Pairing2+FExp checks, %d per step per GPU, "
                        "2 pairings per check; half valid BLS checks, half random (seeded batch, seed 1 + rank)" % n,
            "curve_id": CID, "checks_per_step_per_gpu": n, "parallelism": "index-split x%d, no collective" % world,
            "l2": "256 MiB flush write between timed iterations", "encoding": "reference Bytes() formats"}


def check_scalars(cid, n, seed):
    """the seeded scalars of an n-check batch: a_i, pool index j_i, and k2_i (valid check for even i: k2 = a*b_j)."""
    import random
    with open(os.path.join(ROOT, "tests", "golden", "g2_pool.json")) as f:
        pool = json.load(f)[str({3: 3, 5: 3, 6: 3, 7: 3, 1: 1, 4: 4}[cid])]
    rnd = random.Random(seed)
    r = ORDERS[cid]
    a = [rnd.randrange(1, r) for _ in range(n)]
    j = [rnd.randrange(len(pool)) for _ in range(n)]
    b = [int(pool[x]["b"], 16) for x in j]
    k2 = [(a[i] * b[i]) % r if i % 2 == 0 else rnd.randrange(1, r) for i in range(n)]
    return pool, a, j, k2


def make_inputs(m, cid, n, seed=1):
    """n Pairing2 checks in the reference's BYTES encoding: half valid (e(aG1,Q)e(-abG1,G2) = 1), half random.
    G1 points are made on the GPU (b200_g1_mul_batch); G2 points come from the committed pool
    tests/golden/g2_pool.json (Q_j = [b_j]G2)."""
    import ctypes
    import numpy as np
    c = m.Curves[cid]
    pool, a, j, k2 = check_scalars(cid, n, seed)
    r = c.order
    gen = c.GenG1.Bytes()
    ka = b"".join(x.to_bytes(32, "big") for x in a)
    kb = b"".join((r - x).to_bytes(32, "big") for x in k2)        # -k2 mod r
    lib = m.load()
    g1a = ctypes.create_string_buffer(n * c.G1ByteSize)
    g1b = ctypes.create_string_buffer(n * c.G1ByteSize)
    m.check(lib.b200_g1_mul_batch(cid, n, m.buf_ptr(gen * n), m.buf_ptr(ka), g1a, 0))
    m.check(lib.b200_g1_mul_batch(cid, n, m.buf_ptr(gen * n), m.buf_ptr(kb), g1b, 0))
    g2a = b"".join(bytes.fromhex(pool[x]["g2"]) for x in j)
    g2b = c.GenG2.Bytes() * n
    expect_unity = np.array([1 if i % 2 == 0 else 0 for i in range(n)], dtype=np.uint8)
    return g1a.raw, g2a, g1b.raw, g2b, expect_unity


def make_inputs_cpu(cid, n_total, n_sample, seed=1):
    """the first n_sample checks of the SAME seeded n_total-check batch as make_inputs, built without a GPU
    (G1 points through the CPU oracle) -- inputs of the reference arm."""
    from oracle import cpu_binding as orc
    from oracle.params import CURVE_IDS
    from oracle import codec
    from oracle.curve import Curve as OC
    P, _ = CURVE_IDS[cid]
    oc = OC(P)
    pool, a, j, k2 = check_scalars(cid, n_total, seed)
    r = ORDERS[cid]
    gen = codec.g1_to_bytes(P, oc.g1)
    ka = b"".join(x.to_bytes(32, "big") for x in a[:n_sample])
    kb = b"".join((r - x).to_bytes(32, "big") for x in k2[:n_sample])
    g1a = orc.g1_mul_batch(cid, n_sample, gen * n_sample, ka)
    g1b = orc.g1_mul_batch(cid, n_sample, gen * n_sample, kb)
    g2a = b"".join(bytes.fromhex(pool[x]["g2"]) for x in j[:n_sample])
    g2b = codec.g2_to_bytes(P, oc.g2) * n_sample
    return g1a, g2a, g1b, g2b


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    for nm, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_rate(cid, inputs, n_sample, budget_s, nthreads=0):
    """Pairing2+FExp checks/s of the CPU restatement (oracle/cpu) on `nthreads` host threads (0 = all)."""
    from oracle import cpu_binding as orc
    c_g1 = 96 if cid != 1 else 64
    c_g2 = 2 * c_g1
    g1a, g2a, g1b, g2b = inputs
    nt = nthreads or orc.threads()
    done, t0 = 0, time.perf_counter()
    off = 0
    total = len(g1a) // c_g1
    while True:
        lo = off % max(1, total - n_sample + 1)
        orc.pairing_batch(cid, n_sample, g1a[lo * c_g1:(lo + n_sample) * c_g1], g2a[lo * c_g2:(lo + n_sample) * c_g2],
                          g1b[lo * c_g1:(lo + n_sample) * c_g1], g2b[lo * c_g2:(lo + n_sample) * c_g2],
                          fexp=True, unity=True, nthreads=nt)
        done += n_sample
        off += n_sample
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, nt, done, dt


_REAL_STDOUT = None
_T0 = time.time()


def log(msg):
    """progress on stderr (the JSON line is the only thing on stdout)"""
    sys.stderr.write("[bench %6.1fs rank %s] %s\n" % (time.time() - _T0, os.environ.get("RANK", "0"), msg))
    sys.stderr.flush()


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun warnings) may write to fd 1; the contract is ONE JSON line on stdout.
    Everything written to fd 1 from here on goes to stderr; emit() writes the JSON line to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def msm_work_m(n, bits=255):
    """Algorithmic Fp products of one Pippenger MSM (the plan of msm.cuh msm_plan without GLV: c = 16 from 2^17 points on):
    n*W mixed additions (10 m) + the running-sum reduction of W * 2^(c-1) buckets (2 full additions of 14 m each).  The
    GLV plan the library uses for one-shot BLS12 MSMs does the same number of bucket additions (2n points x W/2 windows)
    over the same number of buckets, so this stays the work figure of the roofline."""
    lg = n.bit_length() - 1
    c = (max(lg + 1, 7) if lg <= 13 else {14: 15, 15: 14, 16: 15}.get(lg, 16))
    w = bits // c + 1
    return n * w * 10 + w * (1 << (c - 1)) * 2 * 14, c, w


def run_reference_arm(args):
    """CPU arm: the reference's algorithm on the host cores, same metric / unit / config as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_binding as orc
    n = args.batch
    per_step = min(REF_SAMPLE, n)
    cols = make_inputs_cpu(CID, n, per_step, seed=1)
    nt = orc.threads()
    for _ in range(args.warmup):
        orc.pairing_batch(CID, per_step, cols[0], cols[1], cols[2], cols[3], fexp=True, unity=True, nthreads=nt)
    t0 = time.perf_counter()
    verdicts = b""
    for _ in range(args.steps):
        verdicts = orc.pairing_batch(CID, per_step, cols[0], cols[1], cols[2], cols[3], fexp=True, unity=True, nthreads=nt)
    dt = time.perf_counter() - t0
    if verdicts != bytes(1 - (i & 1) for i in range(per_step)):
        raise SystemExit("reference arm: CPU verdicts are wrong")
    value = 2.0 * per_step * args.steps / dt
    sample = ("%d steps x %d checks (checks [0,%d) of the same seeded %d-check batch) on %d threads; oracle/cpu C++ "
              "restatement (NOT gnark/kilic assembly: no Go toolchain in this image)" % (args.steps, per_step, per_step, n, nt))
    extra = {}
    if not args.no_extra:
        # the MSM half of the metric on the host cores: configs[2] at full size, all threads (window-parallel Pippenger)
        try:
            import numpy as np
            nm = 1 << 20
            rng = np.random.default_rng(5)
            ks = scalars_mod_r(rng, nm, 5)
            from oracle.params import CURVE_IDS
            from oracle import codec
            from oracle.curve import Curve as OC
            P, _ = CURVE_IDS[5]
            gen = codec.g1_to_bytes(P, OC(P).g1)
            nd = 1 << 14           # distinct points, tiled to 2^20 (CPU timing does not depend on the point values)
            pts = orc.g1_mul_batch(5, nd, gen * nd, scalars_mod_r(rng, nd, 5).tobytes()) * (nm // nd)
            t0 = time.perf_counter()
            orc.g1_msm(5, nm, pts, ks.tobytes())
            t_msm = time.perf_counter() - t0
            extra["config2_bls12_381_g1_msm_2^20_cpu"] = {
                "latency_ms": t_msm * 1e3, "cores": nt, "kind": "port",
                "note": "oracle/cpu Pippenger, one window per worker thread (mirrors ecc.MultiExpConfig{} all cores); "
                        "2^14 distinct random points tiled to 2^20, scalars uniform in [0, r)"}
        except Exception as ex:
            extra = {"error": repr(ex)}
    line = {
        "impl": "reference", "metric": "bls12_381_pairings_per_s", "value": value, "unit": "pairings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (CPU)", "data": "synthetic",
        "config": headline_config(n, max(1, args.gpus)),
        "pairings_per_s_per_host_thread": value / nt,
        "cpu_baseline": {"value": value, "unit": "pairings/s", "cores": nt, "kind": "port", "sample": sample,
                         "per_thread": value / nt},
        "e2e": {"value": value, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "extra": extra,
    }
    emit(line)


def scalars_mod_r(rng, n, cid):
    """n scalars uniform in [0, r) as an (n, 32) big-endian byte array (vectorised rejection sampling)."""
    import numpy as np
    r = ORDERS[cid]
    top_mask = (1 << (r.bit_length() - 248)) - 1
    rw = np.frombuffer(r.to_bytes(32, "big"), dtype=">u8")
    out = np.empty((n, 32), dtype=np.uint8)
    filled = 0
    while filled < n:
        cnt = int((n - filled) * 1.9) + 64
        cand = rng.integers(0, 256, size=(cnt, 32), dtype=np.uint8)
        cand[:, 0] &= top_mask
        w = cand.view(">u8")
        lt = np.zeros(cnt, dtype=bool)
        eq = np.ones(cnt, dtype=bool)
        for j in range(4):
            lt |= eq & (w[:, j] < rw[j])
            eq &= w[:, j] == rw[j]
        good = cand[lt]
        take = min(len(good), n - filled)
        out[filled:filled + take] = good[:take]
        filled += take
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import mathlib_b200 as m

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mathlib_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = m.load()
    m.check(lib.b200_init(0))
    m.check(lib.b200_set_device(local_rank))
    stream = torch.cuda.current_stream()
    m.check(lib.b200_set_stream(stream.cuda_stream))
    c = m.Curves[CID]
    n = args.batch
    warmup = max(3, args.warmup)

    # ---- inputs (each rank its own batch: weak scaling, no data-path collective)
    log("library loaded, making %d checks" % n)
    g1a, g2a, g1b, g2b, expect = make_inputs(m, CID, n, seed=1 + rank)
    host = [torch.frombuffer(bytearray(x), dtype=torch.uint8).pin_memory() for x in (g1a, g2a, g1b, g2b)]
    d_in = [h.to(dev) for h in host]
    d_out = torch.empty(n * c.GtByteSize, dtype=torch.uint8, device=dev)
    h_out = torch.empty(n * c.GtByteSize, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def step_device():
        m.check(lib.b200_pairing2_batch(CID, n, d_in[0].data_ptr(), d_in[1].data_ptr(), d_in[2].data_ptr(),
                                        d_in[3].data_ptr(), d_out.data_ptr(), m.DEVICE_PTRS | m.FEXP))

    def step_e2e():
        m.check(lib.b200_pairing2_batch(CID, n, host[0].data_ptr(), host[1].data_ptr(), host[2].data_ptr(),
                                        host[3].data_ptr(), h_out.data_ptr(), m.FEXP))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness of the timed path before timing it: valid checks are 1, random ones are not
    log("inputs ready, first launch")
    step_device()
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().reshape(n, c.GtByteSize)
    one = np.frombuffer(c._gt_one, dtype=np.uint8)
    is_one = (got == one).all(axis=1).astype(np.uint8)
    if not (is_one == expect).all():
        raise SystemExit("bench: Pairing2+FExp verdicts are wrong -- refusing to time a broken kernel")

    log("verdicts ok, timing")
    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.b200_launch_count()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)                                   # evict inputs from L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    launches = lib.b200_launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop()

    # ---- e2e through the C ABI with host buffers
    log("device-resident: %.2f ms per step; e2e" % (dev_ms / args.steps))
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if not (h_out.numpy().reshape(n, c.GtByteSize) == got).all():
        raise SystemExit("bench: host-buffer path and device-pointer path disagree")

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    value = 2.0 * n * args.steps * world / (dev_ms * 1e-3)
    e2e_value = 2.0 * n * args.steps * world / (e2e_ms * 1e-3)
    del d_in, d_out, flush

    # ---- the BASELINE configs, at every N (collective ops inside: every rank takes part)
    extra = {}
    log("e2e %.2f ms per step; BASELINE configs" % (e2e_ms / args.steps))
    if not args.no_extra:
        try:
            extra = run_configs(m, lib, dev, stream, torch, dist, rank, world, args)
        except Exception as ex:
            if dist is not None:
                raise
            extra = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    m_mult = M_PER_OP["bls381_pairing2_fexp"] - FERMAT_M[12]
    mac_per_check = m_mult * MAC_PER_M[12]
    achieved = (n / (dev_ms / args.steps * 1e-3)) * mac_per_check          # this rank's kernel, MAC32/s
    peak = IMAD_WIDE_PEAK
    try:
        with open(os.path.join(ROOT, "profiles", "peaks_r1.json")) as f:
            peak = float(json.load(f).get("imad_wide_realistic_per_s", peak))
    except Exception:
        pass
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r2_pairing_vm_ncu.json")) as f:
            ncu = json.load(f)
    except Exception:
        ncu = {"sm__pipe_fmaheavy_cycles_active_pct": 59.2, "smsp__issue_active_pct": 37.5,
               "source": "profiles/r1_pairing_vm_ncu.md (round-1 kernel)"}
    traffic_per_check = float(ncu.get("dram_bytes_per_check", 39.7e6 / 65536))
    roofline = {
        "bound": "imad", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "GMAC32/s (IMAD.WIDE.U32)",
        "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one vm_pairing_kernel<BLS381,2> launch, scaled to this batch
        # size from the committed ncu capture; the algorithmic HBM bytes are 1,152 B per check
        "traffic": traffic_per_check * n,
        "traffic_detail": {"algorithmic_bytes_per_launch": 1152 * n,
                           "note": "inputs are read once after the L2 flush; the outputs are still in L2 when the kernel ends",
                           "source": ncu.get("source", "")},
        "kernel": "vm_pairing_kernel<BLS381,2> (warp-cooperative VM: Miller loop x2 + final exponentiation fused)",
        "m_per_check": m_mult, "mac32_per_m": MAC_PER_M[12],
        # from the committed ncu capture of this kernel (not measured in this run): north_star's roofline evidence
        "ncu": ncu,
        "note": "integer-multiply roofline (SURVEY 8d): algorithmic MAC32 = checks/s x %d m x 300 (oracle count %d m minus "
                "the %d m Fermat inversion, which the GPU does on the ALU pipe); peak = measured IMAD.WIDE.U32 issue "
                "rate (tools/imad_peak.cu, profiles/peaks_r1.json; MEASURED_PEAKS.json holds no integer peak). HBM "
                "traffic is 1,152 B per check (<0.01%% of HBM bandwidth), so no HBM roofline applies."
                % (m_mult, M_PER_OP["bls381_pairing2_fexp"], FERMAT_M[12]),
    }

    # ---- CPU baseline (oracle/cpu port, all host threads, bounded sample)
    log("CPU baseline")
    cpu = None
    try:
        rate, nt, done, dt = cpu_reference_rate(CID, (g1a, g2a, g1b, g2b), 256, args.cpu_budget)
        cpu = {"value": 2.0 * rate, "unit": "pairings/s", "cores": nt, "kind": "port", "per_thread": 2.0 * rate / nt,
               "sample": "%d Pairing2+FExp checks of the same batch in %.1f s on %d threads; oracle/cpu C++ restatement "
                         "(NOT gnark/kilic assembly: no Go toolchain in this image)" % (done, dt, nt)}
    except Exception as ex:  # the checker is optional for the number
        cpu = {"value": None, "unit": "pairings/s", "cores": 0, "kind": "port", "sample": "unavailable: %r" % (ex,)}

    line = {
        "metric": "bls12_381_pairings_per_s", "value": value, "unit": "pairings/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (integer Montgomery arithmetic)", "data": "synthetic",
        "config": headline_config(n, world),
        "checks_per_s": value / 2, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairings/s", "h2d_bytes_per_step": sum(h.numel() for h in host),
                "d2h_bytes_per_step": h_out.numel(), "ms_per_step": e2e_ms / args.steps,
                "api": "b200_pairing2_batch(host buffers, B200_FEXP)"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


# =====================================================================================================================
# BASELINE configs[0..4] and the neighbouring rows, run at every N
# =====================================================================================================================
def run_configs(m, lib, dev, stream, torch, dist, rank, world, args):
    import ctypes
    import numpy as np
    out = {}

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def timed(fn, reps=3):
        """device time of fn (CUDA events on the launch stream), barrier on both sides, max over ranks, best of reps"""
        fn()
        best = 1e30
        for _ in range(reps):
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            sync_all()
            best = min(best, max_over_ranks(e0.elapsed_time(e1)))
        return best

    def wall(fn, reps=3):
        """host wall-clock of fn between barriers (host buffers in, host result out), max over ranks, best of reps"""
        fn()
        best = 1e30
        for _ in range(reps):
            sync_all()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, max_over_ranks((time.perf_counter() - t0) * 1e3))
        return best

    def dev_bytes(b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)

    from mathlib_b200 import shard

    # ---------------------------------------------------------------------------------------------- configs[0]
    log("configs[0]")
    # exactly 1,024 BLS12_381 (kilic) Pairing2+FExp checks per GPU
    c = m.Curves[3]
    n0 = 1024
    ins = make_inputs(m, 3, n0, seed=77 + rank)
    d = [dev_bytes(x) for x in ins[:4]]
    hb = [torch.frombuffer(bytearray(x), dtype=torch.uint8).pin_memory() for x in ins[:4]]
    o = torch.empty(n0 * c.GtByteSize, dtype=torch.uint8, device=dev)
    ho = torch.empty(n0 * c.GtByteSize, dtype=torch.uint8).pin_memory()
    ms = timed(lambda: m.check(lib.b200_pairing2_batch(3, n0, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(),
                                                       d[3].data_ptr(), o.data_ptr(), m.DEVICE_PTRS | m.FEXP)), reps=5)
    ms_e = wall(lambda: m.check(lib.b200_pairing2_batch(3, n0, hb[0].data_ptr(), hb[1].data_ptr(), hb[2].data_ptr(),
                                                        hb[3].data_ptr(), ho.data_ptr(), m.FEXP)), reps=5)
    got = o.cpu().numpy().reshape(n0, c.GtByteSize)
    okv = (got == np.frombuffer(c._gt_one, dtype=np.uint8)).all(axis=1).astype(np.uint8)
    if not (okv == ins[4]).all() or not (ho.numpy().reshape(n0, -1) == got).all():
        raise RuntimeError("configs[0]: verdicts wrong or host path differs")
    out["config0_bls12_381_kilic_1024_pairing2_fexp"] = {
        "ms": ms, "checks_per_s": n0 * world / ms * 1e3, "pairings_per_s": 2 * n0 * world / ms * 1e3,
        "e2e": {"ms": ms_e, "pairings_per_s": 2 * n0 * world / ms_e * 1e3, "h2d_bytes_per_step": sum(h.numel() for h in hb),
                "d2h_bytes_per_step": ho.numel(), "api": "b200_pairing2_batch(host buffers)"},
        "imad_frac": n0 / ms * 1e3 * (M_PER_OP["bls381_pairing2_fexp"] - FERMAT_M[12]) * MAC_PER_M[12] / IMAD_WIDE_PEAK,
        "what": "BASELINE configs[0] verbatim, %d checks per GPU x %d GPUs" % (n0, world)}
    del d, o

    # ---------------------------------------------------------------------------------------------- configs[1]
    log("configs[1]")
    if world == 1:
        c = m.Curves[1]
        n1 = 65536
        ins = make_inputs(m, 1, n1, seed=78)
        d = [dev_bytes(x) for x in ins[:2]]
        o = torch.empty(n1 * c.GtByteSize, dtype=torch.uint8, device=dev)
        ms = timed(lambda: m.check(lib.b200_pairing_batch(1, n1, d[0].data_ptr(), d[1].data_ptr(), o.data_ptr(),
                                                          m.DEVICE_PTRS | m.FEXP)), reps=2)
        rate = n1 / ms * 1e3
        out["config1_bn254_65536_pairing_fexp"] = {
            "ms": ms, "pairings_per_s": rate,
            "imad_frac": rate * (M_PER_OP["bn254_pairing_fexp"] - FERMAT_M[8]) * MAC_PER_M[8] / IMAD_WIDE_PEAK}
        del d, o

    # ---------------------------------------------------------------------------------------------- configs[2]
    log("configs[2]")
    # BLS12_381_GURVY G1 MultiScalarMul, 2^20 points / scalars uniform in [0, r), STRONG-scaled: rank k owns the point
    # range [n k / N, n (k+1) / N); one affine partial sum per rank, 96-byte all-gather, b200_g1_sum.
    cid = 5
    c = m.Curves[cid]
    n2 = 1 << 20
    rng = np.random.default_rng(5)
    ks_pts = scalars_mod_r(rng, n2, cid)
    ks = scalars_mod_r(rng, n2, cid)
    lo, hi = n2 * rank // world, n2 * (rank + 1) // world
    nl = hi - lo
    g1sz = c.G1ByteSize
    gen = dev_bytes(c.GenG1.Bytes()).repeat(n2)
    d_kp = torch.from_numpy(ks_pts.reshape(-1)).to(dev)
    pts_bytes = torch.empty(n2 * g1sz, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(cid, n2, gen.data_ptr(), d_kp.data_ptr(), pts_bytes.data_ptr(), m.DEVICE_PTRS))
    torch.cuda.synchronize()
    pts_mont = torch.empty(nl * g1sz, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(cid, nl, gen.data_ptr(), d_kp[lo * 32:].data_ptr(), pts_mont.data_ptr(),
                                  m.DEVICE_PTRS | m.OUT_MONT))
    d_k = torch.from_numpy(ks[lo:hi].reshape(-1)).to(dev)
    d_part = torch.empty(g1sz, dtype=torch.uint8, device=dev)
    d_gath = torch.empty(world * g1sz, dtype=torch.uint8, device=dev)
    d_res = torch.empty(g1sz, dtype=torch.uint8, device=dev)

    scr = (d_part, d_gath, d_res)

    def msm_dev():
        # local MSM -> affine partial (MONT limbs) -> NCCL all-gather -> b200_g1_sum -> Bytes(), all on the stream
        shard.msm_sharded_device(lib, cid, pts_mont, d_k, nl, dist, dev, in_flags=m.IN_MONT, scratch=scr)
    ms = timed(msm_dev, reps=5)
    want = d_res.cpu().numpy().tobytes()
    # only this rank's local MSM, without the combine (what the all-gather + sum add on top)
    ms_local = timed(lambda: m.check(lib.b200_g1_msm(cid, nl, pts_mont.data_ptr(), d_k.data_ptr(), d_part.data_ptr(),
                                                     m.DEVICE_PTRS | m.IN_MONT | m.OUT_MONT)), reps=3)
    # e2e: this rank's points (Bytes() form) and scalars in pinned HOST memory, result back on the host
    h_pts = pts_bytes[lo * g1sz:hi * g1sz].cpu().pin_memory()
    h_k = torch.from_numpy(ks[lo:hi].reshape(-1).copy()).pin_memory()
    h_part = torch.empty(g1sz, dtype=torch.uint8).pin_memory()

    def msm_e2e():
        m.check(lib.b200_g1_msm(cid, nl, h_pts.data_ptr(), h_k.data_ptr(), h_part.data_ptr(), m.OUT_MONT))
        d_part.copy_(h_part, non_blocking=True)
        if dist is not None:
            dist.all_gather_into_tensor(d_gath, d_part)
        m.check(lib.b200_g1_sum(cid, world, (d_gath if dist is not None else d_part).data_ptr(), d_res.data_ptr(),
                                m.DEVICE_PTRS | m.IN_MONT))
        return d_res.cpu()
    ms_e = wall(msm_e2e, reps=3)
    if msm_e2e().numpy().tobytes() != want:
        raise RuntimeError("configs[2]: host-buffer MSM disagrees with the device-resident one")
    work_m, cwin, nwin = msm_work_m(nl)
    res = {"latency_ms": ms, "local_msm_ms": ms_local, "points_per_s": n2 / ms * 1e3,
           "points_per_rank": nl, "scaling": "strong", "n_gpus": world,
           "e2e": {"latency_ms": ms_e, "h2d_bytes_per_rank": nl * (g1sz + 32), "d2h_bytes_per_rank": g1sz,
                   "api": "b200_g1_msm(host buffers) per rank + NCCL all-gather + b200_g1_sum"},
           "roofline": {"bound": "imad", "unit": "GMAC32/s (IMAD.WIDE.U32)",
                        "achieved": work_m * MAC_PER_M[12] / (ms_local * 1e-3) / 1e9, "peak": IMAD_WIDE_PEAK / 1e9,
                        "frac": work_m * MAC_PER_M[12] / (ms_local * 1e-3) / IMAD_WIDE_PEAK,
                        "m_per_launch": work_m, "window_bits": cwin, "windows": nwin,
                        "hbm_GBps_algorithmic": nl * (g1sz + 32) / (ms_local * 1e-3) / 1e9,
                        "note": "per rank: n*W mixed adds x 10 m + 2*W*2^(c-1) adds x 14 m (c-bit windows over the full "
                                "scalar; the GLV plan in use does the same additions over 2n points and W/2 windows), 300 "
                                "MAC32 per m, over this rank's local MSM time; HBM bytes n*(2*FpBytes+32) are a secondary "
                                "figure"},
           "what": "points (Montgomery slab) and scalars already in HBM; scalars uniform in [0, r)"}
    # oracle check + CPU baseline of the same 2^20-point problem (rank 0, all host threads)
    if rank == 0:
        try:
            from oracle import cpu_binding as orc
            hp = pts_bytes.cpu().numpy().tobytes()
            t0 = time.perf_counter()
            ref = orc.g1_msm(cid, n2, hp, ks.tobytes())
            t_cpu = time.perf_counter() - t0
            if want != ref:
                raise RuntimeError("configs[2]: GPU MSM result differs from the CPU oracle at 2^20")
            res["checked_against_oracle"] = True
            nt = orc.threads()
            res["cpu_baseline"] = {"latency_ms": t_cpu * 1e3, "cores": nt, "kind": "port",
                                   "sample": "the same 2^20 points/scalars, oracle/cpu Pippenger with one window per "
                                             "worker thread on %d threads (mirrors ecc.MultiExpConfig{} all cores)" % nt}
            if world == 1:
                n1t = 1 << 17
                t0 = time.perf_counter()
                orc.g1_msm(cid, n1t, hp[:n1t * g1sz], ks[:n1t].tobytes(), nthreads=1)
                res["cpu_baseline"]["one_thread_2^17_points_ms"] = (time.perf_counter() - t0) * 1e3
        except RuntimeError:
            raise
        except Exception as ex:
            res["cpu_baseline"] = {"error": repr(ex)}
    if world == 1:
        # resident bases with per-window tables, scalars uploaded from pinned host memory
        h = ctypes.c_uint64()
        t0 = time.perf_counter()
        m.check(lib.b200_bases_upload(cid, n2, pts_mont.data_ptr(), m.DEVICE_PTRS | m.IN_MONT | m.BASES_TABLES, ctypes.byref(h)))
        torch.cuda.synchronize()
        res["tables_upload_ms"] = (time.perf_counter() - t0) * 1e3
        o = torch.empty(g1sz, dtype=torch.uint8, device=dev)
        res["resident_tables_latency_ms"] = timed(lambda: m.check(lib.b200_g1_msm_resident(
            h.value, n2, d_k.data_ptr(), o.data_ptr(), m.DEVICE_PTRS)))
        if o.cpu().numpy().tobytes() != want:
            raise RuntimeError("window-table MSM disagrees with the one-shot MSM")
        h_o = ctypes.create_string_buffer(g1sz)
        res["resident_tables_scalars_from_host_ms"] = wall(
            lambda: m.check(lib.b200_g1_msm_resident(h.value, n2, h_k.data_ptr(), h_o, 0)))
        if h_o.raw != want:
            raise RuntimeError("host-scalar resident MSM disagrees")
        m.check(lib.b200_bases_free(h.value))
    out["config2_bls12_381_g1_msm_2^20"] = res
    del gen, d_kp, pts_bytes, pts_mont, d_k, h_pts

    # ---------------------------------------------------------------------------------------------- configs[3]
    log("configs[3]")
    # BLS12_377_GURVY MSM, 2^21 points per rank (2^24 over 8 GPUs) + partial-sum combine: WEAK scaling
    cid = 4
    c = m.Curves[cid]
    n3 = 1 << 21
    g1sz = c.G1ByteSize
    rng3 = np.random.default_rng(1000 + rank)
    d_kp = torch.from_numpy(scalars_mod_r(rng3, n3, cid).reshape(-1)).to(dev)
    d_k = torch.from_numpy(scalars_mod_r(rng3, n3, cid).reshape(-1)).to(dev)
    gen = dev_bytes(c.GenG1.Bytes()).repeat(n3)
    pts = torch.empty(n3 * g1sz, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(cid, n3, gen.data_ptr(), d_kp.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS | m.OUT_MONT))
    d_part = torch.empty(g1sz, dtype=torch.uint8, device=dev)
    d_gath = torch.empty(world * g1sz, dtype=torch.uint8, device=dev)
    d_res = torch.empty(g1sz, dtype=torch.uint8, device=dev)

    scr3 = (d_part, d_gath, d_res)

    def msm3():
        shard.msm_sharded_device(lib, cid, pts, d_k, n3, dist, dev, in_flags=m.IN_MONT, out_flags=m.OUT_MONT, scratch=scr3)
    ms = timed(msm3, reps=2)
    # size-independent check of the exchange step: the combined point equals the sum of the gathered partials formed
    # one by one (b200_g1_sum over pairs), and rank 0's partial equals sum_i [k_i * kp_i] G (one Mul of the generator)
    parts = d_gath.cpu().numpy().tobytes() if dist is not None else d_part.cpu().numpy().tobytes()
    acc = ctypes.create_string_buffer(parts[:g1sz], g1sz)
    for k in range(1, world):
        pair = acc.raw + parts[k * g1sz:(k + 1) * g1sz]
        m.check(lib.b200_g1_sum(cid, 2, m.buf_ptr(pair), acc, m.IN_MONT | m.OUT_MONT))
    if acc.raw != d_res.cpu().numpy().tobytes():
        raise RuntimeError("configs[3]: combine disagrees with pairwise sums")
    r = ORDERS[cid]
    kp = np.frombuffer(d_kp.cpu().numpy().tobytes(), dtype=np.uint8).reshape(n3, 32)
    kk = np.frombuffer(d_k.cpu().numpy().tobytes(), dtype=np.uint8).reshape(n3, 32)
    s = 0
    step = 1 << 14
    for lo3 in range(0, n3, step):            # sum k_i * kp_i mod r on Python ints (exact)
        a = [int.from_bytes(x.tobytes(), "big") for x in kp[lo3:lo3 + step]]
        b = [int.from_bytes(x.tobytes(), "big") for x in kk[lo3:lo3 + step]]
        s = (s + sum(x * y for x, y in zip(a, b))) % r
    expect = ctypes.create_string_buffer(g1sz)
    m.check(lib.b200_g1_mul_batch(cid, 1, m.buf_ptr(c.GenG1.Bytes()), m.buf_ptr(s.to_bytes(32, "big")), expect, m.OUT_MONT))
    if expect.raw != d_part.cpu().numpy().tobytes():
        raise RuntimeError("configs[3]: rank %d partial MSM != [sum k_i kp_i]G" % rank)
    work_m, cwin, nwin = msm_work_m(n3, 253)
    out["config3_bls12_377_g1_msm_2^21_per_gpu"] = {
        "latency_ms": ms, "points_total": n3 * world, "points_per_s": n3 * world / ms * 1e3, "scaling": "weak",
        "n_gpus": world, "imad_frac_per_gpu": work_m * MAC_PER_M[12] / (ms * 1e-3) / IMAD_WIDE_PEAK,
        "checked": "partial == [sum k_i kp_i mod r]G on every rank; combine == pairwise sums",
        "what": "2^21 points per rank (2^24 at 8 GPUs), NCCL all-gather of the partials + b200_g1_sum"}
    del gen, pts, d_kp, d_k

    # ---------------------------------------------------------------------------------------------- configs[4]
    log("configs[4]")
    # BLS12_381_BBS: 12,500 BBS-style verifications per rank (100k at 8 GPUs) = Mul2 (B = [e]G1 + [f]A) feeding
    # Pairing2+FExp -> IsUnity on the device, G2 arguments fixed (public key, generator); every second one is valid
    import random
    cid = 6
    c = m.Curves[cid]
    n4 = 12500
    with open(os.path.join(ROOT, "tests", "golden", "g2_pool.json")) as f:
        pk = json.load(f)["3"][0]
    rnd = random.Random(11 + rank)
    r, bj = c.order, int(pk["b"], 16)
    a = [rnd.randrange(1, r) for _ in range(n4)]
    fs = [rnd.randrange(1, r) for _ in range(n4)]
    es = [(-a[i] * bj - fs[i] * a[i]) % r if i % 2 == 0 else rnd.randrange(1, r) for i in range(n4)]
    gen = dev_bytes(c.GenG1.Bytes() * n4)
    d_a = dev_bytes(b"".join(x.to_bytes(32, "big") for x in a))
    h_e = torch.frombuffer(bytearray(b"".join(x.to_bytes(32, "big") for x in es)), dtype=torch.uint8).pin_memory()
    h_f = torch.frombuffer(bytearray(b"".join(x.to_bytes(32, "big") for x in fs)), dtype=torch.uint8).pin_memory()
    d_e, d_f = h_e.to(dev), h_f.to(dev)
    d_pk = dev_bytes(bytes.fromhex(pk["g2"]) * n4)
    d_g2 = dev_bytes(c.GenG2.Bytes() * n4)
    d_A = torch.empty(n4 * c.G1ByteSize, dtype=torch.uint8, device=dev)
    d_B = torch.empty(n4 * c.G1ByteSize, dtype=torch.uint8, device=dev)
    d_v = torch.empty(n4, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(cid, n4, gen.data_ptr(), d_a.data_ptr(), d_A.data_ptr(), m.DEVICE_PTRS))
    h_A = d_A.cpu().pin_memory()
    expect_v = np.array([1 - (i & 1) for i in range(n4)], dtype=np.uint8)

    def verify():
        m.check(lib.b200_g1_mul2_batch(cid, n4, gen.data_ptr(), d_e.data_ptr(), d_A.data_ptr(), d_f.data_ptr(),
                                       d_B.data_ptr(), m.DEVICE_PTRS))
        m.check(lib.b200_pairing2_batch(cid, n4, d_A.data_ptr(), d_pk.data_ptr(), d_B.data_ptr(), d_g2.data_ptr(),
                                        d_v.data_ptr(), m.DEVICE_PTRS | m.FEXP | m.OUT_UNITY_ONLY))
    ms = timed(verify)
    if not (d_v.cpu().numpy() == expect_v).all():
        raise RuntimeError("BBS-style verification verdicts are wrong")
    out["config4_bls12_381_bbs_verify_12500_per_gpu"] = {
        "ms": ms, "verifications_per_s": n4 * world / ms * 1e3, "verifications_total": n4 * world, "scaling": "weak",
        "n_gpus": world, "what": "Mul2 -> Pairing2+FExp -> IsUnity, device resident, general (non-fixed) G2 arguments"}
    # the same verifications with the two fixed G2 arguments (public key, generator) as a resident line table (8f-1)
    hl = ctypes.c_uint64()
    m.check(lib.b200_g2_lines_upload(cid, 2, m.buf_ptr(bytes.fromhex(pk["g2"]) + c.GenG2.Bytes()), 0, ctypes.byref(hl)))

    def verify_fixed():
        m.check(lib.b200_g1_mul2_batch(cid, n4, gen.data_ptr(), d_e.data_ptr(), d_A.data_ptr(), d_f.data_ptr(),
                                       d_B.data_ptr(), m.DEVICE_PTRS))
        m.check(lib.b200_pairing2_fixed_batch(hl.value, n4, d_A.data_ptr(), None, d_B.data_ptr(), None, d_v.data_ptr(),
                                              m.DEVICE_PTRS | m.FEXP | m.OUT_UNITY_ONLY))
    d_v.zero_()
    ms = timed(verify_fixed)
    if not (d_v.cpu().numpy() == expect_v).all():
        raise RuntimeError("fixed-Q verification verdicts differ")
    # e2e: signature points A and scalars e, f come from pinned host memory every step, verdict bytes go back
    h_v = torch.empty(n4, dtype=torch.uint8).pin_memory()

    def verify_e2e():
        d_A.copy_(h_A, non_blocking=True)
        d_e.copy_(h_e, non_blocking=True)
        d_f.copy_(h_f, non_blocking=True)
        verify_fixed()
        h_v.copy_(d_v, non_blocking=True)
    ms_e = wall(verify_e2e)
    if not (h_v.numpy() == expect_v).all():
        raise RuntimeError("fixed-Q e2e verdicts differ")
    out["config4_bls12_381_bbs_verify_12500_fixed_q_tables"] = {
        "ms": ms, "verifications_per_s": n4 * world / ms * 1e3, "verifications_total": n4 * world, "scaling": "weak",
        "n_gpus": world,
        "e2e": {"ms": ms_e, "verifications_per_s": n4 * world / ms_e * 1e3,
                "h2d_bytes_per_step": h_A.numel() + h_e.numel() + h_f.numel(), "d2h_bytes_per_step": n4}}
    if world > 1:
        m.check(lib.b200_g2_lines_free(hl.value))
        return out

    # ---------------------------------------------------------------------------------------------- neighbours (N = 1)
    # fixed-Q Pairing2+FExp at the headline batch size (65,536 checks, same two G2 rows for every check)
    rng = np.random.default_rng(9)
    n5 = 65536
    gen2 = dev_bytes(c.GenG1.Bytes() * n5)
    d_kk = torch.from_numpy(scalars_mod_r(rng, n5, cid).reshape(-1)).to(dev)
    d_P = torch.empty(n5 * c.G1ByteSize, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(cid, n5, gen2.data_ptr(), d_kk.data_ptr(), d_P.data_ptr(), m.DEVICE_PTRS))
    d_o = torch.empty(n5 * c.GtByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_pairing2_fixed_batch(hl.value, n5, d_P.data_ptr(), None, d_P.data_ptr(), None,
                                                             d_o.data_ptr(), m.DEVICE_PTRS | m.FEXP)), reps=2)
    out["next_pairing2_fexp_fixed_q_bls12_381_65536"] = {"ms": ms, "pairings_per_s": 2 * n5 / ms * 1e3}
    m.check(lib.b200_g2_lines_free(hl.value))
    del gen2, d_kk, d_P, d_o
    # SURVEY 8(f) row 3, the callers next to the hot path: Gt.Exp and G2.Mul batches on BLS12-381 (device resident)
    c = m.Curves[5]
    n6 = 16384
    gt = dev_bytes(c.GenGt.Bytes() * n6)
    d_kr = torch.from_numpy(scalars_mod_r(rng, n6, 5).reshape(-1)).to(dev)
    o = torch.empty(n6 * c.GtByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_gt_exp_batch(5, n6, gt.data_ptr(), d_kr.data_ptr(), o.data_ptr(), m.DEVICE_PTRS)),
               reps=2)
    out["next_gt_exp_bls12_381_16384"] = {"ms": ms, "exps_per_s": n6 / ms * 1e3}
    g2 = dev_bytes(c.GenG2.Bytes() * n6)
    o2 = torch.empty(n6 * c.G2ByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_g2_mul_batch(5, n6, g2.data_ptr(), d_kr.data_ptr(), o2.data_ptr(), m.DEVICE_PTRS)),
               reps=2)
    out["next_g2_mul_bls12_381_16384"] = {"ms": ms, "muls_per_s": n6 / ms * 1e3}
    # G2 MSM over the n6 points just produced (distinct multiples of the generator), scalars uniform in [0, r)
    o3 = torch.empty(c.G2ByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_g2_msm(5, n6, o2.data_ptr(), d_kr.data_ptr(), o3.data_ptr(), m.DEVICE_PTRS)), reps=2)
    out["next_g2_msm_bls12_381_16384"] = {"latency_ms": ms, "points_per_s": n6 / ms * 1e3}
    return out


if __name__ == "__main__":
    main()
