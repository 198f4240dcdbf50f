#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 backend for IBM/mathlib's pairing / G1 hot path.

Contract (see task statement):  python bench.py --gpus N --steps K --warmup W [--impl reference]
prints ONE JSON line on rank 0.

Workload (BASELINE.json metric "BLS12-381 pairings/s ..."): one step = one batch of 65,536 independent
BLS12-381 Pairing2+FExp checks (the op of configs[0] / perf_test.go:541-560 at the batch size of configs[1],
because 1,024 checks cannot fill 148 SMs).  Half of the batch are valid BLS-style checks (product == 1), half
random.  `value` counts pairings (2 per check) with inputs resident in HBM; `e2e` is the same metric through
the C ABI with HOST buffers (H2D + kernel + D2H inside the timed region).  Secondary configs (configs[0] exact,
configs[1] BN254, configs[2] MSM 2^20) are reported under "extra".

--impl reference times the reference's CPU algorithm for the same check.  The reference's Go drivers cannot be
built here (no Go toolchain), so this arm runs oracle/cpu (the C++ restatement, kind "port") on all host threads.
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
# Algorithmic work per op in Fp Montgomery products `m`, counted by instrumenting the optimised CPU oracle path
# (oracle/cpu, orc_mul_count; includes the Fermat inversion of the easy part) -- see DESIGN.md "work units".
M_PER_OP = {"bls381_pairing2_fexp": 19825, "bls381_pairing_fexp": 15198, "bn254_pairing_fexp": 17496}
MAC_PER_M = {12: 300, 8: 136}            # 2n^2 + n 32x32->64 multiply-accumulates per Montgomery product
# IMAD.WIDE.U32 issue peak of one B200, measured with tools/imad_variants.cu (profiles/peaks_r1.json):
# fmaheavy pipe, 4 cycles per warp instruction -> 32 lanes/clk/SM * 148 SM * 1.965 GHz = 9.31e12/s nominal.
IMAD_WIDE_PEAK = 8.98e12


def make_inputs(m, cid, n, seed=1):
    """n Pairing2 checks in the reference's BYTES encoding: half valid (e(aG1,Q)e(-abG1,G2) = 1), half random.
    G1 points are made on the GPU (b200_g1_mul_batch); G2 points come from the committed pool
    tests/golden/g2_pool.json (Q_j = [b_j]G2)."""
    import random
    import numpy as np
    c = m.Curves[cid]
    with open(os.path.join(ROOT, "tests", "golden", "g2_pool.json")) as f:
        pool = json.load(f)[str({3: 3, 5: 3, 6: 3, 7: 3, 1: 1, 4: 4}[cid])]
    rnd = random.Random(seed)
    r = c.order
    a = [rnd.randrange(1, r) for _ in range(n)]
    j = [rnd.randrange(len(pool)) for _ in range(n)]
    b = [int(pool[x]["b"], 16) for x in j]
    k2 = [(a[i] * b[i]) % r if i % 2 == 0 else rnd.randrange(1, r) for i in range(n)]
    gen = c.GenG1.Bytes()
    ka = b"".join(x.to_bytes(32, "big") for x in a)
    kb = b"".join((r - x).to_bytes(32, "big") for x in k2)        # -k2 mod r
    lib = m.load()
    import ctypes
    g1a = ctypes.create_string_buffer(n * c.G1ByteSize)
    g1b = ctypes.create_string_buffer(n * c.G1ByteSize)
    m.check(lib.b200_g1_mul_batch(cid, n, m.buf_ptr(gen * n), m.buf_ptr(ka), g1a, 0))
    m.check(lib.b200_g1_mul_batch(cid, n, m.buf_ptr(gen * n), m.buf_ptr(kb), g1b, 0))
    g2a = b"".join(bytes.fromhex(pool[x]["g2"]) for x in j)
    g2b = c.GenG2.Bytes() * n
    expect_unity = np.array([1 if i % 2 == 0 else 0 for i in range(n)], dtype=np.uint8)
    return g1a.raw, g2a, g1b.raw, g2b, expect_unity


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


def run_reference_arm(args):
    """CPU arm: the reference's algorithm on the host cores, same metric/unit/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_binding as orc
    import random
    # inputs without the GPU: tile the golden vectors' pairing2 cases (timing is value independent)
    with open(os.path.join(ROOT, "tests", "golden", "vectors_3.json")) as f:
        v = json.load(f)
    cases = v["pairing2"][:2]
    per_step = 1024
    cols = [b"".join(bytes.fromhex(cases[i % 2][k]) for i in range(per_step)) for k in ("g1a", "g2a", "g1b", "g2b")]
    nt = orc.threads()
    for _ in range(max(1, args.warmup if args.warmup < 2 else 1)):
        orc.pairing_batch(CID, 64, cols[0][:64 * 96], cols[1][:64 * 192], cols[2][:64 * 96], cols[3][:64 * 192], fexp=True,
                          unity=True, nthreads=nt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.pairing_batch(CID, per_step, cols[0], cols[1], cols[2], cols[3], fexp=True, unity=True, nthreads=nt)
    dt = time.perf_counter() - t0
    value = 2.0 * per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "bls12_381_pairings_per_s", "value": value, "unit": "pairings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (CPU)", "data": "synthetic",
        "config": {"workload": "BLS12_381 (kilic semantics) Pairing2+FExp checks, %d per step (bounded sample of the "
                               "65,536-check workload), 2 pairings per check" % per_step,
                   "curve_id": CID, "checks_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "pairings/s", "cores": nt, "kind": "port",
                         "sample": "%d steps x %d checks on %d threads; oracle/cpu C++ restatement "
                                   "(NOT gnark/kilic assembly: no Go toolchain in this image)" % (args.steps, per_step, nt)},
        "e2e": {"value": value, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
    step_device()
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().reshape(n, c.GtByteSize)
    one = np.frombuffer(c._gt_one, dtype=np.uint8)
    is_one = (got == one).all(axis=1).astype(np.uint8)
    if not (is_one == expect).all():
        raise SystemExit("bench: Pairing2+FExp verdicts are wrong -- refusing to time a broken kernel")

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

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    mac_per_check = M_PER_OP["bls381_pairing2_fexp"] * MAC_PER_M[12]
    achieved = (n / (dev_ms / args.steps * 1e-3)) * mac_per_check          # this rank's kernel, MAC32/s
    peak = IMAD_WIDE_PEAK
    try:
        with open(os.path.join(ROOT, "profiles", "peaks_r1.json")) as f:
            peak = float(json.load(f).get("imad_wide_realistic_per_s", peak))
    except Exception:
        pass
    roofline = {
        "bound": "imad", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "GMAC32/s (IMAD.WIDE.U32)",
        "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one vm_pairing_kernel<BLS381,2> launch at this batch size
        # (ncu capture profiles/r1_bench_launches.md); the algorithmic HBM bytes are 1,152 B per check
        "traffic": 39.7e6 * n / 65536,
        "traffic_detail": {"dram_bytes_per_launch_at_65536_checks": 39.7e6, "algorithmic_bytes_per_launch": 1152 * n,
                           "note": "39.3 MB read (the 37.7 MB of inputs, after the L2 flush) + 0.4 MB written: the 37.7 MB "
                                   "of outputs are still in L2 when the kernel ends",
                           "source": "profiles/r1_bench_launches.md"},
        "kernel": "vm_pairing_kernel<BLS381,2> (warp-cooperative VM: Miller loop x2 + final exponentiation fused)",
        # from the committed ncu capture of this kernel (not measured in this run): north_star's roofline evidence
        "ncu": {"sm__pipe_fmaheavy_cycles_active_pct": 59.2, "smsp__issue_active_pct": 37.5,
                "source": "profiles/r1_pairing_vm_ncu.md",
                "miller_loop_only_sm__pipe_fmaheavy_cycles_active_pct": 62.0,
                "miller_loop_source": "profiles/r1_miller_loop_ncu.md"},
        "note": "integer-multiply roofline (SURVEY 8d): algorithmic MAC32 = checks/s x %d m x 300; peak = measured "
                "IMAD.WIDE.U32 issue rate (tools/imad_peak.cu, profiles/peaks_r1.json). HBM traffic is "
                "1,152 B per check (<0.01%% of HBM bandwidth), so no HBM roofline applies." % M_PER_OP["bls381_pairing2_fexp"],
    }

    # ---- CPU baseline (oracle/cpu port, all host threads, bounded sample)
    cpu = None
    try:
        rate, nt, done, dt = cpu_reference_rate(CID, (g1a, g2a, g1b, g2b), 256, args.cpu_budget)
        cpu = {"value": 2.0 * rate, "unit": "pairings/s", "cores": nt, "kind": "port",
               "sample": "%d Pairing2+FExp checks of the same batch in %.1f s on %d threads; oracle/cpu C++ restatement "
                         "(NOT gnark/kilic assembly: no Go toolchain in this image)" % (done, dt, nt)}
    except Exception as ex:  # the checker is optional for the number
        cpu = {"value": None, "unit": "pairings/s", "cores": 0, "kind": "port", "sample": "unavailable: %r" % (ex,)}

    extra = {}
    if not args.no_extra and world == 1:
        try:
            extra = run_extra(m, lib, dev, stream, torch)
        except Exception as ex:
            extra = {"error": repr(ex)}

    line = {
        "metric": "bls12_381_pairings_per_s", "value": value, "unit": "pairings/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (integer Montgomery arithmetic)", "data": "synthetic",
        "config": {"workload": "BLS12_381 (kilic semantics, mathlib CurveID 3) Pairing2+FExp checks, %d per step per GPU, "
                               "2 pairings per check; half valid BLS checks, half random" % n,
                   "curve_id": CID, "checks_per_step_per_gpu": n, "parallelism": "index-split x%d, no collective" % world,
                   "l2": "256 MiB flush write between timed iterations", "encoding": "reference Bytes() formats"},
        "checks_per_s": value / 2, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairings/s", "h2d_bytes_per_step": sum(h.numel() for h in host),
                "d2h_bytes_per_step": h_out.numel(), "ms_per_step": e2e_ms / args.steps,
                "api": "b200_pairing2_batch(host buffers, B200_FEXP)"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def run_extra(m, lib, dev, stream, torch):
    """Secondary BASELINE.json configs, device-resident, a few iterations each."""
    import numpy as np
    out = {}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    # configs[0]: exactly 1,024 BLS12_381 (kilic) Pairing2+FExp checks
    c = m.Curves[3]
    ins = make_inputs(m, 3, 1024, seed=77)
    d = [torch.frombuffer(bytearray(x), dtype=torch.uint8).to(dev) for x in ins[:4]]
    o = torch.empty(1024 * c.GtByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_pairing2_batch(3, 1024, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(),
                                                       d[3].data_ptr(), o.data_ptr(), m.DEVICE_PTRS | m.FEXP)))
    out["config0_bls12_381_kilic_1024_pairing2_fexp"] = {"ms": ms, "checks_per_s": 1024 / ms * 1e3}

    # configs[1]: BN254 Pairing+FExp, 65,536 pairs
    c = m.Curves[1]
    n = 65536
    ins = make_inputs(m, 1, n, seed=78)
    d = [torch.frombuffer(bytearray(x), dtype=torch.uint8).to(dev) for x in ins[:2]]
    o = torch.empty(n * c.GtByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_pairing_batch(1, n, d[0].data_ptr(), d[1].data_ptr(), o.data_ptr(),
                                                      m.DEVICE_PTRS | m.FEXP)), reps=2)
    rate = n / ms * 1e3
    out["config1_bn254_65536_pairing_fexp"] = {
        "ms": ms, "pairings_per_s": rate,
        "imad_frac": rate * M_PER_OP["bn254_pairing_fexp"] * MAC_PER_M[8] / IMAD_WIDE_PEAK}

    # configs[2]: BLS12_381_GURVY G1 MSM, 2^20 uniform scalars, points [k_i]G1 made on the GPU
    c = m.Curves[5]
    n = 1 << 20
    rng = np.random.default_rng(5)
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks[:, 0] &= 0x3F                                           # < 2^254 < r
    d_k = torch.from_numpy(ks.reshape(-1)).to(dev)
    gen = torch.frombuffer(bytearray(c.GenG1.Bytes()), dtype=torch.uint8).to(dev).repeat(n)
    pts = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(5, n, gen.data_ptr(), d_k.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS | m.OUT_MONT))
    ks2 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks2[:, 0] &= 0x3F
    d_k2 = torch.from_numpy(ks2.reshape(-1)).to(dev)
    o = torch.empty(c.G1ByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_g1_msm(5, n, pts.data_ptr(), d_k2.data_ptr(), o.data_ptr(),
                                               m.DEVICE_PTRS | m.IN_MONT)))
    want = o.cpu().numpy().tobytes()
    res = {"latency_ms": ms, "points_per_s": n / ms * 1e3, "hbm_GBps_algorithmic": n * 128 / (ms * 1e-3) / 1e9,
           "what": "points (Montgomery slab) and scalars already in HBM"}
    # (i) bases resident with per-window tables, scalars uploaded from pinned host memory, result read back
    import ctypes
    h = ctypes.c_uint64()
    t0 = time.perf_counter()
    m.check(lib.b200_bases_upload(5, n, pts.data_ptr(), m.DEVICE_PTRS | m.IN_MONT | m.BASES_TABLES, ctypes.byref(h)))
    torch.cuda.synchronize()
    res["tables_upload_ms"] = (time.perf_counter() - t0) * 1e3
    ms_t = timed(lambda: m.check(lib.b200_g1_msm_resident(h.value, n, d_k2.data_ptr(), o.data_ptr(),
                                                          m.DEVICE_PTRS)))
    if o.cpu().numpy().tobytes() != want:
        raise RuntimeError("window-table MSM disagrees with the one-shot MSM")
    res["resident_tables_latency_ms"] = ms_t
    h_k2 = torch.from_numpy(ks2.reshape(-1)).pin_memory()
    h_o = ctypes.create_string_buffer(c.G1ByteSize)

    def wall(fn, reps=3):
        fn()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best * 1e3
    res["resident_tables_scalars_from_host_ms"] = wall(
        lambda: m.check(lib.b200_g1_msm_resident(h.value, n, h_k2.data_ptr(), h_o, 0)))
    if h_o.raw != want:
        raise RuntimeError("host-scalar resident MSM disagrees")
    m.check(lib.b200_bases_free(h.value))
    # (ii) everything from host: 2^20 points in Bytes() form + scalars cross PCIe inside the timed region
    m.check(lib.b200_g1_mul_batch(5, n, gen.data_ptr(), d_k.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS))
    h_pts = pts.cpu().pin_memory()
    res["all_from_host_ms"] = wall(lambda: m.check(lib.b200_g1_msm(5, n, h_pts.data_ptr(), h_k2.data_ptr(), h_o, 0)), reps=2)
    if h_o.raw != want:
        raise RuntimeError("host-buffer MSM disagrees")
    out["config2_bls12_381_g1_msm_2^20"] = res

    # configs[3], one GPU's share: BLS12_377_GURVY MSM over 2^21 of the 2^24 points (range split, one partial sum per GPU)
    c = m.Curves[4]
    n = 1 << 21
    ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks[:, 0] &= 0x0F                                           # < 2^252 < r
    d_k = torch.from_numpy(ks.reshape(-1)).to(dev)
    gen = torch.frombuffer(bytearray(c.GenG1.Bytes()), dtype=torch.uint8).to(dev).repeat(n)
    pts = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(4, n, gen.data_ptr(), d_k.data_ptr(), pts.data_ptr(), m.DEVICE_PTRS | m.OUT_MONT))
    ks2 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks2[:, 0] &= 0x0F
    d_k2 = torch.from_numpy(ks2.reshape(-1)).to(dev)
    o = torch.empty(c.G1ByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_g1_msm(4, n, pts.data_ptr(), d_k2.data_ptr(), o.data_ptr(),
                                               m.DEVICE_PTRS | m.IN_MONT)), reps=2)
    out["config3_bls12_377_g1_msm_2^21_per_gpu_share"] = {"latency_ms": ms, "points_per_s": n / ms * 1e3}
    del gen, pts, d_k, d_k2

    # configs[4], one GPU's share: 12,500 BBS-style verifications = Mul2 (B = [e]G1 + [f]A) feeding Pairing2+FExp ->
    # IsUnity on the device, G2 arguments fixed (public key, generator); every second signature is valid
    import random
    c = m.Curves[6]
    n = 12500
    with open(os.path.join(ROOT, "tests", "golden", "g2_pool.json")) as f:
        pk = json.load(f)["3"][0]
    rnd = random.Random(11)
    r, bj = c.order, int(pk["b"], 16)
    a = [rnd.randrange(1, r) for _ in range(n)]
    fs = [rnd.randrange(1, r) for _ in range(n)]
    es = [(-a[i] * bj - fs[i] * a[i]) % r if i % 2 == 0 else rnd.randrange(1, r) for i in range(n)]

    def dev_bytes(b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
    gen = dev_bytes(c.GenG1.Bytes() * n)
    d_a = dev_bytes(b"".join(x.to_bytes(32, "big") for x in a))
    d_e = dev_bytes(b"".join(x.to_bytes(32, "big") for x in es))
    d_f = dev_bytes(b"".join(x.to_bytes(32, "big") for x in fs))
    d_pk = dev_bytes(bytes.fromhex(pk["g2"]) * n)
    d_g2 = dev_bytes(c.GenG2.Bytes() * n)
    d_A = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
    d_B = torch.empty(n * c.G1ByteSize, dtype=torch.uint8, device=dev)
    d_v = torch.empty(n, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(6, n, gen.data_ptr(), d_a.data_ptr(), d_A.data_ptr(), m.DEVICE_PTRS))

    def verify():
        m.check(lib.b200_g1_mul2_batch(6, n, gen.data_ptr(), d_e.data_ptr(), d_A.data_ptr(), d_f.data_ptr(),
                                       d_B.data_ptr(), m.DEVICE_PTRS))
        m.check(lib.b200_pairing2_batch(6, n, d_A.data_ptr(), d_pk.data_ptr(), d_B.data_ptr(), d_g2.data_ptr(),
                                        d_v.data_ptr(), m.DEVICE_PTRS | m.FEXP | m.OUT_UNITY_ONLY))
    ms = timed(verify)
    v = d_v.cpu().numpy()
    if not (v == np.array([1 - (i & 1) for i in range(n)], dtype=np.uint8)).all():
        raise RuntimeError("BBS-style verification verdicts are wrong")
    out["config4_bls12_381_bbs_verify_12500_per_gpu_share"] = {"ms": ms, "verifications_per_s": n / ms * 1e3,
                                                               "what": "Mul2 -> Pairing2+FExp -> IsUnity, device resident"}
    # the same verifications with the two fixed G2 arguments (public key, generator) as a resident line table (8f-1)
    import ctypes
    hl = ctypes.c_uint64()
    m.check(lib.b200_g2_lines_upload(6, 2, m.buf_ptr(bytes.fromhex(pk["g2"]) + c.GenG2.Bytes()), 0, ctypes.byref(hl)))

    def verify_fixed():
        m.check(lib.b200_g1_mul2_batch(6, n, gen.data_ptr(), d_e.data_ptr(), d_A.data_ptr(), d_f.data_ptr(),
                                       d_B.data_ptr(), m.DEVICE_PTRS))
        m.check(lib.b200_pairing2_fixed_batch(hl.value, n, d_A.data_ptr(), None, d_B.data_ptr(), None, d_v.data_ptr(),
                                              m.DEVICE_PTRS | m.FEXP | m.OUT_UNITY_ONLY))
    d_v.zero_()
    ms = timed(verify_fixed)
    if not (d_v.cpu().numpy() == v).all():
        raise RuntimeError("fixed-Q verification verdicts differ")
    out["config4_bls12_381_bbs_verify_12500_fixed_q_tables"] = {"ms": ms, "verifications_per_s": n / ms * 1e3}
    # fixed-Q Pairing2+FExp at the headline batch size (65,536 checks, same two G2 rows for every check)
    n2 = 65536
    gen2 = dev_bytes(c.GenG1.Bytes() * n2)
    kk = rng.integers(0, 256, size=(n2, 32), dtype=np.uint8)
    kk[:, 0] &= 0x3F
    d_kk = torch.from_numpy(kk.reshape(-1)).to(dev)
    d_P = torch.empty(n2 * c.G1ByteSize, dtype=torch.uint8, device=dev)
    m.check(lib.b200_g1_mul_batch(6, n2, gen2.data_ptr(), d_kk.data_ptr(), d_P.data_ptr(), m.DEVICE_PTRS))
    d_o = torch.empty(n2 * c.GtByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_pairing2_fixed_batch(hl.value, n2, d_P.data_ptr(), None, d_P.data_ptr(), None,
                                                             d_o.data_ptr(), m.DEVICE_PTRS | m.FEXP)), reps=2)
    out["next_pairing2_fexp_fixed_q_bls12_381_65536"] = {"ms": ms, "pairings_per_s": 2 * n2 / ms * 1e3}
    m.check(lib.b200_g2_lines_free(hl.value))
    del gen2, d_kk, d_P, d_o
    # SURVEY 8(f) row 3, the callers next to the hot path: Gt.Exp and G2.Mul batches on BLS12-381 (device resident)
    c = m.Curves[5]
    n = 16384
    gt = torch.frombuffer(bytearray(c.GenGt.Bytes() * n), dtype=torch.uint8).to(dev)
    ksr = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ksr[:, 0] &= 0x3F
    d_kr = torch.from_numpy(ksr.reshape(-1)).to(dev)
    o = torch.empty(n * c.GtByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_gt_exp_batch(5, n, gt.data_ptr(), d_kr.data_ptr(), o.data_ptr(), m.DEVICE_PTRS)),
               reps=2)
    out["next_gt_exp_bls12_381_16384"] = {"ms": ms, "exps_per_s": n / ms * 1e3}
    g2 = torch.frombuffer(bytearray(c.GenG2.Bytes() * n), dtype=torch.uint8).to(dev)
    o2 = torch.empty(n * c.G2ByteSize, dtype=torch.uint8, device=dev)
    ms = timed(lambda: m.check(lib.b200_g2_mul_batch(5, n, g2.data_ptr(), d_kr.data_ptr(), o2.data_ptr(), m.DEVICE_PTRS)),
               reps=2)
    out["next_g2_mul_bls12_381_16384"] = {"ms": ms, "muls_per_s": n / ms * 1e3}
    return out


if __name__ == "__main__":
    main()
