"""GPU parity at sizes the C++ CPU oracle finishes in seconds, on seeded random inputs, plus size-independent
properties at BASELINE.json's full batch sizes (bilinearity-derived unity checks, MSM linearity)."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def m():
    import mathlib_b200
    mathlib_b200.load()
    return mathlib_b200


def rand_inputs(m, cid, n, seed):
    import bench
    return bench.make_inputs(m, cid, n, seed=seed)


@pytest.mark.parametrize("cid", [1, 3, 4, 5, 6, 7])
def test_pairing2_random_vs_cpu_oracle(m, cid):
    """1,000 random Pairing2 (+FExp) per curve id, raw and exponentiated bytes, against oracle/cpu."""
    from oracle import cpu_binding as orc
    c = m.Curves[cid]
    n = 1000
    g1a, g2a, g1b, g2b, expect = rand_inputs(m, cid, n, seed=100 + cid)
    raw = c.Pairing2Batch(g1a, g2a, g1b, g2b, n)
    assert raw == orc.pairing_batch(cid, n, g1a, g2a, g1b, g2b)
    fe = c.FExpBatch(raw, n)
    want = orc.pairing_batch(cid, n, g1a, g2a, g1b, g2b, fexp=True)
    assert fe == want
    assert c.Pairing2Batch(g1a, g2a, g1b, g2b, n, m.FEXP) == want
    ver = c.Pairing2Batch(g1a, g2a, g1b, g2b, n, m.FEXP | m.OUT_UNITY_ONLY)
    assert list(ver) == list(expect)
    # single pairing path
    raw1 = c.PairingBatch(g1a, g2a, n)
    assert raw1 == orc.pairing_batch(cid, n, g1a, g2a)


@pytest.mark.parametrize("cid", [1, 4, 5, 6])
def test_g1_mul_random_vs_cpu_oracle(m, cid):
    from oracle import cpu_binding as orc
    c = m.Curves[cid]
    n = 4096
    rnd = random.Random(7 + cid)
    g1a, _, g1b, _, _ = rand_inputs(m, cid, n, seed=200 + cid)
    ks = b"".join(rnd.randrange(1 << 256).to_bytes(32, "big") for _ in range(n))
    ks2 = b"".join(rnd.randrange(c.order).to_bytes(32, "big") for _ in range(n))
    out = b"".join(p.Bytes() for p in c.G1MulBatch(g1a, ks, n))
    assert out == orc.g1_mul_batch(cid, n, g1a, ks)
    out2 = b"".join(p.Bytes() for p in c.G1Mul2Batch(g1a, ks, g1b, ks2, n))
    assert out2 == orc.g1_mul2_batch(cid, n, g1a, ks, g1b, ks2)


@pytest.mark.parametrize("cid,n", [(5, 1), (5, 7), (5, 1000), (5, 1 << 14), (1, 1 << 14), (4, 1 << 13), (3, 300),
                                   # every window size of msm_plan (c = 7..16, mixed c / c-1 bit windows) on all three orders
                                   (4, 50), (1, 3000), (5, (1 << 13) + 1), (5, 1 << 15), (1, 1 << 16), (4, 1 << 17)])
def test_msm_random_vs_cpu_oracle(m, cid, n):
    from oracle import cpu_binding as orc
    c = m.Curves[cid]
    rnd = random.Random(n + cid)
    g1a, _, _, _, _ = rand_inputs(m, cid, n, seed=300 + cid)
    ks = b"".join(rnd.randrange(c.order).to_bytes(32, "big") for _ in range(n))
    assert c.MsmBatch(g1a, ks, n) == orc.g1_msm(cid, n, g1a, ks)


@pytest.mark.parametrize("cid", [1, 3, 4])
def test_small_batch_split_kernel_is_bit_identical(m, cid):
    """BASELINE configs[0] shape: batches of at most one check per resident warp (<= 1,184) run the three-lanes-per-role
    kernel (vm_pairing_split_kernel), larger ones the five-checks-per-warp kernel.  Same inputs, same bytes -- raw Miller
    value, exponentiated value and unity verdicts, with infinity arguments in both slots."""
    c = m.Curves[cid]
    n = 1024
    g1a, g2a, g1b, g2b, expect = rand_inputs(m, cid, 2 * n, seed=700 + cid)
    gs, qs, ts = c.G1ByteSize, c.G2ByteSize, c.GtByteSize
    inf2 = bytearray(qs)
    if c.fp_bytes == 48:
        inf2[0] = 0x40
    g1a = c._g1_inf + g1a[gs:]                              # dead pair a in check 0
    g2b = g2b[:qs] + bytes(inf2) + g2b[2 * qs:]             # dead pair b in check 1
    for flags in (0, m.FEXP):
        big = c.Pairing2Batch(g1a, g2a, g1b, g2b, 2 * n, flags)                                  # 2,048 checks: 6-lane kernel
        small = c.Pairing2Batch(g1a[:n * gs], g2a[:n * qs], g1b[:n * gs], g2b[:n * qs], n, flags)     # 1,024: split kernel
        assert small == big[:n * ts]
        one = c.Pairing2Batch(g1a[:gs], g2a[:qs], g1b[:gs], g2b[:qs], 1, flags)
        assert one == big[:ts]
        big1 = c.PairingBatch(g1b, g2a, 2 * n, flags)
        assert c.PairingBatch(g1b[:n * gs], g2a[:n * qs], n, flags) == big1[:n * ts]
    ver = c.Pairing2Batch(g1a[:n * gs], g2a[:n * qs], g1b[:n * gs], g2b[:n * qs], n, m.FEXP | m.OUT_UNITY_ONLY)
    want = list(expect[:n])
    want[0] = 0 if cid else 0
    got = list(ver)
    assert got[2:] == want[2:]
    assert ver == c.Pairing2Batch(g1a, g2a, g1b, g2b, 2 * n, m.FEXP | m.OUT_UNITY_ONLY)[:n]


def test_msm_skewed_scalars(m):
    """small / repeated scalars put most points into a few buckets (the reference tests use MaxInt64 scalars,
    math_test.go:749-771): results must still be exact."""
    from oracle import cpu_binding as orc
    c = m.Curves[5]
    n = 2048
    g1a, _, _, _, _ = rand_inputs(m, 5, n, seed=401)
    ks = b"".join((((1 << 63) - 1) if i % 3 else 5).to_bytes(32, "big") for i in range(n))
    assert c.MsmBatch(g1a, ks, n) == orc.g1_msm(5, n, g1a, ks)
    # long runs: 2^16 points, all scalars equal (every point of a window lands in ONE bucket: 128 segments of 512), and a
    # mix where a third of the scalars are equal; runs longer than B200_MSM_SEG are split over extra threads (msm.cuh)
    import time
    n = 1 << 16
    g1a, _, _, _, _ = rand_inputs(m, 5, n, seed=402)
    rnd = random.Random(12)
    k0 = rnd.randrange(c.order)
    for ks in (k0.to_bytes(32, "big") * n,
               b"".join((k0 if i % 3 == 0 else rnd.randrange(c.order)).to_bytes(32, "big") for i in range(n))):
        t0 = time.perf_counter()
        got = c.MsmBatch(g1a, ks, n)
        assert time.perf_counter() - t0 < 2.0          # a single thread walking 65,536 points would take ~1 s per window
        assert got == orc.g1_msm(5, n, g1a, ks)


def test_full_size_properties(m):
    """BASELINE sizes: 65,536 BLS12-381 checks -- every even check is a valid BLS-style equation and must be 1, every odd
    one must not; MSM 2^18 linearity: msm(P, k) + msm(P, k') == msm(P, k + k')."""
    c = m.Curves[3]
    n = 65536
    g1a, g2a, g1b, g2b, expect = rand_inputs(m, 3, n, seed=55)
    ver = c.Pairing2Batch(g1a, g2a, g1b, g2b, n, m.FEXP | m.OUT_UNITY_ONLY)
    assert (np.frombuffer(ver, dtype=np.uint8) == expect).all()
    # the same 65,536 checks against a resident line table of the distinct G2 arguments: same verdicts
    qsz = c.G2ByteSize
    rows, ra = {}, []
    for i in range(n):
        ra.append(rows.setdefault(g2a[i * qsz:(i + 1) * qsz], len(rows)))
    table = list(rows) + [c.GenG2.Bytes()]
    h = c.G2LinesUpload(b"".join(table), len(table))
    assert c.Pairing2FixedBatch(h, g1a, ra, g1b, [len(table) - 1] * n, n, m.FEXP | m.OUT_UNITY_ONLY) == ver
    c.G2LinesFree(h)
    c5 = m.Curves[5]
    nm = 1 << 18
    rnd = np.random.default_rng(9)
    pts = g1a * (nm // n)
    k1 = rnd.integers(0, 256, size=(nm, 32), dtype=np.uint8)
    k2 = rnd.integers(0, 256, size=(nm, 32), dtype=np.uint8)
    k1[:, 0] &= 0x1F
    k2[:, 0] &= 0x1F
    s = (k1.astype(np.uint16)[:, ::-1] + k2.astype(np.uint16)[:, ::-1])
    carry = np.zeros(nm, dtype=np.uint16)
    ksum = np.zeros((nm, 32), dtype=np.uint8)
    for j in range(32):
        t = s[:, j] + carry
        ksum[:, 31 - j] = (t & 0xFF).astype(np.uint8)
        carry = t >> 8
    a = c5.NewG1FromBytes(c5.MsmBatch(pts, k1.tobytes(), nm))
    b = c5.NewG1FromBytes(c5.MsmBatch(pts, k2.tobytes(), nm))
    a.Add(b)
    assert a.Bytes() == c5.MsmBatch(pts, ksum.tobytes(), nm)


def _points_on_gpu(m, cid, n, seed):
    """n distinct G1 points [k_i]G in Bytes() form, made by the GPU (b200_g1_mul_batch is itself oracle-checked above)."""
    import bench
    c = m.Curves[cid]
    ks = bench.scalars_mod_r(np.random.default_rng(seed), n, cid)
    return b"".join(p.Bytes() for p in c.G1MulBatch(c.GenG1.Bytes() * n, ks.tobytes(), n))


@pytest.mark.parametrize("cid,lg", [(5, 20), (4, 21)])
def test_msm_baseline_sizes_vs_cpu_oracle(m, cid, lg):
    """BASELINE configs[2] / one GPU's share of configs[3]: 2^20 BLS12-381 and 2^21 BLS12-377 points, scalars uniform in
    [0, r), against the CPU oracle's Pippenger (reference call site bls12381/bls12-381.go:766-783)."""
    import bench
    from oracle import cpu_binding as orc
    c = m.Curves[cid]
    n = 1 << lg
    nd = 1 << 16                                   # distinct points, tiled (the bucket sums see every point 2^(lg-16) times)
    pts = _points_on_gpu(m, cid, nd, 40 + cid) * (n // nd)
    ks = bench.scalars_mod_r(np.random.default_rng(50 + cid), n, cid).tobytes()
    assert c.MsmBatch(pts, ks, n) == orc.g1_msm(cid, n, pts, ks)


def test_bn254_65536_pairing_fexp_sampled_vs_cpu_oracle(m):
    """BASELINE configs[1]: 65,536 BN254 Pairing+FExp on the GPU, every 32nd result (2,048 samples) compared byte for byte
    with the CPU oracle; the raw Miller bytes of the same samples too."""
    from oracle import cpu_binding as orc
    c = m.Curves[1]
    n, stride = 65536, 32
    g1a, g2a, _, _, _ = rand_inputs(m, 1, n, seed=31)
    out = c.PairingBatch(g1a, g2a, n, m.FEXP)
    gs, qs, ts = c.G1ByteSize, c.G2ByteSize, c.GtByteSize
    idx = range(0, n, stride)
    s1 = b"".join(g1a[i * gs:(i + 1) * gs] for i in idx)
    s2 = b"".join(g2a[i * qs:(i + 1) * qs] for i in idx)
    assert b"".join(out[i * ts:(i + 1) * ts] for i in idx) == orc.pairing_batch(1, len(idx), s1, s2, fexp=True)


def test_multi_gpu_split_matches_single(m):
    """host-buffer batch calls split by index over every visible GPU; results must not depend on the split."""
    lib = m.load()
    if lib.b200_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    c = m.Curves[5]
    n = 4096
    g1a, g2a, g1b, g2b, _ = rand_inputs(m, 5, n, seed=77)
    multi = c.Pairing2Batch(g1a, g2a, g1b, g2b, n, m.FEXP)
    m.check(lib.b200_set_device(0))
    single = c.Pairing2Batch(g1a, g2a, g1b, g2b, n, m.FEXP)
    assert multi == single


def test_resident_bases_and_threads(m):
    """b200_bases_upload / b200_g1_msm_resident against the one-shot MSM, and concurrent calls from several host
    threads sharing read-only inputs (the reference benchmarks call one curve from many goroutines, perf_test.go:392-405)."""
    import ctypes
    import threading
    lib = m.load()
    c = m.Curves[5]
    n = 5000
    g1a, g2a, g1b, g2b, expect = rand_inputs(m, 5, n, seed=91)
    rnd = random.Random(4)
    ks = b"".join(rnd.randrange(c.order).to_bytes(32, "big") for _ in range(n))
    want = c.MsmBatch(g1a, ks, n)
    h = ctypes.c_uint64()
    m.check(lib.b200_bases_upload(5, n, m.buf_ptr(g1a), 0, ctypes.byref(h)))
    out = ctypes.create_string_buffer(c.G1ByteSize)
    m.check(lib.b200_g1_msm_resident(h.value, n, m.buf_ptr(ks), out, 0))
    assert out.raw == want
    m.check(lib.b200_g1_msm_resident(h.value, 100, m.buf_ptr(ks), out, 0))          # prefix of the bases
    assert out.raw == c.MsmBatch(g1a[:100 * c.G1ByteSize], ks[:100 * 32], 100)
    m.check(lib.b200_bases_free(h.value))
    with pytest.raises(m.B200Error):
        m.check(lib.b200_g1_msm_resident(h.value, n, m.buf_ptr(ks), out, 0))
    # concurrency
    ref = c.Pairing2Batch(g1a, g2a, g1b, g2b, 512, m.FEXP | m.OUT_UNITY_ONLY)
    results, errors = [None] * 6, []

    def worker(i):
        try:
            if i % 2:
                results[i] = c.Pairing2Batch(g1a, g2a, g1b, g2b, 512, m.FEXP | m.OUT_UNITY_ONLY)
            else:
                results[i] = c.MsmBatch(g1a, ks, n)
        except Exception as ex:       # noqa: BLE001
            errors.append(ex)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors
    for i in range(6):
        assert results[i] == (ref if i % 2 else want)


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_resident_window_tables(m, cid):
    """B200_BASES_TABLES: resident bases carrying 2^(c*w)*P_i per window give the same MSM bytes as the one-shot path,
    for the table's own plan (n == bases) and for a much smaller prefix (which falls back to the plain points)."""
    import ctypes
    lib = m.load()
    c = m.Curves[cid]
    n = 3000
    g1a = rand_inputs(m, cid, n, seed=17)[0]
    # an infinity base and a repeated base in the table
    sz = c.G1ByteSize
    inf = c.NewG1().Bytes()
    g1a = inf + g1a[sz:2 * sz] + g1a[sz:2 * sz] + g1a[3 * sz:]
    rnd = random.Random(8)
    ks = b"".join(rnd.randrange(c.order).to_bytes(32, "big") for _ in range(n))
    h = ctypes.c_uint64()
    m.check(lib.b200_bases_upload(cid, n, m.buf_ptr(g1a), m.BASES_TABLES, ctypes.byref(h)))
    out = ctypes.create_string_buffer(sz)
    m.check(lib.b200_g1_msm_resident(h.value, n, m.buf_ptr(ks), out, 0))
    assert out.raw == c.MsmBatch(g1a, ks, n)
    m.check(lib.b200_g1_msm_resident(h.value, n - 1, m.buf_ptr(ks), out, 0))       # same plan, shorter run
    assert out.raw == c.MsmBatch(g1a[:(n - 1) * sz], ks[:(n - 1) * 32], n - 1)
    m.check(lib.b200_g1_msm_resident(h.value, 50, m.buf_ptr(ks), out, 0))          # different plan: plain points
    assert out.raw == c.MsmBatch(g1a[:50 * sz], ks[:50 * 32], 50)
    m.check(lib.b200_bases_free(h.value))


def test_mont_slabs_for_pairing(m):
    """IN_MONT / OUT_MONT: the zero-conversion path for gnark-layout slabs gives the same values as BYTES."""
    import ctypes
    lib = m.load()
    c = m.Curves[5]
    n = 64
    g1a, g2a, g1b, g2b, _ = rand_inputs(m, 5, n, seed=12)
    raw_b = c.Pairing2Batch(g1a, g2a, g1b, g2b, n)
    raw_m = c.Pairing2Batch(g1a, g2a, g1b, g2b, n, m.OUT_MONT)
    assert c.FExpBatch(raw_m, n, m.IN_MONT) == c.FExpBatch(raw_b, n)
    # G1 to Montgomery slabs via [1]P, then pairing with IN_MONT for G1 is not mixed-format capable: check G1 Mul only
    one = (1).to_bytes(32, "big") * n
    pm = ctypes.create_string_buffer(n * c.G1ByteSize)
    m.check(lib.b200_g1_mul_batch(5, n, m.buf_ptr(g1a), m.buf_ptr(one), pm, m.OUT_MONT))
    back = ctypes.create_string_buffer(n * c.G1ByteSize)
    m.check(lib.b200_g1_mul_batch(5, n, pm, m.buf_ptr(one), back, m.IN_MONT))
    assert back.raw == g1a


def test_multi_gpu_msm_matches_single(m):
    lib = m.load()
    if lib.b200_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    c = m.Curves[5]
    n = 1 << 17
    g1a, _, _, _, _ = rand_inputs(m, 5, 4096, seed=5)
    pts = g1a * (n // 4096)
    rnd = np.random.default_rng(3)
    ks = rnd.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ks[:, 0] &= 0x3F
    lib2 = m.load()
    import mathlib_b200._lib as L
    # fresh thread: device not pinned -> range split over all GPUs + partial-sum combine
    import threading
    res = {}
    t = threading.Thread(target=lambda: res.setdefault("multi", c.MsmBatch(pts, ks.tobytes(), n)))
    t.start()
    t.join()
    m.check(lib2.b200_set_device(0))
    assert res["multi"] == c.MsmBatch(pts, ks.tobytes(), n)


@pytest.mark.parametrize("cid", [1, 3, 4, 5])
def test_fixed_q_pairings_match_general_path(m, cid):
    """pairings against resident G2 line tables give the same bytes as Pairing / Pairing2 on the same points
    (raw Miller value, exponentiated value and unity verdict), with per-check row indices, the default rows, an infinity
    row and infinity G1 arguments."""
    import json
    import os
    c = m.Curves[cid]
    n = 300
    g1a, g2a, g1b, g2b, _ = rand_inputs(m, cid, n, seed=300 + cid)
    gsz, qsz = c.G1ByteSize, c.G2ByteSize
    # table: 6 distinct public keys out of the batch, the generator, and the point at infinity
    inf2 = bytearray(qsz)
    if c.fp_bytes == 48:
        inf2[0] = 0x40
    rows = [g2a[i * qsz:(i + 1) * qsz] for i in range(6)] + [c.GenG2.Bytes(), bytes(inf2)]
    h = c.G2LinesUpload(b"".join(rows), len(rows))
    rnd = random.Random(5 + cid)
    ra = [rnd.randrange(len(rows)) for _ in range(n)]
    rb = [rnd.randrange(len(rows)) for _ in range(n)]
    g1a = c._g1_inf + g1a[gsz:]                                   # an infinity G1 in slot a of check 0
    qa = b"".join(rows[r] for r in ra)
    qb = b"".join(rows[r] for r in rb)
    for flags in (0, m.FEXP, m.FEXP | m.OUT_UNITY_ONLY):
        assert c.Pairing2FixedBatch(h, g1a, ra, g1b, rb, n, flags) == c.Pairing2Batch(g1a, qa, g1b, qb, n, flags)
        assert c.PairingFixedBatch(h, g1b, rb, n, flags) == c.PairingBatch(g1b, qb, n, flags)
    # default rows: (row 0, row 1) for every check -- the BLS verification shape e(A_i, pk) * e(B_i, g2)
    want = c.Pairing2Batch(g1a, rows[0] * n, g1b, rows[1] * n, n, m.FEXP)
    assert c.Pairing2FixedBatch(h, g1a, None, g1b, None, n, m.FEXP) == want
    # valid BLS-style checks come out as unity: e(aG1, bG2) * e(-ab G1, G2) = 1 with pk = bG2 in row 0, G2 in row 1
    with open(os.path.join(os.path.dirname(__file__), "golden", "g2_pool.json")) as f:
        pk = json.load(f)[str({3: 3, 5: 3, 1: 1, 4: 4}[cid])][0]
    h2 = c.G2LinesUpload(bytes.fromhex(pk["g2"]) + c.GenG2.Bytes(), 2)
    b = int(pk["b"], 16)
    a = [rnd.randrange(1, c.order) for _ in range(64)]
    A = b"".join(p.Bytes() for p in c.G1MulBatch(c.GenG1.Bytes() * 64, b"".join(x.to_bytes(32, "big") for x in a), 64))
    B = b"".join(p.Bytes() for p in c.G1MulBatch(
        c.GenG1.Bytes() * 64, b"".join(((c.order - x * b) % c.order).to_bytes(32, "big") for x in a), 64))
    assert c.Pairing2FixedBatch(h2, A, None, B, None, 64, m.FEXP | m.OUT_UNITY_ONLY) == b"\x01" * 64
    with pytest.raises(m.B200Error):
        c.Pairing2FixedBatch(h2, A, [0] * 64, B, [2] * 64, 64, m.FEXP)          # row index out of range
    c.G2LinesFree(h)
    c.G2LinesFree(h2)
    with pytest.raises(m.B200Error):
        c.Pairing2FixedBatch(h2, A, None, B, None, 64, m.FEXP)


def test_torchrun_two_ranks_sharded_msm(m):
    """SURVEY 8e line 2 on hardware: two processes, one GPU each, range-split MSM + NCCL all-gather of the partial sums
    + b200_g1_sum (mathlib_b200.shard.msm_sharded_device), checked against the CPU oracle inside the worker."""
    import subprocess
    import sys
    lib = m.load()
    if lib.b200_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(HERE)
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(HERE, "dist_msm_worker.py")], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "SHARDED_MSM_OK ranks=2" in out.stdout
