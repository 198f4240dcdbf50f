"""Point (de)compression and validation (mathlib_b200/csrc/points.cuh) through the host emulation, against
oracle/codec.py -- the same per-item function the CUDA kernel runs per thread."""
import ctypes
import random

import pytest

from oracle import codec
from oracle.pairing import Pairing
from oracle.params import CURVE_IDS

EMU_CURVE = {1: 0, 5: 1, 4: 2}
NO_SUBGROUP = 0x40


def _ctx(cid):
    P, _ = CURVE_IDS[cid]
    return P, Pairing(P).C


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_g1_roundtrip_and_rejects(hostemu, cid):
    P, C = _ctx(cid)
    ec, n = EMU_CURVE[cid], P.fp_bytes
    rnd = random.Random(cid)
    pts = [None, C.g1, C.g1_neg(C.g1)] + [C.g1_mul(C.g1, rnd.randrange(1, P.r)) for _ in range(6)]
    unc, cmp_ = ctypes.create_string_buffer(2 * n), ctypes.create_string_buffer(n)
    ok = ctypes.create_string_buffer(1)
    for pt in pts:
        ub, cb = codec.g1_to_bytes(P, pt), codec.g1_to_compressed(P, pt)
        assert hostemu.he_point_codec(ec, 0, 1, ub, cmp_, 0) == 0 and cmp_.raw == cb          # G1.Compressed()
        assert hostemu.he_point_codec(ec, 0, 0, cb, unc, 0) == 0 and unc.raw == ub            # NewG1FromCompressed
        assert hostemu.he_point_codec(ec, 0, 2, ub, ok, 0) == 0 and ok.raw == b"\x01"         # NewG1FromBytes checks
    # x with no point on the curve, x >= p, an uncompressed flag, a dirty infinity
    x = 1
    while pow((x ** 3 + P.b) % P.p, (P.p - 1) // 2, P.p) == 1:
        x += 1
    bad = bytearray(x.to_bytes(n, "big"))
    bad[0] |= 0x80
    assert hostemu.he_point_codec(ec, 0, 0, bytes(bad), unc, 0) == 1
    big = bytearray(P.p.to_bytes(n, "big"))
    big[0] |= 0x80
    assert hostemu.he_point_codec(ec, 0, 0, bytes(big), unc, 0) == 1
    assert hostemu.he_point_codec(ec, 0, 0, codec.g1_to_bytes(P, C.g1)[:n], unc, 0) == 1
    inf = bytearray(codec.g1_to_compressed(P, None))
    inf[-1] = 1
    assert hostemu.he_point_codec(ec, 0, 0, bytes(inf), unc, 0) == 1
    # a point of the curve outside the order-r subgroup (BLS12 only: BN254 G1 has cofactor 1)
    x = 1
    while True:
        rhs = (x ** 3 + P.b) % P.p
        if pow(rhs, (P.p - 1) // 2, P.p) == 1:
            y = _sqrt(rhs, P.p)
            # [r]P computed as [r-1]P + P (the oracle's scalar multiplication reduces its scalar mod r)
            if P.family == "bn" or C.g1_add(C.g1_mul((x, y), P.r - 1), (x, y)) is not None:
                break
        x += 1
    pt = (x, y)
    ub = codec.g1_to_bytes(P, pt)
    hostemu.he_point_codec(ec, 0, 2, ub, ok, 0)
    assert ok.raw == (b"\x01" if P.family == "bn" else b"\x00")
    hostemu.he_point_codec(ec, 0, 2, ub, ok, NO_SUBGROUP)
    assert ok.raw == b"\x01"
    assert hostemu.he_point_codec(ec, 0, 0, codec.g1_to_compressed(P, pt), unc, 0) == (0 if P.family == "bn" else 1)
    assert hostemu.he_point_codec(ec, 0, 0, codec.g1_to_compressed(P, pt), unc, NO_SUBGROUP) == 0 and unc.raw == ub
    off = bytearray(ub)
    off[-1] ^= 1                                                                               # not on the curve
    hostemu.he_point_codec(ec, 0, 2, bytes(off), ok, NO_SUBGROUP)
    assert ok.raw == b"\x00"


def _sqrt(a, p):
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return r


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_g2_roundtrip_and_rejects(hostemu, cid):
    P, C = _ctx(cid)
    ec, n = EMU_CURVE[cid], P.fp_bytes
    rnd = random.Random(10 + cid)
    pts = [None, C.g2, C.g2_neg(C.g2)] + [C.g2_mul(C.g2, rnd.randrange(1, P.r)) for _ in range(3)]
    unc, cmp_ = ctypes.create_string_buffer(4 * n), ctypes.create_string_buffer(2 * n)
    ok = ctypes.create_string_buffer(1)
    for pt in pts:
        ub, cb = codec.g2_to_bytes(P, pt), codec.g2_to_compressed(P, pt)
        assert hostemu.he_point_codec(ec, 1, 1, ub, cmp_, 0) == 0 and cmp_.raw == cb
        assert hostemu.he_point_codec(ec, 1, 0, cb, unc, 0) == 0 and unc.raw == ub
        assert hostemu.he_point_codec(ec, 1, 2, ub, ok, 0) == 0 and ok.raw == b"\x01"
    # a twist point outside the subgroup: decompression without the subgroup check finds *a* point for a random x
    found = 0
    x0 = 5
    while not found:
        cb = bytearray((0).to_bytes(n, "big") + x0.to_bytes(n, "big"))        # X = x0 + 0*u
        cb[0] |= 0x80
        if hostemu.he_point_codec(ec, 1, 0, bytes(cb), unc, NO_SUBGROUP) == 0:
            found = 1
        else:
            x0 += 1
    pt = codec.g2_from_bytes(P, unc.raw)
    assert C.g2_on_curve(pt)
    in_sub = C.g2_add(C.g2_mul(pt, P.r - 1), pt) is None
    assert not in_sub                                     # the twist's cofactor is huge: a random point is outside G2
    hostemu.he_point_codec(ec, 1, 2, unc.raw, ok, NO_SUBGROUP)
    assert ok.raw == b"\x01"
    hostemu.he_point_codec(ec, 1, 2, unc.raw, ok, 0)
    assert ok.raw == b"\x00"
    assert hostemu.he_point_codec(ec, 1, 0, bytes(cb), unc, 0) == 1


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_batch_normalisation_on_host(hostemu, cid):
    """Montgomery-trick normalisation of Jacobian points: random Z, an infinity in the middle, batch sizes
    1..8 -- bytes equal the oracle's affine encoding."""
    import random
    import ctypes
    from oracle import codec
    from oracle.pairing import Pairing
    from oracle.params import CURVE_IDS
    P, _ = CURVE_IDS[cid]
    C = Pairing(P).C
    ec = {1: 0, 5: 1, 4: 2}[cid]
    n32 = P.limbs32
    R = 1 << (32 * n32)
    rnd = random.Random(60 + cid)

    def mont_words(v):
        v = v * R % P.p
        return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n32)]
    for n in (1, 2, 5, 8):
        pts = [C.g1_mul(C.g1, rnd.randrange(1, P.r)) for _ in range(n)]
        if n >= 5:
            pts[2] = None
        words, want = [], b""
        for pt in pts:
            z = rnd.randrange(1, P.p)
            if pt is None:
                words += mont_words(rnd.randrange(P.p)) + mont_words(rnd.randrange(P.p)) + [0] * n32
            else:
                words += mont_words(pt[0] * z * z % P.p) + mont_words(pt[1] * z * z * z % P.p) + mont_words(z)
            want += codec.g1_to_bytes(P, pt)
        arr = (ctypes.c_uint32 * len(words))(*words)
        out = ctypes.create_string_buffer(n * 2 * P.fp_bytes)
        hostemu.he_g1_normalize(ec, ctypes.c_size_t(n), arr, out)
        assert out.raw == want
