"""Hash-to-G1: the oracle against the reference's own constants and RFC 9380's known-answer vectors, and
the device code path compiled for the host (tests/hostemu) against the oracle.  GPU parity: tests/test_gpu_parity.py."""
import ctypes
import random

import pytest

from oracle import hash_to_curve as H
from oracle import codec
from oracle.params import BLS12_381

P = BLS12_381.p
RINV = pow(1 << 384, -1, P)


def from_mont_limbs(limbs):
    return sum(v << (64 * i) for i, v in enumerate(limbs)) * RINV % P


def test_swu_parameters_are_the_reference_literals():
    """reference driver/kilic/custom.go:26-42 (swuParamsForG1, Montgomery form) and :367-374 (F = 2^256 R)"""
    a = from_mont_limbs([0x2f65aa0e9af5aa51, 0x86464c2d1e8416c3, 0xb85ce591b7bd31e2, 0x27e11c91b5f24e7c, 0x28376eda6bfc1835, 0x155455c3e5071d85])
    b = from_mont_limbs([0xfb996971fe22a1e0, 0x9aa93eb35b742d6f, 0x8c476013de99c5c4, 0x873e27c3a221e571, 0xca72b5e45a52d888, 0x06824061418a386b])
    z = from_mont_limbs([0x886c00000023ffdc, 0x0f70008d3090001d, 0x77672417ed5828c3, 0x9dac23e943dc1740, 0x50553f1b9c131521, 0x078c712fbe0ab6e8])
    zinv = from_mont_limbs([0x0e8a2e8ba2e83e10, 0x5b28ba2ca4d745d1, 0x678cd5473847377a, 0x4c506dd8a8076116, 0x9bcb227d79284139, 0x0e8d3154b0ba099a])
    mba = from_mont_limbs([0x052583c93555a7fe, 0x3b40d72430f93c82, 0x1b75faa0105ec983, 0x2527e7dc63851767, 0x99fffd1f34fc181d, 0x097cab54770ca0d3])
    assert (a, b, z) == (H.ISO_A, H.ISO_B, H.SWU_Z)
    assert zinv * z % P == P - 1                   # the literal named zInv is -1/z: x1 = zInv * (-b/a) = b / (z a)
    assert (mba * a + b) % P == 0
    F = sum(v << (64 * i) for i, v in enumerate([0x75b3cd7c5ce820f, 0x3ec6ba621c3edb0b, 0x168a13d82bff6bce,
                                                 0x87663c4bf8c449d2, 0x15f34c83ddc8d830, 0xf9628b49caa2e85]))
    assert F == (1 << 256) * (1 << 384) % P        # from64Bytes: e0 * 2^256 + e1 in Montgomery form


def test_isogeny_is_a_homomorphism_onto_e():
    rnd = random.Random(3)
    pts = []
    while len(pts) < 3:
        x = rnd.randrange(P)
        rhs = (x ** 3 + H.ISO_A * x + H.ISO_B) % P
        y = pow(rhs, (P + 1) // 4, P)
        if y * y % P == rhs:
            pts.append((x, y))
    for pt in pts:
        X, Y = H.iso_map(pt)
        assert (Y * Y - X ** 3 - 4) % P == 0
    s = H._ec_add(pts[0], pts[1], H.ISO_A)
    assert H.iso_map(s) == H._ec_add(H.iso_map(pts[0]), H.iso_map(pts[1]), 0)
    xn, xd, yn, yd = H.isogeny()
    assert (len(xn), len(xd), len(yn), len(yd)) == (12, 11, 16, 16) and xd[-1] == 1 and yd[-1] == 1


# RFC 9380 appendix J.9.1, suite BLS12381G1_XMD:SHA-256_SSWU_RO_
RFC_DST = b"QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_RO_"
RFC_VECTORS = [
    (b"", 0x052926add2207b76ca4fa57a8734416c8dc95e24501772c814278700eed6d1e4e8cf62d9c09db0fac349612b759e79a1,
     0x08ba738453bfed09cb546dbb0783dbb3a5f1f566ed67bb6be0e8c67e2e81a4cc68ee29813bb7994998f3eae0c9c6a265),
    (b"abc", 0x03567bc5ef9c690c2ab2ecdf6a96ef1c139cc0b2f284dca0a9a7943388a49a3aee664ba5379a7655d3c68900be2f6903,
     0x0b9c15f3fe6e5cf4211f346271d7b01c8f3b28be689c8429c85b67af215533311f0b8dfaaa154fa6b88176c229f2885d),
]


def test_rfc9380_known_answers():
    """the standard variant is the RFC suite both libraries implement: published vectors pin hash, SWU, isogeny, cofactor"""
    u = H.hash_to_field('sha256', b"", RFC_DST)
    assert u[0] == 0x0ba14bd907ad64a016293ee7c2d276b8eae71f25a4b941eece7b0d89f17f75cb3ae5438a614fb61d6835ad59f29c564f
    for msg, x, y in RFC_VECTORS:
        assert H.hash_to_g1(msg, RFC_DST, 'standard') == (x, y)


def test_outputs_are_in_g1_and_variants_differ():
    from oracle.curve import Curve
    C = Curve(BLS12_381)
    for msg, dst in ((b"Chase!", b""), (b"CD", b"EF"), (b"Amazing Grace (how sweet the sound)", b"")):      # math_test.go:307, 904-909
        a, b = H.hash_to_g1(msg, dst, 'standard'), H.hash_to_g1(msg, dst, 'bbs')
        assert a != b
        for pt in (a, b):
            assert C.g1_on_curve(pt) and C.g1_add(C.g1_mul(pt, C.r - 1), pt) is None


MESSAGES = [b"", b"a", b"Chase!", b"x" * 55, b"y" * 56, b"z" * 63, b"w" * 64, b"v" * 65, b"u" * 119, b"t" * 127, b"s" * 128,
            b"r" * 129, bytes(range(256)) * 3]
DOMAINS = [b"", b"EF", b"powerplant", b"d" * 200, b"e" * 255]


@pytest.mark.parametrize("cid", [3, 6])
def test_hostemu_hash_to_g1_matches_oracle(hostemu, cid):
    """the per-message device function (hash_to_g1.cuh) compiled for the host: bytes equal the oracle's for every
    padding boundary of SHA-256 / BLAKE2b and for long domains"""
    var = H.variant_of(cid)
    out = ctypes.create_string_buffer(96)
    for msg in MESSAGES:
        for dst in (DOMAINS if len(msg) < 70 else DOMAINS[:2]):
            rc = hostemu.he_hash_to_g1(cid, msg, len(msg), dst, len(dst), out)
            assert rc == 0
            assert out.raw == codec.g1_to_bytes(BLS12_381, H.hash_to_g1(msg, dst, var)), (msg[:8], dst[:8])
