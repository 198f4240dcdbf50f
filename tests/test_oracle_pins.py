"""Pins the oracle to every constant / known answer the reference itself holds for this path."""
import ctypes
import random

import pytest

from oracle.params import BN254, BLS12_381, BLS12_377, CURVES, CURVE_IDS
from oracle.pairing import Pairing
from oracle import codec


def test_group_orders_match_reference_math_test():
    # reference math_test.go:261-270 (expectedModuli)
    assert "%x" % BN254.r == "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001"
    assert "%x" % BLS12_381.r == "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001"
    assert "%x" % BLS12_377.r == "12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001"


def test_g1_generators_match_reference_math_test():
    # reference math_test.go:250-259 (expectedG1Gens); BN254 prints (1,2)
    assert BN254.g1 == (1, 2)
    assert BLS12_381.g1 == (
        3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
        1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569)
    assert BLS12_377.g1 == (
        81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
        241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030)
    for P in CURVES.values():
        pr = Pairing(P)
        assert pr.C.g1_on_curve(P.g1) and pr.C.g2_on_curve(P.g2)
        assert pr.C.g1_mul(P.g1, P.r - 1) == pr.C.g1_neg(P.g1)


def test_kilic_fp_constants():
    # reference driver/kilic/custom.go:26-29 (modulus, r1) and custom_generic.go:64 (-p^-1 mod 2^64)
    limbs = [0xb9feffffffffaaab, 0x1eabfffeb153ffff, 0x6730d2a0f6b0f624, 0x64774b84f38512bf, 0x4b1ba7b6434bacd7,
             0x1a0111ea397fe69a]
    p = sum(l << (64 * i) for i, l in enumerate(limbs))
    assert p == BLS12_381.p
    r1 = [0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745, 0x5c071a97a256ec6d,
          0x15f65ec3fa80e493]
    assert sum(l << (64 * i) for i, l in enumerate(r1)) == (1 << 384) % p
    assert (-pow(p, -1, 1 << 64)) % (1 << 64) == 9940570264628428797


def test_cpu_oracle_cios_matches_reference_routine_constants():
    """oracle/cpu derives -p^-1, R, R^2 itself; they must equal the reference's literals, and its CIOS product
    must equal a*b*R^-1 mod p (what driver/kilic/custom_generic.go:57-175 computes)."""
    from oracle import cpu_binding as orc
    lib = orc.load()
    p = (ctypes.c_uint64 * 6)()
    one = (ctypes.c_uint64 * 6)()
    r2 = (ctypes.c_uint64 * 6)()
    inv = ctypes.c_uint64()
    assert lib.orc_consts(3, p, ctypes.byref(inv), one, r2) == 0
    assert inv.value == 9940570264628428797
    assert list(p)[0] == 0xb9feffffffffaaab and list(one)[0] == 0x760900000002fffd
    P = BLS12_381
    rnd = random.Random(5)
    R = 1 << 384
    for _ in range(200):
        a, b = rnd.randrange(P.p), rnd.randrange(P.p)
        A = (ctypes.c_uint64 * 6)(*[(a >> (64 * i)) & (2**64 - 1) for i in range(6)])
        B = (ctypes.c_uint64 * 6)(*[(b >> (64 * i)) & (2**64 - 1) for i in range(6)])
        O = (ctypes.c_uint64 * 6)()
        lib.orc_fp_mul_raw(3, A, B, O)
        assert sum(int(x) << (64 * i) for i, x in enumerate(O)) == a * b * pow(R, -1, P.p) % P.p


def test_byte_sizes_match_reference_adapters():
    # reference bn254.go:307-329, kilic/bls12-381.go:312-334
    for P, g1, g2, gt in ((BN254, 64, 128, 384), (BLS12_381, 96, 192, 576), (BLS12_377, 96, 192, 576)):
        assert len(codec.g1_to_bytes(P, P.g1)) == g1 and len(codec.g1_to_compressed(P, P.g1)) == g1 // 2
        assert len(codec.g2_to_bytes(P, P.g2)) == g2 and len(codec.g2_to_compressed(P, P.g2)) == g2 // 2
        assert len(codec.zr_to_bytes(P, 5)) == 32
        pr = Pairing(P)
        assert len(codec.gt_to_bytes(P, pr.T.f12_one)) == gt
        assert codec.g1_from_bytes(P, codec.g1_to_bytes(P, P.g1)) == P.g1
        assert codec.g2_from_bytes(P, codec.g2_to_bytes(P, P.g2)) == P.g2
        assert codec.g1_from_bytes(P, codec.g1_to_bytes(P, None)) is None


@pytest.mark.parametrize("name", ["BN254", "BLS12_381", "BLS12_377"])
def test_reference_properties(name):
    """The algebraic assertions of reference math_test.go: bilinearity (423-455), e^r = 1, GenGt (457-470),
    MSM == naive sum (323-346), Mul2 == Mul + Add (290), JointScalarMultiplication restatement (bls12-381.go:869-937)."""
    P = CURVES[name]
    pr = Pairing(P)
    C, T = pr.C, pr.T
    rnd = random.Random(11)
    a, b = rnd.randrange(P.r), rnd.randrange(P.r)
    g = pr.final_exp(pr.miller_projective([(C.g1, C.g2)]))
    assert g != T.f12_one and T.f12_pow(g, P.r) == T.f12_one
    lhs = pr.final_exp(pr.miller_projective([(C.g1_mul(C.g1, a), C.g2_mul(C.g2, b))]))
    assert lhs == T.f12_pow(g, a * b % P.r)
    assert lhs == pr.final_exp(pr.miller_textbook([(C.g1_mul(C.g1, a), C.g2_mul(C.g2, b))]))
    assert lhs == pr.final_exp_plain(pr.miller_textbook([(C.g1_mul(C.g1, a), C.g2_mul(C.g2, b))]))
    pts = [C.g1_mul(C.g1, rnd.randrange(P.r)) for _ in range(10)]
    ks = [rnd.randrange(P.r) for _ in range(10)]
    acc = None
    for pt, k in zip(pts, ks):
        acc = C.g1_add(acc, C.g1_mul(pt, k))
    assert C.g1_msm(pts, ks) == acc
    e, f = (1 << 63) - 1, rnd.randrange(P.r)
    assert C.g1_mul2(pts[0], e, pts[1], f) == C.joint_scalar_mul(pts[0], pts[1], e, f)
    assert C.joint_scalar_mul(pts[0], pts[1], -5, 7) == C.g1_add(C.g1_mul(pts[0], P.r - 5), C.g1_mul(pts[1], 7))


def test_kilic_and_gurvy_semantics_agree_after_fexp():
    """reference math_test.go:879-945 (Test381Compat): both BLS12-381 drivers yield the same bytes after FExp."""
    P, _ = CURVE_IDS[3]
    pr = Pairing(P)
    C = pr.C
    Q, G = C.g2_mul(C.g2, 77), C.g1_mul(C.g1, 99)
    k = pr.fexp(pr.pairing(Q, G, "kilic"), "kilic")
    g = pr.fexp(pr.pairing(Q, G, "gurvy"), "gurvy")
    assert codec.gt_to_bytes(P, k) == codec.gt_to_bytes(P, g)
