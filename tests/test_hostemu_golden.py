"""Kernel logic on the CPU: the device headers (fp/tower/pairing/g1/codecs) compiled for the host with the
PTX carry chains emulated in C (tests/hostemu), checked against the golden vectors.  This is how kernel
changes are validated in the GPU-less build container before spending gpurun time."""
import ctypes
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_CURVE = {1: 0, 3: 1, 5: 1, 4: 2}


def load_vectors(cid):
    with open(os.path.join(HERE, "golden", "vectors_%d.json" % cid)) as f:
        return json.load(f)


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_g1_ops(hostemu, cid):
    v = load_vectors(cid)
    ec = EMU_CURVE[cid]
    n = 32 if cid == 1 else 48
    out = ctypes.create_string_buffer(2 * n)
    for case in v["g1_mul"]:
        rc = hostemu.he_g1_op(ec, 0, bytes.fromhex(case["p"]), bytes.fromhex(case["k"]), None, None, out)
        assert rc == 0 and out.raw.hex() == case["out"]
    for case in v["g1_mul2"]:
        rc = hostemu.he_g1_op(ec, 1, bytes.fromhex(case["p"]), bytes.fromhex(case["e"]), bytes.fromhex(case["q"]),
                              bytes.fromhex(case["f"]), out)
        assert rc == 0 and out.raw.hex() == case["out"]


@pytest.mark.parametrize("cid", [4, 5])
def test_g1_glv_random_scalars(hostemu, cid):
    """The GLV split k = k1 + k2*lambda of g1.cuh on random and extreme 256-bit scalars (any value < 2^256 is legal at
    the ABI), against the C++ oracle's plain double-and-add."""
    import random
    from oracle import cpu_binding as orc
    v = load_vectors(cid)
    ec = EMU_CURVE[cid]
    base = bytes.fromhex(v["g1_mul"][0]["p"])
    q = bytes.fromhex(v["g1_mul2"][0]["q"])
    rnd = random.Random(1234 + cid)
    ks = [rnd.getrandbits(256) for _ in range(40)] + [rnd.getrandbits(b) for b in (1, 64, 120, 121, 127, 128, 129, 136, 200)]
    ks += [(1 << 256) - 1, (1 << 255), (1 << 128) - 1, (1 << 128), (1 << 120) - 1, 3]
    kb = b"".join(k.to_bytes(32, "big") for k in ks)
    want = orc.g1_mul_batch(cid, len(ks), base * len(ks), kb)
    out = ctypes.create_string_buffer(96)
    for i, k in enumerate(ks):
        assert hostemu.he_g1_op(ec, 0, base, k.to_bytes(32, "big"), None, None, out) == 0
        assert out.raw == want[96 * i:96 * (i + 1)], hex(k)
    fs = ks[::-1]
    want2 = orc.g1_mul2_batch(cid, len(ks), base * len(ks), kb, q * len(ks), b"".join(k.to_bytes(32, "big") for k in fs))
    for i in range(0, len(ks), 3):
        assert hostemu.he_g1_op(ec, 1, base, ks[i].to_bytes(32, "big"), q, fs[i].to_bytes(32, "big"), out) == 0
        assert out.raw == want2[96 * i:96 * (i + 1)]


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_pairing(hostemu, cid):
    v = load_vectors(cid)
    ec = EMU_CURVE[cid]
    n = 32 if cid == 1 else 48
    out = ctypes.create_string_buffer(12 * n)
    for case in v["pairing"][:3]:
        rc = hostemu.he_pairing_bytes(ec, 1, bytes.fromhex(case["g1"]), bytes.fromhex(case["g2"]), None, None, out, 0)
        assert rc == 0 and out.raw.hex() == case["pairing"]
        rc = hostemu.he_pairing_bytes(ec, 1, bytes.fromhex(case["g1"]), bytes.fromhex(case["g2"]), None, None, out, 1)
        assert rc == 0 and out.raw.hex() == case["canonical"]
    for case in v["pairing2"][:2]:
        rc = hostemu.he_pairing_bytes(ec, 2, bytes.fromhex(case["g1a"]), bytes.fromhex(case["g2a"]),
                                      bytes.fromhex(case["g1b"]), bytes.fromhex(case["g2b"]), out, 1)
        assert rc == 0 and out.raw.hex() == case["fexp"]


def test_kilic_semantics_vectors_are_post_fexp():
    """curve id 3 (kilic): golden 'pairing' already includes the final exponentiation and FExp is the identity
    (reference driver/kilic/bls12-381.go:260-281)."""
    v3, v5 = load_vectors(3), load_vectors(5)
    for c in v3["pairing"]:
        assert c["pairing"] == c["fexp"] == c["canonical"]
    for c in v5["pairing"][:2]:
        assert c["pairing"] != c["fexp"]


def test_bad_encoding_flag(hostemu):
    out = ctypes.create_string_buffer(96)
    bad = b"\x1f" + b"\xff" * 95
    assert hostemu.he_g1_op(1, 0, bad, (1).to_bytes(32, "big"), None, None, out) == 1


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_vm_pairing(hostemu, cid):
    """the warp-cooperative VM interpreter (vm.cuh + pairing_vm.cuh + microcode.h) on the host: lanes of a phase run
    one after another; raw Miller value and final exponentiation against the golden vectors."""
    v = load_vectors(cid)
    ec = EMU_CURVE[cid]
    n = 32 if cid == 1 else 48
    out = ctypes.create_string_buffer(12 * n)
    for case in v["pairing"]:
        assert hostemu.he_vm_pairing(ec, 1, bytes.fromhex(case["g1"]), bytes.fromhex(case["g2"]), None, None, out, 0) == 0
        assert out.raw.hex() == case["pairing"]
        assert hostemu.he_vm_pairing(ec, 1, bytes.fromhex(case["g1"]), bytes.fromhex(case["g2"]), None, None, out, 1) == 0
        assert out.raw.hex() == case["canonical"]
    for case in v["pairing2"]:
        args = [bytes.fromhex(case[k]) for k in ("g1a", "g2a", "g1b", "g2b")]
        assert hostemu.he_vm_pairing(ec, 2, *args, out, 0) == 0 and out.raw.hex() == case["pairing2"]
        assert hostemu.he_vm_pairing(ec, 2, *args, out, 1) == 0 and out.raw.hex() == case["fexp"]
    # split mode (three lanes per role: the small-batch kernel) gives the same bytes
    for case in v["pairing"][:3]:
        assert hostemu.he_vm_pairing_split(ec, 1, bytes.fromhex(case["g1"]), bytes.fromhex(case["g2"]), None, None, out, 0) == 0
        assert out.raw.hex() == case["pairing"]
    for case in v["pairing2"]:
        args = [bytes.fromhex(case[k]) for k in ("g1a", "g2a", "g1b", "g2b")]
        assert hostemu.he_vm_pairing_split(ec, 2, *args, out, 0) == 0 and out.raw.hex() == case["pairing2"]
        assert hostemu.he_vm_pairing_split(ec, 2, *args, out, 1) == 0 and out.raw.hex() == case["fexp"]


def test_fp_inverse_and_dot(hostemu):
    """binary extended-Euclid inversion == Fermat; multi-operand Montgomery product == sum of products."""
    import random
    from oracle.params import BN254, BLS12_381, BLS12_377
    rnd = random.Random(17)
    for ci, P in enumerate([BN254, BLS12_381, BLS12_377]):
        n = P.limbs32
        R = 1 << (32 * n)
        arr = lambda vs: (ctypes.c_uint32 * (len(vs) * n))(*[(v >> (32 * i)) & 0xFFFFFFFF for v in vs for i in range(n)])
        val = lambda o: sum(int(x) << (32 * i) for i, x in enumerate(o))
        for a in [1, P.p - 1, 0] + [rnd.randrange(1, P.p) for _ in range(40)]:
            o, o2 = (ctypes.c_uint32 * n)(), (ctypes.c_uint32 * n)()
            hostemu.he_fp_op(ci, 5, arr([a]), arr([0]), o)
            hostemu.he_fp_op(ci, 8, arr([a]), arr([0]), o2)
            want = 0 if a == 0 else pow(a * pow(R, -1, P.p) % P.p, -1, P.p) * R % P.p
            assert val(o) == want == val(o2)
        for _ in range(60):
            a = [rnd.randrange(P.p) for _ in range(3)]
            b = [rnd.randrange(P.p) for _ in range(3)]
            for T in (1, 2, 3):
                o = (ctypes.c_uint32 * n)()
                hostemu.he_fp_dot(ci, T, arr(a), arr(b), o)
                assert val(o) == sum(a[t] * b[t] for t in range(T)) * pow(R, -1, P.p) % P.p


def test_fp_dedicated_squaring(hostemu):
    """FpOps::sqr (upper-triangle wide square + word-sliding reduction) == a * a / R mod p, incl. 0, 1, p - 1 and values with
    all-ones limbs (every carry deposit of the wide square is exercised)."""
    import random
    from oracle.params import BN254, BLS12_381, BLS12_377
    rnd = random.Random(23)
    for ci, P in enumerate([BN254, BLS12_381, BLS12_377]):
        n = P.limbs32
        R = 1 << (32 * n)
        arr = lambda v: (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
        val = lambda o: sum(int(x) << (32 * i) for i, x in enumerate(o))
        ones = [((1 << (32 * k)) - 1) % P.p for k in range(1, n + 1)] + [(P.p - 1) ^ ((1 << (32 * k)) - 1) for k in range(1, n)]
        for a in [0, 1, 2, P.p - 1, P.p - 2, R % P.p] + [v % P.p for v in ones] + [rnd.randrange(P.p) for _ in range(200)]:
            o = (ctypes.c_uint32 * n)()
            hostemu.he_fp_op(ci, 9, arr(a), arr(0), o)
            assert val(o) == a * a * pow(R, -1, P.p) % P.p, hex(a)


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_vm_fixed_q_and_gt_exp_on_host(hostemu, cid):
    """The C++ control flow of the fixed-Q kernels (precompute_lines + miller_fixed) and of the Gt.Exp ladder, host-emulated:
    fixed-Q results equal the golden Pairing / Pairing2 bytes (raw and exponentiated, infinity cases included), Gt.Exp
    equals the golden vectors."""
    v = load_vectors(cid)
    ec = EMU_CURVE[cid]
    n = 32 if cid == 1 else 48
    out = ctypes.create_string_buffer(12 * n)
    for case in v["pairing"]:
        args = (bytes.fromhex(case["g1"]), bytes.fromhex(case["g2"]), None, None)
        assert hostemu.he_vm_pairing_fixed(ec, 1, *args, out, 0) == 0 and out.raw.hex() == case["pairing"]
    for case in v["pairing2"]:
        args = [bytes.fromhex(case[k]) for k in ("g1a", "g2a", "g1b", "g2b")]
        assert hostemu.he_vm_pairing_fixed(ec, 2, *args, out, 0) == 0 and out.raw.hex() == case["pairing2"]
        assert hostemu.he_vm_pairing_fixed(ec, 2, *args, out, 1) == 0 and out.raw.hex() == case["fexp"]
    for case in v["gt_exp"][::3]:
        assert hostemu.he_vm_gt_exp(ec, bytes.fromhex(case["a"]), bytes.fromhex(case["k"]), out) == 0
        assert out.raw.hex() == case["out"]
