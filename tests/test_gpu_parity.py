"""GPU parity: every C-ABI hot-path entry point against the committed golden vectors
(tests/golden/vectors_<curve id>.json, produced by the Python oracle) -- bit-exact bytes.

Mirrors the reference's own assertions where they exist: runPairingTest / runMultiScalarMul /
the Mul2 line of runG1Test (reference math_test.go:423-455, 323-346, 290).
"""
import json
import os

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CURVE_IDS = [1, 3, 4, 5]


def load_vectors(cid):
    with open(os.path.join(HERE, "golden", "vectors_%d.json" % cid)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def m():
    import mathlib_b200
    mathlib_b200.load()
    return mathlib_b200


@pytest.mark.parametrize("cid", CURVE_IDS)
def test_pairing_golden(m, cid):
    c = m.Curves[cid]
    v = load_vectors(cid)
    for case in v["pairing"]:
        p1 = c.NewG1FromBytes(bytes.fromhex(case["g1"]))
        p2 = c.NewG2FromBytes(bytes.fromhex(case["g2"]))
        raw = c.Pairing(p2, p1)
        assert raw.Bytes().hex() == case["pairing"]
        fe = c.FExp(raw)
        assert fe.Bytes().hex() == case["fexp"]
        assert fe.Bytes().hex() == case["canonical"]        # Tier A: textbook definition
    # fused Pairing+FExp in one launch gives the same bytes
    n = len(v["pairing"])
    g1 = b"".join(bytes.fromhex(x["g1"]) for x in v["pairing"])
    g2 = b"".join(bytes.fromhex(x["g2"]) for x in v["pairing"])
    out = c.PairingBatch(g1, g2, n, m.FEXP)
    assert out.hex() == "".join(x["canonical"] for x in v["pairing"])


@pytest.mark.parametrize("cid", CURVE_IDS)
def test_pairing2_golden(m, cid):
    c = m.Curves[cid]
    v = load_vectors(cid)
    for case in v["pairing2"]:
        args = [bytes.fromhex(case[k]) for k in ("g1a", "g2a", "g1b", "g2b")]
        raw = c.Pairing2(c.NewG2FromBytes(args[1]), c.NewG2FromBytes(args[3]), c.NewG1FromBytes(args[0]),
                         c.NewG1FromBytes(args[2]))
        assert raw.Bytes().hex() == case["pairing2"]
        fe = c.FExp(raw)
        assert fe.Bytes().hex() == case["fexp"]
        assert fe.IsUnity() == case["unity"]
    n = len(v["pairing2"])
    cols = [b"".join(bytes.fromhex(x[k]) for x in v["pairing2"]) for k in ("g1a", "g2a", "g1b", "g2b")]
    verdict = c.Pairing2Batch(cols[0], cols[1], cols[2], cols[3], n, m.FEXP | m.OUT_UNITY_ONLY)
    assert list(verdict) == [1 if x["unity"] else 0 for x in v["pairing2"]]


@pytest.mark.parametrize("cid", CURVE_IDS)
def test_g1_mul_golden(m, cid):
    c = m.Curves[cid]
    v = load_vectors(cid)
    for case in v["g1_mul"]:
        p = c.NewG1FromBytes(bytes.fromhex(case["p"]))
        before = p.Bytes()
        r = p.Mul(c.NewZrFromBytes(bytes.fromhex(case["k"])))
        assert r.Bytes().hex() == case["out"]
        assert p.Bytes() == before                          # receiver unchanged (math_test.go:93-95)


@pytest.mark.parametrize("cid", CURVE_IDS)
def test_g1_mul2_golden(m, cid):
    c = m.Curves[cid]
    v = load_vectors(cid)
    for case in v["g1_mul2"]:
        p = c.NewG1FromBytes(bytes.fromhex(case["p"]))
        q = c.NewG1FromBytes(bytes.fromhex(case["q"]))
        e = c.NewZrFromBytes(bytes.fromhex(case["e"]))
        f = c.NewZrFromBytes(bytes.fromhex(case["f"]))
        r = p.Mul2(e, q, f)
        assert r.Bytes().hex() == case["out"]
        # Mul2 == Mul + Add (math_test.go:290) and Mul2InPlace (math_test.go:749-771)
        a = p.Mul(e)
        a.Add(q.Mul(f))
        assert a.Equals(r)
        p2 = p.Copy()
        p2.Mul2InPlace(e, q, f)
        assert p2.Equals(r)


@pytest.mark.parametrize("cid", CURVE_IDS)
def test_msm_golden(m, cid):
    c = m.Curves[cid]
    v = load_vectors(cid)
    for case in v["msm"]:
        n = case["n"]
        out = c.MsmBatch(bytes.fromhex(case["points"]), bytes.fromhex(case["scalars"]), n)
        assert out.hex() == case["out"], "n=%d" % n


@pytest.mark.parametrize("cid", CURVE_IDS)
def test_mont_encoding_roundtrip(m, cid):
    """OUT_MONT then IN_MONT must reproduce the BYTES results (zero-conversion path for Go slabs)."""
    c = m.Curves[cid]
    v = load_vectors(cid)
    case = v["pairing"][0]
    g1, g2 = bytes.fromhex(case["g1"]), bytes.fromhex(case["g2"])
    raw_m = c.PairingBatch(g1, g2, 1, m.OUT_MONT)
    fe = c.FExpBatch(raw_m, 1, m.IN_MONT)
    assert fe.hex() == case["fexp"]
    k = bytes.fromhex(v["g1_mul"][6]["k"])
    p = bytes.fromhex(v["g1_mul"][6]["p"])
    import ctypes
    lib = m.load()
    pm = ctypes.create_string_buffer(c.G1ByteSize)
    # [1]P with OUT_MONT is the Montgomery encoding of P
    one = (1).to_bytes(32, "big")
    m.check(lib.b200_g1_mul_batch(c.id, 1, m.buf_ptr(p), m.buf_ptr(one), pm, m.OUT_MONT))
    out = ctypes.create_string_buffer(c.G1ByteSize)
    m.check(lib.b200_g1_mul_batch(c.id, 1, pm, m.buf_ptr(k), out, m.IN_MONT))
    assert out.raw.hex() == v["g1_mul"][6]["out"]


def test_bad_encoding_is_an_error(m):
    c = m.Curves[5]
    bad = b"\x1f" + b"\xff" * 95            # x >= p
    with pytest.raises(m.B200Error):
        c.G1MulBatch(bad, (1).to_bytes(32, "big"), 1)
    with pytest.raises(m.B200Error):
        m.check(m.load().b200_fp_bytes(99))
