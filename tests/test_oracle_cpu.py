"""Pins the C++ CPU oracle (oracle/cpu) to the Python oracle's golden vectors -- every entry point, bit-exact."""
import json
import os

import pytest

from oracle import cpu_binding as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def load_vectors(cid):
    with open(os.path.join(HERE, "golden", "vectors_%d.json" % cid)) as f:
        return json.load(f)


@pytest.mark.parametrize("cid", [1, 3, 4, 5])
def test_pairing(cid):
    v = load_vectors(cid)
    cs = v["pairing"]
    g1 = b"".join(bytes.fromhex(c["g1"]) for c in cs)
    g2 = b"".join(bytes.fromhex(c["g2"]) for c in cs)
    raw = orc.pairing_batch(cid, len(cs), g1, g2)
    assert raw.hex() == "".join(c["pairing"] for c in cs)
    assert orc.fexp_batch(cid, len(cs), raw).hex() == "".join(c["fexp"] for c in cs)
    assert orc.pairing_batch(cid, len(cs), g1, g2, fexp=True).hex() == "".join(c["canonical"] for c in cs)


@pytest.mark.parametrize("cid", [1, 3, 4, 5])
def test_pairing2(cid):
    v = load_vectors(cid)
    cs = v["pairing2"]
    cols = [b"".join(bytes.fromhex(c[k]) for c in cs) for k in ("g1a", "g2a", "g1b", "g2b")]
    raw = orc.pairing_batch(cid, len(cs), cols[0], cols[1], cols[2], cols[3])
    assert raw.hex() == "".join(c["pairing2"] for c in cs)
    ver = orc.pairing_batch(cid, len(cs), cols[0], cols[1], cols[2], cols[3], fexp=True, unity=True)
    assert list(ver) == [1 if c["unity"] else 0 for c in cs]


@pytest.mark.parametrize("cid", [1, 4, 5])
def test_g1(cid):
    v = load_vectors(cid)
    cs = v["g1_mul"]
    out = orc.g1_mul_batch(cid, len(cs), b"".join(bytes.fromhex(c["p"]) for c in cs),
                           b"".join(bytes.fromhex(c["k"]) for c in cs))
    assert out.hex() == "".join(c["out"] for c in cs)
    cs = v["g1_mul2"]
    out = orc.g1_mul2_batch(cid, len(cs), b"".join(bytes.fromhex(c["p"]) for c in cs),
                            b"".join(bytes.fromhex(c["e"]) for c in cs), b"".join(bytes.fromhex(c["q"]) for c in cs),
                            b"".join(bytes.fromhex(c["f"]) for c in cs))
    assert out.hex() == "".join(c["out"] for c in cs)
    for c in v["msm"]:
        out = orc.g1_msm(cid, c["n"], bytes.fromhex(c["points"]), bytes.fromhex(c["scalars"]))
        assert out.hex() == c["out"], c["n"]
