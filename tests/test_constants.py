"""The generated csrc/constants.h (product) against values derived independently by the oracle."""
import os
import re

from oracle.params import BN254, BLS12_381, BLS12_377
from oracle.tower import Tower

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse():
    out = {}
    with open(os.path.join(ROOT, "mathlib_b200", "csrc", "constants.h")) as f:
        for line in f:
            m = re.match(r"#define (\w+) \{(.*)\}", line)
            if m:
                out[m.group(1)] = [int(x.rstrip("u"), 16) for x in m.group(2).split(",")]
            m = re.match(r"#define (\w+_INV32) (0x[0-9a-f]+)u", line)
            if m:
                out[m.group(1)] = int(m.group(2), 16)
    return out


def val(ws):
    return sum(w << (32 * i) for i, w in enumerate(ws))


def test_constants_header():
    k = parse()
    for name, P in (("BN254", BN254), ("BLS381", BLS12_381), ("BLS377", BLS12_377)):
        n = P.limbs32
        R = 1 << (32 * n)
        T = Tower(P)
        assert val(k[name + "_P"]) == P.p
        assert val(k[name + "_ONE"]) == R % P.p
        assert val(k[name + "_R2"]) == R * R % P.p
        assert k[name + "_INV32"] == (-pow(P.p, -1, 1 << 32)) % (1 << 32)
        assert val(k[name + "_ORDER"]) == P.r
        assert val(k[name + "_B"]) == P.b * R % P.p
        b2 = T.f2_mul((P.b, 0), T.f2_inv(T.xi)) if P.twist == "D" else T.f2_mul((P.b, 0), T.xi)
        assert val(k[name + "_BTW"][:n]) == b2[0] * R % P.p and val(k[name + "_BTW"][n:]) == b2[1] * R % P.p
        for kk in (1, 2, 3):
            ws = k[name + "_FROB%d" % kk]
            gam = T.frob_consts(kk)
            for i in range(1, 6):
                c0 = val(ws[(i - 1) * 2 * n:(i - 1) * 2 * n + n])
                c1 = val(ws[(i - 1) * 2 * n + n:i * 2 * n])
                assert (c0, c1) == (gam[i][0] * R % P.p, gam[i][1] * R % P.p)
