"""driver/b200/*.go cannot be compiled here (no Go toolchain); at least its constants and its cgo surface are checked:
generators / orders / moduli against the oracle, and every C function it calls against include/b200.h."""
import os
import re

from oracle.params import BN254, BLS12_381, BLS12_377

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_go_constants_match_oracle():
    s = open(os.path.join(ROOT, "driver", "b200", "params.go")).read()

    def grab(name):
        m = re.search(name + r" = mustBytes\(\d+,(.*?)\)\n", s, re.S)
        return [int(x, 16) for x in re.findall(r'"([0-9a-f]+)"', m.group(1))]

    for name, P in (("BN254", BN254), ("BLS12381", BLS12_381), ("BLS12377", BLS12_377)):
        assert tuple(grab("g1Gen" + name)) == P.g1
        (x0, x1), (y0, y1) = P.g2
        assert grab("g2Gen" + name) == [x1, x0, y1, y0]
        assert int(re.search(r"order%s\s*= mustHex\(\"([0-9a-f]+)\"\)" % name, s).group(1), 16) == P.r
        assert int(re.search(r"mod%s\s*= mustHex\(\"([0-9a-f]+)\"\)" % name, s).group(1), 16) == P.p


def test_go_calls_only_declared_c_functions():
    hdr = open(os.path.join(ROOT, "include", "b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    go = open(os.path.join(ROOT, "driver", "b200", "cabi.go")).read()
    used = set(re.findall(r"C\.(b200_[a-z0-9_]+)\(", go))
    assert used and used <= declared, used - declared
