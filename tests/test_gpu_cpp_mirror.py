"""C++ host mirror (include/b200_driver.hpp) on the GPU: same assertions as the Python mirror, through a compiled
program that only includes the two public headers."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cid", [1, 3, 5])
def test_cpp_mirror(cid):
    from oracle import cpu_binding as orc
    exe = os.path.join(ROOT, "tests", "cpp", "driver_test.bin")
    src_cpp = os.path.join(ROOT, "tests", "cpp", "driver_test.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src_cpp):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                               os.path.join(ROOT, "tests", "cpp", "driver_test.cpp"), "-L" + os.path.join(ROOT, "mathlib_b200"),
                               "-lb200math", "-Wl,-rpath," + os.path.join(ROOT, "mathlib_b200")])
    with open(os.path.join(ROOT, "tests", "golden", "vectors_%d.json" % cid)) as f:
        v = json.load(f)
    c = v["pairing2"][0]
    e, fs = v["g1_mul2"][0]["e"], v["g1_mul2"][0]["f"]
    out = subprocess.run([exe, str(cid), c["g1a"], c["g2a"], c["g1b"], c["g2b"], e, fs], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr
    res = dict(line.split(" ", 1) for line in out.stdout.strip().splitlines())
    assert res["pairing2_fexp"] == c["fexp"]
    assert res["mul2_eq_mul_add"] == "1" and res["receiver_unchanged"] == "1" and res["bad_curve_throws"] == "1"
    assert res["bilinear"] == "1" and res["inverse_ok"] == "1" and res["g2_add_is_double"] == "1" and res["compress_roundtrip"] == "1"
    g1a, g1b = bytes.fromhex(c["g1a"]), bytes.fromhex(c["g1b"])
    want = orc.g1_mul2_batch(cid, 1, g1a, bytes.fromhex(e), g1b, bytes.fromhex(fs))
    assert res["mul2"] == want.hex() and res["msm"] == want.hex()
