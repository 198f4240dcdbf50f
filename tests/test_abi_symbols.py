"""CPU-side checks of the drop-in boundary: the shared object loads, exports every symbol include/b200.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_is_plain_c():
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "b200.h")])


def test_library_exports_every_declared_symbol():
    import mathlib_b200
    syms = header_symbols()
    assert len(syms) >= 17
    assert set(syms) == set(mathlib_b200.PROTOTYPES), "ctypes prototypes and include/b200.h disagree"
    lib = ctypes.CDLL(mathlib_b200.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libb200math.so does not export %s" % s


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device every compute entry point fails loudly (B200_ERR_NOGPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mathlib_b200 as m
    c = m.Curves[5]
    with pytest.raises(m.B200Error) as ei:
        c.Pairing(c.GenG2, c.GenG1)
    assert ei.value.code == m._lib.ERR_NOGPU
    with pytest.raises(m.B200Error):
        c.MultiScalarMul([c.GenG1], [c.NewZrFromInt(3)])


def test_product_does_not_import_the_oracle():
    """Nothing under mathlib_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "mathlib_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "oracle/" not in txt.replace("(does NOT import oracle/)", "") or fn == "gen_constants.py", fn
