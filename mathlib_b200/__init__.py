"""mathlib_b200 -- B200 (sm_100a) backend for IBM/mathlib's batched pairing and G1 hot path.

The product is ``libb200math.so`` (CUDA kernels + C ABI, ``include/b200.h``); this package is the
Python host-side mirror of the reference's ``driver.Curve`` interface for that path.
"""
from ._lib import (B200Error, load, check, buf_ptr, LIB_PATH, PROTOTYPES, FEXP, IN_MONT, OUT_MONT, OUT_UNITY_ONLY,
                   DEVICE_PTRS, BASES_TABLES, NO_SUBGROUP_CHECK)
from .driver import (Curve, Curves, Zr, G1, G2, Gt, BN254, BLS12_381, BLS12_377_GURVY, BLS12_381_GURVY,
                     BLS12_381_BBS, BLS12_381_BBS_GURVY)
