"""CPU oracle for the mathlib pairing / G1 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (``mathlib_b200``,
``libb200math.so``) imports, links or executes this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may use it, and only as the checker or as the timed CPU arm.

What it restates
----------------
IBM/mathlib's hot path (``driver.Curve.Pairing/Pairing2/FExp/MultiScalarMul`` and
``driver.G1.Mul/Mul2``, reference ``driver/math.go:49-180,249-288``) forwards to
three un-vendored Go modules (``go.mod:6,7,14``):

* github.com/consensys/gnark-crypto v0.20.1   (BN254, BLS12-377, BLS12-381 "gurvy")
* github.com/kilic/bls12-381        v0.1.0    (BLS12-381 "kilic")

Neither module, nor a Go toolchain, is present in the build container, so the
reference implementation itself cannot be executed.  This oracle restates the
*published mathematics* those modules implement (optimal-ate pairing with the
standard tower, final exponent s*(p^12-1)/r, short-Weierstrass group law,
ZCash/gnark byte formats) in Python big integers (``oracle/*.py``) and in C++
64-bit limbs (``oracle/cpu/``).

Pinning status
--------------
* Pinned by the reference's own constants / tests (checked in
  ``tests/test_oracle_pins.py``): G1 generators (``math_test.go:250-259``), group
  orders (``math_test.go:261-270``), BLS12-381 Fp modulus / -p^-1 / R / R^2 limbs
  (``driver/kilic/custom.go:26-29``, ``custom_generic.go:57-175`` CIOS routine),
  byte sizes (``bn254.go:307-329``, ``kilic/bls12-381.go:312-334``), and the algebraic
  properties ``math_test.go`` asserts (bilinearity, MSM == naive sum, Mul2 ==
  Mul+Add, e(G2,G1)^r == 1).
* Tier A (canonical group elements; any correct algorithm is byte-exact): G1/G2
  bytes, Mul/Mul2/MSM results, Gt *after* FExp.  The textbook affine Miller loop
  and a plain ``pow`` by s*(p^12-1)/r are the authority here.
* Tier B -- **parity unpinned**: the raw (pre-FExp) Miller value returned by the
  Gurvy ``Pairing``/``Pairing2``.  It is formula dependent; the formulas follow the
  gnark projective doubling/addition steps as recalled and could
  not be compared with gnark v0.20.1 output.  Only ``FExp(raw)`` is verified.
"""
