"""Host-side mirror of mathlib's driver interface for the hot path, on top of the C ABI.

Names, argument order and error behaviour follow the reference so that tests read like
``math_test.go``:

* ``Curve.Pairing(g2, g1)`` / ``Pairing2(p2a, p2b, p1a, p1b)`` / ``FExp(gt)`` /
  ``MultiScalarMul(points, scalars)``        -- reference driver/math.go:49-57,170
* ``G1.Mul(zr)`` / ``G1.Mul2(e, Q, f)`` / ``G1.Mul2InPlace`` / ``G1.Add`` -- reference driver/math.go:249-288
* failures raise (the Go drivers panic: reference driver/gurvy/bn254.go:249-251)

Elements hold the reference's serialized form (``Bytes()``), so ``x.Bytes()`` is directly
comparable with the reference drivers' output.  The ``*Batch`` methods are the batch entry
points ``driver/b200`` adds: contiguous slabs in, contiguous slabs out, one launch.

All arithmetic runs in the CUDA library; nothing here computes field or group operations
on the CPU (point negation below is the one exception: ``p - y`` on a Python int, host
bookkeeping exactly like the Go wrapper's ``big.Int`` handling of ``Zr``).
"""
import ctypes

from . import _lib
from ._lib import FEXP, IN_MONT, OUT_MONT, OUT_UNITY_ONLY, DEVICE_PTRS, check, load, buf_ptr

# mathlib CurveID values (reference math.go:70-103)
BN254 = 1
BLS12_381 = 3
BLS12_377_GURVY = 4
BLS12_381_GURVY = 5
BLS12_381_BBS = 6
BLS12_381_BBS_GURVY = 7

_ORDERS = {
    BN254: 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
    BLS12_381: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    BLS12_377_GURVY: 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
}
_MODULI = {
    BN254: 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
    BLS12_381: 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    BLS12_377_GURVY: 0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
}
# generators: G1 from reference math_test.go:250-259; G2 the standard ones
_G1 = {
    BN254: (1, 2),
    BLS12_381: (3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
                1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569),
    BLS12_377_GURVY: (81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
                      241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030),
}
_G2 = {
    BN254: ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
             11559732032986387107991004021392285783925812861821192530917403151452391805634),
            (8495653923123431417604973247489272438418190587263600148770280649306958101930,
             4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    BLS12_381: ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
                 0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
                (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
                 0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    BLS12_377_GURVY: ((233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
                       140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118),
                      (63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
                       149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491)),
}
_BASE = {BN254: BN254, BLS12_381: BLS12_381, BLS12_377_GURVY: BLS12_377_GURVY, BLS12_381_GURVY: BLS12_381,
         BLS12_381_BBS: BLS12_381, BLS12_381_BBS_GURVY: BLS12_381}
_NAMES = {BN254: "BN254", BLS12_381: "BLS12_381", BLS12_377_GURVY: "BLS12_377_GURVY",
          BLS12_381_GURVY: "BLS12_381_GURVY", BLS12_381_BBS: "BLS12_381_BBS",
          BLS12_381_BBS_GURVY: "BLS12_381_BBS_GURVY"}


class Zr:
    """Scalar; value kept as a Python int (may be unreduced, like common.BaseZr: reference
    driver/common/big.go:60-72).  ``Bytes()`` reduces: 32-byte big-endian (big.go:101-113)."""

    def __init__(self, curve, v):
        self.curve = curve
        self.v = int(v)

    def Bytes(self):
        return (self.v % self.curve.order).to_bytes(32, "big")

    def Plus(self, o):
        return Zr(self.curve, self.v + o.v)

    def Minus(self, o):
        return Zr(self.curve, self.v - o.v)

    def Mul(self, o):
        return Zr(self.curve, self.v * o.v % self.curve.order)

    def Equals(self, o):
        return self.v % self.curve.order == o.v % self.curve.order

    def Copy(self):
        return Zr(self.curve, self.v)


class G1:
    def __init__(self, curve, raw):
        self.curve = curve
        self.raw = bytes(raw)

    def Bytes(self):
        return self.raw

    def Copy(self):
        return G1(self.curve, self.raw)

    def Compressed(self):
        """driver.G1.Compressed (reference driver/gurvy/bn254.go:82-86)."""
        return self.curve.PointCodecBatch(0, 1, self.raw, 1)

    def Equals(self, o):
        return self.raw == o.raw

    def IsInfinity(self):
        return self.raw == self.curve._g1_inf

    def Mul(self, k):
        """driver.G1.Mul: fresh value, receiver untouched."""
        return self.curve.G1MulBatch(self.raw, k.Bytes(), 1)[0]

    def Mul2(self, e, Q, f):
        return self.curve.G1Mul2Batch(self.raw, e.Bytes(), Q.raw, f.Bytes(), 1)[0]

    def Mul2InPlace(self, e, Q, f):
        self.raw = self.Mul2(e, Q, f).raw

    def Add(self, o):
        """mutates the receiver (reference driver/math.go:256)."""
        self.raw = self.curve.G1Sum([self, o]).raw

    def Neg(self):
        if self.IsInfinity():
            return
        n = self.curve.fp_bytes
        y = int.from_bytes(self.raw[n:], "big")
        self.raw = self.raw[:n] + ((self.curve.modulus - y) % self.curve.modulus).to_bytes(n, "big")

    def Sub(self, o):
        t = o.Copy()
        t.Neg()
        self.Add(t)


class G2:
    def __init__(self, curve, raw):
        self.curve = curve
        self.raw = bytes(raw)

    def Bytes(self):
        return self.raw

    def Copy(self):
        return G2(self.curve, self.raw)

    def Compressed(self):
        """driver.G2.Compressed (reference driver/gurvy/bn254.go:167-171)."""
        return self.curve.PointCodecBatch(1, 1, self.raw, 1)

    def Equals(self, o):
        return self.raw == o.raw

    def Mul(self, k):
        """driver.G2.Mul: fresh value, receiver untouched (reference driver/math.go:307)."""
        return G2(self.curve, self.curve.G2MulBatch(self.raw, k.Bytes(), 1))

    def Add(self, o):
        """mutates the receiver (reference driver/math.go:310)."""
        self.raw = self.curve.G2Sum([self, o]).raw


class Gt:
    def __init__(self, curve, raw):
        self.curve = curve
        self.raw = bytes(raw)

    def Bytes(self):
        return self.raw

    def Equals(self, o):
        return self.raw == o.raw

    def IsUnity(self):
        return self.raw == self.curve._gt_one

    def Exp(self, k):
        """driver.Gt.Exp: fresh value (reference driver/math.go:359)."""
        return Gt(self.curve, self.curve.GtExpBatch(self.raw, k.Bytes(), 1))

    def Mul(self, o):
        """mutates the receiver (reference driver/math.go:347)."""
        self.raw = self.curve.GtMulBatch(self.raw, o.raw, 1)

    def Inverse(self):
        """mutates the receiver (reference driver/math.go:344)."""
        self.raw = self.curve.GtInvBatch(self.raw, 1)


class Curve:
    """driver.Curve for one mathlib CurveID, hot-path methods only."""

    def __init__(self, curve_id):
        if curve_id not in _BASE:
            raise ValueError("unsupported curve id %r" % (curve_id,))
        self.id = curve_id
        self.name = _NAMES[curve_id]
        base = _BASE[curve_id]
        self.order = _ORDERS[base]
        self.modulus = _MODULI[base]
        self.fp_bytes = 32 if base == BN254 else 48
        self.kilic = curve_id in (BLS12_381, BLS12_381_BBS)
        n = self.fp_bytes
        self.G1ByteSize, self.G2ByteSize, self.GtByteSize = 2 * n, 4 * n, 12 * n
        self.CoordinateByteSize, self.ScalarByteSize = n, 32
        self.CompressedG1ByteSize, self.CompressedG2ByteSize = n, 2 * n
        inf1 = bytearray(2 * n)
        if base != BN254:
            inf1[0] = 0x40
        self._g1_inf = bytes(inf1)
        self._gt_one = bytes(12 * n - 1) + b"\x01"
        gx, gy = _G1[base]
        self.GenG1 = G1(self, gx.to_bytes(n, "big") + gy.to_bytes(n, "big"))
        (x0, x1), (y0, y1) = _G2[base]
        self.GenG2 = G2(self, b"".join(v.to_bytes(n, "big") for v in (x1, x0, y1, y0)))
        self.GroupOrder = Zr(self, self.order)
        self._gen_gt = None

    # ---- constructors (reference driver/math.go:96-131) ----
    def NewZrFromInt(self, i):
        return Zr(self, i)

    def NewZrFromBytes(self, b):
        return Zr(self, int.from_bytes(b, "big"))

    def NewG1(self):
        return G1(self, self._g1_inf)

    def NewG1FromBytes(self, b):
        if len(b) != self.G1ByteSize:
            raise ValueError("failure [invalid G1 length %d]" % len(b))
        return G1(self, b)

    def NewG2FromBytes(self, b):
        if len(b) != self.G2ByteSize:
            raise ValueError("failure [invalid G2 length %d]" % len(b))
        return G2(self, b)

    def NewG1FromCompressed(self, b):
        """decompression + on-curve + subgroup checks on the device (reference driver/gurvy/bn254.go:359-367)."""
        if len(b) != self.CompressedG1ByteSize:
            raise ValueError("failure [invalid compressed G1 length %d]" % len(b))
        return G1(self, self.PointCodecBatch(0, 0, b, 1))

    def NewG2FromCompressed(self, b):
        if len(b) != self.CompressedG2ByteSize:
            raise ValueError("failure [invalid compressed G2 length %d]" % len(b))
        return G2(self, self.PointCodecBatch(1, 0, b, 1))

    def NewGtFromBytes(self, b):
        if len(b) != self.GtByteSize:
            raise ValueError("failure [invalid Gt length %d]" % len(b))
        return Gt(self, b)

    @property
    def GenGt(self):
        """FExp(Pairing(GenG2, GenG1)), computed once (reference bn254.go:298-305)."""
        if self._gen_gt is None:
            self._gen_gt = self.FExp(self.Pairing(self.GenG2, self.GenG1))
        return self._gen_gt

    # ---- single-op driver methods = the n == 1 case of the batch entry points ----
    def Pairing(self, p2, p1):
        return Gt(self, self.PairingBatch(p1.raw, p2.raw, 1))

    def Pairing2(self, p2a, p2b, p1a, p1b):
        """e(p2a,p1a)*e(p2b,p1b) -- driver argument order (reference driver/math.go:54)."""
        return Gt(self, self.Pairing2Batch(p1a.raw, p2a.raw, p1b.raw, p2b.raw, 1))

    def FExp(self, gt):
        return Gt(self, self.FExpBatch(gt.raw, 1))

    def MultiScalarMul(self, points, scalars):
        """sum [b_i]a_i.  Length mismatch yields infinity: gnark's error is discarded
        (reference bn254.go:242)."""
        if len(points) != len(scalars):
            return self.NewG1()
        pts = b"".join(p.raw for p in points)
        sc = b"".join(s.Bytes() for s in scalars)
        return G1(self, self.MsmBatch(pts, sc, len(points)))

    def G1Sum(self, points):
        lib = load()
        out = ctypes.create_string_buffer(self.G1ByteSize)
        pts = b"".join(p.raw for p in points)
        check(lib.b200_g1_sum(self.id, len(points), buf_ptr(pts), out, 0))
        return G1(self, out.raw)

    # ---- batch entry points (contiguous slabs, BYTES encoding unless flags say otherwise) ----
    def PairingBatch(self, g1, g2, n, flags=0):
        lib = load()
        osz = n if flags & OUT_UNITY_ONLY else n * self.GtByteSize
        out = ctypes.create_string_buffer(max(osz, 1))
        check(lib.b200_pairing_batch(self.id, n, buf_ptr(g1), buf_ptr(g2), out, flags))
        return out.raw[:osz]

    def Pairing2Batch(self, g1a, g2a, g1b, g2b, n, flags=0):
        lib = load()
        osz = n if flags & OUT_UNITY_ONLY else n * self.GtByteSize
        out = ctypes.create_string_buffer(max(osz, 1))
        check(lib.b200_pairing2_batch(self.id, n, buf_ptr(g1a), buf_ptr(g2a), buf_ptr(g1b), buf_ptr(g2b), out, flags))
        return out.raw[:osz]

    def FExpBatch(self, gt, n, flags=0):
        lib = load()
        osz = n if flags & OUT_UNITY_ONLY else n * self.GtByteSize
        out = ctypes.create_string_buffer(max(osz, 1))
        check(lib.b200_fexp_batch(self.id, n, buf_ptr(gt), out, flags))
        return out.raw[:osz]

    def G1MulBatch(self, pts, scalars, n, flags=0):
        lib = load()
        out = ctypes.create_string_buffer(max(n * self.G1ByteSize, 1))
        check(lib.b200_g1_mul_batch(self.id, n, buf_ptr(pts), buf_ptr(scalars), out, flags))
        sz = self.G1ByteSize
        return [G1(self, out.raw[i * sz:(i + 1) * sz]) for i in range(n)]

    def G1Mul2Batch(self, P, e, Q, f, n, flags=0):
        lib = load()
        out = ctypes.create_string_buffer(max(n * self.G1ByteSize, 1))
        check(lib.b200_g1_mul2_batch(self.id, n, buf_ptr(P), buf_ptr(e), buf_ptr(Q), buf_ptr(f), out, flags))
        sz = self.G1ByteSize
        return [G1(self, out.raw[i * sz:(i + 1) * sz]) for i in range(n)]

    # ---- callers next to the hot path ----
    def G2MulBatch(self, pts, scalars, n, flags=0):
        lib = load()
        out = ctypes.create_string_buffer(max(n * self.G2ByteSize, 1))
        check(lib.b200_g2_mul_batch(self.id, n, buf_ptr(pts), buf_ptr(scalars), out, flags))
        return out.raw[:n * self.G2ByteSize]

    def G2Sum(self, points):
        lib = load()
        out = ctypes.create_string_buffer(self.G2ByteSize)
        pts = b"".join(p.raw for p in points)
        check(lib.b200_g2_sum(self.id, len(points), buf_ptr(pts), out, 0))
        return G2(self, out.raw)

    def GtExpBatch(self, gt, scalars, n, flags=0):
        lib = load()
        out = ctypes.create_string_buffer(max(n * self.GtByteSize, 1))
        check(lib.b200_gt_exp_batch(self.id, n, buf_ptr(gt), buf_ptr(scalars), out, flags))
        return out.raw[:n * self.GtByteSize]

    def GtMulBatch(self, a, b, n, flags=0):
        lib = load()
        out = ctypes.create_string_buffer(max(n * self.GtByteSize, 1))
        check(lib.b200_gt_mul_batch(self.id, n, buf_ptr(a), buf_ptr(b), out, flags))
        return out.raw[:n * self.GtByteSize]

    def GtInvBatch(self, a, n, flags=0):
        lib = load()
        out = ctypes.create_string_buffer(max(n * self.GtByteSize, 1))
        check(lib.b200_gt_inv_batch(self.id, n, buf_ptr(a), out, flags))
        return out.raw[:n * self.GtByteSize]

    # ---- fixed-Q pairings against resident G2 line tables ----
    def G2LinesUpload(self, g2_pts, n_q, flags=0):
        lib = load()
        h = ctypes.c_uint64()
        check(lib.b200_g2_lines_upload(self.id, n_q, buf_ptr(g2_pts), flags, ctypes.byref(h)))
        return h.value

    def G2LinesFree(self, handle):
        check(load().b200_g2_lines_free(handle))

    @staticmethod
    def _idx(rows):
        if rows is None:
            return None
        import array
        return array.array("I", rows).tobytes()

    def PairingFixedBatch(self, handle, g1, rows, n, flags=0):
        lib = load()
        osz = n if flags & OUT_UNITY_ONLY else n * self.GtByteSize
        out = ctypes.create_string_buffer(max(osz, 1))
        r = self._idx(rows)
        check(lib.b200_pairing_fixed_batch(handle, n, buf_ptr(g1), buf_ptr(r) if r else None, out, flags))
        return out.raw[:osz]

    def Pairing2FixedBatch(self, handle, g1a, rows_a, g1b, rows_b, n, flags=0):
        lib = load()
        osz = n if flags & OUT_UNITY_ONLY else n * self.GtByteSize
        out = ctypes.create_string_buffer(max(osz, 1))
        ra, rb = self._idx(rows_a), self._idx(rows_b)
        check(lib.b200_pairing2_fixed_batch(handle, n, buf_ptr(g1a), buf_ptr(ra) if ra else None, buf_ptr(g1b),
                                            buf_ptr(rb) if rb else None, out, flags))
        return out.raw[:osz]

    def PointCodecBatch(self, g2, op, data, n, flags=0):
        """op 0: compressed -> Bytes(); 1: Bytes() -> compressed; 2: Bytes() -> verdict bytes."""
        lib = load()
        unc, cmp_ = (self.G2ByteSize, self.CompressedG2ByteSize) if g2 else (self.G1ByteSize, self.CompressedG1ByteSize)
        osz = n * (unc if op == 0 else cmp_ if op == 1 else 1)
        out = ctypes.create_string_buffer(max(osz, 1))
        fn = [[lib.b200_g1_decompress_batch, lib.b200_g1_compress_batch, lib.b200_g1_validate_batch],
              [lib.b200_g2_decompress_batch, lib.b200_g2_compress_batch, lib.b200_g2_validate_batch]][g2][op]
        check(fn(self.id, n, buf_ptr(data), out, flags))
        return out.raw[:osz]

    def G1NormalizeBatch(self, jacobian_mont, n, flags=0):
        """n Jacobian points (Montgomery limbs X|Y|Z) -> affine Bytes() (batch normalisation)"""
        lib = load()
        out = ctypes.create_string_buffer(max(n * self.G1ByteSize, 1))
        check(lib.b200_g1_normalize_batch(self.id, n, buf_ptr(jacobian_mont), out, flags))
        return out.raw[:n * self.G1ByteSize]

    # ---- hash-to-G1 (reference driver/math.go:120-131) ----
    def HashToG1Batch(self, messages, domain=b"", flags=0):
        """one G1 point per message (list of bytes); curve ids 3 / 5: RFC 9380 SHA-256 suite, 6 / 7: the BBS variant"""
        import array
        lib = load()
        n = len(messages)
        offs = array.array("Q", [0])
        for msg in messages:
            offs.append(offs[-1] + len(msg))
        blob = b"".join(messages)
        out = ctypes.create_string_buffer(max(n * self.G1ByteSize, 1))
        check(lib.b200_hash_to_g1_batch(self.id, n, buf_ptr(blob) if blob else None, buf_ptr(offs.tobytes()),
                                        buf_ptr(domain) if domain else None, len(domain), out, flags))
        sz = self.G1ByteSize
        return [G1(self, out.raw[i * sz:(i + 1) * sz]) for i in range(n)]

    def HashToG1(self, data):
        return self.HashToG1Batch([bytes(data)])[0]

    def HashToG1WithDomain(self, data, domain):
        return self.HashToG1Batch([bytes(data)], bytes(domain))[0]

    def MsmBatch(self, pts, scalars, n, flags=0):
        lib = load()
        out = ctypes.create_string_buffer(self.G1ByteSize)
        check(lib.b200_g1_msm(self.id, n, buf_ptr(pts) if n else None, buf_ptr(scalars) if n else None, out, flags))
        return out.raw

    def G2MsmBatch(self, pts, scalars, n, flags=0):
        """sum_i [k_i] Q_i over G2 (G2 MSM); pts = n G2.Bytes() encodings, scalars = n x 32 bytes big-endian"""
        lib = load()
        out = ctypes.create_string_buffer(self.G2ByteSize)
        check(lib.b200_g2_msm(self.id, n, buf_ptr(pts) if n else None, buf_ptr(scalars) if n else None, out, flags))
        return out.raw


# mathlib.Curves analogue: index by CurveID (reference math.go:142-255); unsupported ids are None
Curves = [None] * 8
for _cid in _BASE:
    Curves[_cid] = Curve(_cid)
