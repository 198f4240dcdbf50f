package b200

import (
	"crypto/rand"
	"crypto/sha256"
	"io"
	"math/big"

	"github.com/IBM/mathlib/driver"
	"github.com/IBM/mathlib/driver/common"
)

// mathlib CurveID values this driver serves (reference math.go:70-103).
const (
	idBN254          = 1
	idBLS12381       = 3 // kilic semantics: Pairing includes FExp, FExp is the identity
	idBLS12377Gurvy  = 4
	idBLS12381Gurvy  = 5
	idBLS12381BBS    = 6
	idBLS12381BBSGur = 7
)

// Curve implements driver.Curve (reference driver/math.go:49-180) for one CurveID.
type Curve struct {
	common.CurveBase // ModAdd/ModSub/ModMul/... on big.Int scalars: host bookkeeping, reference driver/common/curve.go
	id        int
	fpBytes   int
	g1Gen     []byte
	g2Gen     []byte
	genGt     *Gt
}

func newCurve(id, fpBytes int, order *big.Int, g1Gen, g2Gen []byte) *Curve {
	return &Curve{CurveBase: common.CurveBase{Modulus: *order}, id: id, fpBytes: fpBytes, g1Gen: g1Gen, g2Gen: g2Gen}
}

// NewBn254, NewBls12_381 (kilic semantics), NewBls12_381Gurvy, NewBls12_377 and the two BBS flavours mirror the reference
// constructors (reference driver/gurvy/bn254.go:289, driver/kilic/bls12-381.go:294 and :298 (NewBls12_381BBS),
// driver/gurvy/bls12381/bls12-381.go:441 and :789 (NewBBSCurve)).  The BBS ids differ from 3 / 5 only in HashToG1
// (BLAKE2b-512 + big-endian sign rule, reference driver/kilic/custom.go:205-237).
func NewBn254() *Curve            { return newCurve(idBN254, 32, orderBN254, g1GenBN254, g2GenBN254) }
func NewBls12_381() *Curve        { return newCurve(idBLS12381, 48, orderBLS12381, g1GenBLS12381, g2GenBLS12381) }
func NewBls12_381Gurvy() *Curve   { return newCurve(idBLS12381Gurvy, 48, orderBLS12381, g1GenBLS12381, g2GenBLS12381) }
func NewBls12_377() *Curve        { return newCurve(idBLS12377Gurvy, 48, orderBLS12377, g1GenBLS12377, g2GenBLS12377) }
func NewBls12_381BBS() *Curve     { return newCurve(idBLS12381BBS, 48, orderBLS12381, g1GenBLS12381, g2GenBLS12381) }
func NewBls12_381BBSGurvy() *Curve { return newCurve(idBLS12381BBSGur, 48, orderBLS12381, g1GenBLS12381, g2GenBLS12381) }

func (c *Curve) g1Size() int { return 2 * c.fpBytes }
func (c *Curve) g2Size() int { return 4 * c.fpBytes }
func (c *Curve) gtSize() int { return 12 * c.fpBytes }

// ---- hot path -------------------------------------------------------------------------------------------------

// Pairing computes e(p2, p1) (raw Miller value for gurvy ids, exponentiated for kilic ids): the n == 1 case of
// PairingBatch.  reference driver/math.go:51.
func (c *Curve) Pairing(p2 driver.G2, p1 driver.G1) driver.Gt {
	return &Gt{c: c, raw: pairingBatch(c.id, 1, p1.(*G1).raw, p2.(*G2).raw, c.gtSize(), 0)}
}

// Pairing2 computes e(p2a,p1a)*e(p2b,p1b) with one shared squaring chain.  reference driver/math.go:54.
func (c *Curve) Pairing2(p2a, p2b driver.G2, p1a, p1b driver.G1) driver.Gt {
	return &Gt{c: c, raw: pairing2Batch(c.id, 1, p1a.(*G1).raw, p2a.(*G2).raw, p1b.(*G1).raw, p2b.(*G2).raw, c.gtSize(), 0)}
}

// FExp: final exponentiation (identity for kilic ids).  reference driver/math.go:57.
func (c *Curve) FExp(a driver.Gt) driver.Gt {
	return &Gt{c: c, raw: fexpBatch(c.id, 1, a.(*Gt).raw, 0)}
}

// MultiScalarMul: sum [b_i]a_i; a length mismatch yields infinity because the reference discards gnark's error
// (reference driver/gurvy/bn254.go:242).  Scalars are reduced mod r (Euclidean) before upload.
func (c *Curve) MultiScalarMul(a []driver.G1, b []driver.Zr) driver.G1 {
	if len(a) != len(b) {
		return c.NewG1()
	}
	pts := make([]byte, 0, len(a)*c.g1Size())
	ks := make([]byte, 0, len(a)*32)
	for i := range a {
		pts = append(pts, a[i].(*G1).raw...)
		ks = append(ks, b[i].Bytes()...)
	}
	return &G1{c: c, raw: g1Msm(c.id, len(a), pts, ks, c.g1Size(), 0)}
}

// ---- batch entry points added by this driver (contiguous slabs in, contiguous slabs out) ------------------------

// PairingBatch / Pairing2Batch / FExpBatch / VerifyBatch amortise one launch over n independent operations.
func (c *Curve) PairingBatch(n int, g1, g2 []byte, fexp bool) []byte {
	var f uint32
	if fexp {
		f = flagFExp
	}
	return pairingBatch(c.id, n, g1, g2, c.gtSize(), f)
}

func (c *Curve) Pairing2Batch(n int, g1a, g2a, g1b, g2b []byte, fexp bool) []byte {
	var f uint32
	if fexp {
		f = flagFExp
	}
	return pairing2Batch(c.id, n, g1a, g2a, g1b, g2b, c.gtSize(), f)
}

// VerifyBatch returns one byte per item: 1 iff FExp(Pairing2(...)).IsUnity() (the check of reference perf_test.go:254-259).
func (c *Curve) VerifyBatch(n int, g1a, g2a, g1b, g2b []byte) []byte {
	return pairing2Batch(c.id, n, g1a, g2a, g1b, g2b, 1, flagFExp|flagUnityOnly)
}

func (c *Curve) G1MulBatch(n int, pts, scalars []byte) []byte { return g1MulBatch(c.id, n, pts, scalars, 0) }

// MultiScalarMulG2: sum [b_i]a_i over G2 in one Pippenger run (SURVEY 8f-3); a length mismatch yields infinity like
// MultiScalarMul.
func (c *Curve) MultiScalarMulG2(a []driver.G2, b []driver.Zr) driver.G2 {
	if len(a) != len(b) {
		return c.NewG2()
	}
	pts := make([]byte, 0, len(a)*c.g2Size())
	ks := make([]byte, 0, len(a)*32)
	for i := range a {
		pts = append(pts, a[i].(*G2).raw...)
		ks = append(ks, b[i].Bytes()...)
	}
	return &G2{c: c, raw: g2Msm(c.id, len(a), pts, ks, c.g2Size(), 0)}
}

// G1NormalizeBatch: Jacobian Montgomery slabs (kilic PointG1 / gnark G1Jac memory) -> affine Bytes(), batched inversion.
func (c *Curve) G1NormalizeBatch(n int, jacobianMont []byte) []byte {
	return g1NormalizeBatch(c.id, n, jacobianMont, c.g1Size())
}

// ---- the remaining driver.Curve methods: host bookkeeping ---------------------------------------------------------

func (c *Curve) GenG1() driver.G1 { return &G1{c: c, raw: append([]byte(nil), c.g1Gen...)} }
func (c *Curve) GenG2() driver.G2 { return &G2{c: c, raw: append([]byte(nil), c.g2Gen...)} }
func (c *Curve) GenGt() driver.Gt {
	if c.genGt == nil { // FExp(Pairing(GenG2, GenG1)) once, reference driver/gurvy/bn254.go:298-305
		c.genGt = c.FExp(c.Pairing(c.GenG2(), c.GenG1())).(*Gt)
	}
	return c.genGt
}
func (c *Curve) CoordinateByteSize() int   { return c.fpBytes }
func (c *Curve) G1ByteSize() int           { return c.g1Size() }
func (c *Curve) CompressedG1ByteSize() int { return c.fpBytes }
func (c *Curve) G2ByteSize() int           { return c.g2Size() }
func (c *Curve) CompressedG2ByteSize() int { return 2 * c.fpBytes }
func (c *Curve) ScalarByteSize() int       { return 32 }

func (c *Curve) NewG1() driver.G1 {
	raw := make([]byte, c.g1Size())
	if c.fpBytes == 48 {
		raw[0] = 0x40 // uncompressed-infinity flag (SURVEY A.3)
	}
	return &G1{c: c, raw: raw}
}
func (c *Curve) NewG2() driver.G2 {
	raw := make([]byte, c.g2Size())
	if c.fpBytes == 48 {
		raw[0] = 0x40
	}
	return &G2{c: c, raw: raw}
}
func (c *Curve) NewG1FromBytes(b []byte) driver.G1 {
	if len(b) != c.g1Size() {
		panic("failure [invalid G1 length]")
	}
	// gnark SetBytes / kilic FromUncompressed reject non-canonical, off-curve and out-of-subgroup points
	// (reference driver/gurvy/bn254.go:339-347, kilic/bls12-381.go:344-355)
	if pointCodec(c.id, 0, 2, 1, b, 1, 0)[0] != 1 {
		panic("set bytes failed [invalid point]")
	}
	return &G1{c: c, raw: append([]byte(nil), b...)}
}
func (c *Curve) NewG2FromBytes(b []byte) driver.G2 {
	if len(b) != c.g2Size() {
		panic("failure [invalid G2 length]")
	}
	if pointCodec(c.id, 1, 2, 1, b, 1, 0)[0] != 1 {
		panic("set bytes failed [invalid point]")
	}
	return &G2{c: c, raw: append([]byte(nil), b...)}
}
func (c *Curve) NewGtFromBytes(b []byte) driver.Gt {
	if len(b) != c.gtSize() {
		panic("failure [invalid Gt length]")
	}
	return &Gt{c: c, raw: append([]byte(nil), b...)}
}

// NewG1FromCompressed / NewG2FromCompressed decompress, and check curve and subgroup membership, on the device
// (SURVEY 8f-2; reference driver/gurvy/bn254.go:359-377, kilic/bls12-381.go:370-394).  A bad encoding panics with the
// reference's message style.
func (c *Curve) NewG1FromCompressed(b []byte) driver.G1 {
	if len(b) != c.fpBytes {
		panic("failure [invalid compressed G1 length]")
	}
	return &G1{c: c, raw: pointCodec(c.id, 0, 0, 1, b, c.g1Size(), 0)}
}
func (c *Curve) NewG2FromCompressed(b []byte) driver.G2 {
	if len(b) != 2*c.fpBytes {
		panic("failure [invalid compressed G2 length]")
	}
	return &G2{c: c, raw: pointCodec(c.id, 1, 0, 1, b, c.g2Size(), 0)}
}

func (c *Curve) HashToZr(data []byte) driver.Zr {
	digest := sha256.Sum256(data) // reference driver/common/curve.go:86-92
	return c.NewZrFromBytes(digest[:])
}

// HashToG1 / HashToG1WithDomain (SURVEY 8f-4): on the BLS12-381 ids the whole pipeline -- expand_message_xmd, SWU,
// 11-isogeny, cofactor clearing -- runs on the device (b200_hash_to_g1_batch): RFC 9380's SHA-256 suite for ids 3 / 5
// (reference driver/kilic/bls12-381.go:410-447, driver/gurvy/bls12381/bls12-381.go:652-677), the BLAKE2b / big-endian-sign
// variant for the BBS ids (reference driver/kilic/bls12-381.go:462-497).  BN254 / BLS12-377 hashing is not on this path.
func (c *Curve) HashToG1(data []byte) driver.G1 { return c.HashToG1WithDomain(data, []byte{}) }
func (c *Curve) HashToG1WithDomain(data, domain []byte) driver.G1 {
	if c.fpBytes != 48 || c.id == idBLS12377Gurvy {
		panic("b200: HashToG1 is built for the BLS12-381 curve ids only (SURVEY 8f-4)")
	}
	return &G1{c: c, raw: hashToG1Batch(c.id, [][]byte{data}, domain, c.g1Size())}
}

// HashToG1Batch hashes every message with one domain in a single launch (the BBS verify loop of reference
// perf_test.go:250-261 hashes one message per signature).
func (c *Curve) HashToG1Batch(msgs [][]byte, domain []byte) []byte {
	return hashToG1Batch(c.id, msgs, domain, c.g1Size())
}
func (c *Curve) HashToG2(data []byte) driver.G2                   { panic("b200: HashToG2 is out of scope (SURVEY 8f-4)") }
func (c *Curve) HashToG2WithDomain(data, domain []byte) driver.G2 { panic("b200: HashToG2 is out of scope (SURVEY 8f-4)") }
func (c *Curve) Rand() (io.Reader, error)                          { return rand.Reader, nil }
