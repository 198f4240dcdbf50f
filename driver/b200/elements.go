package b200

import (
	"bytes"
	"encoding/hex"
	"math/big"

	"github.com/IBM/mathlib/driver"
)

// Elements hold the reference's serialized form (what Bytes() returns), so Bytes() is free and batches are
// contiguous slabs.  Scalars are *common.BaseZr exactly as in the gurvy BN254 / kilic drivers
// (reference driver/gurvy/bn254.go:51), supplied by the embedded common.CurveBase.

// G1 implements driver.G1 (reference driver/math.go:249-288).
type G1 struct {
	c   *Curve
	raw []byte // X || Y big-endian, infinity flagged (SURVEY A.3)
}

func (g *G1) Bytes() []byte      { return append([]byte(nil), g.raw...) }
func (g *G1) Compressed() []byte { return pointCodec(g.c.id, 0, 1, 1, g.raw, g.c.fpBytes, 0) }
func (g *G1) String() string     { return hex.EncodeToString(g.raw) }
func (g *G1) Copy() driver.G1    { return &G1{c: g.c, raw: append([]byte(nil), g.raw...)} }
func (g *G1) Clone(a driver.G1)  { g.raw = append(g.raw[:0], a.(*G1).raw...) }
func (g *G1) Equals(a driver.G1) bool { return bytes.Equal(g.raw, a.(*G1).raw) }
func (g *G1) IsInfinity() bool {
	for i, b := range g.raw {
		if i == 0 {
			b &^= 0x40
		}
		if b != 0 {
			return false
		}
	}
	return true
}

// Mul returns [a]g; the receiver is untouched (reference math_test.go:93-95).  n == 1 case of b200_g1_mul_batch.
func (g *G1) Mul(a driver.Zr) driver.G1 {
	return &G1{c: g.c, raw: g1MulBatch(g.c.id, 1, g.raw, a.Bytes(), 0)}
}

// Mul2 returns [e]g + [f]Q (reference driver/math.go:263).
func (g *G1) Mul2(e driver.Zr, Q driver.G1, f driver.Zr) driver.G1 {
	return &G1{c: g.c, raw: g1Mul2Batch(g.c.id, 1, g.raw, e.Bytes(), Q.(*G1).raw, f.Bytes(), 0)}
}

// Mul2InPlace stores [e]g + [f]Q in the receiver (reference driver/math.go:266).
func (g *G1) Mul2InPlace(e driver.Zr, Q driver.G1, f driver.Zr) {
	g.raw = g1Mul2Batch(g.c.id, 1, g.raw, e.Bytes(), Q.(*G1).raw, f.Bytes(), 0)
}

// Add / Sub / Neg mutate the receiver (reference driver/math.go:256-258).
func (g *G1) Add(a driver.G1) {
	g.raw = g1Sum(g.c.id, 2, append(append([]byte(nil), g.raw...), a.(*G1).raw...), g.c.g1Size())
}
func (g *G1) Neg() {
	if g.IsInfinity() {
		return
	}
	n := g.c.fpBytes
	y := new(big.Int).SetBytes(g.raw[n:])
	y.Sub(fieldModulus(g.c.id), y)
	y.Mod(y, fieldModulus(g.c.id))
	y.FillBytes(g.raw[n:])
}
func (g *G1) Sub(a driver.G1) {
	t := a.Copy().(*G1)
	t.Neg()
	g.Add(t)
}

// G2 implements driver.G2 (reference driver/math.go:299-329).  Group operations are the n == 1 (Mul) / n == 2 (Add)
// cases of b200_g2_mul_batch / b200_g2_sum (SURVEY 8f-3).
type G2 struct {
	c   *Curve
	raw []byte // X.A1 || X.A0 || Y.A1 || Y.A0
}

func (g *G2) Bytes() []byte           { return append([]byte(nil), g.raw...) }
func (g *G2) Compressed() []byte      { return pointCodec(g.c.id, 1, 1, 1, g.raw, 2*g.c.fpBytes, 0) }
func (g *G2) String() string          { return hex.EncodeToString(g.raw) }
func (g *G2) Copy() driver.G2         { return &G2{c: g.c, raw: append([]byte(nil), g.raw...)} }
func (g *G2) Clone(a driver.G2)       { g.raw = append(g.raw[:0], a.(*G2).raw...) }
func (g *G2) Equals(a driver.G2) bool { return bytes.Equal(g.raw, a.(*G2).raw) }
func (g *G2) Affine()                 {}

// Mul returns [a]g; the receiver is untouched (reference driver/math.go:307).
func (g *G2) Mul(a driver.Zr) driver.G2 {
	return &G2{c: g.c, raw: g2MulBatch(g.c.id, 1, g.raw, a.Bytes(), 0)}
}

// Add / Sub mutate the receiver (reference driver/math.go:310-313).
func (g *G2) Add(a driver.G2) {
	g.raw = g2Sum(g.c.id, 2, append(append([]byte(nil), g.raw...), a.(*G2).raw...), g.c.g2Size())
}
func (g *G2) Sub(a driver.G2) {
	// -(x, y) = (x, -y): negate both Fp components of Y (bytes 2n..4n hold Y.A1 || Y.A0)
	t := a.Copy().(*G2)
	n := g.c.fpBytes
	zero := true
	for i, b := range t.raw {
		if (i == 0 && b&^0x40 != 0) || (i != 0 && b != 0) {
			zero = false
		}
	}
	if !zero {
		for k := 2; k < 4; k++ {
			y := new(big.Int).SetBytes(t.raw[k*n : (k+1)*n])
			y.Sub(fieldModulus(g.c.id), y)
			y.Mod(y, fieldModulus(g.c.id))
			y.FillBytes(t.raw[k*n : (k+1)*n])
		}
	}
	g.Add(t)
}

// Gt implements driver.Gt (reference driver/math.go:339-360).
type Gt struct {
	c   *Curve
	raw []byte // 12 Fp, C1.B2.A1 first
}

func (g *Gt) Bytes() []byte           { return append([]byte(nil), g.raw...) }
func (g *Gt) ToString() string        { return hex.EncodeToString(g.raw) }
func (g *Gt) Equals(a driver.Gt) bool { return bytes.Equal(g.raw, a.(*Gt).raw) }
func (g *Gt) IsUnity() bool {
	for i, b := range g.raw {
		if (i == len(g.raw)-1 && b != 1) || (i != len(g.raw)-1 && b != 0) {
			return false
		}
	}
	return true
}

// Inverse / Mul mutate the receiver, Exp returns a fresh value (reference driver/math.go:344-359); n == 1 cases of
// b200_gt_inv_batch / b200_gt_mul_batch / b200_gt_exp_batch (SURVEY 8f-3).
func (g *Gt) Inverse()        { g.raw = gtInvBatch(g.c.id, 1, g.raw, 0) }
func (g *Gt) Mul(a driver.Gt) { g.raw = gtMulBatch(g.c.id, 1, g.raw, a.(*Gt).raw, 0) }
// Exp: the device ladder takes a 32-byte exponent.  z.Bytes() is the scalar reduced to [0, r) (reference
// driver/common/big.go:101-113), whereas the reference drivers exponentiate by the raw big.Int (bn254.go:187-191).  The two
// agree on every element of order r -- i.e. every Gt a protocol sees after FExp; they differ only for raw (pre-FExp)
// Miller values combined with exponents >= r or negative, which no reference test or caller forms.
func (g *Gt) Exp(z driver.Zr) driver.Gt {
	return &Gt{c: g.c, raw: gtExpBatch(g.c.id, 1, g.raw, z.Bytes(), 0)}
}
