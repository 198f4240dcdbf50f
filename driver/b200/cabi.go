// Package b200 is a mathlib driver backed by hand-written sm_100a CUDA kernels (libb200math.so).
//
// It implements driver.Curve / driver.G1 / driver.G2 / driver.Gt / driver.Zr (reference driver/math.go:49-360)
// and forwards the data-parallel hot path -- Pairing, Pairing2, FExp, MultiScalarMul, G1.Mul, G1.Mul2 and their
// batch forms -- through cgo to the C ABI declared in include/b200.h.  Everything else (Zr arithmetic, hashing,
// element bookkeeping) is thin host code.
//
// NOTE: this package was written against the header without a Go toolchain (none exists in the build image), so it
// has never been compiled.  It is deliberately mechanical: each method marshals byte slabs and calls one C function.
package b200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../mathlib_b200 -lb200math -Wl,-rpath,${SRCDIR}/../../mathlib_b200
#include <stdlib.h>
#include "b200.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// flags (include/b200.h)
const (
	flagFExp      = 0x1
	flagInMont    = 0x2
	flagOutMont   = 0x4
	flagUnityOnly = 0x8
)

func lastError() string { return C.GoString(C.b200_last_error()) }

// check turns a non-zero return code into the panic the reference drivers raise
// (reference driver/gurvy/bn254.go:249-251: panic(fmt.Sprintf("pairing failed [%s]", err))).
func check(op string, rc C.int) {
	if rc != 0 {
		panic(fmt.Sprintf("%s failed [%s]", op, lastError()))
	}
}

func ptr(b []byte) unsafe.Pointer {
	if len(b) == 0 {
		return nil
	}
	return unsafe.Pointer(&b[0])
}

// Init selects the GPUs (bit i = CUDA device i; 0 = all).  Optional: the first call initialises lazily.
func Init(deviceMask uint32) { check("b200 init", C.b200_init(C.uint32_t(deviceMask))) }

// pairingBatch: n x Pairing(G2,G1); g1 / g2 are contiguous slabs in the reference Bytes() encoding.
func pairingBatch(curve int, n int, g1, g2 []byte, gtSize int, flags uint32) []byte {
	out := make([]byte, n*gtSize)
	check("pairing", C.b200_pairing_batch(C.int(curve), C.size_t(n), ptr(g1), ptr(g2), ptr(out), C.uint32_t(flags)))
	return out
}

func pairing2Batch(curve int, n int, g1a, g2a, g1b, g2b []byte, gtSize int, flags uint32) []byte {
	out := make([]byte, n*gtSize)
	check("pairing 2", C.b200_pairing2_batch(C.int(curve), C.size_t(n), ptr(g1a), ptr(g2a), ptr(g1b), ptr(g2b), ptr(out), C.uint32_t(flags)))
	return out
}

func fexpBatch(curve int, n int, gt []byte, flags uint32) []byte {
	out := make([]byte, len(gt))
	check("final exponentiation", C.b200_fexp_batch(C.int(curve), C.size_t(n), ptr(gt), ptr(out), C.uint32_t(flags)))
	return out
}

func g1MulBatch(curve int, n int, pts, scalars []byte, flags uint32) []byte {
	out := make([]byte, len(pts))
	check("g1 mul", C.b200_g1_mul_batch(C.int(curve), C.size_t(n), ptr(pts), ptr(scalars), ptr(out), C.uint32_t(flags)))
	return out
}

func g1Mul2Batch(curve int, n int, p, e, q, f []byte, flags uint32) []byte {
	out := make([]byte, len(p))
	check("g1 mul2", C.b200_g1_mul2_batch(C.int(curve), C.size_t(n), ptr(p), ptr(e), ptr(q), ptr(f), ptr(out), C.uint32_t(flags)))
	return out
}

func g1Msm(curve int, n int, pts, scalars []byte, g1Size int, flags uint32) []byte {
	out := make([]byte, g1Size)
	check("multi scalar mul", C.b200_g1_msm(C.int(curve), C.size_t(n), ptr(pts), ptr(scalars), ptr(out), C.uint32_t(flags)))
	return out
}

func g1Sum(curve int, n int, pts []byte, g1Size int) []byte {
	out := make([]byte, g1Size)
	check("g1 sum", C.b200_g1_sum(C.int(curve), C.size_t(n), ptr(pts), ptr(out), 0))
	return out
}

// ---- callers next to the hot path (SURVEY 8f-3) ----

func g2MulBatch(curve int, n int, pts, scalars []byte, flags uint32) []byte {
	out := make([]byte, len(pts))
	check("g2 mul", C.b200_g2_mul_batch(C.int(curve), C.size_t(n), ptr(pts), ptr(scalars), ptr(out), C.uint32_t(flags)))
	return out
}

func g2Msm(curve int, n int, pts, scalars []byte, g2Size int, flags uint32) []byte {
	out := make([]byte, g2Size)
	check("g2 multi scalar mul", C.b200_g2_msm(C.int(curve), C.size_t(n), ptr(pts), ptr(scalars), ptr(out), C.uint32_t(flags)))
	return out
}

// g1NormalizeBatch: n Jacobian points as Montgomery limbs X|Y|Z -> n affine Bytes() encodings (one inversion per 8 points).
func g1NormalizeBatch(curve int, n int, jac []byte, g1Size int) []byte {
	out := make([]byte, n*g1Size)
	check("g1 normalize", C.b200_g1_normalize_batch(C.int(curve), C.size_t(n), ptr(jac), ptr(out), 0))
	return out
}

func g2Sum(curve int, n int, pts []byte, g2Size int) []byte {
	out := make([]byte, g2Size)
	check("g2 sum", C.b200_g2_sum(C.int(curve), C.size_t(n), ptr(pts), ptr(out), 0))
	return out
}

func gtMulBatch(curve int, n int, a, b []byte, flags uint32) []byte {
	out := make([]byte, len(a))
	check("gt mul", C.b200_gt_mul_batch(C.int(curve), C.size_t(n), ptr(a), ptr(b), ptr(out), C.uint32_t(flags)))
	return out
}

func gtInvBatch(curve int, n int, a []byte, flags uint32) []byte {
	out := make([]byte, len(a))
	check("gt inverse", C.b200_gt_inv_batch(C.int(curve), C.size_t(n), ptr(a), ptr(out), C.uint32_t(flags)))
	return out
}

func gtExpBatch(curve int, n int, a, scalars []byte, flags uint32) []byte {
	out := make([]byte, len(a))
	check("gt exp", C.b200_gt_exp_batch(C.int(curve), C.size_t(n), ptr(a), ptr(scalars), ptr(out), C.uint32_t(flags)))
	return out
}

// hashToG1Batch: one G1 point per message, one domain separation tag for the batch (SURVEY 8f-4).
func hashToG1Batch(curve int, msgs [][]byte, domain []byte, g1Size int) []byte {
	offs := make([]uint64, len(msgs)+1)
	total := 0
	for i, m := range msgs {
		total += len(m)
		offs[i+1] = uint64(total)
	}
	blob := make([]byte, 0, total)
	for _, m := range msgs {
		blob = append(blob, m...)
	}
	out := make([]byte, len(msgs)*g1Size)
	check("HashToG1", C.b200_hash_to_g1_batch(C.int(curve), C.size_t(len(msgs)), ptr(blob), (*C.uint64_t)(unsafe.Pointer(&offs[0])),
		ptr(domain), C.size_t(len(domain)), ptr(out), 0))
	return out
}

// TakeError reports (and clears) the per-device error flag raised by B200_DEVICE_PTRS calls: those calls are
// asynchronous, so a rejected input zeroes its own item's output instead of failing the call (include/b200.h).
func TakeError() bool {
	var had C.int
	check("take error", C.b200_take_error(&had))
	return had != 0
}

// ---- point (de)serialisation and validation batches (SURVEY 8f-2) ----

// pointCodec: op 0 decompress, 1 compress, 2 validate; outElem = bytes per output element.
func pointCodec(curve, g2, op, n int, in []byte, outElem int, flags uint32) []byte {
	out := make([]byte, n*outElem)
	var rc C.int
	switch {
	case g2 == 0 && op == 0:
		rc = C.b200_g1_decompress_batch(C.int(curve), C.size_t(n), ptr(in), ptr(out), C.uint32_t(flags))
	case g2 == 0 && op == 1:
		rc = C.b200_g1_compress_batch(C.int(curve), C.size_t(n), ptr(in), ptr(out), C.uint32_t(flags))
	case g2 == 0:
		rc = C.b200_g1_validate_batch(C.int(curve), C.size_t(n), ptr(in), ptr(out), C.uint32_t(flags))
	case op == 0:
		rc = C.b200_g2_decompress_batch(C.int(curve), C.size_t(n), ptr(in), ptr(out), C.uint32_t(flags))
	case op == 1:
		rc = C.b200_g2_compress_batch(C.int(curve), C.size_t(n), ptr(in), ptr(out), C.uint32_t(flags))
	default:
		rc = C.b200_g2_validate_batch(C.int(curve), C.size_t(n), ptr(in), ptr(out), C.uint32_t(flags))
	}
	check("set bytes", rc)
	return out
}

// ---- fixed-Q pairings against resident G2 line tables (SURVEY 8f-1) ----

// LineTable is a resident table of Miller-loop line coefficients for a fixed set of G2 points (public keys, the
// generator).  Free releases the device memory; a finalizer is not installed because the table is usually process-wide.
type LineTable struct {
	curve  int
	handle C.uint64_t
	gtSize int
}

// NewLineTable precomputes the G2 side of the Miller loop for the given points (reference Bytes() encoding, concatenated).
func NewLineTable(c *Curve, g2Points []byte) *LineTable {
	var h C.uint64_t
	n := len(g2Points) / c.g2Size()
	check("line table", C.b200_g2_lines_upload(C.int(c.id), C.size_t(n), ptr(g2Points), 0, &h))
	return &LineTable{curve: c.id, handle: h, gtSize: c.gtSize()}
}

func (t *LineTable) Free() { check("line table free", C.b200_g2_lines_free(t.handle)) }

func idxPtr(rows []uint32) *C.uint32_t {
	if len(rows) == 0 {
		return nil
	}
	return (*C.uint32_t)(unsafe.Pointer(&rows[0]))
}

// Pairing2Batch: n x e(rowA_i, p1a_i) * e(rowB_i, p1b_i); nil row slices mean row 0 / row 1 for every check.
// With verdictOnly the result is one byte per check (Gt.IsUnity after FExp).
func (t *LineTable) Pairing2Batch(n int, g1a []byte, rowsA []uint32, g1b []byte, rowsB []uint32, flags uint32) []byte {
	size := t.gtSize
	if flags&flagUnityOnly != 0 {
		size = 1
	}
	out := make([]byte, n*size)
	check("pairing 2", C.b200_pairing2_fixed_batch(t.handle, C.size_t(n), ptr(g1a), idxPtr(rowsA), ptr(g1b), idxPtr(rowsB),
		ptr(out), C.uint32_t(flags)))
	return out
}

func (t *LineTable) PairingBatch(n int, g1 []byte, rows []uint32, flags uint32) []byte {
	size := t.gtSize
	if flags&flagUnityOnly != 0 {
		size = 1
	}
	out := make([]byte, n*size)
	check("pairing", C.b200_pairing_fixed_batch(t.handle, C.size_t(n), ptr(g1), idxPtr(rows), ptr(out), C.uint32_t(flags)))
	return out
}
