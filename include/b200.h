/* libb200math -- C ABI of the B200 (sm_100a) backend for IBM/mathlib's batched pairing and G1 hot path.
 *
 * This is the drop-in boundary: what a `driver/b200` Go package binds through cgo (see INTEGRATION.md and
 * driver/b200/*.go) to implement mathlib's `driver.Curve` / `driver.G1` hot-path methods
 * (reference driver/math.go:49-180 and :249-288).  Plain pointers and sizes only; no torch / CUDA types.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative B200_ERR_* code otherwise; b200_last_error() gives the
 *    thread-local message.  The library never aborts the process; the Go wrapper turns a non-zero return
 *    into the same panic the reference drivers raise (reference driver/gurvy/bn254.go:249-251).
 *  - Every function is re-entrant; concurrent calls that share read-only inputs are legal (the reference
 *    benchmarks call one curve from many goroutines: perf_test.go:392-405).  b200_init / b200_shutdown must not run
 *    concurrently with other calls.
 *  - Host buffers belong to the caller and are only read / written during the call (cgo pointer rules).
 *  - `curve` is the mathlib CurveID (reference math.go:70-103): 1 BN254, 3 BLS12_381 (kilic semantics),
 *    4 BLS12_377_GURVY, 5 BLS12_381_GURVY, 6 BLS12_381_BBS (kilic semantics), 7 BLS12_381_BBS_GURVY.
 *    "kilic semantics": Pairing/Pairing2 include the final exponentiation and FExp is the identity
 *    (reference driver/kilic/bls12-381.go:260-281); gurvy: Pairing* is the raw Miller value and FExp
 *    exponentiates (reference driver/gurvy/bls12381/bls12-381.go:448-468).
 *
 * Element encodings (selected per call with B200_IN_* / B200_OUT_* flags)
 *  - BYTES (default): exactly what the reference's Bytes() returns / New*FromBytes() accepts
 *      Fp   : FpBytes big-endian canonical (32 for BN254, 48 for BLS12-381/377)
 *      G1   : X || Y                      (reference bn254.go:76-80; kilic/bls12-381.go:74-78), infinity flagged
 *      G2   : X.A1 || X.A0 || Y.A1 || Y.A0 (reference bn254.go:147-151)
 *      Gt   : 12 Fp, C1.B2.A1 first ... C0.B0.A0 last (reference bn254.go:216-220, kilic/bls12-381.go:224-231)
 *      Zr   : 32 bytes big-endian (reference driver/common/big.go:101-113); any value < 2^256 is accepted and
 *             used mod r
 *  - MONT: in-memory layout of gnark-crypto / kilic elements -- little-endian 32-bit limbs (== 64-bit limbs on
 *      little-endian hosts) in Montgomery form, R = 2^256 / 2^384 (reference driver/gurvy/custom.go:30-40 shows the
 *      two libraries share it).  G1 = X|Y, G2 = X.A0|X.A1|Y.A0|Y.A1, Gt = C0.B0.A0|C0.B0.A1|C0.B1.A0|...|C1.B2.A1,
 *      infinity = all zero.  Scalars stay 32-byte big-endian.
 */
#ifndef B200_H
#define B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* curve ids == mathlib CurveID (reference math.go:70-103) */
#define B200_BN254 1
#define B200_BLS12_381 3
#define B200_BLS12_377_GURVY 4
#define B200_BLS12_381_GURVY 5
#define B200_BLS12_381_BBS 6
#define B200_BLS12_381_BBS_GURVY 7

/* flags */
#define B200_FEXP 0x1u           /* pairing*: also apply the final exponentiation (no-op for kilic ids) */
#define B200_IN_MONT 0x2u        /* group/Gt inputs are MONT limbs instead of BYTES */
#define B200_OUT_MONT 0x4u       /* group/Gt outputs are MONT limbs instead of BYTES */
#define B200_OUT_UNITY_ONLY 0x8u /* pairing / fexp: write one byte per item: 1 if the result is 1 (Gt.IsUnity), else 0 */
#define B200_DEVICE_PTRS 0x10u   /* all buffers are device pointers on the current device (no copies, async on the
                                    stream set with b200_set_stream; caller synchronises).  The call returns before its
                                    kernels ran, so a rejected input cannot fail it with B200_ERR_ENCODING: see
                                    b200_take_error(). */
#define B200_BASES_TABLES 0x20u  /* b200_bases_upload: also store 2^(c*w) * P_i for every Pippenger window w (W x the
                                    memory, one-time cost of c*(W-1) doublings per point); MSMs against the handle then
                                    skip the Horner tail.  Ignored (plain bases kept) when the table would exceed 1/4 of
                                    the device memory. */
#define B200_NO_SUBGROUP_CHECK 0x40u /* decompress / validate: skip the [r]P == O test (on-curve only) */

/* error codes */
#define B200_OK 0
#define B200_ERR_CUDA -1     /* CUDA runtime error (message has the cudaError string) */
#define B200_ERR_ARG -2      /* bad curve id / null pointer / unsupported flag combination */
#define B200_ERR_ENCODING -3 /* an input coordinate is not a canonical field element (>= p) */
#define B200_ERR_NOGPU -4    /* no CUDA device / library not initialised: there is NO CPU fallback */

/* Library lifetime.  device_mask bit i selects CUDA device i (0 = all visible devices). */
int b200_init(uint32_t device_mask);
void b200_shutdown(void);
const char* b200_last_error(void);
int b200_device_count(void);
/* Per calling thread: device used by subsequent calls (default: first device of the mask; batch calls with host
   buffers split across ALL devices of the mask when b200_set_device was never called on the thread). */
int b200_set_device(int device);
/* Per calling thread: CUDA stream (cudaStream_t as void*) used with B200_DEVICE_PTRS; NULL = default stream. */
int b200_set_stream(void* cuda_stream);

/* Error reporting of B200_DEVICE_PTRS calls.  With host buffers a non-canonical coordinate (or any other rejected
   input) fails the call with B200_ERR_ENCODING.  With device pointers the failure is PER ITEM: the offending item's
   output is written as all-zero bytes (verdict 0 with B200_OUT_UNITY_ONLY -- never a stale or passing value), every
   other item of the batch is computed normally, and a per-device flag is raised.  b200_take_error synchronises the
   calling thread's stream, stores 1 in *had_error if any B200_DEVICE_PTRS call on the current device raised the flag
   since the last take (0 otherwise) and clears it.  Row indices of b200_pairing*_fixed_batch outside the table are
   reported the same way. */
int b200_take_error(int* had_error);

/* sizes in bytes of the BYTES / MONT encodings for a curve (FpBytes = 32 or 48) */
int b200_fp_bytes(int curve);

/* driver.Curve.Pairing(G2, G1) for n independent pairs (reference driver/math.go:51).
   g1: n G1 elements, g2: n G2 elements, gt_out: n Gt elements (or n bytes with B200_OUT_UNITY_ONLY). */
int b200_pairing_batch(int curve, size_t n, const void* g1, const void* g2, void* gt_out, uint32_t flags);

/* driver.Curve.Pairing2(p2a, p2b, p1a, p1b) = e(p2a,p1a)*e(p2b,p1b) with one shared squaring chain
   (reference driver/math.go:54; facade argument reorder math.go:869-871). */
int b200_pairing2_batch(int curve, size_t n, const void* g1a, const void* g2a, const void* g1b, const void* g2b,
                        void* gt_out, uint32_t flags);

/* driver.Curve.FExp (reference driver/math.go:57).  Identity copy for kilic-semantics ids. */
int b200_fexp_batch(int curve, size_t n, const void* gt_in, void* gt_out, uint32_t flags);

/* driver.G1.Mul (reference driver/math.go:260): out[i] = [scalars[i]] pts[i], affine. */
int b200_g1_mul_batch(int curve, size_t n, const void* pts, const void* scalars_be32, void* out, uint32_t flags);

/* driver.G1.Mul2 / Mul2InPlace (reference driver/math.go:263-266): out[i] = [e[i]]P[i] + [f[i]]Q[i]. */
int b200_g1_mul2_batch(int curve, size_t n, const void* P, const void* e_be32, const void* Q, const void* f_be32,
                       void* out, uint32_t flags);

/* driver.Curve.MultiScalarMul (reference driver/math.go:170): out = sum_i [scalars[i]] pts[i], one affine G1.
   n == 0 gives infinity.  Scalars are any 256-bit values (reduced mod r on the device).  The points must lie in the
   prime-order subgroup -- what mathlib's deserialisers guarantee and gnark's MultiExp assumes as well: on the BLS12 curves
   the scalars are split with the GLV endomorphism (k = k1 + lambda k2, [lambda](x, y) = (beta x, y)). */
int b200_g1_msm(int curve, size_t n, const void* pts, const void* scalars_be32, void* out, uint32_t flags);

/* Resident base points: upload once, run many MSMs against them (scalars only cross PCIe). */
int b200_bases_upload(int curve, size_t n, const void* pts, uint32_t flags, uint64_t* handle);
int b200_g1_msm_resident(uint64_t handle, size_t n, const void* scalars_be32, void* out, uint32_t flags);
int b200_bases_free(uint64_t handle);

/* sum of n G1 points (used to combine per-GPU MSM partial sums; n is small). */
int b200_g1_sum(int curve, size_t n, const void* pts, void* out, uint32_t flags);

/* ---- callers next to the hot path; same encodings, flags and error behaviour as above ---- */

/* driver.G2.Mul(Zr) for n independent (point, scalar) pairs (reference driver/math.go:307; impls bn254.go:134-139,
   bls12-377.go:131-136, bls12381/bls12-381.go:342-351, kilic/bls12-381.go:127-137).  Output: affine G2 elements. */
int b200_g2_mul_batch(int curve, size_t n, const void* g2_pts, const void* scalars_be32, void* out, uint32_t flags);
/* sum of n G2 points; n = 2 is driver.G2.Add (reference driver/math.go:310; bn254.go:141, kilic/bls12-381.go:139). */
int b200_g2_sum(int curve, size_t n, const void* g2_pts, void* out, uint32_t flags);
/* driver.Gt.Mul / Gt.Inverse / Gt.Exp(Zr) batches (reference driver/math.go:339-360; impls bn254.go:187-203,
   bls12-377.go:184-200, bls12381/bls12-381.go:399-419, kilic/bls12-381.go:185-210).  Exponents are 32 bytes big-endian, used as given. */
int b200_gt_mul_batch(int curve, size_t n, const void* gt_a, const void* gt_b, void* gt_out, uint32_t flags);
int b200_gt_inv_batch(int curve, size_t n, const void* gt_a, void* gt_out, uint32_t flags);
int b200_gt_exp_batch(int curve, size_t n, const void* gt_a, const void* scalars_be32, void* gt_out, uint32_t flags);

/* ---- fixed-Q pairings: the BLS / BBS verification pattern, where the G2 arguments are a small set of
   public keys and the generator (reference perf_test.go:248-259).  The G2 side of the Miller loop is computed once per
   point into a resident line table; pairings against table rows then do no G2 arithmetic.  Results are bit-identical to
   b200_pairing_batch / b200_pairing2_batch on the same points (same driver semantics per curve id, same flags).
   q_idx / qa_idx / qb_idx: one uint32 row index per item; NULL = row 0 (pair a) and row 1 (pair b) for every item. */
int b200_g2_lines_upload(int curve, size_t n_q, const void* g2_pts, uint32_t flags, uint64_t* handle);
int b200_g2_lines_free(uint64_t handle);
int b200_pairing_fixed_batch(uint64_t lines, size_t n, const void* g1, const uint32_t* q_idx, void* gt_out, uint32_t flags);
int b200_pairing2_fixed_batch(uint64_t lines, size_t n, const void* g1a, const uint32_t* qa_idx, const void* g1b,
                              const uint32_t* qb_idx, void* gt_out, uint32_t flags);

/* ---- point (de)serialisation and validation for whole batches ----
   NewG1FromCompressed / NewG2FromCompressed (reference driver/gurvy/bn254.go:359-377, bls12381/bls12-381.go:551-569,
   kilic/bls12-381.go:370-394): compressed -> uncompressed Bytes() (or MONT limbs with B200_OUT_MONT).  A coordinate >= p,
   an x with no point, wrong flag bits, or (unless B200_NO_SUBGROUP_CHECK) a point outside the order-r subgroup fails the
   call with B200_ERR_ENCODING, as gnark's SetBytes / kilic's FromCompressed fail. */
int b200_g1_decompress_batch(int curve, size_t n, const void* compressed, void* out, uint32_t flags);
int b200_g2_decompress_batch(int curve, size_t n, const void* compressed, void* out, uint32_t flags);
/* G1.Compressed / G2.Compressed (reference bn254.go:82-86,167-171, bls12381/bls12-381.go:292-296,379-383, kilic/bls12-381.go:81-85,159-163). */
int b200_g1_compress_batch(int curve, size_t n, const void* pts, void* compressed_out, uint32_t flags);
int b200_g2_compress_batch(int curve, size_t n, const void* pts, void* compressed_out, uint32_t flags);
/* The checks of NewG1FromBytes / NewG2FromBytes (reference bn254.go:339-357, kilic/bls12-381.go:344-368): one verdict
   byte per point, 1 = canonical coordinates, on the curve and (unless B200_NO_SUBGROUP_CHECK) in the subgroup. */
int b200_g1_validate_batch(int curve, size_t n, const void* pts, void* ok_out, uint32_t flags);
int b200_g2_validate_batch(int curve, size_t n, const void* pts, void* ok_out, uint32_t flags);

/* G2 MultiScalarMul: out = sum_i [k_i] Q_i on the twist E'(Fp2) -- the G2 counterpart of driver.Curve.MultiScalarMul
   (reference driver/math.go:170) built from driver.G2.Mul / driver.G2.Add (reference driver/math.go:307-310; gnark
   G2Affine.MultiExp is what a gurvy-style adapter would forward to).  n G2 elements (G2.Bytes() layout, or MONT with
   B200_IN_MONT) and n 32-byte big-endian scalars -> one G2 element.  Same Pippenger pipeline as b200_g1_msm with the
   XYZZ formulas over Fp2; one device per call (combine per-GPU partial sums with b200_g2_sum_batch-style addition). */
int b200_g2_msm(int curve, size_t n, const void* pts, const void* scalars, void* out, uint32_t flags);

/* Batch affine normalisation (Montgomery's trick, one inversion per 8 points): n Jacobian G1 points given as MONT limbs
   X | Y | Z (kilic PointG1 [3]fe, reference driver/kilic/bls12-381.go:20-23; gnark G1Jac) -> n affine G1 elements
   (X / Z^2, Y / Z^3) in BYTES form (what G1.Bytes() returns, reference kilic/bls12-381.go:74-78) or MONT with
   B200_OUT_MONT.  Z = 0 is the point at infinity. */
int b200_g1_normalize_batch(int curve, size_t n, const void* jacobian_mont, void* out, uint32_t flags);

/* ---- hash-to-G1 for whole batches: driver.Curve.HashToG1 / HashToG1WithDomain
   (reference driver/math.go:120-131).  BLS12-381 curve ids only:
     3, 5  kilic g1.HashToCurve / gnark bls12381.HashToG1 (reference driver/kilic/bls12-381.go:410-447,
           driver/gurvy/bls12381/bls12-381.go:652-677) = RFC 9380 BLS12381G1_XMD:SHA-256_SSWU_RO_ with `domain` as DST;
     6, 7  the BBS variant HashToG1GenericBESwu (reference driver/kilic/custom.go:205-237, driver/gurvy/custom.go:152-193):
           BLAKE2b-512 in expand_message_xmd and the big-endian sign rule.
   Message i is msgs[offsets[i] .. offsets[i+1]) (n + 1 offsets); one domain (<= 255 bytes, may be empty) for the batch.
   out: n affine G1 elements (BYTES, or MONT limbs with B200_OUT_MONT).  Other curve ids fail with B200_ERR_ARG. */
int b200_hash_to_g1_batch(int curve, size_t n, const void* msgs, const uint64_t* offsets, const void* domain,
                          size_t domain_len, void* out, uint32_t flags);

/* Number of kernel launches issued by this library in the calling process since load (bench.py's gpu_launches). */
uint64_t b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_H */
