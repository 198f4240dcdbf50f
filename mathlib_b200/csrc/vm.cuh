// Interpreter of the warp-cooperative pairing VM (microcode compiled by mathlib_b200/vm/compiler.py).
//
// A group of VM_G = 6 lanes owns one pairing product.  All Fp2 values live in a per-group slot file in shared
// memory; each lane executes one op per phase:
//   DOT : up to 6 Fp2 products accumulated UNREDUCED in two wide register accumulators (real / imaginary part,
//         Karatsuba: 3 wide N x N multiplies per product), then ONE Montgomery reduction per part -- "lazy
//         reduction" over a whole dot product, so an Fp12 multiplication costs 6x(18 wide mul + 2 redc) instead
//         of 54 full Montgomery multiplications, and every lane of the group is busy;
//   LIN : small-integer linear combination (no multiplier use);  INV : Fp2 inversion on one lane.
// Host-emulable like fp.cuh (tests/hostemu runs the lanes of a phase one after another).
#pragma once
#include "tower.cuh"
#include "microcode.h"

namespace b200 {

enum : uint32_t { VM_NEG = 1, VM_CONJ = 2, VM_XI = 4, VM_DBL = 8, VM_REAL0 = 4, VM_REAL1 = 8 };
enum : uint32_t { VM_C_ABS = 0, VM_C_B1 = 1, VM_C_B2 = 2, VM_C_B3 = 3, VM_C_CONST = 4 };
enum : uint32_t { VM_NOP = 0, VM_DOT = 1, VM_LIN = 2, VM_INV = 3 };

template <class C>
struct Vm {
    static constexpr int N = C::N;
    static constexpr int W = 2 * N + 1;          // wide accumulator words
    static constexpr int SLOT_WORDS = 2 * N;
    typedef FpOps<C> F;
    typedef Tower<C> T;
    typedef Fp<N> E1;
    typedef Fp2<N> E2;

    struct Ctx {
        uint32_t* slots;            // this group's slot file (shared memory)
        const uint32_t* kbank;      // constant bank (Fp2 each, Montgomery)
        uint32_t base[3];           // B1, B2, B3 (slot units)
        uint32_t live;              // bit k: pair k is live
    };

    // slots are 16-byte aligned (slot = 2N words, N a multiple of 4): 128-bit accesses
    struct alignas(16) Q4 { uint32_t x, y, z, w; };
    static B200_HD void load2(E2& r, const uint32_t* p) {
        const Q4* q = reinterpret_cast<const Q4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            Q4 a = q[i], b = q[N / 4 + i];
            r.c0.l[4 * i] = a.x; r.c0.l[4 * i + 1] = a.y; r.c0.l[4 * i + 2] = a.z; r.c0.l[4 * i + 3] = a.w;
            r.c1.l[4 * i] = b.x; r.c1.l[4 * i + 1] = b.y; r.c1.l[4 * i + 2] = b.z; r.c1.l[4 * i + 3] = b.w;
        }
    }
    static B200_HD void store2(uint32_t* p, const E2& r) {
        Q4* q = reinterpret_cast<Q4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            Q4 a = {r.c0.l[4 * i], r.c0.l[4 * i + 1], r.c0.l[4 * i + 2], r.c0.l[4 * i + 3]};
            Q4 b = {r.c1.l[4 * i], r.c1.l[4 * i + 1], r.c1.l[4 * i + 2], r.c1.l[4 * i + 3]};
            q[i] = a; q[N / 4 + i] = b;
        }
    }
    // modifiers in the order CONJ, XI, DBL, NEG; result fully reduced
    static B200_HD void apply_mod(E2& x, uint32_t m) {
        if (m & VM_CONJ) F::neg(x.c1, x.c1);
        if (m & VM_XI) T::f2_mul_xi(x, x);
        if (m & VM_DBL) T::f2_dbl(x, x);
        if (m & VM_NEG) T::f2_neg(x, x);
    }

    // ---------------------------------------------------------------------------------------------------
    // wide arithmetic
    // ---------------------------------------------------------------------------------------------------
    static constexpr int AW = 2 * N + 2;         // words of an even / odd accumulator array (redc)
    // T[0..2N) += a * b, in place.  Operand scanning with the partial products split by the parity of their word position:
    // a[j]*b[i] lands on word i+j; every 32x32 product is one IMAD.WIDE.U32.X inside a carry chain (two independent chains
    // per row) -- one instruction per multiply-accumulate, against two for a product-scanning column sum.  T itself is the
    // even-aligned accumulator (its words may be full: every chain below is a carry chain over the words it touches); the
    // odd-aligned partial sums go to a fresh Od (Od[k] = word k + 1) that is merged once at the end.  The carry that leaves
    // the top of a row's chain -- T's or Od's -- is deposited in Od one word above the chain: up there Od holds nothing but
    // earlier carries, so the deposit cannot overflow and the next rows' chains absorb it.  144 IMAD.WIDE + 22 deposits +
    // 23 merge additions per product; round 1 formed the product in a separate array and added it (25 more instructions,
    // 24 more registers).  The sum stays below 2^(64N) by the compiler's bounds (dot_bounds): nothing leaves word 2N - 1.
    static B200_HD void wide_mac(uint32_t* T, const uint32_t* a, const uint32_t* b) {
        uint32_t Od[2 * N - 1];
        // row 0: T's chain accumulates, Od starts from plain products
        T[0] = mad_lo_cc(a[0], b[0], T[0]);
        T[1] = madc_hi_cc(a[0], b[0], T[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            T[j] = madc_lo_cc(a[j], b[0], T[j]);
            T[j + 1] = madc_hi_cc(a[j], b[0], T[j + 1]);
        }
        const uint32_t c0 = addc(0, 0);               // word N
#pragma unroll
        for (int j = 1; j < N; j += 2) {
            Od[j - 1] = mul_lo(a[j], b[0]);
            Od[j] = mul_hi(a[j], b[0]);
        }
        Od[N - 1] += c0;                              // hi <= 2^32 - 2: no overflow
#pragma unroll
        for (int k = N; k < 2 * N - 1; k++) Od[k] = 0;
#pragma unroll
        for (int i = 1; i < N; i++) {
            const int je = i & 1;          // first j with (i+j) even
            const int jo = je ^ 1;         // first j with (i+j) odd
            {
                const int w0 = i + je;
                T[w0] = mad_lo_cc(a[je], b[i], T[w0]);
                T[w0 + 1] = madc_hi_cc(a[je], b[i], T[w0 + 1]);
#pragma unroll
                for (int j = je + 2; j < N; j += 2) {
                    T[i + j] = madc_lo_cc(a[j], b[i], T[i + j]);
                    T[i + j + 1] = madc_hi_cc(a[j], b[i], T[i + j + 1]);
                }
                if (w0 + N - 1 < 2 * N - 1) Od[w0 + N - 1] = addc(Od[w0 + N - 1], 0);        // word w0 + N
            }
            {
                const int w0 = i + jo - 1;
                Od[w0] = mad_lo_cc(a[jo], b[i], Od[w0]);
                Od[w0 + 1] = madc_hi_cc(a[jo], b[i], Od[w0 + 1]);
#pragma unroll
                for (int j = jo + 2; j < N; j += 2) {
                    Od[i + j - 1] = madc_lo_cc(a[j], b[i], Od[i + j - 1]);
                    Od[i + j] = madc_hi_cc(a[j], b[i], Od[i + j]);
                }
                if (w0 + N < 2 * N - 1) Od[w0 + N] = addc(Od[w0 + N], 0);
            }
        }
        T[1] = add_cc(T[1], Od[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; k++) T[k] = addc_cc(T[k], Od[k - 1]);
        T[2 * N - 1] = addc(T[2 * N - 1], Od[2 * N - 2]);
    }
    // one half (N words) of a slot
    static B200_HD void load1(E1& r, const uint32_t* p) {
        const Q4* q = reinterpret_cast<const Q4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            Q4 a = q[i];
            r.l[4 * i] = a.x; r.l[4 * i + 1] = a.y; r.l[4 * i + 2] = a.z; r.l[4 * i + 3] = a.w;
        }
    }
    // plain integer helpers on N words (no reduction): the lazy operand modifiers below
    static B200_HD void add_nr(E1& r, const E1& a, const E1& b) {
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
        r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);
    }
    static B200_HD void p_minus(E1& r, const E1& a) {       // p - a for a <= p
        const uint32_t* p = C::p();
        r.l[0] = sub_cc(p[0], a.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(p[i], a.l[i]);
        r.l[N - 1] = subc(p[N - 1], a.l[N - 1]);
    }
    static B200_HD void p_minus_c(E1& x) { p_minus(x, x); }
    static B200_HD void shl1(E1& r) {
#pragma unroll
        for (int i = N - 1; i > 0; i--) r.l[i] = (r.l[i] << 1) | (r.l[i - 1] >> 31);
        r.l[0] <<= 1;
    }
    // Operand modifiers WITHOUT reduction (curves whose modulus leaves >= 3 spare bits in N words and xi = 1 + u): the halves
    // stay plain integers below 4p (compiler.operand_bounds tracks the bound; the wide accumulators and the offset k p^2
    // are sized for it).  Negations first -- they commute with the linear maps that follow and p - x keeps the bound 1:
    //   NEG: (p - x0, p - x1)   CONJ: (x0, p - x1)   XI: (x0 - x1 + p, x0 + x1)   DBL: (2 x0, 2 x1)
    static B200_HD void lazy_mods(E1& a0, E1& a1, uint32_t m) {
        if (m & VM_NEG) p_minus(a0, a0);
        if (((m & VM_NEG) != 0) != ((m & VM_CONJ) != 0)) p_minus(a1, a1);
        if (m & VM_XI) {
            E1 s;
            add_nr(s, a0, a1);
            const uint32_t* p = C::p();                 // (a0 + p) - a1: two plain chains, no complement of a1 needed
            a0.l[0] = add_cc(a0.l[0], p[0]);
#pragma unroll
            for (int i = 1; i < N - 1; i++) a0.l[i] = addc_cc(a0.l[i], p[i]);
            a0.l[N - 1] = addc(a0.l[N - 1], p[N - 1]);
            a0.l[0] = sub_cc(a0.l[0], a1.l[0]);
#pragma unroll
            for (int i = 1; i < N - 1; i++) a0.l[i] = subc_cc(a0.l[i], a1.l[i]);
            a0.l[N - 1] = subc(a0.l[N - 1], a1.l[N - 1]);
            a1 = s;
        }
        if (m & VM_DBL) { shl1(a0); shl1(a1); }
    }
    // Montgomery reduction of a wide value T < 4 p R (T[0..2N), top word zero) -> canonical residue T / R mod p.
    // Word-sliding reduction on an even / odd split of T: no data movement, the window offsets are compile-time.
    // levels: the compiler's bound on the result before canonicalisation -- T < (2^levels - 1) p R, so T/R + p < 2^levels p
    // and `levels` conditional subtractions (of 4p, 2p, p) make it canonical (microcode header bits 18..19, per phase:
    // vm/compiler.py dot_bounds)
    static B200_HD void redc(E1& r, const uint32_t* Tw, uint32_t levels) {
        uint32_t X[AW], Y[AW];
#pragma unroll
        for (int k = 0; k < 2 * N; k += 2) { X[k] = Tw[k]; X[k + 1] = 0; Y[k] = Tw[k + 1]; Y[k + 1] = 0; }
        X[2 * N] = 0; X[2 * N + 1] = 0; Y[2 * N] = 0; Y[2 * N + 1] = 0;
        const uint32_t* p = C::p();
        {
            uint32_t m = mul_lo(X[0], C::inv32());
            chain_odd_full(Y, p, m, false);
            chain_even_full(X, p, m);
        }
        // after row i the even accumulator's word 0 is zero; its word 1 is folded into the other accumulator and
        // the roles swap with the old even accumulator sliding up by two words
#pragma unroll
        for (int i = 1; i < N; i++) {
            const int ao = 2 * ((i - 1) / 2) + ((i & 1) ? 0 : 0);
            if (i & 1) {      // A = X (window at xo), B = Y (window at yo)
                const int xo = i - 1, yo = i - 1;
                Y[yo] = add_cc(Y[yo], X[xo + 1]);
                uint32_t m = mul_lo(Y[yo], C::inv32());
                chain_odd_full(X + xo + 2, p, m, true);
                chain_even_full(Y + yo, p, m);
            } else {          // A = Y, B = X
                const int yo = i - 2, xo = i;
                X[xo] = add_cc(X[xo], Y[yo + 1]);
                uint32_t m = mul_lo(X[xo], C::inv32());
                chain_odd_full(Y + yo + 2, p, m, true);
                chain_even_full(X + xo, p, m);
            }
            (void)ao;
        }
        // N rows done (N even): the last row had A = X at offset N-2 -> X slid to N, B = Y at offset N-2.
        // result word k = Yw[k+1] + Xw[k] with Yw = Y + (N-2) (even accumulator), Xw = X + N (odd accumulator)
        const uint32_t* Yw = Y + (N - 2);
        const uint32_t* Xw = X + N;
        r.l[0] = add_cc(Xw[0], Yw[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(Xw[k], Yw[k + 1]);
        r.l[N - 1] = addc(Xw[N - 1], Yw[N]);
        if (levels >= 3) cond_sub_kp(r, 2);
        if (levels >= 2) cond_sub_kp(r, 1);
        cond_sub_kp(r, 0);
    }
    // acc (even aligned, full-size words above) += v_even * m, carry rippled one word up
    static B200_HD void chain_even_full(uint32_t* acc, const uint32_t* v, uint32_t m) {
        acc[0] = mad_lo_cc(v[0], m, acc[0]);
        acc[1] = madc_hi_cc(v[0], m, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = madc_lo_cc(v[j], m, acc[j]);
            acc[j + 1] = madc_hi_cc(v[j], m, acc[j + 1]);
        }
        acc[N] = addc_cc(acc[N], 0);
        acc[N + 1] = addc(acc[N + 1], 0);
    }
    // acc (odd aligned) += v_odd * m; optional carry-in from the preceding stray-word addition
    static B200_HD void chain_odd_full(uint32_t* acc, const uint32_t* v, uint32_t m, bool carry_in) {
        if (carry_in) acc[0] = madc_lo_cc(v[1], m, acc[0]); else acc[0] = mad_lo_cc(v[1], m, acc[0]);
        acc[1] = madc_hi_cc(v[1], m, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = madc_lo_cc(v[j + 1], m, acc[j]);
            acc[j + 1] = madc_hi_cc(v[j + 1], m, acc[j + 1]);
        }
        acc[N] = addc_cc(acc[N], 0);
        acc[N + 1] = addc(acc[N + 1], 0);
    }
    // r -= (p << sh) if r >= (p << sh), sh = 0, 1, 2 (2p and 4p come from the constant table)
    static B200_HD void cond_sub_kp(E1& r, int sh) {
        const uint32_t* q = sh == 0 ? C::p() : C::K().pk + (sh - 1) * N;
        uint32_t t[N];
        t[0] = sub_cc(r.l[0], q[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(r.l[i], q[i]);
        uint32_t borrow = subc(0, 0);
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? r.l[i] : t[i];
    }

    // x * c mod p for a small positive c (1..15)
    static B200_HD void mul_small(E1& r, const E1& x, uint32_t c) {
        E1 acc = x;
        int top = 3;
        while (top > 0 && !((c >> top) & 1)) top--;
        for (int b = top - 1; b >= 0; b--) {
            F::dbl(acc, acc);
            if ((c >> b) & 1) F::add(acc, acc, x);
        }
        r = acc;
    }

    // ---------------------------------------------------------------------------------------------------
    // one op
    // ---------------------------------------------------------------------------------------------------
    // acc +/-= v over 2N words (no carry out by construction)
    static B200_HD void wide_add2n(uint32_t* acc, const uint32_t* v) {
        acc[0] = add_cc(acc[0], v[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) acc[k] = addc_cc(acc[k], v[k]);
        acc[2 * N - 1] = addc(acc[2 * N - 1], v[2 * N - 1]);
    }
    static B200_HD void wide_sub2n(uint32_t* acc, const uint32_t* v) {
        acc[0] = sub_cc(acc[0], v[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) acc[k] = subc_cc(acc[k], v[k]);
        acc[2 * N - 1] = subc(acc[2 * N - 1], v[2 * N - 1]);
    }
    // v (2N words) *= c for the small scales the programs use (1, 2, 3, 4, 6); the compiler's bound keeps c v below 2^(64N)
    static B200_HD void wide_scale(uint32_t* v, uint32_t c) {
        if (c == 3 || c == 6) {
            uint32_t d[2 * N];
#pragma unroll
            for (int k = 2 * N - 1; k > 0; k--) d[k] = (v[k] << 1) | (v[k - 1] >> 31);
            d[0] = v[0] << 1;
            wide_add2n(v, d);
        }
        if (c == 2 || c == 6 || c == 4) {
#pragma unroll
            for (int k = 2 * N - 1; k > 0; k--) v[k] = (v[k] << 1) | (v[k - 1] >> 31);
            v[0] <<= 1;
        }
        if (c == 4) {
#pragma unroll
            for (int k = 2 * N - 1; k > 0; k--) v[k] = (v[k] << 1) | (v[k - 1] >> 31);
            v[0] <<= 1;
        }
    }
    // x (N words, plain integer) *= c, c in 1..4, no reduction
    static B200_HD void small_multiple_nr(E1& x, uint32_t c) {
        if (c == 3) {
            E1 d = x;
            shl1(d);
            add_nr(x, x, d);
        } else {
            if (c >= 2) shl1(x);
            if (c == 4) shl1(x);
        }
    }
    // r += mag * x for mag in 0..7 as ONE instruction sequence for every lane (the lanes of a phase carry different
    // coefficients; a branch per coefficient would serialise them): r += (x & m0) + (2x & m1) + (4x & m2), plain integers
    static B200_HD void add_small_multiple_uniform(E1& r, const E1& x, uint32_t mag) {
        E1 v = x;
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const uint32_t mask = 0u - ((mag >> b) & 1u);
            r.l[0] = add_cc(r.l[0], v.l[0] & mask);
#pragma unroll
            for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], v.l[i] & mask);
            r.l[N - 1] = addc(r.l[N - 1], v.l[N - 1] & mask);
            if (b < 2) shl1(v);
        }
    }
    // v[N..2N) += x: adds x R to the value under reduction, i.e. x to the reduced result
    static B200_HD void wide_add_high(uint32_t* v, const E1& x) {
        v[N] = add_cc(v[N], x.l[0]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) v[N + k] = addc_cc(v[N + k], x.l[k]);
        v[2 * N - 1] = addc(v[2 * N - 1], x.l[N - 1]);
    }
    // slot address without branches: the three run-time bases travel packed in one word (byte k = base of class k, byte 0 = 0)
#if defined(__CUDA_ARCH__)
    typedef uint32_t kboff_t;        // shared-memory word offset of the constant bank from the group's slot file
#else
    typedef long long kboff_t;       // host emulation: two unrelated allocations
#endif
    static B200_HD const uint32_t* slot_ptr(const uint32_t* slots, kboff_t kb_off, uint32_t bases, uint32_t o) {
        const uint32_t cls = (o >> 8) & 7, idx = o & 255;
        const uint32_t b = (bases >> (8 * (cls & 3))) & 255;
        return slots + (b + idx) * SLOT_WORDS + (cls == VM_C_CONST ? kb_off : (kboff_t)0);
    }

    static B200_HD_NOINLINE void exec_op(uint32_t* slots, const uint32_t* kbank, uint32_t b1, uint32_t b2, uint32_t b3,
                                         uint32_t live, const uint32_t* w) {
        const uint32_t bases = (b1 << 8) | (b2 << 16) | (b3 << 24);
        const kboff_t kb_off = (kboff_t)(kbank - slots);       // the constant bank lies behind the slot files
        const uint32_t hdr = w[0];
        const uint32_t kind = hdr & 15;
        if (kind == VM_NOP) return;
        const uint32_t nt = (hdr >> 4) & 15, nl = (hdr >> 8) & 15;
        const uint32_t scale = (hdr >> 12) & 7, halve = (hdr >> 15) & 1, pred = (hdr >> 16) & 3;
        uint32_t* dst = const_cast<uint32_t*>(slot_ptr(slots, kb_off, bases, w[1] & 0x7FF));
        if (pred && !((live >> (pred - 1)) & 1)) {
            E2 t;
            load2(t, slot_ptr(slots, kb_off, bases, (w[1] >> 16) & 0x7FF));
            store2(dst, t);
            return;
        }
        E2 res;
        if (kind == VM_INV) {
            E2 a;
            load2(a, slot_ptr(slots, kb_off, bases, w[2] & 0x7FF));
            T::f2_inv(res, a);
            store2(dst, res);
            return;
        }
        bool folded = false;
        if (kind == VM_DOT) {
            // Karatsuba over Fp2 with lazy reduction: per term v0 = a0 b0, v1 = a1 b1, v2 = (a0+a1)(b0+b1) (unreduced); every
            // product is accumulated IN PLACE into its own wide accumulator (T0 = sum v0, T1 = sum v1, T2 = sum v2; wide_mac)
            // and the combination
            //   RE = off p^2 + T0 - |BETA| T1        IM = T2 - T0 - T1
            // happens once per op, followed by ONE Montgomery reduction each.  off p^2 (header bits 24..28: the compiler's
            // bound on |BETA| sum a1 b1) keeps RE non-negative.
            // The SM sub-partition dispatches one IMAD.WIDE per 4 cycles and one other integer instruction per ~2 cycles and
            // does not overlap the two (measured: run time = 4 x IMAD.WIDE + 1.9 x others, summed over the resident warps,
            // from two warps per sub-partition on), so every instruction removed here is run time.
            uint32_t T0[2 * N], T1[2 * N], T2[2 * N];
#pragma unroll
            for (int k = 0; k < 2 * N; k++) { T0[k] = 0; T1[k] = 0; T2[k] = 0; }
            for (uint32_t t = 0; t < nt; t++) {
                const uint32_t tw = w[2 + t];
                const uint32_t* pa = slot_ptr(slots, kb_off, bases, tw & 0x7FF);
                const uint32_t* pb = slot_ptr(slots, kb_off, bases, (tw >> 11) & 0x7FF);
                const uint32_t am = (tw >> 22) & 15, bm = (tw >> 26) & 15;
                E1 a0, a1, b0;
                if (C::LAZY_MODS) {
                    load1(a0, pa);
                    load1(a1, pa + N);
                    if (am) lazy_mods(a0, a1, am);
                } else {
                    E2 a;
                    load2(a, pa);
                    if (am) apply_mod(a, am);
                    a0 = a.c0; a1 = a.c1;
                }
                const bool real_b = (bm & (VM_REAL0 | VM_REAL1)) != 0;
                load1(b0, pb + ((bm & VM_REAL1) ? N : 0));
                if (!real_b && (bm & VM_NEG)) p_minus_c(b0);
                wide_mac(T0, a0.l, b0.l);                       // v0 = a0 b0
                if (!real_b) {
                    E1 b1;
                    load1(b1, pb + N);
                    if (((bm & VM_NEG) != 0) != ((bm & VM_CONJ) != 0)) p_minus_c(b1);
                    wide_mac(T1, a1.l, b1.l);                   // v1 = a1 b1
                    add_nr(b0, b0, b1);                         // plain sums: below 2^(32N) by the compiler's bounds
                }
                add_nr(a0, a0, a1);
                wide_mac(T2, a0.l, b0.l);                       // v2 = (a0 + a1)(b0 + b1), or (a0 + a1) s for a real scalar s
            }
            {
                const uint32_t* off = C::K().p2 + ((hdr >> 24) & 31) * (2 * N);
                // T0 := RE = off + T0 - |BETA| T1,  T2 := IM = T2 - (T0 + T1)
                wide_sub2n(T2, T0);
                wide_sub2n(T2, T1);
                {
                    uint32_t o[2 * N];
#pragma unroll
                    for (int k = 0; k < 2 * N; k++) o[k] = off[k];
                    wide_add2n(T0, o);
                }
                wide_sub2n(T0, T1);
                if (C::BETA == -5) { wide_sub2n(T0, T1); wide_sub2n(T0, T1); wide_sub2n(T0, T1); wide_sub2n(T0, T1); }
                folded = ((hdr >> 29) & 1) != 0;
                if (folded) {
                    // scale and linear tail BEFORE the reduction (compiler.fold_bounds): c redc(V) = redc(c V), and adding
                    // x R to V adds x to redc(V) -- small multiples of unreduced operands into the high halves of RE / IM
                    // instead of modular doublings, additions and subtractions on the results
                    if (scale != 1) { wide_scale(T0, scale); wide_scale(T2, scale); }
                    for (uint32_t t = 0; t < nl; t++) {
                        const uint32_t lw = w[8 + t];
                        const uint32_t* px = slot_ptr(slots, kb_off, bases, lw & 0x7FF);
                        int coef = (int)((lw >> 11) & 31);
                        if (coef >= 16) coef -= 32;
                        uint32_t m = (lw >> 16) & 15;
                        if (coef < 0) m ^= VM_NEG;
                        const uint32_t mag = (uint32_t)(coef < 0 ? -coef : coef);
                        E1 x0, x1;
                        if (C::LAZY_MODS) {
                            load1(x0, px);
                            load1(x1, px + N);
                            if (m) lazy_mods(x0, x1, m);
                        } else {
                            E2 x;
                            load2(x, px);
                            if (m & ~VM_NEG) apply_mod(x, m & ~VM_NEG);
                            x0 = x.c0; x1 = x.c1;
                            if (m & VM_NEG) { p_minus_c(x0); p_minus_c(x1); }
                        }
                        if (mag != 1) { small_multiple_nr(x0, mag); small_multiple_nr(x1, mag); }
                        wide_add_high(T0, x0);
                        wide_add_high(T2, x1);
                    }
                }
                const uint32_t levels = (hdr >> 18) & 3;
                redc(res.c0, T0, levels);
                redc(res.c1, T2, levels);
            }
            if (!folded && scale != 1) { mul_small(res.c0, res.c0, scale); mul_small(res.c1, res.c1, scale); }
        } else {
            // LIN: sum c_i mod(x_i) as plain integers -- a negative coefficient becomes |c| (p - x), small multiples are shifts
            // and masked additions, ONE instruction sequence whatever the coefficients (the six lanes of a linear phase carry
            // different ones: with a modular doubling / addition / subtraction per coefficient the warp ran every variant in
            // turn, ~2,000 instructions per phase) -- then `levels` conditional subtractions (4p, 2p, p): the compiler's
            // bound  sum |c_i| bound(x_i) <= 2^levels  (compiler.lin_levels).
            E1 r0, r1;
            F::zero(r0);
            F::zero(r1);
            for (uint32_t t = 0; t < nl; t++) {
                const uint32_t lw = w[8 + t];
                const uint32_t* px = slot_ptr(slots, kb_off, bases, lw & 0x7FF);
                int coef = (int)((lw >> 11) & 31);
                if (coef >= 16) coef -= 32;
                uint32_t m = (lw >> 16) & 15;
                if (coef < 0) m ^= VM_NEG;
                const uint32_t mag = (uint32_t)(coef < 0 ? -coef : coef);
                E1 x0, x1;
                if (C::LAZY_MODS) {
                    load1(x0, px);
                    load1(x1, px + N);
                    if (m) lazy_mods(x0, x1, m);
                } else {
                    E2 x;
                    load2(x, px);
                    if (m & ~VM_NEG) apply_mod(x, m & ~VM_NEG);
                    x0 = x.c0; x1 = x.c1;
                    if (m & VM_NEG) { p_minus_c(x0); p_minus_c(x1); }
                }
                add_small_multiple_uniform(r0, x0, mag);
                add_small_multiple_uniform(r1, x1, mag);
            }
            const uint32_t levels = (hdr >> 18) & 3;
            if (levels >= 3) { cond_sub_kp(r0, 2); cond_sub_kp(r1, 2); }
            if (levels >= 2) { cond_sub_kp(r0, 1); cond_sub_kp(r1, 1); }
            cond_sub_kp(r0, 0);
            cond_sub_kp(r1, 0);
            res.c0 = r0; res.c1 = r1;
            folded = true;                          // the linear terms are consumed
        }
        if (!folded) {
            for (uint32_t t = 0; t < nl; t++) {
                const uint32_t lw = w[8 + t];
                E2 x;
                load2(x, slot_ptr(slots, kb_off, bases, lw & 0x7FF));
                apply_mod(x, (lw >> 16) & 15);
                int coef = (int)((lw >> 11) & 31);
                if (coef >= 16) coef -= 32;
                const uint32_t mag = (uint32_t)(coef < 0 ? -coef : coef);
                if (mag != 1) { mul_small(x.c0, x.c0, mag); mul_small(x.c1, x.c1, mag); }
                if (coef < 0) T::f2_sub(res, res, x); else T::f2_add(res, res, x);
            }
        }
        if (halve) T::f2_halve(res, res);
        store2(dst, res);
    }

    // ---------------------------------------------------------------------------------------------------
    // split mode (small batches, vm_pairing_split_kernel): THREE lanes per role.  Lane `sub` of a role accumulates ONE of the
    // three Karatsuba products of every term (sub 0: a0 b0, 1: a1 b1, 2: (a0+a1)(b0+b1)); the wide sums are exchanged
    // through a shared scratch area, then sub 0 reduces the real part and sub 1 the imaginary part, and each finishes its
    // own component (scale, linear tail, halving, store).  Same values as exec_op, a third of the multiplier chain per lane.
    // stage1, a warp barrier, stage2 -- the host emulation runs the three lanes of a stage one after another.
    // xch: this role's scratch, 3 * W words.
    // ---------------------------------------------------------------------------------------------------
    static B200_HD_NOINLINE void split_stage1(uint32_t* slots, const uint32_t* kbank, uint32_t b1, uint32_t b2, uint32_t b3,
                                              uint32_t live, const uint32_t* w, int sub, uint32_t* xch) {
        const uint32_t bases = (b1 << 8) | (b2 << 16) | (b3 << 24);
        const kboff_t kb_off = (kboff_t)(kbank - slots);
        const uint32_t hdr = w[0];
        if ((hdr & 15) != VM_DOT) return;
        const uint32_t nt = (hdr >> 4) & 15, pred = (hdr >> 16) & 3;
        if (pred && !((live >> (pred - 1)) & 1)) return;
        uint32_t Tacc[2 * N];
#pragma unroll
        for (int k = 0; k < 2 * N; k++) Tacc[k] = 0;
        for (uint32_t t = 0; t < nt; t++) {
            const uint32_t tw = w[2 + t];
            const uint32_t* pa = slot_ptr(slots, kb_off, bases, tw & 0x7FF);
            const uint32_t* pb = slot_ptr(slots, kb_off, bases, (tw >> 11) & 0x7FF);
            const uint32_t am = (tw >> 22) & 15, bm = (tw >> 26) & 15;
            E1 a0, a1, b0, b1x, sa, sb;
            if (C::LAZY_MODS) {
                load1(a0, pa);
                load1(a1, pa + N);
                if (am) lazy_mods(a0, a1, am);
            } else {
                E2 a;
                load2(a, pa);
                if (am) apply_mod(a, am);
                a0 = a.c0; a1 = a.c1;
            }
            if (bm & (VM_REAL0 | VM_REAL1)) {               // real scalar s = (s, 0)
                load1(b0, pb + ((bm & VM_REAL1) ? N : 0));
                F::zero(b1x);
            } else {
                load1(b0, pb);
                load1(b1x, pb + N);
                if (bm & VM_NEG) p_minus_c(b0);
                if (((bm & VM_NEG) != 0) != ((bm & VM_CONJ) != 0)) p_minus_c(b1x);
            }
            // this lane's factor pair: (a0, b0), (a1, b1) or the plain sums (a0 + a1, b0 + b1)
            add_nr(sa, a0, a1);
            add_nr(sb, b0, b1x);
            uint32_t xa[N], xb[N];
#pragma unroll
            for (int i = 0; i < N; i++) {
                xa[i] = sub == 0 ? a0.l[i] : (sub == 1 ? a1.l[i] : sa.l[i]);
                xb[i] = sub == 0 ? b0.l[i] : (sub == 1 ? b1x.l[i] : sb.l[i]);
            }
            wide_mac(Tacc, xa, xb);
        }
        uint32_t* mine = xch + sub * W;
#pragma unroll
        for (int k = 0; k < 2 * N; k++) mine[k] = Tacc[k];
    }
    static B200_HD_NOINLINE void split_stage2(uint32_t* slots, const uint32_t* kbank, uint32_t b1, uint32_t b2, uint32_t b3,
                                              uint32_t live, const uint32_t* w, int sub, const uint32_t* xch) {
        const uint32_t bases = (b1 << 8) | (b2 << 16) | (b3 << 24);
        const kboff_t kb_off = (kboff_t)(kbank - slots);
        const uint32_t hdr = w[0];
        const uint32_t kind = hdr & 15;
        if (kind == VM_NOP || sub == 2) return;
        const uint32_t nl = (hdr >> 8) & 15;
        const uint32_t scale = (hdr >> 12) & 7, halve = (hdr >> 15) & 1, pred = (hdr >> 16) & 3;
        uint32_t* dst = const_cast<uint32_t*>(slot_ptr(slots, kb_off, bases, w[1] & 0x7FF)) + sub * N;   // this lane's component
        if (pred && !((live >> (pred - 1)) & 1)) {
            const uint32_t* src = slot_ptr(slots, kb_off, bases, (w[1] >> 16) & 0x7FF) + sub * N;
            for (int i = 0; i < N; i++) dst[i] = src[i];
            return;
        }
        if (kind == VM_INV) {
            if (sub == 0) {
                E2 a, res;
                load2(a, slot_ptr(slots, kb_off, bases, w[2] & 0x7FF));
                T::f2_inv(res, a);
                store2(dst, res);
            }
            return;
        }
        E1 r;
        bool folded = false;
        if (kind == VM_DOT) {
            uint32_t V[2 * N], U[2 * N];
            const uint32_t levels = (hdr >> 18) & 3;
            if (sub == 0) {                                  // RE = off p^2 + T0 - |BETA| T1
                const uint32_t* off = C::K().p2 + ((hdr >> 24) & 31) * (2 * N);
#pragma unroll
                for (int k = 0; k < 2 * N; k++) { V[k] = off[k]; U[k] = xch[k]; }
                wide_add2n(V, U);
#pragma unroll
                for (int k = 0; k < 2 * N; k++) U[k] = xch[W + k];
                wide_sub2n(V, U);
                if (C::BETA == -5) { wide_sub2n(V, U); wide_sub2n(V, U); wide_sub2n(V, U); wide_sub2n(V, U); }
            } else {                                         // IM = T2 - T0 - T1
#pragma unroll
                for (int k = 0; k < 2 * N; k++) { V[k] = xch[2 * W + k]; U[k] = xch[k]; }
                wide_sub2n(V, U);
#pragma unroll
                for (int k = 0; k < 2 * N; k++) U[k] = xch[W + k];
                wide_sub2n(V, U);
            }
            folded = ((hdr >> 29) & 1) != 0;
            if (folded) {                                    // scale and linear tail before the reduction, as in exec_op
                if (scale != 1) wide_scale(V, scale);
                for (uint32_t t = 0; t < nl; t++) {
                    const uint32_t lw = w[8 + t];
                    const uint32_t* px = slot_ptr(slots, kb_off, bases, lw & 0x7FF);
                    int coef = (int)((lw >> 11) & 31);
                    if (coef >= 16) coef -= 32;
                    uint32_t m = (lw >> 16) & 15;
                    if (coef < 0) m ^= VM_NEG;
                    const uint32_t mag = (uint32_t)(coef < 0 ? -coef : coef);
                    E1 x0, x1;
                    if (C::LAZY_MODS) {
                        load1(x0, px);
                        load1(x1, px + N);
                        if (m) lazy_mods(x0, x1, m);
                    } else {
                        E2 x;
                        load2(x, px);
                        if (m & ~VM_NEG) apply_mod(x, m & ~VM_NEG);
                        x0 = x.c0; x1 = x.c1;
                        if (m & VM_NEG) { p_minus_c(x0); p_minus_c(x1); }
                    }
                    E1 xc = sub == 0 ? x0 : x1;
                    if (mag != 1) small_multiple_nr(xc, mag);
                    wide_add_high(V, xc);
                }
            }
            redc(r, V, levels);
            if (!folded && scale != 1) mul_small(r, r, scale);
        } else {
            // LIN as a plain integer sum with one instruction sequence for every coefficient, as in exec_op
            F::zero(r);
            for (uint32_t t = 0; t < nl; t++) {
                const uint32_t lw = w[8 + t];
                const uint32_t* px = slot_ptr(slots, kb_off, bases, lw & 0x7FF);
                int coef = (int)((lw >> 11) & 31);
                if (coef >= 16) coef -= 32;
                uint32_t m = (lw >> 16) & 15;
                if (coef < 0) m ^= VM_NEG;
                const uint32_t mag = (uint32_t)(coef < 0 ? -coef : coef);
                E1 x0, x1;
                if (C::LAZY_MODS) {
                    load1(x0, px);
                    load1(x1, px + N);
                    if (m) lazy_mods(x0, x1, m);
                } else {
                    E2 x;
                    load2(x, px);
                    if (m & ~VM_NEG) apply_mod(x, m & ~VM_NEG);
                    x0 = x.c0; x1 = x.c1;
                    if (m & VM_NEG) { p_minus_c(x0); p_minus_c(x1); }
                }
                add_small_multiple_uniform(r, sub == 0 ? x0 : x1, mag);
            }
            const uint32_t levels = (hdr >> 18) & 3;
            if (levels >= 3) cond_sub_kp(r, 2);
            if (levels >= 2) cond_sub_kp(r, 1);
            cond_sub_kp(r, 0);
            folded = true;
        }
        if (!folded) {
            for (uint32_t t = 0; t < nl; t++) {
                const uint32_t lw = w[8 + t];
                E2 x;
                load2(x, slot_ptr(slots, kb_off, bases, lw & 0x7FF));
                apply_mod(x, (lw >> 16) & 15);
                E1 xc = sub == 0 ? x.c0 : x.c1;
                int coef = (int)((lw >> 11) & 31);
                if (coef >= 16) coef -= 32;
                const uint32_t mag = (uint32_t)(coef < 0 ? -coef : coef);
                if (mag != 1) mul_small(xc, xc, mag);
                if (coef < 0) F::sub(r, r, xc); else F::add(r, r, xc);
            }
        }
        if (halve) F::halve(r, r);
        for (int i = 0; i < N; i++) dst[i] = r.l[i];
    }
};

}  // namespace b200
