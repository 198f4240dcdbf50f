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

    static B200_HD const uint32_t* operand_ptr(const Ctx& c, uint32_t o) {
        uint32_t cls = (o >> 8) & 7, idx = o & 255;
        if (cls == VM_C_CONST) return c.kbank + idx * SLOT_WORDS;
        uint32_t b = cls == VM_C_ABS ? 0u : c.base[cls - 1];
        return c.slots + (b + idx) * SLOT_WORDS;
    }
    static B200_HD void load2(E2& r, const uint32_t* p) {
#pragma unroll
        for (int i = 0; i < N; i++) { r.c0.l[i] = p[i]; r.c1.l[i] = p[N + i]; }
    }
    static B200_HD void store2(uint32_t* p, const E2& r) {
#pragma unroll
        for (int i = 0; i < N; i++) { p[i] = r.c0.l[i]; p[N + i] = r.c1.l[i]; }
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
    // v[0..2N) = a * b (fresh), a, b < 2^(32N)
    static B200_HD void wide_mul(uint32_t* v, const uint32_t* a, const uint32_t* b) {
        uint32_t Ev[2 * N + 2], Od[2 * N + 2];
#pragma unroll
        for (int i = 0; i < 2 * N + 2; i++) { Ev[i] = 0; Od[i] = 0; }
#pragma unroll
        for (int i = 0; i < N; i++) {
            // products a[j]*b[i] land on word i+j: even (i+j) -> Ev, odd -> Od (Od[k] holds word k+1)
            const int je = i & 1;          // first j with (i+j) even
            const int jo = je ^ 1;         // first j with (i+j) odd
            {   // even-aligned chain, words i+je .. i+je+N-1
                const int w0 = i + je;
                Ev[w0] = mad_lo_cc(a[je], b[i], Ev[w0]);
                Ev[w0 + 1] = madc_hi_cc(a[je], b[i], Ev[w0 + 1]);
#pragma unroll
                for (int j = je + 2; j < N; j += 2) {
                    Ev[i + j] = madc_lo_cc(a[j], b[i], Ev[i + j]);
                    Ev[i + j + 1] = madc_hi_cc(a[j], b[i], Ev[i + j + 1]);
                }
                Ev[w0 + N] = addc(Ev[w0 + N], 0);
            }
            {   // odd-aligned chain: product at word i+jo (odd) is stored at Od[i+jo-1]
                const int w0 = i + jo - 1;
                Od[w0] = mad_lo_cc(a[jo], b[i], Od[w0]);
                Od[w0 + 1] = madc_hi_cc(a[jo], b[i], Od[w0 + 1]);
#pragma unroll
                for (int j = jo + 2; j < N; j += 2) {
                    Od[i + j - 1] = madc_lo_cc(a[j], b[i], Od[i + j - 1]);
                    Od[i + j] = madc_hi_cc(a[j], b[i], Od[i + j]);
                }
                Od[w0 + N] = addc(Od[w0 + N], 0);
            }
        }
        // v = Ev + (Od << 32)
        v[0] = Ev[0];
        v[1] = add_cc(Ev[1], Od[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; k++) v[k] = addc_cc(Ev[k], Od[k - 1]);
        v[2 * N - 1] = addc(Ev[2 * N - 1], Od[2 * N - 2]);
    }
    // acc[0..W) += v[0..2N)
    static B200_HD void wide_add(uint32_t* acc, const uint32_t* v) {
        acc[0] = add_cc(acc[0], v[0]);
#pragma unroll
        for (int k = 1; k < 2 * N; k++) acc[k] = addc_cc(acc[k], v[k]);
        acc[2 * N] = addc(acc[2 * N], 0);
    }
    static B200_HD void wide_sub(uint32_t* acc, const uint32_t* v) {
        acc[0] = sub_cc(acc[0], v[0]);
#pragma unroll
        for (int k = 1; k < 2 * N; k++) acc[k] = subc_cc(acc[k], v[k]);
        acc[2 * N] = subc(acc[2 * N], 0);
    }
    // Montgomery reduction of a wide value T < 4 p R  ->  canonical residue T / R mod p
    static B200_HD void redc(E1& r, const uint32_t* Tw) {
        const uint32_t* p = C::p();
        // REDC(T_low) via the reduction rows of FpOps::mul (no a*b part), then + T_high
        uint32_t X[N + 2], Y[N + 2];
#pragma unroll
        for (int j = 0; j < N; j += 2) { X[j] = Tw[j]; X[j + 1] = 0; Y[j] = Tw[j + 1]; Y[j + 1] = 0; }
        X[N] = 0; X[N + 1] = 0; Y[N] = 0; Y[N + 1] = 0;
        {
            uint32_t m = mul_lo(X[0], C::inv32());
            F::chain_odd(Y, p, m);
            F::chain_even(X, p, m);
            X[N] = addc(0, 0);
        }
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            redc_row(X, Y);
            if (i + 1 < N) redc_row(Y, X);
        }
        E1 lo;
        lo.l[0] = add_cc(X[0], Y[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) lo.l[i] = addc_cc(X[i], Y[i + 1]);
        lo.l[N - 1] = addc(X[N - 1], Y[N]);
        // + T_high (N words; T < 4pR so the sum is < 5p < 2^(32N))
        r.l[0] = add_cc(lo.l[0], Tw[N]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(lo.l[i], Tw[N + i]);
        r.l[N - 1] = addc(lo.l[N - 1], Tw[2 * N - 1]);
        // canonicalise: r < 4p -> subtract 2p, then p, conditionally
        cond_sub_kp(r, 1);
        cond_sub_kp(r, 0);
    }
    // row of the reduction without a multiplicand part: A = previous even accumulator, B = previous odd accumulator
    static B200_HD void redc_row(uint32_t* A, uint32_t* B) {
        const uint32_t* p = C::p();
        B[0] = add_cc(B[0], A[1]);
#pragma unroll
        for (int j = 0; j < N - 2; j++) A[j] = addc_cc(A[j + 2], 0);
        A[N - 2] = addc(A[N], 0);
        A[N - 1] = 0;
        B[N] = 0;
        uint32_t m = mul_lo(B[0], C::inv32());
        F::chain_odd(A, p, m);
        F::chain_even(B, p, m);
        B[N] = addc(B[N], 0);
    }
    // r -= (p << sh) if r >= (p << sh)
    static B200_HD void cond_sub_kp(E1& r, int sh) {
        const uint32_t* p = C::p();
        uint32_t t[N];
        uint32_t q0 = p[0] << sh;
        t[0] = sub_cc(r.l[0], q0);
#pragma unroll
        for (int i = 1; i < N; i++) {
            uint32_t qi = sh ? ((p[i] << sh) | (p[i - 1] >> (32 - sh))) : p[i];
            t[i] = subc_cc(r.l[i], qi);
        }
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
    static B200_HD_NOINLINE void exec_op(const Ctx& c, const uint32_t* w) {
        const uint32_t hdr = w[0];
        const uint32_t kind = hdr & 15;
        if (kind == VM_NOP) return;
        const uint32_t nt = (hdr >> 4) & 15, nl = (hdr >> 8) & 15;
        const uint32_t scale = (hdr >> 12) & 7, halve = (hdr >> 15) & 1, pred = (hdr >> 16) & 3;
        uint32_t* dst = const_cast<uint32_t*>(operand_ptr(c, w[1] & 0x7FF));
        if (pred && !((c.live >> (pred - 1)) & 1)) {
            E2 t;
            load2(t, operand_ptr(c, (w[1] >> 16) & 0x7FF));
            store2(dst, t);
            return;
        }
        E2 res;
        if (kind == VM_INV) {
            E2 a;
            load2(a, operand_ptr(c, w[2] & 0x7FF));
            T::f2_inv(res, a);
            store2(dst, res);
            return;
        }
        if (kind == VM_DOT) {
            uint32_t RE[W], IM[W];
            // RE starts at nt * |BETA| * p^2 so that the subtractions below never underflow
            {
                const uint32_t* off = C::K().p2 + (nt * (C::BETA == -1 ? 1 : 5)) * (2 * N);
#pragma unroll
                for (int k = 0; k < 2 * N; k++) { RE[k] = off[k]; IM[k] = 0; }
                RE[2 * N] = 0; IM[2 * N] = 0;
            }
            for (uint32_t t = 0; t < nt; t++) {
                const uint32_t tw = w[2 + t];
                E2 a, b;
                load2(a, operand_ptr(c, tw & 0x7FF));
                apply_mod(a, (tw >> 22) & 15);
                const uint32_t bm = (tw >> 26) & 15;
                load2(b, operand_ptr(c, (tw >> 11) & 0x7FF));
                const bool real_b = (bm & (VM_REAL0 | VM_REAL1)) != 0;
                if (real_b) {
                    if (bm & VM_REAL1) b.c0 = b.c1;
                } else {
                    apply_mod(b, bm & 3);
                }
                uint32_t v[2 * N];
                wide_mul(v, a.c0.l, b.c0.l);               // v0 = a0 b0
                wide_add(RE, v);
                if (!real_b) {
                    wide_sub(IM, v);
                    wide_mul(v, a.c1.l, b.c1.l);           // v1 = a1 b1
                    wide_sub(IM, v);
                    wide_sub(RE, v);
                    if (C::BETA == -5) { wide_sub(RE, v); wide_sub(RE, v); wide_sub(RE, v); wide_sub(RE, v); }
                    // (a0+a1)(b0+b1): sums stay below 2p < 2^(32N)
                    a.c0.l[0] = add_cc(a.c0.l[0], a.c1.l[0]);
#pragma unroll
                    for (int i = 1; i < N; i++) a.c0.l[i] = addc_cc(a.c0.l[i], a.c1.l[i]);
                    b.c0.l[0] = add_cc(b.c0.l[0], b.c1.l[0]);
#pragma unroll
                    for (int i = 1; i < N; i++) b.c0.l[i] = addc_cc(b.c0.l[i], b.c1.l[i]);
                    wide_mul(v, a.c0.l, b.c0.l);
                    wide_add(IM, v);
                } else {
                    wide_mul(v, a.c1.l, b.c0.l);           // imaginary part a1 * s
                    wide_add(IM, v);
                }
            }
            redc(res.c0, RE);
            redc(res.c1, IM);
            if (scale != 1) { mul_small(res.c0, res.c0, scale); mul_small(res.c1, res.c1, scale); }
        } else {
            T::f2_zero(res);
        }
        for (uint32_t t = 0; t < nl; t++) {
            const uint32_t lw = w[8 + t];
            E2 x;
            load2(x, operand_ptr(c, lw & 0x7FF));
            apply_mod(x, (lw >> 16) & 15);
            int coef = (int)((lw >> 11) & 31);
            if (coef >= 16) coef -= 32;
            const uint32_t mag = (uint32_t)(coef < 0 ? -coef : coef);
            if (mag != 1) { mul_small(x.c0, x.c0, mag); mul_small(x.c1, x.c1, mag); }
            if (coef < 0) T::f2_sub(res, res, x); else T::f2_add(res, res, x);
        }
        if (halve) T::f2_halve(res, res);
        store2(dst, res);
    }
};

}  // namespace b200
