// Fp2 -> Fp6 -> Fp12 tower on top of fp.cuh (one element per thread).
//
//   Fp2  = Fp[u]/(u^2 - BETA)            BETA = -1 (BN254, BLS12-381), -5 (BLS12-377)
//   Fp6  = Fp2[v]/(v^3 - xi)             xi = 9+u, 1+u, u
//   Fp12 = Fp6[w]/(w^2 - v)
// Same element naming as the types on the reference path (gnark E2/E6/E12 embedded at reference
// driver/gurvy/bn254.go:183-185, kilic fe12 at driver/kilic/bls12-381.go:179-183): C0/C1 . B0/B1/B2 . A0/A1.
// All values are canonical Montgomery residues in [0,p), so raw outputs are bit-comparable.
#pragma once
#include "curves.cuh"

namespace b200 {

template <int N> struct Fp2 { Fp<N> c0, c1; };
template <int N> struct Fp6 { Fp2<N> c0, c1, c2; };
template <int N> struct Fp12 { Fp6<N> c0, c1; };

template <class C>
struct Tower {
    static constexpr int N = C::N;
    typedef FpOps<C> F;
    typedef Fp<N> E1;
    typedef Fp2<N> E2;
    typedef Fp6<N> E6;
    typedef Fp12<N> E12;

    // ------------------------------------------------------------------ Fp2
    static B200_HD void f2_zero(E2& r) { F::zero(r.c0); F::zero(r.c1); }
    static B200_HD void f2_one(E2& r) { F::one(r.c0); F::zero(r.c1); }
    static B200_HD bool f2_is_zero(const E2& a) { return F::is_zero(a.c0) && F::is_zero(a.c1); }
    static B200_HD bool f2_eq(const E2& a, const E2& b) { return F::eq(a.c0, b.c0) && F::eq(a.c1, b.c1); }
    static B200_HD void f2_add(E2& r, const E2& a, const E2& b) { F::add(r.c0, a.c0, b.c0); F::add(r.c1, a.c1, b.c1); }
    static B200_HD void f2_sub(E2& r, const E2& a, const E2& b) { F::sub(r.c0, a.c0, b.c0); F::sub(r.c1, a.c1, b.c1); }
    static B200_HD void f2_dbl(E2& r, const E2& a) { F::dbl(r.c0, a.c0); F::dbl(r.c1, a.c1); }
    static B200_HD void f2_neg(E2& r, const E2& a) { F::neg(r.c0, a.c0); F::neg(r.c1, a.c1); }
    static B200_HD void f2_conj(E2& r, const E2& a) { r.c0 = a.c0; F::neg(r.c1, a.c1); }
    static B200_HD void f2_halve(E2& r, const E2& a) { F::halve(r.c0, a.c0); F::halve(r.c1, a.c1); }
    static B200_HD void f2_triple(E2& r, const E2& a) { E2 t; f2_dbl(t, a); f2_add(r, t, a); }
    static B200_HD void f2_cmov(E2& r, const E2& a, bool c) { F::cmov(r.c0, a.c0, c); F::cmov(r.c1, a.c1, c); }

    // r = BETA * a  (a in Fp)
    static B200_HD void fp_mul_beta(E1& r, const E1& a) {
        if (C::BETA == -1) {
            F::neg(r, a);
        } else {  // -5
            E1 t;
            F::dbl(t, a); F::dbl(t, t); F::add(t, t, a);
            F::neg(r, t);
        }
    }

    // Fp2 product and square, out of line with operands and result BY VALUE (48 words in, 24 out in registers: references
    // would push every operand through a local-memory stack slot, as the round-1 G1 code did).  The product is two
    // two-operand Montgomery products (FpOps::mul_dot<2>: one reduction per component, 6N^2 + 2N multiply-accumulates and
    // no Karatsuba additions) instead of three full ones:  c1 = a0 b1 + a1 b0,  c0 = a0 b0 + (BETA a1) b1.
    static B200_HD_NOINLINE E2 f2_mulv(E2 a, E2 b) {
        E2 r;
        E1 n1;
        F::mul_dot2(r.c1, a.c0, b.c1, a.c1, b.c0);
        fp_mul_beta(n1, a.c1);
        F::mul_dot2(r.c0, a.c0, b.c0, n1, b.c1);
        return r;
    }
    static B200_HD void f2_mul(E2& r, const E2& a, const E2& b) { r = f2_mulv(a, b); }
    static B200_HD_NOINLINE E2 f2_sqrv(E2 a) {
        E2 r;
        E1 s, d, v;
        F::mul(v, a.c0, a.c1);
        F::add(s, a.c0, a.c1);
        if (C::BETA == -1) {
            F::sub(d, a.c0, a.c1);
            F::mul(r.c0, s, d);
        } else {  // (a0+a1)(a0-5a1) + 4v
            E1 t;
            fp_mul_beta(t, a.c1);
            F::add(d, a.c0, t);
            F::mul(t, s, d);
            F::dbl(d, v); F::dbl(d, d);
            F::add(r.c0, t, d);
        }
        F::dbl(r.c1, v);
        return r;
    }
    static B200_HD void f2_sqr(E2& r, const E2& a) { r = f2_sqrv(a); }
    // r = a * s, s in Fp
    static B200_HD void f2_mul_fp(E2& r, const E2& a, const E1& s) {
        F::mul(r.c0, a.c0, s);
        F::mul(r.c1, a.c1, s);
    }
    static B200_HD void f2_mul_xi(E2& r, const E2& a) {
        if (C::XI0 == 1 && C::XI1 == 1) {        // (1+u): BETA = -1
            E1 t;
            F::sub(t, a.c0, a.c1);
            F::add(r.c1, a.c0, a.c1);
            r.c0 = t;
        } else if (C::XI0 == 9) {                 // (9+u): BETA = -1 -> (9a0 - a1, 9a1 + a0)
            E1 t0, t1, n0, n1;
            F::dbl(t0, a.c0); F::dbl(t0, t0); F::dbl(t0, t0); F::add(t0, t0, a.c0);   // 9a0
            F::dbl(t1, a.c1); F::dbl(t1, t1); F::dbl(t1, t1); F::add(t1, t1, a.c1);   // 9a1
            F::sub(n0, t0, a.c1);
            F::add(n1, t1, a.c0);
            r.c0 = n0; r.c1 = n1;
        } else {                                  // u, u^2 = BETA
            E1 t;
            fp_mul_beta(t, a.c1);
            r.c1 = a.c0;
            r.c0 = t;
        }
    }
    static B200_HD void f2_inv(E2& r, const E2& a) {
        // 1/(a0 + a1 u) = (a0 - a1 u)/(a0^2 - BETA a1^2)
        E1 n0, n1;
        F::sqr(n0, a.c0);
        F::sqr(n1, a.c1);
        fp_mul_beta(n1, n1);
        F::sub(n0, n0, n1);
        F::inv(n0, n0);
        F::mul(r.c0, a.c0, n0);
        F::mul(n1, a.c1, n0);
        F::neg(r.c1, n1);
    }

    // ------------------------------------------------------------------ Fp6
    static B200_HD void f6_zero(E6& r) { f2_zero(r.c0); f2_zero(r.c1); f2_zero(r.c2); }
    static B200_HD void f6_one(E6& r) { f2_one(r.c0); f2_zero(r.c1); f2_zero(r.c2); }
    static B200_HD void f6_add(E6& r, const E6& a, const E6& b) { f2_add(r.c0, a.c0, b.c0); f2_add(r.c1, a.c1, b.c1); f2_add(r.c2, a.c2, b.c2); }
    static B200_HD void f6_sub(E6& r, const E6& a, const E6& b) { f2_sub(r.c0, a.c0, b.c0); f2_sub(r.c1, a.c1, b.c1); f2_sub(r.c2, a.c2, b.c2); }
    static B200_HD void f6_neg(E6& r, const E6& a) { f2_neg(r.c0, a.c0); f2_neg(r.c1, a.c1); f2_neg(r.c2, a.c2); }
    static B200_HD void f6_dbl(E6& r, const E6& a) { f2_dbl(r.c0, a.c0); f2_dbl(r.c1, a.c1); f2_dbl(r.c2, a.c2); }
    static B200_HD bool f6_eq(const E6& a, const E6& b) { return f2_eq(a.c0, b.c0) && f2_eq(a.c1, b.c1) && f2_eq(a.c2, b.c2); }
    // r = v * a
    static B200_HD void f6_mul_v(E6& r, const E6& a) {
        E2 t;
        f2_mul_xi(t, a.c2);
        r.c2 = a.c1;
        r.c1 = a.c0;
        r.c0 = t;
    }
    static B200_HD_NOINLINE void f6_mul(E6& r, const E6& a, const E6& b) {
        E2 t0, t1, t2, s, u, x;
        f2_mul(t0, a.c0, b.c0);
        f2_mul(t1, a.c1, b.c1);
        f2_mul(t2, a.c2, b.c2);
        E6 o;
        // c0 = ((a1+a2)(b1+b2) - t1 - t2) xi + t0
        f2_add(s, a.c1, a.c2); f2_add(u, b.c1, b.c2);
        f2_mul(x, s, u);
        f2_sub(x, x, t1); f2_sub(x, x, t2);
        f2_mul_xi(x, x);
        f2_add(o.c0, x, t0);
        // c1 = (a0+a1)(b0+b1) - t0 - t1 + xi t2
        f2_add(s, a.c0, a.c1); f2_add(u, b.c0, b.c1);
        f2_mul(x, s, u);
        f2_sub(x, x, t0); f2_sub(x, x, t1);
        f2_mul_xi(s, t2);
        f2_add(o.c1, x, s);
        // c2 = (a0+a2)(b0+b2) - t0 - t2 + t1
        f2_add(s, a.c0, a.c2); f2_add(u, b.c0, b.c2);
        f2_mul(x, s, u);
        f2_sub(x, x, t0); f2_sub(x, x, t2);
        f2_add(o.c2, x, t1);
        r = o;
    }
    // r = a * (c0 + c1 v)
    static B200_HD_NOINLINE void f6_mul_by_01(E6& r, const E6& a, const E2& c0, const E2& c1) {
        E2 t0, t1, t2, s, u, x;
        E6 o;
        f2_mul(t0, a.c0, c0);
        f2_mul(t1, a.c1, c1);
        f2_mul(t2, a.c2, c1);
        f2_mul_xi(t2, t2);
        f2_add(o.c0, t0, t2);
        f2_add(s, a.c0, a.c1); f2_add(u, c0, c1);
        f2_mul(x, s, u);
        f2_sub(x, x, t0);
        f2_sub(o.c1, x, t1);
        f2_mul(x, a.c2, c0);
        f2_add(o.c2, x, t1);
        r = o;
    }
    // r = a * (c1 v)
    static B200_HD void f6_mul_by_1(E6& r, const E6& a, const E2& c1) {
        E2 t0, t1, t2;
        f2_mul(t2, a.c2, c1);
        f2_mul(t0, a.c0, c1);
        f2_mul(t1, a.c1, c1);
        f2_mul_xi(r.c0, t2);
        r.c1 = t0;
        r.c2 = t1;
    }
    // r = a * c0 (c0 in Fp2)
    static B200_HD void f6_mul_by_0(E6& r, const E6& a, const E2& c0) {
        f2_mul(r.c0, a.c0, c0);
        f2_mul(r.c1, a.c1, c0);
        f2_mul(r.c2, a.c2, c0);
    }
    static B200_HD_NOINLINE void f6_inv(E6& r, const E6& a) {
        E2 t0, t1, t2, x, y, d;
        f2_sqr(t0, a.c0); f2_mul(x, a.c1, a.c2); f2_mul_xi(x, x); f2_sub(t0, t0, x);        // a0^2 - xi a1 a2
        f2_sqr(t1, a.c2); f2_mul_xi(t1, t1); f2_mul(x, a.c0, a.c1); f2_sub(t1, t1, x);      // xi a2^2 - a0 a1
        f2_sqr(t2, a.c1); f2_mul(x, a.c0, a.c2); f2_sub(t2, t2, x);                         // a1^2 - a0 a2
        f2_mul(x, a.c2, t1); f2_mul(y, a.c1, t2); f2_add(x, x, y); f2_mul_xi(x, x);
        f2_mul(d, a.c0, t0); f2_add(d, d, x);
        f2_inv(d, d);
        f2_mul(r.c0, t0, d); f2_mul(r.c1, t1, d); f2_mul(r.c2, t2, d);
    }

    // ------------------------------------------------------------------ Fp12
    static B200_HD void f12_one(E12& r) { f6_one(r.c0); f6_zero(r.c1); }
    static B200_HD bool f12_eq(const E12& a, const E12& b) { return f6_eq(a.c0, b.c0) && f6_eq(a.c1, b.c1); }
    static B200_HD bool f12_is_one(const E12& a) { E12 o; f12_one(o); return f12_eq(a, o); }
    static B200_HD void f12_conj(E12& r, const E12& a) { r.c0 = a.c0; f6_neg(r.c1, a.c1); }
    static B200_HD_NOINLINE void f12_mul(E12& r, const E12& a, const E12& b) {
        E6 t0, t1, s, u;
        f6_mul(t0, a.c0, b.c0);
        f6_mul(t1, a.c1, b.c1);
        f6_add(s, a.c0, a.c1);
        f6_add(u, b.c0, b.c1);
        f6_mul(s, s, u);
        f6_sub(s, s, t0);
        f6_sub(r.c1, s, t1);
        f6_mul_v(t1, t1);
        f6_add(r.c0, t0, t1);
    }
    static B200_HD_NOINLINE void f12_sqr(E12& r, const E12& a) {
        // complex method: t = a0 a1; c0 = (a0+a1)(a0 + v a1) - t - v t; c1 = 2t
        E6 t, s, u;
        f6_mul(t, a.c0, a.c1);
        f6_add(s, a.c0, a.c1);
        f6_mul_v(u, a.c1);
        f6_add(u, u, a.c0);
        f6_mul(s, s, u);
        f6_sub(s, s, t);
        f6_mul_v(u, t);
        f6_sub(r.c0, s, u);
        f6_dbl(r.c1, t);
    }
    static B200_HD_NOINLINE void f12_inv(E12& r, const E12& a) {
        E6 t0, t1;
        f6_mul(t0, a.c0, a.c0);
        f6_mul(t1, a.c1, a.c1);
        f6_mul_v(t1, t1);
        f6_sub(t0, t0, t1);
        f6_inv(t0, t0);
        f6_mul(r.c0, a.c0, t0);
        f6_mul(t1, a.c1, t0);
        f6_neg(r.c1, t1);
    }
    // sparse multiplications by a line
    //   M-type "014": line = c0 + c1 v + c4 v w      (C0.B0, C0.B1, C1.B1)
    static B200_HD_NOINLINE void f12_mul_by_014(E12& f, const E2& c0, const E2& c1, const E2& c4) {
        E6 a, b, e;
        E2 d;
        f6_mul_by_01(a, f.c0, c0, c1);
        f6_mul_by_1(b, f.c1, c4);
        f2_add(d, c1, c4);
        f6_add(e, f.c0, f.c1);
        f6_mul_by_01(e, e, c0, d);
        f6_sub(e, e, a);
        f6_sub(f.c1, e, b);
        f6_mul_v(b, b);
        f6_add(f.c0, a, b);
    }
    //   D-type "034": line = c0 + c3 w + c4 v w      (C0.B0, C1.B0, C1.B1)
    static B200_HD_NOINLINE void f12_mul_by_034(E12& f, const E2& c0, const E2& c3, const E2& c4) {
        E6 a, b, e;
        E2 d;
        f6_mul_by_0(a, f.c0, c0);
        f6_mul_by_01(b, f.c1, c3, c4);
        f2_add(d, c0, c3);
        f6_add(e, f.c0, f.c1);
        f6_mul_by_01(e, e, d, c4);
        f6_sub(e, e, a);
        f6_sub(f.c1, e, b);
        f6_mul_v(b, b);
        f6_add(f.c0, a, b);
    }

    // Frobenius a -> a^(p^k), k = 1,2,3.  w-basis coefficient i is (conjugated for odd k and)
    // multiplied by gamma_{k,i} = xi^(i(p^k-1)/6).
    static B200_HD const E2& frob_const(int k, int i) {
        const uint32_t* base = (k == 1) ? C::K().frob1 : (k == 2) ? C::K().frob2 : C::K().frob3;
        return *reinterpret_cast<const E2*>(base + (i - 1) * 2 * N);
    }
    static B200_HD_NOINLINE void f12_frob(E12& r, const E12& a, int k) {
        E2* ro[6] = {&r.c0.c0, &r.c1.c0, &r.c0.c1, &r.c1.c1, &r.c0.c2, &r.c1.c2};
        const E2* ai[6] = {&a.c0.c0, &a.c1.c0, &a.c0.c1, &a.c1.c1, &a.c0.c2, &a.c1.c2};
        for (int i = 0; i < 6; i++) {
            E2 t = *ai[i];
            if (k & 1) f2_conj(t, t);
            if (i > 0) {
                E2 g = frob_const(k, i);
                f2_mul(t, t, g);
            }
            *ro[i] = t;
        }
    }

    // Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part).
    static B200_HD void fp4_sqr(E2& r0, E2& r1, const E2& a, const E2& b) {
        // (a + b s)^2, s^2 = xi:  r0 = a^2 + xi b^2, r1 = 2ab
        E2 t0, t1, t2;
        f2_sqr(t0, a);
        f2_sqr(t1, b);
        f2_add(t2, a, b);
        f2_sqr(t2, t2);
        f2_sub(t2, t2, t0);
        f2_sub(r1, t2, t1);
        f2_mul_xi(t1, t1);
        f2_add(r0, t0, t1);
    }
    static B200_HD_NOINLINE void f12_cyclo_sqr(E12& r, const E12& a) {
        // w-basis g0..g5 = C0.B0, C1.B0, C0.B1, C1.B1, C0.B2, C1.B2 ; Fp4 pairs (g0,g3), (g1,g4), (g2,g5)
        E2 a0, a1, b0, b1, c0, c1, t;
        fp4_sqr(a0, a1, a.c0.c0, a.c1.c1);   // (g0,g3)^2
        fp4_sqr(b0, b1, a.c1.c0, a.c0.c2);   // (g1,g4)^2
        fp4_sqr(c0, c1, a.c0.c1, a.c1.c2);   // (g2,g5)^2
        E12 o;
        // g0' = 3 a0 - 2 g0 ; g3' = 3 a1 + 2 g3
        f2_sub(t, a0, a.c0.c0); f2_dbl(t, t); f2_add(o.c0.c0, t, a0);
        f2_add(t, a1, a.c1.c1); f2_dbl(t, t); f2_add(o.c1.c1, t, a1);
        // g1' = 3 xi c1 + 2 g1 ; g4' = 3 c0 - 2 g4
        f2_mul_xi(c1, c1);
        f2_add(t, c1, a.c1.c0); f2_dbl(t, t); f2_add(o.c1.c0, t, c1);
        f2_sub(t, c0, a.c0.c2); f2_dbl(t, t); f2_add(o.c0.c2, t, c0);
        // g2' = 3 b0 - 2 g2 ; g5' = 3 b1 + 2 g5
        f2_sub(t, b0, a.c0.c1); f2_dbl(t, t); f2_add(o.c0.c1, t, b0);
        f2_add(t, b1, a.c1.c2); f2_dbl(t, t); f2_add(o.c1.c2, t, b1);
        r = o;
    }
};

}  // namespace b200
