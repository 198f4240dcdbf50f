// Optimal-ate Miller loop and final exponentiation, one pairing product per thread.
//
// Replaces what the reference adapters forward to gnark-crypto / kilic:
//   Pairing / Pairing2  reference driver/gurvy/bn254.go:247-263, bls12-377.go:244-260,
//                       bls12381/bls12-381.go:448-464, driver/kilic/bls12-381.go:260-277
//   FExp                reference bn254.go:265-267, bls12-377.go:262-264, bls12-381.go:466-468
// The G2 accumulator uses the homogeneous projective doubling / mixed-addition steps and sparse
// line slots of SURVEY A.5, so the raw (pre-FExp) value is the one the Gurvy drivers expose.
// Final exponent: 3(p^12-1)/r for BLS12 (Hayashida-Hayasaka-Teruya chain),
// 2x(6x^2+3x+1)(p^12-1)/r for BN254 (Fuentes-Castaneda chain) -- SURVEY A.2/A.4.
#pragma once
#include "tower.cuh"

namespace b200 {

template <int N> struct G1Aff { Fp<N> x, y; };           // (0,0) = infinity (gnark convention)
template <int N> struct G2Aff { Fp2<N> x, y; };
template <int N> struct G2Proj { Fp2<N> x, y, z; };

template <class C>
struct PairingOps {
    static constexpr int N = C::N;
    typedef Tower<C> T;
    typedef FpOps<C> F;
    typedef Fp<N> E1;
    typedef Fp2<N> E2;
    typedef Fp12<N> E12;

    struct Line { E2 r0, r1, r2; };

    static B200_HD bool g1_is_inf(const G1Aff<N>& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
    static B200_HD bool g2_is_inf(const G2Aff<N>& q) { return T::f2_is_zero(q.x) && T::f2_is_zero(q.y); }

    static B200_HD const E2& btw() { return *reinterpret_cast<const E2*>(C::K().btw); }

    // T <- 2T, tangent line coefficients
    static B200_HD_NOINLINE void double_step(G2Proj<N>& t, Line& l) {
        E2 A, B, Cc, D, E, Fv, G, H, I, J, K, x;
        T::f2_mul(A, t.x, t.y);
        T::f2_halve(A, A);
        T::f2_sqr(B, t.y);
        T::f2_sqr(Cc, t.z);
        T::f2_triple(D, Cc);
        E2 bt = btw();
        T::f2_mul(E, D, bt);
        T::f2_triple(Fv, E);
        T::f2_add(G, B, Fv);
        T::f2_halve(G, G);
        T::f2_add(H, t.y, t.z);
        T::f2_sqr(H, H);
        T::f2_add(x, B, Cc);
        T::f2_sub(H, H, x);
        T::f2_sub(I, E, B);
        T::f2_sqr(J, t.x);
        T::f2_sqr(K, E);
        T::f2_triple(K, K);
        T::f2_sub(x, B, Fv);
        T::f2_mul(t.x, x, A);
        T::f2_sqr(G, G);
        T::f2_sub(t.y, G, K);
        T::f2_mul(t.z, B, H);
        if (C::TWIST == TWIST_M) {
            l.r0 = I;
            T::f2_triple(l.r1, J);
            T::f2_neg(l.r2, H);
        } else {
            T::f2_neg(l.r0, H);
            T::f2_triple(l.r1, J);
            l.r2 = I;
        }
    }
    // T <- T + Q (Q affine), chord line coefficients; update=false computes the line only
    static B200_HD_NOINLINE void add_step(G2Proj<N>& t, Line& l, const G2Aff<N>& q, bool update) {
        E2 O, L, J, x, y;
        T::f2_mul(x, q.y, t.z);
        T::f2_sub(O, t.y, x);
        T::f2_mul(x, q.x, t.z);
        T::f2_sub(L, t.x, x);
        T::f2_mul(x, q.x, O);
        T::f2_mul(y, L, q.y);
        T::f2_sub(J, x, y);
        if (update) {
            E2 Cc, D, E, Fv, G, H;
            T::f2_sqr(Cc, O);
            T::f2_sqr(D, L);
            T::f2_mul(E, L, D);
            T::f2_mul(Fv, t.z, Cc);
            T::f2_mul(G, t.x, D);
            T::f2_add(H, E, Fv);
            T::f2_dbl(x, G);
            T::f2_sub(H, H, x);
            T::f2_mul(t.x, L, H);
            T::f2_sub(x, G, H);
            T::f2_mul(x, x, O);
            T::f2_mul(y, t.y, E);
            T::f2_sub(t.y, x, y);
            T::f2_mul(t.z, E, t.z);
        }
        if (C::TWIST == TWIST_M) {
            l.r0 = J;
            T::f2_neg(l.r1, O);
            l.r2 = L;
        } else {
            l.r0 = L;
            T::f2_neg(l.r1, O);
            l.r2 = J;
        }
    }
    // f <- f * line(P)
    static B200_HD void mul_line(E12& f, const Line& l, const G1Aff<N>& p) {
        E2 a, b;
        if (C::TWIST == TWIST_M) {          // C0.B0 = r0, C0.B1 = r1*xP, C1.B1 = r2*yP
            T::f2_mul_fp(a, l.r1, p.x);
            T::f2_mul_fp(b, l.r2, p.y);
            T::f12_mul_by_014(f, l.r0, a, b);
        } else {                            // C0.B0 = r0*yP, C1.B0 = r1*xP, C1.B1 = r2
            T::f2_mul_fp(a, l.r0, p.y);
            T::f2_mul_fp(b, l.r1, p.x);
            T::f12_mul_by_034(f, a, b, l.r2);
        }
    }

    // digit i of the loop scalar (LSB = 0): BLS12 bits of |x|; BN254 NAF of 6x+2
    static B200_HD int loop_len() { return C::FAMILY == FAMILY_BLS12 ? 64 : 66; }
    static B200_HD int loop_digit(int i) {
        if (C::FAMILY == FAMILY_BLS12) return (int)((C::X_ABS >> i) & 1);
        // NAF(6x+2), x = 4965661367192848881, LSB first
        const signed char naf[66] = {0, 0, 0, 1, 0, 1, 0, -1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, -1, 0, 0,
                                     0, 1, 0, -1, 0, 0, 0, 0, -1, 0, 0, 1, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0,
                                     -1, 0, 0, -1, 0, 1, 0, -1, 0, 0, 0, -1, 0, -1, 0, 0, 0, 1, 0, -1, 0, 1};
        return naf[i];
    }

    // prod_k f_{lambda,Q_k}(P_k) for NP (1 or 2) pairs; pairs with an infinity member are skipped
    template <int NP>
    static B200_HD void miller_loop(E12& f, const G1Aff<N>* P, const G2Aff<N>* Q) {
        G2Proj<N> t[NP];
        G2Aff<N> nq[NP];
        bool live[NP];
        T::f12_one(f);
        for (int k = 0; k < NP; k++) {
            live[k] = !(g1_is_inf(P[k]) || g2_is_inf(Q[k]));
            t[k].x = Q[k].x;
            t[k].y = Q[k].y;
            T::f2_one(t[k].z);
            nq[k].x = Q[k].x;
            T::f2_neg(nq[k].y, Q[k].y);
        }
        Line l;
        for (int i = loop_len() - 2; i >= 0; i--) {
            T::f12_sqr(f, f);
            int d = loop_digit(i);
            for (int k = 0; k < NP; k++) {
                if (!live[k]) continue;
                double_step(t[k], l);
                mul_line(f, l, P[k]);
                if (d != 0) {
                    add_step(t[k], l, d > 0 ? Q[k] : nq[k], true);
                    mul_line(f, l, P[k]);
                }
            }
        }
        if (C::FAMILY == FAMILY_BN) {
            for (int k = 0; k < NP; k++) {
                if (!live[k]) continue;
                G2Aff<N> q1, q2;
                // Q1 = pi(Q) = (conj(x) g_{1,2}, conj(y) g_{1,3});  -pi^2(Q) = (x g_{2,2}, y)
                T::f2_conj(q1.x, Q[k].x);
                T::f2_mul(q1.x, q1.x, T::frob_const(1, 2));
                T::f2_conj(q1.y, Q[k].y);
                T::f2_mul(q1.y, q1.y, T::frob_const(1, 3));
                T::f2_mul(q2.x, Q[k].x, T::frob_const(2, 2));
                q2.y = Q[k].y;
                add_step(t[k], l, q1, true);
                mul_line(f, l, P[k]);
                add_step(t[k], l, q2, false);
                mul_line(f, l, P[k]);
            }
        }
        if (C::X_NEG) T::f12_conj(f, f);
    }

    // z^|x| with cyclotomic squarings (conjugated when x < 0 so that it is z^x)
    static B200_HD_NOINLINE void exp_by_x(E12& r, const E12& z) {
        E12 acc = z;
        int top = 63;
        while (!((C::X_ABS >> top) & 1)) top--;
        for (int i = top - 1; i >= 0; i--) {
            T::f12_cyclo_sqr(acc, acc);
            if ((C::X_ABS >> i) & 1) T::f12_mul(acc, acc, z);
        }
        if (C::X_NEG) T::f12_conj(acc, acc);
        r = acc;
    }

    static B200_HD void final_exp(E12& r, const E12& in) {
        E12 f, t0, t1, t2;
        // easy part: f^((p^6-1)(p^2+1))
        T::f12_inv(t0, in);
        T::f12_conj(t1, in);
        T::f12_mul(t0, t1, t0);
        T::f12_frob(t1, t0, 2);
        T::f12_mul(f, t1, t0);
        if (C::FAMILY == FAMILY_BLS12) {
            T::f12_cyclo_sqr(t0, f);
            exp_by_x(t1, f);
            T::f12_conj(t2, f);
            T::f12_mul(t1, t1, t2);
            exp_by_x(t2, t1);
            T::f12_conj(t1, t1);
            T::f12_mul(t1, t1, t2);
            exp_by_x(t2, t1);
            T::f12_frob(t1, t1, 1);
            T::f12_mul(t1, t1, t2);
            T::f12_mul(f, f, t0);
            exp_by_x(t0, t1);
            exp_by_x(t2, t0);
            T::f12_frob(t0, t1, 2);
            T::f12_conj(t1, t1);
            T::f12_mul(t1, t1, t2);
            T::f12_mul(t1, t1, t0);
            T::f12_mul(f, f, t1);
            r = f;
        } else {
            E12 t3, t4;
            exp_by_x(t0, f); T::f12_conj(t0, t0);
            T::f12_cyclo_sqr(t0, t0);
            T::f12_cyclo_sqr(t1, t0);
            T::f12_mul(t1, t0, t1);
            exp_by_x(t2, t1); T::f12_conj(t2, t2);
            T::f12_conj(t3, t1);
            T::f12_mul(t1, t2, t3);
            T::f12_cyclo_sqr(t3, t2);
            exp_by_x(t4, t3);
            T::f12_mul(t4, t1, t4);
            T::f12_mul(t3, t0, t4);
            T::f12_mul(t0, t2, t4);
            T::f12_mul(t0, f, t0);
            T::f12_frob(t2, t3, 1);
            T::f12_mul(t0, t2, t0);
            T::f12_frob(t2, t4, 2);
            T::f12_mul(t0, t2, t0);
            T::f12_conj(t2, f);
            T::f12_mul(t2, t2, t3);
            T::f12_frob(t2, t2, 3);
            T::f12_mul(t0, t2, t0);
            r = t0;
        }
    }
};

}  // namespace b200
