// G1 arithmetic in extended-Jacobian (XYZZ) coordinates for curves y^2 = x^3 + b (a = 0).
//
// Replaces what the reference adapters forward to gnark-crypto / kilic:
//   G1.Mul            reference driver/gurvy/bn254.go:49-54, bls12-377.go:48-53,
//                     bls12381/bls12-381.go:238-247, driver/kilic/bls12-381.go:40-50
//   G1.Mul2 / InPlace reference bn254.go:56-70, bls12-381.go:249-280 + 869-937, kilic/bls12-381.go:52-66
//   MultiScalarMul    reference bn254.go:232-245, bls12-377.go:229-242, bls12-381.go:766-783
// Results are canonical group elements (affine, (0,0) = infinity as in gnark), so any correct
// addition chain is byte-exact.  The adders are complete: equal inputs fall through to
// the doubling, opposite inputs to infinity.
#pragma once
#include "curves.cuh"

namespace b200 {

template <int N> struct G1Affine { Fp<N> x, y; };                    // (0,0) = infinity
template <int N> struct G1XYZZ { Fp<N> x, y, zz, zzz; };             // zz == 0  <=> infinity

// B200_G1_FORMULAS_NOINLINE (diagnostic builds): compile the point formulas out of line
#if defined(B200_G1_FORMULAS_NOINLINE)
#define B200_G1_FN B200_HD_NOINLINE
#else
#define B200_G1_FN B200_HD
#endif

template <class C>
struct G1Ops {
    static constexpr int N = C::N;
    typedef FpOps<C> F;
    typedef Fp<N> E;
    typedef G1Affine<N> Aff;
    typedef G1XYZZ<N> Pt;

    static B200_HD bool aff_is_inf(const Aff& a) { return F::is_zero(a.x) && F::is_zero(a.y); }
    static B200_HD bool is_inf(const Pt& p) { return F::is_zero(p.zz); }
    static B200_HD void set_inf(Pt& p) { F::zero(p.x); F::zero(p.y); F::zero(p.zz); F::zero(p.zzz); }
    static B200_HD void from_affine(Pt& p, const Aff& a) {
        if (aff_is_inf(a)) { set_inf(p); return; }
        p.x = a.x; p.y = a.y; F::one(p.zz); F::one(p.zzz);
    }
    static B200_HD void neg_affine(Aff& r, const Aff& a) { r.x = a.x; F::neg(r.y, a.y); }
    static B200_HD void neg_y(Aff& a) { F::neg(a.y, a.y); }                    // group-generic MSM kernels (msm.cuh)
    static B200_HD void neg(Pt& r, const Pt& a) { r.x = a.x; F::neg(r.y, a.y); r.zz = a.zz; r.zzz = a.zzz; }

    // p <- 2a for affine a (mdbl-2008-s-1)
    static B200_G1_FN void dbl_affine(Pt& p, const Aff& a) {
        if (aff_is_inf(a) || F::is_zero(a.y)) { set_inf(p); return; }
        E U, V, W, S, M, t;
        F::dbl(U, a.y);
        F::sqrx(V, U);
        F::mulx(W, U, V);
        F::mulx(S, a.x, V);
        F::sqrx(M, a.x);
        F::dbl(t, M); F::add(M, M, t);
        F::sqrx(p.x, M);
        F::sub(p.x, p.x, S); F::sub(p.x, p.x, S);
        F::sub(t, S, p.x);
        F::neg(U, a.y);
        F::mulx2(p.y, M, t, W, U);     // M (S - X3) - W Y1: two products, one reduction
        p.zz = V;
        p.zzz = W;
    }
    // p <- 2p (dbl-2008-s-1)
    static B200_G1_FN void dbl(Pt& p) {
        if (is_inf(p)) return;
        E U, V, W, S, M, t;
        F::dbl(U, p.y);
        F::sqrx(V, U);
        F::mulx(W, U, V);
        F::mulx(S, p.x, V);
        F::sqrx(M, p.x);
        F::dbl(t, M); F::add(M, M, t);
        F::neg(U, p.y);
        F::sqrx(p.x, M);
        F::sub(p.x, p.x, S); F::sub(p.x, p.x, S);
        F::sub(t, S, p.x);
        F::mulx2(p.y, M, t, W, U);     // M (S - X3) - W Y1: two products, one reduction
        F::mulx(p.zz, V, p.zz);
        F::mulx(p.zzz, W, p.zzz);
    }
    // p <- p + a, a affine (madd-2008-s), complete
    static B200_G1_FN void madd(Pt& p, const Aff& a) {
        if (aff_is_inf(a)) return;
        if (is_inf(p)) { from_affine(p, a); return; }
        E U2, S2, Pp, R, PP, PPP, Q, t;
        F::mulx(U2, a.x, p.zz);
        F::mulx(S2, a.y, p.zzz);
        F::sub(Pp, U2, p.x);
        F::sub(R, S2, p.y);
        if (F::is_zero(Pp)) {
            if (F::is_zero(R)) dbl_affine(p, a); else set_inf(p);
            return;
        }
        F::sqrx(PP, Pp);
        F::mulx(PPP, Pp, PP);
        F::mulx(Q, p.x, PP);
        F::sqrx(t, R);
        F::sub(t, t, PPP); F::sub(t, t, Q); F::sub(t, t, Q);   // X3
        F::sub(Q, Q, t);
        // Y3 = R (Q - X3) - Y PPP = R (Q - X3) + (p - Y) PPP: two products, one reduction
        F::neg(S2, p.y);
        F::mulx2(p.y, R, Q, S2, PPP);
        p.x = t;
        F::mulx(p.zz, p.zz, PP);
        F::mulx(p.zzz, p.zzz, PPP);
    }
    // p <- p + q (add-2008-s), complete
    static B200_G1_FN void add(Pt& p, const Pt& q) {
        if (is_inf(q)) return;
        if (is_inf(p)) { p = q; return; }
        E U1, U2, S1, S2, Pp, R, PP, PPP, Q, t;
        F::mulx(U1, p.x, q.zz);
        F::mulx(U2, q.x, p.zz);
        F::mulx(S1, p.y, q.zzz);
        F::mulx(S2, q.y, p.zzz);
        F::sub(Pp, U2, U1);
        F::sub(R, S2, S1);
        if (F::is_zero(Pp)) {
            if (F::is_zero(R)) dbl(p); else set_inf(p);
            return;
        }
        F::sqrx(PP, Pp);
        F::mulx(PPP, Pp, PP);
        F::mulx(Q, U1, PP);
        F::sqrx(t, R);
        F::sub(t, t, PPP); F::sub(t, t, Q); F::sub(t, t, Q);   // X3
        F::sub(Q, Q, t);
        F::neg(S1, S1);
        F::mulx2(p.y, R, Q, S1, PPP);  // R (Q - X3) - S1 PPP: two products, one reduction
        p.x = t;
        F::mulx(p.zz, p.zz, q.zz);
        F::mulx(p.zz, p.zz, PP);
        F::mulx(p.zzz, p.zzz, q.zzz);
        F::mulx(p.zzz, p.zzz, PPP);
    }
    // affine (x,y) = (X/ZZ, Y/ZZZ); infinity -> (0,0)
    static B200_HD void to_affine(Aff& a, const Pt& p) {
        if (is_inf(p)) { F::zero(a.x); F::zero(a.y); return; }
        E t, i;
        F::mulx(t, p.zz, p.zzz);
        F::inv(i, t);
        F::mulx(t, i, p.zzz);      // 1/ZZ
        F::mulx(a.x, p.x, t);
        F::mulx(t, i, p.zz);       // 1/ZZZ
        F::mulx(a.y, p.y, t);
    }

    // scalar: 8 little-endian 32-bit words (any 256-bit value; [k]P = [k mod r]P in the r-torsion)
    static B200_HD int scalar_bit(const uint32_t* k, int i) { return (k[i >> 5] >> (i & 31)) & 1; }

    // ---- GLV (BLS12 curves): k = k1 + k2*lambda over the integers with lambda = x^2 - 1 (lambda^2 + lambda + 1 = 0 mod
    // r) and [lambda](X, Y) = (beta X, Y) on the order-r subgroup, so [k]P = [k1]P + [k2]phi(P) with ~128-bit halves: half
    // the doublings.  P + phi(P) = -phi^2(P) = (-(X + beta X), -Y) is free, so the joint ladder adds one of three AFFINE
    // points per step.  (gnark's ScalarMultiplication -- what the reference forwards to -- is a GLV ladder too; like it,
    // this assumes subgroup points, which is what mathlib's deserialisers guarantee.)
    // k2 = ((k >> 120) * floor(2^256/lambda)) >> 136 under-estimates floor(k/lambda) by at most 3, so k1 = k - k2*lambda
    // is in [0, 4*lambda): both halves fit 5 words for any 256-bit k.
    static B200_HD_NOINLINE void glv_split(uint32_t* k1, uint32_t* k2, const uint32_t* k) {
        const uint32_t* lam = C::K().glv_lambda;
        const uint32_t* M = C::K().glv_m;
        uint32_t kh[5], prod[11];
        for (int j = 0; j < 5; j++) kh[j] = (k[3 + j] >> 24) | (j + 4 < 8 ? (k[4 + j] << 8) : 0u);
        for (int j = 0; j < 11; j++) prod[j] = 0;
        for (int i = 0; i < 5; i++) {
            uint64_t carry = 0;
            for (int j = 0; j < 5; j++) {
                uint64_t t = (uint64_t)kh[i] * M[j] + prod[i + j] + carry;
                prod[i + j] = (uint32_t)t;
                carry = t >> 32;
            }
            prod[i + 5] = (uint32_t)carry;
        }
        for (int j = 0; j < 5; j++) k2[j] = (prod[4 + j] >> 8) | (prod[5 + j] << 24);
        uint32_t t5[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5; i++) {
            uint64_t carry = 0;
            for (int j = 0; j < 4 && i + j < 5; j++) {
                uint64_t t = (uint64_t)k2[i] * lam[j] + t5[i + j] + carry;
                t5[i + j] = (uint32_t)t;
                carry = t >> 32;
            }
            if (i == 0) t5[4] = (uint32_t)carry;          // rows i >= 1 end at or beyond word 4: their carry is 2^160 * ...
        }
        uint64_t borrow = 0;
        for (int j = 0; j < 5; j++) {
            uint64_t d = (uint64_t)k[j] - t5[j] - borrow;
            k1[j] = (uint32_t)d;
            borrow = (d >> 32) & 1;
        }
    }
    struct GlvTable { E x1, x2, x3, y, ny; };       // P = (x1, y), phi(P) = (x2, y), P + phi(P) = (x3, -y)
    static B200_HD_NOINLINE void glv_table(GlvTable& t, const Aff& base) {
        E beta;
        const uint32_t* bw = C::K().glv_beta;
        for (int i = 0; i < N; i++) beta.l[i] = bw[i];
        t.x1 = base.x;
        t.y = base.y;
        F::mulx(t.x2, base.x, beta);
        F::add(t.x3, t.x1, t.x2);
        F::neg(t.x3, t.x3);
        F::neg(t.ny, base.y);
    }
    static B200_HD void glv_step(Pt& acc, const GlvTable& t, uint32_t b) {    // b in 1..3
        Aff s;
        s.x = b == 1 ? t.x1 : (b == 2 ? t.x2 : t.x3);
        s.y = b == 3 ? t.ny : t.y;
        madd(acc, s);
    }
    static B200_HD uint32_t glv_bits(const uint32_t* k1, const uint32_t* k2, int i) {
        return ((k1[i >> 5] >> (i & 31)) & 1u) | (((k2[i >> 5] >> (i & 31)) & 1u) << 1);
    }

    static B200_HD void scalar_mul(Pt& acc, const Aff& base, const uint32_t* k) {
        set_inf(acc);
        if (C::FAMILY == FAMILY_BLS12) {
            uint32_t k1[5], k2[5];
            glv_split(k1, k2, k);
            GlvTable t;
            glv_table(t, base);
            int top = 159;
            while (top >= 0 && !glv_bits(k1, k2, top)) top--;
            for (int i = top; i >= 0; i--) {
                dbl(acc);
                const uint32_t b = glv_bits(k1, k2, i);
                if (b) glv_step(acc, t, b);
            }
            return;
        }
        for (int i = 255; i >= 0; i--) {
            dbl(acc);
            if (scalar_bit(k, i)) madd(acc, base);
        }
    }
    // [e]P + [f]Q.  BLS12: two GLV tables, one shared doubling chain of ~130 steps, at most two mixed additions per step.
    // BN254: Strauss-Shamir with a joint 1-bit window.
    static B200_HD void scalar_mul2(Pt& acc, const Aff& P, const uint32_t* e, const Aff& Q, const uint32_t* f) {
        set_inf(acc);
        if (C::FAMILY == FAMILY_BLS12) {
            uint32_t e1[5], e2[5], f1[5], f2[5];
            glv_split(e1, e2, e);
            glv_split(f1, f2, f);
            GlvTable tp, tq;
            glv_table(tp, P);
            glv_table(tq, Q);
            int top = 159;
            while (top >= 0 && !(glv_bits(e1, e2, top) | glv_bits(f1, f2, top))) top--;
            for (int i = top; i >= 0; i--) {
                dbl(acc);
                const uint32_t bp = glv_bits(e1, e2, i), bq = glv_bits(f1, f2, i);
                if (bp) glv_step(acc, tp, bp);
                if (bq) glv_step(acc, tq, bq);
            }
            return;
        }
        Pt pq;
        from_affine(pq, P);
        madd(pq, Q);
        for (int i = 255; i >= 0; i--) {
            dbl(acc);
            int b = scalar_bit(e, i) | (scalar_bit(f, i) << 1);
            if (b == 1) madd(acc, P);
            else if (b == 2) madd(acc, Q);
            else if (b == 3) add(acc, pq);
        }
    }
};

}  // namespace b200
