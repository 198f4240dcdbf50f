// G2 arithmetic on the sextic twist E'(Fp2): y^2 = x^3 + b' (a = 0), extended-Jacobian (XYZZ) coordinates over Fp2.
//
// SURVEY 8(f) row 3 -- the callers next to the hot path:
//   driver.G2.Mul(Zr)   reference driver/math.go:307; impls bn254.go:134-139, bls12-377.go:131-136,
//                       bls12381/bls12-381.go:342-351, kilic/bls12-381.go:127-137
//   driver.G2.Add       reference driver/math.go:310; impls bn254.go:141, bls12-377.go:138, kilic/bls12-381.go:139
// Results are canonical group elements (affine, (0,0) = infinity), so any correct addition chain is byte-exact with
// the reference's `Bytes()`; the formulas are the a = 0 XYZZ ones of g1.cuh with Fp2 in place of Fp (they never use
// the curve constant, so the same code serves every twist).
#pragma once
#include "kernels.cuh"

namespace b200 {

template <int N> struct G2XYZZ { Fp2<N> x, y, zz, zzz; };            // zz == 0  <=> infinity

template <class C>
struct G2Ops {
    static constexpr int N = C::N;
    typedef Tower<C> T;
    typedef Fp2<N> E;
    typedef G2Aff<N> Aff;
    typedef G2XYZZ<N> Pt;

    static B200_HD bool aff_is_inf(const Aff& a) { return T::f2_is_zero(a.x) && T::f2_is_zero(a.y); }
    static B200_HD bool is_inf(const Pt& p) { return T::f2_is_zero(p.zz); }
    static B200_HD void set_inf(Pt& p) { T::f2_zero(p.x); T::f2_zero(p.y); T::f2_zero(p.zz); T::f2_zero(p.zzz); }
    static B200_HD void from_affine(Pt& p, const Aff& a) {
        if (aff_is_inf(a)) { set_inf(p); return; }
        p.x = a.x; p.y = a.y; T::f2_one(p.zz); T::f2_one(p.zzz);
    }
    // p <- 2a for affine a (mdbl-2008-s-1)
    static B200_HD_NOINLINE void dbl_affine(Pt& p, const Aff& a) {
        if (aff_is_inf(a) || T::f2_is_zero(a.y)) { set_inf(p); return; }
        E U, V, W, S, M, t;
        T::f2_dbl(U, a.y);
        T::f2_sqr(V, U);
        T::f2_mul(W, U, V);
        T::f2_mul(S, a.x, V);
        T::f2_sqr(M, a.x);
        T::f2_triple(M, M);
        T::f2_sqr(p.x, M);
        T::f2_sub(p.x, p.x, S); T::f2_sub(p.x, p.x, S);
        T::f2_sub(t, S, p.x);
        T::f2_mul(t, M, t);
        T::f2_mul(U, W, a.y);
        T::f2_sub(p.y, t, U);
        p.zz = V;
        p.zzz = W;
    }
    // p <- 2p (dbl-2008-s-1)
    static B200_HD_NOINLINE void dbl(Pt& p) {
        if (is_inf(p)) return;
        if (T::f2_is_zero(p.y)) { set_inf(p); return; }
        E U, V, W, S, M, t;
        T::f2_dbl(U, p.y);
        T::f2_sqr(V, U);
        T::f2_mul(W, U, V);
        T::f2_mul(S, p.x, V);
        T::f2_sqr(M, p.x);
        T::f2_triple(M, M);
        T::f2_mul(U, W, p.y);
        T::f2_sqr(p.x, M);
        T::f2_sub(p.x, p.x, S); T::f2_sub(p.x, p.x, S);
        T::f2_sub(t, S, p.x);
        T::f2_mul(t, M, t);
        T::f2_sub(p.y, t, U);
        T::f2_mul(p.zz, V, p.zz);
        T::f2_mul(p.zzz, W, p.zzz);
    }
    // p <- p + a, a affine (madd-2008-s), complete
    static B200_HD_NOINLINE void madd(Pt& p, const Aff& a) {
        if (aff_is_inf(a)) return;
        if (is_inf(p)) { from_affine(p, a); return; }
        E U2, S2, Pp, R, PP, PPP, Q, t;
        T::f2_mul(U2, a.x, p.zz);
        T::f2_mul(S2, a.y, p.zzz);
        T::f2_sub(Pp, U2, p.x);
        T::f2_sub(R, S2, p.y);
        if (T::f2_is_zero(Pp)) {
            if (T::f2_is_zero(R)) dbl_affine(p, a); else set_inf(p);
            return;
        }
        T::f2_sqr(PP, Pp);
        T::f2_mul(PPP, Pp, PP);
        T::f2_mul(Q, p.x, PP);
        T::f2_sqr(t, R);
        T::f2_sub(t, t, PPP); T::f2_sub(t, t, Q); T::f2_sub(t, t, Q);   // X3
        T::f2_sub(Q, Q, t);
        T::f2_mul(Q, R, Q);
        T::f2_mul(S2, p.y, PPP);
        T::f2_sub(p.y, Q, S2);
        p.x = t;
        T::f2_mul(p.zz, p.zz, PP);
        T::f2_mul(p.zzz, p.zzz, PPP);
    }
    static B200_HD void neg_y(Aff& a) { T::f2_neg(a.y, a.y); }
    // p <- p + q (add-2008-s), complete: the bucket sums of the G2 MSM (msm.cuh)
    static B200_HD_NOINLINE void add(Pt& p, const Pt& q) {
        if (is_inf(q)) return;
        if (is_inf(p)) { p = q; return; }
        E U1, U2, S1, S2, Pp, R, PP, PPP, Q, t;
        T::f2_mul(U1, p.x, q.zz);
        T::f2_mul(U2, q.x, p.zz);
        T::f2_mul(S1, p.y, q.zzz);
        T::f2_mul(S2, q.y, p.zzz);
        T::f2_sub(Pp, U2, U1);
        T::f2_sub(R, S2, S1);
        if (T::f2_is_zero(Pp)) {
            if (T::f2_is_zero(R)) dbl(p); else set_inf(p);
            return;
        }
        T::f2_sqr(PP, Pp);
        T::f2_mul(PPP, Pp, PP);
        T::f2_mul(Q, U1, PP);
        T::f2_sqr(t, R);
        T::f2_sub(t, t, PPP); T::f2_sub(t, t, Q); T::f2_sub(t, t, Q);   // X3
        T::f2_sub(Q, Q, t);
        T::f2_mul(Q, R, Q);
        T::f2_mul(S1, S1, PPP);
        T::f2_sub(p.y, Q, S1);
        p.x = t;
        T::f2_mul(p.zz, p.zz, q.zz);
        T::f2_mul(p.zz, p.zz, PP);
        T::f2_mul(p.zzz, p.zzz, q.zzz);
        T::f2_mul(p.zzz, p.zzz, PPP);
    }
    // affine (x,y) = (X/ZZ, Y/ZZZ); infinity -> (0,0)
    static B200_HD void to_affine(Aff& a, const Pt& p) {
        if (is_inf(p)) { T::f2_zero(a.x); T::f2_zero(a.y); return; }
        E t, i;
        T::f2_mul(t, p.zz, p.zzz);
        T::f2_inv(i, t);
        T::f2_mul(t, i, p.zzz);      // 1/ZZ
        T::f2_mul(a.x, p.x, t);
        T::f2_mul(t, i, p.zz);       // 1/ZZZ
        T::f2_mul(a.y, p.y, t);
    }
    // [k]P, k = 8 little-endian words (any 256-bit value: G2.Mul takes a Zr, already < r)
    static B200_HD void scalar_mul(Pt& acc, const Aff& base, const uint32_t* k) {
        set_inf(acc);
        int top = 255;
        while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1)) top--;
        for (int i = top; i >= 0; i--) {
            dbl(acc);
            if ((k[i >> 5] >> (i & 31)) & 1) madd(acc, base);
        }
    }
};

template <class C>
struct G2Codec {
    static constexpr int N = C::N;
    typedef Codec<C> CD;
    // G2 on the wire: X.A1|X.A0|Y.A1|Y.A0; infinity flag 0x40 on the 3-flag-bit curves, all-zero on BN254
    static B200_HD void g2_store(uint8_t* d, const G2Aff<N>& q, bool mont) {
        if (mont) {
            uint32_t* w = (uint32_t*)d;
            CD::fp_to_mont_words(w, q.x.c0); CD::fp_to_mont_words(w + N, q.x.c1);
            CD::fp_to_mont_words(w + 2 * N, q.y.c0); CD::fp_to_mont_words(w + 3 * N, q.y.c1);
            return;
        }
        const int FB = C::FP_BYTES;
        CD::fp_to_bytes(d, q.x.c1);
        CD::fp_to_bytes(d + FB, q.x.c0);
        CD::fp_to_bytes(d + 2 * FB, q.y.c1);
        CD::fp_to_bytes(d + 3 * FB, q.y.c0);
        if (C::FLAG_BITS == 3 && Tower<C>::f2_is_zero(q.x) && Tower<C>::f2_is_zero(q.y)) d[0] |= 0x40;
    }
};

#if defined(__CUDACC__)
#define B200_G2_THREADS 64

template <class C>
__global__ void __launch_bounds__(B200_G2_THREADS)
g2_mul_kernel(size_t n, const uint8_t* pts, const uint8_t* scalars, uint8_t* out, uint32_t flags, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef Codec<C> CD;
    typedef G2Ops<C> G;
    int e = 0;
    typename G::Aff a;
    CD::g2_load(a, pts + i * CD::g2_size(), flags & FLAG_IN_MONT, &e);
    if (e) { atomicExch(err, 1); CD::store_zero(out + i * CD::g2_size(), CD::g2_size()); return; }
    uint32_t k[8];
    CD::scalar_load(k, scalars + i * 32);
    typename G::Pt acc;
    G::scalar_mul(acc, a, k);
    G::to_affine(a, acc);
    G2Codec<C>::g2_store(out + i * CD::g2_size(), a, flags & FLAG_OUT_MONT);
}

// sum of n (small) affine G2 points -> one affine point; single thread (G2.Add is the n = 2 case)
template <class C>
__global__ void g2_sum_kernel(size_t n, const uint8_t* pts, uint8_t* out, uint32_t flags, int* err) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    typedef Codec<C> CD;
    typedef G2Ops<C> G;
    typename G::Pt acc;
    G::set_inf(acc);
    int e = 0;
    for (size_t i = 0; i < n; i++) {
        typename G::Aff a;
        CD::g2_load(a, pts + i * CD::g2_size(), flags & FLAG_IN_MONT, &e);
        G::madd(acc, a);
    }
    if (e) { atomicExch(err, 1); Codec<C>::store_zero(out, Codec<C>::g2_size()); return; }
    typename G::Aff r;
    G::to_affine(r, acc);
    G2Codec<C>::g2_store(out, r, flags & FLAG_OUT_MONT);
}
#endif  // __CUDACC__

}  // namespace b200
