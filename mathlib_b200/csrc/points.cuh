// Point (de)serialisation and validation for whole batches: the device side of
//   NewG1FromCompressed / NewG2FromCompressed / NewG1FromBytes / NewG2FromBytes
//       reference driver/gurvy/bn254.go:339-377, bls12381/bls12-381.go:531-569, kilic/bls12-381.go:344-394
//       (gnark `SetBytes`, kilic `FromCompressed` / `FromUncompressed`: canonical coordinates, on-curve, subgroup)
//   G1.Compressed / G2.Compressed
//       reference bn254.go:82-86,167-171, bls12381/bls12-381.go:292-296,379-383, kilic/bls12-381.go:81-85,159-163
// Encodings: SURVEY A.3.  The square root is Tonelli-Shanks over p-1 = 2^s q
// (s = 1 for BN254 and BLS12-381, where it degenerates to a^((p+1)/4); s = 46 for BLS12-377); Fp2 roots go through the
// norm.  Which root is returned never matters: the sign flag selects y or -y, so the output bytes are unique.
#pragma once
#include "g2.cuh"

namespace b200 {

enum : uint32_t { FLAG_NO_SUBGROUP = 0x40u };      // B200_NO_SUBGROUP_CHECK

template <class C>
struct PointCodec {
    static constexpr int N = C::N;
    static constexpr int FB = C::FP_BYTES;
    typedef FpOps<C> F;
    typedef Tower<C> T;
    typedef Codec<C> CD;
    typedef Fp<N> E;
    typedef Fp2<N> E2;

    static B200_HD void load_const(E& r, const uint32_t* w) {
        for (int i = 0; i < N; i++) r.l[i] = w[i];
    }
    // r = a^e, e = nw little-endian words
    static B200_HD_NOINLINE void fp_pow(E& r, const E& a, const uint32_t* e, int nw) {
        E acc;
        F::one(acc);
        bool started = false;
        for (int i = nw * 32 - 1; i >= 0; i--) {
            if (started) F::sqrx(acc, acc);
            if ((e[i >> 5] >> (i & 31)) & 1) {
                if (started) F::mulx(acc, acc, a); else { acc = a; started = true; }
            }
        }
        r = acc;
    }
    // square root in Fp; returns false if a is not a square
    static B200_HD_NOINLINE bool fp_sqrt(E& r, const E& a) {
        if (F::is_zero(a)) { F::zero(r); return true; }
        E w, x, t, c, one;
        F::one(one);
        fp_pow(w, a, C::K().ts_exp, N);          // a^((q-1)/2)
        F::mulx(x, a, w);                        // a^((q+1)/2)
        F::mulx(t, x, w);                        // a^q
        load_const(c, C::K().ts_z);
        int m = C::TS_S;
        while (!F::eq(t, one)) {
            // least i, 0 < i < m, with t^(2^i) = 1
            E t2 = t;
            int i = 0;
            while (!F::eq(t2, one)) {
                F::sqrx(t2, t2);
                i++;
                if (i == m) return false;        // t has order 2^m: a is a non-residue
            }
            E b = c;
            for (int j = 0; j < m - i - 1; j++) F::sqrx(b, b);
            F::mulx(x, x, b);
            F::sqrx(c, b);
            F::mulx(t, t, c);
            m = i;
        }
        r = x;
        return true;
    }
    // square root in Fp2 = Fp[u]/(u^2 - BETA); returns false if a is not a square
    static B200_HD_NOINLINE bool f2_sqrt(E2& r, const E2& a) {
        if (T::f2_is_zero(a)) { T::f2_zero(r); return true; }
        E2 cand;
        if (F::is_zero(a.c1)) {
            if (fp_sqrt(cand.c0, a.c0)) { F::zero(cand.c1); r = cand; return true; }
            // a0 is a non-residue: a0 = BETA * (a0 / BETA) and u^2 = BETA, so sqrt = sqrt(a0 / BETA) * u
            E bi, beta;
            F::one(beta);
            T::fp_mul_beta(beta, beta);
            F::inv(bi, beta);
            F::mulx(bi, bi, a.c0);
            if (!fp_sqrt(cand.c1, bi)) return false;
            F::zero(cand.c0);
            r = cand;
            return true;
        }
        E n, s, d, t;
        F::sqrx(n, a.c0);
        F::sqrx(t, a.c1);
        T::fp_mul_beta(t, t);
        F::sub(n, n, t);                         // norm a0^2 - BETA a1^2
        if (!fp_sqrt(s, n)) return false;
        F::add(d, a.c0, s);
        F::halve(d, d);
        if (!fp_sqrt(cand.c0, d)) {
            F::sub(d, a.c0, s);
            F::halve(d, d);
            if (!fp_sqrt(cand.c0, d)) return false;
        }
        F::dbl(t, cand.c0);
        F::inv(t, t);
        F::mulx(cand.c1, a.c1, t);
        E2 chk;
        T::f2_sqr(chk, cand);
        if (!T::f2_eq(chk, a)) return false;
        r = cand;
        return true;
    }
    // canonical value > (p-1)/2 ?
    static B200_HD bool fp_is_largest(const E& a_mont) {
        E a;
        F::from_mont(a, a_mont);
        const uint32_t* h = C::K().half_p;
        for (int i = N - 1; i >= 0; i--)
            if (a.l[i] != h[i]) return a.l[i] > h[i];
        return false;
    }
    static B200_HD bool f2_is_largest(const E2& y) { return F::is_zero(y.c1) ? fp_is_largest(y.c0) : fp_is_largest(y.c1); }

    // flag bytes
    static B200_HD uint8_t fl_small() { return 0x80; }
    static B200_HD uint8_t fl_large() { return C::FLAG_BITS == 3 ? 0xA0 : 0xC0; }
    static B200_HD uint8_t fl_cinf() { return C::FLAG_BITS == 3 ? 0xC0 : 0x40; }

    static B200_HD bool g1_on_curve(const E& x, const E& y) {
        E l, r, b;
        F::sqrx(l, y);
        F::sqrx(r, x);
        F::mulx(r, r, x);
        load_const(b, C::K().b);
        F::add(r, r, b);
        return F::eq(l, r);
    }
    static B200_HD bool g2_on_curve(const E2& x, const E2& y) {
        E2 l, r, b;
        T::f2_sqr(l, y);
        T::f2_sqr(r, x);
        T::f2_mul(r, r, x);
        load_const(b.c0, C::K().btw);
        load_const(b.c1, C::K().btw + N);
        T::f2_add(r, r, b);
        return T::f2_eq(l, r);
    }
    // [r]P == O with the plain ladder (the GLV ladder of g1.cuh presupposes the subgroup)
    static B200_HD bool g1_in_subgroup(const G1Affine<N>& a) {
        if (C::FAMILY == FAMILY_BN) return true;            // BN254 G1 has cofactor 1
        typedef G1Ops<C> G;
        typename G::Pt acc;
        G::set_inf(acc);
        const uint32_t* r = C::K().order;
        for (int i = 255; i >= 0; i--) {
            G::dbl(acc);
            if ((r[i >> 5] >> (i & 31)) & 1) G::madd(acc, a);
        }
        return G::is_inf(acc);
    }
    static B200_HD bool g2_in_subgroup(const G2Aff<N>& a) {
        typedef G2Ops<C> G;
        typename G::Pt acc;
        G::scalar_mul(acc, a, C::K().order);
        return G::is_inf(acc);
    }

    // compressed G1 -> affine; returns 0 ok, 1 bad encoding / not on curve / not in subgroup
    static B200_HD int g1_decompress(G1Affine<N>& a, const uint8_t* s, bool check_subgroup) {
        const uint8_t fl = s[0] & CD::flag_mask();
        if (fl == fl_cinf()) {
            for (int i = 0; i < FB; i++)
                if ((i == 0 ? (s[0] & (uint8_t)~CD::flag_mask()) : s[i]) != 0) return 1;
            F::zero(a.x); F::zero(a.y);
            return 0;
        }
        if (fl != fl_small() && fl != fl_large()) return 1;
        int e = 0;
        CD::fp_from_bytes(a.x, s, (uint8_t)~CD::flag_mask(), &e);
        if (e) return 1;
        E rhs, b;
        F::sqrx(rhs, a.x);
        F::mulx(rhs, rhs, a.x);
        load_const(b, C::K().b);
        F::add(rhs, rhs, b);
        if (!fp_sqrt(a.y, rhs)) return 1;
        if (fp_is_largest(a.y) != (fl == fl_large())) F::neg(a.y, a.y);
        if (check_subgroup && !g1_in_subgroup(a)) return 1;
        return 0;
    }
    static B200_HD int g2_decompress(G2Aff<N>& q, const uint8_t* s, bool check_subgroup) {
        const uint8_t fl = s[0] & CD::flag_mask();
        if (fl == fl_cinf()) {
            for (int i = 0; i < 2 * FB; i++)
                if ((i == 0 ? (s[0] & (uint8_t)~CD::flag_mask()) : s[i]) != 0) return 1;
            T::f2_zero(q.x); T::f2_zero(q.y);
            return 0;
        }
        if (fl != fl_small() && fl != fl_large()) return 1;
        int e = 0;
        CD::fp_from_bytes(q.x.c1, s, (uint8_t)~CD::flag_mask(), &e);
        CD::fp_from_bytes(q.x.c0, s + FB, 0xFF, &e);
        if (e) return 1;
        E2 rhs, b;
        T::f2_sqr(rhs, q.x);
        T::f2_mul(rhs, rhs, q.x);
        load_const(b.c0, C::K().btw);
        load_const(b.c1, C::K().btw + N);
        T::f2_add(rhs, rhs, b);
        if (!f2_sqrt(q.y, rhs)) return 1;
        if (f2_is_largest(q.y) != (fl == fl_large())) T::f2_neg(q.y, q.y);
        if (check_subgroup && !g2_in_subgroup(q)) return 1;
        return 0;
    }
    static B200_HD void g1_compress(uint8_t* d, const G1Affine<N>& a) {
        if (F::is_zero(a.x) && F::is_zero(a.y)) {
            for (int i = 0; i < FB; i++) d[i] = 0;
            d[0] = fl_cinf();
            return;
        }
        CD::fp_to_bytes(d, a.x);
        d[0] |= fp_is_largest(a.y) ? fl_large() : fl_small();
    }
    static B200_HD void g2_compress(uint8_t* d, const G2Aff<N>& q) {
        if (T::f2_is_zero(q.x) && T::f2_is_zero(q.y)) {
            for (int i = 0; i < 2 * FB; i++) d[i] = 0;
            d[0] = fl_cinf();
            return;
        }
        CD::fp_to_bytes(d, q.x.c1);
        CD::fp_to_bytes(d + FB, q.x.c0);
        d[0] |= f2_is_largest(q.y) ? fl_large() : fl_small();
    }
};

// one item of a batch; op: 0 decompress (compressed -> uncompressed / MONT), 1 compress (uncompressed / MONT ->
// compressed), 2 validate (uncompressed / MONT -> one verdict byte: canonical, on curve, in the subgroup).
// Returns 1 on an encoding error (ops 0, 1), else 0.
template <class C, int G2>
B200_HD int point_codec_item(int op, const uint8_t* in, uint8_t* out, uint32_t flags) {
    typedef Codec<C> CD;
    typedef PointCodec<C> PC;
    const bool subgroup = !(flags & FLAG_NO_SUBGROUP);
    if (G2) {
        G2Aff<C::N> q;
        if (op == 0) {
            if (PC::g2_decompress(q, in, subgroup)) return 1;
            G2Codec<C>::g2_store(out, q, flags & FLAG_OUT_MONT);
            return 0;
        }
        int e = 0;
        CD::g2_load(q, in, flags & FLAG_IN_MONT, &e);
        if (op == 1) {
            if (e) return 1;
            PC::g2_compress(out, q);
            return 0;
        }
        const bool inf = Tower<C>::f2_is_zero(q.x) && Tower<C>::f2_is_zero(q.y);
        out[0] = (!e && (inf || (PC::g2_on_curve(q.x, q.y) && (!subgroup || PC::g2_in_subgroup(q))))) ? 1 : 0;
        return 0;
    }
    G1Affine<C::N> a;
    if (op == 0) {
        if (PC::g1_decompress(a, in, subgroup)) return 1;
        CD::g1_store(out, a.x, a.y, flags & FLAG_OUT_MONT);
        return 0;
    }
    int e = 0;
    CD::g1_load(a.x, a.y, in, flags & FLAG_IN_MONT, &e);
    if (op == 1) {
        if (e) return 1;
        PC::g1_compress(out, a);
        return 0;
    }
    const bool inf = FpOps<C>::is_zero(a.x) && FpOps<C>::is_zero(a.y);
    out[0] = (!e && (inf || (PC::g1_on_curve(a.x, a.y) && (!subgroup || PC::g1_in_subgroup(a))))) ? 1 : 0;
    return 0;
}

// Batch affine normalisation with Montgomery's trick: Jacobian (X, Y, Z) -> affine (X / Z^2, Y / Z^3)
// for BATCH points per thread with ONE field inversion -- the device side of `Bytes()` / `BatchJacobianToAffine` over slabs of
// kilic PointG1 [3]fe (reference driver/kilic/bls12-381.go:20-23, 74-78) or gnark G1Jac values.  Input: Montgomery limbs
// X | Y | Z per point; Z = 0 is the point at infinity (-> (0, 0)).  Output: G1 BYTES or MONT.
#define B200_NORM_BATCH 8
template <class C>
B200_HD void g1_normalize_items(size_t n_items, const uint32_t* in, uint8_t* out, bool out_mont) {
    typedef FpOps<C> F;
    typedef Fp<C::N> E;
    constexpr int N = C::N;
    E z[B200_NORM_BATCH], pre[B200_NORM_BATCH], acc, inv;
    F::one(acc);
    for (size_t k = 0; k < n_items; k++) {
        for (int i = 0; i < N; i++) z[k].l[i] = in[(k * 3 + 2) * N + i];
        pre[k] = acc;                                   // product of the non-zero Z's before this one
        if (!F::is_zero(z[k])) F::mulx(acc, acc, z[k]);
    }
    F::inv(inv, acc);
    for (size_t kk = n_items; kk > 0; kk--) {
        const size_t k = kk - 1;
        E x, y;
        F::zero(x); F::zero(y);
        if (!F::is_zero(z[k])) {
            E zi, zi2, zi3, X, Y;
            F::mulx(zi, inv, pre[k]);                   // 1 / Z_k
            F::mulx(inv, inv, z[k]);                    // drop Z_k from the running inverse
            for (int i = 0; i < N; i++) { X.l[i] = in[(k * 3) * N + i]; Y.l[i] = in[(k * 3 + 1) * N + i]; }
            F::sqrx(zi2, zi);
            F::mulx(zi3, zi2, zi);
            F::mulx(x, X, zi2);
            F::mulx(y, Y, zi3);
        }
        Codec<C>::g1_store(out + k * Codec<C>::g1_size(), x, y, out_mont);
    }
}

#if defined(__CUDACC__)
template <class C>
__global__ void __launch_bounds__(64)
g1_normalize_kernel(size_t n, const uint32_t* in, uint8_t* out, uint32_t flags) {
    const size_t first = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * B200_NORM_BATCH;
    if (first >= n) return;
    const size_t cnt = n - first < B200_NORM_BATCH ? n - first : B200_NORM_BATCH;
    g1_normalize_items<C>(cnt, in + first * 3 * C::N, out + first * Codec<C>::g1_size(), flags & FLAG_OUT_MONT);
}
#endif

#if defined(__CUDACC__)
#define B200_PT_THREADS 64
template <class C, int G2>
__global__ void __launch_bounds__(B200_PT_THREADS)
point_codec_kernel(int op, size_t n, const uint8_t* in, uint8_t* out, uint32_t flags, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr size_t FB = C::FP_BYTES;
    const size_t usz = (G2 ? 4 : 2) * FB, csz = (G2 ? 2 : 1) * FB;
    const size_t isz = op == 0 ? csz : usz, osz = op == 0 ? usz : (op == 1 ? csz : 1);
    if (point_codec_item<C, G2>(op, in + i * isz, out + i * osz, flags)) {
        atomicExch(err, 1);
        for (size_t b = 0; b < osz; b++) out[i * osz + b] = 0;          // defined output for the rejected item
    }
}
#endif

}  // namespace b200
