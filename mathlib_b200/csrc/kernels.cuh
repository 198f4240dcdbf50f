// Batch kernels behind the C ABI (include/b200.h): one independent operation per thread.
//   pairing_kernel  -> b200_pairing_batch / b200_pairing2_batch  (driver.Curve.Pairing / Pairing2, + fused FExp)
//   fexp_kernel     -> b200_fexp_batch                           (driver.Curve.FExp)
//   g1_mul_kernel   -> b200_g1_mul_batch                         (driver.G1.Mul)
//   g1_mul2_kernel  -> b200_g1_mul2_batch                        (driver.G1.Mul2 / Mul2InPlace)
//   msm_*           -> b200_g1_msm                               (driver.Curve.MultiScalarMul), see msm.cuh
// plus the on-device codecs between the reference's Bytes() formats and Montgomery limbs.
#pragma once
#include "pairing.cuh"
#include "g1.cuh"

namespace b200 {

enum : uint32_t { FLAG_FEXP = 1u, FLAG_IN_MONT = 2u, FLAG_OUT_MONT = 4u, FLAG_UNITY = 8u };

// ------------------------------------------------------------------------------------------
// codecs (device)
// ------------------------------------------------------------------------------------------
template <class C>
struct Codec {
    static constexpr int N = C::N;
    static constexpr int FB = C::FP_BYTES;
    typedef FpOps<C> F;
    typedef Fp<N> E;

    static B200_HD uint8_t flag_mask() { return C::FLAG_BITS == 3 ? 0xE0 : 0xC0; }

    // big-endian canonical bytes -> Montgomery. top_mask clears encoding flag bits of byte 0.
    static B200_HD void fp_from_bytes(E& r, const uint8_t* s, uint8_t top_mask, int* err) {
        E t;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint8_t* q = s + FB - 4 * (i + 1);
            uint32_t b0 = q[0];
            if (i == N - 1) b0 &= top_mask;
            t.l[i] = (b0 << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
        }
        // canonical check: t < p
        const uint32_t* p = C::p();
        bool lt = false, decided = false;
        for (int i = N - 1; i >= 0; i--) {
            if (!decided && t.l[i] != p[i]) { lt = t.l[i] < p[i]; decided = true; }
        }
        if (!lt) { *err = 1; F::zero(r); return; }
        F::to_mont(r, t);
    }
    static B200_HD void fp_to_bytes(uint8_t* d, const E& a) {
        E t;
        F::from_mont(t, a);
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint8_t* q = d + FB - 4 * (i + 1);
            q[0] = (uint8_t)(t.l[i] >> 24); q[1] = (uint8_t)(t.l[i] >> 16); q[2] = (uint8_t)(t.l[i] >> 8); q[3] = (uint8_t)t.l[i];
        }
    }
    static B200_HD void fp_from_mont_words(E& r, const uint32_t* s) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = s[i];
    }
    static B200_HD void fp_to_mont_words(uint32_t* d, const E& a) {
#pragma unroll
        for (int i = 0; i < N; i++) d[i] = a.l[i];
    }

    static B200_HD size_t g1_size() { return 2 * (size_t)FB; }
    static B200_HD size_t g2_size() { return 4 * (size_t)FB; }
    static B200_HD size_t gt_size() { return 12 * (size_t)FB; }

    // G1: x|y ; infinity -> (0,0)
    static B200_HD void g1_load(Fp<N>& x, Fp<N>& y, const uint8_t* s, bool mont, int* err) {
        if (mont) {
            fp_from_mont_words(x, (const uint32_t*)s);
            fp_from_mont_words(y, (const uint32_t*)s + N);
            return;
        }
        uint8_t fl = s[0] & flag_mask();
        if (C::FLAG_BITS == 3 && fl == 0x40) { F::zero(x); F::zero(y); return; }
        if (fl != 0) { *err = 1; F::zero(x); F::zero(y); return; }   // compressed encodings are not accepted here
        fp_from_bytes(x, s, (uint8_t)~flag_mask(), err);
        fp_from_bytes(y, s + FB, 0xFF, err);
    }
    // defined output of an item whose input was rejected (B200_ERR_ENCODING): all-zero bytes
    static B200_HD void store_zero(uint8_t* d, size_t bytes) {
        for (size_t i = 0; i < bytes; i++) d[i] = 0;
    }
    static B200_HD void g1_store(uint8_t* d, const Fp<N>& x, const Fp<N>& y, bool mont) {
        if (mont) {
            fp_to_mont_words((uint32_t*)d, x);
            fp_to_mont_words((uint32_t*)d + N, y);
            return;
        }
        fp_to_bytes(d, x);
        fp_to_bytes(d + FB, y);
        if (C::FLAG_BITS == 3 && F::is_zero(x) && F::is_zero(y)) d[0] |= 0x40;
    }
    // G2: X.A1|X.A0|Y.A1|Y.A0 on the wire
    static B200_HD void g2_load(G2Aff<N>& q, const uint8_t* s, bool mont, int* err) {
        if (mont) {
            const uint32_t* w = (const uint32_t*)s;
            fp_from_mont_words(q.x.c0, w); fp_from_mont_words(q.x.c1, w + N);
            fp_from_mont_words(q.y.c0, w + 2 * N); fp_from_mont_words(q.y.c1, w + 3 * N);
            return;
        }
        uint8_t fl = s[0] & flag_mask();
        if (C::FLAG_BITS == 3 && fl == 0x40) { Tower<C>::f2_zero(q.x); Tower<C>::f2_zero(q.y); return; }
        if (fl != 0) { *err = 1; Tower<C>::f2_zero(q.x); Tower<C>::f2_zero(q.y); return; }
        fp_from_bytes(q.x.c1, s, (uint8_t)~flag_mask(), err);
        fp_from_bytes(q.x.c0, s + FB, 0xFF, err);
        fp_from_bytes(q.y.c1, s + 2 * FB, 0xFF, err);
        fp_from_bytes(q.y.c0, s + 3 * FB, 0xFF, err);
    }
    // Gt: bytes are the 12 Fp of the struct in reverse order
    static B200_HD void gt_load(Fp12<N>& f, const uint8_t* s, bool mont, int* err) {
        Fp<N>* e = reinterpret_cast<Fp<N>*>(&f);
        if (mont) {
            for (int k = 0; k < 12; k++) fp_from_mont_words(e[k], (const uint32_t*)s + k * N);
            return;
        }
        for (int k = 0; k < 12; k++) fp_from_bytes(e[11 - k], s + k * FB, 0xFF, err);
    }
    static B200_HD void gt_store(uint8_t* d, const Fp12<N>& f, bool mont) {
        const Fp<N>* e = reinterpret_cast<const Fp<N>*>(&f);
        if (mont) {
            for (int k = 0; k < 12; k++) fp_to_mont_words((uint32_t*)d + k * N, e[k]);
            return;
        }
        for (int k = 0; k < 12; k++) fp_to_bytes(d + k * FB, e[11 - k]);
    }
    // 32-byte big-endian scalar -> 8 little-endian words
    static B200_HD void scalar_load(uint32_t* k, const uint8_t* s) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint8_t* q = s + 32 - 4 * (i + 1);
            k[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
        }
    }
};

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
#define B200_PAIR_THREADS 64

template <class C, int NP, int THREADS = B200_PAIR_THREADS, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
pairing_kernel(size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
               uint8_t* out, uint32_t flags, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef Codec<C> CD;
    typedef PairingOps<C> PO;
    constexpr int N = C::N;
    const bool in_mont = flags & FLAG_IN_MONT;
    int e = 0;
    G1Aff<N> P[NP];
    G2Aff<N> Q[NP];
    CD::g1_load(P[0].x, P[0].y, g1a + i * CD::g1_size(), in_mont, &e);
    CD::g2_load(Q[0], g2a + i * CD::g2_size(), in_mont, &e);
    if (NP == 2) {
        CD::g1_load(P[NP - 1].x, P[NP - 1].y, g1b + i * CD::g1_size(), in_mont, &e);
        CD::g2_load(Q[NP - 1], g2b + i * CD::g2_size(), in_mont, &e);
    }
    if (e) {
        atomicExch(err, 1);
        if (flags & FLAG_UNITY) out[i] = 0; else CD::store_zero(out + i * CD::gt_size(), CD::gt_size());
        return;
    }
    Fp12<N> f;
    PO::template miller_loop<NP>(f, P, Q);
    if (flags & FLAG_FEXP) PO::final_exp(f, f);
    if (flags & FLAG_UNITY) out[i] = Tower<C>::f12_is_one(f) ? 1 : 0;
    else CD::gt_store(out + i * CD::gt_size(), f, flags & FLAG_OUT_MONT);
}

template <class C>
__global__ void __launch_bounds__(B200_PAIR_THREADS)
fexp_kernel(size_t n, const uint8_t* in, uint8_t* out, uint32_t flags, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef Codec<C> CD;
    int e = 0;
    Fp12<C::N> f;
    CD::gt_load(f, in + i * CD::gt_size(), flags & FLAG_IN_MONT, &e);
    if (e) {
        atomicExch(err, 1);
        if (flags & FLAG_UNITY) out[i] = 0; else CD::store_zero(out + i * CD::gt_size(), CD::gt_size());
        return;
    }
    if (flags & FLAG_FEXP) PairingOps<C>::final_exp(f, f);
    if (flags & FLAG_UNITY) out[i] = Tower<C>::f12_is_one(f) ? 1 : 0;
    else CD::gt_store(out + i * CD::gt_size(), f, flags & FLAG_OUT_MONT);
}

#define B200_G1_THREADS 128
#ifndef B200_G1_MIN_BLOCKS
#define B200_G1_MIN_BLOCKS 3        // 12 warps per SM at <= 168 registers: the accumulator point and the formula temporaries stay in registers
#endif

template <class C>
__global__ void __launch_bounds__(B200_G1_THREADS, B200_G1_MIN_BLOCKS)
g1_mul_kernel(size_t n, const uint8_t* pts, const uint8_t* scalars, uint8_t* out, uint32_t flags, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef Codec<C> CD;
    typedef G1Ops<C> G;
    int e = 0;
    typename G::Aff a;
    CD::g1_load(a.x, a.y, pts + i * CD::g1_size(), flags & FLAG_IN_MONT, &e);
    if (e) { atomicExch(err, 1); CD::store_zero(out + i * CD::g1_size(), CD::g1_size()); return; }
    uint32_t k[8];
    CD::scalar_load(k, scalars + i * 32);
    typename G::Pt acc;
    G::scalar_mul(acc, a, k);
    G::to_affine(a, acc);
    CD::g1_store(out + i * CD::g1_size(), a.x, a.y, flags & FLAG_OUT_MONT);
}

template <class C>
__global__ void __launch_bounds__(B200_G1_THREADS, B200_G1_MIN_BLOCKS)
g1_mul2_kernel(size_t n, const uint8_t* P, const uint8_t* es, const uint8_t* Q, const uint8_t* fs, uint8_t* out,
               uint32_t flags, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef Codec<C> CD;
    typedef G1Ops<C> G;
    int e = 0;
    typename G::Aff a, b;
    CD::g1_load(a.x, a.y, P + i * CD::g1_size(), flags & FLAG_IN_MONT, &e);
    CD::g1_load(b.x, b.y, Q + i * CD::g1_size(), flags & FLAG_IN_MONT, &e);
    if (e) { atomicExch(err, 1); CD::store_zero(out + i * CD::g1_size(), CD::g1_size()); return; }
    uint32_t ke[8], kf[8];
    CD::scalar_load(ke, es + i * 32);
    CD::scalar_load(kf, fs + i * 32);
    typename G::Pt acc;
    G::scalar_mul2(acc, a, ke, b, kf);
    G::to_affine(a, acc);
    CD::g1_store(out + i * CD::g1_size(), a.x, a.y, flags & FLAG_OUT_MONT);
}

// sum of n (small) affine points -> one affine point; single thread
template <class C>
__global__ void g1_sum_kernel(size_t n, const uint8_t* pts, uint8_t* out, uint32_t flags, int* err) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    typedef Codec<C> CD;
    typedef G1Ops<C> G;
    typename G::Pt acc;
    G::set_inf(acc);
    int e = 0;
    for (size_t i = 0; i < n; i++) {
        typename G::Aff a;
        CD::g1_load(a.x, a.y, pts + i * CD::g1_size(), flags & FLAG_IN_MONT, &e);
        G::madd(acc, a);
    }
    if (e) { atomicExch(err, 1); CD::store_zero(out, CD::g1_size()); return; }
    typename G::Aff r;
    G::to_affine(r, acc);
    CD::g1_store(out, r.x, r.y, flags & FLAG_OUT_MONT);
}
#endif  // __CUDACC__

}  // namespace b200
