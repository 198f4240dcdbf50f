// G1 Mul / Mul2 for small batches: three lanes per operation.
//
// One scalar multiplication is a dependency chain of ~230 point operations; with one thread per operation a small batch
// cannot fill the GPU and the call time is the latency of that chain (g1.cuh).  Inside a point operation most
// field products are independent, so here three consecutive lanes share them: every lane keeps a full copy of the
// state, computes one product per level and the three results are exchanged through shared memory -- 4 product latencies
// per doubling or mixed addition instead of 9 / 10.  Same results as G1Ops::scalar_mul / scalar_mul2 (canonical group
// elements; reference driver/math.go:260-266).  Control flow is warp-uniform (the ladder runs to the warp's longest
// scalar, additions are computed every step and committed by selection) because the exchange synchronises the warp.
// Used for batches of at most 2,048 operations (launch.cuh: p3_max): 1,000 Mul take 2.5 ms instead of 3.4 ms; from a few
// thousand operations on the one-lane kernels are faster.
#pragma once
#include "kernels.cuh"

namespace b200 {

#if defined(__CUDACC__)
#define B200_P3_THREADS 128                     // 4 warps; 10 operations per warp (lanes 30, 31 idle)
#define B200_P3_OPS_PER_BLOCK (B200_P3_THREADS / 32 * 10)

template <class C>
struct G1P3 {
    static constexpr int N = C::N;
    typedef FpOps<C> F;
    typedef G1Ops<C> G;
    typedef Fp<N> E;
    typedef typename G::Aff Aff;
    typedef typename G::Pt Pt;

    E* sh;          // this group's 3 exchange slots
    int sel;        // lane within the group (0..2)

    // r0, r1, r2 = a0*b0, a1*b1, a2*b2 -- one product per lane
    __device__ __forceinline__ void par3(E& r0, E& r1, E& r2, const E& a0, const E& b0, const E& a1, const E& b1, const E& a2,
                                         const E& b2) {
        E a = sel == 0 ? a0 : (sel == 1 ? a1 : a2);
        E b = sel == 0 ? b0 : (sel == 1 ? b1 : b2);
        E r;
        F::mulx(r, a, b);
        sh[sel] = r;
        __syncwarp();
        r0 = sh[0]; r1 = sh[1]; r2 = sh[2];
        __syncwarp();
    }
    // p <- 2p (dbl-2008-s-1); the point at infinity (all zero) maps to itself
    __device__ void dbl(Pt& p) {
        E U, V, W, S, M, M2, t, u, d0;
        F::dbl(U, p.y);
        par3(V, M, d0, U, U, p.x, p.x, U, U);                 // V = U^2, M = X^2
        F::dbl(t, M); F::add(M, M, t);                        // M = 3 X^2
        par3(W, S, d0, U, V, p.x, V, U, V);                   // W = U V, S = X V
        par3(M2, u, p.zz, M, M, W, p.y, V, p.zz);             // M^2, W*Y, ZZ3 = V*ZZ
        F::sub(p.x, M2, S); F::sub(p.x, p.x, S);              // X3
        F::sub(t, S, p.x);
        par3(t, p.zzz, d0, M, t, W, p.zzz, M, t);             // M*(S - X3), ZZZ3 = W*ZZZ
        F::sub(p.y, t, u);
    }
    // p <- p + a (madd-2008-s) if commit; complete (infinity, equal and opposite points) through the serial adder
    __device__ void madd(Pt& p, const Aff& a, bool commit) {
        E U2, S2, Pp, R, PP, R2, PPP, Q, ZZ3, t, u, v, d0;
        par3(U2, S2, d0, a.x, p.zz, a.y, p.zzz, a.x, p.zz);
        F::sub(Pp, U2, p.x);
        F::sub(R, S2, p.y);
        par3(PP, R2, d0, Pp, Pp, R, R, Pp, Pp);
        par3(PPP, Q, ZZ3, Pp, PP, p.x, PP, p.zz, PP);
        F::sub(t, R2, PPP); F::sub(t, t, Q); F::sub(t, t, Q); // X3
        F::sub(Q, Q, t);
        par3(u, v, d0, R, Q, p.y, PPP, p.zzz, PPP);           // R*(Q - X3), Y*PPP, ZZZ3 = ZZZ*PPP
        if (!commit || G::aff_is_inf(a)) return;
        if (G::is_inf(p) || F::is_zero(Pp)) {                 // same in the three lanes of the group: no exchange inside
            G::madd(p, a);
            return;
        }
        p.x = t;
        F::sub(p.y, u, v);
        p.zz = ZZ3;
        p.zzz = d0;
    }
};

// ops: 1 = [k]P (Mul), 2 = [e]P + [f]Q (Mul2).  Operation i is handled by lanes 3*(i%10) .. +2 of warp i/10.
template <class C, int OPS>
__global__ void __launch_bounds__(B200_P3_THREADS)
g1_mul_p3_kernel(size_t n, const uint8_t* P, const uint8_t* es, const uint8_t* Q, const uint8_t* fs, uint8_t* out,
                 uint32_t flags, int* err) {
    typedef Codec<C> CD;
    typedef G1Ops<C> G;
    typedef G1P3<C> P3;
    __shared__ typename P3::E shbuf[B200_P3_THREADS / 32][11][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / 3;                                  // 0..10 (group 10 = lanes 30, 31: spare)
    const size_t item = ((size_t)blockIdx.x * (B200_P3_THREADS / 32) + warp) * 10 + grp;
    const bool active = grp < 10 && item < n;
    const size_t it = active ? item : 0;
    P3 X;
    X.sh = shbuf[warp][grp];
    X.sel = lane % 3;
    int e = 0;
    typename G::Aff a, b;
    CD::g1_load(a.x, a.y, P + it * CD::g1_size(), flags & FLAG_IN_MONT, &e);
    uint32_t ke[8], kf[8];
    CD::scalar_load(ke, es + it * 32);
    if (OPS == 2) {
        CD::g1_load(b.x, b.y, Q + it * CD::g1_size(), flags & FLAG_IN_MONT, &e);
        CD::scalar_load(kf, fs + it * 32);
    }
    // a rejected encoding fails its own operation only (flag raised, all-zero output); g1_load has zeroed the point, so
    // the three lanes run the ladder on infinity in lock-step with the rest of the warp
    const bool bad = active && e != 0;
    if (bad && X.sel == 0) atomicExch(err, 1);
    typename G::Pt acc;
    G::set_inf(acc);
    if (C::FAMILY == FAMILY_BLS12) {
        uint32_t e1[5], e2[5], f1[5], f2[5];
        typename G::GlvTable tp, tq;
        G::glv_split(e1, e2, ke);
        G::glv_table(tp, a);
        if (OPS == 2) {
            G::glv_split(f1, f2, kf);
            G::glv_table(tq, b);
        }
        int top = 159;
        while (top >= 0 && !(G::glv_bits(e1, e2, top) | (OPS == 2 ? G::glv_bits(f1, f2, top) : 0u))) top--;
        top = __reduce_max_sync(0xffffffffu, active ? top : -1);
        for (int i = top; i >= 0; i--) {
            X.dbl(acc);
            const uint32_t bp = G::glv_bits(e1, e2, i);
            typename G::Aff s;
            s.x = bp == 2 ? tp.x2 : (bp == 3 ? tp.x3 : tp.x1);
            s.y = bp == 3 ? tp.ny : tp.y;
            X.madd(acc, s, bp != 0);
            if (OPS == 2) {
                const uint32_t bq = G::glv_bits(f1, f2, i);
                s.x = bq == 2 ? tq.x2 : (bq == 3 ? tq.x3 : tq.x1);
                s.y = bq == 3 ? tq.ny : tq.y;
                X.madd(acc, s, bq != 0);
            }
        }
    } else {
        int top = 255;
        while (top >= 0 && !(G::scalar_bit(ke, top) | (OPS == 2 ? G::scalar_bit(kf, top) : 0))) top--;
        top = __reduce_max_sync(0xffffffffu, active ? top : -1);
        for (int i = top; i >= 0; i--) {
            X.dbl(acc);
            X.madd(acc, a, G::scalar_bit(ke, i) != 0);
            if (OPS == 2) X.madd(acc, b, G::scalar_bit(kf, i) != 0);
        }
    }
    typename G::Aff r;
    G::to_affine(r, acc);
    if (active && X.sel == 0) {
        if (bad) CD::store_zero(out + item * CD::g1_size(), CD::g1_size());
        else CD::g1_store(out + item * CD::g1_size(), r.x, r.y, flags & FLAG_OUT_MONT);
    }
}
#endif  // __CUDACC__

}  // namespace b200
