// Montgomery Fp arithmetic on 32-bit limbs for sm_100a.
//
// Replaces the Fp layer under gnark-crypto / kilic that mathlib's adapters call (reference
// driver/gurvy/bn254.go:248-267, driver/kilic/bls12-381.go:260-277); the one Fp multiply that
// lives inside the reference, driver/kilic/custom_generic.go:57-175 (6x64 CIOS), computes the
// same function a*b*R^-1 mod p with R = 2^384 -- in-memory limbs are bit-identical
// (12 x u32 little-endian == 6 x u64 little-endian).
//
// Design: one field element per thread, N = 8 (BN254) or 12 (BLS12-381/377) limbs in
// registers.  Multiplication is operand-scanning Montgomery with the partial products split
// into an even-aligned and an odd-aligned accumulator so that every 32x32->64 product is a
// mad.lo.cc/madc.hi.cc pair (one IMAD.WIDE.U32 in SASS) on one of two independent carry
// chains -- the chains interleave in the FMA pipe instead of serialising on the carry flag.
//
// The same source compiles for the host (tests/hostemu): the PTX carry-chain primitives
// have a C emulation below, used ONLY by the CPU-side test harness to check kernel logic
// without a GPU.  No product entry point reaches the host path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_HD_NOINLINE __host__ __device__ __noinline__
#else
#define B200_HD inline
#define B200_HD_NOINLINE
#endif

namespace b200 {

// ---------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define B200_ASM_R3(name, ptx)                                                              \
    static __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b) {               \
        uint32_t r;                                                                         \
        asm volatile(ptx " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                        \
        return r;                                                                           \
    }
#define B200_ASM_R4(name, ptx)                                                              \
    static __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b, uint32_t c) {   \
        uint32_t r;                                                                         \
        asm volatile(ptx " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));            \
        return r;                                                                           \
    }
B200_ASM_R3(add_cc, "add.cc.u32")
B200_ASM_R3(addc_cc, "addc.cc.u32")
B200_ASM_R3(addc, "addc.u32")
B200_ASM_R3(sub_cc, "sub.cc.u32")
B200_ASM_R3(subc_cc, "subc.cc.u32")
B200_ASM_R3(subc, "subc.u32")
B200_ASM_R4(mad_lo_cc, "mad.lo.cc.u32")
B200_ASM_R4(madc_lo_cc, "madc.lo.cc.u32")
B200_ASM_R4(mad_hi_cc, "mad.hi.cc.u32")
B200_ASM_R4(madc_hi_cc, "madc.hi.cc.u32")
B200_ASM_R4(madc_hi, "madc.hi.u32")
B200_ASM_R4(madc_lo, "madc.lo.u32")
static __device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
static __device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
// Host emulation of the PTX condition-code register (test harness only).
static thread_local uint32_t g_cc = 0;
static inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cc; }
static inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; g_cc = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
static inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; g_cc = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
static inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cc; }
static inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
static inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)(((uint64_t)a * b) >> 32) + c + g_cc; }
static inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((uint64_t)a * b) + c + g_cc; }
static inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
static inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
#endif

// ---------------------------------------------------------------------------------------
// field element
// ---------------------------------------------------------------------------------------
template <int N>
struct alignas(16) Fp {
    uint32_t l[N];
};

// C: curve traits (curves.cuh) -- C::N, C::p() (modulus limbs), C::inv32(), C::one()
template <class C>
struct FpOps {
    static constexpr int N = C::N;
    typedef Fp<N> E;

    static B200_HD void zero(E& r) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
    }
    static B200_HD void one(E& r) {
        const uint32_t* o = C::one();
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = o[i];
    }
    static B200_HD bool is_zero(const E& a) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= a.l[i];
        return t == 0;
    }
    static B200_HD bool eq(const E& a, const E& b) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= a.l[i] ^ b.l[i];
        return t == 0;
    }
    // r = c ? a : r
    static B200_HD void cmov(E& r, const E& a, bool c) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = c ? a.l[i] : r.l[i];
    }

    // conditional final subtraction: r in [0, 2p) -> [0, p)
    static B200_HD void reduce_once(E& r) {
        const uint32_t* p = C::p();
        uint32_t t[N];
        t[0] = sub_cc(r.l[0], p[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(r.l[i], p[i]);
        uint32_t borrow = subc(0, 0);  // 0 or 0xffffffff
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? r.l[i] : t[i];
    }

    static B200_HD void add(E& r, const E& a, const E& b) {
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
        // p has >= 2 spare bits in the top limb: no carry out of N limbs
        reduce_once(r);
    }
    static B200_HD void dbl(E& r, const E& a) { add(r, a, a); }

    static B200_HD void sub(E& r, const E& a, const E& b) {
        const uint32_t* p = C::p();
        r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = subc(0, 0);
        // add p back masked
        r.l[0] = add_cc(r.l[0], p[0] & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], p[i] & borrow);
        r.l[N - 1] = addc(r.l[N - 1], p[N - 1] & borrow);
    }
    static B200_HD void neg(E& r, const E& a) {
        const uint32_t* p = C::p();
        bool z = is_zero(a);
        E t;
        t.l[0] = sub_cc(p[0], a.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) t.l[i] = subc_cc(p[i], a.l[i]);
        t.l[N - 1] = subc(p[N - 1], a.l[N - 1]);
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = z ? 0u : t.l[i];
    }
    // r = a/2 mod p
    static B200_HD void halve(E& r, const E& a) {
        const uint32_t* p = C::p();
        uint32_t m = 0u - (a.l[0] & 1u);
        uint32_t t[N];
        t[0] = add_cc(a.l[0], p[0] & m);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = addc_cc(a.l[i], p[i] & m);
#pragma unroll
        for (int i = 0; i < N - 1; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
        r.l[N - 1] = t[N - 1] >> 1;
    }

    // ----------------------------------------------------------------------------------
    // Montgomery multiplication r = a*b/R mod p, fully reduced.
    // Even/odd split accumulators; see file header.  X = even-aligned words 0..N (X[N] is
    // the carry word), Y = odd-aligned words 1..N held in Y[0..N-1].
    // ----------------------------------------------------------------------------------
    // acc[0..N) += v_even * m   (v[0], v[2], ...), returns with CC = carry out of word N-1
    static B200_HD void chain_even(uint32_t* acc, const uint32_t* v, uint32_t m) {
        acc[0] = mad_lo_cc(v[0], m, acc[0]);
        acc[1] = madc_hi_cc(v[0], m, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = madc_lo_cc(v[j], m, acc[j]);
            acc[j + 1] = madc_hi_cc(v[j], m, acc[j + 1]);
        }
    }
    // acc[0..N) (odd aligned) += v_odd * m   (v[1], v[3], ...); top limb of v is < 2^30 so no carry out
    static B200_HD void chain_odd(uint32_t* acc, const uint32_t* v, uint32_t m) {
        acc[0] = mad_lo_cc(v[1], m, acc[0]);
        acc[1] = madc_hi_cc(v[1], m, acc[1]);
#pragma unroll
        for (int j = 2; j < N - 2; j += 2) {
            acc[j] = madc_lo_cc(v[j + 1], m, acc[j]);
            acc[j + 1] = madc_hi_cc(v[j + 1], m, acc[j + 1]);
        }
        acc[N - 2] = madc_lo_cc(v[N - 1], m, acc[N - 2]);
        acc[N - 1] = madc_hi(v[N - 1], m, acc[N - 1]);
    }

    // One row: A = previous even-aligned accumulator (A[0]==0, stray A[1], carry word A[N]),
    // B = previous odd-aligned accumulator.  After the 32-bit shift B is the even-aligned one
    // and A (shifted down by two words) the odd-aligned one.
    static B200_HD void row(uint32_t* A, uint32_t* B, const uint32_t* a, uint32_t bi) {
        const uint32_t* p = C::p();
        // stray word of A lands on word 0 of the new even accumulator; its carry feeds word 1
        B[0] = add_cc(B[0], A[1]);
        // A'[j] = A[j+2] + a_odd*bi   (shift fused into the multiply-accumulate)
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            A[j] = madc_lo_cc(a[j + 1], bi, A[j + 2]);
            A[j + 1] = madc_hi_cc(a[j + 1], bi, A[j + 3]);
        }
        A[N - 2] = madc_lo_cc(a[N - 1], bi, A[N]);
        A[N - 1] = madc_hi(a[N - 1], bi, 0);
        // B += a_even*bi
        chain_even(B, a, bi);
        B[N] = addc(0, 0);
        uint32_t m = mul_lo(B[0], C::inv32());
        chain_odd(A, p, m);
        chain_even(B, p, m);
        B[N] = addc(B[N], 0);
    }

    static B200_HD void mul(E& r, const E& a, const E& b) {
        const uint32_t* p = C::p();
        uint32_t X[N + 2], Y[N + 2];
        // row 0
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            X[j] = mul_lo(a.l[j], b.l[0]);
            X[j + 1] = mul_hi(a.l[j], b.l[0]);
            Y[j] = mul_lo(a.l[j + 1], b.l[0]);
            Y[j + 1] = mul_hi(a.l[j + 1], b.l[0]);
        }
        X[N] = 0; X[N + 1] = 0; Y[N] = 0; Y[N + 1] = 0;
        {
            uint32_t m = mul_lo(X[0], C::inv32());
            chain_odd(Y, p, m);
            chain_even(X, p, m);
            X[N] = addc(0, 0);
        }
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            row(X, Y, a.l, b.l[i]);
            if (i + 1 < N) row(Y, X, a.l, b.l[i + 1]);
        }
        // N rows done; N even => last row was row(X, Y): even accumulator is Y, odd is X
        // result = (Y >> 32) + X
        r.l[0] = add_cc(X[0], Y[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(X[i], Y[i + 1]);
        r.l[N - 1] = addc(X[N - 1], Y[N]);
        reduce_once(r);
    }
    // ---- dedicated Montgomery squaring -----------------------------------------------------------------------------
    // a^2 = 2 * sum_{i<j} a_i a_j B^(i+j) + sum_i a_i^2 B^(2i): N(N-1)/2 + N products for the wide square instead of N^2, then
    // the word-sliding reduction below -- 2N^2 + N -> (3N^2 + 3N)/2 multiply-accumulates (300 -> 234 for N = 12).
    // Cross products by operand scanning with the even / odd position split of the wide MAC (vm.cuh wide_mac): T is the
    // even-aligned accumulator, Od (Od[k] = word k + 1) the odd-aligned one and also takes the carries that leave the top of
    // a chain.
    static B200_HD void wide_sqr(uint32_t* T, const uint32_t* a) {
        uint32_t Od[2 * N - 1];
#pragma unroll
        for (int k = 0; k < 2 * N; k++) T[k] = 0;
#pragma unroll
        for (int k = 0; k < 2 * N - 1; k++) Od[k] = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            // even positions: j = i + 2, i + 4, ...   (words i + j, i + j + 1 of T)
            if (i + 2 < N) {
                T[2 * i + 2] = mad_lo_cc(a[i + 2], a[i], T[2 * i + 2]);
                T[2 * i + 3] = madc_hi_cc(a[i + 2], a[i], T[2 * i + 3]);
                int top = 2 * i + 3;
#pragma unroll
                for (int j = i + 4; j < N; j += 2) {
                    T[i + j] = madc_lo_cc(a[j], a[i], T[i + j]);
                    T[i + j + 1] = madc_hi_cc(a[j], a[i], T[i + j + 1]);
                    top = i + j + 1;
                }
                Od[top] = addc(Od[top], 0);                      // word top + 1
            }
            // odd positions: j = i + 1, i + 3, ...    (words i + j, i + j + 1 = Od[i + j - 1], Od[i + j])
            {
                Od[2 * i] = mad_lo_cc(a[i + 1], a[i], Od[2 * i]);
                Od[2 * i + 1] = madc_hi_cc(a[i + 1], a[i], Od[2 * i + 1]);
                int top = 2 * i + 1;
#pragma unroll
                for (int j = i + 3; j < N; j += 2) {
                    Od[i + j - 1] = madc_lo_cc(a[j], a[i], Od[i + j - 1]);
                    Od[i + j] = madc_hi_cc(a[j], a[i], Od[i + j]);
                    top = i + j;
                }
                if (top + 1 < 2 * N - 1) Od[top + 1] = addc(Od[top + 1], 0);
            }
        }
        // T += Od << 32, doubled, plus the squares on the even word pairs
        T[1] = add_cc(T[1], Od[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; k++) T[k] = addc_cc(T[k], Od[k - 1]);
        T[2 * N - 1] = addc(T[2 * N - 1], Od[2 * N - 2]);
#pragma unroll
        for (int k = 2 * N - 1; k > 0; k--) T[k] = (T[k] << 1) | (T[k - 1] >> 31);
        T[0] <<= 1;
        T[0] = mad_lo_cc(a[0], a[0], T[0]);
        T[1] = madc_hi_cc(a[0], a[0], T[1]);
#pragma unroll
        for (int i = 1; i < N; i++) {
            T[2 * i] = madc_lo_cc(a[i], a[i], T[2 * i]);
            if (i < N - 1) T[2 * i + 1] = madc_hi_cc(a[i], a[i], T[2 * i + 1]);
            else T[2 * i + 1] = madc_hi(a[i], a[i], T[2 * i + 1]);
        }
    }
    // acc (even aligned, full-size words above) += v_even * m, carry rippled one word up
    static B200_HD void redc_chain_even(uint32_t* acc, const uint32_t* v, uint32_t m) {
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
    static B200_HD void redc_chain_odd(uint32_t* acc, const uint32_t* v, uint32_t m, bool carry_in) {
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
    // Montgomery reduction of a wide value T < p R (2N words) -> canonical T / R mod p.  Word-sliding reduction on an even /
    // odd split of T (the scheme of the pairing VM's redc, vm.cuh): no data movement, the window offsets are compile-time.
    static B200_HD void redc_wide(E& r, const uint32_t* Tw) {
        uint32_t X[2 * N + 2], Y[2 * N + 2];
#pragma unroll
        for (int k = 0; k < 2 * N; k += 2) { X[k] = Tw[k]; X[k + 1] = 0; Y[k] = Tw[k + 1]; Y[k + 1] = 0; }
        X[2 * N] = 0; X[2 * N + 1] = 0; Y[2 * N] = 0; Y[2 * N + 1] = 0;
        const uint32_t* p = C::p();
        {
            uint32_t m = mul_lo(X[0], C::inv32());
            redc_chain_odd(Y, p, m, false);
            redc_chain_even(X, p, m);
        }
#pragma unroll
        for (int i = 1; i < N; i++) {
            if (i & 1) {
                const int xo = i - 1, yo = i - 1;
                Y[yo] = add_cc(Y[yo], X[xo + 1]);
                uint32_t m = mul_lo(Y[yo], C::inv32());
                redc_chain_odd(X + xo + 2, p, m, true);
                redc_chain_even(Y + yo, p, m);
            } else {
                const int yo = i - 2, xo = i;
                X[xo] = add_cc(X[xo], Y[yo + 1]);
                uint32_t m = mul_lo(X[xo], C::inv32());
                redc_chain_odd(Y + yo + 2, p, m, true);
                redc_chain_even(X + xo, p, m);
            }
        }
        const uint32_t* Yw = Y + (N - 2);
        const uint32_t* Xw = X + N;
        r.l[0] = add_cc(Xw[0], Yw[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(Xw[k], Yw[k + 1]);
        r.l[N - 1] = addc(Xw[N - 1], Yw[N]);
        reduce_once(r);
    }
    static B200_HD void sqr(E& r, const E& a) {
        uint32_t T[2 * N];
        wide_sqr(T, a.l);
        redc_wide(r, T);
    }
    // Multi-operand Montgomery product ("Montgomery dot product"):  r = (sum_{t<T} a[t]*b[t]) / R  mod p, fully reduced.
    // Same row structure as mul(): per row the T multiplicand chains accumulate into the SAME even/odd accumulators and
    // ONE reduction row follows, so T products cost T*N^2 + N^2 + N multiply-accumulates instead of T*(2N^2+N) --
    // the lazy reduction of the pairing VM without any wide (2N-word) arithmetic.  Requires T <= 3 (top-limb headroom:
    // (T+1) * 2^30 <= 2^32 in the odd accumulator) and canonical inputs; the unreduced result is < (T*p/R + 1) p < 2p.
    template <int T>
    static B200_HD void mul_dot(E& r, const E* a, const E* b) {
        static_assert(T >= 1 && T <= 3, "mul_dot: 1..3 operands");
        const uint32_t* p = C::p();
        uint32_t X[N + 2], Y[N + 2];
        // row 0
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            X[j] = mul_lo(a[0].l[j], b[0].l[0]);
            X[j + 1] = mul_hi(a[0].l[j], b[0].l[0]);
            Y[j] = mul_lo(a[0].l[j + 1], b[0].l[0]);
            Y[j + 1] = mul_hi(a[0].l[j + 1], b[0].l[0]);
        }
        X[N] = 0; X[N + 1] = 0; Y[N] = 0; Y[N + 1] = 0;
#pragma unroll
        for (int t = 1; t < T; t++) {
            chain_odd(Y, a[t].l, b[t].l[0]);
            chain_even(X, a[t].l, b[t].l[0]);
            X[N] = addc(X[N], 0);
        }
        {
            uint32_t m = mul_lo(X[0], C::inv32());
            chain_odd(Y, p, m);
            chain_even(X, p, m);
            X[N] = addc(X[N], 0);
        }
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            row_dot<T>(X, Y, a, b, i);
            if (i + 1 < N) row_dot<T>(Y, X, a, b, i + 1);
        }
        r.l[0] = add_cc(X[0], Y[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(X[i], Y[i + 1]);
        r.l[N - 1] = addc(X[N - 1], Y[N]);
        reduce_once(r);
    }
    // register-operand front ends (arrays of operands get demoted to local memory by the compiler)
    static B200_HD void mul_dot2(E& r, const E& a0, const E& b0, const E& a1, const E& b1) {
        const E a[2] = {a0, a1}, b[2] = {b0, b1};
        mul_dot<2>(r, a, b);
    }
    static B200_HD void mul_dot3(E& r, const E& a0, const E& b0, const E& a1, const E& b1, const E& a2, const E& b2) {
        const E a[3] = {a0, a1, a2}, b[3] = {b0, b1, b2};
        mul_dot<3>(r, a, b);
    }
    // one row of mul_dot: A = previous even accumulator (A[0] == 0, stray A[1], carry word A[N]), B = previous odd one
    template <int T>
    static B200_HD void row_dot(uint32_t* A, uint32_t* B, const E* a, const E* b, int i) {
        const uint32_t* p = C::p();
        const uint32_t b0 = b[0].l[i];
        B[0] = add_cc(B[0], A[1]);
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            A[j] = madc_lo_cc(a[0].l[j + 1], b0, A[j + 2]);
            A[j + 1] = madc_hi_cc(a[0].l[j + 1], b0, A[j + 3]);
        }
        A[N - 2] = madc_lo_cc(a[0].l[N - 1], b0, A[N]);
        A[N - 1] = madc_hi(a[0].l[N - 1], b0, 0);
        chain_even(B, a[0].l, b0);
        B[N] = addc(0, 0);
#pragma unroll
        for (int t = 1; t < T; t++) {
            const uint32_t bt = b[t].l[i];
            chain_odd(A, a[t].l, bt);
            chain_even(B, a[t].l, bt);
            B[N] = addc(B[N], 0);
        }
        uint32_t m = mul_lo(B[0], C::inv32());
        chain_odd(A, p, m);
        chain_even(B, p, m);
        B[N] = addc(B[N], 0);
    }

    // out-of-line product for callers whose hot loop must stay inside the instruction cache (G1 point formulas).
    // Operands and result travel BY VALUE: the device ABI passes them in registers (24 words in, 12 out), whereas
    // references would force every operand through a local-memory stack slot -- with ~200 KB of shared memory per SM the
    // L1 is nearly gone and that traffic went to L2 / DRAM (round 1: 39 GB written by one 2^21-point g1_mul launch).
    static B200_HD_NOINLINE E mulv(E a, E b) { E r; mul(r, a, b); return r; }
    static B200_HD void mulx(E& r, const E& a, const E& b) { r = mulv(a, b); }
    static B200_HD_NOINLINE E sqrv(E a) { E r; sqr(r, a); return r; }
    static B200_HD void sqrx(E& r, const E& a) { r = sqrv(a); }
    // a0 b0 + a1 b1 with ONE Montgomery reduction (mul_dot<2>: 3N^2 + N multiply-accumulates instead of 4N^2 + 2N), out of
    // line and by value like mulv
    static B200_HD_NOINLINE E mulv2(E a0, E b0, E a1, E b1) { E r; mul_dot2(r, a0, b0, a1, b1); return r; }
    static B200_HD void mulx2(E& r, const E& a0, const E& b0, const E& a1, const E& b1) { r = mulv2(a0, b0, a1, b1); }

    // Montgomery form conversions
    static B200_HD void to_mont(E& r, const E& a) {
        E r2;
        const uint32_t* q = C::r2();
#pragma unroll
        for (int i = 0; i < N; i++) r2.l[i] = q[i];
        mul(r, a, r2);
    }
    static B200_HD void from_mont(E& r, const E& a) {
        E o;
        zero(o);
        o.l[0] = 1;
        mul(r, a, o);
    }

    // r = a^(p-2) (Fermat).  Kept as the cross-check of inv() in the tests.
    static B200_HD_NOINLINE void inv_fermat(E& r, const E& a) {
        const uint32_t* p = C::p();
        E acc, base = a;
        one(acc);
        uint32_t borrow2 = 2;
        for (int i = 0; i < N; i++) {
            uint32_t w = p[i];
            uint32_t e = w - borrow2;
            borrow2 = (w < borrow2) ? 1 : 0;
            for (int bit = 0; bit < 32; bit++) {
                if ((e >> bit) & 1) mul(acc, acc, base);
                sqr(base, base);
            }
        }
        r = acc;
    }

    // helpers on plain N-word integers
    static B200_HD bool geq(const uint32_t* a, const uint32_t* b) {      // a >= b
        uint32_t t = sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t = subc_cc(a[i], b[i]);
        (void)t;
        return subc(0, 0) == 0;
    }
    static B200_HD void shr1(uint32_t* a) {
#pragma unroll
        for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
        a[N - 1] >>= 1;
    }
    static B200_HD void sub_n(uint32_t* a, const uint32_t* b) {           // a -= b (a >= b)
        a[0] = sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) a[i] = subc_cc(a[i], b[i]);
        a[N - 1] = subc(a[N - 1], b[N - 1]);
    }

    // r = a^-1 in Montgomery form (a = xR -> x^-1 R), binary extended Euclid on the ALU pipe:
    // ~2*bits shift / subtract steps instead of ~1.5*bits Montgomery products, and it leaves the multiplier free.
    // inv(0) = 0 (same convention as gnark's Inverse).  Not constant time; inputs are public here.
    static B200_HD_NOINLINE void inv(E& r, const E& a) {
        if (is_zero(a)) { zero(r); return; }
        const uint32_t* p = C::p();
        uint32_t u[N], v[N];
        E x1, x2;
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = a.l[i]; v[i] = p[i]; x1.l[i] = 0; x2.l[i] = 0; }
        x1.l[0] = 1;
        // invariants: x1 * a == u, x2 * a == v (mod p)
        for (;;) {
            uint32_t u_is_one = u[0] ^ 1u, v_is_one = v[0] ^ 1u;
#pragma unroll
            for (int i = 1; i < N; i++) { u_is_one |= u[i]; v_is_one |= v[i]; }
            if (u_is_one == 0 || v_is_one == 0) {
                E res = (u_is_one == 0) ? x1 : x2;
                // res = a^-1 (as integers mod p) = x^-1 R^-1 ; times R^3 / R -> x^-1 R
                E r3;
                const uint32_t* q = C::K().r3;
#pragma unroll
                for (int i = 0; i < N; i++) r3.l[i] = q[i];
                mul(r, res, r3);
                return;
            }
            while (!(u[0] & 1)) { shr1(u); halve(x1, x1); }
            while (!(v[0] & 1)) { shr1(v); halve(x2, x2); }
            if (geq(u, v)) { sub_n(u, v); sub(x1, x1, x2); }
            else { sub_n(v, u); sub(x2, x2, x1); }
        }
    }
};

}  // namespace b200
