// Curve traits: compile-time parameters + constant-memory tables for BN254, BLS12-381, BLS12-377
// (the three curves of BASELINE.json; mathlib CurveIDs 1, 3/5/6/7, 4 -- reference math.go:70-103).
#pragma once
#include "fp.cuh"
#include "constants.h"

namespace b200 {

template <int N>
struct CurveConsts {
    uint32_t p[N];
    uint32_t one[N];
    uint32_t r2[N];
    uint32_t b[N];          // G1 curve coefficient b (Montgomery)
    uint32_t b3[N];         // 3b
    uint32_t btw[2 * N];    // twist coefficient b' in Fp2 (Montgomery)
    uint32_t order[8];      // group order r, little-endian words
    uint32_t frob1[10 * N]; // gamma_{1,i} = xi^(i(p-1)/6),   i=1..5, Fp2 each
    uint32_t frob2[10 * N]; // gamma_{2,i} = xi^(i(p^2-1)/6)
    uint32_t frob3[10 * N]; // gamma_{3,i} = xi^(i(p^3-1)/6)
    uint32_t p2[31 * 2 * N]; // k * p^2, k = 0..30, 2N words each (lazy-reduction offsets, vm.cuh)
    uint32_t r3[N];          // R^3 mod p (binary inversion fix-up)
    uint32_t glv_lambda[4];  // BLS12 G1 endomorphism eigenvalue lambda = x^2 - 1 (zeros: no GLV)
    uint32_t glv_m[5];       // floor(2^256 / lambda)
    uint32_t glv_beta[N];    // cube root of unity with [lambda](X, Y) = (beta X, Y), Montgomery
    uint32_t ts_exp[N];      // Tonelli-Shanks: (q-1)/2 with p-1 = 2^s q
    uint32_t ts_z[N];        // g^q for a quadratic non-residue g (Montgomery)
    uint32_t half_p[N];      // (p-1)/2: y is "lexicographically largest" iff y > half_p
    uint32_t pk[2 * N];      // 2p, 4p
};

#define B200_DEFINE_CONSTS(NAME, NL)                                                        \
    static const CurveConsts<NL> H_##NAME = {NAME##_P, NAME##_ONE, NAME##_R2, NAME##_B,     \
                                             NAME##_B3, NAME##_BTW, NAME##_ORDER,           \
                                             NAME##_FROB1, NAME##_FROB2, NAME##_FROB3, NAME##_P2, \
                                             NAME##_R3, NAME##_GLV_LAMBDA, NAME##_GLV_M,    \
                                             NAME##_GLV_BETA, NAME##_TS_EXP, NAME##_TS_Z, NAME##_HALF_P, NAME##_PK};

B200_DEFINE_CONSTS(BN254, 8)
B200_DEFINE_CONSTS(BLS381, 12)
B200_DEFINE_CONSTS(BLS377, 12)

#if defined(__CUDACC__)
#define B200_DEFINE_DCONSTS(NAME, NL)                                                       \
    static __device__ __constant__ CurveConsts<NL> D_##NAME = {NAME##_P, NAME##_ONE, NAME##_R2,    \
                                                        NAME##_B, NAME##_B3, NAME##_BTW,    \
                                                        NAME##_ORDER, NAME##_FROB1,         \
                                                        NAME##_FROB2, NAME##_FROB3, NAME##_P2,      \
                                                        NAME##_R3, NAME##_GLV_LAMBDA,       \
                                                        NAME##_GLV_M, NAME##_GLV_BETA,      \
                                                        NAME##_TS_EXP, NAME##_TS_Z, NAME##_HALF_P, NAME##_PK};
B200_DEFINE_DCONSTS(BN254, 8)
B200_DEFINE_DCONSTS(BLS381, 12)
B200_DEFINE_DCONSTS(BLS377, 12)
#endif

#if defined(__CUDA_ARCH__)
#define B200_K(NAME) D_##NAME
#else
#define B200_K(NAME) H_##NAME
#endif

enum TwistType { TWIST_D = 0, TWIST_M = 1 };
enum Family { FAMILY_BN = 0, FAMILY_BLS12 = 1 };

struct BN254 {
    static constexpr int N = 8;
    static constexpr int FP_BYTES = 32;
    static constexpr int BETA = -1;                   // u^2
    static constexpr int XI0 = 9, XI1 = 1;            // xi = 9 + u
    static constexpr bool LAZY_MODS = false;          // VM operand modifiers stay canonical (2 spare bits, xi = 9 + u)
    static constexpr TwistType TWIST = TWIST_D;
    static constexpr Family FAMILY = FAMILY_BN;
    static constexpr uint64_t X_ABS = 4965661367192848881ull;
    static constexpr bool X_NEG = false;
    static constexpr int FLAG_BITS = 2;
    static constexpr int SCALAR_BITS = 254;
    static constexpr int GLV_BITS = 0;                // no GLV constants
    static B200_HD const CurveConsts<8>& K() { return B200_K(BN254); }
    static B200_HD const uint32_t* p() { return K().p; }
    static B200_HD const uint32_t* one() { return K().one; }
    static B200_HD const uint32_t* r2() { return K().r2; }
    static B200_HD uint32_t inv32() { return BN254_INV32; }
    static constexpr int TS_S = BN254_TS_S;
};

struct BLS381 {
    static constexpr int N = 12;
    static constexpr int FP_BYTES = 48;
    static constexpr int BETA = -1;
    static constexpr int XI0 = 1, XI1 = 1;            // xi = 1 + u
    static constexpr bool LAZY_MODS = true;           // VM operand modifiers unreduced: 8p < 2^384 (vm.cuh lazy_mods)
    static constexpr TwistType TWIST = TWIST_M;
    static constexpr Family FAMILY = FAMILY_BLS12;
    static constexpr uint64_t X_ABS = 0xd201000000010000ull;
    static constexpr bool X_NEG = true;
    static constexpr int FLAG_BITS = 3;
    static constexpr int SCALAR_BITS = 255;
    static constexpr int GLV_BITS = 128;              // bit length of lambda = x^2 - 1: both halves of an exact GLV split fit
    static B200_HD const CurveConsts<12>& K() { return B200_K(BLS381); }
    static B200_HD const uint32_t* p() { return K().p; }
    static B200_HD const uint32_t* one() { return K().one; }
    static B200_HD const uint32_t* r2() { return K().r2; }
    static B200_HD uint32_t inv32() { return BLS381_INV32; }
    static constexpr int TS_S = BLS381_TS_S;
};

struct BLS377 {
    static constexpr int N = 12;
    static constexpr int FP_BYTES = 48;
    static constexpr int BETA = -5;
    static constexpr int XI0 = 0, XI1 = 1;            // xi = u
    static constexpr bool LAZY_MODS = false;
    static constexpr TwistType TWIST = TWIST_D;
    static constexpr Family FAMILY = FAMILY_BLS12;
    static constexpr uint64_t X_ABS = 0x8508c00000000001ull;
    static constexpr bool X_NEG = false;
    static constexpr int FLAG_BITS = 3;
    static constexpr int SCALAR_BITS = 253;
    static constexpr int GLV_BITS = 127;
    static B200_HD const CurveConsts<12>& K() { return B200_K(BLS377); }
    static B200_HD const uint32_t* p() { return K().p; }
    static B200_HD const uint32_t* one() { return K().one; }
    static B200_HD const uint32_t* r2() { return K().r2; }
    static B200_HD uint32_t inv32() { return BLS377_INV32; }
    static constexpr int TS_S = BLS377_TS_S;
};

}  // namespace b200
