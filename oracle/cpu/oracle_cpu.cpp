// CPU oracle for the mathlib pairing / G1 hot path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// A C++ restatement, on 64-bit limbs, of the algorithms the reference's CPU drivers run for this path.
// The reference forwards to un-vendored Go modules (gnark-crypto v0.20.1, kilic/bls12-381 v0.1.0; go.mod:6,14)
// that cannot be built here (no Go toolchain), so this file restates their published algorithms:
//   * Fp: CIOS Montgomery multiplication on 64-bit limbs with a 128-bit accumulator -- the routine the reference
//     itself carries at driver/kilic/custom_generic.go:57-175 (same p, same -p^-1 = 0x89f3fffcfffcfffd for BLS12-381).
//   * tower Fp2/Fp6/Fp12 with Karatsuba, sparse line multiplication, cyclotomic squaring
//   * optimal-ate Miller loop in homogeneous projective coordinates with shared squaring for multi-pairings
//     (gnark MillerLoop: call sites driver/gurvy/bn254.go:248,257; bls12-377.go:245,254; bls12381/bls12-381.go:449,458)
//   * final exponentiation: easy part + Hayashida-Hayasaka-Teruya (BLS12) / Fuentes-Castaneda (BN254) hard part
//     (gnark FinalExponentiation: bn254.go:266; bls12-377.go:263; bls12-381.go:467)
//   * G1 Jacobian arithmetic, windowed scalar multiplication (ScalarMultiplication: bn254.go:51), Strauss-Shamir Mul2
//     (bls12-381.go:869-937), Pippenger bucket MSM with signed digits, one window per worker thread
//     (gnark MultiExp: bn254.go:242; bls12-381.go:777 with ecc.MultiExpConfig{} = all cores)
//   * the Bytes() codecs of SURVEY A.3
// Every derived constant (R, R^2, -p^-1, Frobenius coefficients, twist b') is computed at start-up from the
// modulus and the tower definition, independently of the product's generated constants.h.
// It is labelled everywhere as "restated CPU baseline (not gnark/kilic assembly)".
// Pinned by tests/test_oracle_cpu.py against the Python oracle's golden vectors.
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>
#include <mutex>

typedef uint64_t u64;
typedef unsigned __int128 u128;

namespace {

// ------------------------------------------------------------------------------------------ parameters
struct BN254P {
    static constexpr int L = 4; static constexpr int FB = 32; static constexpr int BETA = -1;
    static constexpr int XI0 = 9, XI1 = 1; static constexpr bool TWIST_M = false; static constexpr bool BN = true;
    static constexpr u64 X = 4965661367192848881ull; static constexpr bool XNEG = false; static constexpr int B = 3;
    static constexpr int FLAGBITS = 2;
    static const char* p_hex() { return "30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47"; }
    static const char* r_hex() { return "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001"; }
};
struct BLS381P {
    static constexpr int L = 6; static constexpr int FB = 48; static constexpr int BETA = -1;
    static constexpr int XI0 = 1, XI1 = 1; static constexpr bool TWIST_M = true; static constexpr bool BN = false;
    static constexpr u64 X = 0xd201000000010000ull; static constexpr bool XNEG = true; static constexpr int B = 4;
    static constexpr int FLAGBITS = 3;
    static const char* p_hex() { return "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab"; }
    static const char* r_hex() { return "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001"; }
};
struct BLS377P {
    static constexpr int L = 6; static constexpr int FB = 48; static constexpr int BETA = -5;
    static constexpr int XI0 = 0, XI1 = 1; static constexpr bool TWIST_M = false; static constexpr bool BN = false;
    static constexpr u64 X = 0x8508c00000000001ull; static constexpr bool XNEG = false; static constexpr int B = 1;
    static constexpr int FLAGBITS = 3;
    static const char* p_hex() { return "01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001"; }
    static const char* r_hex() { return "12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001"; }
};

// number of Fp Montgomery products executed by the calling thread (algorithmic work unit `m`)
static thread_local u64 t_mul_count = 0;

static void hex_to_limbs(const char* h, u64* out, int L) {
    for (int i = 0; i < L; i++) out[i] = 0;
    int n = (int)strlen(h);
    for (int i = 0; i < n; i++) {
        char c = h[n - 1 - i];
        u64 v = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
        out[i / 16] |= v << (4 * (i % 16));
    }
}

// ------------------------------------------------------------------------------------------ Fp
template <class C>
struct Fp {
    static constexpr int L = C::L;
    u64 v[L];

    static u64 P[L], INV, ONE[L], R2[L];
    static bool ready;

    static bool geq_p(const u64* a) {
        for (int i = L - 1; i >= 0; i--) { if (a[i] != P[i]) return a[i] > P[i]; }
        return true;
    }
    static void sub_p(u64* a) {
        u64 br = 0;
        for (int i = 0; i < L; i++) { u128 t = (u128)a[i] - P[i] - br; a[i] = (u64)t; br = (u64)(t >> 64) & 1; }
    }
    static void init() {
        if (ready) return;
        hex_to_limbs(C::p_hex(), P, L);
        u64 x = 1;                                   // Newton: x = p^-1 mod 2^64
        for (int i = 0; i < 6; i++) x *= 2 - P[0] * x;
        INV = (u64)0 - x;
        u64 t[L]; memset(t, 0, sizeof t); t[0] = 1;   // 2^k mod p by doubling
        auto dbl = [&](u64* a) {
            u64 c = 0;
            for (int i = 0; i < L; i++) { u64 n = (a[i] << 1) | c; c = a[i] >> 63; a[i] = n; }
            if (c || geq_p(a)) sub_p(a);
        };
        for (int i = 0; i < 64 * L; i++) dbl(t);
        memcpy(ONE, t, sizeof t);
        for (int i = 0; i < 64 * L; i++) dbl(t);
        memcpy(R2, t, sizeof t);
        ready = true;
    }
    static Fp zero() { Fp r; memset(r.v, 0, sizeof r.v); return r; }
    static Fp one() { Fp r; memcpy(r.v, ONE, sizeof r.v); return r; }
    static Fp from_int(long k) {
        Fp r = zero();
        r.v[0] = (u64)(k < 0 ? -k : k);
        r = r.to_mont();
        return k < 0 ? r.neg() : r;
    }
    bool is_zero() const { u64 t = 0; for (int i = 0; i < L; i++) t |= v[i]; return t == 0; }
    bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof v) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    // CIOS Montgomery product (cf. reference driver/kilic/custom_generic.go:57-175)
    Fp operator*(const Fp& o) const {
        t_mul_count++;
        u64 t[L + 2];
        memset(t, 0, sizeof t);
        for (int i = 0; i < L; i++) {
            u64 c = 0;
            for (int j = 0; j < L; j++) {
                u128 s = (u128)v[j] * o.v[i] + t[j] + c;
                t[j] = (u64)s; c = (u64)(s >> 64);
            }
            u128 s = (u128)t[L] + c;
            t[L] = (u64)s; t[L + 1] = (u64)(s >> 64);
            u64 m = t[0] * INV;
            s = (u128)m * P[0] + t[0];
            c = (u64)(s >> 64);
            for (int j = 1; j < L; j++) {
                s = (u128)m * P[j] + t[j] + c;
                t[j - 1] = (u64)s; c = (u64)(s >> 64);
            }
            s = (u128)t[L] + c;
            t[L - 1] = (u64)s;
            t[L] = t[L + 1] + (u64)(s >> 64);
        }
        Fp r;
        memcpy(r.v, t, sizeof r.v);
        if (t[L] || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    Fp sqr() const { return *this * *this; }
    Fp operator+(const Fp& o) const {
        Fp r; u64 c = 0;
        for (int i = 0; i < L; i++) { u128 s = (u128)v[i] + o.v[i] + c; r.v[i] = (u64)s; c = (u64)(s >> 64); }
        if (c || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r; u64 br = 0;
        for (int i = 0; i < L; i++) { u128 s = (u128)v[i] - o.v[i] - br; r.v[i] = (u64)s; br = (u64)(s >> 64) & 1; }
        if (br) { u64 c = 0; for (int i = 0; i < L; i++) { u128 s = (u128)r.v[i] + P[i] + c; r.v[i] = (u64)s; c = (u64)(s >> 64); } }
        return r;
    }
    Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    Fp dbl() const { return *this + *this; }
    Fp halve() const {
        Fp r = *this; u64 c = 0;
        if (r.v[0] & 1) { for (int i = 0; i < L; i++) { u128 s = (u128)r.v[i] + P[i] + c; r.v[i] = (u64)s; c = (u64)(s >> 64); } }
        for (int i = 0; i < L - 1; i++) r.v[i] = (r.v[i] >> 1) | (r.v[i + 1] << 63);
        r.v[L - 1] = (r.v[L - 1] >> 1) | (c << 63);
        return r;
    }
    Fp to_mont() const { Fp r2; memcpy(r2.v, R2, sizeof r2.v); return *this * r2; }
    Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return *this * o; }
    Fp pow_limbs(const u64* e, int n) const {
        Fp acc = one();
        for (int i = n * 64 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) acc = acc * *this;
        }
        return acc;
    }
    Fp inv() const {                                   // a^(p-2)
        u64 e[L]; memcpy(e, P, sizeof e);
        u64 br = 2;
        for (int i = 0; i < L && br; i++) { u64 o = e[i]; e[i] = o - br; br = o < br ? 1 : 0; }
        return pow_limbs(e, L);
    }
    // canonical big-endian bytes
    void to_bytes(uint8_t* out) const {
        Fp t = from_mont();
        for (int i = 0; i < L; i++) for (int b = 0; b < 8; b++) out[C::FB - 1 - (8 * i + b)] = (uint8_t)(t.v[i] >> (8 * b));
    }
    static bool from_bytes(Fp& r, const uint8_t* in, uint8_t topmask) {
        Fp t = zero();
        for (int i = 0; i < C::FB; i++) {
            uint8_t b = in[C::FB - 1 - i];
            if (i == C::FB - 1) b &= topmask;
            t.v[i / 8] |= (u64)b << (8 * (i % 8));
        }
        if (geq_p(t.v)) return false;
        r = t.to_mont();
        return true;
    }
};
template <class C> u64 Fp<C>::P[Fp<C>::L];
template <class C> u64 Fp<C>::INV;
template <class C> u64 Fp<C>::ONE[Fp<C>::L];
template <class C> u64 Fp<C>::R2[Fp<C>::L];
template <class C> bool Fp<C>::ready = false;

// ------------------------------------------------------------------------------------------ tower
template <class C>
struct Fp2 {
    typedef Fp<C> F;
    F a0, a1;
    static Fp2 zero() { return {F::zero(), F::zero()}; }
    static Fp2 one() { return {F::one(), F::zero()}; }
    bool is_zero() const { return a0.is_zero() && a1.is_zero(); }
    bool operator==(const Fp2& o) const { return a0 == o.a0 && a1 == o.a1; }
    Fp2 operator+(const Fp2& o) const { return {a0 + o.a0, a1 + o.a1}; }
    Fp2 operator-(const Fp2& o) const { return {a0 - o.a0, a1 - o.a1}; }
    Fp2 neg() const { return {a0.neg(), a1.neg()}; }
    Fp2 conj() const { return {a0, a1.neg()}; }
    Fp2 dbl() const { return {a0.dbl(), a1.dbl()}; }
    Fp2 halve() const { return {a0.halve(), a1.halve()}; }
    Fp2 triple() const { return dbl() + *this; }
    static F mul_beta(const F& x) {
        if (C::BETA == -1) return x.neg();
        F t = x.dbl().dbl() + x;
        return t.neg();
    }
    Fp2 operator*(const Fp2& o) const {
        F t0 = a0 * o.a0, t1 = a1 * o.a1;
        F t2 = (a0 + a1) * (o.a0 + o.a1);
        return {t0 + mul_beta(t1), t2 - t0 - t1};
    }
    Fp2 sqr() const {                                  // complex squaring: 2 products
        F v = a0 * a1;
        if (C::BETA == -1) return {(a0 + a1) * (a0 - a1), v.dbl()};
        // (a0+a1)(a0+BETA a1) - (1+BETA) v   with BETA = -5
        F t = (a0 + a1) * (a0 + mul_beta(a1));
        return {t + v.dbl().dbl(), v.dbl()};
    }
    Fp2 mul_fp(const F& s) const { return {a0 * s, a1 * s}; }
    Fp2 mul_xi() const {
        if (C::XI0 == 1) return {a0 - a1, a0 + a1};
        if (C::XI0 == 9) {
            F n0 = a0.dbl().dbl().dbl() + a0, n1 = a1.dbl().dbl().dbl() + a1;
            return {n0 - a1, n1 + a0};
        }
        return {mul_beta(a1), a0};
    }
    Fp2 inv() const {
        F n = a0.sqr() - mul_beta(a1.sqr());
        F i = n.inv();
        return {a0 * i, (a1 * i).neg()};
    }
    Fp2 pow_u64s(const u64* e, int n) const {
        Fp2 acc = one();
        for (int i = n * 64 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) acc = acc * *this;
        }
        return acc;
    }
};

template <class C>
struct Fp6 {
    typedef Fp2<C> F2;
    F2 b0, b1, b2;
    static Fp6 zero() { return {F2::zero(), F2::zero(), F2::zero()}; }
    static Fp6 one() { return {F2::one(), F2::zero(), F2::zero()}; }
    bool operator==(const Fp6& o) const { return b0 == o.b0 && b1 == o.b1 && b2 == o.b2; }
    Fp6 operator+(const Fp6& o) const { return {b0 + o.b0, b1 + o.b1, b2 + o.b2}; }
    Fp6 operator-(const Fp6& o) const { return {b0 - o.b0, b1 - o.b1, b2 - o.b2}; }
    Fp6 neg() const { return {b0.neg(), b1.neg(), b2.neg()}; }
    Fp6 dbl() const { return {b0.dbl(), b1.dbl(), b2.dbl()}; }
    Fp6 mul_v() const { return {b2.mul_xi(), b0, b1}; }
    Fp6 operator*(const Fp6& o) const {
        F2 t0 = b0 * o.b0, t1 = b1 * o.b1, t2 = b2 * o.b2;
        F2 c0 = ((b1 + b2) * (o.b1 + o.b2) - t1 - t2).mul_xi() + t0;
        F2 c1 = (b0 + b1) * (o.b0 + o.b1) - t0 - t1 + t2.mul_xi();
        F2 c2 = (b0 + b2) * (o.b0 + o.b2) - t0 - t2 + t1;
        return {c0, c1, c2};
    }
    Fp6 mul_by_01(const F2& c0, const F2& c1) const {
        F2 t0 = b0 * c0, t1 = b1 * c1;
        F2 r0 = (b2 * c1).mul_xi() + t0;
        F2 r1 = (b0 + b1) * (c0 + c1) - t0 - t1;
        F2 r2 = b2 * c0 + t1;
        return {r0, r1, r2};
    }
    Fp6 mul_by_1(const F2& c1) const { return {(b2 * c1).mul_xi(), b0 * c1, b1 * c1}; }
    Fp6 mul_by_0(const F2& c0) const { return {b0 * c0, b1 * c0, b2 * c0}; }
    Fp6 inv() const {
        F2 t0 = b0.sqr() - (b1 * b2).mul_xi();
        F2 t1 = b2.sqr().mul_xi() - b0 * b1;
        F2 t2 = b1.sqr() - b0 * b2;
        F2 d = b0 * t0 + (b2 * t1 + b1 * t2).mul_xi();
        F2 di = d.inv();
        return {t0 * di, t1 * di, t2 * di};
    }
};

template <class C>
struct Fp12 {
    typedef Fp2<C> F2;
    typedef Fp6<C> F6;
    F6 c0, c1;
    static F2 FROB[4][6];     // FROB[k][i] = xi^(i(p^k-1)/6)
    static F2 BTW;            // twist coefficient
    static bool ready;

    static Fp12 one() { return {F6::one(), F6::zero()}; }
    bool operator==(const Fp12& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fp12 operator*(const Fp12& o) const {
        F6 t0 = c0 * o.c0, t1 = c1 * o.c1;
        return {t0 + t1.mul_v(), (c0 + c1) * (o.c0 + o.c1) - t0 - t1};
    }
    Fp12 sqr() const {
        F6 t = c0 * c1;
        return {(c0 + c1) * (c0 + c1.mul_v()) - t - t.mul_v(), t.dbl()};
    }
    Fp12 conj() const { return {c0, c1.neg()}; }
    Fp12 inv() const {
        F6 d = (c0 * c0 - (c1 * c1).mul_v()).inv();
        return {c0 * d, (c1 * d).neg()};
    }
    Fp12 mul_by_014(const F2& l0, const F2& l1, const F2& l4) const {
        F6 a = c0.mul_by_01(l0, l1), b = c1.mul_by_1(l4);
        F6 e = (c0 + c1).mul_by_01(l0, l1 + l4);
        return {a + b.mul_v(), e - a - b};
    }
    Fp12 mul_by_034(const F2& l0, const F2& l3, const F2& l4) const {
        F6 a = c0.mul_by_0(l0), b = c1.mul_by_01(l3, l4);
        F6 e = (c0 + c1).mul_by_01(l0 + l3, l4);
        return {a + b.mul_v(), e - a - b};
    }
    Fp12 frob(int k) const {
        const F2* in[6] = {&c0.b0, &c1.b0, &c0.b1, &c1.b1, &c0.b2, &c1.b2};
        F2 o[6];
        for (int i = 0; i < 6; i++) {
            F2 t = (k & 1) ? in[i]->conj() : *in[i];
            o[i] = i ? t * FROB[k][i] : t;
        }
        return {{o[0], o[2], o[4]}, {o[1], o[3], o[5]}};
    }
    static void fp4_sqr(F2& r0, F2& r1, const F2& a, const F2& b) {
        F2 t0 = a.sqr(), t1 = b.sqr();
        r1 = (a + b).sqr() - t0 - t1;
        r0 = t0 + t1.mul_xi();
    }
    Fp12 cyclo_sqr() const {                     // Granger-Scott
        F2 a0, a1, b0, b1, d0, d1;
        fp4_sqr(a0, a1, c0.b0, c1.b1);
        fp4_sqr(b0, b1, c1.b0, c0.b2);
        fp4_sqr(d0, d1, c0.b1, c1.b2);
        F2 x = d1.mul_xi();
        Fp12 r;
        r.c0.b0 = (a0 - c0.b0).dbl() + a0;
        r.c1.b1 = (a1 + c1.b1).dbl() + a1;
        r.c1.b0 = (x + c1.b0).dbl() + x;
        r.c0.b2 = (d0 - c0.b2).dbl() + d0;
        r.c0.b1 = (b0 - c0.b1).dbl() + b0;
        r.c1.b2 = (b1 + c1.b2).dbl() + b1;
        return r;
    }
    Fp12 exp_x() const {                          // this^x (x signed)
        Fp12 acc = *this;
        int top = 63;
        while (!((C::X >> top) & 1)) top--;
        for (int i = top - 1; i >= 0; i--) {
            acc = acc.cyclo_sqr();
            if ((C::X >> i) & 1) acc = acc * *this;
        }
        return C::XNEG ? acc.conj() : acc;
    }
    static void init() {
        if (ready) return;
        Fp<C>::init();
        typedef Fp<C> F;
        F2 xi = {F::from_int(C::XI0), F::from_int(C::XI1)};
        // (p-1)/6 by long division
        u64 e[C::L];
        memcpy(e, F::P, sizeof e);
        e[0] -= 1;
        u64 rem = 0;
        for (int i = C::L - 1; i >= 0; i--) { u128 cur = ((u128)rem << 64) | e[i]; e[i] = (u64)(cur / 6); rem = (u64)(cur % 6); }
        F2 g1 = xi.pow_u64s(e, C::L);
        F2 g2 = g1 * g1.conj();                   // xi^((p^2-1)/6) = g1^(p+1)
        F2 g3 = g1 * g2;                          // xi^((p^3-1)/6) = g1^(p^2+p+1) = g1 * g2
        F2 base[4] = {F2::one(), g1, g2, g3};
        for (int k = 1; k <= 3; k++) {
            FROB[k][0] = F2::one();
            for (int i = 1; i < 6; i++) FROB[k][i] = FROB[k][i - 1] * base[k];
        }
        F2 b = {F::from_int(C::B), F::zero()};
        BTW = C::TWIST_M ? b * xi : b * xi.inv();
        ready = true;
    }
    // Gt bytes: 12 Fp, highest coefficient first at every level (reverse of memory order)
    void to_bytes(uint8_t* out) const {
        const Fp<C>* e = reinterpret_cast<const Fp<C>*>(this);
        for (int k = 0; k < 12; k++) e[11 - k].to_bytes(out + k * C::FB);
    }
    static bool from_bytes(Fp12& r, const uint8_t* in) {
        Fp<C>* e = reinterpret_cast<Fp<C>*>(&r);
        bool ok = true;
        for (int k = 0; k < 12; k++) ok &= Fp<C>::from_bytes(e[11 - k], in + k * C::FB, 0xFF);
        return ok;
    }
};
template <class C> Fp2<C> Fp12<C>::FROB[4][6];
template <class C> Fp2<C> Fp12<C>::BTW;
template <class C> bool Fp12<C>::ready = false;

// ------------------------------------------------------------------------------------------ G1 / G2 points
template <class C> struct G1A { Fp<C> x, y; bool inf() const { return x.is_zero() && y.is_zero(); } };
template <class C> struct G2A { Fp2<C> x, y; bool inf() const { return x.is_zero() && y.is_zero(); } };

template <class C>
struct Codec {
    typedef Fp<C> F;
    static constexpr uint8_t MASK = C::FLAGBITS == 3 ? 0xE0 : 0xC0;
    static bool g1_load(G1A<C>& p, const uint8_t* in) {
        uint8_t fl = in[0] & MASK;
        if (C::FLAGBITS == 3 && fl == 0x40) { p.x = F::zero(); p.y = F::zero(); return true; }
        if (fl) return false;
        return F::from_bytes(p.x, in, (uint8_t)~MASK) & F::from_bytes(p.y, in + C::FB, 0xFF);
    }
    static void g1_store(uint8_t* out, const G1A<C>& p) {
        p.x.to_bytes(out); p.y.to_bytes(out + C::FB);
        if (C::FLAGBITS == 3 && p.inf()) out[0] |= 0x40;
    }
    static bool g2_load(G2A<C>& q, const uint8_t* in) {
        uint8_t fl = in[0] & MASK;
        if (C::FLAGBITS == 3 && fl == 0x40) { q.x = Fp2<C>::zero(); q.y = Fp2<C>::zero(); return true; }
        if (fl) return false;
        return F::from_bytes(q.x.a1, in, (uint8_t)~MASK) & F::from_bytes(q.x.a0, in + C::FB, 0xFF) &
               F::from_bytes(q.y.a1, in + 2 * C::FB, 0xFF) & F::from_bytes(q.y.a0, in + 3 * C::FB, 0xFF);
    }
};

// ------------------------------------------------------------------------------------------ pairing
template <class C>
struct Pairing {
    typedef Fp<C> F; typedef Fp2<C> F2; typedef Fp12<C> F12;
    struct Proj { F2 x, y, z; };
    struct Line { F2 r0, r1, r2; };

    static void double_step(Proj& t, Line& l) {
        F2 A = (t.x * t.y).halve(), B = t.y.sqr(), Cc = t.z.sqr();
        F2 E = Cc.triple() * F12::BTW, Fv = E.triple();
        F2 G = (B + Fv).halve();
        F2 H = (t.y + t.z).sqr() - (B + Cc);
        F2 I = E - B, J = t.x.sqr(), K = E.sqr().triple();
        t.x = (B - Fv) * A;
        t.y = G.sqr() - K;
        t.z = B * H;
        if (C::TWIST_M) l = {I, J.triple(), H.neg()}; else l = {H.neg(), J.triple(), I};
    }
    static void add_step(Proj& t, Line& l, const G2A<C>& q, bool update) {
        F2 O = t.y - q.y * t.z, Lv = t.x - q.x * t.z;
        F2 J = q.x * O - Lv * q.y;
        if (update) {
            F2 Cc = O.sqr(), D = Lv.sqr(), E = Lv * D, Fv = t.z * Cc, G = t.x * D;
            F2 H = E + Fv - G.dbl();
            F2 ny = (G - H) * O - t.y * E;
            t.x = Lv * H; t.y = ny; t.z = E * t.z;
        }
        if (C::TWIST_M) l = {J, O.neg(), Lv}; else l = {Lv, O.neg(), J};
    }
    static F12 mul_line(const F12& f, const Line& l, const G1A<C>& p) {
        if (C::TWIST_M) return f.mul_by_014(l.r0, l.r1.mul_fp(p.x), l.r2.mul_fp(p.y));
        return f.mul_by_034(l.r0.mul_fp(p.y), l.r1.mul_fp(p.x), l.r2);
    }
    static void loop_digits(std::vector<int>& d) {
        d.clear();
        if (!C::BN) { for (int i = 0; i < 64; i++) d.push_back((int)((C::X >> i) & 1)); return; }
        u128 n = (u128)6 * C::X + 2;                       // NAF(6x+2)
        while (n) {
            int di = 0;
            if (n & 1) { di = 2 - (int)(n & 3); n -= di; }
            d.push_back(di);
            n >>= 1;
        }
    }
    static F12 miller(int np, const G1A<C>* P, const G2A<C>* Q) {
        static std::vector<int> digits;
        static std::once_flag once;
        std::call_once(once, [] { loop_digits(digits); });
        Proj t[2]; G2A<C> nq[2]; bool live[2];
        for (int k = 0; k < np; k++) {
            live[k] = !(P[k].inf() || Q[k].inf());
            t[k] = {Q[k].x, Q[k].y, F2::one()};
            nq[k] = {Q[k].x, Q[k].y.neg()};
        }
        F12 f = F12::one();
        Line l;
        int top = (int)digits.size() - 1;
        while (digits[top] == 0) top--;
        for (int i = top - 1; i >= 0; i--) {
            f = f.sqr();
            for (int k = 0; k < np; k++) {
                if (!live[k]) continue;
                double_step(t[k], l);
                f = mul_line(f, l, P[k]);
                if (digits[i]) {
                    add_step(t[k], l, digits[i] > 0 ? Q[k] : nq[k], true);
                    f = mul_line(f, l, P[k]);
                }
            }
        }
        if (C::BN) {
            for (int k = 0; k < np; k++) {
                if (!live[k]) continue;
                G2A<C> q1 = {Q[k].x.conj() * F12::FROB[1][2], Q[k].y.conj() * F12::FROB[1][3]};
                G2A<C> q2 = {Q[k].x * F12::FROB[2][2], Q[k].y};
                add_step(t[k], l, q1, true);
                f = mul_line(f, l, P[k]);
                add_step(t[k], l, q2, false);
                f = mul_line(f, l, P[k]);
            }
        }
        return C::XNEG ? f.conj() : f;
    }
    static F12 final_exp(const F12& in) {
        F12 t0 = in.conj() * in.inv();
        F12 f = t0.frob(2) * t0;
        if (!C::BN) {
            F12 a = f.cyclo_sqr();
            F12 t1 = f.exp_x() * f.conj();
            F12 t2 = t1.exp_x();
            t1 = t1.conj() * t2;
            t2 = t1.exp_x();
            t1 = t1.frob(1) * t2;
            f = f * a;
            a = t1.exp_x();
            t2 = a.exp_x();
            a = t1.frob(2);
            t1 = t1.conj() * t2 * a;
            return f * t1;
        }
        F12 t[5];
        t[0] = f.exp_x().conj().cyclo_sqr();
        t[1] = t[0].cyclo_sqr();
        t[1] = t[0] * t[1];
        t[2] = t[1].exp_x().conj();
        t[3] = t[1].conj();
        t[1] = t[2] * t[3];
        t[3] = t[2].cyclo_sqr();
        t[4] = t[3].exp_x();
        t[4] = t[1] * t[4];
        t[3] = t[0] * t[4];
        t[0] = t[2] * t[4];
        t[0] = f * t[0];
        t[2] = t[3].frob(1);
        t[0] = t[2] * t[0];
        t[2] = t[4].frob(2);
        t[0] = t[2] * t[0];
        t[2] = f.conj() * t[3];
        t[2] = t[2].frob(3);
        return t[2] * t[0];
    }
};

// ------------------------------------------------------------------------------------------ G1 Jacobian
template <class C>
struct G1J {
    typedef Fp<C> F;
    F x, y, z;
    static G1J inf() { return {F::one(), F::one(), F::zero()}; }
    bool is_inf() const { return z.is_zero(); }
    static G1J from_affine(const G1A<C>& a) { return a.inf() ? inf() : G1J{a.x, a.y, F::one()}; }
    G1J neg() const { return {x, y.neg(), z}; }
    G1J dbl() const {                               // dbl-2009-l (a = 0)
        if (is_inf()) return *this;
        F A = x.sqr(), B = y.sqr(), Cc = B.sqr();
        F D = ((x + B).sqr() - A - Cc).dbl();
        F E = A.dbl() + A, Fv = E.sqr();
        F x3 = Fv - D.dbl();
        F y3 = E * (D - x3) - Cc.dbl().dbl().dbl();
        F z3 = (y * z).dbl();
        return {x3, y3, z3};
    }
    G1J add(const G1J& o) const {                   // add-2007-bl, complete
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        F z1z1 = z.sqr(), z2z2 = o.z.sqr();
        F u1 = x * z2z2, u2 = o.x * z1z1;
        F s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
        if (u1 == u2) return s1 == s2 ? dbl() : inf();
        F H = u2 - u1, I = H.dbl().sqr(), J = H * I, r = (s2 - s1).dbl(), V = u1 * I;
        F x3 = r.sqr() - J - V.dbl();
        F y3 = r * (V - x3) - (s1 * J).dbl();
        F z3 = ((z + o.z).sqr() - z1z1 - z2z2) * H;
        return {x3, y3, z3};
    }
    G1J add_affine(const G1A<C>& a) const { return add(from_affine(a)); }
    G1A<C> to_affine() const {
        if (is_inf()) return {F::zero(), F::zero()};
        F zi = z.inv(), zi2 = zi.sqr();
        return {x * zi2, y * zi2 * zi};
    }
};

static inline int sc_bit(const u64* k, int i) { return (int)((k[i / 64] >> (i % 64)) & 1); }
static void sc_load(u64* k, const uint8_t* be32) {
    for (int i = 0; i < 4; i++) { k[i] = 0; for (int b = 0; b < 8; b++) k[i] |= (u64)be32[31 - (8 * i + b)] << (8 * b); }
}
template <class C> static void sc_reduce(u64* k) {
    u64 r[4]; hex_to_limbs(C::r_hex(), r, 4);
    for (;;) {
        u64 t[4]; u64 br = 0;
        for (int i = 0; i < 4; i++) { u128 s = (u128)k[i] - r[i] - br; t[i] = (u64)s; br = (u64)(s >> 64) & 1; }
        if (br) break;
        memcpy(k, t, sizeof t);
    }
}

template <class C>
static G1J<C> scalar_mul(const G1A<C>& p, const u64* k) {     // fixed 4-bit window
    G1J<C> tab[16];
    tab[0] = G1J<C>::inf();
    tab[1] = G1J<C>::from_affine(p);
    for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? tab[i - 1].add(tab[1]) : tab[i / 2].dbl();
    G1J<C> acc = G1J<C>::inf();
    for (int w = 63; w >= 0; w--) {
        for (int j = 0; j < 4; j++) acc = acc.dbl();
        int d = (int)((k[w / 16] >> (4 * (w % 16))) & 15);
        if (d) acc = acc.add(tab[d]);
    }
    return acc;
}

// Strauss-Shamir, 2-bit joint window, 15-entry table (reference bls12-381.go:869-937)
template <class C>
static G1J<C> joint_mul(const G1A<C>& a1, const u64* k1, const G1A<C>& a2, const u64* k2) {
    G1J<C> t[15];
    t[0] = G1J<C>::from_affine(a1);
    t[3] = G1J<C>::from_affine(a2);
    t[1] = t[0].dbl(); t[2] = t[1].add(t[0]);
    t[4] = t[3].add(t[0]); t[5] = t[3].add(t[1]); t[6] = t[3].add(t[2]);
    t[7] = t[3].dbl(); t[8] = t[7].add(t[0]); t[9] = t[7].add(t[1]); t[10] = t[7].add(t[2]);
    t[11] = t[7].add(t[3]); t[12] = t[11].add(t[0]); t[13] = t[11].add(t[1]); t[14] = t[11].add(t[2]);
    G1J<C> res = G1J<C>::inf();
    for (int i = 3; i >= 0; i--)
        for (int j = 0; j < 32; j++) {
            res = res.dbl().dbl();
            int sh = 62 - 2 * j;
            int b1 = (int)((k1[i] >> sh) & 3), b2 = (int)((k2[i] >> sh) & 3);
            if (b1 | b2) res = res.add(t[(b2 << 2 | b1) - 1]);
        }
    return res;
}

// Pippenger bucket method, signed digits, one window per task (gnark MultiExp structure)
template <class C>
static G1J<C> msm(size_t n, const G1A<C>* pts, const u64* ks /* n*4 reduced */, int nthreads) {
    if (n == 0) return G1J<C>::inf();
    int c = 2;
    { size_t t = n; int lg = 0; while (t >>= 1) lg++; c = std::max(2, std::min(16, lg - 2)); }
    int W = 256 / c + 1;
    size_t B = (size_t)1 << (c - 1);
    std::vector<G1J<C>> wsum(W, G1J<C>::inf());
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<G1J<C>> buckets(B);
        for (;;) {
            int w = next.fetch_add(1);
            if (w >= W) break;
            for (auto& b : buckets) b = G1J<C>::inf();
            for (size_t i = 0; i < n; i++) {
                // signed digit w of scalar i (recomputed with its carry chain)
                const u64* k = ks + 4 * i;
                int carry = 0, d = 0;
                for (int ww = 0; ww <= w; ww++) {
                    int bit = ww * c;
                    u64 v = 0;
                    if (bit < 256) {
                        v = k[bit / 64] >> (bit % 64);
                        if (bit % 64 + c > 64 && bit / 64 + 1 < 4) v |= k[bit / 64 + 1] << (64 - bit % 64);
                        v &= ((u64)1 << c) - 1;
                    }
                    d = (int)v + carry;
                    if (d > (int)B) { d -= (1 << c); carry = 1; } else carry = 0;
                }
                if (d == 0 || pts[i].inf()) continue;
                if (d > 0) buckets[d - 1] = buckets[d - 1].add_affine(pts[i]);
                else { G1A<C> m = {pts[i].x, pts[i].y.neg()}; buckets[-d - 1] = buckets[-d - 1].add_affine(m); }
            }
            G1J<C> run = G1J<C>::inf(), acc = G1J<C>::inf();
            for (size_t b = B; b-- > 0;) { run = run.add(buckets[b]); acc = acc.add(run); }
            wsum[w] = acc;
        }
    };
    int nt = std::max(1, std::min(nthreads, W));
    std::vector<std::thread> th;
    for (int i = 1; i < nt; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    G1J<C> acc = G1J<C>::inf();
    for (int w = W - 1; w >= 0; w--) {
        for (int j = 0; j < c; j++) acc = acc.dbl();
        acc = acc.add(wsum[w]);
    }
    return acc;
}

template <class Fn> static void parallel_for(size_t n, int nthreads, Fn fn) {
    int nt = (int)std::max<size_t>(1, std::min<size_t>(nthreads, n));
    std::atomic<size_t> next(0);
    auto worker = [&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; fn(i); } };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------------ per-curve entry points
template <class C>
struct Api {
    static void init() { Fp12<C>::init(); }
    static int pairing(int np, size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                       uint8_t* out, int fexp, int unity, int nthreads) {
        init();
        std::atomic<int> bad(0);
        parallel_for(n, nthreads, [&](size_t i) {
            G1A<C> P[2]; G2A<C> Q[2];
            bool ok = Codec<C>::g1_load(P[0], g1a + i * 2 * C::FB) & Codec<C>::g2_load(Q[0], g2a + i * 4 * C::FB);
            if (np == 2) ok &= Codec<C>::g1_load(P[1], g1b + i * 2 * C::FB) & Codec<C>::g2_load(Q[1], g2b + i * 4 * C::FB);
            if (!ok) { bad = 1; return; }
            Fp12<C> f = Pairing<C>::miller(np, P, Q);
            if (fexp) f = Pairing<C>::final_exp(f);
            if (unity) out[i] = f == Fp12<C>::one() ? 1 : 0; else f.to_bytes(out + i * 12 * C::FB);
        });
        return bad ? -3 : 0;
    }
    static int fexp(size_t n, const uint8_t* in, uint8_t* out, int do_exp, int nthreads) {
        init();
        std::atomic<int> bad(0);
        parallel_for(n, nthreads, [&](size_t i) {
            Fp12<C> f;
            if (!Fp12<C>::from_bytes(f, in + i * 12 * C::FB)) { bad = 1; return; }
            if (do_exp) f = Pairing<C>::final_exp(f);
            f.to_bytes(out + i * 12 * C::FB);
        });
        return bad ? -3 : 0;
    }
    static int g1_mul(size_t n, const uint8_t* pts, const uint8_t* ks, uint8_t* out, int nthreads) {
        init();
        std::atomic<int> bad(0);
        parallel_for(n, nthreads, [&](size_t i) {
            G1A<C> p;
            if (!Codec<C>::g1_load(p, pts + i * 2 * C::FB)) { bad = 1; return; }
            u64 k[4]; sc_load(k, ks + 32 * i);
            Codec<C>::g1_store(out + i * 2 * C::FB, scalar_mul<C>(p, k).to_affine());
        });
        return bad ? -3 : 0;
    }
    static int g1_mul2(size_t n, const uint8_t* P, const uint8_t* e, const uint8_t* Q, const uint8_t* f, uint8_t* out,
                       int nthreads) {
        init();
        std::atomic<int> bad(0);
        parallel_for(n, nthreads, [&](size_t i) {
            G1A<C> p, q;
            if (!(Codec<C>::g1_load(p, P + i * 2 * C::FB) & Codec<C>::g1_load(q, Q + i * 2 * C::FB))) { bad = 1; return; }
            u64 k1[4], k2[4]; sc_load(k1, e + 32 * i); sc_load(k2, f + 32 * i);
            sc_reduce<C>(k1); sc_reduce<C>(k2);
            Codec<C>::g1_store(out + i * 2 * C::FB, joint_mul<C>(p, k1, q, k2).to_affine());
        });
        return bad ? -3 : 0;
    }
    static int g1_msm(size_t n, const uint8_t* pts, const uint8_t* ks, uint8_t* out, int nthreads) {
        init();
        std::vector<G1A<C>> P(n);
        std::vector<u64> K(4 * n);
        std::atomic<int> bad(0);
        parallel_for((n + 1023) / 1024, nthreads, [&](size_t blk) {
            for (size_t i = blk * 1024; i < std::min(n, (blk + 1) * 1024); i++) {
                if (!Codec<C>::g1_load(P[i], pts + i * 2 * C::FB)) bad = 1;
                sc_load(&K[4 * i], ks + 32 * i);
                sc_reduce<C>(&K[4 * i]);
            }
        });
        if (bad) return -3;
        Codec<C>::g1_store(out, msm<C>(n, P.data(), K.data(), nthreads).to_affine());
        return 0;
    }
    static void fp_mul_raw(const u64* a, const u64* b, u64* o) {
        init();
        Fp<C> x, y; memcpy(x.v, a, sizeof x.v); memcpy(y.v, b, sizeof y.v);
        Fp<C> z = x * y; memcpy(o, z.v, sizeof z.v);
    }
    static void consts(u64* p, u64* inv, u64* one, u64* r2) {
        init();
        memcpy(p, Fp<C>::P, sizeof(u64) * C::L); *inv = Fp<C>::INV;
        memcpy(one, Fp<C>::ONE, sizeof(u64) * C::L); memcpy(r2, Fp<C>::R2, sizeof(u64) * C::L);
    }
};

#define DISPATCH(curve, CALL)                                                   \
    switch (curve) {                                                            \
        case 1: return Api<BN254P>::CALL;                                       \
        case 3: case 5: case 6: case 7: return Api<BLS381P>::CALL;              \
        case 4: return Api<BLS377P>::CALL;                                      \
        default: return -2;                                                     \
    }
static bool is_kilic(int curve) { return curve == 3 || curve == 6; }

}  // namespace

// curve = mathlib CurveID; BYTES encodings; same driver semantics as include/b200.h
extern "C" {
// Fp multiplications executed by the calling thread since the last reset (use nthreads = 1)
void orc_mul_count_reset() { t_mul_count = 0; }
unsigned long long orc_mul_count() { return t_mul_count; }
int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
int orc_pairing_batch(int curve, int np, size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b,
                      const uint8_t* g2b, uint8_t* out, int fexp, int unity, int nthreads) {
    int fe = fexp || is_kilic(curve);
    DISPATCH(curve, pairing(np, n, g1a, g2a, g1b, g2b, out, fe, unity, nthreads));
}
int orc_fexp_batch(int curve, size_t n, const uint8_t* in, uint8_t* out, int nthreads) {
    int fe = !is_kilic(curve);
    DISPATCH(curve, fexp(n, in, out, fe, nthreads));
}
int orc_g1_mul_batch(int curve, size_t n, const uint8_t* pts, const uint8_t* ks, uint8_t* out, int nthreads) {
    DISPATCH(curve, g1_mul(n, pts, ks, out, nthreads));
}
int orc_g1_mul2_batch(int curve, size_t n, const uint8_t* P, const uint8_t* e, const uint8_t* Q, const uint8_t* f,
                      uint8_t* out, int nthreads) {
    DISPATCH(curve, g1_mul2(n, P, e, Q, f, out, nthreads));
}
int orc_g1_msm(int curve, size_t n, const uint8_t* pts, const uint8_t* ks, uint8_t* out, int nthreads) {
    DISPATCH(curve, g1_msm(n, pts, ks, out, nthreads));
}
int orc_fp_mul_raw(int curve, const u64* a, const u64* b, u64* o) {
    switch (curve) {
        case 1: Api<BN254P>::fp_mul_raw(a, b, o); return 0;
        case 3: case 5: Api<BLS381P>::fp_mul_raw(a, b, o); return 0;
        case 4: Api<BLS377P>::fp_mul_raw(a, b, o); return 0;
    }
    return -2;
}
int orc_consts(int curve, u64* p, u64* inv, u64* one, u64* r2) {
    switch (curve) {
        case 1: Api<BN254P>::consts(p, inv, one, r2); return 0;
        case 3: case 5: Api<BLS381P>::consts(p, inv, one, r2); return 0;
        case 4: Api<BLS377P>::consts(p, inv, one, r2); return 0;
    }
    return -2;
}
}
