// Host-emulation harness (TEST INFRASTRUCTURE): compiles the device headers for the CPU with the
// PTX carry-chain primitives emulated in C, so kernel logic can be checked without a GPU.
#include <string.h>
#include "../../mathlib_b200/csrc/curves.cuh"
using namespace b200;
template <class C> static void t_mul(const uint32_t* a, const uint32_t* b, uint32_t* o, int op) {
    typedef FpOps<C> F; typename F::E x, y, z;
    for (int i = 0; i < C::N; i++) { x.l[i] = a[i]; y.l[i] = b[i]; }
    switch (op) {
        case 0: F::mul(z, x, y); break;
        case 1: F::add(z, x, y); break;
        case 2: F::sub(z, x, y); break;
        case 3: F::neg(z, x); break;
        case 4: F::halve(z, x); break;
        case 5: F::inv(z, x); break;
        case 8: F::inv_fermat(z, x); break;
        case 6: F::to_mont(z, x); break;
        case 7: F::from_mont(z, x); break;
        case 9: F::sqr(z, x); break;
    }
    for (int i = 0; i < C::N; i++) o[i] = z.l[i];
}
extern "C" void he_fp_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t* o) {
    if (curve == 0) t_mul<BN254>(a, b, o, op);
    else if (curve == 1) t_mul<BLS381>(a, b, o, op);
    else t_mul<BLS377>(a, b, o, op);
}

#include "../../mathlib_b200/csrc/pairing.cuh"
// layout of inputs: Montgomery limbs, G1 = x|y (2N words), G2 = x.c0|x.c1|y.c0|y.c1 (4N words),
// Gt = C0.B0.A0, C0.B0.A1, C0.B1.A0, ... C1.B2.A1 (12N words, struct order)
template <class C> static void t_pair(int np, const uint32_t* g1, const uint32_t* g2, uint32_t* out, int mode) {
    typedef PairingOps<C> PO;
    constexpr int N = C::N;
    G1Aff<N> P[2]; G2Aff<N> Q[2];
    memcpy(P, g1, sizeof(G1Aff<N>) * np);
    memcpy(Q, g2, sizeof(G2Aff<N>) * np);
    Fp12<N> f;
    if (np == 1) PO::template miller_loop<1>(f, P, Q); else PO::template miller_loop<2>(f, P, Q);
    if (mode) PO::final_exp(f, f);
    memcpy(out, &f, sizeof(f));
}
extern "C" void he_pairing(int curve, int np, const uint32_t* g1, const uint32_t* g2, uint32_t* out, int mode) {
    if (curve == 0) t_pair<BN254>(np, g1, g2, out, mode);
    else if (curve == 1) t_pair<BLS381>(np, g1, g2, out, mode);
    else t_pair<BLS377>(np, g1, g2, out, mode);
}
template <class C> static void t_f12(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    typedef Tower<C> T; typedef PairingOps<C> PO;
    Fp12<C::N> x, y, z;
    memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y));
    switch (op) {
        case 0: T::f12_mul(z, x, y); break;
        case 1: T::f12_sqr(z, x); break;
        case 2: T::f12_inv(z, x); break;
        case 3: T::f12_frob(z, x, 1); break;
        case 4: T::f12_frob(z, x, 2); break;
        case 5: T::f12_frob(z, x, 3); break;
        case 6: T::f12_cyclo_sqr(z, x); break;
        case 7: PO::final_exp(z, x); break;
        case 8: T::f12_conj(z, x); break;
    }
    memcpy(out, &z, sizeof(z));
}
extern "C" void he_f12_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    if (curve == 0) t_f12<BN254>(op, a, b, out);
    else if (curve == 1) t_f12<BLS381>(op, a, b, out);
    else t_f12<BLS377>(op, a, b, out);
}

#include "../../mathlib_b200/csrc/kernels.cuh"
// G1 ops through the BYTES codecs (same code path the kernels run per thread)
template <class C> static int t_g1(int op, const uint8_t* p, const uint8_t* e, const uint8_t* q, const uint8_t* f, uint8_t* out) {
    typedef Codec<C> CD; typedef G1Ops<C> G;
    int err = 0;
    typename G::Aff a, b;
    CD::g1_load(a.x, a.y, p, false, &err);
    uint32_t ke[8], kf[8];
    CD::scalar_load(ke, e);
    typename G::Pt acc;
    if (op == 0) {
        G::scalar_mul(acc, a, ke);
    } else if (op == 1) {
        CD::g1_load(b.x, b.y, q, false, &err);
        CD::scalar_load(kf, f);
        G::scalar_mul2(acc, a, ke, b, kf);
    } else {  // sum of two points
        CD::g1_load(b.x, b.y, q, false, &err);
        G::from_affine(acc, a);
        G::madd(acc, b);
    }
    G::to_affine(a, acc);
    CD::g1_store(out, a.x, a.y, false);
    return err;
}
extern "C" int he_g1_op(int curve, int op, const uint8_t* p, const uint8_t* e, const uint8_t* q, const uint8_t* f, uint8_t* out) {
    if (curve == 0) return t_g1<BN254>(op, p, e, q, f, out);
    if (curve == 1) return t_g1<BLS381>(op, p, e, q, f, out);
    return t_g1<BLS377>(op, p, e, q, f, out);
}
// pairing through the BYTES codecs
template <class C> static int t_pair_bytes(int np, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                                           uint8_t* out, int fexp) {
    typedef Codec<C> CD; typedef PairingOps<C> PO;
    constexpr int N = C::N;
    int err = 0;
    G1Aff<N> P[2]; G2Aff<N> Q[2];
    CD::g1_load(P[0].x, P[0].y, g1a, false, &err);
    CD::g2_load(Q[0], g2a, false, &err);
    if (np == 2) { CD::g1_load(P[1].x, P[1].y, g1b, false, &err); CD::g2_load(Q[1], g2b, false, &err); }
    Fp12<N> f;
    if (np == 1) PO::template miller_loop<1>(f, P, Q); else PO::template miller_loop<2>(f, P, Q);
    if (fexp) PO::final_exp(f, f);
    CD::gt_store(out, f, false);
    return err;
}
extern "C" int he_pairing_bytes(int curve, int np, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                                uint8_t* out, int fexp) {
    if (curve == 0) return t_pair_bytes<BN254>(np, g1a, g2a, g1b, g2b, out, fexp);
    if (curve == 1) return t_pair_bytes<BLS381>(np, g1a, g2a, g1b, g2b, out, fexp);
    return t_pair_bytes<BLS377>(np, g1a, g2a, g1b, g2b, out, fexp);
}

#include "../../mathlib_b200/csrc/pairing_vm.cuh"
#include <vector>
// VM pairing on the host: lanes of a phase are executed one after another (pairing_vm.cuh, host path of run());
// SPLIT = the three-lanes-per-role mode of the small-batch kernel
template <class C, bool SPLIT = false> static int t_vm_pair(int np, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b,
                                                           const uint8_t* g2b, uint8_t* out, int fexp) {
    constexpr int N = C::N;
    typedef VmTables<C> TB;
    std::vector<uint32_t> slots((size_t)TB::NSLOTS * 2 * N, 0u), kb((size_t)VM_KBANK * 2 * N), xch((size_t)VM_G * 3 * (2 * N + 1), 0u);
    for (int i = 0; i < VM_KBANK; i++) vm_fill_kbank<C>(kb.data() + (size_t)i * 2 * N, i);
    VmDriver<C, SPLIT> D;
    D.xch = xch.data();
    D.ctx.slots = slots.data();
    D.ctx.kbank = kb.data();
    D.words = TB::host_words();
    D.dir = TB::host_dir();
    D.role = 0;
    int err = 0;
    unsigned m0 = 0, m1 = 0;
    for (int r = 0; r < VM_G; r++) {
        m0 |= (unsigned)D.load_coord(r, 0, g1a, g2a, false, &err) << r;
        if (np == 2) m1 |= (unsigned)D.load_coord(r, 1, g1b, g2b, false, &err) << r;
    }
    if (err) return 1;
    bool dead0 = ((m0 & 3u) == 3u) || ((m0 & 60u) == 60u);
    bool dead1 = np == 2 ? (((m1 & 3u) == 3u) || ((m1 & 60u) == 60u)) : true;
    D.ctx.live = (dead0 ? 0u : 1u) | (dead1 ? 0u : 2u);
    uint32_t fb = np == 1 ? D.template miller<1>() : D.template miller<2>();
    if (fexp) fb = D.final_exp(fb);
    for (int r = 0; r < VM_G; r++) D.store_coeff(r, fb, out, false);
    return 0;
}
extern "C" int he_vm_pairing(int curve, int np, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                             uint8_t* out, int fexp) {
    if (curve == 0) return t_vm_pair<BN254>(np, g1a, g2a, g1b, g2b, out, fexp);
    if (curve == 1) return t_vm_pair<BLS381>(np, g1a, g2a, g1b, g2b, out, fexp);
    return t_vm_pair<BLS377>(np, g1a, g2a, g1b, g2b, out, fexp);
}

extern "C" int he_vm_pairing_split(int curve, int np, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                                   uint8_t* out, int fexp) {
    if (curve == 0) return t_vm_pair<BN254, true>(np, g1a, g2a, g1b, g2b, out, fexp);
    if (curve == 1) return t_vm_pair<BLS381, true>(np, g1a, g2a, g1b, g2b, out, fexp);
    return t_vm_pair<BLS377, true>(np, g1a, g2a, g1b, g2b, out, fexp);
}

// fixed-Q pairing on the host: line tables of the G2 arguments (precompute_lines), then miller_fixed -- the control flow of
// vm_lines_kernel + vm_pairing_fixed_kernel
template <class C> static int t_vm_pair_fixed(int np, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b,
                                              const uint8_t* g2b, uint8_t* out, int fexp) {
    constexpr int N = C::N;
    typedef VmTables<C> TB;
    std::vector<uint32_t> kb((size_t)VM_KBANK * 2 * N);
    for (int i = 0; i < VM_KBANK; i++) vm_fill_kbank<C>(kb.data() + (size_t)i * 2 * N, i);
    const size_t rowsz = (size_t)VmDriver<C>::nlines() * 3 * 2 * N;
    std::vector<uint32_t> rows[2];
    bool qinf[2] = {false, false};
    int err = 0;
    for (int k = 0; k < np; k++) {
        std::vector<uint32_t> slots((size_t)TB::NSLOTS * 2 * N, 0u);
        VmDriver<C> D;
        D.ctx.slots = slots.data(); D.ctx.kbank = kb.data(); D.ctx.live = 3;
        D.words = TB::host_words(); D.dir = TB::host_dir(); D.role = 0;
        unsigned z = 0;
        for (int r = 2; r < VM_G; r++) z |= (unsigned)D.load_coord(r, 0, nullptr, k ? g2b : g2a, false, &err) << r;
        qinf[k] = (z & 60u) == 60u;
        rows[k].assign(rowsz, 0u);
        D.precompute_lines(rows[k].data());
    }
    if (err) return 1;
    std::vector<uint32_t> slots((size_t)TB::NSLOTS * 2 * N, 0u);
    VmDriver<C> D;
    D.ctx.slots = slots.data(); D.ctx.kbank = kb.data();
    D.words = TB::host_words(); D.dir = TB::host_dir(); D.role = 0;
    unsigned m0 = 0, m1 = 0;
    for (int r = 0; r < 2; r++) {
        m0 |= (unsigned)D.load_coord(r, 0, g1a, nullptr, false, &err) << r;
        if (np == 2) m1 |= (unsigned)D.load_coord(r, 1, g1b, nullptr, false, &err) << r;
    }
    if (err) return 1;
    const bool dead0 = ((m0 & 3u) == 3u) || qinf[0];
    const bool dead1 = np == 2 ? (((m1 & 3u) == 3u) || qinf[1]) : true;
    D.ctx.live = (dead0 ? 0u : 1u) | (dead1 ? 0u : 2u);
    const uint32_t* r1 = np == 2 ? rows[1].data() : rows[0].data();
    uint32_t fb = np == 1 ? D.template miller_fixed<1>(rows[0].data(), r1) : D.template miller_fixed<2>(rows[0].data(), r1);
    if (fexp) fb = D.final_exp(fb);
    for (int r = 0; r < VM_G; r++) D.store_coeff(r, fb, out, false);
    return 0;
}
extern "C" int he_vm_pairing_fixed(int curve, int np, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b,
                                   const uint8_t* g2b, uint8_t* out, int fexp) {
    if (curve == 0) return t_vm_pair_fixed<BN254>(np, g1a, g2a, g1b, g2b, out, fexp);
    if (curve == 1) return t_vm_pair_fixed<BLS381>(np, g1a, g2a, g1b, g2b, out, fexp);
    return t_vm_pair_fixed<BLS377>(np, g1a, g2a, g1b, g2b, out, fexp);
}
// Gt.Exp ladder on the host (VmDriver::gt_exp)
template <class C> static int t_vm_gt_exp(const uint8_t* gt, const uint8_t* k, uint8_t* out) {
    constexpr int N = C::N;
    typedef VmTables<C> TB;
    std::vector<uint32_t> slots((size_t)TB::NSLOTS * 2 * N, 0u), kb((size_t)VM_KBANK * 2 * N);
    for (int i = 0; i < VM_KBANK; i++) vm_fill_kbank<C>(kb.data() + (size_t)i * 2 * N, i);
    VmDriver<C> D;
    D.ctx.slots = slots.data(); D.ctx.kbank = kb.data(); D.ctx.live = 3;
    D.words = TB::host_words(); D.dir = TB::host_dir(); D.role = 0;
    int err = 0;
    for (int r = 0; r < VM_G; r++) D.load_coeff(r, 0, gt, false, &err);
    if (err) return 1;
    uint32_t fb = D.gt_exp(k, VmDriver<C>::scalar_bitlen(k));
    for (int r = 0; r < VM_G; r++) D.store_coeff(r, fb, out, false);
    return 0;
}
extern "C" int he_vm_gt_exp(int curve, const uint8_t* gt, const uint8_t* k, uint8_t* out) {
    if (curve == 0) return t_vm_gt_exp<BN254>(gt, k, out);
    if (curve == 1) return t_vm_gt_exp<BLS381>(gt, k, out);
    return t_vm_gt_exp<BLS377>(gt, k, out);
}

// mul_dot<T>: r = sum a[t]*b[t] / R mod p  (operands as 3 consecutive Fp each)
template <class C> static void t_dot(int T, const uint32_t* a, const uint32_t* b, uint32_t* o) {
    typedef FpOps<C> F; typename F::E x[3], y[3], z;
    for (int t = 0; t < 3; t++) for (int i = 0; i < C::N; i++) { x[t].l[i] = a[t * C::N + i]; y[t].l[i] = b[t * C::N + i]; }
    if (T == 1) F::template mul_dot<1>(z, x, y); else if (T == 2) F::template mul_dot<2>(z, x, y); else F::template mul_dot<3>(z, x, y);
    for (int i = 0; i < C::N; i++) o[i] = z.l[i];
}
extern "C" void he_fp_dot(int curve, int T, const uint32_t* a, const uint32_t* b, uint32_t* o) {
    if (curve == 0) t_dot<BN254>(T, a, b, o);
    else if (curve == 1) t_dot<BLS381>(T, a, b, o);
    else t_dot<BLS377>(T, a, b, o);
}

#include "../../mathlib_b200/csrc/points.cuh"
// point (de)compression / validation, one item (points.cuh: the function the kernel runs per thread)
extern "C" int he_point_codec(int curve, int g2, int op, const uint8_t* in, uint8_t* out, uint32_t flags) {
    if (curve == 0) return g2 ? point_codec_item<BN254, 1>(op, in, out, flags) : point_codec_item<BN254, 0>(op, in, out, flags);
    if (curve == 1) return g2 ? point_codec_item<BLS381, 1>(op, in, out, flags) : point_codec_item<BLS381, 0>(op, in, out, flags);
    return g2 ? point_codec_item<BLS377, 1>(op, in, out, flags) : point_codec_item<BLS377, 0>(op, in, out, flags);
}

#include "../../mathlib_b200/csrc/hash_to_g1.cuh"
// hash-to-G1, one message (hash_to_g1.cuh: the function the kernel runs per thread); curve ids 3 / 5 standard, 6 / 7 BBS
extern "C" int he_hash_to_g1(int cid, const uint8_t* msg, size_t mlen, const uint8_t* dst, size_t dlen, uint8_t* out) {
    if (cid != 3 && cid != 5 && cid != 6 && cid != 7) return 1;
    HashToG1::item(cid == 6 || cid == 7, msg, mlen, dst, dlen, out, false);
    return 0;
}

// batch affine normalisation of Jacobian points (points.cuh: one thread's batch)
extern "C" void he_g1_normalize(int curve, size_t n, const uint32_t* jac, uint8_t* out) {
    if (curve == 0) g1_normalize_items<BN254>(n, jac, out, false);
    else if (curve == 1) g1_normalize_items<BLS381>(n, jac, out, false);
    else g1_normalize_items<BLS377>(n, jac, out, false);
}
