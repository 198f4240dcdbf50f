// Warp-cooperative pairing kernel: 6 lanes per pairing product, state in shared memory, control flow below,
// arithmetic in the VM interpreter (vm.cuh) running microcode compiled from mathlib_b200/vm/programs.py.
// The sequence of program runs is the one of mathlib_b200/vm/driver_ref.py (checked against the oracle on the CPU).
//
// Same results as pairing.cuh (thread-per-pairing): raw Miller value per SURVEY A.5, canonical value after FExp.
// Replaces the same reference calls: driver.Curve.Pairing / Pairing2 / FExp (reference driver/math.go:51-57).
#pragma once
#include "vm.cuh"
#include "kernels.cuh"

namespace b200 {

template <class C> struct VmTables;
#define B200_VM_TABLES(NAME, CURVE)                                                                       \
    static const uint32_t H_VMW_##NAME[] = VM_##NAME##_WORDS;                                             \
    static const VmDirEntry H_VMD_##NAME[VP_COUNT] = VM_##NAME##_DIR;                                     \
    template <> struct VmTables<CURVE> {                                                                  \
        static constexpr int NSLOTS = VM_##NAME##_NSLOTS;                                                 \
        static constexpr int NREGS = VM_##NAME##_NREGS;                                                   \
        static constexpr int NWORDS = VM_##NAME##_NWORDS;                                                 \
        static constexpr int NWORDS_CORE = VM_##NAME##_NWORDS_CORE;                                       \
        static constexpr int QSTRIDE = VM_##NAME##_QSTRIDE;                                               \
        static constexpr int QBASE = VM_##NAME##_QBASE;                                                   \
        static constexpr int PBASE = VM_##NAME##_PBASE;                                                   \
        static constexpr int TBASE = VM_##NAME##_TBASE;                                                   \
        static constexpr int LINEOUT = VM_##NAME##_LINEOUT;                                               \
        static const uint32_t* host_words() { return H_VMW_##NAME; }                                      \
        static const VmDirEntry* host_dir() { return H_VMD_##NAME; }                                      \
    };
B200_VM_TABLES(BN254, BN254)
B200_VM_TABLES(BLS381, BLS381)
B200_VM_TABLES(BLS377, BLS377)

constexpr int VM_KBANK = 18;      // one, b', zero, 15 Frobenius constants

// constant bank in Montgomery form: [one, btw, zero, frob1[1..5], frob2[1..5], frob3[1..5]]
template <class C>
B200_HD void vm_fill_kbank(uint32_t* kb, int idx) {
    constexpr int N = C::N;
    const uint32_t* src = nullptr;
    if (idx == 0) {
        for (int i = 0; i < N; i++) { kb[i] = C::one()[i]; kb[N + i] = 0; }
        return;
    }
    if (idx == 2) {
        for (int i = 0; i < 2 * N; i++) kb[i] = 0;
        return;
    }
    if (idx == 1) src = C::K().btw;
    else {
        int k = (idx - 3) / 5, i = (idx - 3) % 5;
        src = (k == 0 ? C::K().frob1 : k == 1 ? C::K().frob2 : C::K().frob3) + i * 2 * N;
    }
    for (int i = 0; i < 2 * N; i++) kb[i] = src[i];
}

// SPLIT: three lanes per role (vm.cuh split_stage1 / split_stage2) -- the small-batch kernel; `sub` is the lane's index
// within its role and `xch` the group's exchange scratch (VM_G * 3 * W words)
template <class C, bool SPLIT = false>
struct VmDriver {
    static constexpr int N = C::N;
    static constexpr int SW = 2 * N;
    typedef Vm<C> M;
    typedef VmTables<C> TB;
    typedef Codec<C> CD;
    typedef FpOps<C> F;

    typename M::Ctx ctx;
    const uint32_t* words;
    const VmDirEntry* dir;
    int role;                 // device: this lane's role (0..5, or -1 idle); host emulation: ignored
    int sub = 0;              // SPLIT: lane within the role (0..2)
    uint32_t* xch = nullptr;  // SPLIT: exchange scratch of the group

    B200_HD void sync() {
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
    B200_HD void run(int prog, uint32_t b1, uint32_t b2, uint32_t b3) {
        ctx.base[0] = b1; ctx.base[1] = b2; ctx.base[2] = b3;
        const uint32_t* w = words + dir[prog].offset;
        const uint32_t nph = dir[prog].phases;
        for (uint32_t ph = 0; ph < nph; ph++) {
#if defined(__CUDA_ARCH__)
            if (SPLIT) {
                const uint32_t* ww = w + (ph * VM_G + (role >= 0 ? role : 0)) * VM_OP_WORDS;
                uint32_t* x = xch + (role >= 0 ? role : 0) * 3 * M::W;
                if (role >= 0) M::split_stage1(ctx.slots, ctx.kbank, b1, b2, b3, ctx.live, ww, sub, x);
                __syncwarp();
                if (role >= 0) M::split_stage2(ctx.slots, ctx.kbank, b1, b2, b3, ctx.live, ww, sub, x);
            } else if (role >= 0) {
                M::exec_op(ctx.slots, ctx.kbank, b1, b2, b3, ctx.live, w + (ph * VM_G + role) * VM_OP_WORDS);
            }
            __syncwarp();
#else
            for (int r = 0; r < VM_G; r++) {
                const uint32_t* ww = w + (ph * VM_G + r) * VM_OP_WORDS;
                if (SPLIT) {
                    uint32_t* x = xch + r * 3 * M::W;
                    for (int sl = 0; sl < 3; sl++) M::split_stage1(ctx.slots, ctx.kbank, b1, b2, b3, ctx.live, ww, sl, x);
                    for (int sl = 0; sl < 3; sl++) M::split_stage2(ctx.slots, ctx.kbank, b1, b2, b3, ctx.live, ww, sl, x);
                } else {
                    M::exec_op(ctx.slots, ctx.kbank, b1, b2, b3, ctx.live, ww);
                }
            }
#endif
        }
    }

    // ---- Miller loop; returns the slot base of f
    template <int NP>
    B200_HD uint32_t miller() {
        uint32_t cur = 0, nxt = 6;
        run(NP == 1 ? VP_INIT1 : VP_INIT2, cur, nxt, 0);
        const bool dbl_swaps = ((1 + NP) & 1) != 0, add_swaps = (NP & 1) != 0;
        const int len = PairingOps<C>::loop_len();
        int top = len - 1;
        while (PairingOps<C>::loop_digit(top) == 0) top--;
        for (int i = top - 1; i >= 0; i--) {
            run(NP == 1 ? VP_DBL1 : VP_DBL2, cur, nxt, 0);
            if (dbl_swaps) { uint32_t t = cur; cur = nxt; nxt = t; }
            int d = PairingOps<C>::loop_digit(i);
            if (d) {
                run(NP == 1 ? VP_ADD1 : VP_ADD2, cur, nxt, d > 0 ? 0u : 1u);
                if (add_swaps) { uint32_t t = cur; cur = nxt; nxt = t; }
            }
        }
        if (C::FAMILY == FAMILY_BN) run(NP == 1 ? VP_BNTAIL1 : VP_BNTAIL2, cur, nxt, 0);
        if (C::X_NEG) {
            run(VP_CONJ, nxt, cur, 0);
            uint32_t t = cur; cur = nxt; nxt = t;
        }
        return cur;
    }

    // ---- fixed-Q Miller loop: the G2 side of every step comes from a precomputed line table -- per step
    // three P-independent Fp2 coefficients per pair, copied into the pair's T slots -- and the SQRLINE / LINE programs
    // evaluate them at P and fold them into f.  Same f as miller<NP>() (vm/driver_ref.py: miller_fixed).
    // rowK: table row of pair K (nlines * 3 slots); lane `role` copies coefficient role % 3 of pair role / 3.
    template <int NP>
    B200_HD void load_lines(const uint32_t* row0, const uint32_t* row1, int step) {
#if defined(__CUDA_ARCH__)
        if (role >= 0 && role < 3 * NP) {
            const int k = role / 3, j = role % 3;
            const uint4* src = reinterpret_cast<const uint4*>((k ? row1 : row0) + ((size_t)step * 3 + j) * SW);
            uint4* dst = reinterpret_cast<uint4*>(ctx.slots + (TB::TBASE + 3 * k + j) * SW);
#pragma unroll
            for (int i = 0; i < SW / 4; i++) dst[i] = src[i];
        }
        __syncwarp();
#else
        for (int r = 0; r < 3 * NP; r++) {
            const int k = r / 3, j = r % 3;
            const uint32_t* src = (k ? row1 : row0) + ((size_t)step * 3 + j) * SW;
            uint32_t* dst = ctx.slots + (TB::TBASE + 3 * k + j) * SW;
            for (int i = 0; i < SW; i++) dst[i] = src[i];
        }
#endif
    }
    template <int NP>
    B200_HD uint32_t miller_fixed(const uint32_t* row0, const uint32_t* row1) {
        uint32_t cur = 0, nxt = 6;
        run(NP == 1 ? VP_INIT1 : VP_INIT2, cur, nxt, 0);
        const bool sq_swaps = ((1 + NP) & 1) != 0, ln_swaps = (NP & 1) != 0;
        const int len = PairingOps<C>::loop_len();
        int top = len - 1;
        while (PairingOps<C>::loop_digit(top) == 0) top--;
        int step = 0;
        for (int i = top - 1; i >= 0; i--) {
            load_lines<NP>(row0, row1, step++);
            run(NP == 1 ? VP_SQRLINE1 : VP_SQRLINE2, cur, nxt, 0);
            if (sq_swaps) { uint32_t t = cur; cur = nxt; nxt = t; }
            if (PairingOps<C>::loop_digit(i)) {
                load_lines<NP>(row0, row1, step++);
                run(NP == 1 ? VP_LINE1 : VP_LINE2, cur, nxt, 0);
                if (ln_swaps) { uint32_t t = cur; cur = nxt; nxt = t; }
            }
        }
        if (C::FAMILY == FAMILY_BN)
            for (int t2 = 0; t2 < 2; t2++) {
                load_lines<NP>(row0, row1, step++);
                run(NP == 1 ? VP_LINE1 : VP_LINE2, cur, nxt, 0);
                if (ln_swaps) { uint32_t t = cur; cur = nxt; nxt = t; }
            }
        if (C::X_NEG) {
            run(VP_CONJ, nxt, cur, 0);
            uint32_t t = cur; cur = nxt; nxt = t;
        }
        return cur;
    }
    // number of line evaluations of one Miller loop = rows of 3 slots per table entry
    static B200_HD int nlines() {
        const int len = PairingOps<C>::loop_len();
        int top = len - 1;
        while (PairingOps<C>::loop_digit(top) == 0) top--;
        int n = 0;
        for (int i = top - 1; i >= 0; i--) n += 1 + (PairingOps<C>::loop_digit(i) ? 1 : 0);
        return n + (C::FAMILY == FAMILY_BN ? 2 : 0);
    }
    // table of one Q: runs the G2 side of the loop on pair 0 and writes (r0, r1, r2) of every step to `row`
    B200_HD void precompute_lines(uint32_t* row) {
        run(VP_INIT1, 0, 6, 0);
        const int len = PairingOps<C>::loop_len();
        int top = len - 1;
        while (PairingOps<C>::loop_digit(top) == 0) top--;
        int step = 0;
        for (int i = top - 1; i >= 0; i--) {
            run(VP_PRE_DBL, 0, 0, 0);
            store_line(row, step++);
            const int d = PairingOps<C>::loop_digit(i);
            if (d) {
                run(VP_PRE_ADD, 0, 0, d > 0 ? 0u : 1u);
                store_line(row, step++);
            }
        }
        if (C::FAMILY == FAMILY_BN) {
            run(VP_PRE_TAIL1, 0, 0, 0);
            store_line(row, step++);
            run(VP_PRE_TAIL2, 0, 0, 0);
            store_line(row, step++);
        }
    }
    B200_HD void store_line(uint32_t* row, int step) {
#if defined(__CUDA_ARCH__)
        if (role >= 0 && role < 3) {
            const uint32_t* src = ctx.slots + (TB::LINEOUT + role) * SW;
            uint32_t* dst = row + ((size_t)step * 3 + role) * SW;
            for (int i = 0; i < SW; i++) dst[i] = src[i];
        }
        __syncwarp();
#else
        for (int r = 0; r < 3; r++)
            for (int i = 0; i < SW; i++) row[((size_t)step * 3 + r) * SW + i] = ctx.slots[(TB::LINEOUT + r) * SW + i];
#endif
    }

    // ---- final exponentiation over NREGS Fp12 registers; returns the slot base of the result
    struct Regs {
        uint32_t freemask;
        B200_HD int alloc() {
            int r = 0;
            while (!((freemask >> r) & 1)) r++;
            freemask &= ~(1u << r);
            return r;
        }
        B200_HD void release(int r) { freemask |= 1u << r; }
    };
    Regs R;
    B200_HD int op(int prog, int a, int b = 0) {
        int d = R.alloc();
        run(prog, 6 * d, 6 * a, 6 * b);
        return d;
    }
    B200_HD int expx(int z) {
        int top = 63;
        while (!((C::X_ABS >> top) & 1)) top--;
        int acc = -1;
        for (int i = top - 1; i >= 0; i--) {
            int n = op(VP_CYCLO_SQR, acc < 0 ? z : acc);
            if (acc >= 0) R.release(acc);
            acc = n;
            if ((C::X_ABS >> i) & 1) {
                n = op(VP_F12_MUL, acc, z);
                R.release(acc);
                acc = n;
            }
        }
        if (C::X_NEG) {
            int n = op(VP_CONJ, acc);
            R.release(acc);
            acc = n;
        }
        return acc;
    }
    B200_HD uint32_t final_exp(uint32_t f_base) {
        int f = (int)(f_base / 6);
        R.freemask = ((1u << TB::NREGS) - 1) & ~(1u << f);
        int x;
        int ri = op(VP_F12_INV, f);
        int c = op(VP_CONJ, f);
        R.release(f);
        int t = op(VP_F12_MUL, c, ri);
        R.release(c); R.release(ri);
        int u = op(VP_FROB2, t);
        f = op(VP_F12_MUL, u, t);
        R.release(u); R.release(t);
        if (C::FAMILY == FAMILY_BLS12) {
            int t0 = op(VP_CYCLO_SQR, f);
            int t1 = expx(f);
            int t2 = op(VP_CONJ, f);
            x = op(VP_F12_MUL, t1, t2); R.release(t1); R.release(t2); t1 = x;
            t2 = expx(t1);
            x = op(VP_CONJ, t1); R.release(t1); t1 = x;
            x = op(VP_F12_MUL, t1, t2); R.release(t1); R.release(t2); t1 = x;
            t2 = expx(t1);
            x = op(VP_FROB1, t1); R.release(t1); t1 = x;
            x = op(VP_F12_MUL, t1, t2); R.release(t1); R.release(t2); t1 = x;
            x = op(VP_F12_MUL, f, t0); R.release(f); R.release(t0); f = x;
            t0 = expx(t1);
            t2 = expx(t0);
            R.release(t0);
            t0 = op(VP_FROB2, t1);
            x = op(VP_CONJ, t1); R.release(t1); t1 = x;
            x = op(VP_F12_MUL, t1, t2); R.release(t1); R.release(t2); t1 = x;
            x = op(VP_F12_MUL, t1, t0); R.release(t1); R.release(t0); t1 = x;
            x = op(VP_F12_MUL, f, t1); R.release(f); R.release(t1); f = x;
            return 6u * f;
        }
        int e = expx(f);
        int t0 = op(VP_CONJ, e); R.release(e);
        x = op(VP_CYCLO_SQR, t0); R.release(t0); t0 = x;
        int t1 = op(VP_CYCLO_SQR, t0);
        x = op(VP_F12_MUL, t0, t1); R.release(t1); t1 = x;
        e = expx(t1);
        int t2 = op(VP_CONJ, e); R.release(e);
        int t3 = op(VP_CONJ, t1);
        x = op(VP_F12_MUL, t2, t3); R.release(t1); R.release(t3); t1 = x;
        t3 = op(VP_CYCLO_SQR, t2);
        int t4 = expx(t3);
        x = op(VP_F12_MUL, t1, t4); R.release(t4); R.release(t1); t4 = x;
        x = op(VP_F12_MUL, t0, t4); R.release(t3); t3 = x;
        x = op(VP_F12_MUL, t2, t4); R.release(t0); R.release(t2); t0 = x;
        x = op(VP_F12_MUL, f, t0); R.release(t0); t0 = x;
        t2 = op(VP_FROB1, t3);
        x = op(VP_F12_MUL, t2, t0); R.release(t2); R.release(t0); t0 = x;
        t2 = op(VP_FROB2, t4); R.release(t4);
        x = op(VP_F12_MUL, t2, t0); R.release(t2); R.release(t0); t0 = x;
        t2 = op(VP_CONJ, f); R.release(f);
        x = op(VP_F12_MUL, t2, t3); R.release(t2); R.release(t3); t2 = x;
        x = op(VP_FROB3, t2); R.release(t2); t2 = x;
        x = op(VP_F12_MUL, t2, t0); R.release(t2); R.release(t0); t0 = x;
        return 6u * t0;
    }

    // ---- Gt.Exp (reference driver/math.go:359; impls bn254.go:187-191, kilic/bls12-381.go:185-199): the Fp12 in
    // register 0 raised to a 256-bit exponent (32 bytes big-endian, used as given) by a left-to-right ladder of generic
    // squarings and multiplies predicated on the group's exponent digit -- the groups of a warp run in lock-step, each on
    // its own exponent.  `top` = warp-uniform bit length (>= this exponent's).  Same sequence as vm/driver_ref.py:gt_exp.
    // Returns the slot base of the result.
    B200_HD uint32_t gt_exp(const uint8_t* k_be32, int top) {
        // registers: a = 0, a^2 = 18, a^3 = 24; accumulator ping-pongs between 6 and 12.  2-bit windows: two squarings and
        // one multiply by a^d per digit d, the table entry picked through the group's own B3 base (0 / 18 / 24), d = 0
        // predicated off.
        uint32_t cur = 6, oth = 12;
        run(VP_F12_SQR, 18, 0, 0);
        run(VP_F12_MUL, 24, 18, 0);
        run(VP_GT_ONE, cur, 0, 0);
        for (int j = (top + 1) / 2 - 1; j >= 0; j--) {
            const int bit = 2 * j;
            const uint32_t d = (k_be32[31 - (bit >> 3)] >> (bit & 7)) & 3u;
            run(VP_F12_SQR, oth, cur, 0);
            run(VP_F12_SQR, cur, oth, 0);
            ctx.live = d != 0 ? 1u : 0u;
            run(VP_F12_MULP, oth, cur, d == 2 ? 18u : (d == 3 ? 24u : 0u));
            const uint32_t t = cur; cur = oth; oth = t;
        }
        return cur;
    }
    static B200_HD int scalar_bitlen(const uint8_t* k_be32) {
        for (int b = 0; b < 32; b++)
            if (k_be32[b]) {
                int l = 8;
                while (!((k_be32[b] >> (l - 1)) & 1)) l--;
                return (31 - b) * 8 + l;
            }
        return 0;
    }

    // ---- I/O: lane `r` of the group converts coordinate r of pair k (P.x, P.y, Q.x.c0, Q.x.c1, Q.y.c0, Q.y.c1)
    // returns 1 if the coordinate is zero; *err set on a bad encoding
    B200_HD int load_coord(int r, int k, const uint8_t* g1, const uint8_t* g2, bool mont, int* err) {
        Fp<N> v;
        const int FB = C::FP_BYTES;
        if (r < 2) {
            if (mont) CD::fp_from_mont_words(v, (const uint32_t*)g1 + r * N);
            else {
                uint8_t fl = g1[0] & CD::flag_mask();
                if (C::FLAG_BITS == 3 && fl == 0x40) F::zero(v);
                else if (fl != 0) { *err = 1; F::zero(v); }
                else CD::fp_from_bytes(v, g1 + r * FB, r == 0 ? (uint8_t)~CD::flag_mask() : 0xFF, err);
            }
            uint32_t* dst = ctx.slots + (TB::PBASE + k) * SW + r * N;
            for (int i = 0; i < N; i++) dst[i] = v.l[i];
        } else {
            int q = r - 2;                    // 0: x.c0, 1: x.c1, 2: y.c0, 3: y.c1
            if (mont) CD::fp_from_mont_words(v, (const uint32_t*)g2 + q * N);
            else {
                uint8_t fl = g2[0] & CD::flag_mask();
                // wire order X.A1 | X.A0 | Y.A1 | Y.A0
                int pos = (q ^ 1);
                if (C::FLAG_BITS == 3 && fl == 0x40) F::zero(v);
                else if (fl != 0) { *err = 1; F::zero(v); }
                else CD::fp_from_bytes(v, g2 + pos * FB, pos == 0 ? (uint8_t)~CD::flag_mask() : 0xFF, err);
            }
            uint32_t* dst = ctx.slots + (TB::QBASE + TB::QSTRIDE * k + (q >> 1)) * SW + (q & 1) * N;
            for (int i = 0; i < N; i++) dst[i] = v.l[i];
        }
        return F::is_zero(v) ? 1 : 0;
    }
    // lane r stores w-basis coefficient r (2 Fp) of the Fp12 at slot base fb
    B200_HD void store_coeff(int r, uint32_t fb, uint8_t* out, bool mont) {
        const uint32_t* src = ctx.slots + (fb + r) * SW;
        for (int a = 0; a < 2; a++) {
            Fp<N> v;
            for (int i = 0; i < N; i++) v.l[i] = src[a * N + i];
            int e = ((r & 1) * 3 + (r >> 1)) * 2 + a;       // index in the C0.B0.A0 ... C1.B2.A1 order
            if (mont) CD::fp_to_mont_words((uint32_t*)out + e * N, v);
            else CD::fp_to_bytes(out + (11 - e) * C::FP_BYTES, v);
        }
    }
    // lane r loads w-basis coefficient r (2 Fp) of a Gt element into the Fp12 register at slot base fb
    B200_HD void load_coeff(int r, uint32_t fb, const uint8_t* in, bool mont, int* err) {
        uint32_t* dst = ctx.slots + (fb + r) * SW;
        for (int a = 0; a < 2; a++) {
            Fp<N> v;
            int e = ((r & 1) * 3 + (r >> 1)) * 2 + a;
            if (mont) CD::fp_from_mont_words(v, (const uint32_t*)in + e * N);
            else CD::fp_from_bytes(v, in + (11 - e) * C::FP_BYTES, 0xFF, err);
            for (int i = 0; i < N; i++) dst[a * N + i] = v.l[i];
        }
    }
    // defined output of an item whose input was rejected: lane r zeroes its share (2 Fp) of the Gt element
    B200_HD void store_zero_coeff(int r, uint8_t* out, bool mont) {
        for (int a = 0; a < 2; a++) {
            int e = ((r & 1) * 3 + (r >> 1)) * 2 + a;
            if (mont) { for (int i = 0; i < N; i++) ((uint32_t*)out)[e * N + i] = 0; }
            else { for (int i = 0; i < C::FP_BYTES; i++) out[(11 - e) * C::FP_BYTES + i] = 0; }
        }
    }
    B200_HD bool coeff_is_one_part(int r, uint32_t fb) {
        const uint32_t* src = ctx.slots + (fb + r) * SW;
        uint32_t d = 0;
        for (int i = 0; i < N; i++) {
            d |= src[i] ^ (r == 0 ? C::one()[i] : 0u);
            d |= src[N + i];
        }
        return d == 0;
    }
};

#if defined(__CUDACC__)
// Block shape.  BLS12 curves: 4 warps (20 groups) per block, two blocks per SM -- the operand-scanning wide product needs
// ~200 registers, so 8 resident warps is the register-file limit (12 warps need <= 168 registers, which the slower
// product-scanning variant reaches: measured 103.7 ms vs 99.9 ms per 65,536 checks for this shape).
#ifndef B200_VM_WARPS_MAX
#define B200_VM_WARPS_MAX 4
#endif
#ifndef B200_VM_WARPS_BN
#define B200_VM_WARPS_BN 11
#endif
// BN254 (8-limb operands, ~160 registers): one block of 11 warps per SM for the Pairing / FExp kernels (55 slot files +
// the core microcode = 220 KB); the kernels that stage the full microcode (fixed-Q, Gt ops) use 10
template <class C> __host__ __device__ constexpr int vm_warps() { return C::N == 8 ? B200_VM_WARPS_BN : B200_VM_WARPS_MAX; }
template <class C> __host__ __device__ constexpr int vm_warps_x() { return C::N == 8 ? 10 : 4; }
#define B200_VM_GROUPS_PER_WARP 5
#define B200_VM_GROUP_PAD 4          // words; staggers the groups across shared-memory banks

template <class C>
__host__ __device__ constexpr size_t vm_group_stride() { return (size_t)VmTables<C>::NSLOTS * 2 * C::N + B200_VM_GROUP_PAD; }
// small batches use 4-warp blocks (20 groups) so that they spread over more SMs
#define B200_VM_WARPS_SMALL 4
// CORE = true: only the programs of the Pairing / FExp kernels are staged (the microcode words before F12_SQR); the
// fixed-Q, Gt and line-table kernels stage everything.  Keeps the headline kernel's shared-memory footprint unchanged
// by the neighbouring rows' programs.
template <class C, bool CORE>
__host__ __device__ constexpr int vm_nwords() { return CORE ? VmTables<C>::NWORDS_CORE : VmTables<C>::NWORDS; }
template <class C, int WARPS, bool CORE = false>
__host__ __device__ constexpr size_t vm_smem_bytes() {
    return 4 * ((size_t)WARPS * B200_VM_GROUPS_PER_WARP * vm_group_stride<C>() + (size_t)VM_KBANK * 2 * C::N +
                (size_t)vm_nwords<C, CORE>() + VP_COUNT * 2);
}

template <class C, int NP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 4 ? 2 : 1)
vm_pairing_kernel(size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                  uint8_t* out, uint32_t flags, int* err, const uint32_t* mc_words, const VmDirEntry* mc_dir) {
    extern __shared__ uint32_t smem[];
    constexpr int N = C::N;
    constexpr int GPB = WARPS * B200_VM_GROUPS_PER_WARP;
    uint32_t* s_slots = smem;
    uint32_t* s_kbank = s_slots + GPB * vm_group_stride<C>();
    uint32_t* s_words = s_kbank + VM_KBANK * 2 * N;
    VmDirEntry* s_dir = reinterpret_cast<VmDirEntry*>(s_words + VmTables<C>::NWORDS_CORE);
    for (int i = threadIdx.x; i < VmTables<C>::NWORDS_CORE; i += blockDim.x) s_words[i] = mc_words[i];
    if (threadIdx.x < VP_COUNT) s_dir[threadIdx.x] = mc_dir[threadIdx.x];
    if (threadIdx.x < VM_KBANK) vm_fill_kbank<C>(s_kbank + threadIdx.x * 2 * N, threadIdx.x);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = lane / VM_G;                          // group within warp (5 = idle lanes 30, 31)
    const int role = gw < B200_VM_GROUPS_PER_WARP ? lane % VM_G : -1;
    const int gblock = warp * B200_VM_GROUPS_PER_WARP + (gw < B200_VM_GROUPS_PER_WARP ? gw : 0);
    const size_t item = (size_t)blockIdx.x * GPB + gblock;
    const bool active = role >= 0 && item < n;

    VmDriver<C> D;
    D.ctx.slots = s_slots + (size_t)gblock * vm_group_stride<C>();
    D.ctx.kbank = s_kbank;
    D.words = s_words;
    D.dir = s_dir;
    D.role = active ? role : -1;
    typedef Codec<C> CD;
    const bool in_mont = flags & FLAG_IN_MONT;
    int e = 0, z0 = 0, z1 = 0;
    if (active) {
        z0 = D.load_coord(role, 0, g1a + item * CD::g1_size(), g2a + item * CD::g2_size(), in_mont, &e);
        if (NP == 2) z1 = D.load_coord(role, 1, g1b + item * CD::g1_size(), g2b + item * CD::g2_size(), in_mont, &e);
    }
    // group-wide view of the zero flags: pair k is dead if P (roles 0,1) or Q (roles 2..5) is all zero
    const unsigned b0 = __ballot_sync(0xffffffffu, z0 != 0), b1 = __ballot_sync(0xffffffffu, z1 != 0);
    const unsigned any_err = __ballot_sync(0xffffffffu, e != 0);
    const int sh = (gw < B200_VM_GROUPS_PER_WARP ? gw : 0) * VM_G;
    // a rejected encoding fails ITS item only: the flag is raised, the group runs on with both pairs dead and writes a
    // defined output (zero element / verdict 0); the other groups of the warp are unaffected
    const bool bad = active && ((any_err >> sh) & 63u) != 0;
    if (bad && role == 0) atomicExch(err, 1);
    const unsigned m0 = (b0 >> sh) & 63u, m1 = (b1 >> sh) & 63u;
    const bool dead0 = bad || ((m0 & 3u) == 3u) || ((m0 & 60u) == 60u);
    const bool dead1 = NP == 2 ? (bad || ((m1 & 3u) == 3u) || ((m1 & 60u) == 60u)) : true;
    D.ctx.live = (dead0 ? 0u : 1u) | (dead1 ? 0u : 2u);
    __syncwarp();
    uint32_t fb = D.template miller<NP>();
    if (flags & FLAG_FEXP) fb = D.final_exp(fb);
    if (flags & FLAG_UNITY) {
        const bool ok = active ? D.coeff_is_one_part(role, fb) : true;
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (active && role == 0) out[item] = (!bad && ((okm >> sh) & 63u) == 63u) ? 1 : 0;
    } else if (active) {
        if (bad) D.store_zero_coeff(role, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
        else D.store_coeff(role, fb, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
    }
}
// Small batches (BASELINE configs[0]: 1,024 checks): fewer checks than the GPU has warps, so the run time is the length of
// ONE check's dependency chain.  This kernel spends three lanes per role -- one Karatsuba product each (vm.cuh split
// mode) -- and one check per warp: 18 lanes, a third of the multiplier chain per lane.  Bit-identical to vm_pairing_kernel.
#define B200_VM_SPLIT_WARPS 4
template <class C>
__host__ __device__ constexpr size_t vm_split_smem_bytes() {
    return 4 * ((size_t)B200_VM_SPLIT_WARPS * (vm_group_stride<C>() + VM_G * 3 * (2 * C::N + 1) + 3) + (size_t)VM_KBANK * 2 * C::N +
                (size_t)VmTables<C>::NWORDS_CORE + VP_COUNT * 2);
}
template <class C, int NP>
__global__ void __launch_bounds__(B200_VM_SPLIT_WARPS * 32, 2)
vm_pairing_split_kernel(size_t n, const uint8_t* g1a, const uint8_t* g2a, const uint8_t* g1b, const uint8_t* g2b,
                        uint8_t* out, uint32_t flags, int* err, const uint32_t* mc_words, const VmDirEntry* mc_dir) {
    extern __shared__ uint32_t smem[];
    constexpr int N = C::N;
    constexpr int WARPS = B200_VM_SPLIT_WARPS;
    constexpr int XW = VM_G * 3 * (2 * N + 1) + 3;       // exchange words per group (padded to a multiple of 4)
    uint32_t* s_slots = smem;
    uint32_t* s_xch = s_slots + WARPS * vm_group_stride<C>();
    uint32_t* s_kbank = s_xch + WARPS * XW;
    uint32_t* s_words = s_kbank + VM_KBANK * 2 * N;
    VmDirEntry* s_dir = reinterpret_cast<VmDirEntry*>(s_words + VmTables<C>::NWORDS_CORE);
    for (int i = threadIdx.x; i < VmTables<C>::NWORDS_CORE; i += blockDim.x) s_words[i] = mc_words[i];
    if (threadIdx.x < VP_COUNT) s_dir[threadIdx.x] = mc_dir[threadIdx.x];
    if (threadIdx.x < VM_KBANK) vm_fill_kbank<C>(s_kbank + threadIdx.x * 2 * N, threadIdx.x);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = lane < 3 * VM_G ? lane / 3 : -1;     // lanes 18..31 idle
    const int sub = lane % 3;
    const size_t item = (size_t)blockIdx.x * WARPS + warp;
    const bool active = role >= 0 && item < n;
    const bool lead = active && sub == 0;                  // the lane of a role that does its I/O

    VmDriver<C, true> D;
    D.ctx.slots = s_slots + (size_t)warp * vm_group_stride<C>();
    D.ctx.kbank = s_kbank;
    D.words = s_words;
    D.dir = s_dir;
    D.role = active ? role : -1;
    D.sub = sub;
    D.xch = s_xch + (size_t)warp * XW;
    typedef Codec<C> CD;
    const bool in_mont = flags & FLAG_IN_MONT;
    int e = 0, z0 = 0, z1 = 0;
    if (lead) {
        z0 = D.load_coord(role, 0, g1a + item * CD::g1_size(), g2a + item * CD::g2_size(), in_mont, &e);
        if (NP == 2) z1 = D.load_coord(role, 1, g1b + item * CD::g1_size(), g2b + item * CD::g2_size(), in_mont, &e);
    }
    // lane 3 r speaks for role r
    const unsigned b0 = __ballot_sync(0xffffffffu, z0 != 0), b1 = __ballot_sync(0xffffffffu, z1 != 0);
    const unsigned any_err = __ballot_sync(0xffffffffu, e != 0);
    const bool bad = active && any_err != 0;
    if (bad && lane == 0) atomicExch(err, 1);
    auto zero_role = [](unsigned m, int r) { return ((m >> (3 * r)) & 1u) != 0; };
    const bool pz0 = zero_role(b0, 0) && zero_role(b0, 1), qz0 = zero_role(b0, 2) && zero_role(b0, 3) && zero_role(b0, 4) && zero_role(b0, 5);
    const bool pz1 = zero_role(b1, 0) && zero_role(b1, 1), qz1 = zero_role(b1, 2) && zero_role(b1, 3) && zero_role(b1, 4) && zero_role(b1, 5);
    const bool dead0 = bad || pz0 || qz0;
    const bool dead1 = NP == 2 ? (bad || pz1 || qz1) : true;
    D.ctx.live = (dead0 ? 0u : 1u) | (dead1 ? 0u : 2u);
    __syncwarp();
    uint32_t fb = D.template miller<NP>();
    if (flags & FLAG_FEXP) fb = D.final_exp(fb);
    if (flags & FLAG_UNITY) {
        const bool ok = lead ? D.coeff_is_one_part(role, fb) : true;
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (lead && role == 0) out[item] = (!bad && okm == 0xffffffffu) ? 1 : 0;
    } else if (lead) {
        if (bad) D.store_zero_coeff(role, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
        else D.store_coeff(role, fb, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
    }
}

// standalone driver.Curve.FExp on the VM (same slot file / microcode as the pairing kernel)
template <class C, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 4 ? 2 : 1)
vm_fexp_kernel(size_t n, const uint8_t* in, uint8_t* out, uint32_t flags, int* err, const uint32_t* mc_words,
               const VmDirEntry* mc_dir) {
    extern __shared__ uint32_t smem[];
    constexpr int N = C::N;
    constexpr int GPB = WARPS * B200_VM_GROUPS_PER_WARP;
    uint32_t* s_slots = smem;
    uint32_t* s_kbank = s_slots + GPB * vm_group_stride<C>();
    uint32_t* s_words = s_kbank + VM_KBANK * 2 * N;
    VmDirEntry* s_dir = reinterpret_cast<VmDirEntry*>(s_words + VmTables<C>::NWORDS_CORE);
    for (int i = threadIdx.x; i < VmTables<C>::NWORDS_CORE; i += blockDim.x) s_words[i] = mc_words[i];
    if (threadIdx.x < VP_COUNT) s_dir[threadIdx.x] = mc_dir[threadIdx.x];
    if (threadIdx.x < VM_KBANK) vm_fill_kbank<C>(s_kbank + threadIdx.x * 2 * N, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = lane / VM_G;
    const int role = gw < B200_VM_GROUPS_PER_WARP ? lane % VM_G : -1;
    const int gblock = warp * B200_VM_GROUPS_PER_WARP + (gw < B200_VM_GROUPS_PER_WARP ? gw : 0);
    const size_t item = (size_t)blockIdx.x * GPB + gblock;
    const bool active = role >= 0 && item < n;
    VmDriver<C> D;
    D.ctx.slots = s_slots + (size_t)gblock * vm_group_stride<C>();
    D.ctx.kbank = s_kbank;
    D.ctx.live = 3;
    D.words = s_words;
    D.dir = s_dir;
    D.role = active ? role : -1;
    typedef Codec<C> CD;
    int e = 0;
    if (active) D.load_coeff(role, 0, in + item * CD::gt_size(), flags & FLAG_IN_MONT, &e);
    const int sh = (gw < B200_VM_GROUPS_PER_WARP ? gw : 0) * VM_G;
    const unsigned errm = __ballot_sync(0xffffffffu, e != 0);      // every lane votes: no short-circuit around it
    const bool bad = active && ((errm >> sh) & 63u) != 0;      // per item, see vm_pairing_kernel
    if (bad && role == 0) atomicExch(err, 1);
    __syncwarp();
    uint32_t fb = 0;
    if (flags & FLAG_FEXP) fb = D.final_exp(0);
    if (flags & FLAG_UNITY) {
        const bool ok = active ? D.coeff_is_one_part(role, fb) : true;
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (active && role == 0) out[item] = (!bad && ((okm >> sh) & 63u) == 63u) ? 1 : 0;
    } else if (active) {
        if (bad) D.store_zero_coeff(role, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
        else D.store_coeff(role, fb, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
    }
}

// ---- fixed-Q pairings --------------------------------------------------------------------------------
// line tables of n_q G2 points: lines[q] = nlines * 3 slots (Montgomery), qinf[q] = 1 for the point at infinity
template <class C, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 4 ? 2 : 1)
vm_lines_kernel(size_t n_q, const uint8_t* g2, uint32_t* lines, uint8_t* qinf, uint32_t flags, int* err,
                const uint32_t* mc_words, const VmDirEntry* mc_dir) {
    extern __shared__ uint32_t smem[];
    constexpr int N = C::N;
    constexpr int GPB = WARPS * B200_VM_GROUPS_PER_WARP;
    uint32_t* s_slots = smem;
    uint32_t* s_kbank = s_slots + GPB * vm_group_stride<C>();
    uint32_t* s_words = s_kbank + VM_KBANK * 2 * N;
    VmDirEntry* s_dir = reinterpret_cast<VmDirEntry*>(s_words + VmTables<C>::NWORDS);
    for (int i = threadIdx.x; i < VmTables<C>::NWORDS; i += blockDim.x) s_words[i] = mc_words[i];
    if (threadIdx.x < VP_COUNT) s_dir[threadIdx.x] = mc_dir[threadIdx.x];
    if (threadIdx.x < VM_KBANK) vm_fill_kbank<C>(s_kbank + threadIdx.x * 2 * N, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = lane / VM_G;
    const int role = gw < B200_VM_GROUPS_PER_WARP ? lane % VM_G : -1;
    const int gblock = warp * B200_VM_GROUPS_PER_WARP + (gw < B200_VM_GROUPS_PER_WARP ? gw : 0);
    const size_t item = (size_t)blockIdx.x * GPB + gblock;
    const bool active = role >= 0 && item < n_q;
    VmDriver<C> D;
    D.ctx.slots = s_slots + (size_t)gblock * vm_group_stride<C>();
    D.ctx.kbank = s_kbank;
    D.ctx.live = 3;
    D.words = s_words;
    D.dir = s_dir;
    D.role = active ? role : -1;
    typedef Codec<C> CD;
    int e = 0, z = 0;
    if (active) {
        if (role >= 2) z = D.load_coord(role, 0, nullptr, g2 + item * CD::g2_size(), flags & FLAG_IN_MONT, &e);
        else {
            uint32_t* dst = D.ctx.slots + (VmTables<C>::PBASE) * 2 * N + role * N;
            for (int i = 0; i < N; i++) dst[i] = 0;
        }
    }
    const unsigned bz = __ballot_sync(0xffffffffu, z != 0);
    const int sh = (gw < B200_VM_GROUPS_PER_WARP ? gw : 0) * VM_G;
    const unsigned errm = __ballot_sync(0xffffffffu, e != 0);      // every lane votes: no short-circuit around it
    const bool bad = active && ((errm >> sh) & 63u) != 0;
    if (bad && role == 0) atomicExch(err, 1);        // the upload call reads the flag back and fails
    if (active && role == 0) qinf[item] = (bad || ((bz >> sh) & 60u) == 60u) ? 1 : 0;
    __syncwarp();
    D.precompute_lines(lines + (active ? item : 0) * (size_t)VmDriver<C>::nlines() * 3 * 2 * N);
}

// Pairing / Pairing2 with every G2 argument taken from a line table: check i uses rows qa_idx[i] (and qb_idx[i]); a null
// index array means row 0 (pair a) / row 1 (pair b) for every check -- the BLS / BBS verification pattern.
template <class C, int NP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 4 ? 2 : 1)
vm_pairing_fixed_kernel(size_t n, const uint8_t* g1a, const uint32_t* qa_idx, const uint8_t* g1b, const uint32_t* qb_idx,
                        const uint32_t* lines, const uint8_t* qinf, uint32_t n_q, uint8_t* out, uint32_t flags, int* err,
                        const uint32_t* mc_words, const VmDirEntry* mc_dir) {
    extern __shared__ uint32_t smem[];
    constexpr int N = C::N;
    constexpr int GPB = WARPS * B200_VM_GROUPS_PER_WARP;
    uint32_t* s_slots = smem;
    uint32_t* s_kbank = s_slots + GPB * vm_group_stride<C>();
    uint32_t* s_words = s_kbank + VM_KBANK * 2 * N;
    VmDirEntry* s_dir = reinterpret_cast<VmDirEntry*>(s_words + VmTables<C>::NWORDS);
    for (int i = threadIdx.x; i < VmTables<C>::NWORDS; i += blockDim.x) s_words[i] = mc_words[i];
    if (threadIdx.x < VP_COUNT) s_dir[threadIdx.x] = mc_dir[threadIdx.x];
    if (threadIdx.x < VM_KBANK) vm_fill_kbank<C>(s_kbank + threadIdx.x * 2 * N, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = lane / VM_G;
    const int role = gw < B200_VM_GROUPS_PER_WARP ? lane % VM_G : -1;
    const int gblock = warp * B200_VM_GROUPS_PER_WARP + (gw < B200_VM_GROUPS_PER_WARP ? gw : 0);
    const size_t item = (size_t)blockIdx.x * GPB + gblock;
    const bool active = role >= 0 && item < n;
    VmDriver<C> D;
    D.ctx.slots = s_slots + (size_t)gblock * vm_group_stride<C>();
    D.ctx.kbank = s_kbank;
    D.words = s_words;
    D.dir = s_dir;
    D.role = active ? role : -1;
    typedef Codec<C> CD;
    const bool in_mont = flags & FLAG_IN_MONT;
    int e = 0, z0 = 0, z1 = 0;
    if (active && role < 2) {
        z0 = D.load_coord(role, 0, g1a + item * CD::g1_size(), nullptr, in_mont, &e);
        if (NP == 2) z1 = D.load_coord(role, 1, g1b + item * CD::g1_size(), nullptr, in_mont, &e);
    }
    const unsigned b0 = __ballot_sync(0xffffffffu, z0 != 0), b1 = __ballot_sync(0xffffffffu, z1 != 0);
    const int sh = (gw < B200_VM_GROUPS_PER_WARP ? gw : 0) * VM_G;
    const size_t it = active ? item : 0;
    uint32_t ra = qa_idx ? qa_idx[it] : 0u, rb = NP == 2 ? (qb_idx ? qb_idx[it] : 1u) : 0u;
    // row indices may come straight from device memory (B200_DEVICE_PTRS): an index outside the table is an error of
    // its item (flag raised, verdict 0 / zero element), never an address
    if (ra >= n_q || rb >= n_q) { e = 1; ra = 0; rb = 0; }
    const unsigned errm = __ballot_sync(0xffffffffu, e != 0);      // every lane votes: no short-circuit around it
    const bool bad = active && ((errm >> sh) & 63u) != 0;
    if (bad && role == 0) atomicExch(err, 1);
    const bool dead0 = bad || (((b0 >> sh) & 3u) == 3u) || qinf[ra] != 0;
    const bool dead1 = NP == 2 ? (bad || (((b1 >> sh) & 3u) == 3u) || qinf[rb] != 0) : true;
    D.ctx.live = (dead0 ? 0u : 1u) | (dead1 ? 0u : 2u);
    __syncwarp();
    const size_t rowsz = (size_t)VmDriver<C>::nlines() * 3 * 2 * N;
    uint32_t fb = D.template miller_fixed<NP>(lines + ra * rowsz, lines + rb * rowsz);
    if (flags & FLAG_FEXP) fb = D.final_exp(fb);
    if (flags & FLAG_UNITY) {
        const bool ok = active ? D.coeff_is_one_part(role, fb) : true;
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (active && role == 0) out[item] = (!bad && ((okm >> sh) & 63u) == 63u) ? 1 : 0;
    } else if (active) {
        if (bad) D.store_zero_coeff(role, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
        else D.store_coeff(role, fb, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
    }
}

// Gt.Mul / Gt.Inverse / Gt.Exp batches on the VM (reference driver/math.go:339-360, impls
// bn254.go:187-203, kilic/bls12-381.go:185-210).  b = second operand (MUL) or 32-byte big-endian exponents (EXP).
enum : int { GT_OP_MUL = 0, GT_OP_INV = 1, GT_OP_EXP = 2 };
template <class C, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 4 ? 2 : 1)
vm_gt_kernel(int opk, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out, uint32_t flags, int* err,
             const uint32_t* mc_words, const VmDirEntry* mc_dir) {
    extern __shared__ uint32_t smem[];
    constexpr int N = C::N;
    constexpr int GPB = WARPS * B200_VM_GROUPS_PER_WARP;
    uint32_t* s_slots = smem;
    uint32_t* s_kbank = s_slots + GPB * vm_group_stride<C>();
    uint32_t* s_words = s_kbank + VM_KBANK * 2 * N;
    VmDirEntry* s_dir = reinterpret_cast<VmDirEntry*>(s_words + VmTables<C>::NWORDS);
    for (int i = threadIdx.x; i < VmTables<C>::NWORDS; i += blockDim.x) s_words[i] = mc_words[i];
    if (threadIdx.x < VP_COUNT) s_dir[threadIdx.x] = mc_dir[threadIdx.x];
    if (threadIdx.x < VM_KBANK) vm_fill_kbank<C>(s_kbank + threadIdx.x * 2 * N, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = lane / VM_G;
    const int role = gw < B200_VM_GROUPS_PER_WARP ? lane % VM_G : -1;
    const int gblock = warp * B200_VM_GROUPS_PER_WARP + (gw < B200_VM_GROUPS_PER_WARP ? gw : 0);
    const size_t item = (size_t)blockIdx.x * GPB + gblock;
    const bool active = role >= 0 && item < n;
    VmDriver<C> D;
    D.ctx.slots = s_slots + (size_t)gblock * vm_group_stride<C>();
    D.ctx.kbank = s_kbank;
    D.ctx.live = 3;
    D.words = s_words;
    D.dir = s_dir;
    D.role = active ? role : -1;
    typedef Codec<C> CD;
    int e = 0;
    if (active) {
        D.load_coeff(role, 0, a + item * CD::gt_size(), flags & FLAG_IN_MONT, &e);
        if (opk == GT_OP_MUL) D.load_coeff(role, 6, b + item * CD::gt_size(), flags & FLAG_IN_MONT, &e);
    }
    const int sh = (gw < B200_VM_GROUPS_PER_WARP ? gw : 0) * VM_G;
    const unsigned errm = __ballot_sync(0xffffffffu, e != 0);      // every lane votes: no short-circuit around it
    const bool bad = active && ((errm >> sh) & 63u) != 0;      // per item, see vm_pairing_kernel
    if (bad && role == 0) atomicExch(err, 1);
    __syncwarp();
    uint32_t fb = 12;
    if (opk == GT_OP_MUL) D.run(VP_F12_MUL, 12, 0, 6);
    else if (opk == GT_OP_INV) D.run(VP_F12_INV, 12, 0, 0);
    else {
        const uint8_t* k = b + (active ? item : 0) * 32;
        const int len = active ? VmDriver<C>::scalar_bitlen(k) : 0;
        const int top = __reduce_max_sync(0xffffffffu, len);
        fb = D.gt_exp(k, top);
    }
    if (active) {
        if (bad) D.store_zero_coeff(role, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
        else D.store_coeff(role, fb, out + item * CD::gt_size(), flags & FLAG_OUT_MONT);
    }
}
#endif

}  // namespace b200
