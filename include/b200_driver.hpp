// Header-only C++ mirror of mathlib's driver interface for the hot path, over the C ABI (b200.h).
// Same names / argument order / error behaviour as the reference: driver.Curve.Pairing / Pairing2 / FExp /
// MultiScalarMul (reference driver/math.go:49-57,170), driver.G1.Mul / Mul2 / Mul2InPlace / Add (driver/math.go:249-288).
// Failures throw std::runtime_error where the Go drivers panic (reference driver/gurvy/bn254.go:249-251).
// Elements hold the reference's Bytes() encoding, so bytes() is directly comparable with the reference drivers.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "b200.h"

namespace b200drv {

typedef std::vector<unsigned char> Bytes;

inline void check(const char* op, int rc) {
    if (rc != 0) throw std::runtime_error(std::string(op) + " failed [" + b200_last_error() + "]");
}

struct Curve;
struct Zr { Bytes be32; const Bytes& bytes() const { return be32; } };
struct G2 { Bytes raw; const Bytes& bytes() const { return raw; } };
struct Gt {
    Bytes raw;
    const Bytes& bytes() const { return raw; }
    bool isUnity() const {
        for (size_t i = 0; i + 1 < raw.size(); i++) if (raw[i]) return false;
        return !raw.empty() && raw.back() == 1;
    }
    bool equals(const Gt& o) const { return raw == o.raw; }
};
struct G1 {
    const Curve* c;
    Bytes raw;
    const Bytes& bytes() const { return raw; }
    bool equals(const G1& o) const { return raw == o.raw; }
    G1 mul(const Zr& a) const;                               // driver.G1.Mul: receiver untouched
    G1 mul2(const Zr& e, const G1& Q, const Zr& f) const;    // driver.G1.Mul2
    void mul2InPlace(const Zr& e, const G1& Q, const Zr& f) { raw = mul2(e, Q, f).raw; }
    void add(const G1& o);                                   // mutates the receiver
};

struct Curve {
    int id;
    int fp;
    explicit Curve(int curve_id) : id(curve_id), fp(b200_fp_bytes(curve_id)) {
        if (fp <= 0) throw std::runtime_error(std::string("unknown curve [") + b200_last_error() + "]");
    }
    size_t g1Size() const { return 2 * (size_t)fp; }
    size_t g2Size() const { return 4 * (size_t)fp; }
    size_t gtSize() const { return 12 * (size_t)fp; }

    Gt pairing(const G2& p2, const G1& p1) const {
        Gt r; r.raw.resize(gtSize());
        check("pairing", b200_pairing_batch(id, 1, p1.raw.data(), p2.raw.data(), r.raw.data(), 0));
        return r;
    }
    Gt pairing2(const G2& p2a, const G2& p2b, const G1& p1a, const G1& p1b) const {
        Gt r; r.raw.resize(gtSize());
        check("pairing 2", b200_pairing2_batch(id, 1, p1a.raw.data(), p2a.raw.data(), p1b.raw.data(), p2b.raw.data(),
                                               r.raw.data(), 0));
        return r;
    }
    Gt fexp(const Gt& a) const {
        Gt r; r.raw.resize(gtSize());
        check("final exponentiation", b200_fexp_batch(id, 1, a.raw.data(), r.raw.data(), 0));
        return r;
    }
    // a length mismatch yields infinity: the reference discards gnark's error (driver/gurvy/bn254.go:242)
    G1 multiScalarMul(const std::vector<G1>& a, const std::vector<Zr>& b) const {
        G1 r{this, Bytes(g1Size(), 0)};
        if (a.size() != b.size()) { if (fp == 48) r.raw[0] = 0x40; return r; }
        Bytes pts, ks;
        for (size_t i = 0; i < a.size(); i++) {
            pts.insert(pts.end(), a[i].raw.begin(), a[i].raw.end());
            ks.insert(ks.end(), b[i].be32.begin(), b[i].be32.end());
        }
        check("multi scalar mul", b200_g1_msm(id, a.size(), pts.data(), ks.data(), r.raw.data(), 0));
        return r;
    }
    // G2 counterpart of multiScalarMul: sum [b_i]a_i over the twist, one Pippenger run on the device
    G2 multiScalarMulG2(const std::vector<G2>& a, const std::vector<Zr>& b) const {
        G2 r; r.raw.assign(g2Size(), 0);
        if (a.size() != b.size()) { if (fp == 48) r.raw[0] = 0x40; return r; }
        Bytes pts, ks;
        for (size_t i = 0; i < a.size(); i++) {
            pts.insert(pts.end(), a[i].raw.begin(), a[i].raw.end());
            ks.insert(ks.end(), b[i].be32.begin(), b[i].be32.end());
        }
        check("g2 multi scalar mul", b200_g2_msm(id, a.size(), pts.data(), ks.data(), r.raw.data(), 0));
        return r;
    }
    // callers next to the hot path (reference driver/math.go:307-310, 344-359)
    G2 g2Mul(const G2& p, const Zr& a) const {
        G2 r; r.raw.resize(g2Size());
        check("g2 mul", b200_g2_mul_batch(id, 1, p.raw.data(), a.be32.data(), r.raw.data(), 0));
        return r;
    }
    G2 g2Add(const G2& p, const G2& q) const {
        Bytes two(p.raw);
        two.insert(two.end(), q.raw.begin(), q.raw.end());
        G2 r; r.raw.resize(g2Size());
        check("g2 add", b200_g2_sum(id, 2, two.data(), r.raw.data(), 0));
        return r;
    }
    Gt gtExp(const Gt& a, const Zr& k) const {
        Gt r; r.raw.resize(gtSize());
        check("gt exp", b200_gt_exp_batch(id, 1, a.raw.data(), k.be32.data(), r.raw.data(), 0));
        return r;
    }
    Gt gtMul(const Gt& a, const Gt& b) const {
        Gt r; r.raw.resize(gtSize());
        check("gt mul", b200_gt_mul_batch(id, 1, a.raw.data(), b.raw.data(), r.raw.data(), 0));
        return r;
    }
    Gt gtInverse(const Gt& a) const {
        Gt r; r.raw.resize(gtSize());
        check("gt inverse", b200_gt_inv_batch(id, 1, a.raw.data(), r.raw.data(), 0));
        return r;
    }
    // (de)serialisation with the reference's checks (driver/gurvy/bn254.go:339-377): throws on a bad encoding
    G1 newG1FromCompressed(const Bytes& b) const {
        G1 r{this, Bytes(g1Size())};
        check("set bytes", b200_g1_decompress_batch(id, 1, b.data(), r.raw.data(), 0));
        return r;
    }
    G2 newG2FromCompressed(const Bytes& b) const {
        G2 r; r.raw.resize(g2Size());
        check("set bytes", b200_g2_decompress_batch(id, 1, b.data(), r.raw.data(), 0));
        return r;
    }
    Bytes g1Compressed(const G1& p) const {
        Bytes out(fp);
        check("compress", b200_g1_compress_batch(id, 1, p.raw.data(), out.data(), 0));
        return out;
    }
    Bytes g2Compressed(const G2& p) const {
        Bytes out(2 * (size_t)fp);
        check("compress", b200_g2_compress_batch(id, 1, p.raw.data(), out.data(), 0));
        return out;
    }
    bool g1IsValid(const G1& p) const {
        unsigned char ok = 0;
        check("validate", b200_g1_validate_batch(id, 1, p.raw.data(), &ok, 0));
        return ok == 1;
    }
    // batch entry points: contiguous slabs
    Bytes pairing2Batch(size_t n, const Bytes& g1a, const Bytes& g2a, const Bytes& g1b, const Bytes& g2b, unsigned flags) const {
        Bytes out((flags & B200_OUT_UNITY_ONLY) ? n : n * gtSize());
        check("pairing 2", b200_pairing2_batch(id, n, g1a.data(), g2a.data(), g1b.data(), g2b.data(), out.data(), flags));
        return out;
    }
};

inline G1 G1::mul(const Zr& a) const {
    G1 r{c, Bytes(raw.size())};
    check("g1 mul", b200_g1_mul_batch(c->id, 1, raw.data(), a.be32.data(), r.raw.data(), 0));
    return r;
}
inline G1 G1::mul2(const Zr& e, const G1& Q, const Zr& f) const {
    G1 r{c, Bytes(raw.size())};
    check("g1 mul2", b200_g1_mul2_batch(c->id, 1, raw.data(), e.be32.data(), Q.raw.data(), f.be32.data(), r.raw.data(), 0));
    return r;
}
inline void G1::add(const G1& o) {
    Bytes two(raw);
    two.insert(two.end(), o.raw.begin(), o.raw.end());
    Bytes out(raw.size());
    check("g1 add", b200_g1_sum(c->id, 2, two.data(), out.data(), 0));
    raw = out;
}

}  // namespace b200drv
