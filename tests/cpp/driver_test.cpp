// C++ mirror smoke test (run by tests/test_gpu_cpp_mirror.py on the GPU box): the assertions of reference
// math_test.go runPairingTest (423-455) and the Mul2 line of runG1Test (290) through include/b200_driver.hpp,
// on the golden inputs passed as hex on the command line.
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/b200_driver.hpp"
using namespace b200drv;

static Bytes unhex(const char* s) {
    Bytes out;
    size_t n = strlen(s);
    for (size_t i = 0; i + 1 < n; i += 2) {
        unsigned v;
        sscanf(s + i, "%2x", &v);
        out.push_back((unsigned char)v);
    }
    return out;
}
static std::string hex(const Bytes& b) {
    std::string s;
    char buf[3];
    for (unsigned char c : b) { snprintf(buf, sizeof buf, "%02x", c); s += buf; }
    return s;
}

// argv: curve g1a g2a g1b g2b scalar_e scalar_f
int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage\n"); return 2; }
    try {
        Curve c(atoi(argv[1]));
        G1 p1a{&c, unhex(argv[2])}, p1b{&c, unhex(argv[4])};
        G2 p2a{unhex(argv[3])}, p2b{unhex(argv[5])};
        Zr e{unhex(argv[6])}, f{unhex(argv[7])};
        Gt both = c.fexp(c.pairing2(p2a, p2b, p1a, p1b));
        printf("pairing2_fexp %s\n", hex(both.bytes()).c_str());
        printf("pairing_fexp %s\n", hex(c.fexp(c.pairing(p2a, p1a)).bytes()).c_str());
        G1 m2 = p1a.mul2(e, p1b, f);
        G1 sum = p1a.mul(e);
        sum.add(p1b.mul(f));
        printf("mul2 %s\n", hex(m2.bytes()).c_str());
        printf("mul2_eq_mul_add %d\n", (int)m2.equals(sum));
        G1 before = p1a;
        (void)p1a.mul(e);
        printf("receiver_unchanged %d\n", (int)before.equals(p1a));
        printf("msm %s\n", hex(c.multiScalarMul({p1a, p1b}, {e, f}).bytes()).c_str());
        // bilinearity through the neighbouring ops (math_test.go:399-420): e([e]P, [f]Q) == e(P, Q)^e^f, and a * a^-1 == 1
        Gt base = c.fexp(c.pairing(p2a, p1a));
        Gt lhs = c.fexp(c.pairing(c.g2Mul(p2a, f), p1a.mul(e)));
        Gt rhs = c.gtExp(c.gtExp(base, e), f);
        printf("bilinear %d\n", (int)lhs.equals(rhs));
        printf("inverse_ok %d\n", (int)c.gtMul(base, c.gtInverse(base)).isUnity());
        printf("g2_add_is_double %d\n", (int)(c.g2Add(p2a, p2a).raw == c.g2Mul(p2a, Zr{unhex("0000000000000000000000000000000000000000000000000000000000000002")}).raw));
        // compressed round trip and validation
        printf("compress_roundtrip %d\n", (int)(c.newG1FromCompressed(c.g1Compressed(p1a)).raw == p1a.raw &&
                                                c.newG2FromCompressed(c.g2Compressed(p2a)).raw == p2a.raw && c.g1IsValid(p1a)));
        // error behaviour: bad curve id throws
        bool threw = false;
        try { Curve bad(99); } catch (const std::runtime_error&) { threw = true; }
        printf("bad_curve_throws %d\n", (int)threw);
    } catch (const std::exception& ex) {
        fprintf(stderr, "exception: %s\n", ex.what());
        return 1;
    }
    return 0;
}
