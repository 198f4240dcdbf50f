#!/usr/bin/env python3
"""Generate csrc/constants.h: per-curve Montgomery constants for the CUDA kernels.

Stand-alone (does NOT import oracle/): the numbers are derived from the curve seeds with
the published BN / BLS12 polynomials.  tests/test_constants.py cross-checks
the generated header against the oracle and against the limbs the reference holds
(driver/kilic/custom.go:26-29).  Run:  python mathlib_b200/csrc/gen_constants.py
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def bn(x):
    return (36 * x**4 + 36 * x**3 + 24 * x**2 + 6 * x + 1, 36 * x**4 + 36 * x**3 + 18 * x**2 + 6 * x + 1)


def bls12(x):
    r = x**4 - x**2 + 1
    return ((x - 1) ** 2 * r // 3 + x, r)


CURVES = [
    # name, family, x, limbs, b, beta, xi, twist, g1, g2
    dict(name='BN254', fam='bn', x=4965661367192848881, n=8, b=3, beta=-1, xi=(9, 1), twist='D'),
    dict(name='BLS381', fam='bls12', x=-0xd201000000010000, n=12, b=4, beta=-1, xi=(1, 1), twist='M'),
    dict(name='BLS377', fam='bls12', x=0x8508c00000000001, n=12, b=1, beta=-5, xi=(0, 1), twist='D'),
]


def glv_params(p, r, x, b):
    """GLV endomorphism constants of a BLS12 G1 (y^2 = x^3 + b): lambda = x^2 - 1 satisfies lambda^2 + lambda + 1 = 0
    mod r, and [lambda](X, Y) = (beta * X, Y) for one of the two primitive cube roots of unity beta in Fp -- picked here
    by checking the identity on a point of order r."""
    lam = x * x - 1
    assert (lam * lam + lam + 1) % r == 0

    def add(P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if (P[1] + Q[1]) % p == 0:
                return None
            m = 3 * P[0] * P[0] * pow(2 * P[1], -1, p) % p
        else:
            m = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
        x3 = (m * m - P[0] - Q[0]) % p
        return (x3, (m * (P[0] - x3) - P[1]) % p)

    def mul(P, k):
        R_ = None
        while k:
            if k & 1:
                R_ = add(R_, P)
            P = add(P, P)
            k >>= 1
        return R_

    assert p % 4 == 3 or True
    h = (p - x) // r                       # #E(Fp) = p + 1 - t, t = x + 1
    assert h * r == p - x
    G = None
    xx = 1
    while G is None:
        rhs = (xx ** 3 + b) % p
        if pow(rhs, (p - 1) // 2, p) == 1:
            # square root: p = 3 mod 4 for BLS12-381; Tonelli-Shanks otherwise
            y = tonelli(rhs, p)
            G = mul((xx, y), h)
        xx += 1
    assert mul(G, r) is None
    L = mul(G, lam)
    assert L[1] == G[1]
    beta = L[0] * pow(G[0], -1, p) % p
    assert beta != 1 and pow(beta, 3, p) == 1
    return lam, beta


def tonelli(a, p):
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, rr = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        bb = pow(c, 1 << (m - i - 1), p)
        m, c = i, bb * bb % p
        t, rr = t * c % p, rr * bb % p
    return rr


def h2c_bls381(p, r, x):
    """Constants of BLS12-381 hash-to-G1 (hash_to_g1.cuh): the simplified-SWU parameters of the 11-isogenous curve
    E': y^2 = x^3 + A x + B (the values the reference holds at driver/kilic/custom.go:26-42) and the 11-isogeny E' -> E,
    DERIVED here with Velu's formulas: E'(Fp) has one rational subgroup of order 11, its quotient is y^2 = x^3 + 4 * 11^6,
    and (x, y) -> (x / 11^2, y / 11^3) lands on E.  tests/test_hash_to_g1.py checks the resulting map against RFC 9380's
    known answers.  Returns (A, B, Z, xnum, xden, ynum, yden) with coefficient lists low degree first (dens monic)."""
    A = 0x144698a3b8e9433d693a02c96d4982b0ea985383ee66a8d8e8981aefd881ac98936f8da0e0f97f5cf428082d584c1d
    B = 0x12e2908d11688030018b12e8753eee3b2016c1f0f24f4070a0b9c14fcef35ef55a23215a316ceaa5d1cc48e98e172be0

    def add(P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if (P[1] + Q[1]) % p == 0:
                return None
            m = (3 * P[0] * P[0] + A) * pow(2 * P[1], -1, p) % p
        else:
            m = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
        x3 = (m * m - P[0] - Q[0]) % p
        return (x3, (m * (P[0] - x3) - P[1]) % p)

    def mul(P, k):
        R_ = None
        while k:
            if k & 1:
                R_ = add(R_, P)
            P = add(P, P)
            k >>= 1
        return R_

    def pmul(a, b):
        o = [0] * (len(a) + len(b) - 1)
        for i, u in enumerate(a):
            for j, v in enumerate(b):
                o[i + j] = (o[i + j] + u * v) % p
        return o

    def padd(a, b):
        n = max(len(a), len(b))
        return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % p for i in range(n)]

    def pder(a):
        return [a[i] * i % p for i in range(1, len(a))]

    cof = (x - 1) ** 2 // 3 * r // 121
    xx = 0
    while True:
        xx += 1
        rhs = (xx ** 3 + A * xx + B) % p
        y = pow(rhs, (p + 1) // 4, p)
        if y * y % p != rhs:
            continue
        Q = mul((xx, y), cof)
        if Q is None:
            continue
        if mul(Q, 11) is not None:
            Q = mul(Q, 11)
        break
    assert mul(Q, 11) is None
    terms, h, v, w = [], [1], 0, 0
    for i in range(1, 6):
        xi, yi = mul(Q, i)
        vi, ui = 2 * (3 * xi * xi + A) % p, 4 * yi * yi % p
        v, w = (v + vi) % p, (w + ui + xi * vi) % p
        terms.append((xi, vi, ui))
        h = pmul(h, [(-xi) % p, 1])
    assert (A - 5 * v) % p == 0 and (B - 7 * w) % p == 4 * 11 ** 6
    h2 = pmul(h, h)
    N = pmul([0, 1], h2)
    for xi, vi, ui in terms:
        others = [1]
        for xj, _, _ in terms:
            if xj != xi:
                others = pmul(others, [(-xj) % p, 1])
        o2 = pmul(others, others)
        N = padd(N, [c * vi % p for c in pmul([(-xi) % p, 1], o2)])
        N = padd(N, [c * ui % p for c in o2])
    Yn = padd(pmul(pder(N), h), [c * (p - 2) % p for c in pmul(N, pder(h))])
    h3 = pmul(h2, h)
    i2, i3 = pow(121, -1, p), pow(1331, -1, p)
    return A, B, 11, [c * i2 % p for c in N], h2, [c * i3 % p for c in Yn], h3


def limbs(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def fmt(vs):
    return '{' + ','.join('0x%08xu' % v for v in vs) + '}'


def main():
    out = ['// GENERATED by gen_constants.py -- do not edit.',
           '#pragma once',
           '// Field order: p, one(R), r2(R^2), p_minus_2 (unused slot kept 0), inv32, then Fp2 constants',
           '']
    for c in CURVES:
        p, r = (bn if c['fam'] == 'bn' else bls12)(c['x'])
        n = c['n']
        R = 1 << (32 * n)
        beta = c['beta'] % p
        xi = (c['xi'][0] % p, c['xi'][1] % p)

        def f2mul(a, b):
            return ((a[0] * b[0] + beta * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

        def f2pow(a, e):
            res = (1, 0)
            while e:
                if e & 1:
                    res = f2mul(res, a)
                a = f2mul(a, a)
                e >>= 1
            return res

        def f2inv(a):
            nrm = (a[0] * a[0] - beta * a[1] * a[1]) % p
            ni = pow(nrm, -1, p)
            return (a[0] * ni % p, -a[1] * ni % p)

        def mont(v):
            return limbs(v * R % p, n)

        def mont2(a):
            return mont(a[0]) + mont(a[1])

        if c['twist'] == 'D':
            b2 = f2mul((c['b'], 0), f2inv(xi))
        else:
            b2 = f2mul((c['b'], 0), xi)
        inv32 = (-pow(p, -1, 1 << 32)) % (1 << 32)
        name = c['name']
        out.append('// %s: p = 0x%x' % (name, p))
        out.append('//        r = 0x%x' % r)
        out.append('#define %s_P %s' % (name, fmt(limbs(p, n))))
        out.append('#define %s_ONE %s' % (name, fmt(mont(1))))
        out.append('#define %s_R2 %s' % (name, fmt(limbs(R * R % p, n))))
        out.append('#define %s_INV32 0x%08xu' % (name, inv32))
        out.append('#define %s_B %s' % (name, fmt(mont(c['b']))))
        out.append('#define %s_B3 %s' % (name, fmt(mont(3 * c['b']))))
        out.append('#define %s_BTW %s' % (name, fmt(mont2(b2))))
        out.append('#define %s_ORDER %s' % (name, fmt(limbs(r, 8))))
        # k * p^2 for k = 0..30 as 2n-word integers (offsets that keep the lazy-reduction accumulators non-negative)
        p2 = []
        for k in range(31):
            assert k * p * p < (1 << (64 * n)) or k > (6 if c['beta'] == -1 else 30)
            p2 += limbs((k * p * p) % (1 << (64 * n)), 2 * n)
        out.append('#define %s_P2 %s' % (name, fmt(p2)))
        out.append('#define %s_R3 %s' % (name, fmt(limbs(pow(R, 3, p), n))))
        # 2p and 4p (vm.cuh: conditional subtractions that canonicalise a lazily reduced result)
        assert 4 * p < (1 << (32 * n))
        out.append('#define %s_PK %s' % (name, fmt(limbs(2 * p, n) + limbs(4 * p, n))))
        # GLV (g1.cuh): lambda (4 words), M = floor(2^256 / lambda) (5 words), beta (Montgomery); zeros = no GLV (BN254)
        if c['fam'] == 'bls12':
            lam, gbeta = glv_params(p, r, c['x'], c['b'])
            assert (1 << 120) < lam < (1 << 128)
            out.append('#define %s_GLV_LAMBDA %s' % (name, fmt(limbs(lam, 4))))
            out.append('#define %s_GLV_M %s' % (name, fmt(limbs((1 << 256) // lam, 5))))
            out.append('#define %s_GLV_BETA %s' % (name, fmt(mont(gbeta))))
        else:
            out.append('#define %s_GLV_LAMBDA %s' % (name, fmt([0] * 4)))
            out.append('#define %s_GLV_M %s' % (name, fmt([0] * 5)))
            out.append('#define %s_GLV_BETA %s' % (name, fmt([0] * n)))
        # square roots (codec decompression): p - 1 = 2^s * q, Tonelli-Shanks constants -- TS_EXP = (q-1)/2,
        # TS_Z = g^q for the smallest non-residue g (Montgomery), HALF_P = (p-1)/2 ("lexicographically largest" test)
        q_, s_ = p - 1, 0
        while q_ % 2 == 0:
            q_ //= 2
            s_ += 1
        g_ = 2
        while pow(g_, (p - 1) // 2, p) != p - 1:
            g_ += 1
        out.append('#define %s_TS_S %d' % (name, s_))
        out.append('#define %s_TS_EXP %s' % (name, fmt(limbs((q_ - 1) // 2, n))))
        out.append('#define %s_TS_Z %s' % (name, fmt(mont(pow(g_, q_, p)))))
        out.append('#define %s_HALF_P %s' % (name, fmt(limbs((p - 1) // 2, n))))
        if name == 'BLS381':
            hA, hB, hZ, xnum, xden, ynum, yden = h2c_bls381(p, r, c['x'])
            assert len(xnum) == 12 and len(xden) == 11 and len(ynum) == 16 and len(yden) == 16 and xden[-1] == 1 and yden[-1] == 1
            out.append('#define BLS381_H2C_A %s' % fmt(mont(hA)))
            out.append('#define BLS381_H2C_B %s' % fmt(mont(hB)))
            out.append('#define BLS381_H2C_MBA %s' % fmt(mont(-hB * pow(hA, -1, p) % p)))
            out.append('#define BLS381_H2C_X1EXC %s' % fmt(mont(hB * pow(hZ * hA, -1, p) % p)))
            out.append('#define BLS381_H2C_Z %s' % fmt(mont(hZ)))
            out.append('#define BLS381_H2C_CF %s' % fmt(limbs((1 << 256) * R * R % p, n)))
            for nm, cs in (('XNUM', xnum), ('XDEN', xden[:-1]), ('YNUM', ynum), ('YDEN', yden[:-1])):
                vals = []
                for cc in cs:
                    vals += mont(cc)
                out.append('#define BLS381_H2C_%s %s' % (nm, fmt(vals)))
        # Frobenius: gamma[k][i] = xi^(i*(p^k-1)/6), k=1..3, i=1..5
        for k in (1, 2, 3):
            e = (p ** k - 1) // 6
            vals = []
            for i in range(1, 6):
                vals += mont2(f2pow(xi, i * e))
            out.append('#define %s_FROB%d %s' % (name, k, fmt(vals)))
        out.append('')
    with open(os.path.join(HERE, 'constants.h'), 'w') as f:
        f.write('\n'.join(out) + '\n')
    print('wrote constants.h')


if __name__ == '__main__':
    main()
