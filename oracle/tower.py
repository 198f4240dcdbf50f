"""Fp2 / Fp6 / Fp12 big-integer arithmetic (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Tower: Fp2 = Fp[u]/(u^2-beta), Fp6 = Fp2[v]/(v^3-xi), Fp12 = Fp6[w]/(w^2-v).
Element shapes: Fp2 = (a0,a1); Fp6 = (b0,b1,b2) of Fp2; Fp12 = (c0,c1) of Fp6 --
the same C0/C1 . B0/B1/B2 . A0/A1 naming gnark's E12/E6/E2 and kilic's fe12/fe6/fe2
use (element types on the path: reference bn254.go:183-185, kilic/bls12-381.go:179-183).

Deliberately schoolbook: no Karatsuba, no lazy reduction -- this file is the
authority the optimised C++/CUDA code is compared against, so it stays obvious.
"""


class Tower:
    def __init__(self, params):
        self.P = params
        self.p = params.p
        self.beta = params.beta % params.p
        self.xi = (params.xi[0] % params.p, params.xi[1] % params.p)
        self.f2_zero = (0, 0)
        self.f2_one = (1, 0)
        self.f6_zero = ((0, 0),) * 3
        self.f6_one = ((1, 0), (0, 0), (0, 0))
        self.f12_one = (self.f6_one, self.f6_zero)
        self.f12_zero = (self.f6_zero, self.f6_zero)
        self._frob = {}

    # ---------------- Fp ----------------
    def inv(self, a):
        return pow(a, -1, self.p)

    # ---------------- Fp2 ----------------
    def f2_add(self, a, b):
        p = self.p
        return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)

    def f2_sub(self, a, b):
        p = self.p
        return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)

    def f2_neg(self, a):
        p = self.p
        return (-a[0] % p, -a[1] % p)

    def f2_conj(self, a):
        return (a[0], -a[1] % self.p)

    def f2_mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.beta * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def f2_sqr(self, a):
        return self.f2_mul(a, a)

    def f2_muls(self, a, s):
        p = self.p
        return (a[0] * s % p, a[1] * s % p)

    def f2_mul_xi(self, a):
        return self.f2_mul(a, self.xi)

    def f2_inv(self, a):
        p = self.p
        n = (a[0] * a[0] - self.beta * a[1] * a[1]) % p
        ni = self.inv(n)
        return (a[0] * ni % p, -a[1] * ni % p)

    def f2_pow(self, a, e):
        r = self.f2_one
        while e:
            if e & 1:
                r = self.f2_mul(r, a)
            a = self.f2_mul(a, a)
            e >>= 1
        return r

    def f2_is_zero(self, a):
        return a[0] == 0 and a[1] == 0

    def f2_sqrt(self, a):
        """Any square root in Fp2, or None. Brute simple: via norm trick; p % 4 == 3 or generic Tonelli in Fp."""
        if self.f2_is_zero(a):
            return (0, 0)
        # generic: a^((p^2+1)/... ) is awkward for beta != -1; use the 'complex' method.
        p = self.p
        a0, a1 = a
        if a1 == 0:
            s = self.fp_sqrt(a0)
            if s is not None:
                return (s, 0)
            # sqrt(a0) = t*u with t^2*beta = a0
            t = self.fp_sqrt(a0 * self.inv(self.beta) % p)
            return None if t is None else (0, t)
        n = (a0 * a0 - self.beta * a1 * a1) % p
        s = self.fp_sqrt(n)
        if s is None:
            return None
        half = self.inv(2)
        for sg in (s, -s % p):
            x2 = (a0 + sg) * half % p
            x = self.fp_sqrt(x2)
            if x is None or x == 0:
                continue
            y = a1 * self.inv(2 * x % p) % p
            cand = (x, y)
            if self.f2_sqr(cand) == (a0 % p, a1 % p):
                return cand
        return None

    def fp_sqrt(self, a):
        p = self.p
        a %= p
        if a == 0:
            return 0
        if pow(a, (p - 1) // 2, p) != 1:
            return None
        if p % 4 == 3:
            return pow(a, (p + 1) // 4, p)
        # Tonelli-Shanks
        q, s = p - 1, 0
        while q % 2 == 0:
            q //= 2
            s += 1
        z = 2
        while pow(z, (p - 1) // 2, p) != p - 1:
            z += 1
        m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
        while t != 1:
            i, t2 = 0, t
            while t2 != 1:
                t2 = t2 * t2 % p
                i += 1
            b = pow(c, 1 << (m - i - 1), p)
            m, c = i, b * b % p
            t, r = t * c % p, r * b % p
        return r

    # ---------------- Fp6 ----------------
    def f6_add(self, a, b):
        return tuple(self.f2_add(x, y) for x, y in zip(a, b))

    def f6_sub(self, a, b):
        return tuple(self.f2_sub(x, y) for x, y in zip(a, b))

    def f6_neg(self, a):
        return tuple(self.f2_neg(x) for x in a)

    def f6_mul(self, a, b):
        m, ad, xi = self.f2_mul, self.f2_add, self.f2_mul_xi
        a0, a1, a2 = a
        b0, b1, b2 = b
        c0 = ad(m(a0, b0), xi(ad(m(a1, b2), m(a2, b1))))
        c1 = ad(ad(m(a0, b1), m(a1, b0)), xi(m(a2, b2)))
        c2 = ad(ad(m(a0, b2), m(a1, b1)), m(a2, b0))
        return (c0, c1, c2)

    def f6_mul_v(self, a):
        """multiply by v: (a0,a1,a2) -> (xi*a2, a0, a1)"""
        return (self.f2_mul_xi(a[2]), a[0], a[1])

    def f6_inv(self, a):
        m, sb, xi = self.f2_mul, self.f2_sub, self.f2_mul_xi
        a0, a1, a2 = a
        t0 = sb(m(a0, a0), xi(m(a1, a2)))
        t1 = sb(xi(m(a2, a2)), m(a0, a1))
        t2 = sb(m(a1, a1), m(a0, a2))
        d = self.f2_add(m(a0, t0), xi(self.f2_add(m(a2, t1), m(a1, t2))))
        di = self.f2_inv(d)
        return (m(t0, di), m(t1, di), m(t2, di))

    # ---------------- Fp12 ----------------
    def f12_mul(self, a, b):
        a0, a1 = a
        b0, b1 = b
        t0 = self.f6_mul(a0, b0)
        t1 = self.f6_mul(a1, b1)
        c0 = self.f6_add(t0, self.f6_mul_v(t1))
        c1 = self.f6_add(self.f6_mul(a0, b1), self.f6_mul(a1, b0))
        return (c0, c1)

    def f12_sqr(self, a):
        return self.f12_mul(a, a)

    def f12_conj(self, a):
        return (a[0], self.f6_neg(a[1]))

    def f12_inv(self, a):
        a0, a1 = a
        d = self.f6_sub(self.f6_mul(a0, a0), self.f6_mul_v(self.f6_mul(a1, a1)))
        di = self.f6_inv(d)
        return (self.f6_mul(a0, di), self.f6_neg(self.f6_mul(a1, di)))

    def f12_pow(self, a, e):
        if e < 0:
            a, e = self.f12_inv(a), -e
        r = self.f12_one
        for bit in bin(e)[2:]:
            r = self.f12_sqr(r)
            if bit == '1':
                r = self.f12_mul(r, a)
        return r

    def f12_is_one(self, a):
        return a == self.f12_one

    # coefficient view: Fp12 as sum_{k<6} g_k w^k with g_k in Fp2 (SURVEY A.1 slot map)
    #   C0.B0->w^0, C1.B0->w^1, C0.B1->w^2, C1.B1->w^3, C0.B2->w^4, C1.B2->w^5
    def f12_to_w(self, a):
        (c00, c01, c02), (c10, c11, c12) = a
        return [c00, c10, c01, c11, c02, c12]

    def f12_from_w(self, g):
        return ((g[0], g[2], g[4]), (g[1], g[3], g[5]))

    def frob_consts(self, k):
        """gamma_{k,i} = xi^(i*(p^k-1)/6), i=0..5"""
        if k not in self._frob:
            e = (self.p ** k - 1) // 6
            self._frob[k] = [self.f2_pow(self.xi, i * e) for i in range(6)]
        return self._frob[k]

    def f12_frob(self, a, k=1):
        """a -> a^(p^k).  u^(p) = u * beta^((p-1)/2) = -u for non-residue beta."""
        g = self.f12_to_w(a)
        gam = self.frob_consts(k)
        out = []
        for i in range(6):
            c = g[i] if k % 2 == 0 else self.f2_conj(g[i])
            out.append(self.f2_mul(c, gam[i]))
        return self.f12_from_w(out)
