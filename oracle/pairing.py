"""Optimal-ate pairing oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Two Miller loops and two final exponentiations are restated:

``miller_textbook``   affine chord-and-tangent on the twist, exact (unscaled) untwisted
                      lines -- the mathematical definition, SURVEY A.2.  Authority for
                      Tier A (everything after the final exponentiation).
``miller_projective`` the homogeneous-projective doubling / mixed-addition steps and
                      sparse line slots gnark-crypto's ``MillerLoop`` uses (call sites:
                      reference bn254.go:248,257; bls12-377.go:245,254;
                      bls12-381.go:449,458), restated from SURVEY A.5.  Its raw value is
                      what the Gurvy ``Pairing``/``Pairing2`` return (Tier B, PARITY
                      UNPINNED: formulas recalled, not compared with gnark v0.20.1).
``final_exp_plain``   f^(s*(p^12-1)/r) by square-and-multiply: the definition.
``final_exp``         easy part by conjugate/inverse/Frobenius and hard part by plain pow
                      of s*(p^4-p^2+1)/r -- same value, ~10x faster.

Driver conventions:  gurvy Pairing* = raw Miller value, FExp = final exp
(reference bls12-381.go:448-468);  kilic Pairing* = Miller + final exp, FExp = identity
(reference kilic/bls12-381.go:260-281).
"""
from .curve import Curve


def naf(n):
    out = []
    while n:
        if n & 1:
            d = 2 - (n % 4)
            n -= d
        else:
            d = 0
        out.append(d)
        n >>= 1
    return out


class Pairing:
    def __init__(self, params):
        self.P = params
        self.C = Curve(params)
        self.T = self.C.T
        x = params.x
        if params.family == 'bls12':
            self.loop = abs(x)
            self.loop_digits = [int(b) for b in bin(self.loop)[2:]][::-1]   # LSB first, 0/1
            self.loop_neg = x < 0
        else:
            self.loop = 6 * x + 2
            self.loop_digits = naf(self.loop)                                # LSB first, -1/0/1
            self.loop_neg = False
        p, r = params.p, params.r
        self.hard_exp = params.fexp_scale * ((p ** 4 - p ** 2 + 1) // r)
        assert (p ** 4 - p ** 2 + 1) % r == 0
        self.full_exp = params.fexp_scale * ((p ** 12 - 1) // r)

    # ------------------------------------------------------------------ lines
    def _place(self, coeffs):
        """coeffs: dict power-of-w -> Fp2; returns Fp12"""
        g = [(0, 0)] * 6
        for k, v in coeffs.items():
            g[k] = v
        return self.T.f12_from_w(g)

    def _line_textbook(self, lam, Tpt, Ppt):
        """exact line through T' with twist-slope lam, at P:
        M-type: y_P - lam*x_P/w + (lam*x_T - y_T)/w^3
        D-type: y_P - lam*x_P*w + (lam*x_T - y_T)*w^3
        1/w = w^5/xi, 1/w^3 = w^3/xi   (w^6 = xi)."""
        T = self.T
        xP, yP = Ppt
        xT, yT = Tpt
        c = T.f2_sub(T.f2_mul(lam, xT), yT)
        mlx = T.f2_neg(T.f2_muls(lam, xP))
        if self.P.twist == 'M':
            xi_inv = T.f2_inv(T.xi)
            return self._place({0: (yP % self.P.p, 0), 5: T.f2_mul(mlx, xi_inv), 3: T.f2_mul(c, xi_inv)})
        return self._place({0: (yP % self.P.p, 0), 1: mlx, 3: c})

    # ------------------------------------------------------------------ textbook
    def miller_textbook(self, pairs):
        """pairs: list of (P in G1 affine|None, Q in G2 affine|None). Returns prod f_{lambda,Q}(P)."""
        T, C = self.T, self.C
        pairs = [(P, Q) for (P, Q) in pairs if P is not None and Q is not None]
        f = T.f12_one
        if not pairs:
            return f
        Ts = [Q for (_, Q) in pairs]
        digits = self.loop_digits
        for i in range(len(digits) - 2, -1, -1):
            f = T.f12_sqr(f)
            for k, (P, Q) in enumerate(pairs):
                Tk = Ts[k]
                lam = T.f2_mul(T.f2_muls(T.f2_sqr(Tk[0]), 3), T.f2_inv(T.f2_muls(Tk[1], 2)))
                f = T.f12_mul(f, self._line_textbook(lam, Tk, P))
                Tk = C.g2_add(Tk, Tk)
                d = digits[i]
                if d:
                    A = Q if d == 1 else C.g2_neg(Q)
                    lam = T.f2_mul(T.f2_sub(A[1], Tk[1]), T.f2_inv(T.f2_sub(A[0], Tk[0])))
                    f = T.f12_mul(f, self._line_textbook(lam, Tk, P))
                    Tk = C.g2_add(Tk, A)
                Ts[k] = Tk
        if self.P.family == 'bn':
            for k, (P, Q) in enumerate(pairs):
                Q1, Q2n = self.bn_frobenius_points(Q)
                Tk = Ts[k]
                lam = T.f2_mul(T.f2_sub(Q1[1], Tk[1]), T.f2_inv(T.f2_sub(Q1[0], Tk[0])))
                f = T.f12_mul(f, self._line_textbook(lam, Tk, P))
                Tk = C.g2_add(Tk, Q1)
                lam = T.f2_mul(T.f2_sub(Q2n[1], Tk[1]), T.f2_inv(T.f2_sub(Q2n[0], Tk[0])))
                f = T.f12_mul(f, self._line_textbook(lam, Tk, P))
        if self.loop_neg:
            f = T.f12_conj(f)
        return f

    def bn_frobenius_points(self, Q):
        """Q1 = pi(Q), and -pi^2(Q)  (SURVEY A.2)."""
        T = self.T
        p = self.P.p
        g12 = T.f2_pow(T.xi, (p - 1) // 3)
        g13 = T.f2_pow(T.xi, (p - 1) // 2)
        g22 = T.f2_pow(T.xi, (p * p - 1) // 3)
        Q1 = (T.f2_mul(T.f2_conj(Q[0]), g12), T.f2_mul(T.f2_conj(Q[1]), g13))
        Q2n = (T.f2_mul(Q[0], g22), Q[1])
        return Q1, Q2n

    # ------------------------------------------------------------------ gnark-style projective
    def _double_step(self, Tp):
        """SURVEY A.5 (eprint 2013/722 sect. 4.3). Tp = (X,Y,Z) Fp2 homogeneous projective.
        Returns new Tp and (r0,r1,r2)."""
        T = self.T
        X, Y, Z = Tp
        half = T.inv(2)
        A = T.f2_muls(T.f2_mul(X, Y), half)
        B = T.f2_sqr(Y)
        Cc = T.f2_sqr(Z)
        D = T.f2_muls(Cc, 3)
        E = T.f2_mul(D, self.C.b2)
        F = T.f2_muls(E, 3)
        G = T.f2_muls(T.f2_add(B, F), half)
        H = T.f2_sub(T.f2_sqr(T.f2_add(Y, Z)), T.f2_add(B, Cc))
        I = T.f2_sub(E, B)
        J = T.f2_sqr(X)
        K = T.f2_muls(T.f2_sqr(E), 3)
        X3 = T.f2_mul(T.f2_sub(B, F), A)
        Y3 = T.f2_sub(T.f2_sqr(G), K)
        Z3 = T.f2_mul(B, H)
        if self.P.twist == 'M':
            line = (I, T.f2_muls(J, 3), T.f2_neg(H))
        else:
            line = (T.f2_neg(H), T.f2_muls(J, 3), I)
        return (X3, Y3, Z3), line

    def _add_step(self, Tp, Q, update=True):
        T = self.T
        X, Y, Z = Tp
        xq, yq = Q
        O = T.f2_sub(Y, T.f2_mul(yq, Z))
        L = T.f2_sub(X, T.f2_mul(xq, Z))
        J = T.f2_sub(T.f2_mul(xq, O), T.f2_mul(L, yq))
        if update:
            Cc = T.f2_sqr(O)
            D = T.f2_sqr(L)
            E = T.f2_mul(L, D)
            F = T.f2_mul(Z, Cc)
            G = T.f2_mul(X, D)
            H = T.f2_sub(T.f2_add(E, F), T.f2_muls(G, 2))
            X3 = T.f2_mul(L, H)
            Y3 = T.f2_sub(T.f2_mul(T.f2_sub(G, H), O), T.f2_mul(Y, E))
            Z3 = T.f2_mul(E, Z)
            Tp = (X3, Y3, Z3)
        if self.P.twist == 'M':
            line = (J, T.f2_neg(O), L)
        else:
            line = (L, T.f2_neg(O), J)
        return Tp, line

    def _line_eval(self, line, Ppt):
        """M-type '014': C0.B0=r0, C0.B1=r1*xP, C1.B1=r2*yP  -> w^0, w^2, w^3
           D-type '034': C0.B0=r0*yP, C1.B0=r1*xP, C1.B1=r2  -> w^0, w^1, w^3"""
        T = self.T
        r0, r1, r2 = line
        xP, yP = Ppt
        if self.P.twist == 'M':
            return self._place({0: r0, 2: T.f2_muls(r1, xP), 3: T.f2_muls(r2, yP)})
        return self._place({0: T.f2_muls(r0, yP), 1: T.f2_muls(r1, xP), 3: r2})

    def miller_projective(self, pairs):
        T, C = self.T, self.C
        pairs = [(P, Q) for (P, Q) in pairs if P is not None and Q is not None]
        f = T.f12_one
        if not pairs:
            return f
        Ts = [(Q[0], Q[1], (1, 0)) for (_, Q) in pairs]
        digits = self.loop_digits
        for i in range(len(digits) - 2, -1, -1):
            f = T.f12_sqr(f)
            for k, (P, Q) in enumerate(pairs):
                Ts[k], ln = self._double_step(Ts[k])
                f = T.f12_mul(f, self._line_eval(ln, P))
                d = digits[i]
                if d:
                    A = Q if d == 1 else C.g2_neg(Q)
                    Ts[k], ln = self._add_step(Ts[k], A)
                    f = T.f12_mul(f, self._line_eval(ln, P))
        if self.P.family == 'bn':
            for k, (P, Q) in enumerate(pairs):
                Q1, Q2n = self.bn_frobenius_points(Q)
                Ts[k], ln = self._add_step(Ts[k], Q1)
                f = T.f12_mul(f, self._line_eval(ln, P))
                _, ln = self._add_step(Ts[k], Q2n, update=False)
                f = T.f12_mul(f, self._line_eval(ln, P))
        if self.loop_neg:
            f = T.f12_conj(f)
        return f

    # ------------------------------------------------------------------ final exponentiation
    def final_exp_plain(self, f):
        return self.T.f12_pow(f, self.full_exp)

    def final_exp(self, f):
        T = self.T
        t = T.f12_mul(T.f12_conj(f), T.f12_inv(f))         # f^(p^6-1)
        t = T.f12_mul(T.f12_frob(t, 2), t)                 # ^(p^2+1)
        return T.f12_pow(t, self.hard_exp)

    # ------------------------------------------------------------------ driver-level API
    def pairing(self, Q, P, semantics='gurvy'):
        """driver.Curve.Pairing(G2, G1) (reference driver/math.go:51)."""
        f = self.miller_projective([(P, Q)])
        return self.final_exp(f) if semantics == 'kilic' else f

    def pairing2(self, Qa, Qb, Pa, Pb, semantics='gurvy'):
        """driver.Curve.Pairing2(p2a, p2b, p1a, p1b) (reference driver/math.go:54)."""
        f = self.miller_projective([(Pa, Qa), (Pb, Qb)])
        return self.final_exp(f) if semantics == 'kilic' else f

    def fexp(self, f, semantics='gurvy'):
        """driver.Curve.FExp (reference driver/math.go:57; identity for kilic: kilic/bls12-381.go:279-281)."""
        return f if semantics == 'kilic' else self.final_exp(f)
