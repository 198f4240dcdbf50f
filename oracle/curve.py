"""G1 / G2 affine group law, scalar multiplication, MSM (TEST INFRASTRUCTURE).

Semantics restated (outputs are canonical group elements, so any correct algorithm is
byte-exact -- SURVEY A.6):
* ``G1.Mul``   reference bn254.go:49-54, bls12-381.go:238-247, kilic/bls12-381.go:40-50
* ``G1.Mul2``  reference bn254.go:56-62 (two Mul + Add), bls12-381.go:249-265 + 869-937
               (Strauss-Shamir: restated op-for-op in ``joint_scalar_mul``)
* ``MultiScalarMul`` reference bn254.go:232-245, bls12-381.go:766-783, kilic/bls12-381.go:247-254
Points are ``None`` (infinity) or ``(x, y)``; G2 coordinates are Fp2 tuples.
"""
from .tower import Tower


class Curve:
    def __init__(self, params):
        self.P = params
        self.T = Tower(params)
        self.p = params.p
        self.r = params.r
        T = self.T
        if params.twist == 'D':
            self.b2 = T.f2_mul((params.b, 0), T.f2_inv(T.xi))
        else:
            self.b2 = T.f2_mul((params.b, 0), T.xi)
        self.g1 = params.g1
        self.g2 = params.g2

    # ---------------- G1 ----------------
    def g1_on_curve(self, P):
        if P is None:
            return True
        x, y = P
        return (y * y - x * x * x - self.P.b) % self.p == 0

    def g1_neg(self, P):
        return None if P is None else (P[0], -P[1] % self.p)

    def g1_add(self, P, Q):
        p = self.p
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return (x3, (lam * (x1 - x3) - y1) % p)

    def g1_mul(self, P, k):
        """[k mod r]P -- the value every driver's G1.Mul returns for P in the r-torsion."""
        k %= self.r
        R = None
        for bit in bin(k)[2:] if k else '':
            R = self.g1_add(R, R)
            if bit == '1':
                R = self.g1_add(R, P)
        return R

    def g1_mul2(self, P, e, Q, f):
        return self.g1_add(self.g1_mul(P, e), self.g1_mul(Q, f))

    def joint_scalar_mul(self, a1, a2, s1, s2):
        """Op-for-op restatement of reference bls12-381.go:869-937 (Strauss-Shamir, 2-bit
        joint window, whole 64-bit words scanned from the top word down, negative
        scalars negate the base). Scalars are Python ints (may be negative / >= r
        exactly like the big.Int the reference receives)."""
        add, neg = self.g1_add, self.g1_neg
        table = [None] * 15
        k1, k2 = s1, s2
        if s1 < 0:
            k1 = -s1
            table[0] = neg(a1)
        else:
            table[0] = a1
        if s2 < 0:
            k2 = -s2
            table[3] = neg(a2)
        else:
            table[3] = a2
        table[1] = add(table[0], table[0])
        table[2] = add(table[1], table[0])
        table[4] = add(table[3], table[0])
        table[5] = add(table[3], table[1])
        table[6] = add(table[3], table[2])
        table[7] = add(table[3], table[3])
        table[8] = add(table[7], table[0])
        table[9] = add(table[7], table[1])
        table[10] = add(table[7], table[2])
        table[11] = add(table[7], table[3])
        table[12] = add(table[11], table[0])
        table[13] = add(table[11], table[1])
        table[14] = add(table[11], table[2])
        # fr.Element.SetBigInt reduces mod r; .Bits() gives the regular-form words
        w1 = k1 % self.r
        w2 = k2 % self.r
        max_bit = max(k1.bit_length(), k2.bit_length())
        hi_word = (max_bit - 1) // 64 if max_bit > 0 else -1
        # Go: (maxBit-1)/64 with maxBit==0 -> (-1)/64 == 0 (truncation), so word 0 is still scanned
        if max_bit == 0:
            hi_word = 0
        res = None
        for i in range(hi_word, -1, -1):
            for j in range(32):
                res = add(res, res)
                res = add(res, res)
                sh = 62 - 2 * j
                b1 = (w1 >> (64 * i + sh)) & 3
                b2 = (w2 >> (64 * i + sh)) & 3
                if b1 | b2:
                    res = add(res, table[(b2 << 2 | b1) - 1])
        return res

    def g1_msm(self, points, scalars):
        """sum_i [s_i mod r]P_i ; n==0 -> infinity; length mismatch -> infinity
        (gnark MultiExp errors are discarded: reference bn254.go:242)."""
        if len(points) != len(scalars):
            return None
        acc = None
        for P, s in zip(points, scalars):
            acc = self.g1_add(acc, self.g1_mul(P, s))
        return acc

    # ---------------- G2 (on the twist E'/Fp2) ----------------
    def g2_on_curve(self, Q):
        if Q is None:
            return True
        T = self.T
        x, y = Q
        return T.f2_sub(T.f2_sqr(y), T.f2_add(T.f2_mul(T.f2_sqr(x), x), self.b2)) == (0, 0)

    def g2_neg(self, Q):
        return None if Q is None else (Q[0], self.T.f2_neg(Q[1]))

    def g2_add(self, P, Q):
        T = self.T
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if T.f2_add(y1, y2) == (0, 0):
                return None
            lam = T.f2_mul(T.f2_muls(T.f2_sqr(x1), 3), T.f2_inv(T.f2_muls(y1, 2)))
        else:
            lam = T.f2_mul(T.f2_sub(y2, y1), T.f2_inv(T.f2_sub(x2, x1)))
        x3 = T.f2_sub(T.f2_sub(T.f2_sqr(lam), x1), x2)
        return (x3, T.f2_sub(T.f2_mul(lam, T.f2_sub(x1, x3)), y1))

    def g2_mul(self, Q, k):
        k %= self.r
        R = None
        for bit in bin(k)[2:] if k else '':
            R = self.g2_add(R, R)
            if bit == '1':
                R = self.g2_add(R, Q)
        return R
