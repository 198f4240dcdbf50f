"""Byte formats of Zr / G1 / G2 / Gt (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates SURVEY A.3: what ``Bytes()`` / ``Compressed()`` return in the reference adapters
(bn254.go:76-86,216-220; bls12-381.go:286-296,432-436; kilic/bls12-381.go:74-85,224-231;
common/big.go:101-113) -- gnark ``RawBytes``/``Bytes`` and the ZCash format kilic uses.
Kilic == gnark on BLS12-381 is pinned by reference math_test.go:879-945.
"""


def fp_to_bytes(P, a):
    return int(a % P.p).to_bytes(P.fp_bytes, 'big')


def fp_from_bytes(P, b):
    return int.from_bytes(b, 'big')


def zr_to_bytes(P, k):
    """32-byte big-endian of k mod r (common/big.go:101-113; the group order itself encodes as r)."""
    return int(k % P.r).to_bytes(32, 'big')


def _mask(P):
    return {2: 0xC0, 3: 0xE0}[P.flag_bits]


def _flags(P, kind):
    if P.flag_bits == 3:      # BLS12: ZCash flags
        return {'unc': 0x00, 'unc_inf': 0x40, 'small': 0x80, 'large': 0xA0, 'cmp_inf': 0xC0}[kind]
    return {'unc': 0x00, 'unc_inf': 0x00, 'small': 0x80, 'large': 0xC0, 'cmp_inf': 0x40}[kind]


def _fp_lex_largest(P, y):
    return y > (P.p - 1) // 2


def _fp2_lex_largest(P, y):
    return _fp_lex_largest(P, y[1]) if y[1] != 0 else _fp_lex_largest(P, y[0])


def g1_to_bytes(P, pt):
    """uncompressed: X || Y big-endian; infinity = flag byte then zeros."""
    n = P.fp_bytes
    if pt is None:
        out = bytearray(2 * n)
        out[0] |= _flags(P, 'unc_inf')
        return bytes(out)
    return fp_to_bytes(P, pt[0]) + fp_to_bytes(P, pt[1])


def g1_to_compressed(P, pt):
    n = P.fp_bytes
    if pt is None:
        out = bytearray(n)
        out[0] |= _flags(P, 'cmp_inf')
        return bytes(out)
    out = bytearray(fp_to_bytes(P, pt[0]))
    out[0] |= _flags(P, 'large' if _fp_lex_largest(P, pt[1]) else 'small')
    return bytes(out)


def g2_to_bytes(P, pt):
    """X.A1 || X.A0 || Y.A1 || Y.A0"""
    n = P.fp_bytes
    if pt is None:
        out = bytearray(4 * n)
        out[0] |= _flags(P, 'unc_inf')
        return bytes(out)
    (x0, x1), (y0, y1) = pt
    return fp_to_bytes(P, x1) + fp_to_bytes(P, x0) + fp_to_bytes(P, y1) + fp_to_bytes(P, y0)


def g2_to_compressed(P, pt):
    n = P.fp_bytes
    if pt is None:
        out = bytearray(2 * n)
        out[0] |= _flags(P, 'cmp_inf')
        return bytes(out)
    (x0, x1), y = pt
    out = bytearray(fp_to_bytes(P, x1) + fp_to_bytes(P, x0))
    out[0] |= _flags(P, 'large' if _fp2_lex_largest(P, y) else 'small')
    return bytes(out)


def g1_from_bytes(P, b):
    n = P.fp_bytes
    assert len(b) == 2 * n
    flag = b[0] & _mask(P)
    if P.flag_bits == 3 and flag == 0x40:
        return None
    raw = bytes([b[0] & ~_mask(P) & 0xFF]) + bytes(b[1:])
    x = int.from_bytes(raw[:n], 'big')
    y = int.from_bytes(raw[n:], 'big')
    if P.flag_bits == 2 and x == 0 and y == 0:
        return None
    return (x, y)


def g2_from_bytes(P, b):
    n = P.fp_bytes
    assert len(b) == 4 * n
    flag = b[0] & _mask(P)
    if P.flag_bits == 3 and flag == 0x40:
        return None
    raw = bytes([b[0] & ~_mask(P) & 0xFF]) + bytes(b[1:])
    v = [int.from_bytes(raw[i * n:(i + 1) * n], 'big') for i in range(4)]
    if P.flag_bits == 2 and not any(v):
        return None
    return ((v[1], v[0]), (v[3], v[2]))


GT_ORDER = [(1, 2, 1), (1, 2, 0), (1, 1, 1), (1, 1, 0), (1, 0, 1), (1, 0, 0),
            (0, 2, 1), (0, 2, 0), (0, 1, 1), (0, 1, 0), (0, 0, 1), (0, 0, 0)]


def gt_to_bytes(P, f):
    """C1.B2.A1, C1.B2.A0, ..., C0.B0.A1, C0.B0.A0 -- highest coefficient first at every level."""
    return b''.join(fp_to_bytes(P, f[c][b][a]) for (c, b, a) in GT_ORDER)


def gt_from_bytes(P, raw):
    n = P.fp_bytes
    assert len(raw) == 12 * n
    vals = {}
    for i, key in enumerate(GT_ORDER):
        vals[key] = int.from_bytes(raw[i * n:(i + 1) * n], 'big')
    return tuple(tuple((vals[(c, b, 0)], vals[(c, b, 1)]) for b in range(3)) for c in range(2))
