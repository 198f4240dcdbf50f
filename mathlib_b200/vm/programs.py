"""Straight-line Fp2-level programs of the pairing VM, written against compiler.py's DSL.

Fp12 elements are held in the w-basis: g_k (k = 0..5) is the Fp2 coefficient of w^k, w^6 = xi, i.e.
g0..g5 = C0.B0, C1.B0, C0.B1, C1.B1, C0.B2, C1.B2 of the gnark/kilic E12 layout.

Slot maps
  Miller loop: F double buffer at base registers B1 (current) / B2 (next), 6 slots each (absolute 0..5 / 6..11);
               pair k: T = (X,Y,Z) at 12+3k, Q = (x, y, -y) at 18+3k (y is addressed through B3 = 0 / 1 to pick y / -y
               for the signed NAF digits of BN254), P at 24+k (c0 = xP, c1 = yP); temporaries 26..47.
  Final exp.:  NREGS Fp12 registers at 0,6,12,...; temporaries after them; programs take (B1,B2,B3) = (dst, a, b) bases.
The formulas are the ones of csrc/pairing.cuh (SURVEY A.5 doubling / addition steps, sparse line slots, Granger-Scott
squaring); tests/test_vm_programs.py checks every compiled program against the oracle.
"""
from .compiler import (ref, const, dot, lin, inv, mul, sqr, compile_program, set_field, NEG, CONJ, XI, DBL, REAL0,
                       REAL1, C_ABS, C_B1, C_B2, C_B3, C_CONST)

T_BASE, Q_BASE = 12, 18
# per-curve slot file: BLS12 curves (96-byte slots): 48 slots, 5 Fp12 registers for the final exponentiation;
# BN254 (64-byte slots): 72 slots, 8 registers (the Fuentes-Castaneda chain keeps more values alive).  Both are 4,608 B.
SLOTCFG = {'BLS381': (36, 5), 'BLS377': (36, 5), 'BN254': (54, 8)}
_cfg = {'nslots': 36, 'nregs': 5, 'qstride': 2}


def q_stride():
    """slots per Q: (x, y) -- plus -y for BN254, whose signed NAF digits subtract Q"""
    return _cfg['qstride']


def p_base():
    return Q_BASE + 2 * q_stride()


def miller_temp0():
    return p_base() + 2


def miller_temps():
    return list(range(miller_temp0(), _cfg['nslots']))


def fexp_temps():
    return list(range(6 * _cfg['nregs'], _cfg['nslots']))

# constant bank indices
K_ONE, K_BTW, K_ZERO = 0, 1, 2
K_FROB = 3          # K_FROB + 5*(k-1) + (i-1), k = 1..3, i = 1..5


def k_frob(k, i):
    return K_FROB + 5 * (k - 1) + (i - 1)


class Curve:
    def __init__(self, name, twist, family, btw_is_4xi):
        self.name, self.twist, self.family, self.btw_is_4xi = name, twist, family, btw_is_4xi
        self.supp = (0, 2, 3) if twist == 'M' else (0, 1, 3)     # powers of w carrying the line coefficients


# name -> (p, 32-bit limbs, u^2, lazy operand modifiers)
FIELDS = {
    'BN254': (0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47, 8, -1, False),
    'BLS381': (0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 12, -1, True),
    'BLS377': (0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001, 12, -5, False),
}

CURVES = {
    'BN254': Curve('BN254', 'D', 'bn', False),
    'BLS381': Curve('BLS381', 'M', 'bls12', True),
    'BLS377': Curve('BLS377', 'D', 'bls12', False),
}


# ------------------------------------------------------------------------------------------------ Fp12 building blocks
def f12_mul_nodes(a, b):
    out = []
    for k in range(6):
        terms = []
        for i in range(6):
            j = (k - i) % 6
            am = XI if i + j >= 6 else 0
            terms.append(((a[i], am), b[j]))
        out.append(dot(terms, name='m%d' % k))
    return out


def f12_sqr_nodes(a):
    out = []
    for k in range(6):
        terms = []
        for i in range(6):
            j = (k - i) % 6
            if i > j:
                continue
            am = (XI if i + j >= 6 else 0) | (DBL if i != j else 0)
            terms.append(((a[i], am), a[j]))
        out.append(dot(terms, name='s%d' % k))
    return out


def sparse_mul_nodes(a, line, supp, pred, alt):
    """a * (sum_s line[s] w^s); line: dict power -> Val"""
    out = []
    for k in range(6):
        terms = []
        for s in supp:
            i = (k - s) % 6
            am = XI if k - s < 0 else 0
            terms.append(((a[i], am), line[s]))
        out.append(dot(terms, pred=pred, alt=alt[k], name='sp%d' % k))
    return out


def cyclo_sqr_nodes(g):
    """Granger-Scott; Fp4 pairs (g0,g3), (g1,g4), (g2,g5)."""
    o = [None] * 6
    # g0' = 3(g0^2 + xi g3^2) - 2 g0 ; g3' = 3*(2 g0 g3) + 2 g3
    o[0] = dot([(g[0], g[0]), ((g[3], XI), g[3])], scale=3, lin=[(g[0], -2)])
    o[3] = dot([((g[0], DBL), g[3])], scale=3, lin=[(g[3], 2)])
    # g1' = 3 xi (2 g2 g5) + 2 g1 ; g4' = 3(g2^2 + xi g5^2) - 2 g4
    o[1] = dot([((g[2], XI | DBL), g[5])], scale=3, lin=[(g[1], 2)])
    o[4] = dot([(g[2], g[2]), ((g[5], XI), g[5])], scale=3, lin=[(g[4], -2)])
    # g2' = 3(g1^2 + xi g4^2) - 2 g2 ; g5' = 3*(2 g1 g4) + 2 g5
    o[2] = dot([(g[1], g[1]), ((g[4], XI), g[4])], scale=3, lin=[(g[2], -2)])
    o[5] = dot([((g[1], DBL), g[4])], scale=3, lin=[(g[5], 2)])
    return o


def f12_inv_nodes(g):
    x = [g[0], g[2], g[4]]          # C0 = x0 + x1 v + x2 v^2
    y = [g[1], g[3], g[5]]          # C1
    # d = C0^2 - v C1^2
    d0 = dot([(x[0], x[0]), ((x[1], XI | DBL), x[2]), ((y[0], XI | DBL | NEG), y[2]), ((y[1], XI | NEG), y[1])])
    d1 = dot([((x[0], DBL), x[1]), ((x[2], XI), x[2]), ((y[0], NEG), y[0]), ((y[1], XI | DBL | NEG), y[2])])
    d2 = dot([((x[0], DBL), x[2]), (x[1], x[1]), ((y[0], DBL | NEG), y[1]), ((y[2], XI | NEG), y[2])])
    t0 = dot([(d0, d0), ((d1, XI | NEG), d2)])
    t1 = dot([((d2, XI), d2), ((d0, NEG), d1)])
    t2 = dot([(d1, d1), ((d0, NEG), d2)])
    n = dot([(d0, t0), ((d2, XI), t1), ((d1, XI), t2)], name='norm')
    ni = inv(n)
    e = [mul(t0, ni), mul(t1, ni), mul(t2, ni)]

    def f6mul(u, v, neg):
        m = NEG if neg else 0
        z0 = dot([((u[0], m), v[0]), ((u[1], XI | m), v[2]), ((u[2], XI | m), v[1])])
        z1 = dot([((u[0], m), v[1]), ((u[1], m), v[0]), ((u[2], XI | m), v[2])])
        z2 = dot([((u[0], m), v[2]), ((u[1], m), v[1]), ((u[2], m), v[0])])
        return [z0, z1, z2]
    c0 = f6mul(x, e, False)
    c1 = f6mul(y, e, True)
    return [c0[0], c1[0], c0[1], c1[1], c0[2], c1[2]], n, ni


def set_group(nodes, g):
    """assign scheduling group g to every not-yet-grouped non-reference node reachable from `nodes`"""
    seen = set()

    def visit(v):
        if v.id in seen or v.kind == 'ref':
            return
        seen.add(v.id)
        if v.group == 0:
            v.group = g
        for o in v.operands():
            visit(o)
    for n in nodes:
        visit(n)


def regs(cls, base=0):
    return [ref(cls, base + k) for k in range(6)]


def bind(nodes, cls, base=0):
    return [(n, (cls, base + k)) for k, n in enumerate(nodes)]


# ------------------------------------------------------------------------------------------------ final-exp programs
def prog_f12_mul():
    return compile_program('F12_MUL', bind(f12_mul_nodes(regs(C_B2), regs(C_B3)), C_B1), fexp_temps())


def prog_f12_sqr():
    """generic Fp12 squaring (Gt.Exp on elements that need not be cyclotomic)"""
    return compile_program('F12_SQR', bind(f12_sqr_nodes(regs(C_B2)), C_B1), fexp_temps())


def prog_f12_mulp():
    """B1 = live[0] ? B2 * B3 : B2   -- the conditional multiply of a square-and-multiply ladder whose exponent bit
    differs between the groups of a warp (the bit is handed to the interpreter as the group's `live` flag)"""
    a, b = regs(C_B2), regs(C_B3)
    nodes = f12_mul_nodes(a, b)
    for k, n in enumerate(nodes):
        n.pred, n.alt = 1, a[k]
    return compile_program('F12_MULP', bind(nodes, C_B1), fexp_temps())


def prog_gt_one():
    one, zero = const(K_ONE), const(K_ZERO)
    return compile_program('GT_ONE', [(lin([(one if k == 0 else zero, 1)]), (C_B1, k)) for k in range(6)], fexp_temps())


def prog_cyclo():
    return compile_program('CYCLO_SQR', bind(cyclo_sqr_nodes(regs(C_B2)), C_B1), fexp_temps())


def prog_conj():
    a = regs(C_B2)
    return compile_program('CONJ', bind([lin([((a[k], NEG if k & 1 else 0), 1)]) for k in range(6)], C_B1), fexp_temps())


def prog_frob(k):
    a = regs(C_B2)
    cm = CONJ if k & 1 else 0
    nodes = [lin([((a[0], cm), 1)])]
    for i in range(1, 6):
        nodes.append(dot([((a[i], cm), const(k_frob(k, i)))]))
    return compile_program('FROB%d' % k, bind(nodes, C_B1), fexp_temps())


def prog_inv():
    nodes, n, ni = f12_inv_nodes(regs(C_B2))
    # the norm and its inverse are parked in two slots of the destination register (written for real only in the last
    # phase), which keeps the program within 6 temporaries
    outs = [(n, (C_B1, 4)), (ni, (C_B1, 5))] + bind(nodes, C_B1)
    return compile_program('F12_INV', outs, fexp_temps())


# ------------------------------------------------------------------------------------------------ Miller-loop programs
def _line_and_sparse(cv, f_cur, lines_per_pair, f_slots_cycle):
    """apply one sparse multiplication per pair, ping-ponging between the two f buffers.
    f_slots_cycle: list of classes the successive results go to."""
    outs = []
    for k, line in enumerate(lines_per_pair):
        cls = f_slots_cycle[k]
        nodes = sparse_mul_nodes(f_cur, line, cv.supp, pred=k + 1, alt=f_cur)
        outs += bind(nodes, cls)
        f_cur = nodes
    return outs, f_cur


def _double_pair(cv, k, raw=False):
    X, Y, Z = ref(C_ABS, T_BASE + 3 * k), ref(C_ABS, T_BASE + 3 * k + 1), ref(C_ABS, T_BASE + 3 * k + 2)
    P = ref(C_ABS, p_base() + k)
    XY = mul(X, Y, name='XY')
    B = sqr(Y, name='B')
    H = dot([((Y, DBL), Z)], name='H')
    J = sqr(X, name='J')
    # E = 3 b' Z^2 straight out of a multiplier op (scale applied after the reduction): no separate linear phase
    if cv.btw_is_4xi:
        E = dot([((Z, XI | DBL), Z)], scale=6, name='E')            # b' = 4 xi: 3 b' Z^2 = 6 * (2 xi Z) * Z
    else:
        C = sqr(Z, name='C')
        E = dot([(C, const(K_BTW))], scale=3, name='E')
    F3 = lin([(E, 3)], name='F3')
    Gv = lin([(B, 1), (E, 3)], halve=True, name='G')
    BmF = lin([(B, 1), (E, -3)], name='BmF')
    I = lin([(E, 1), (B, -1)], name='I')
    X3 = dot([(XY, BmF)], halve=True, name='X3')
    Y3 = dot([(Gv, Gv), (F3, (E, NEG))], name='Y3')
    Z3 = mul(B, H, name='Z3')
    if cv.twist == 'M':      # (r0,r1,r2) = (I, 3J, -H): w^0 = r0, w^2 = r1 xP, w^3 = r2 yP
        line = {0: I, 2: dot([(J, (P, REAL0))], scale=3), 3: dot([((H, NEG), (P, REAL1))])}
    else:                    # (r0,r1,r2) = (-H, 3J, I): w^0 = r0 yP, w^1 = r1 xP, w^3 = r2
        line = {0: dot([((H, NEG), (P, REAL1))]), 1: dot([(J, (P, REAL0))], scale=3), 3: I}
    tout = [(X3, (C_ABS, T_BASE + 3 * k)), (Y3, (C_ABS, T_BASE + 3 * k + 1)), (Z3, (C_ABS, T_BASE + 3 * k + 2))]
    if raw:          # the P-independent coefficients (r0, r1, r2) of the comments above, for the fixed-Q line tables
        r = (I, lin([(J, 3)]), lin([(H, -1)])) if cv.twist == 'M' else (lin([(H, -1)]), lin([(J, 3)]), I)
        return tout, r
    return tout, line


def _add_pair(cv, k, qx, qy, update=True, raw=False):
    X, Y, Z = ref(C_ABS, T_BASE + 3 * k), ref(C_ABS, T_BASE + 3 * k + 1), ref(C_ABS, T_BASE + 3 * k + 2)
    P = ref(C_ABS, p_base() + k)
    O = dot([((qy, NEG), Z)], lin=[(Y, 1)], name='O')
    L = dot([((qx, NEG), Z)], lin=[(X, 1)], name='L')
    J = dot([(qx, O), ((L, NEG), qy)], name='J')
    tout = []
    if update:
        Cc = sqr(O)
        D = sqr(L)
        E = mul(L, D)
        Fv = mul(Z, Cc)
        Gv = mul(X, D)
        H = lin([(E, 1), (Fv, 1), (Gv, -2)])
        GmH = lin([(Gv, 3), (E, -1), (Fv, -1)])          # = Gv - H, written on H's inputs: same phase as H
        X3 = mul(L, H)
        Y3 = dot([(GmH, O), ((Y, NEG), E)])
        Z3 = mul(E, Z)
        tout = [(X3, (C_ABS, T_BASE + 3 * k)), (Y3, (C_ABS, T_BASE + 3 * k + 1)), (Z3, (C_ABS, T_BASE + 3 * k + 2))]
    if raw:
        r = (J, lin([(O, -1)]), lin([(L, 1)])) if cv.twist == 'M' else (lin([(L, 1)]), lin([(O, -1)]), J)
        return tout, r
    if cv.twist == 'M':      # (J, -O, L)
        line = {0: J, 2: dot([((O, NEG), (P, REAL0))]), 3: dot([(L, (P, REAL1))])}
    else:                    # (L, -O, J)
        line = {0: dot([(L, (P, REAL1))]), 1: dot([((O, NEG), (P, REAL0))]), 3: J}
    return tout, line


def _f_cycle(n):
    """buffers the successive f results are written to, starting from a read of B1"""
    return [C_B2 if i % 2 == 0 else C_B1 for i in range(n)]


def prog_dbl(cv, np_):
    f = regs(C_B1)
    f2 = f12_sqr_nodes(f)
    set_group(f2, 1)
    cyc = _f_cycle(1 + np_)
    outs = bind(f2, cyc[0])
    f_cur = f2
    for k in range(np_):
        t, line = _double_pair(cv, k)
        nodes = sparse_mul_nodes(f_cur, line, cv.supp, pred=k + 1, alt=f_cur)
        set_group([v for v, _ in t] + nodes, 2 + k)       # one pair at a time: bounds the live temporaries
        outs += t
        outs += bind(nodes, cyc[1 + k])
        f_cur = nodes
    return compile_program('DBL%d' % np_, outs, miller_temps())


def prog_add(cv, np_):
    f = regs(C_B1)
    outs = []
    cyc = _f_cycle(np_)
    f_cur = f
    for k in range(np_):
        qx = ref(C_ABS, Q_BASE + q_stride() * k)
        qy = ref(C_B3, Q_BASE + q_stride() * k + 1)
        t, line = _add_pair(cv, k, qx, qy)
        nodes = sparse_mul_nodes(f_cur, line, cv.supp, pred=k + 1, alt=f_cur)
        set_group([v for v, _ in t] + nodes, 1 + k)
        outs += t
        outs += bind(nodes, cyc[k])
        f_cur = nodes
    return compile_program('ADD%d' % np_, outs, miller_temps())


def prog_bn_tail(cv, np_):
    """f *= l_{T,pi(Q)}; T += pi(Q); f *= l_{T,-pi^2(Q)}   """
    f = regs(C_B1)
    f_cur = f
    outs = []
    cyc = _f_cycle(2 * np_)
    ci = 0
    for k in range(np_):
        qx, qy = ref(C_ABS, Q_BASE + q_stride() * k), ref(C_ABS, Q_BASE + q_stride() * k + 1)
        q1x = dot([((qx, CONJ), const(k_frob(1, 2)))])
        q1y = dot([((qy, CONJ), const(k_frob(1, 3)))])
        q2x = dot([(qx, const(k_frob(2, 2)))])
        X, Y, Z = ref(C_ABS, T_BASE + 3 * k), ref(C_ABS, T_BASE + 3 * k + 1), ref(C_ABS, T_BASE + 3 * k + 2)
        P = ref(C_ABS, p_base() + k)
        # first step (with update), written out here because the second step needs the NEW T as DAG nodes
        O = dot([((q1y, NEG), Z)], lin=[(Y, 1)])
        L = dot([((q1x, NEG), Z)], lin=[(X, 1)])
        J = dot([(q1x, O), ((L, NEG), q1y)])
        Cc, D = sqr(O), sqr(L)
        E, Fv, Gv = mul(L, D), mul(Z, Cc), mul(X, D)
        H = lin([(E, 1), (Fv, 1), (Gv, -2)])
        GmH = lin([(Gv, 3), (E, -1), (Fv, -1)])          # = Gv - H, written on H's inputs: same phase as H
        X3, Y3, Z3 = mul(L, H), dot([(GmH, O), ((Y, NEG), E)]), mul(E, Z)
        line1 = {0: dot([(L, (P, REAL1))]), 1: dot([((O, NEG), (P, REAL0))]), 3: J}
        # second step, line only, through (X3,Y3,Z3) and (q2x, qy)
        O2 = dot([((qy, NEG), Z3)], lin=[(Y3, 1)])
        L2 = dot([((q2x, NEG), Z3)], lin=[(X3, 1)])
        J2 = dot([(q2x, O2), ((L2, NEG), qy)])
        line2 = {0: dot([(L2, (P, REAL1))]), 1: dot([((O2, NEG), (P, REAL0))]), 3: J2}
        for line in (line1, line2):
            nodes = sparse_mul_nodes(f_cur, line, cv.supp, pred=k + 1, alt=f_cur)
            set_group(nodes, 1 + k)
            outs += bind(nodes, cyc[ci])
            ci += 1
            f_cur = nodes
    return compile_program('BNTAIL%d' % np_, outs, miller_temps())


# ------------------------------------------------------------------------------------------------ fixed-Q programs
# when the G2 arguments are fixed (public keys, the generator) the G2 side of every Miller step is
# computed once into a LINE TABLE: per step three P-independent Fp2 coefficients (r0, r1, r2).  At run time the driver
# copies a step's coefficients of pair k into the pair's T slots and the programs below evaluate them at P and fold them
# into f -- same values as the DBL / ADD / BNTAIL programs, without any G2 arithmetic.
LINE_OUT = T_BASE + 3          # precompute programs work on pair 0 and leave (r0, r1, r2) in pair 1's T slots


def _fixed_line(cv, k):
    r0, r1, r2 = (ref(C_ABS, T_BASE + 3 * k + j) for j in range(3))
    P = ref(C_ABS, p_base() + k)
    if cv.twist == 'M':      # w^0 = r0, w^2 = r1 xP, w^3 = r2 yP
        return {0: r0, 2: dot([(r1, (P, REAL0))]), 3: dot([(r2, (P, REAL1))])}
    return {0: dot([(r0, (P, REAL1))]), 1: dot([(r1, (P, REAL0))]), 3: r2}


def prog_fixed(cv, np_, square):
    """f <- [f^2] * prod_k line_k(P_k)   (SQRLINE = the DBL programs' counterpart, LINE = ADD's / each BN tail line)"""
    f = regs(C_B1)
    outs = []
    if square:
        f_cur = f12_sqr_nodes(f)
        set_group(f_cur, 1)
        cyc = _f_cycle(1 + np_)
        outs += bind(f_cur, cyc[0])
        cyc = cyc[1:]
    else:
        f_cur = f
        cyc = _f_cycle(np_)
    lines = [_fixed_line(cv, k) for k in range(np_)]
    for k in range(np_):
        set_group([v for v in lines[k].values() if v.kind != 'ref'], 2)        # all P-scalings share one phase
    for k in range(np_):
        nodes = sparse_mul_nodes(f_cur, lines[k], cv.supp, pred=k + 1, alt=f_cur)
        set_group(nodes, 3 + k)
        outs += bind(nodes, cyc[k])
        f_cur = nodes
    return compile_program(('SQRLINE%d' if square else 'LINE%d') % np_, outs, miller_temps())


def _bind_raw(r):
    outs = []
    for j, v in enumerate(r):
        if v.out is not None or v.kind == 'ref':
            v = lin([(v, 1)])
        outs.append((v, (C_ABS, LINE_OUT + j)))
    return outs


def prog_pre_dbl(cv):
    t, r = _double_pair(cv, 0, raw=True)
    return compile_program('PRE_DBL', t + _bind_raw(r), miller_temps())


def prog_pre_add(cv):
    qx, qy = ref(C_ABS, Q_BASE), ref(C_B3, Q_BASE + 1)
    t, r = _add_pair(cv, 0, qx, qy, raw=True)
    return compile_program('PRE_ADD', t + _bind_raw(r), miller_temps())


def prog_pre_tail(cv, second):
    """BN254 tail: line through T and pi(Q) (with T += pi(Q)), then the line through the new T and -pi^2(Q)"""
    qx, qy = ref(C_ABS, Q_BASE), ref(C_ABS, Q_BASE + 1)
    if not second:
        q1x = dot([((qx, CONJ), const(k_frob(1, 2)))])
        q1y = dot([((qy, CONJ), const(k_frob(1, 3)))])
        t, r = _add_pair(cv, 0, q1x, q1y, raw=True)
        return compile_program('PRE_TAIL1', t + _bind_raw(r), miller_temps())
    q2x = dot([(qx, const(k_frob(2, 2)))])
    t, r = _add_pair(cv, 0, q2x, qy, update=False, raw=True)
    return compile_program('PRE_TAIL2', _bind_raw(r), miller_temps())


def prog_init(np_):
    """f = 1 (into B1), Z_k = 1, -y_k"""
    one, zero = const(K_ONE), const(K_ZERO)
    outs = [(lin([(one if k == 0 else zero, 1)]), (C_B1, k)) for k in range(6)]
    for k in range(np_):
        outs.append((lin([(one, 1)]), (C_ABS, T_BASE + 3 * k + 2)))
        outs.append((lin([(ref(C_ABS, Q_BASE + q_stride() * k), 1)]), (C_ABS, T_BASE + 3 * k)))
        outs.append((lin([(ref(C_ABS, Q_BASE + q_stride() * k + 1), 1)]), (C_ABS, T_BASE + 3 * k + 1)))
        if q_stride() == 3:
            outs.append((lin([((ref(C_ABS, Q_BASE + q_stride() * k + 1), NEG), 1)]), (C_ABS, Q_BASE + q_stride() * k + 2)))
    return compile_program('INIT%d' % np_, outs, miller_temps())


def build_all(curve_name):
    """name -> Program for one curve"""
    cv = CURVES[curve_name]
    _cfg['nslots'], _cfg['nregs'] = SLOTCFG[curve_name]
    _cfg['qstride'] = 3 if cv.family == 'bn' else 2
    # field parameters for the lazy-reduction bounds (compiler.dot_bounds).  BLS12-381 (xi = 1 + u, three spare bits above p)
    # runs the interpreter's unreduced operand modifiers (csrc/vm.cuh lazy_mods, curves.cuh LAZY_MODS); the other two curves
    # keep canonical modifiers
    set_field(*FIELDS[curve_name])
    progs = {}
    for np_ in (1, 2):
        progs['INIT%d' % np_] = prog_init(np_)
        progs['DBL%d' % np_] = prog_dbl(cv, np_)
        progs['ADD%d' % np_] = prog_add(cv, np_)
        if cv.family == 'bn':
            progs['BNTAIL%d' % np_] = prog_bn_tail(cv, np_)
    progs['F12_MUL'] = prog_f12_mul()
    progs['CYCLO_SQR'] = prog_cyclo()
    progs['CONJ'] = prog_conj()
    for k in (1, 2, 3):
        progs['FROB%d' % k] = prog_frob(k)
    progs['F12_INV'] = prog_inv()
    progs['F12_SQR'] = prog_f12_sqr()
    progs['F12_MULP'] = prog_f12_mulp()
    progs['GT_ONE'] = prog_gt_one()
    for np_ in (1, 2):
        progs['SQRLINE%d' % np_] = prog_fixed(cv, np_, True)
        progs['LINE%d' % np_] = prog_fixed(cv, np_, False)
    progs['PRE_DBL'] = prog_pre_dbl(cv)
    progs['PRE_ADD'] = prog_pre_add(cv)
    if cv.family == 'bn':
        progs['PRE_TAIL1'] = prog_pre_tail(cv, False)
        progs['PRE_TAIL2'] = prog_pre_tail(cv, True)
    return progs
