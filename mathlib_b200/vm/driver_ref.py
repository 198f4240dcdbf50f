"""Reference control flow of the VM pairing kernel in Python (executes the compiled microcode with compiler.simulate).

csrc/pairing_vm.cuh follows exactly this sequence of program runs; the tests run it against the oracle, so a
microcode or sequencing bug is caught on the CPU.  Values here are canonical residues (the CUDA side holds the same
values in Montgomery form).
"""
from . import programs as PR
from .compiler import Field, simulate


class CurveCtx:
    def __init__(self, name, p, beta, xi, x_abs, x_neg, naf, btw, frob):
        self.name = name
        self.F = Field(p, beta, xi)
        self.p = p
        self.x_abs, self.x_neg, self.naf = x_abs, x_neg, naf
        self.cv = PR.CURVES[name]
        self.progs = PR.build_all(name)
        self.nslots, self.nregs = PR.SLOTCFG[name]
        self.qstride = 3 if self.cv.family == 'bn' else 2
        self.pbase = PR.Q_BASE + 2 * self.qstride
        # constant bank: one, b', zero, then frobenius gamma_{k,i}
        self.consts = [(1, 0), btw, (0, 0)] + [frob[k][i] for k in (1, 2, 3) for i in range(1, 6)]

    def run(self, name, slots, bases=(0, 0, 0), live=(True, True)):
        simulate(self.F, self.progs[name].words, slots, self.consts, bases, live)


def loop_digits(ctx):
    if ctx.cv.family == 'bls12':
        return [(ctx.x_abs >> i) & 1 for i in range(64)]
    return ctx.naf


def miller(ctx, pairs):
    """pairs: list of (P=(x,y)|None, Q=((x0,x1),(y0,y1))|None), 1 or 2 entries. Returns (slots, f_base)."""
    np_ = len(pairs)
    slots = [(0, 0)] * ctx.nslots
    live = [True, True]
    for k, (P, Q) in enumerate(pairs):
        live[k] = P is not None and Q is not None
        P = P or (0, 0)
        Q = Q or ((0, 0), (0, 0))
        slots[PR.Q_BASE + ctx.qstride * k] = Q[0]
        slots[PR.Q_BASE + ctx.qstride * k + 1] = Q[1]
        slots[ctx.pbase + k] = (P[0], P[1])
    cur, nxt = 0, 6
    ctx.run('INIT%d' % np_, slots, (cur, nxt, 0), live)
    digits = loop_digits(ctx)
    top = len(digits) - 1
    while digits[top] == 0:
        top -= 1
    dbl_swaps = (1 + np_) % 2 == 1
    add_swaps = np_ % 2 == 1
    for i in range(top - 1, -1, -1):
        ctx.run('DBL%d' % np_, slots, (cur, nxt, 0), live)
        if dbl_swaps:
            cur, nxt = nxt, cur
        d = digits[i]
        if d:
            ctx.run('ADD%d' % np_, slots, (cur, nxt, 0 if d > 0 else 1), live)
            if add_swaps:
                cur, nxt = nxt, cur
    if ctx.cv.family == 'bn':
        ctx.run('BNTAIL%d' % np_, slots, (cur, nxt, 0), live)
        # 2*np_ rewrites: even -> no swap
    if ctx.x_neg:
        ctx.run('CONJ', slots, (nxt, cur, 0), live)
        cur, nxt = nxt, cur
    return slots, cur


def precompute_lines(ctx, Q):
    """Line table of a fixed G2 point: one (r0, r1, r2) triple per Miller step, in loop order.
    Same control flow as csrc/pairing_vm.cuh: vm_lines_kernel."""
    slots = [(0, 0)] * ctx.nslots
    slots[PR.Q_BASE], slots[PR.Q_BASE + 1] = Q[0], Q[1]
    ctx.run('INIT1', slots, (0, 6, 0))
    digits = loop_digits(ctx)
    top = len(digits) - 1
    while digits[top] == 0:
        top -= 1
    table = []

    def emit():
        table.append(tuple(slots[PR.LINE_OUT + j] for j in range(3)))
    for i in range(top - 1, -1, -1):
        ctx.run('PRE_DBL', slots)
        emit()
        d = digits[i]
        if d:
            ctx.run('PRE_ADD', slots, (0, 0, 0 if d > 0 else 1))
            emit()
    if ctx.cv.family == 'bn':
        ctx.run('PRE_TAIL1', slots)
        emit()
        ctx.run('PRE_TAIL2', slots)
        emit()
    return table


def miller_fixed(ctx, pairs):
    """pairs: list of (P | None, line table | None).  Returns (slots, f_base): the same f as miller() on (P, Q)."""
    np_ = len(pairs)
    slots = [(0, 0)] * ctx.nslots
    live = [True, True]
    for k, (P, tab) in enumerate(pairs):
        live[k] = P is not None and tab is not None
        slots[ctx.pbase + k] = P or (0, 0)
    cur, nxt = 0, 6
    ctx.run('INIT%d' % np_, slots, (cur, nxt, 0), live)
    digits = loop_digits(ctx)
    top = len(digits) - 1
    while digits[top] == 0:
        top -= 1
    step = 0

    def load():
        nonlocal step
        for k, (P, tab) in enumerate(pairs):
            for j in range(3):
                slots[PR.T_BASE + 3 * k + j] = tab[step][j] if tab is not None else (0, 0)
        step += 1
    sq_swaps = (1 + np_) % 2 == 1
    ln_swaps = np_ % 2 == 1
    for i in range(top - 1, -1, -1):
        load()
        ctx.run('SQRLINE%d' % np_, slots, (cur, nxt, 0), live)
        if sq_swaps:
            cur, nxt = nxt, cur
        if digits[i]:
            load()
            ctx.run('LINE%d' % np_, slots, (cur, nxt, 0), live)
            if ln_swaps:
                cur, nxt = nxt, cur
    if ctx.cv.family == 'bn':
        for _ in range(2):
            load()
            ctx.run('LINE%d' % np_, slots, (cur, nxt, 0), live)
            if ln_swaps:
                cur, nxt = nxt, cur
    if ctx.x_neg:
        ctx.run('CONJ', slots, (nxt, cur, 0), live)
        cur, nxt = nxt, cur
    return slots, cur


class Regs:
    def __init__(self, n, used):
        self.free = [r for r in range(n) if r not in used]

    def alloc(self):
        return self.free.pop(0)

    def release(self, *rs):
        for r in rs:
            self.free.append(r)
        self.free.sort()


def final_exp(ctx, slots, f_base):
    """slots: register file with f at f_base (0 or 6). Returns base of the result."""
    R = Regs(ctx.nregs, [f_base // 6])

    def op(name, a, b=None):
        d = R.alloc()
        ctx.run(name, slots, (6 * d, 6 * a, 6 * (b if b is not None else 0)))
        return d

    def expx(z):
        """returns register holding z^x; z is kept"""
        top = 63
        while not (ctx.x_abs >> top) & 1:
            top -= 1
        acc = None
        for i in range(top - 1, -1, -1):
            n = op('CYCLO_SQR', z if acc is None else acc)
            if acc is not None:
                R.release(acc)
            acc = n
            if (ctx.x_abs >> i) & 1:
                n = op('F12_MUL', acc, z)
                R.release(acc)
                acc = n
        if ctx.x_neg:
            n = op('CONJ', acc)
            R.release(acc)
            acc = n
        return acc

    f = f_base // 6
    ri = op('F12_INV', f)
    c = op('CONJ', f)
    R.release(f)
    t = op('F12_MUL', c, ri)
    R.release(c, ri)
    u = op('FROB2', t)
    f = op('F12_MUL', u, t)
    R.release(u, t)
    if ctx.cv.family == 'bls12':
        t0 = op('CYCLO_SQR', f)
        t1 = expx(f)
        t2 = op('CONJ', f)
        x = op('F12_MUL', t1, t2); R.release(t1, t2); t1 = x
        t2 = expx(t1)
        x = op('CONJ', t1); R.release(t1); t1 = x
        x = op('F12_MUL', t1, t2); R.release(t1, t2); t1 = x
        t2 = expx(t1)
        x = op('FROB1', t1); R.release(t1); t1 = x
        x = op('F12_MUL', t1, t2); R.release(t1, t2); t1 = x
        x = op('F12_MUL', f, t0); R.release(f, t0); f = x
        t0 = expx(t1)
        t2 = expx(t0)
        R.release(t0)
        t0 = op('FROB2', t1)
        x = op('CONJ', t1); R.release(t1); t1 = x
        x = op('F12_MUL', t1, t2); R.release(t1, t2); t1 = x
        x = op('F12_MUL', t1, t0); R.release(t1, t0); t1 = x
        x = op('F12_MUL', f, t1); R.release(f, t1); f = x
        return 6 * f
    # BN254: Fuentes-Castaneda chain
    e = expx(f)
    t0 = op('CONJ', e); R.release(e)
    x = op('CYCLO_SQR', t0); R.release(t0); t0 = x
    t1 = op('CYCLO_SQR', t0)
    x = op('F12_MUL', t0, t1); R.release(t1); t1 = x
    e = expx(t1)
    t2 = op('CONJ', e); R.release(e)
    t3 = op('CONJ', t1)
    x = op('F12_MUL', t2, t3); R.release(t1, t3); t1 = x
    t3 = op('CYCLO_SQR', t2)
    t4 = expx(t3)
    x = op('F12_MUL', t1, t4); R.release(t4, t1); t4 = x
    x = op('F12_MUL', t0, t4); R.release(t3); t3 = x
    x = op('F12_MUL', t2, t4); R.release(t0, t2); t0 = x
    x = op('F12_MUL', f, t0); R.release(t0); t0 = x
    t2 = op('FROB1', t3)
    x = op('F12_MUL', t2, t0); R.release(t2, t0); t0 = x
    t2 = op('FROB2', t4); R.release(t4)
    x = op('F12_MUL', t2, t0); R.release(t2, t0); t0 = x
    t2 = op('CONJ', f); R.release(f)
    x = op('F12_MUL', t2, t3); R.release(t2, t3); t2 = x
    x = op('FROB3', t2); R.release(t2); t2 = x
    x = op('F12_MUL', t2, t0); R.release(t2, t0); t0 = x
    return 6 * t0


def gt_exp(ctx, slots, k, top=None):
    """Gt.Exp: the Fp12 in register 0 (slots 0..5) raised to the integer k >= 0 by a left-to-right ladder over 2-bit
    digits: two generic squarings and one multiply by a^d per digit, d = 0 predicated off (the kernel runs the groups of a
    warp in lock-step, each with its own exponent, so the multiply is always executed).  `top` = bit length used by the
    ladder (the kernel takes the warp maximum).  Returns the slot base of the result."""
    cur, oth = 6, 12
    ctx.run('F12_SQR', slots, (18, 0, 0))
    ctx.run('F12_MUL', slots, (24, 18, 0))
    ctx.run('GT_ONE', slots, (cur, 0, 0))
    top = k.bit_length() if top is None else top
    for j in range((top + 1) // 2 - 1, -1, -1):
        d = (k >> (2 * j)) & 3
        ctx.run('F12_SQR', slots, (oth, cur, 0))
        ctx.run('F12_SQR', slots, (cur, oth, 0))
        ctx.run('F12_MULP', slots, (oth, cur, {0: 0, 1: 0, 2: 18, 3: 24}[d]), (d != 0, True))
        cur, oth = oth, cur
    return cur


def f12_from_slots(slots, base):
    """w-basis g0..g5 -> ((C0.B0,C0.B1,C0.B2),(C1.B0,C1.B1,C1.B2))"""
    g = slots[base:base + 6]
    return ((g[0], g[2], g[4]), (g[1], g[3], g[5]))
