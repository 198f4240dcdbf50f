"""Microcode compiler for the warp-cooperative pairing VM (build-time tool; no oracle dependency).

Machine model (csrc/vm.cuh is the interpreter):
  * a GROUP of G = 6 lanes works on one pairing product; all big state lives in a per-group file of Fp2 SLOTS in
    shared memory;
  * a PROGRAM is a list of PHASES; in a phase every lane executes one OP (or idles), then the group synchronises;
  * two op kinds:
      DOT  dst = [halve]( scale * redc( sum_t  modA(A_t) * modB(B_t) )  +  sum_i c_i * mod(L_i) )
           -- up to 6 Fp2 products accumulated UNREDUCED in wide registers (lazy reduction), one Montgomery reduction
              per component at the end;
      LIN  dst = [halve]( sum_i c_i * mod(L_i) )        (no multiplier use)
      INV  dst = 1 / A                                   (Fp2 inversion, one lane)
  * operand = (class, index): class selects a base (absolute slot, one of three run-time base registers, or the
    read-only constant bank), so that one program serves every Fp12 register / f double-buffer.
A DOT never writes a slot that any op of the same phase reads (lanes are not in lock-step inside a phase).

The DSL below builds a DAG of Fp2 values; `compile_program` schedules it into phases (ASAP list scheduling, full
phases first), allocates slots with liveness, and encodes the phases as 32-bit words.  `simulate` executes the
encoded words on Python integers -- used by the tests to check every program against the textbook formulas before
any CUDA is involved.
"""
from dataclasses import dataclass, field

G = 6                   # lanes per group
MAX_TERMS = 6
MAX_LIN = 4

# operand modifiers
NEG, CONJ, XI, DBL = 1, 2, 4, 8            # applied in the order CONJ, XI, DBL, NEG
REAL0, REAL1 = 4, 8                        # B operand only: use component 0 / 1 of the slot as an Fp scalar

# operand classes
C_ABS, C_B1, C_B2, C_B3, C_CONST = 0, 1, 2, 3, 4

KIND_NOP, KIND_DOT, KIND_LIN, KIND_INV = 0, 1, 2, 3
# ---- bounds of the lazy reduction (csrc/vm.cuh: dot_operand / exec_op) ------------------------------------------------
# A DOT accumulates  RE = off p^2 + sum A0 B0 - |beta| sum A1 B1  and  IM = sum (A0 B1 + A1 B0)  as plain integers and
# reduces each ONCE.  A, B are the operands after their modifiers; on a field with head room above p (BLS12-381) the
# interpreter leaves additions unreduced, so A0, A1 are bounded by small multiples of p:
#     plain (1, 1);  XI (xi = 1 + u): A0 = x0 - x1 + p < 2p, A1 = x0 + x1 < 2p;  DBL doubles both.
# B operands are canonical (bound 1 per half; a REAL scalar has no second half).  From these the compiler derives, per op,
#     off    = ceil bound of |beta| sum A1 B1 in units of p^2 (index into the table of k p^2, k <= 30), and
#     levels = conditional subtractions (of 4p, 2p, p) that canonicalise redc's result T/R + (< p) < 2^levels p,
# written into the op header (bits 24..28 and 18..19; `levels` is the maximum over the phase so that the lanes agree).
_field = {'p': None, 'limbs': None, 'beta': -1, 'lazy': False}


def set_field(p, limbs32, beta, lazy):
    """field parameters of the curve being compiled (programs.build_all): modulus, 32-bit limbs, u^2, lazy modifiers"""
    _field.update(p=p, limbs=limbs32, beta=beta, lazy=lazy)


def operand_bounds(am):
    """(bound A0, bound A1) in units of p for an A operand with modifiers am"""
    b0 = b1 = 1
    if _field['lazy']:
        if am & XI:
            b0 = b1 = 2
        if am & DBL:
            b0, b1 = 2 * b0, 2 * b1
    return b0, b1


def _dot_k(terms):
    """(off, k): offset index and the bound k of the reduction input in units of p^2; terms = [(a, am, b, bm)]"""
    p, R = _field['p'], 1 << (32 * _field['limbs'])
    ab = -_field['beta']
    off = kre = kim = ksum = 0
    for (_, am, _, bm) in terms:
        a0, a1 = operand_bounds(am)
        b0, b1 = (1, 0) if bm & (REAL0 | REAL1) else (1, 1)
        off += ab * a1 * b1
        kre += a0 * b0
        kim += a0 * b1 + a1 * b0
        ksum += (a0 + a1) * (b0 + b1)
    assert off <= 30, "offset table of k p^2 ends at k = 30"
    assert ksum * p * p < R * R, "T2 overflows 2N words"
    return off, max(off + kre, kim)                       # redc input < k p^2


def dot_bounds(terms):
    """(off, levels) of a dot product; terms = [(a, am, b, bm)]"""
    p, R = _field['p'], 1 << (32 * _field['limbs'])
    off, k = _dot_k(terms)
    for levels in (1, 2, 3):
        # result < k p^2 / R + p  <  2^levels p   <=>   k p < (2^levels - 1) R
        if k * p < ((1 << levels) - 1) * R and (1 << levels) * p <= R:
            return off, levels
    raise AssertionError("dot product too wide for the field's head room: k = %d" % k)


def lin_levels(v):
    """`levels` of a LIN op evaluated as a plain integer sum (csrc/vm.cuh exec_op): the value stays below
    sum |c_i| bound(mod(x_i)) p, which must fit N words and be at most 2^levels p (levels <= 3; the interpreter subtracts
    4p, 2p, p conditionally)."""
    p, R = _field['p'], 1 << (32 * _field['limbs'])
    tot = 0
    for (_, c, m) in v.lin:
        assert abs(c) <= 7, "LIN coefficient beyond the 3-bit small multiple"
        tot += abs(c) * max(operand_bounds(m))
    assert tot * p < R, "LIN sum overflows N words: %d p" % tot
    for levels in (1, 2, 3):
        if tot <= (1 << levels) and (1 << (levels - 1)) * p < R:
            return levels
    raise AssertionError("LIN sum too wide: %d p" % tot)


FOLD_SCALES = (1, 2, 3, 4, 6)


def fold_bounds(v):
    """Can the scale and the linear tail of DOT `v` be applied to the wide values BEFORE the reduction (csrc/vm.cuh exec_op:
    scale * V, then  + |c| * mod(x) * R  into the high half)?  Returns the `levels` this needs, or None if the head room of
    the field does not allow it.  Reduction input  s k p^2 + L p R  with L = sum |c_i| bound(mod(x_i)); the result is below
    (s k p / R + L + 1) p."""
    if v.scale == 1 and not v.lin:
        return None
    p, R = _field['p'], 1 << (32 * _field['limbs'])
    _, k = _dot_k(v.terms)
    s = v.scale
    if s not in FOLD_SCALES:
        return None
    lin = 0
    for (_, c, m) in v.lin:
        mag = abs(c)
        if mag > 4:
            return None
        mb = max(operand_bounds(m))
        if mag * mb * p >= R:
            return None
        lin += mag * mb
    if s * k * p * p + lin * p * R >= R * R:
        return None
    for levels in (1, 2, 3):
        if s * k * p + (lin + 1) * R <= (1 << levels) * R and (1 << levels) * p <= R:
            return levels
    return None


OP_WORDS = 12            # header + 6 terms + 4 lin + 1 spare  (fixed size keeps the interpreter trivial)


class Val:
    """A node of the DAG (an Fp2 value)."""
    _n = 0

    def __init__(self, kind, **kw):
        self.kind = kind                 # 'ref' | 'dot' | 'lin' | 'inv'
        self.terms = kw.get('terms', [])  # [(Val, amod, Val, bmod)]
        self.lin = kw.get('lin', [])      # [(Val, coef, mod)]
        self.scale = kw.get('scale', 1)
        self.halve = kw.get('halve', False)
        self.ref = kw.get('ref')          # (class, index) for kind == 'ref'
        self.pred = kw.get('pred', 0)     # 0 always, 1 / 2: only if pair 0 / 1 is live, else dst = alt
        self.alt = kw.get('alt')
        self.name = kw.get('name', '')
        self.out = None                   # (class, index) if bound to an output location
        self.group = kw.get('group', 0)   # scheduling group: lower groups are scheduled first (bounds live temporaries)
        Val._n += 1
        self.id = Val._n

    def operands(self):
        ops = [t[0] for t in self.terms] + [t[2] for t in self.terms] + [l[0] for l in self.lin]
        if self.alt is not None:
            ops.append(self.alt)
        return ops


def ref(cls, idx, name=''):
    return Val('ref', ref=(cls, idx), name=name)


def const(idx, name=''):
    return Val('ref', ref=(C_CONST, idx), name=name)


def _as_term(x):
    """accepts Val or (Val, mod)"""
    if isinstance(x, tuple):
        return x
    return (x, 0)


def dot(terms, lin=(), scale=1, halve=False, pred=0, alt=None, name=''):
    """terms: [(a, b)] with a, b = Val or (Val, mod)"""
    tt = []
    for a, b in terms:
        a, am = _as_term(a)
        b, bm = _as_term(b)
        tt.append((a, am, b, bm))
    ll = [(_as_term(v)[0], c, _as_term(v)[1]) for v, c in lin]
    assert 1 <= len(tt) <= MAX_TERMS and len(ll) <= MAX_LIN
    return Val('dot', terms=tt, lin=ll, scale=scale, halve=halve, pred=pred, alt=alt, name=name)


def lin(terms, halve=False, name=''):
    """terms: [(Val | (Val, mod), coef)]"""
    ll = [(_as_term(v)[0], c, _as_term(v)[1]) for v, c in terms]
    assert 1 <= len(ll) <= MAX_LIN
    return Val('lin', lin=ll, halve=halve, name=name)


def inv(a, name=''):
    return Val('inv', terms=[(a, 0, a, 0)], name=name)


def mul(a, b, **kw):
    return dot([(a, b)], **kw)


def sqr(a, **kw):
    return dot([(a, a)], **kw)


# ---------------------------------------------------------------------------------------------------------------
# scheduling + slot allocation + encoding
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Program:
    name: str
    words: list                  # encoded phases
    nphases: int
    temps_used: int
    stats: dict = field(default_factory=dict)


def _cost(v):
    if v.kind == 'inv':
        return 1000
    c = 0
    for (_, _, _, bm) in v.terms:
        c += 2 if bm & (REAL0 | REAL1) else 3
    return c


def compile_program(name, outputs, temp_slots, pinned_reads=()):
    """outputs: list of (Val, (class, index)) -- where each result must end up.
    temp_slots: list of absolute slot indices the allocator may use for temporaries.
    Returns Program."""
    # ---- collect DAG
    order, seen = [], set()

    def visit(v):
        if v.id in seen:
            return
        seen.add(v.id)
        for o in v.operands():
            visit(o)
        order.append(v)

    for v, loc in outputs:
        assert v.kind != 'ref', "output %s is a plain reference; wrap it in lin([(x,1)])" % name
        v.out = loc
        visit(v)
    nodes = [v for v in order if v.kind != 'ref']
    users = {v.id: [] for v in order}
    for v in nodes:
        for o in v.operands():
            users[o.id].append(v)

    # ---- list scheduling into phases
    done = set(v.id for v in order if v.kind == 'ref')
    remaining = list(nodes)
    phases = []                      # list of (kind, [Val])
    while remaining:
        ready = [v for v in remaining if all(o.id in done for o in v.operands())]
        assert ready, "cycle in program " + name
        gmin = min(v.group for v in ready)
        ready = [v for v in ready if v.group == gmin]
        r_lin = [v for v in ready if v.kind == 'lin']
        r_dot = [v for v in ready if v.kind == 'dot']
        r_inv = [v for v in ready if v.kind == 'inv']
        emitted = []
        if r_lin:
            for i in range(0, len(r_lin), G):
                phases.append(('lin', r_lin[i:i + G]))
            emitted = r_lin
        elif r_inv:
            phases.append(('inv', r_inv[:1]))
            emitted = r_inv[:1]
        else:
            r_dot.sort(key=lambda v: -_cost(v))
            nfull = len(r_dot) // G
            take = r_dot[:nfull * G] if nfull else r_dot
            for i in range(0, len(take), G):
                phases.append(('dot', take[i:i + G]))
            emitted = take
        for v in emitted:
            done.add(v.id)
        em = set(v.id for v in emitted)
        remaining = [v for v in remaining if v.id not in em]

    # ---- slot allocation
    phase_of = {}
    for pi, (_, vs) in enumerate(phases):
        for v in vs:
            phase_of[v.id] = pi
    last_use = {}
    for v in nodes:
        lu = phase_of[v.id]
        for u in users[v.id]:
            lu = max(lu, phase_of[u.id])
        last_use[v.id] = lu
    # absolute output slots are busy while an earlier value that lives there is still read: we only know about reads
    # through 'ref' nodes, so track the last phase in which each referenced location is read
    loc_last_read = {}
    for v in nodes:
        for o in v.operands():
            if o.kind == 'ref':
                loc_last_read[o.ref] = max(loc_last_read.get(o.ref, -1), phase_of[v.id])
    for loc in pinned_reads:
        loc_last_read[loc] = len(phases)
    free = list(temp_slots)
    loc_of = {}
    release_at = {}
    copies = []
    max_temps = 0
    busy_until = dict(loc_last_read)          # location -> last phase in which its current content is read
    for pi, (kind, vs) in enumerate(phases):
        # free temps whose last use is before this phase
        for vid, slot in list(release_at.items()):
            if last_use[vid] < pi:
                free.append(slot)
                del release_at[vid]
        for v in vs:
            if v.out is not None and busy_until.get(v.out, -1) < pi:
                loc_of[v.id] = v.out
                busy_until[v.out] = last_use[v.id]
                continue
            assert free, "program %s: out of temp slots (have %d)" % (name, len(temp_slots))
            slot = free.pop(0)
            loc_of[v.id] = (C_ABS, slot)
            release_at[v.id] = slot
            max_temps = max(max_temps, len(temp_slots) - len(free))
            if v.out is not None:
                copies.append((v, v.out))
                last_use[v.id] = len(phases)      # keep the temp alive until the final copy
    if copies:
        # final copy phase(s): dst = 1 * src
        for i in range(0, len(copies), G):
            vs = []
            for v, loc in copies[i:i + G]:
                c = Val('lin', lin=[(v, 1, 0)], name='copy')
                loc_of[c.id] = loc
                vs.append(c)
            phases.append(('lin', vs))

    def loc(v):
        return v.ref if v.kind == 'ref' else loc_of[v.id]

    # ---- hazard check: no op writes a location that the same phase reads
    for pi, (kind, vs) in enumerate(phases):
        reads = set()
        for v in vs:
            for o in v.operands():
                reads.add(loc(o))
        for v in vs:
            assert loc(v) not in reads, "program %s phase %d: write-after-read hazard on %s" % (name, pi, loc(v),)
        assert len(set(loc(v) for v in vs)) == len(vs)

    # ---- encode
    def enc_operand(l):
        cls, idx = l
        assert 0 <= idx < 256 and 0 <= cls < 8
        return (cls << 8) | idx

    words = []
    for kind, vs in phases:
        lanes = list(vs) + [None] * (G - len(vs))
        # the widest dot product of the phase: every lane runs the same multi-operand product variant (zero padded),
        # so lanes with fewer terms do not serialise against the others
        pmax = max([len(v.terms) for v in vs if v.kind == 'dot'] + [0])
        fold = {id(v): fold_bounds(v) for v in vs if v.kind == 'dot'}
        levels = max([fold[id(v)] or dot_bounds(v.terms)[1] for v in vs if v.kind == 'dot'] +
                     [lin_levels(v) for v in vs if v.kind == 'lin'] + [1])
        for v in lanes:
            w = [0] * OP_WORDS
            if v is not None:
                k = {'dot': KIND_DOT, 'lin': KIND_LIN, 'inv': KIND_INV}[v.kind]
                nt = len(v.terms) if v.kind != 'lin' else 0
                nl = len(v.lin)
                alt = enc_operand(loc(v.alt)) if v.alt is not None else 0
                assert 1 <= v.scale <= 7
                off = dot_bounds(v.terms)[0] if v.kind == 'dot' else 0
                w[0] = (k | (nt << 4) | (nl << 8) | (v.scale << 12) | ((1 if v.halve else 0) << 15) | (v.pred << 16) |
                        (levels << 18) | (pmax << 20) | (off << 24) |
                        ((1 if v.kind == 'dot' and fold[id(v)] else 0) << 29))
                w[1] = enc_operand(loc(v)) | (alt << 16)
                for t, (a, am, b, bm) in enumerate(v.terms):
                    w[2 + t] = enc_operand(loc(a)) | (enc_operand(loc(b)) << 11) | (am << 22) | (bm << 26)
                for t, (x, c, m) in enumerate(v.lin):
                    assert -16 <= c <= 15 and c != 0
                    w[8 + t] = enc_operand(loc(x)) | ((c & 31) << 11) | (m << 16)
            words += w
    st = {'phases': len(phases), 'dot_phases': sum(1 for k, _ in phases if k == 'dot'),
          'lin_phases': sum(1 for k, _ in phases if k == 'lin'),
          'dot_ops': sum(len(vs) for k, vs in phases if k == 'dot'),
          'cost': sum(max(_cost(v) for v in vs) for k, vs in phases if k in ('dot', 'inv'))}
    return Program(name, words, len(phases), max_temps, st)


# ---------------------------------------------------------------------------------------------------------------
# reference simulator on Python integers (canonical residues, not Montgomery)
# ---------------------------------------------------------------------------------------------------------------
class Field:
    def __init__(self, p, beta, xi):
        self.p, self.beta, self.xi = p, beta % p, (xi[0] % p, xi[1] % p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.beta * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def mod(self, x, m):
        p = self.p
        if m & CONJ:
            x = (x[0], -x[1] % p)
        if m & XI:
            x = self.mul(x, self.xi)
        if m & DBL:
            x = (2 * x[0] % p, 2 * x[1] % p)
        if m & NEG:
            x = (-x[0] % p, -x[1] % p)
        return x

    def inv(self, a):
        p = self.p
        n = pow((a[0] * a[0] - self.beta * a[1] * a[1]) % p, -1, p)
        return (a[0] * n % p, -a[1] * n % p)


def simulate(F, words, slots, consts, bases=(0, 0, 0), live=(True, True)):
    """Execute encoded phases on `slots` (list of Fp2 tuples, modified in place)."""
    p = F.p

    def rd(o):
        cls, idx = (o >> 8) & 7, o & 255
        if cls == C_CONST:
            return consts[idx]
        base = 0 if cls == C_ABS else bases[cls - 1]
        return slots[base + idx]

    def wr_index(o):
        cls, idx = (o >> 8) & 7, o & 255
        assert cls != C_CONST
        base = 0 if cls == C_ABS else bases[cls - 1]
        return base + idx

    nops = len(words) // OP_WORDS
    for ph in range(nops // G):
        pending = []
        for lane in range(G):
            w = words[(ph * G + lane) * OP_WORDS:(ph * G + lane + 1) * OP_WORDS]
            k = w[0] & 15
            if k == KIND_NOP:
                continue
            nt, nl = (w[0] >> 4) & 15, (w[0] >> 8) & 15
            scale, halve, pred = (w[0] >> 12) & 7, (w[0] >> 15) & 1, (w[0] >> 16) & 3
            dst, alt = w[1] & 0x7FF, (w[1] >> 16) & 0x7FF
            if pred and not live[pred - 1]:
                pending.append((wr_index(dst), rd(alt)))
                continue
            if k == KIND_INV:
                pending.append((wr_index(dst), F.inv(rd(w[2] & 0x7FF))))
                continue
            acc = (0, 0)
            for t in range(nt):
                a = F.mod(rd(w[2 + t] & 0x7FF), (w[2 + t] >> 22) & 15)
                braw = rd((w[2 + t] >> 11) & 0x7FF)
                bm = (w[2 + t] >> 26) & 15
                if bm & REAL0:
                    b = (braw[0], 0)
                elif bm & REAL1:
                    b = (braw[1], 0)
                else:
                    b = F.mod(braw, bm & 3)
                pr = F.mul(a, b)
                acc = ((acc[0] + pr[0]) % p, (acc[1] + pr[1]) % p)
            if k == KIND_DOT:
                acc = (acc[0] * scale % p, acc[1] * scale % p)
            for t in range(nl):
                x = F.mod(rd(w[8 + t] & 0x7FF), (w[8 + t] >> 16) & 15)
                c = (w[8 + t] >> 11) & 31
                if c >= 16:
                    c -= 32
                acc = ((acc[0] + c * x[0]) % p, (acc[1] + c * x[1]) % p)
            if halve:
                i2 = pow(2, -1, p)
                acc = (acc[0] * i2 % p, acc[1] * i2 % p)
            pending.append((wr_index(dst), acc))
        for i, v in pending:
            slots[i] = v
