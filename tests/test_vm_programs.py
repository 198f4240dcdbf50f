"""The VM microcode (mathlib_b200/vm) executed by the Python simulator must reproduce the oracle's Miller loop
(raw, Tier-B formulas) and final exponentiation for every curve -- checks programs, scheduling, slot allocation
and the kernel's control flow on the CPU."""
import random

import pytest

from mathlib_b200.vm import driver_ref as DR
from oracle.params import CURVES
from oracle.pairing import Pairing

NAMES = {'BN254': 'BN254', 'BLS381': 'BLS12_381', 'BLS377': 'BLS12_377'}


def make_ctx(vmname):
    P = CURVES[NAMES[vmname]]
    pr = Pairing(P)
    T = pr.T
    frob = {k: T.frob_consts(k) for k in (1, 2, 3)}
    naf = pr.loop_digits if P.family == 'bn' else None
    ctx = DR.CurveCtx(vmname, P.p, P.beta, P.xi, abs(P.x), P.x < 0, naf, pr.C.b2, frob)
    return ctx, pr


@pytest.mark.parametrize("vmname", ['BLS381', 'BN254', 'BLS377'])
def test_vm_pairing_matches_oracle(vmname):
    ctx, pr = make_ctx(vmname)
    C, P = pr.C, pr.P
    rnd = random.Random(42)
    Pa, Qa = C.g1_mul(C.g1, rnd.randrange(P.r)), C.g2_mul(C.g2, rnd.randrange(P.r))
    Pb, Qb = C.g1_mul(C.g1, rnd.randrange(P.r)), C.g2_mul(C.g2, rnd.randrange(P.r))
    for pairs in ([(Pa, Qa)], [(Pa, Qa), (Pb, Qb)], [(None, Qa), (Pb, Qb)], [(Pa, Qa), (Pb, None)]):
        slots, fb = DR.miller(ctx, pairs)
        raw = DR.f12_from_slots(slots, fb)
        assert raw == pr.miller_projective(pairs)
        ob = DR.final_exp(ctx, slots, fb)
        assert DR.f12_from_slots(slots, ob) == pr.final_exp(pr.miller_textbook(pairs))


@pytest.mark.parametrize("vmname", ['BLS381', 'BN254', 'BLS377'])
def test_vm_gt_exp_matches_oracle(vmname):
    """Gt.Exp ladder (F12_SQR + predicated F12_MULP) against the oracle's Fp12 pow, on a non-cyclotomic element too."""
    ctx, pr = make_ctx(vmname)
    T, P = pr.T, pr.P
    rnd = random.Random(5)
    for k in (0, 1, 2, rnd.randrange(1 << 64), P.r - 1):
        g = [(rnd.randrange(P.p), rnd.randrange(P.p)) for _ in range(6)]
        slots = [(0, 0)] * ctx.nslots
        slots[0:6] = g
        ob = DR.gt_exp(ctx, slots, k, top=k.bit_length() + rnd.randrange(0, 4))     # longer ladders (warp maximum) too
        f = ((g[0], g[2], g[4]), (g[1], g[3], g[5]))
        assert DR.f12_from_slots(slots, ob) == T.f12_pow(f, k)


@pytest.mark.parametrize("vmname", ['BLS381', 'BN254', 'BLS377'])
def test_vm_fixed_q_lines_match_miller(vmname):
    """Fixed-Q line tables: PRE_* programs + SQRLINE / LINE reproduce the raw Miller value of the DBL / ADD
    programs bit for bit, for one and two pairs and with a dead pair."""
    ctx, pr = make_ctx(vmname)
    C, P = pr.C, pr.P
    rnd = random.Random(43)
    Pa, Qa = C.g1_mul(C.g1, rnd.randrange(P.r)), C.g2_mul(C.g2, rnd.randrange(P.r))
    Pb, Qb = C.g1_mul(C.g1, rnd.randrange(P.r)), C.g2_mul(C.g2, rnd.randrange(P.r))
    ta, tb = DR.precompute_lines(ctx, Qa), DR.precompute_lines(ctx, Qb)
    for pairs, fixed in (([(Pa, Qa)], [(Pa, ta)]), ([(Pa, Qa), (Pb, Qb)], [(Pa, ta), (Pb, tb)]),
                         ([(None, Qa), (Pb, Qb)], [(None, ta), (Pb, tb)]), ([(Pa, Qa), (Pb, None)], [(Pa, ta), (Pb, None)])):
        slots, fb = DR.miller_fixed(ctx, fixed)
        assert DR.f12_from_slots(slots, fb) == pr.miller_projective(pairs)


def test_program_budgets():
    from mathlib_b200.vm import programs as PR
    for name in PR.CURVES:
        progs = PR.build_all(name)
        nslots, nregs = PR.SLOTCFG[name]
        for pn, p in progs.items():
            assert len(p.words) % (12 * 6) == 0
        assert progs['DBL2'].temps_used <= nslots - PR.miller_temp0()
        assert progs['F12_INV'].temps_used <= nslots - 6 * nregs
