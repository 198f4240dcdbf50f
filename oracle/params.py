"""Curve parameters for the oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Everything is derived from the seed x with the published BN / BLS12 polynomials and
checked in tests/test_oracle_pins.py against the constants the reference holds:
group orders ``math_test.go:261-270``, G1 generators ``math_test.go:250-259``,
BLS12-381 modulus limbs ``driver/kilic/custom.go:26``.
"""
from dataclasses import dataclass, field


@dataclass(frozen=True)
class CurveParams:
    name: str
    x: int                      # seed (signed)
    p: int
    r: int
    b: int                      # E: y^2 = x^3 + b
    beta: int                   # Fp2 = Fp[u]/(u^2 - beta)
    xi: tuple                   # Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v)
    twist: str                  # 'D' (b' = b/xi) or 'M' (b' = b*xi)
    g1: tuple
    g2: tuple                   # ((x0,x1),(y0,y1))
    fp_bytes: int
    limbs32: int
    family: str                 # 'bn' | 'bls12'
    flag_bits: int              # free bits in the top byte of an encoded Fp (2 for BN254, 3 for BLS12)

    @property
    def fexp_scale(self):
        """s in the final exponent s*(p^12-1)/r."""
        if self.family == 'bls12':
            return 3
        x = self.x
        return 2 * x * (6 * x * x + 3 * x + 1)


def _bn(x):
    p = 36 * x**4 + 36 * x**3 + 24 * x**2 + 6 * x + 1
    r = 36 * x**4 + 36 * x**3 + 18 * x**2 + 6 * x + 1
    return p, r


def _bls12(x):
    r = x**4 - x**2 + 1
    p = (x - 1) ** 2 * r // 3 + x
    return p, r


_x_bn254 = 4965661367192848881
_p, _r = _bn(_x_bn254)
BN254 = CurveParams(
    name='BN254', x=_x_bn254, p=_p, r=_r, b=3, beta=-1, xi=(9, 1), twist='D',
    g1=(1, 2),
    g2=((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    fp_bytes=32, limbs32=8, family='bn', flag_bits=2)

_x_381 = -0xd201000000010000
_p, _r = _bls12(_x_381)
BLS12_381 = CurveParams(
    name='BLS12_381', x=_x_381, p=_p, r=_r, b=4, beta=-1, xi=(1, 1), twist='M',
    g1=(3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
        1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569),
    g2=((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
         0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
        (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
         0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    fp_bytes=48, limbs32=12, family='bls12', flag_bits=3)

_x_377 = 0x8508c00000000001
_p, _r = _bls12(_x_377)
BLS12_377 = CurveParams(
    name='BLS12_377', x=_x_377, p=_p, r=_r, b=1, beta=-5, xi=(0, 1), twist='D',
    g1=(81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
        241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030),
    g2=((233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
         140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118),
        (63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
         149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491)),
    fp_bytes=48, limbs32=12, family='bls12', flag_bits=3)

CURVES = {c.name: c for c in (BN254, BLS12_381, BLS12_377)}

# mathlib CurveID values (reference math.go:70-103) -> (params, driver semantics)
#   'gurvy': Pairing = raw Miller loop, FExp = final exponentiation
#   'kilic': Pairing = Miller loop + final exponentiation, FExp = identity
CURVE_IDS = {
    1: (BN254, 'gurvy'),
    3: (BLS12_381, 'kilic'),
    4: (BLS12_377, 'gurvy'),
    5: (BLS12_381, 'gurvy'),
    6: (BLS12_381, 'kilic'),   # BLS12_381_BBS (kilic)
    7: (BLS12_381, 'gurvy'),   # BLS12_381_BBS_GURVY
}
