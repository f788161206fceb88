"""ORACLE (test infrastructure, never shipped, never on the product path).

The reference's ONLY golden outputs of the MSM / NTT chain are the commitments of the pinned verifying key in
/root/reference/halo2_proofs/tests/plonk_api.rs:994-1017 -- keygen_vk over IPA / Vesta: fixed columns ->
`commit_lagrange` -> `best_multiexp`, with `g_lagrange = g_to_lagrange(g)` (best_fft over curve points) and `g`,
`w` drawn by pasta's hash-to-curve (poly/ipa/commitment.rs:158-207).  The reference is generic over the curve, and so
are the restatements in oracle/bn256.py, oracle/plonk.py and oracle/prover.py: this module

  * loads a second instance of those three source files with Vesta's constants (`load_vesta`), and
  * restates what that path needs from the absent dependency `pasta_curves` (0.4, re-exported by halo2curves 0.3.1
    as `halo2curves::pasta`): `CurveExt::hash_to_curve` = hash_to_field (expand_message_xmd over BLAKE2b, 128-byte
    block, zero personalisation) -> simplified SWU on the isogenous curve iso-Vesta (Z = -13) -> point addition ->
    the 3-isogeny to Vesta.  The isogeny is not copied from anywhere: it is Velu's formula for the one rational
    subgroup of order 3 of iso-Vesta (image y^2 = x^3 + 5 * 3^6), followed by (x, y) -> (x / 9, y / 27).

PINNED: `hash_to_curve("Halo2-Parameters")(&[1])` equals the golden key's commitment to its all-zero fixed column
(= 1 * w, poly/ipa/commitment.rs:92-107), and the other 6 fixed and 12 permutation commitments come out of the
restated `ParamsIPA::new` -> `g_to_lagrange` -> keygen -> `commit_lagrange` -> `best_multiexp` chain bit for bit
(tests/test_oracle.py::test_reference_golden_commitments).

Only tests/ may import this module.
"""
from __future__ import annotations

import hashlib
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))

# pasta_curves: Fp = Pallas base field = Vesta scalar field; Fq = Vesta base field = Pallas scalar field
FP = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
FQ = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
VESTA = {"R_MOD": FP, "Q_MOD": FQ, "S": 32, "MULT_GEN": 5, "CURVE_B": 5, "G1_GEN": (FQ - 1, 2)}
# iso-Vesta: y^2 = x^3 + A x + B, 3-isogenous to Vesta (pasta_curves curves.rs `IsoEqAffine`)
ISO_A = 0x267F9B2EE592271A81639C4D96F787739673928C7D01B212C515AD7242EAA6B1
ISO_B = 1265
SWU_Z = FQ - 13


def load_vesta():
    """(bn256-module, plonk-module, prover-module) re-instantiated over Vesta: the same source files executed in a
    package `oracle_vesta` whose `bn256` sees `_CURVE_OVERRIDE = VESTA`."""
    if "oracle_vesta.prover" in sys.modules:
        return tuple(sys.modules["oracle_vesta." + m] for m in ("bn256", "plonk", "prover"))
    pkg = types.ModuleType("oracle_vesta")
    pkg.__path__ = [HERE]
    sys.modules["oracle_vesta"] = pkg
    spec = importlib.util.spec_from_file_location("oracle_vesta.bn256", os.path.join(HERE, "bn256.py"))
    mod = importlib.util.module_from_spec(spec)
    mod._CURVE_OVERRIDE = dict(VESTA)
    sys.modules["oracle_vesta.bn256"] = mod
    spec.loader.exec_module(mod)
    assert mod.R_MOD == FP and mod.Q_MOD == FQ and mod.S == 32
    plonk = importlib.import_module("oracle_vesta.plonk")
    prover = importlib.import_module("oracle_vesta.prover")
    assert prover.O is mod and plonk.R_MOD == FP
    return mod, plonk, prover


# --------------------------------------------------------------------------
# field helpers over Fq (the base field of Vesta)
# --------------------------------------------------------------------------
def _sqrt(a: int, m: int = FQ):
    """Tonelli-Shanks; None if a is not a square."""
    a %= m
    if a == 0:
        return 0
    if pow(a, (m - 1) // 2, m) != 1:
        return None
    s, t = 0, m - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    z = 2
    while pow(z, (m - 1) // 2, m) != m - 1:
        z += 1
    c, x, b, mm = pow(z, t, m), pow(a, (t + 1) // 2, m), pow(a, t, m), s
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2, i = b2 * b2 % m, i + 1
        g = pow(c, 1 << (mm - i - 1), m)
        x, c = x * g % m, g * g % m
        b, mm = b * c % m, i
    return x


def hash_to_field(curve_id: str, domain_prefix: str, message: bytes):
    """pasta_curves hashtocurve.rs `hash_to_field`: two field elements from expand_message_xmd with BLAKE2b
    (64-byte digests, 128-byte block of zeros in front, personalisation = 16 zero bytes), DST =
    domain_prefix || "-" || curve_id || "_XMD:BLAKE2b_SSWU_RO_" || len; each 64-byte block read big-endian mod q."""
    def H(*parts):
        h = hashlib.blake2b(digest_size=64, person=bytes(16))
        for x in parts:
            h.update(x)
        return h.digest()
    dst = domain_prefix.encode() + b"-" + curve_id.encode() + b"_XMD:BLAKE2b_SSWU_RO_"
    dst += bytes([22 + len(curve_id) + len(domain_prefix)])
    b0 = H(bytes(128), message, bytes([0, 128, 0]), dst)
    b1 = H(b0, b"\x01", dst)
    b2 = H(bytes(x ^ y for x, y in zip(b0, b1)), b"\x02", dst)
    return [int.from_bytes(b, "big") % FQ for b in (b1, b2)]


def map_to_curve_simple_swu(u: int):
    """Simplified SWU onto iso-Vesta (draft-irtf-cfrg-hash-to-curve, as pasta_curves implements it; the sign of y is
    the parity of u)."""
    q, A, B, Z = FQ, ISO_A, ISO_B, SWU_Z
    tv1 = (Z * Z * pow(u, 4, q) + Z * u * u) % q
    if tv1 == 0:
        x1 = B * pow(Z * A % q, -1, q) % q
    else:
        x1 = (-B) * pow(A, -1, q) % q * (1 + pow(tv1, -1, q)) % q
    gx1 = (pow(x1, 3, q) + A * x1 + B) % q
    y = _sqrt(gx1)
    if y is not None:
        x = x1
    else:
        x = Z * u * u % q * x1 % q
        y = _sqrt((pow(x, 3, q) + A * x + B) % q)
    if (u & 1) != (y & 1):
        y = (-y) % q
    return (x, y)


def _iso_add(P, Q):
    """Affine addition on iso-Vesta (a != 0, so not the a = 0 formulas of the instance module)."""
    q = FQ
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % q == 0:
            return None
        lam = (3 * P[0] * P[0] + ISO_A) * pow(2 * P[1], -1, q) % q
    else:
        lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, q) % q
    x = (lam * lam - P[0] - Q[0]) % q
    return (x, (lam * (P[0] - x) - P[1]) % q)


def _psi3(x: int) -> int:
    return (3 * pow(x, 4, FQ) + 6 * ISO_A * x * x + 12 * ISO_B * x - ISO_A * ISO_A) % FQ


def _find_three_torsion() -> int:
    """gcd(psi_3(x), x^q - x) over Fq: iso-Vesta has exactly one rational x-coordinate of order 3."""
    q = FQ
    m = [(-ISO_A * ISO_A) % q, 12 * ISO_B % q, 6 * ISO_A % q, 0, 3]

    def pmod(a, mm):
        a = a[:]
        while len(a) >= len(mm):
            c = a[-1] * pow(mm[-1], -1, q) % q
            for i in range(len(mm)):
                a[len(a) - len(mm) + i] = (a[len(a) - len(mm) + i] - c * mm[i]) % q
            a.pop()
        while a and a[-1] == 0:
            a.pop()
        return a

    def pmul(a, b, mm):
        r = [0] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                r[i + j] = (r[i + j] + x * y) % q
        return pmod(r, mm)

    r, base, e = [1], [0, 1], q
    while e:
        if e & 1:
            r = pmul(r, base, m)
        base = pmul(base, base, m)
        e >>= 1
    r = r + [0] * (2 - len(r))
    r[1] = (r[1] - 1) % q
    while r and r[-1] == 0:
        r.pop()
    a, b = m, r
    while b:
        a, b = b, pmod(a, b)
    assert len(a) == 2, "expected exactly one rational 3-torsion abscissa"
    return (-a[0]) * pow(a[1], -1, q) % q


_X0 = None


def iso_map(P):
    """The 3-isogeny iso-Vesta -> Vesta: Velu's formulas for the kernel {O, (x0, +-y0)} (image y^2 = x^3 + 3645 =
    x^3 + 5 * 3^6), then the isomorphism (x, y) -> (x / 9, y / 27) onto y^2 = x^3 + 5."""
    global _X0
    if P is None:
        return None
    q = FQ
    if _X0 is None:
        _X0 = _find_three_torsion()
        assert _psi3(_X0) == 0
    x0 = _X0
    t = 2 * (3 * x0 * x0 + ISO_A) % q
    u = 4 * (pow(x0, 3, q) + ISO_A * x0 + ISO_B) % q
    assert (ISO_A - 5 * t) % q == 0 and (ISO_B - 7 * (u + x0 * t)) % q == 3645  # the image curve
    x, y = P
    di = pow((x - x0) % q, -1, q)
    X = (x + t * di + u * di * di) % q
    Y = y * (1 - t * di * di - 2 * u * di * di * di) % q
    i3 = pow(3, -1, q)
    return (X * i3 * i3 % q, Y * i3 * i3 * i3 % q)


def hash_to_curve(domain_prefix: str, curve_id: str = "vesta"):
    """`C::CurveExt::hash_to_curve(domain_prefix)` -> closure over messages (poly/ipa/commitment.rs:171, 196)."""
    def hasher(message: bytes):
        u0, u1 = hash_to_field(curve_id, domain_prefix, message)
        r = _iso_add(map_to_curve_simple_swu(u0), map_to_curve_simple_swu(u1))
        return iso_map(r)  # cofactor 1: no clearing
    return hasher


class ParamsIPA:
    """poly/ipa/commitment.rs:24-232 for C = EqAffine (Vesta), on the Vesta instance `V` of oracle/bn256.py."""

    def __init__(self, V, k: int):
        assert k < 32  # :161
        self.V, self.k, self.n = V, k, 1 << k
        hasher = hash_to_curve("Halo2-Parameters")
        self.g = [hasher(b"\x00" + i.to_bytes(4, "little")) for i in range(self.n)]  # :166-183
        self.g_lagrange = V.g_to_lagrange(self.g, k)  # :194
        self.w = hasher(b"\x01")  # :196-198
        self.u = hasher(b"\x02")

    def commit_lagrange(self, poly, blind: int = 1):
        """:92-107; `blind` defaults to Blind::default() = 1 (poly/commitment.rs:195-199), what keygen passes."""
        return self.V.best_multiexp(list(poly) + [blind], self.g_lagrange + [self.w])

    def commit(self, poly, blind: int = 1):
        """:212-224"""
        return self.V.best_multiexp(list(poly) + [blind], self.g[:len(poly)] + [self.w]) if len(poly) == self.n else \
            self.V.best_multiexp(list(poly) + [0] * (self.n - len(poly)) + [blind], self.g + [self.w])
