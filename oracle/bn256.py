"""ORACLE (test infrastructure, never shipped, never on the product path).

Big-integer restatement of the two hot paths of halo2_proofs over bn256:

* ``best_multiexp``  -- /root/reference/halo2_proofs/src/arithmetic.rs:13-159
* ``best_fft``       -- /root/reference/halo2_proofs/src/arithmetic.rs:171-274
* ``EvaluationDomain`` -- /root/reference/halo2_proofs/src/poly/domain.rs:39-361
* ``ParamsKZG::setup/commit/commit_lagrange``
                     -- /root/reference/halo2_proofs/src/poly/kzg/commitment.rs:61-129,281-292,327-334

The field / curve arithmetic itself lives in a third-party crate that is NOT in
the reference tree: ``halo2curves`` (git privacy-scaling-explorations/halo2curves,
tag 0.3.1, /root/reference/halo2_proofs/Cargo.toml:51) with ``ff 0.12`` and
``group 0.12``.  It is restated here from the published definition of bn256
(alt_bn128): y^2 = x^3 + 3 over Fq, generator (1, 2), scalar field Fr with
2-adicity 28 and multiplicative generator 7.

PARITY PINNED against reference-held outputs, through the curve the reference holds them for.  The reference's only
golden vectors are the 19 commitment points (and the domain constants) of the pinned verifying key in
tests/plonk_api.rs:624-1020 -- IPA over Vesta.  Its code is generic over the curve, and so is this file: every
constant below can be overridden (`_CURVE_OVERRIDE`), oracle/pasta.py executes this very source a second time with
Vesta's constants, and tests/test_oracle.py::test_reference_golden_commitments reproduces all 19 points bit for bit
through `g_to_lagrange` (best_fft over curve points, :277-301), `best_multiexp` / `multiexp_serial` (:13-159, every
thread count), `small_multiexp` (:105-125), `EvaluationDomain::new` (omega, extended_k) and the keygen of
oracle/prover.py; the field transform is tied to them by commit(lagrange_to_coeff(a)) == commit_lagrange(a) on those
parameters.  What remains recalled rather than verifiable is bn256-specific DATA, not algorithms: the constants of
halo2curves 0.3.1 (moduli, generator (1, 2), 2-adicity 28, multiplicative generator 7 -- re-derived numerically in
tests/test_oracle.py) and its byte encodings (assumptions A1-A4 in DESIGN.md), for which the reference tree holds no
bytes.  Definitional known answers: tests/golden/kat_bn256.json.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# bn256 constants (re-derived in tests/test_oracle.py)
# --------------------------------------------------------------------------
# The restatements below are generic over the curve, like the reference's own code (`C: CurveAffine`).  oracle/pasta.py
# loads a SECOND instance of this very source with Vesta's constants (the curve of the reference's only golden
# vectors, tests/plonk_api.rs:624-1020) by setting `_CURVE_OVERRIDE` in the module namespace before executing it.
_C = globals().get("_CURVE_OVERRIDE") or {}
R_MOD = _C.get("R_MOD", 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001)  # Fr
Q_MOD = _C.get("Q_MOD", 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47)  # Fq
S = _C.get("S", 28)
MULT_GEN = _C.get("MULT_GEN", 7)
ROOT_OF_UNITY = pow(MULT_GEN, (R_MOD - 1) >> S, R_MOD)
ROOT_OF_UNITY_INV = pow(ROOT_OF_UNITY, -1, R_MOD)
ZETA = pow(MULT_GEN, (R_MOD - 1) // 3, R_MOD)
TWO_INV = pow(2, -1, R_MOD)
MONT_R_FR = (1 << 256) % R_MOD
MONT_R_FQ = (1 << 256) % Q_MOD
CURVE_B = _C.get("CURVE_B", 3)  # y^2 = x^3 + b (a = 0 for bn256 and for both pasta curves)
G1_GEN = _C.get("G1_GEN", (1, 2))

Point = Optional[Tuple[int, int]]  # affine; None = identity


# --------------------------------------------------------------------------
# G1 (affine API, Jacobian inside for speed)
# --------------------------------------------------------------------------
def g1_is_on_curve(p: Point) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - CURVE_B) % Q_MOD == 0


def g1_neg(p: Point) -> Point:
    if p is None:
        return None
    return (p[0], (-p[1]) % Q_MOD)


def _jac_double(P):
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return (0, 1, 0)
    q = Q_MOD
    A = X * X % q
    B = Y * Y % q
    C = B * B % q
    D = 2 * ((X + B) * (X + B) - A - C) % q
    E = 3 * A % q
    F = E * E % q
    X3 = (F - 2 * D) % q
    Y3 = (E * (D - X3) - 8 * C) % q
    Z3 = 2 * Y * Z % q
    return (X3, Y3, Z3)


def _jac_add(P, Qp):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Qp
    if Z1 == 0:
        return Qp
    if Z2 == 0:
        return P
    q = Q_MOD
    Z1Z1 = Z1 * Z1 % q
    Z2Z2 = Z2 * Z2 % q
    U1 = X1 * Z2Z2 % q
    U2 = X2 * Z1Z1 % q
    S1 = Y1 * Z2 * Z2Z2 % q
    S2 = Y2 * Z1 * Z1Z1 % q
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (0, 1, 0)
    H = (U2 - U1) % q
    Rr = (S2 - S1) % q
    HH = H * H % q
    HHH = H * HH % q
    V = U1 * HH % q
    X3 = (Rr * Rr - HHH - 2 * V) % q
    Y3 = (Rr * (V - X3) - S1 * HHH) % q
    Z3 = Z1 * Z2 * H % q
    return (X3, Y3, Z3)


def _to_jac(p: Point):
    return (0, 1, 0) if p is None else (p[0], p[1], 1)


def _to_affine(P) -> Point:
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, Q_MOD)
    zi2 = zi * zi % Q_MOD
    return (X * zi2 % Q_MOD, Y * zi2 * zi % Q_MOD)


def g1_add(a: Point, b: Point) -> Point:
    return _to_affine(_jac_add(_to_jac(a), _to_jac(b)))


def g1_double(a: Point) -> Point:
    return _to_affine(_jac_double(_to_jac(a)))


def _jac_mul(P, k: int):
    k %= R_MOD
    acc = (0, 1, 0)
    for bit in bin(k)[2:] if k else "":
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add(acc, P)
    return acc


def g1_mul(p: Point, k: int) -> Point:
    return _to_affine(_jac_mul(_to_jac(p), k))


def batch_to_affine(ps) -> List[Point]:
    """Jacobian -> affine with one inversion (mirrors Curve::batch_normalize)."""
    q = Q_MOD
    prefix = []
    acc = 1
    for (_, _, Z) in ps:
        prefix.append(acc)
        if Z:
            acc = acc * Z % q
    inv = pow(acc, -1, q)
    out: List[Point] = [None] * len(ps)
    for i in range(len(ps) - 1, -1, -1):
        X, Y, Z = ps[i]
        if Z == 0:
            continue
        zi = inv * prefix[i] % q
        inv = inv * Z % q
        zi2 = zi * zi % q
        out[i] = (X * zi2 % q, Y * zi2 * zi % q)
    return out


# --------------------------------------------------------------------------
# MSM
# --------------------------------------------------------------------------
def msm_naive(coeffs: Sequence[int], bases: Sequence[Point]) -> Point:
    """The mathematical definition  sum_i coeffs[i] * bases[i]."""
    assert len(coeffs) == len(bases)
    acc = (0, 1, 0)
    for c, b in zip(coeffs, bases):
        if c % R_MOD and b is not None:
            acc = _jac_add(acc, _jac_mul(_to_jac(b), c))
    return _to_affine(acc)


def _window_size(nbases: int) -> int:
    # arithmetic.rs:16-22
    if nbases < 4:
        return 1
    if nbases < 32:
        return 3
    return math.ceil(math.log(nbases))


def _get_at(segment: int, c: int, repr_le: bytes) -> int:
    # arithmetic.rs:24-42
    skip_bits = segment * c
    skip_bytes = skip_bits // 8
    if skip_bytes >= 32:
        return 0
    v = repr_le[skip_bytes:skip_bytes + 8].ljust(8, b"\0")
    tmp = int.from_bytes(v, "little")
    tmp >>= skip_bits - skip_bytes * 8
    return tmp % (1 << c)


def multiexp_serial(coeffs: Sequence[int], bases: Sequence[Point], acc):
    """arithmetic.rs:13-101 (unsigned windows, 2^c-1 buckets, running sum)."""
    reprs = [int(c % R_MOD).to_bytes(32, "little") for c in coeffs]  # to_repr(), :14
    c = _window_size(len(bases))
    segments = 256 // c + 1  # :44
    for seg in range(segments - 1, -1, -1):
        for _ in range(c):
            acc = _jac_double(acc)  # :47-49
        buckets = [(0, 1, 0)] * ((1 << c) - 1)  # :82
        for rp, base in zip(reprs, bases):
            d = _get_at(seg, c, rp)
            if d != 0 and base is not None:  # :86-88
                buckets[d - 1] = _jac_add(buckets[d - 1], _to_jac(base))
        running = (0, 1, 0)
        for b in reversed(buckets):  # :95-99
            running = _jac_add(b, running)
            acc = _jac_add(acc, running)
    return acc


def small_multiexp(coeffs: Sequence[int], bases: Sequence[Point]) -> Point:
    """arithmetic.rs:105-125: double-and-add, doublings shared across the points."""
    reprs = [(c % R_MOD).to_bytes(32, "little") for c in coeffs]  # to_repr(), :106
    acc = (0, 1, 0)  # C::Curve::identity(), :107
    for byte_idx in range(31, -1, -1):  # :110
        for bit_idx in range(7, -1, -1):  # :112
            acc = _jac_double(acc)  # :113
            for coeff_idx in range(len(reprs)):  # :115
                if (reprs[coeff_idx][byte_idx] >> bit_idx) & 1:  # :116-117
                    acc = _jac_add(acc, _to_jac(bases[coeff_idx]))  # :118
    return _to_affine(acc)


def best_multiexp(coeffs: Sequence[int], bases: Sequence[Point], num_threads: int = 1) -> Point:
    """arithmetic.rs:132-159.  `num_threads` plays rayon's current_num_threads();
    it changes the chunking and the window size but never the result."""
    assert len(coeffs) == len(bases)  # :133
    n = len(coeffs)
    if n > num_threads:
        chunk = n // num_threads
        parts = []
        for s in range(0, n, chunk):
            parts.append(multiexp_serial(coeffs[s:s + chunk], bases[s:s + chunk], (0, 1, 0)))
        acc = (0, 1, 0)
        for p in parts:
            acc = _jac_add(acc, p)
        return _to_affine(acc)
    return _to_affine(multiexp_serial(coeffs, bases, (0, 1, 0)))


# --------------------------------------------------------------------------
# NTT
# --------------------------------------------------------------------------
def omega_for(k: int) -> int:
    """The primitive 2^k-th root every non-bench caller passes (domain.rs:54-73)."""
    assert 0 <= k <= S
    return pow(ROOT_OF_UNITY, 1 << (S - k), R_MOD)


def _bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def best_fft(a: List[int], omega: int, log_n: int) -> None:
    """arithmetic.rs:171-234, the iterative branch (:202-230); the recursive
    branch (:237-274) computes the same butterflies in another order."""
    n = len(a)
    assert n == 1 << log_n  # :184
    r = R_MOD
    for k in range(n):  # :186-191
        rk = _bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)  # :194-200
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % r
    chunk = 2
    twiddle_chunk = n // 2
    for _ in range(log_n):
        half = chunk // 2
        for start in range(0, n, chunk):
            for i in range(half):
                t = a[start + half + i] * tw[i * twiddle_chunk] % r
                u = a[start + i]
                a[start + i] = (u + t) % r
                a[start + half + i] = (u - t) % r
        chunk *= 2
        twiddle_chunk //= 2


def best_fft_group(a: list, omega: int, log_n: int) -> None:
    """best_fft over curve points (G = C::Curve, arithmetic.rs:171-234): the same butterflies with
    group_add / group_sub and group_scale = point * scalar.  `a`: Jacobian points (None = identity)."""
    n = len(a)
    assert n == 1 << log_n  # :184
    r = R_MOD
    for k in range(n):  # :186-191
        rk = _bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)  # :194-200
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % r
    chunk, twiddle_chunk = 2, n // 2
    for _ in range(log_n):  # :202-230
        half = chunk // 2
        for start in range(0, n, chunk):
            for i in range(half):
                t = _jac_mul(a[start + half + i], tw[i * twiddle_chunk])
                u = a[start + i]
                a[start + i] = _jac_add(u, t)
                a[start + half + i] = _jac_add(u, (t[0], (-t[1]) % Q_MOD, t[2]))
        chunk *= 2
        twiddle_chunk //= 2


def g_to_lagrange(g: Sequence[Point], k: int) -> List[Point]:
    """arithmetic.rs:277-301."""
    r = R_MOD
    n_inv = pow(pow(2, -1, r), k, r)  # TWO_INV.pow_vartime([k]), :278
    omega_inv = ROOT_OF_UNITY_INV  # :279-282
    for _ in range(k, S):
        omega_inv = omega_inv * omega_inv % r
    pts = [_to_jac(p) for p in g]
    best_fft_group(pts, omega_inv, k)  # :285
    pts = [_jac_mul(p, n_inv) for p in pts]  # :286-290
    return batch_to_affine(pts)  # :292-298


def dft_naive(a: Sequence[int], omega: int) -> List[int]:
    """O(n^2) definition: out[i] = sum_j a[j] * omega^(i*j)."""
    n = len(a)
    r = R_MOD
    out = []
    for i in range(n):
        wi = pow(omega, i, r)
        acc = 0
        x = 1
        for j in range(n):
            acc = (acc + a[j] * x) % r
            x = x * wi % r
        out.append(acc)
    return out


def domain_roots(j: int, k: int, modulus: int = R_MOD, root_of_unity: int = ROOT_OF_UNITY, s: int = S):
    """(extended_k, extended_omega, omega) of EvaluationDomain::new(j, k), domain.rs:39-73, for any FieldExt
    given by (modulus, ROOT_OF_UNITY, S).  The bn256 domain below goes through here; so does the pin against
    the reference's golden verifying key, which is over the Vesta scalar field (tests/test_oracle.py)."""
    n = 1 << k
    ek = k
    while (1 << ek) < n * (j - 1):  # :49-52
        ek += 1
    assert ek <= s
    extended_omega = pow(root_of_unity, 1 << (s - ek), modulus)  # :54-61
    omega = pow(extended_omega, 1 << (ek - k), modulus)  # :70-73
    return ek, extended_omega, omega


class EvaluationDomain:
    """domain.rs:19-361, field-element (G = Fr) instantiation."""

    def __init__(self, j: int, k: int):
        r = R_MOD
        self.quotient_poly_degree = j - 1  # :41
        self.n = 1 << k
        self.k = k
        ek, self.extended_omega, self.omega = domain_roots(j, k)
        self.extended_k = ek
        self.g_coset = ZETA  # :81
        self.g_coset_inv = ZETA * ZETA % r  # :82
        orig = pow(ZETA, self.n, r)  # :88
        step = pow(self.extended_omega, self.n, r)
        t = []
        cur = orig
        while True:  # :91-97
            t.append(cur)
            cur = cur * step % r
            if cur == orig:
                break
        assert len(t) == 1 << (ek - k)  # :98
        self.t_evaluations = [pow((x - 1) % r, -1, r) for x in t]  # :101-124
        self.ifft_divisor = pow(1 << k, -1, r)
        self.extended_ifft_divisor = pow(1 << ek, -1, r)
        self.barycentric_weight = pow(self.n, -1, r)
        self.omega_inv = pow(self.omega, -1, r)
        self.extended_omega_inv = pow(self.extended_omega, -1, r)

    def extended_len(self) -> int:
        return 1 << self.extended_k

    def _distribute_powers_zeta(self, a: List[int], into_coset: bool) -> None:
        # :335-351
        cp = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        for idx in range(len(a)):
            i = idx % 3
            if i:
                a[idx] = a[idx] * cp[i - 1] % R_MOD

    @staticmethod
    def _ifft(a: List[int], omega_inv: int, log_n: int, divisor: int) -> None:
        # :353-361
        best_fft(a, omega_inv, log_n)
        for i in range(len(a)):
            a[i] = a[i] * divisor % R_MOD

    def lagrange_to_coeff(self, a: Sequence[int]) -> List[int]:
        a = list(a)
        assert len(a) == 1 << self.k  # :227
        self._ifft(a, self.omega_inv, self.k, self.ifft_divisor)
        return a

    def coeff_to_extended(self, a: Sequence[int]) -> List[int]:
        a = list(a)
        assert len(a) == 1 << self.k  # :244
        self._distribute_powers_zeta(a, True)
        a += [0] * (self.extended_len() - len(a))  # :247
        best_fft(a, self.extended_omega, self.extended_k)
        return a

    def extended_to_coeff(self, a: Sequence[int]) -> List[int]:
        a = list(a)
        assert len(a) == self.extended_len()  # :282
        self._ifft(a, self.extended_omega_inv, self.extended_k, self.extended_ifft_divisor)
        self._distribute_powers_zeta(a, False)
        return a[: self.n * self.quotient_poly_degree]  # :299-300

    def divide_by_vanishing_poly(self, a: Sequence[int]) -> List[int]:
        a = list(a)
        assert len(a) == self.extended_len()  # :311
        m = len(self.t_evaluations)
        return [x * self.t_evaluations[i % m] % R_MOD for i, x in enumerate(a)]


def eval_polynomial(poly: Sequence[int], x: int) -> int:
    """arithmetic.rs:304-329 (Horner)."""
    acc = 0
    for c in reversed(poly):
        acc = (acc * x + c) % R_MOD
    return acc


def compute_inner_product(a: Sequence[int], b: Sequence[int]) -> int:
    """arithmetic.rs:331-345."""
    assert len(a) == len(b)  # :334
    acc = 0
    for x, y in zip(a, b):
        acc = (acc + x * y) % R_MOD
    return acc


def kate_division(a: Sequence[int], b: int) -> List[int]:
    """arithmetic.rs:348-367: divides a(X) by X - b, no remainder kept."""
    b = (-b) % R_MOD  # :352
    q = [0] * (len(a) - 1)
    tmp = 0
    for i in range(len(q) - 1, -1, -1):  # q.iter_mut().rev().zip(a.rev())
        lead = (a[i + 1] - tmp) % R_MOD
        q[i] = lead
        tmp = lead * b % R_MOD
    return q


# --------------------------------------------------------------------------
# KZG params (kzg/commitment.rs:61-129) -- only what commit/commit_lagrange need
# --------------------------------------------------------------------------
class ParamsKZG:
    def __init__(self, k: int, g: List[Point], g_lagrange: List[Point]):
        self.k = k
        self.n = 1 << k
        self.g = g
        self.g_lagrange = g_lagrange

    @classmethod
    def setup(cls, k: int, s: int) -> "ParamsKZG":
        """`s` is the toxic secret the reference draws from its rng (:68)."""
        assert k <= S
        n = 1 << k
        r = R_MOD
        G = _to_jac(G1_GEN)
        # g[i] = [s^i] G  (:71-87)
        gj = []
        sp = 1
        for _ in range(n):
            gj.append(_jac_mul(G, sp))
            sp = sp * s % r
        g = batch_to_affine(gj)
        # g_lagrange[i] = [ (s^n - 1)/n * w^i / (s - w^i) ] G   (:89-116)
        root = pow(ROOT_OF_UNITY_INV, -1, r)
        for _ in range(k, S):
            root = root * root % r
        n_inv = pow(n, -1, r)
        multiplier = (pow(s, n, r) - 1) * n_inv % r
        glj = []
        for i in range(n):
            rp = pow(root, i, r)
            scalar = multiplier * rp % r * pow((s - rp) % r, -1, r) % r
            glj.append(_jac_mul(G, scalar))
        return cls(k, g, batch_to_affine(glj))

    def downsize(self, k: int) -> None:
        """kzg/commitment.rs:267-275"""
        assert k <= self.k  # :268
        self.k, self.n = k, 1 << k
        self.g = self.g[: self.n]  # :273
        self.g_lagrange = g_to_lagrange(self.g, k)  # :274

    def commit(self, poly: Sequence[int]) -> Point:
        # :327-334 (blind ignored)
        assert len(self.g) >= len(poly)
        return best_multiexp(list(poly), self.g[: len(poly)])

    def commit_lagrange(self, poly: Sequence[int]) -> Point:
        # :281-292 (blind ignored)
        assert len(self.g_lagrange) >= len(poly)
        return best_multiexp(list(poly), self.g_lagrange[: len(poly)])


# --------------------------------------------------------------------------
# Boundary encodings: Montgomery residues, 4 x u64 little-endian limbs
# --------------------------------------------------------------------------
def fr_to_mont_bytes(x: int) -> bytes:
    return (x % R_MOD * MONT_R_FR % R_MOD).to_bytes(32, "little")


def fr_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(MONT_R_FR, -1, R_MOD) % R_MOD


def fq_to_mont_bytes(x: int) -> bytes:
    return (x % Q_MOD * MONT_R_FQ % Q_MOD).to_bytes(32, "little")


def fq_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(MONT_R_FQ, -1, Q_MOD) % Q_MOD


def g1_to_bytes(p: Point) -> bytes:
    """G1Affine boundary layout: {x: Fq, y: Fq} Montgomery, identity = (0, 0)."""
    if p is None:
        return b"\0" * 64
    return fq_to_mont_bytes(p[0]) + fq_to_mont_bytes(p[1])


def g1_from_bytes(b: bytes) -> Point:
    if b == b"\0" * 64:
        return None
    return (fq_from_mont_bytes(b[:32]), fq_from_mont_bytes(b[32:64]))


def frs_to_bytes(xs: Sequence[int]) -> bytes:
    return b"".join(fr_to_mont_bytes(x) for x in xs)


def frs_from_bytes(b: bytes) -> List[int]:
    rinv = pow(MONT_R_FR, -1, R_MOD)
    return [int.from_bytes(b[i:i + 32], "little") * rinv % R_MOD for i in range(0, len(b), 32)]


def g1s_to_bytes(ps: Sequence[Point]) -> bytes:
    return b"".join(g1_to_bytes(p) for p in ps)
