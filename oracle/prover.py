"""ORACLE (test infrastructure, never shipped, never on the product path).

Big-integer restatement of the PLONK prover and verifier of halo2_proofs for KZG/bn256 with the
GWC multi-opening, Blake2b transcript and `Challenge255`:

* keygen_vk / keygen_pk            -- /root/reference/halo2_proofs/src/plonk/keygen.rs:203-367
* permutation keygen (Assembly)    -- src/plonk/permutation/keygen.rs:27-242
* VerifyingKey::from_parts (hash)  -- src/plonk.rs:171-206, pinned Debug form :220-230
* create_proof                     -- src/plonk/prover.rs:37-651
* permutation argument (prover)    -- src/plonk/permutation/prover.rs:44-328
* vanishing argument (prover)      -- src/plonk/vanishing/prover.rs:36-173
* GWC multiopen (prover/verifier)  -- src/poly/kzg/multiopen/gwc.rs, gwc/prover.rs:31-92, gwc/verifier.rs:33-130
* Blake2b transcript               -- src/transcript.rs:282-430, 486-514
* verify_proof                     -- src/plonk/verifier.rs:27-399, permutation/verifier.rs, vanishing/verifier.rs
                                      (the final pairing check e(L,[s]H) e(R,-H) = 1 of kzg/msm.rs:151-169 is
                                      replaced by the equivalent G1 equation [s]L = R, since tests know s)

* SHPLONK multiopen (prover/verifier) -- src/poly/kzg/multiopen/shplonk.rs:55-134, shplonk/prover.rs:26-285,
                                      shplonk/verifier.rs:52-148; arithmetic.rs:405-478
* lookup argument (prover/verifier) -- src/plonk/lookup/prover.rs:55-475, src/plonk/lookup/verifier.rs:35-210

PARITY.  PINNED against the reference's golden verifying key (tests/plonk_api.rs:626-1019), through the Vesta instance
of this same source that oracle/pasta.py loads: `keygen` (fixed-column commitments, the permutation Assembly's cycle
merging, delta^i * omega^j, `commit_lagrange`) reproduces all 19 commitment points, and `pinned_vk_debug` -- the Debug
rendering of PinnedVerificationKey / PinnedConstraintSystem / Expression / Column / Rotation, field elements and
points (assumption A3 below), i.e. the input of the verifying-key hash that seeds every transcript -- the whole
string, character for character (tests/test_oracle.py::test_reference_golden_commitments,
::test_pinned_vk_of_the_reference).  UNPINNED by reference bytes: proofs themselves -- the reference cannot be built
here (no Rust toolchain), draws its blinding from OsRng and holds no proof fixtures.  Encodings that live in the absent
crate halo2curves 0.3.1 are restated from its published source and are ASSUMPTIONS, listed in DESIGN.md: (A2)
G1Affine::to_bytes = 32-byte LE x with bit 7 of byte 31 = parity of y, identity = zeros; (A3) `{:?}` of Fr/Fq = "0x" +
64 lowercase hex digits, big-endian, of a point "(x, y)" (confirmed for pasta by the golden key); (A4)
Fr::random(rng) and from_bytes_wide = the 512-bit little-endian integer mod r, eight rng.next_u64() draws, low limb
first.  Proofs are pinned by the restated reference verifier: every proof is accepted, tampered ones are rejected
(tests/test_oracle_prover.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import hashlib
from typing import Callable, List, Optional, Sequence, Tuple

from . import bn256 as O
from .bn256 import Q_MOD, R_MOD
from .plonk import ADVICE, DELTA, FIXED, INSTANCE, evaluate_expression

SIGN_BIT = 7  # assumption A2


# --------------------------------------------------------------------------
# encodings
# --------------------------------------------------------------------------
def fr_to_repr(x: int) -> bytes:
    return (x % R_MOD).to_bytes(32, "little")


def fr_from_bytes_wide(b: bytes) -> int:
    assert len(b) == 64
    return int.from_bytes(b, "little") % R_MOD


def g1_to_bytes(p) -> bytes:
    if p is None:
        return bytes(32)
    x, y = p
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= (y & 1) << SIGN_BIT
    return bytes(b)


def g1_from_bytes(b: bytes):
    if b == bytes(32):
        return None
    bb = bytearray(b)
    sign = (bb[31] >> SIGN_BIT) & 1
    bb[31] &= ~(1 << SIGN_BIT) & 0xFF
    x = int.from_bytes(bb, "little")
    assert x < Q_MOD
    y = pow((x * x * x + 3) % Q_MOD, (Q_MOD + 1) // 4, Q_MOD)
    assert y * y % Q_MOD == (x * x * x + 3) % Q_MOD, "not on curve"
    if (y & 1) != sign:
        y = Q_MOD - y
    return (x, y)


def fr_random(rng) -> int:
    """halo2curves Fr::random: from_u512 of eight next_u64 draws (assumption A4)."""
    v = 0
    for i in range(8):
        v |= rng.next_u64() << (64 * i)
    return v % R_MOD


class XorShiftRng:
    """rand_xorshift 0.3 XorShiftRng (the seeded RngCore of the zcash/halo2 ecosystem's tests)."""

    def __init__(self, seed: bytes):
        assert len(seed) == 16
        s = [int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(4)]
        if not any(s):
            s = [0x0BAD5EED, 0x0BAD5EED, 0x0BAD5EED, 0x0BAD5EED]
        self.x, self.y, self.z, self.w = s

    def next_u32(self) -> int:
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def next_u64(self) -> int:  # rand_core::impls::next_u64_via_u32
        lo = self.next_u32()
        hi = self.next_u32()
        return (hi << 32) | lo


# --------------------------------------------------------------------------
# transcript (transcript.rs)
# --------------------------------------------------------------------------
class Blake2bWrite:
    def __init__(self):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.proof = bytearray()

    def squeeze_challenge_scalar(self) -> int:
        self.state.update(b"\x00")
        return fr_from_bytes_wide(self.state.copy().digest())

    def common_point(self, p) -> None:
        assert p is not None, "cannot write points at infinity to the transcript"
        self.state.update(b"\x01")
        self.state.update(p[0].to_bytes(32, "little"))
        self.state.update(p[1].to_bytes(32, "little"))

    def common_scalar(self, s: int) -> None:
        self.state.update(b"\x02")
        self.state.update(fr_to_repr(s))

    def write_point(self, p) -> None:
        self.common_point(p)
        self.proof += g1_to_bytes(p)

    def write_scalar(self, s: int) -> None:
        self.common_scalar(s)
        self.proof += fr_to_repr(s)

    def finalize(self) -> bytes:
        return bytes(self.proof)


class Blake2bRead(Blake2bWrite):
    def __init__(self, proof: bytes):
        super().__init__()
        self.buf = bytes(proof)
        self.pos = 0

    def read_point(self):
        b = self.buf[self.pos:self.pos + 32]
        assert len(b) == 32, "proof truncated"
        self.pos += 32
        p = g1_from_bytes(b)
        self.common_point(p)
        return p

    def read_scalar(self) -> int:
        b = self.buf[self.pos:self.pos + 32]
        assert len(b) == 32, "proof truncated"
        self.pos += 32
        v = int.from_bytes(b, "little")
        assert v < R_MOD
        self.common_scalar(v)
        return v


# --------------------------------------------------------------------------
# constraint system description (plain data) and its pinned Debug form
# --------------------------------------------------------------------------
class CS:
    """Plain-data view of the ConstraintSystem fields the prover and verifier read.
    Expression leaves carry the query index as a 4th element: ("advice", column, rotation, query_index)."""

    def __init__(self, *, num_fixed_columns, num_advice_columns, num_instance_columns, gates, advice_queries,
                 instance_queries, fixed_queries, perm_columns, num_advice_queries, minimum_degree=None,
                 num_challenges=0, advice_column_phase=None, challenge_phase=None, lookups=()):
        self.num_fixed_columns = num_fixed_columns
        self.num_advice_columns = num_advice_columns
        self.num_instance_columns = num_instance_columns
        self.num_challenges = num_challenges
        self.advice_column_phase = list(advice_column_phase or [0] * num_advice_columns)
        self.challenge_phase = list(challenge_phase or [])
        self.gates = gates  # [[expr, ...], ...]
        self.advice_queries = list(advice_queries)  # [(column_index, rotation)]
        self.instance_queries = list(instance_queries)
        self.fixed_queries = list(fixed_queries)
        self.perm_columns = list(perm_columns)  # [(column_type, index)]
        self.num_advice_queries = list(num_advice_queries)
        self.minimum_degree = minimum_degree
        self.lookups = list(lookups)

    @staticmethod
    def _expr_degree(e) -> int:
        k = e[0]
        if k in ("constant", "challenge"):
            return 0
        if k in ("fixed", "advice", "instance"):
            return 1
        if k in ("negated", "scaled"):
            return CS._expr_degree(e[1])
        if k == "sum":
            return max(CS._expr_degree(e[1]), CS._expr_degree(e[2]))
        return CS._expr_degree(e[1]) + CS._expr_degree(e[2])

    def degree(self) -> int:  # circuit.rs:1974-2002
        d = 3
        for inp, tab in self.lookups:
            d = max(d, max(4, 2 + max([1] + [self._expr_degree(e) for e in inp])
                           + max([1] + [self._expr_degree(e) for e in tab])))
        for polys in self.gates:
            for p in polys:
                d = max(d, self._expr_degree(p))
        return max(d, self.minimum_degree or 1)

    def blinding_factors(self) -> int:  # circuit.rs:2006-2031
        factors = max(self.num_advice_queries) if self.num_advice_queries else 1
        return max(3, factors) + 2

    def phases(self):
        return list(range(max([0] + self.advice_column_phase) + 1))

    def query_index(self, column, rot=0) -> int:
        qs = {ADVICE: self.advice_queries, FIXED: self.fixed_queries, INSTANCE: self.instance_queries}[column[0]]
        return qs.index((column[1], rot))


def _dbg_fr(x: int) -> str:
    return "0x%064x" % (x % R_MOD)


def _dbg_point(p) -> str:
    return "Infinity" if p is None else "(0x%064x, 0x%064x)" % p


def _dbg_expr(e) -> str:  # circuit.rs:1083-1137, compact `{:?}`
    k = e[0]
    if k == "constant":
        return "Constant(%s)" % _dbg_fr(e[1])
    if k in ("fixed", "advice", "instance"):
        return "%s { query_index: %d, column_index: %d, rotation: Rotation(%d) }" % (k.capitalize(), e[3], e[1], e[2])
    if k == "challenge":
        return "Challenge(Challenge { index: %d, phase: Phase(%d) })" % (e[1], e[2] if len(e) > 2 else 0)
    if k == "negated":
        return "Negated(%s)" % _dbg_expr(e[1])
    if k == "sum":
        return "Sum(%s, %s)" % (_dbg_expr(e[1]), _dbg_expr(e[2]))
    if k == "product":
        return "Product(%s, %s)" % (_dbg_expr(e[1]), _dbg_expr(e[2]))
    if k == "scaled":
        return "Scaled(%s, %s)" % (_dbg_expr(e[1]), _dbg_fr(e[2]))
    raise ValueError(k)


_TYPE_NAME = {ADVICE: "Advice", FIXED: "Fixed", INSTANCE: "Instance"}


def _dbg_list(items) -> str:
    return "[" + ", ".join(items) + "]"


def pinned_vk_debug(cs: CS, k: int, extended_k: int, omega: int, fixed_commitments, perm_commitments,
                    base_modulus: int = Q_MOD, scalar_modulus: int = R_MOD) -> str:
    """format!("{:?}", vk.pinned())   plonk.rs:197, :220-230; circuit.rs:1398-1452; domain.rs:470-486
    PINNED: tests/test_oracle.py::test_pinned_vk_of_the_reference compares it character for character with the
    reference's golden verifying key (tests/plonk_api.rs:626-1019; over Vesta, hence the modulus parameters)."""
    def col(t, i):
        return "Column { index: %d, column_type: %s }" % (i, _TYPE_NAME[t])

    def queries(t, qs):
        return _dbg_list("(%s, Rotation(%d))" % (col(t, c), r) for c, r in qs)

    f = ["num_fixed_columns: %d" % cs.num_fixed_columns, "num_advice_columns: %d" % cs.num_advice_columns,
         "num_instance_columns: %d" % cs.num_instance_columns, "num_selectors: 0"]
    if cs.num_challenges > 0:
        f += ["num_challenges: %d" % cs.num_challenges,
              "advice_column_phase: " + _dbg_list("Phase(%d)" % p for p in cs.advice_column_phase),
              "challenge_phase: " + _dbg_list("Phase(%d)" % p for p in cs.challenge_phase)]
    f += ["gates: " + _dbg_list(_dbg_expr(p) for polys in cs.gates for p in polys),
          "advice_queries: " + queries(ADVICE, cs.advice_queries),
          "instance_queries: " + queries(INSTANCE, cs.instance_queries),
          "fixed_queries: " + queries(FIXED, cs.fixed_queries),
          "permutation: Argument { columns: %s }" % _dbg_list(col(t, i) for t, i in cs.perm_columns),
          "lookups: " + _dbg_list("Argument { input_expressions: %s, table_expressions: %s }"
                                  % (_dbg_list(_dbg_expr(e) for e in inp), _dbg_list(_dbg_expr(e) for e in tab))
                                  for inp, tab in cs.lookups),
          "constants: []",
          "minimum_degree: " + ("None" if cs.minimum_degree is None else "Some(%d)" % cs.minimum_degree)]
    return ("PinnedVerificationKey { base_modulus: \"0x%064x\", scalar_modulus: \"0x%064x\", "
            "domain: PinnedEvaluationDomain { k: %d, extended_k: %d, omega: %s }, "
            "cs: PinnedConstraintSystem { %s }, fixed_commitments: %s, "
            "permutation: VerifyingKey { commitments: %s } }"
            % (base_modulus, scalar_modulus, k, extended_k, _dbg_fr(omega), ", ".join(f),
               _dbg_list(_dbg_point(p) for p in fixed_commitments), _dbg_list(_dbg_point(p) for p in perm_commitments)))


def vk_transcript_repr(debug: str) -> int:
    """plonk.rs:192-203"""
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(len(debug).to_bytes(8, "little"))
    h.update(debug.encode())
    return fr_from_bytes_wide(h.digest())


# --------------------------------------------------------------------------
# keygen
# --------------------------------------------------------------------------
class PermutationAssembly:
    """permutation/keygen.rs:16-107"""

    def __init__(self, n: int, columns):
        self.columns = list(columns)
        self.mapping = [[(i, j) for j in range(n)] for i in range(len(columns))]
        self.aux = [[(i, j) for j in range(n)] for i in range(len(columns))]
        self.sizes = [[1] * n for _ in columns]

    def copy(self, left_column, left_row, right_column, right_row) -> None:
        lc = self.columns.index(tuple(left_column))
        rc = self.columns.index(tuple(right_column))
        assert left_row < len(self.mapping[lc]) and right_row < len(self.mapping[rc])
        left_cycle = self.aux[lc][left_row]
        right_cycle = self.aux[rc][right_row]
        if left_cycle == right_cycle:
            return
        if self.sizes[left_cycle[0]][left_cycle[1]] < self.sizes[right_cycle[0]][right_cycle[1]]:
            left_cycle, right_cycle = right_cycle, left_cycle
        self.sizes[left_cycle[0]][left_cycle[1]] += self.sizes[right_cycle[0]][right_cycle[1]]
        i = right_cycle
        while True:
            self.aux[i[0]][i[1]] = left_cycle
            i = self.mapping[i[0]][i[1]]
            if i == right_cycle:
                break
        self.mapping[lc][left_row], self.mapping[rc][right_row] = self.mapping[rc][right_row], self.mapping[lc][left_row]


class ProvingKey:
    pass


def keygen(params: O.ParamsKZG, cs: CS, fixed_values: Sequence[Sequence[int]], copies=()) -> ProvingKey:
    """keygen_vk + keygen_pk (keygen.rs:203-367) for a circuit given as its assigned fixed columns and the
    list of copy constraints ((column_type, index), row, (column_type, index), row)."""
    n, k = params.n, params.k
    dom = O.EvaluationDomain(cs.degree(), k)
    assert n >= cs.blinding_factors() + 3  # minimum_rows
    fixed = [list(col) + [0] * (n - len(col)) for col in fixed_values]
    assert len(fixed) == cs.num_fixed_columns
    asm = PermutationAssembly(n, cs.perm_columns)
    for c in copies:
        asm.copy(*c)
    # permutation/keygen.rs:109-160, 162-241
    omega_powers = [pow(dom.omega, i, R_MOD) for i in range(n)]
    deltaomega = [[w * pow(DELTA, j, R_MOD) % R_MOD for w in omega_powers] for j in range(len(cs.perm_columns))]
    permutations = [[deltaomega[asm.mapping[i][j][0]][asm.mapping[i][j][1]] for j in range(n)]
                    for i in range(len(cs.perm_columns))]
    pk = ProvingKey()
    pk.cs, pk.domain, pk.k, pk.n = cs, dom, k, n
    pk.fixed_commitments = [params.commit_lagrange(p) for p in fixed]
    pk.perm_commitments = [params.commit_lagrange(p) for p in permutations]
    pk.fixed_values = fixed
    pk.fixed_polys = [dom.lagrange_to_coeff(p) for p in fixed]
    pk.fixed_cosets = [dom.coeff_to_extended(p) for p in pk.fixed_polys]
    pk.permutations = permutations
    pk.perm_polys = [dom.lagrange_to_coeff(p) for p in permutations]
    pk.perm_cosets = [dom.coeff_to_extended(p) for p in pk.perm_polys]
    bf = cs.blinding_factors()
    l0 = [0] * n
    l0[0] = 1
    l_blind = [0] * n
    for i in range(bf):
        l_blind[n - 1 - i] = 1
    l_last = [0] * n
    l_last[n - bf - 1] = 1
    ext = lambda v: dom.coeff_to_extended(dom.lagrange_to_coeff(v))  # noqa: E731
    pk.l0, l_blind_e, pk.l_last = ext(l0), ext(l_blind), ext(l_last)
    pk.l_active_row = [(1 - (a + b)) % R_MOD for a, b in zip(pk.l_last, l_blind_e)]
    pk.debug = pinned_vk_debug(cs, k, dom.extended_k, dom.omega, pk.fixed_commitments, pk.perm_commitments)
    pk.transcript_repr = vk_transcript_repr(pk.debug)
    return pk


# --------------------------------------------------------------------------
# prover
# --------------------------------------------------------------------------
def rotate_omega(dom, value: int, rotation: int) -> int:  # domain.rs:396-406
    if rotation >= 0:
        return value * pow(dom.omega, rotation, R_MOD) % R_MOD
    return value * pow(dom.omega_inv, -rotation, R_MOD) % R_MOD


def _poly_scale(p, s):
    return [c * s % R_MOD for c in p]


def _poly_add(a, b):
    return [(x + y) % R_MOD for x, y in zip(a, b)]


def _evaluate_h(pk, advice_polys_all, instance_polys_all, challenges, y, beta, gamma, theta, perm_sets_all,
                lookups_all):
    from .plonk import evaluate_h
    cs, dom = pk.cs, pk.domain
    strip = lambda e: e  # noqa: E731  (evaluate_expression ignores the query index)
    circuits = [dict(advice=[dom.coeff_to_extended(p) for p in adv], instance=[dom.coeff_to_extended(p) for p in ins],
                     perm_sets=[s["coset"] for s in sets],
                     lookups=[dict(product=dom.coeff_to_extended(lk["product_poly"]),
                                   permuted_input=dom.coeff_to_extended(lk["permuted_input_poly"]),
                                   permuted_table=dom.coeff_to_extended(lk["permuted_table_poly"])) for lk in lks])
                for adv, ins, sets, lks in zip(advice_polys_all, instance_polys_all, perm_sets_all, lookups_all)]
    return evaluate_h(k=pk.k, extended_k=dom.extended_k, extended_omega=dom.extended_omega,
                      gates=[[strip(p) for p in polys] for polys in cs.gates], lookups=cs.lookups,
                      perm_columns=cs.perm_columns, chunk_len=cs.degree() - 2, blinding_factors=cs.blinding_factors(),
                      fixed=pk.fixed_cosets, l0=pk.l0, l_last=pk.l_last, l_active_row=pk.l_active_row,
                      sigma_cosets=pk.perm_cosets, circuits=circuits, challenges=challenges, y=y, beta=beta,
                      gamma=gamma, theta=theta)


def create_proof(params: O.ParamsKZG, pk: ProvingKey, witnesses: Sequence[Callable], instances, rng,
                 transcript: Blake2bWrite, multiopen: str = "gwc") -> None:
    """plonk/prover.rs:37-651 with P = ProverGWC (QUERY_INSTANCE = false).
    witnesses[i](phase, challenges: dict) -> {advice column index: [values]} for the columns of that phase
    (the role of Circuit::synthesize); instances[i] = list of instance columns (lists of ints)."""
    cs, dom, n = pk.cs, pk.domain, pk.n
    for inst in instances:
        assert len(inst) == cs.num_instance_columns  # Error::InvalidInstances
    transcript.common_scalar(pk.transcript_repr)  # :64
    bf = cs.blinding_factors()
    # instances (:79-138)
    instance_values, instance_polys = [], []
    for inst in instances:
        vals = []
        for values in inst:
            assert len(values) <= n - (bf + 1)  # Error::InstanceTooLarge
            poly = [0] * n
            for i, v in enumerate(values):
                transcript.common_scalar(v)
                poly[i] = v % R_MOD
            vals.append(poly)
        instance_values.append(vals)
        instance_polys.append([dom.lagrange_to_coeff(p) for p in vals])
    # advice (:287-405)
    unusable_rows_start = n - (bf + 1)
    advice_values = [[[0] * n for _ in range(cs.num_advice_columns)] for _ in instances]
    challenges = {}
    for phase in cs.phases():
        column_indices = [i for i, p in enumerate(cs.advice_column_phase) if p == phase]
        for ci, witness in enumerate(witnesses):
            cols = witness(phase, dict(challenges))
            phase_values = []
            for idx in column_indices:
                v = [x % R_MOD for x in cols.get(idx, [])]
                assert len(v) <= unusable_rows_start  # not_enough_rows_available
                phase_values.append(v + [0] * (n - len(v)))
            for v in phase_values:  # blinding factors (:364-368)
                for r in range(unusable_rows_start, n):
                    v[r] = fr_random(rng)
            for _ in phase_values:  # blinds are drawn (and ignored by KZG) (:371-374)
                fr_random(rng)
            for v in phase_values:
                transcript.write_point(params.commit_lagrange(v))
            for idx, v in zip(column_indices, phase_values):
                advice_values[ci][idx] = v
        for index, p in enumerate(cs.challenge_phase):
            if p == phase:
                challenges[index] = transcript.squeeze_challenge_scalar()
    challenges = [challenges[i] for i in range(cs.num_challenges)]
    theta = transcript.squeeze_challenge_scalar()  # :410
    # lookups: permuted columns (:412-437, lookup/prover.rs:55-140)
    lookups_all = []
    for ci in range(len(instances)):
        lks = []
        for inp, tab in cs.lookups:
            def compress(exprs):
                acc = [0] * n
                for e in exprs:
                    vals = [evaluate_expression(e, i, 1, n, pk.fixed_values, advice_values[ci], instance_values[ci],
                                                challenges) for i in range(n)]
                    acc = [(a * theta + v) % R_MOD for a, v in zip(acc, vals)]
                return acc
            lk = dict(compressed_input=compress(inp), compressed_table=compress(tab))
            lk["permuted_input"], lk["permuted_table"] = permute_expression_pair(
                lk["compressed_input"], lk["compressed_table"], n - (bf + 1), bf, rng)
            for name in ("permuted_input", "permuted_table"):  # commit_values (:114-125)
                lk[name + "_poly"] = dom.lagrange_to_coeff(lk[name])
                fr_random(rng)  # blind
                lk[name + "_commitment"] = params.commit_lagrange(lk[name])
            transcript.write_point(lk["permuted_input_commitment"])
            transcript.write_point(lk["permuted_table_commitment"])
            lks.append(lk)
        lookups_all.append(lks)
    beta = transcript.squeeze_challenge_scalar()   # :440
    gamma = transcript.squeeze_challenge_scalar()  # :443
    # permutation argument commit (permutation/prover.rs:44-190)
    chunk_len = cs.degree() - 2
    perm_sets_all = []
    for ci in range(len(instances)):
        def column_values(c):
            return {ADVICE: advice_values[ci], FIXED: pk.fixed_values, INSTANCE: instance_values[ci]}[c[0]][c[1]]
        deltaomega = 1
        last_z = 1
        sets = []
        for s0 in range(0, len(cs.perm_columns), chunk_len):
            columns = cs.perm_columns[s0:s0 + chunk_len]
            perms = pk.permutations[s0:s0 + chunk_len]
            modified = [1] * n
            for c, perm in zip(columns, perms):
                vals = column_values(c)
                for i in range(n):
                    modified[i] = modified[i] * ((beta * perm[i] + gamma + vals[i]) % R_MOD) % R_MOD
            modified = [pow(m, -1, R_MOD) if m else 0 for m in modified]  # batch_invert: zeros stay zero
            for c in columns:
                vals = column_values(c)
                dw = deltaomega
                for i in range(n):
                    modified[i] = modified[i] * ((dw * beta + gamma + vals[i]) % R_MOD) % R_MOD
                    dw = dw * dom.omega % R_MOD
                deltaomega = deltaomega * DELTA % R_MOD
            z = [last_z]
            for row in range(1, n):
                z.append(z[row - 1] * modified[row - 1] % R_MOD)
            for r in range(n - bf, n):
                z[r] = fr_random(rng)
            last_z = z[n - (bf + 1)]
            fr_random(rng)  # blind
            commitment = params.commit_lagrange(z)
            poly = dom.lagrange_to_coeff(z)
            coset = dom.coeff_to_extended(poly)
            transcript.write_point(commitment)
            sets.append(dict(poly=poly, coset=coset))
        perm_sets_all.append(sets)
    # lookups: grand products (:466-475, lookup/prover.rs:146-250)
    for lks in lookups_all:
        for lk in lks:
            prod = [(beta + a) * (gamma + t) % R_MOD for a, t in zip(lk["permuted_input"], lk["permuted_table"])]
            prod = [pow(v, -1, R_MOD) if v else 0 for v in prod]
            prod = [p * ((a + beta) % R_MOD) % R_MOD * ((t + gamma) % R_MOD) % R_MOD
                    for p, a, t in zip(prod, lk["compressed_input"], lk["compressed_table"])]
            z, state = [], 1
            for cur in [1] + prod:
                state = state * cur % R_MOD
                z.append(state)
            z = z[:n - bf] + [fr_random(rng) for _ in range(bf)]
            assert len(z) == n
            fr_random(rng)  # product_blind
            commitment = params.commit_lagrange(z)
            lk["product_poly"] = dom.lagrange_to_coeff(z)
            transcript.write_point(commitment)
    # vanishing argument: random polynomial (vanishing/prover.rs:36-66)
    random_poly = [fr_random(rng) for _ in range(n)]
    fr_random(rng)  # random_blind
    transcript.write_point(params.commit(random_poly))
    y = transcript.squeeze_challenge_scalar()  # :478
    advice_polys = [[dom.lagrange_to_coeff(v) for v in adv] for adv in advice_values]
    h_ext = _evaluate_h(pk, advice_polys, instance_polys, challenges, y, beta, gamma, theta, perm_sets_all,
                        lookups_all)
    # vanishing.construct (vanishing/prover.rs:69-121)
    h_coeff = dom.extended_to_coeff(dom.divide_by_vanishing_poly(h_ext))
    h_pieces = [h_coeff[i:i + n] for i in range(0, len(h_coeff) - n + 1, n)]  # chunks_exact(n)
    for _ in h_pieces:
        fr_random(rng)  # h_blinds
    for piece in h_pieces:
        transcript.write_point(params.commit(piece))
    x = transcript.squeeze_challenge_scalar()  # :525
    xn = pow(x, n, R_MOD)
    # evals (:548-581)
    for adv in advice_polys:
        for col, rot in cs.advice_queries:
            transcript.write_scalar(O.eval_polynomial(adv[col], rotate_omega(dom, x, rot)))
    for col, rot in cs.fixed_queries:
        transcript.write_scalar(O.eval_polynomial(pk.fixed_polys[col], rotate_omega(dom, x, rot)))
    # vanishing.evaluate (vanishing/prover.rs:124-152)
    h_poly = [0] * n
    for piece in reversed(h_pieces):
        h_poly = _poly_add(_poly_scale(h_poly, xn), piece)
    transcript.write_scalar(O.eval_polynomial(random_poly, x))
    # pk.permutation.evaluate (permutation/prover.rs:208-219)
    for poly in pk.perm_polys:
        transcript.write_scalar(O.eval_polynomial(poly, x))
    # permutation product evals (permutation/prover.rs:222-266)
    x_next = rotate_omega(dom, x, 1)
    x_last = rotate_omega(dom, x, -(bf + 1))
    for sets in perm_sets_all:
        for si, s in enumerate(sets):
            transcript.write_scalar(O.eval_polynomial(s["poly"], x))
            transcript.write_scalar(O.eval_polynomial(s["poly"], x_next))
            if si + 1 < len(sets):
                transcript.write_scalar(O.eval_polynomial(s["poly"], x_last))
    # lookup evals (:588-595, lookup/prover.rs:253-283)
    x_inv = rotate_omega(dom, x, -1)
    for lks in lookups_all:
        for lk in lks:
            for poly, pt in ((lk["product_poly"], x), (lk["product_poly"], x_next), (lk["permuted_input_poly"], x),
                             (lk["permuted_input_poly"], x_inv), (lk["permuted_table_poly"], x)):
                transcript.write_scalar(O.eval_polynomial(poly, pt))
    # the opening queries, in the reference's order (:596-645)
    queries: List[Tuple[int, List[int]]] = []
    for ci in range(len(instances)):
        for col, rot in cs.advice_queries:
            queries.append((rotate_omega(dom, x, rot), advice_polys[ci][col]))
        sets = perm_sets_all[ci]
        for s in sets:
            queries.append((x, s["poly"]))
            queries.append((x_next, s["poly"]))
        for s in list(reversed(sets))[1:]:
            queries.append((x_last, s["poly"]))
        for lk in lookups_all[ci]:  # lookup/prover.rs:286-323
            queries.append((x, lk["product_poly"]))
            queries.append((x, lk["permuted_input_poly"]))
            queries.append((x, lk["permuted_table_poly"]))
            queries.append((x_inv, lk["permuted_input_poly"]))
            queries.append((x_next, lk["product_poly"]))
    for col, rot in cs.fixed_queries:
        queries.append((rotate_omega(dom, x, rot), pk.fixed_polys[col]))
    for poly in pk.perm_polys:
        queries.append((x, poly))
    queries.append((x, h_poly))
    queries.append((x, random_poly))
    if multiopen == "gwc":
        gwc_create_proof(params, transcript, queries)
    else:
        shplonk_create_proof(params, transcript, queries)


def permute_expression_pair(input_expression, table_expression, usable_rows: int, blinding_factors: int, rng):
    """lookup/prover.rs:390-475: sort the input, put each first occurrence's table value beside it, spread the
    unused table values over the repeated rows (largest row first, values ascending), then blinding rows."""
    permuted_input = sorted(input_expression[:usable_rows])
    leftover = {}
    for c in table_expression[:usable_rows]:
        leftover[c] = leftover.get(c, 0) + 1
    permuted_table = [0] * usable_rows
    repeated = []
    for row, v in enumerate(permuted_input):
        if row == 0 or v != permuted_input[row - 1]:
            permuted_table[row] = v
            assert leftover.get(v, 0) > 0, "Error::ConstraintSystemFailure"
            leftover[v] -= 1
        else:
            repeated.append(row)
    for coeff in sorted(leftover):
        for _ in range(leftover[coeff]):
            permuted_table[repeated.pop()] = coeff
    assert not repeated
    permuted_input += [fr_random(rng) for _ in range(blinding_factors + 1)]
    permuted_table += [fr_random(rng) for _ in range(blinding_factors + 1)]
    return permuted_input, permuted_table


def construct_intermediate_sets(queries):
    """gwc.rs:36-61: group by point, first-occurrence order."""
    out: List[Tuple[int, list]] = []
    for q in queries:
        for point, qs in out:
            if point == q[0]:
                qs.append(q)
                break
        else:
            out.append((q[0], [q]))
    return out


def gwc_create_proof(params, transcript, queries) -> None:
    """gwc/prover.rs:42-91"""
    v = transcript.squeeze_challenge_scalar()
    for z, qs in construct_intermediate_sets(queries):
        poly_batch, eval_batch, pw = None, 0, 1
        for _, poly in qs:
            ev = O.eval_polynomial(poly, z)
            scaled = _poly_scale(poly, pw)
            poly_batch = scaled if poly_batch is None else _poly_add(poly_batch, scaled)
            eval_batch = (eval_batch + ev * pw) % R_MOD
            pw = pw * v % R_MOD
        poly_batch[0] = (poly_batch[0] - eval_batch) % R_MOD  # &poly_batch - eval_batch (poly.rs:258-268)
        witness = O.kate_division(poly_batch, z)
        transcript.write_point(params.commit(witness))


def lagrange_interpolate(points: Sequence[int], evals: Sequence[int]) -> List[int]:
    """arithmetic.rs:405-458 (coefficients of the interpolant, lowest degree first)."""
    assert len(points) == len(evals)
    if len(points) == 1:
        return [evals[0] % R_MOD]
    final = [0] * len(points)
    for j, (xj, ev) in enumerate(zip(points, evals)):
        tmp = [1]
        for k, xk in enumerate(points):
            if k == j:
                continue
            denom = pow((xj - xk) % R_MOD, -1, R_MOD)
            tmp = [(a * (-denom * xk) + b * denom) % R_MOD for a, b in zip(tmp + [0], [0] + tmp)]
        final = [(f + c * ev) % R_MOD for f, c in zip(final, tmp)]
    return final


def evaluate_vanishing_polynomial(roots: Sequence[int], z: int) -> int:
    acc = 1
    for r in roots:
        acc = acc * (z - r) % R_MOD
    return acc


def shplonk_intermediate_sets(queries, key=lambda q: id(q[1]), get_eval=None):
    """shplonk.rs:55-134.  queries: (point, commitment, ...); `key` identifies a commitment the way the
    reference's pointer equality does.  -> (rotation_sets [(points sorted, [(commitment query, evals)])],
    super_point_set sorted)."""
    super_points = sorted({q[0] for q in queries})
    commitment_rotations = []  # [(key, representative query, set of points)]
    for q in queries:
        for entry in commitment_rotations:
            if entry[0] == key(q):
                entry[2].add(q[0])
                break
        else:
            commitment_rotations.append((key(q), q, {q[0]}))
    rotation_sets = []  # [(frozenset, [representative queries])]
    for kk, q, rots in commitment_rotations:
        for entry in rotation_sets:
            if entry[0] == rots:
                entry[1].append(q)
                break
        else:
            rotation_sets.append((rots, [q]))
    out = []
    for rots, qs in rotation_sets:
        pts = sorted(rots)
        coms = []
        for q in qs:
            evals = []
            for pt in pts:
                match = next(qq for qq in queries if key(qq) == key(q) and qq[0] == pt)
                evals.append(get_eval(match))
            coms.append((q, evals))
        out.append((pts, coms))
    return out, super_points


def shplonk_create_proof(params, transcript, queries) -> None:
    """shplonk/prover.rs:108-285; queries: (point, coefficient list)."""
    n = params.n
    y = transcript.squeeze_challenge_scalar()
    rotation_sets, super_points = shplonk_intermediate_sets(
        queries, key=lambda q: id(q[1]), get_eval=lambda q: O.eval_polynomial(q[1], q[0]))
    ext = []  # per rotation set: (points, [(poly, low_degree_equivalent)])
    for pts, coms in rotation_sets:
        ext.append((pts, [(q[1], lagrange_interpolate(pts, evals)) for q, evals in coms]))
    v = transcript.squeeze_challenge_scalar()
    h_x = None
    pv = 1
    for pts, coms in ext:
        n_x, py = None, 1
        for poly, low in coms:
            num = list(poly)
            for i, c in enumerate(low):
                num[i] = (num[i] - c) % R_MOD
            num = _poly_scale(num, py)
            n_x = num if n_x is None else _poly_add(n_x, num)
            py = py * y % R_MOD
        for pt in pts:  # div_by_vanishing
            n_x = O.kate_division(n_x, pt)
        n_x = n_x + [0] * (n - len(n_x))
        n_x = _poly_scale(n_x, pv)
        h_x = n_x if h_x is None else _poly_add(h_x, n_x)
        pv = pv * v % R_MOD
    transcript.write_point(params.commit(h_x))
    u = transcript.squeeze_challenge_scalar()
    l_x, z_diffs, pv = None, [], 1
    for pts, coms in ext:
        diffs = [p for p in super_points if p not in pts]
        z_i = evaluate_vanishing_polynomial(diffs, u)
        inner, py = None, 1
        for poly, low in coms:
            c = list(poly)
            c[0] = (c[0] - O.eval_polynomial(low, u)) % R_MOD
            c = _poly_scale(c, py)
            inner = c if inner is None else _poly_add(inner, c)
            py = py * y % R_MOD
        inner = _poly_scale(_poly_scale(inner, z_i), pv)
        l_x = inner if l_x is None else _poly_add(l_x, inner)
        z_diffs.append(z_i)
        pv = pv * v % R_MOD
    zt_eval = evaluate_vanishing_polynomial(super_points, u)
    l_x = [(a - b * zt_eval) % R_MOD for a, b in zip(l_x, h_x)]
    assert O.eval_polynomial(l_x, u) == 0
    h2 = O.kate_division(l_x, u)
    h2 = _poly_scale(h2, pow(z_diffs[0], -1, R_MOD))
    transcript.write_point(params.commit(h2))


# --------------------------------------------------------------------------
# verifier
# --------------------------------------------------------------------------
def l_i_range(dom, x: int, xn: int, rotations: Sequence[int]) -> List[int]:
    """domain.rs:435-460"""
    res = [(x - rotate_omega(dom, 1, r)) % R_MOD for r in rotations]
    res = [pow(v, -1, R_MOD) if v else 0 for v in res]
    common = (xn - 1) * dom.barycentric_weight % R_MOD
    return [rotate_omega(dom, v * common % R_MOD, r) for v, r in zip(res, rotations)]


def _eval_expr_at(e, fixed_evals, advice_evals, instance_evals, challenges) -> int:
    k = e[0]
    if k == "constant":
        return e[1] % R_MOD
    if k == "fixed":
        return fixed_evals[e[3]]
    if k == "advice":
        return advice_evals[e[3]]
    if k == "instance":
        return instance_evals[e[3]]
    if k == "challenge":
        return challenges[e[1]]
    f = lambda t: _eval_expr_at(t, fixed_evals, advice_evals, instance_evals, challenges)  # noqa: E731
    if k == "negated":
        return -f(e[1]) % R_MOD
    if k == "sum":
        return (f(e[1]) + f(e[2])) % R_MOD
    if k == "product":
        return f(e[1]) * f(e[2]) % R_MOD
    return f(e[1]) * e[2] % R_MOD


def verify_proof(params: O.ParamsKZG, s: int, vk: ProvingKey, instances, proof: bytes, multiopen: str = "gwc") -> bool:
    """plonk/verifier.rs:27-399 + GWC verifier; `vk` is the verifying-key half of the ProvingKey object
    (cs, domain, commitments, transcript_repr); `s` replaces the pairing (see the module docstring)."""
    cs, dom, n = vk.cs, vk.domain, vk.n
    t = Blake2bRead(proof)
    try:
        num_proofs = len(instances)
        t.common_scalar(vk.transcript_repr)
        for inst in instances:
            assert len(inst) == cs.num_instance_columns
            for col in inst:
                for v in col:
                    t.common_scalar(v)
        advice_commitments = [[None] * cs.num_advice_columns for _ in range(num_proofs)]
        challenges = [0] * cs.num_challenges
        for phase in cs.phases():
            for ac in advice_commitments:
                for i, p in enumerate(cs.advice_column_phase):
                    if p == phase:
                        ac[i] = t.read_point()
            for i, p in enumerate(cs.challenge_phase):
                if p == phase:
                    challenges[i] = t.squeeze_challenge_scalar()
        theta = t.squeeze_challenge_scalar()
        lookups_permuted = [[(t.read_point(), t.read_point()) for _ in cs.lookups] for _ in range(num_proofs)]
        beta = t.squeeze_challenge_scalar()
        gamma = t.squeeze_challenge_scalar()
        chunk_len = cs.degree() - 2
        n_sets = (len(cs.perm_columns) + chunk_len - 1) // chunk_len
        perm_commitments = [[t.read_point() for _ in range(n_sets)] for _ in range(num_proofs)]
        lookups_product = [[t.read_point() for _ in cs.lookups] for _ in range(num_proofs)]
        random_poly_commitment = t.read_point()
        y = t.squeeze_challenge_scalar()
        h_commitments = [t.read_point() for _ in range(dom.quotient_poly_degree)]
        x = t.squeeze_challenge_scalar()
        xn = pow(x, n, R_MOD)
        # instance evals from the public inputs (verifier.rs:169-207)
        min_rot = min([0] + [r for _, r in cs.instance_queries])
        max_rot = max([0] + [r for _, r in cs.instance_queries])
        max_len = max([0] + [len(col) for inst in instances for col in inst])
        l_i_s = l_i_range(dom, x, xn, list(range(-max_rot, max_len + abs(min_rot))))
        instance_evals = []
        for inst in instances:
            evs = []
            for col, rot in cs.instance_queries:
                vals = inst[col]
                off = max_rot - rot
                evs.append(sum(a * b for a, b in zip(vals, l_i_s[off:off + len(vals)])) % R_MOD)
            instance_evals.append(evs)
        advice_evals = [[t.read_scalar() for _ in cs.advice_queries] for _ in range(num_proofs)]
        fixed_evals = [t.read_scalar() for _ in cs.fixed_queries]
        random_eval = t.read_scalar()
        perm_common = [t.read_scalar() for _ in cs.perm_columns]
        perm_evals = []
        for _ in range(num_proofs):
            sets = []
            for si in range(n_sets):
                e = dict(eval=t.read_scalar(), next=t.read_scalar())
                e["last"] = t.read_scalar() if si + 1 < n_sets else None
                sets.append(e)
            perm_evals.append(sets)
        lookup_evals = [[dict(product=t.read_scalar(), product_next=t.read_scalar(), permuted_input=t.read_scalar(),
                              permuted_input_inv=t.read_scalar(), permuted_table=t.read_scalar())
                         for _ in cs.lookups] for _ in range(num_proofs)]
        # expected h(x) (verifier.rs:240-327, permutation/verifier.rs:101-196)
        bf = cs.blinding_factors()
        l_evals = l_i_range(dom, x, xn, list(range(-(bf + 1), 1)))
        assert len(l_evals) == 2 + bf
        l_last, l_blind, l_0 = l_evals[0], sum(l_evals[1:1 + bf]) % R_MOD, l_evals[1 + bf]
        exprs = []
        for pi in range(num_proofs):
            ae, ie, sets = advice_evals[pi], instance_evals[pi], perm_evals[pi]
            for polys in cs.gates:
                for p in polys:
                    exprs.append(_eval_expr_at(p, fixed_evals, ae, ie, challenges))
            if sets:
                exprs.append(l_0 * (1 - sets[0]["eval"]) % R_MOD)
                exprs.append((sets[-1]["eval"] ** 2 - sets[-1]["eval"]) * l_last % R_MOD)
                for si in range(1, len(sets)):
                    exprs.append((sets[si]["eval"] - sets[si - 1]["last"]) * l_0 % R_MOD)
                for ci, st in enumerate(sets):
                    columns = cs.perm_columns[ci * chunk_len:(ci + 1) * chunk_len]
                    pevals = perm_common[ci * chunk_len:(ci + 1) * chunk_len]

                    def col_eval(c):
                        src = {ADVICE: ae, FIXED: fixed_evals, INSTANCE: ie}[c[0]]
                        return src[cs.query_index(c, 0)]
                    left = st["next"]
                    for c, pe in zip(columns, pevals):
                        left = left * ((col_eval(c) + beta * pe + gamma) % R_MOD) % R_MOD
                    right = st["eval"]
                    current_delta = beta * x % R_MOD * pow(DELTA, ci * chunk_len, R_MOD) % R_MOD
                    for c in columns:
                        right = right * ((col_eval(c) + current_delta + gamma) % R_MOD) % R_MOD
                        current_delta = current_delta * DELTA % R_MOD
                    exprs.append((left - right) * (1 - (l_last + l_blind)) % R_MOD)
            for (inp, tab), le in zip(cs.lookups, lookup_evals[pi]):  # lookup/verifier.rs:92-163
                active_rows = (1 - (l_last + l_blind)) % R_MOD

                def compress(es):
                    acc = 0
                    for e in es:
                        acc = (acc * theta + _eval_expr_at(e, fixed_evals, ae, ie, challenges)) % R_MOD
                    return acc
                left = le["product_next"] * (le["permuted_input"] + beta) % R_MOD * (le["permuted_table"] + gamma) % R_MOD
                right = le["product"] * (compress(inp) + beta) % R_MOD * (compress(tab) + gamma) % R_MOD
                a_minus_s = (le["permuted_input"] - le["permuted_table"]) % R_MOD
                exprs.append(l_0 * (1 - le["product"]) % R_MOD)
                exprs.append(l_last * (le["product"] ** 2 - le["product"]) % R_MOD)
                exprs.append((left - right) * active_rows % R_MOD)
                exprs.append(l_0 * a_minus_s % R_MOD)
                exprs.append(a_minus_s * (le["permuted_input"] - le["permuted_input_inv"]) % R_MOD * active_rows % R_MOD)
        expected_h_eval = 0
        for v in exprs:
            expected_h_eval = (expected_h_eval * y + v) % R_MOD
        expected_h_eval = expected_h_eval * pow((xn - 1) % R_MOD, -1, R_MOD) % R_MOD
        h_commitment = None  # vanishing/verifier.rs:92-103
        for c in reversed(h_commitments):
            h_commitment = O.g1_add(O.g1_mul(h_commitment, xn), c)
        # queries (verifier.rs:329-389)
        x_next = rotate_omega(dom, x, 1)
        x_last = rotate_omega(dom, x, -(bf + 1))
        queries = []
        for pi in range(num_proofs):
            for qi, (col, rot) in enumerate(cs.advice_queries):
                queries.append((rotate_omega(dom, x, rot), advice_commitments[pi][col], advice_evals[pi][qi]))
            for c, e in zip(perm_commitments[pi], perm_evals[pi]):
                queries.append((x, c, e["eval"]))
                queries.append((x_next, c, e["next"]))
            for c, e in list(zip(perm_commitments[pi], perm_evals[pi]))[::-1][1:]:
                queries.append((x_last, c, e["last"]))
            x_inv = rotate_omega(dom, x, -1)
            for (ci_, ct_), cp_, le in zip(lookups_permuted[pi], lookups_product[pi], lookup_evals[pi]):
                queries.append((x, cp_, le["product"]))          # lookup/verifier.rs:166-208
                queries.append((x, ci_, le["permuted_input"]))
                queries.append((x, ct_, le["permuted_table"]))
                queries.append((x_inv, ci_, le["permuted_input_inv"]))
                queries.append((x_next, cp_, le["product_next"]))
        for qi, (col, rot) in enumerate(cs.fixed_queries):
            queries.append((rotate_omega(dom, x, rot), vk.fixed_commitments[col], fixed_evals[qi]))
        for c, e in zip(vk.perm_commitments, perm_common):
            queries.append((x, c, e))
        queries.append((x, h_commitment, expected_h_eval))
        queries.append((x, random_poly_commitment, random_eval))
        if multiopen != "gwc":
            return _shplonk_verify(params, s, t, queries)
        # GWC verifier (gwc/verifier.rs:60-129)
        v = t.squeeze_challenge_scalar()
        sets = construct_intermediate_sets(queries)
        ws = [t.read_point() for _ in sets]
        u = t.squeeze_challenge_scalar()
        commitment_multi, eval_multi, witness, witness_with_aux = None, 0, None, None
        pu = 1
        for (z, qs), w in zip(sets, ws):
            cb, eb, pv = None, 0, 1
            for _, c, e in qs:
                cb = O.g1_add(cb, O.g1_mul(c, pv))
                eb = (eb + pv * e) % R_MOD
                pv = pv * v % R_MOD
            commitment_multi = O.g1_add(commitment_multi, O.g1_mul(cb, pu))
            eval_multi = (eval_multi + pu * eb) % R_MOD
            witness_with_aux = O.g1_add(witness_with_aux, O.g1_mul(w, pu * z % R_MOD))
            witness = O.g1_add(witness, O.g1_mul(w, pu))
            pu = pu * u % R_MOD
        left = witness
        right = O.g1_add(O.g1_add(witness_with_aux, commitment_multi),
                         O.g1_mul(O.g1_neg(params.g[0]), eval_multi))
        if t.pos != len(t.buf):
            return False
        return O.g1_mul(left, s) == right  # e(left, [s]H) = e(right, H)
    except AssertionError:
        return False


def _shplonk_verify(params, s: int, t: Blake2bRead, queries) -> bool:
    """shplonk/verifier.rs:52-148; queries: (point, commitment, eval).  Commitments are identified by the
    position of their first query with the same point object identity rules as the reference's pointer
    equality: here by the identity of the commitment value held in the verifier's tables."""
    # the reference compares commitment *references*; equal values from different columns stay distinct, so
    # key on (id of the tuple object) -- every column's commitment is a distinct object read from the proof
    # or the vk, and repeated queries of one column reuse that object
    rotation_sets, super_points = shplonk_intermediate_sets(queries, key=lambda q: id(q[1]), get_eval=lambda q: q[2])
    y = t.squeeze_challenge_scalar()
    v = t.squeeze_challenge_scalar()
    h1 = t.read_point()
    u = t.squeeze_challenge_scalar()
    h2 = t.read_point()
    z_0_diff_inverse = z_0 = 0
    outer, r_outer_acc, pv = None, 0, 1
    for i, (pts, coms) in enumerate(rotation_sets):
        diffs = [p for p in super_points if p not in pts]
        z_diff_i = evaluate_vanishing_polynomial(diffs, u)
        if i == 0:
            z_0 = evaluate_vanishing_polynomial(pts, u)
            z_0_diff_inverse = pow(z_diff_i, -1, R_MOD)
            z_diff_i = 1
        else:
            z_diff_i = z_diff_i * z_0_diff_inverse % R_MOD
        inner, r_inner_acc, py = None, 0, 1
        for q, evals in coms:
            r_x = lagrange_interpolate(pts, evals)
            r_inner_acc = (r_inner_acc + py * O.eval_polynomial(r_x, u)) % R_MOD
            inner = O.g1_add(inner, O.g1_mul(q[1], py))
            py = py * y % R_MOD
        outer = O.g1_add(outer, O.g1_mul(inner, pv * z_diff_i % R_MOD))
        r_outer_acc = (r_outer_acc + pv * r_inner_acc % R_MOD * z_diff_i) % R_MOD
        pv = pv * v % R_MOD
    outer = O.g1_add(outer, O.g1_mul(params.g[0], -r_outer_acc % R_MOD))
    outer = O.g1_add(outer, O.g1_mul(h1, -z_0 % R_MOD))
    outer = O.g1_add(outer, O.g1_mul(h2, u))
    if t.pos != len(t.buf):
        return False
    return O.g1_mul(h2, s) == outer  # e(h2, [s]H) = e(right, H)
