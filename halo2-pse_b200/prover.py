"""Host-side mirror of the reference's key generation and prover for KZG/bn256 (GWC multi-opening,
Blake2b transcript, Challenge255), every polynomial device-resident.

Names, argument meaning, transcript order and RNG-draw order follow halo2_proofs (paths relative to
/root/reference/halo2_proofs/src):

* ``keygen``                       -- plonk/keygen.rs:203-367, plonk/permutation/keygen.rs:27-242
* ``VerifyingKey`` hash            -- plonk.rs:171-206 (Blake2b of the pinned Debug form)
* ``create_proof``                 -- plonk/prover.rs:37-651
* permutation / vanishing argument -- plonk/permutation/prover.rs:44-328, plonk/vanishing/prover.rs:36-173
* ``ProverGWC``                    -- poly/kzg/multiopen/gwc.rs:36-61, gwc/prover.rs:42-91
* ``Blake2bWrite``                 -- transcript.rs:282-430, 486-514

The host does what the reference's host code does besides arithmetic over rows: transcript hashing, the
Fiat-Shamir schedule, O(#columns) scalar bookkeeping and the copy-constraint union-find of keygen.  All
row-wise work (MSMs, NTTs, quotient evaluation, grand products, batched inversion, Horner evaluations,
Kate division, linear combinations, the 512-bit reduction of random scalars) runs in libhalo2b200 on the
GPU; there is no CPU fallback.  The lookup argument's permuted columns (plonk/lookup/prover.rs:55-475) are a
device radix sort + scans (h2b_lookup_permute).

Encodings that live in halo2curves 0.3.1 (absent from the reference tree) are assumptions A2-A4 of DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import hashlib
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from ._ffi import H2B_DEVICE, H2B_HOST, H2BError
from .api import (Q_MOD, R_MOD, Context, DeviceBuffer, EvaluationDomain, ParamsKZG, fr_decode, fr_encode)
from .plonk import (ADVICE, C_HORNER, FIXED, INSTANCE, VS_CONSTANT, VS_THETA, ConstraintSystem, Evaluator, Expression,
                    GraphEvaluator, _ptr_array, _vs, make_eval_columns)

DELTA = pow(7, 1 << 28, R_MOD)  # Fr::DELTA
SIGN_BIT = 7                    # A2: G1Affine::to_bytes keeps the parity of y in bit 7 of byte 31
_MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------
# encodings and transcript
# --------------------------------------------------------------------------
def fr_to_repr(x: int) -> bytes:
    return (x % R_MOD).to_bytes(32, "little")


def g1_to_bytes(p) -> bytes:
    if p is None:
        return bytes(32)
    b = bytearray(p[0].to_bytes(32, "little"))
    b[31] |= (p[1] & 1) << SIGN_BIT
    return bytes(b)


class Blake2bWrite:
    """transcript.rs:282-430 with Challenge255 (:486-514)."""

    def __init__(self):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.writer = bytearray()

    def squeeze_challenge_scalar(self) -> int:
        self.state.update(b"\x00")  # BLAKE2B_PREFIX_CHALLENGE stays absorbed
        return int.from_bytes(self.state.copy().digest(), "little") % R_MOD  # from_bytes_wide

    def common_point(self, p) -> None:
        if p is None:
            raise H2BError(_ffi.H2B_ERR_ARG, "cannot write points at infinity to the transcript")
        self.state.update(b"\x01" + p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little"))

    def common_scalar(self, s: int) -> None:
        self.state.update(b"\x02" + fr_to_repr(s))

    def write_point(self, p) -> None:
        self.common_point(p)
        self.writer += g1_to_bytes(p)

    def write_scalar(self, s: int) -> None:
        self.common_scalar(s)
        self.writer += fr_to_repr(s)

    def finalize(self) -> bytes:
        return bytes(self.writer)

    def state_digest(self) -> bytes:
        """A digest of everything absorbed so far (the sharded prover compares it across ranks)."""
        return self.state.copy().digest()[:32]


def rng_state_digest(rng) -> Optional[bytes]:
    """A digest of the rng's position in its stream, or None for an rng that cannot tell (OsRng-like)."""
    f = getattr(rng, "state_digest", None)
    return f() if f is not None else None


# --------------------------------------------------------------------------
# RngCore implementations (the reference only ever passes OsRng; a seeded rng is what makes proof
# bytes reproducible)
# --------------------------------------------------------------------------
class XorShiftRng:
    """rand_xorshift::XorShiftRng::from_seed([u8; 16]); next_u64 via two next_u32 (low word first)."""

    def __init__(self, seed: bytes):
        if len(seed) != 16:
            raise ValueError("XorShiftRng seed is 16 bytes")
        s = [int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(4)]
        if not any(s):
            s = [0x0BAD5EED] * 4
        self.x, self.y, self.z, self.w = s

    def next_u32(self) -> int:
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def next_u64(self) -> int:
        lo = self.next_u32()
        return (self.next_u32() << 32) | lo

    def state_digest(self) -> bytes:
        return hashlib.sha256(b"xorshift" + b"".join(v.to_bytes(4, "little") for v in (self.x, self.y, self.z, self.w))).digest()

    def fill_u64(self, count: int) -> np.ndarray:
        return np.fromiter((self.next_u64() for _ in range(count)), dtype=np.uint64, count=count)


class CounterRng:
    """A counter-mode RngCore (splitmix64 of seed + index): the same stream whether drawn one u64 at a time
    or in bulk with numpy, so that the 2^k draws of the vanishing argument's random polynomial do not
    serialise on the host.  Not a reference type: any RngCore may be passed to create_proof."""

    def __init__(self, seed: int):
        self.seed = seed & _MASK64
        self.ctr = 0

    def fill_u64(self, count: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            i = np.arange(self.ctr + 1, self.ctr + 1 + count, dtype=np.uint64)
            z = np.uint64(self.seed) + i * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        self.ctr += count
        return z

    def next_u64(self) -> int:
        return int(self.fill_u64(1)[0])

    def state_digest(self) -> bytes:
        return hashlib.sha256(b"counter" + self.seed.to_bytes(8, "little") + self.ctr.to_bytes(16, "little")).digest()

    def fill_fr_device(self, ctx: Context, n: int, out: DeviceBuffer) -> None:
        """n draws of Fr::random straight into device memory: the stream is generated by the same
        counter function on the GPU (h2b_fr_random_counter), then the host counter skips 8 n words."""
        ctx._check(ctx.lib.h2b_fr_random_counter(ctx.h, self.seed, self.ctr, n, out.ptr))
        self.ctr += 8 * n


def fr_random(rng) -> int:
    """Fr::random(rng) = from_u512 of eight next_u64 draws, low limb first (A4)."""
    v = 0
    for i in range(8):
        v |= rng.next_u64() << (64 * i)
    return v % R_MOD


def fr_random_device(ctx: Context, rng, n: int, out: Optional[DeviceBuffer] = None) -> DeviceBuffer:
    """n draws of Fr::random into a device array: the rng words are produced on the host (they are the
    caller's entropy), the 512-bit reductions run on the GPU."""
    out = out or ctx.alloc(max(n, 1) * 32)
    if hasattr(rng, "fill_fr_device"):
        rng.fill_fr_device(ctx, n, out)
        return out
    wide = np.ascontiguousarray(rng.fill_u64(8 * n), dtype=np.uint64)
    ctx._check(ctx.lib.h2b_fr_from_u512(ctx.h, C.c_void_p(wide.ctypes.data), H2B_HOST, n, out.ptr))
    return out


# --------------------------------------------------------------------------
# the pinned Debug form of a verifying key (plonk.rs:192-203, 220-230; circuit.rs:1083-1137, 1398-1452)
# --------------------------------------------------------------------------
_TYPE_NAME = {ADVICE: "Advice", FIXED: "Fixed", INSTANCE: "Instance"}


def _hex(x: int) -> str:
    return "0x%064x" % x


def _fmt_expr(e: Expression) -> str:
    k = e.node[0]
    if k == "constant":
        return "Constant(%s)" % _hex(e.node[1])
    if k in ("fixed", "advice", "instance"):
        return "%s { query_index: %d, column_index: %d, rotation: Rotation(%d) }" % (
            k.capitalize(), e.node[3], e.node[1], e.node[2])
    if k == "challenge":
        return "Challenge(Challenge { index: %d, phase: Phase(%d) })" % (e.node[1], e.node[2])
    if k == "negated":
        return "Negated(%s)" % _fmt_expr(e.node[1])
    if k == "scaled":
        return "Scaled(%s, %s)" % (_fmt_expr(e.node[1]), _hex(e.node[2]))
    return "%s(%s, %s)" % ("Sum" if k == "sum" else "Product", _fmt_expr(e.node[1]), _fmt_expr(e.node[2]))


def pinned_debug(cs: ConstraintSystem, k: int, extended_k: int, omega: int, fixed_commitments, perm_commitments,
                 base_modulus: int = Q_MOD, scalar_modulus: int = R_MOD) -> str:
    """format!("{:?}", vk.pinned()) (plonk.rs:192-203, circuit.rs:1399-1448): the string whose Blake2b hash
    seeds every transcript.  The CPU test-suite checks it, character for character, against the reference's
    golden verifying key of tests/plonk_api.rs (the moduli are parameters only for that test: it is over Vesta)."""
    lst = lambda items: "[" + ", ".join(items) + "]"  # noqa: E731
    col = lambda c: "Column { index: %d, column_type: %s }" % (c.index, _TYPE_NAME[c.column_type])  # noqa: E731
    qs = lambda q: lst("(%s, Rotation(%d))" % (col(c), r) for c, r in q)  # noqa: E731
    pt = lambda p: "Infinity" if p is None else "(%s, %s)" % (_hex(p[0]), _hex(p[1]))  # noqa: E731
    f = ["num_fixed_columns: %d" % cs.num_fixed_columns, "num_advice_columns: %d" % cs.num_advice_columns,
         "num_instance_columns: %d" % cs.num_instance_columns, "num_selectors: 0"]
    if cs.num_challenges > 0:  # multi-phase fields only when used (circuit.rs:1424-1430)
        f += ["num_challenges: %d" % cs.num_challenges,
              "advice_column_phase: " + lst("Phase(%d)" % p for p in cs.advice_column_phase),
              "challenge_phase: " + lst("Phase(%d)" % p for p in cs.challenge_phase)]
    f += ["gates: " + lst(_fmt_expr(p) for _, polys in cs.gates for p in polys),
          "advice_queries: " + qs(cs.advice_queries), "instance_queries: " + qs(cs.instance_queries),
          "fixed_queries: " + qs(cs.fixed_queries),
          "permutation: Argument { columns: %s }" % lst(col(c) for c in cs.permutation.columns),
          "lookups: " + lst("Argument { input_expressions: %s, table_expressions: %s }" % (
              lst(_fmt_expr(e) for e in l.input_expressions), lst(_fmt_expr(e) for e in l.table_expressions))
              for l in cs.lookups),
          "constants: []",
          "minimum_degree: " + ("None" if cs.minimum_degree is None else "Some(%d)" % cs.minimum_degree)]
    return ("PinnedVerificationKey { base_modulus: \"%s\", scalar_modulus: \"%s\", domain: PinnedEvaluationDomain "
            "{ k: %d, extended_k: %d, omega: %s }, cs: PinnedConstraintSystem { %s }, fixed_commitments: %s, "
            "permutation: VerifyingKey { commitments: %s } }" % (
                _hex(base_modulus), _hex(scalar_modulus), k, extended_k, _hex(omega), ", ".join(f),
                lst(pt(p) for p in fixed_commitments), lst(pt(p) for p in perm_commitments)))


# --------------------------------------------------------------------------
# keygen
# --------------------------------------------------------------------------
class PermutationAssembly:
    """plonk/permutation/keygen.rs:16-107: cycles of equal cells, merged smaller-into-larger."""

    def __init__(self, n: int, columns: Sequence):
        m = len(columns)
        self.columns = [tuple(c) for c in columns]
        self.n = n
        # cell (i, j) as the flat index i * n + j
        self.mapping = np.arange(m * n, dtype=np.int64)
        self.aux = np.arange(m * n, dtype=np.int64)
        self.sizes = np.ones(m * n, dtype=np.int64)

    def copy(self, left_column, left_row: int, right_column, right_row: int) -> None:
        try:
            lc, rc = self.columns.index(tuple(left_column)), self.columns.index(tuple(right_column))
        except ValueError:
            raise H2BError(_ffi.H2B_ERR_ARG, "Error::ColumnNotInPermutation")
        if left_row >= self.n or right_row >= self.n:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "Error::BoundsFailure")
        mapping, aux, sizes = self.mapping, self.aux, self.sizes
        left, right = lc * self.n + left_row, rc * self.n + right_row
        left_cycle, right_cycle = int(aux[left]), int(aux[right])
        if left_cycle == right_cycle:
            return
        if sizes[left_cycle] < sizes[right_cycle]:
            left_cycle, right_cycle = right_cycle, left_cycle
        sizes[left_cycle] += sizes[right_cycle]
        i = right_cycle
        while True:
            aux[i] = left_cycle
            i = int(mapping[i])
            if i == right_cycle:
                break
        mapping[left], mapping[right] = mapping[right], mapping[left]


class ProvingKey:
    """plonk.rs:258-305 (ProvingKey + the VerifyingKey inside it), polynomials as DeviceBuffers."""

    def free(self) -> None:
        for name in ("fixed_values", "fixed_polys", "fixed_cosets", "permutations", "permutation_polys",
                     "permutation_cosets"):
            for b in getattr(self, name, []):
                b.free()
        for name in ("l0", "l_last", "l_active_row"):
            if getattr(self, name, None) is not None:
                getattr(self, name).free()
        self.ev.free()
        for pair in getattr(self, "lookup_compress", []):
            for g in pair:
                g.free()
        self.domain.free()


def _as_limbs(values, n: int) -> np.ndarray:
    """An assigned column (ints or (m, 4) limbs, m <= n) zero-padded to n rows of Montgomery limbs."""
    if isinstance(values, np.ndarray) and values.dtype == np.uint64:
        a = values.reshape(-1, 4)
    else:
        a = fr_encode(list(values))
    if a.shape[0] > n:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "Error::not_enough_rows_available")
    if a.shape[0] == n:
        return np.ascontiguousarray(a)
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:a.shape[0]] = a
    return out


def _lookup_compress_graphs(cs: ConstraintSystem):
    """theta-compression of every lookup's input and table expressions over the Lagrange rows
    (lookup/prover.rs:82-104), as interpreter graphs."""
    out = []
    for lookup in cs.lookups:
        pair = []
        for exprs in (lookup.input_expressions, lookup.table_expressions):
            g = GraphEvaluator()
            parts = tuple(g.add_expression(e) for e in exprs)
            g.add_calculation((C_HORNER, _vs(VS_CONSTANT, 0), parts, _vs(VS_THETA)))
            pair.append(g)
        out.append(tuple(pair))
    return out


def keygen(params: ParamsKZG, cs: ConstraintSystem, fixed_values: Sequence, copies: Sequence = ()) -> ProvingKey:
    """keygen_vk + keygen_pk for a circuit handed over as its assigned fixed columns (what
    Assembly::assign_fixed collects) and its copy constraints (Assembly::copy):
    copies = [((column_type, index), row, (column_type, index), row), ...]."""
    ctx, n, k = params.ctx, params.n, params.k
    if n < cs.minimum_rows():
        raise H2BError(_ffi.H2B_ERR_LENGTH, "Error::not_enough_rows_available")
    if len(fixed_values) != cs.num_fixed_columns:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "one assignment per fixed column expected")
    dom = EvaluationDomain(ctx, cs.degree(), k)
    pk = ProvingKey()
    pk.params, pk.cs, pk.domain, pk.k, pk.n = params, cs, dom, k, n
    ext = dom.extended_len()

    def to_coeff(lagrange: DeviceBuffer) -> DeviceBuffer:
        out = ctx.clone(lagrange, n * 32)
        dom.lagrange_to_coeff_device(out)
        return out

    def to_coset(coeff: DeviceBuffer) -> DeviceBuffer:
        out = ctx.alloc(ext * 32)
        dom.coeff_to_extended_device(coeff, out)
        return out

    # fixed columns (keygen.rs:237-258, 300-316)
    pk.fixed_values = [ctx.upload_fr(_as_limbs(v, n)) for v in fixed_values]
    pk.fixed_commitments = params.g_lagrange.msm_many([(b, n) for b in pk.fixed_values])
    pk.fixed_polys = [to_coeff(b) for b in pk.fixed_values]
    pk.fixed_cosets = [to_coset(b) for b in pk.fixed_polys]

    # permutation (permutation/keygen.rs:109-241): sigma_i[j] = delta^i' omega^j' for mapping[i][j] = (i', j')
    pc = cs.permutation.columns
    asm = PermutationAssembly(n, pc)
    for c in copies:
        asm.copy(*c)
    omega = dom.constant("omega")
    ones = np.tile(fr_encode([omega]), (n, 1))
    ones_dev = ctx.upload_fr(ones)
    omega_powers = ctx.running_product(ones_dev, 1, n)  # [omega^0 .. omega^(n-1)]
    ones_dev.free()
    table = np.empty((max(len(pc), 1) * n, 4), dtype=np.uint64)   # deltaomega, flat (i', j')
    for i in range(len(pc)):
        col = ctx.clone(omega_powers, n * 32)
        ctx.poly_scale(col, pow(DELTA, i, R_MOD), n)
        table[i * n:(i + 1) * n] = col.download(n)
        col.free()
    omega_powers.free()
    pk.permutations = [ctx.upload_fr(table[asm.mapping[i * n:(i + 1) * n]]) for i in range(len(pc))]
    pk.perm_commitments = params.g_lagrange.msm_many([(b, n) for b in pk.permutations])
    pk.permutation_polys = [to_coeff(b) for b in pk.permutations]
    pk.permutation_cosets = [to_coset(b) for b in pk.permutation_polys]

    # l_0, l_blind, l_last, l_active_row (keygen.rs:322-350)
    bf = cs.blinding_factors()
    one = fr_encode([1])[0]

    def indicator(rows) -> DeviceBuffer:
        v = np.zeros((n, 4), dtype=np.uint64)
        v[list(rows)] = one
        lag = ctx.upload_fr(v)
        dom.lagrange_to_coeff_device(lag)
        out = to_coset(lag)
        lag.free()
        return out

    pk.l0 = indicator([0])
    l_blind = indicator(range(n - bf, n))
    pk.l_last = indicator([n - bf - 1])
    # l_active_row = one - (l_last + l_blind)
    pk.l_active_row = ctx.upload_fr(np.tile(one, (ext, 1)))
    ctx.poly_sub(pk.l_active_row, pk.l_last, ext)
    ctx.poly_sub(pk.l_active_row, l_blind, ext)
    l_blind.free()

    pk.ev = Evaluator(cs)  # keygen.rs:353
    pk.lookup_compress = _lookup_compress_graphs(cs)
    pk.pinned = pinned_debug(cs, k, dom.extended_k, omega, pk.fixed_commitments, pk.perm_commitments)
    hsh = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    hsh.update(len(pk.pinned).to_bytes(8, "little") + pk.pinned.encode())
    pk.transcript_repr = int.from_bytes(hsh.digest(), "little") % R_MOD  # plonk.rs:192-203
    return pk


# --------------------------------------------------------------------------
# create_proof
# --------------------------------------------------------------------------
class _LookupPolys:
    """The three coefficient-form polynomials Evaluator.evaluate_h reads of a committed lookup."""

    def __init__(self, lk):
        self.product_poly = lk.product_poly.buf
        self.permuted_input_poly = lk.permuted_input_poly.buf
        self.permuted_table_poly = lk.permuted_table_poly.buf


class _Poly:
    """A coefficient-form polynomial on the device with the evaluations already computed for it
    (ProverQuery::get_eval recomputes eval_polynomial; the value is the same)."""

    def __init__(self, ctx: Context, buf: DeviceBuffer, n: int):
        self.ctx, self.buf, self.n = ctx, buf, n
        self.evals: Dict[int, int] = {}

    def eval(self, point: int) -> int:
        if point not in self.evals:
            self.evals[point] = self.ctx.eval_polynomial(self.buf, point, self.n)
        return self.evals[point]


def rotate_omega(dom: EvaluationDomain, omega: int, omega_inv: int, value: int, rotation: int) -> int:
    """poly/domain.rs:396-406"""
    if rotation >= 0:
        return value * pow(omega, rotation, R_MOD) % R_MOD
    return value * pow(omega_inv, -rotation, R_MOD) % R_MOD


def create_proof(params: ParamsKZG, pk: ProvingKey, witnesses: Sequence[Callable], instances: Sequence[Sequence],
                 rng, transcript: Blake2bWrite, timings: Optional[dict] = None, prover=None) -> None:
    """plonk/prover.rs:37-651 with Scheme = KZGCommitmentScheme<Bn256>, E = Challenge255 and P = `prover`
    (ProverGWC by default, or ProverSHPLONK).
    witnesses[i](phase, challenges) -> {advice column index: assigned values (ints or (m, 4) limbs)}: the
    role of Circuit::synthesize through WitnessCollection (:143-285); instances[i] = the instance columns."""
    import time
    t_last = [time.perf_counter()]

    def lap(name: str) -> None:
        if timings is not None:
            ctx.sync()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (now - t_last[0])
            t_last[0] = now

    cs, dom, ctx, n = pk.cs, pk.domain, params.ctx, pk.n
    if len(witnesses) != len(instances):
        raise H2BError(_ffi.H2B_ERR_LENGTH, "one instance list per circuit")
    for inst in instances:
        if len(inst) != cs.num_instance_columns:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "Error::InvalidInstances")  # :55-59
    omega, omega_inv = dom.constant("omega"), dom.constant("omega_inv")
    rot = lambda v, r: rotate_omega(dom, omega, omega_inv, v, r)  # noqa: E731
    bf = cs.blinding_factors()
    ext = dom.extended_len()

    def to_coeff(lagrange: DeviceBuffer) -> DeviceBuffer:
        out = ctx.clone(lagrange, n * 32)
        dom.lagrange_to_coeff_device(out)
        return out

    # ONE proof on several GPUs (dist.ShardedBases): every rank must hold the same polynomials, hence draw the
    # same blinding factors.  Checked here (rng position) and after the openings (transcript), never assumed.
    replicated = getattr(params.g, "check_replicated", None)
    if replicated is not None:
        replicated("the rng passed to create_proof (same seed and position on every rank: dist.broadcast_seed)",
                   rng_state_digest(rng))
    transcript.common_scalar(pk.transcript_repr)  # :62

    # ---- instances (:79-138; QUERY_INSTANCE = false) ----
    instance_values: List[List[DeviceBuffer]] = []
    instance_polys: List[List[DeviceBuffer]] = []
    for inst in instances:
        vals = []
        for values in inst:
            values = [int(v) % R_MOD for v in values]
            if len(values) > n - (bf + 1):
                raise H2BError(_ffi.H2B_ERR_LENGTH, "Error::InstanceTooLarge")
            for v in values:
                transcript.common_scalar(v)
            vals.append(ctx.upload_fr(_as_limbs(values, n)))
        instance_values.append(vals)
        instance_polys.append([to_coeff(b) for b in vals])
    lap("instances")

    # ---- advice (:287-405) ----
    unusable_rows_start = n - (bf + 1)
    advice_values: List[List[Optional[DeviceBuffer]]] = [[None] * cs.num_advice_columns for _ in instances]
    challenges: Dict[int, int] = {}
    for phase in cs.phases():
        column_indices = [i for i, p in enumerate(cs.advice_column_phase) if p == phase]
        for ci, witness in enumerate(witnesses):
            assigned = witness(phase, dict(challenges))
            host_cols = []
            for idx in column_indices:
                col = assigned.get(idx, [])
                if not (isinstance(col, np.ndarray) and col.dtype == np.uint64):
                    col = fr_encode(list(col))
                col = np.ascontiguousarray(col).reshape(-1, 4)
                if col.shape[0] > unusable_rows_start:
                    raise H2BError(_ffi.H2B_ERR_LENGTH, "Error::not_enough_rows_available")  # :228-230
                host_cols.append(col)
            lap("witness")
            bufs, uploads = [], []
            for col in host_cols:  # assigned rows, zero padding, then the blinding factors (:364-368)
                b = ctx.alloc(n * 32)
                if col.shape[0] < unusable_rows_start:
                    ctx.memset(b, 0, (unusable_rows_start - col.shape[0]) * 32, col.shape[0] * 32)
                b.upload(fr_encode([fr_random(rng) for _ in range(n - unusable_rows_start)]), unusable_rows_start * 32)
                bufs.append(b)
                # the assigned rows travel right before the column's commitment, on the context that computes
                # it: the copy of the next column overlaps the commitment of this one
                uploads.append((lambda c, b=b, col=col: b.upload(col, ctx=c)) if col.shape[0] else None)
            for _ in host_cols:    # Blind(Scalar::random(rng)) per column, ignored by KZG (:371-374)
                fr_random(rng)
            spread = getattr(params.g_lagrange, "upload_columns", None)
            if spread is not None:  # bases sharded over GPUs: each rank copies its own rows, NVLink does the rest
                spread(bufs, host_cols)
                uploads = None
            lap("advice_upload")
            for point in params.g_lagrange.msm_many([(b, n) for b in bufs], pre=uploads):  # :375-392, independent commitments
                transcript.write_point(point)
            for idx, b in zip(column_indices, bufs):
                advice_values[ci][idx] = b
            lap("advice_commit")
        for index, p in enumerate(cs.challenge_phase):
            if p == phase:
                challenges[index] = transcript.squeeze_challenge_scalar()
    challenge_list = [challenges[i] for i in range(cs.num_challenges)]

    theta = transcript.squeeze_challenge_scalar()  # :410

    # ---- lookups: permuted columns (:412-437, lookup/prover.rs:55-140) ----
    class _Lookup:
        pass

    lookups: List[List[_Lookup]] = []
    for ci in range(len(instances)):
        cols, keep = make_eval_columns(pk.fixed_values, advice_values[ci], instance_values[ci], challenge_list, 0, 0,
                                       theta, 0)
        lks = []
        for g_in, g_tab in pk.lookup_compress:
            lk = _Lookup()
            compressed = []
            for g in (g_in, g_tab):
                buf = ctx.alloc(n * 32)
                ctx.memset(buf, 0)
                ctx._check(ctx.lib.h2b_graph_evaluate_lagrange(dom.h, g.compile(ctx), C.byref(cols), buf.ptr))
                compressed.append(buf)
            lk.compressed_input, lk.compressed_table = compressed
            lk.permuted_input, lk.permuted_table = ctx.alloc(n * 32), ctx.alloc(n * 32)
            ctx._check(ctx.lib.h2b_lookup_permute(ctx.h, lk.compressed_input.ptr, lk.compressed_table.ptr,
                                                  unusable_rows_start, lk.permuted_input.ptr, lk.permuted_table.ptr))
            for b in (lk.permuted_input, lk.permuted_table):  # blinding rows (:447-449)
                b.upload(fr_encode([fr_random(rng) for _ in range(bf + 1)]), unusable_rows_start * 32)
            for name in ("permuted_input", "permuted_table"):  # commit_values (:114-125)
                values = getattr(lk, name)
                setattr(lk, name + "_poly", _Poly(ctx, to_coeff(values), n))
                fr_random(rng)  # Blind
            commitments = params.g_lagrange.msm_many([(lk.permuted_input, n), (lk.permuted_table, n)])
            transcript.write_point(commitments[0])
            transcript.write_point(commitments[1])
            lks.append(lk)
        del keep
        lookups.append(lks)
    lap("lookup_permuted")

    beta = transcript.squeeze_challenge_scalar()   # :440
    gamma = transcript.squeeze_challenge_scalar()  # :443
    beta_l, gamma_l = fr_encode([beta]), fr_encode([gamma])

    # ---- permutation argument (permutation/prover.rs:44-190) ----
    chunk_len = cs.degree() - 2
    pcols = cs.permutation.columns

    class _Set:
        pass

    class _Committed:
        pass

    # The permutation products, the lookup products and the vanishing argument's random polynomial are
    # committed between the same two challenges (gamma and y): their commitments are independent, so they are
    # computed together after the last of them exists and written to the transcript in the reference's order.
    deferred: List[tuple] = []   # (bases, Lagrange / coefficient buffer, n, free afterwards?)
    permutations = []
    for ci in range(len(instances)):
        def column_values(c):
            return {ADVICE: advice_values[ci], FIXED: pk.fixed_values, INSTANCE: instance_values[ci]}[c.column_type][c.index]
        last_z = 1
        committed = _Committed()
        committed.sets = []
        for s0 in range(0, len(pcols), chunk_len):
            columns = pcols[s0:s0 + chunk_len]
            vals = _ptr_array([column_values(c) for c in columns])
            sig = _ptr_array(pk.permutations[s0:s0 + chunk_len])
            frac = ctx.alloc(n * 32)
            ctx._check(ctx.lib.h2b_permutation_fractions(dom.h, vals, sig, len(columns), s0,
                                                         C.c_void_p(beta_l.ctypes.data),
                                                         C.c_void_p(gamma_l.ctypes.data), frac.ptr))
            z = ctx.running_product(frac, last_z, n)  # z[0] = last_z, z[i] = z[i-1] * frac[i-1]   (:150-158)
            frac.free()
            blinds = fr_encode([fr_random(rng) for _ in range(bf)])  # :161-163
            z.upload(blinds, (n - bf) * 32)
            last_z = fr_decode(z.download(1, (n - (bf + 1)) * 32))[0]  # :165
            fr_random(rng)  # Blind (:167)
            deferred.append((params.g_lagrange, z, n, True))
            zc = to_coeff(z)
            st = _Set()
            st.poly = _Poly(ctx, zc, n)
            st.permutation_product_coset = ctx.alloc(ext * 32)
            dom.coeff_to_extended_device(zc, st.permutation_product_coset)
            committed.sets.append(st)
        permutations.append(committed)
    lap("permutation_commit")

    # ---- lookups: grand products (:466-475, lookup/prover.rs:146-250) ----
    for lks in lookups:
        for lk in lks:
            frac = ctx.alloc(n * 32)
            ctx._check(ctx.lib.h2b_lookup_product_fractions(
                ctx.h, lk.permuted_input.ptr, lk.permuted_table.ptr, lk.compressed_input.ptr, lk.compressed_table.ptr,
                C.c_void_p(beta_l.ctypes.data), C.c_void_p(gamma_l.ctypes.data), n, frac.ptr))
            z = ctx.running_product(frac, 1, n)  # [1, f0, f0 f1, ...], the first n - bf of them kept (:201-209)
            frac.free()
            z.upload(fr_encode([fr_random(rng) for _ in range(bf)]), (n - bf) * 32)
            fr_random(rng)  # product_blind
            deferred.append((params.g_lagrange, z, n, True))
            lk.product_poly = _Poly(ctx, to_coeff(z), n)
            for b in (lk.permuted_input, lk.permuted_table, lk.compressed_input, lk.compressed_table):
                b.free()
    lap("lookup_product")

    # ---- vanishing argument: random polynomial (vanishing/prover.rs:36-66) ----
    random_poly = _Poly(ctx, fr_random_device(ctx, rng, n), n)
    fr_random(rng)  # random_blind
    lap("random_poly")
    deferred.append((params.g, random_poly.buf, n, False))
    for point in type(params.g).msm_many_mixed([d[:3] for d in deferred]):
        transcript.write_point(point)
    for _, buf, _, release in deferred:
        if release:
            buf.free()
    lap("random_commit")

    y = transcript.squeeze_challenge_scalar()  # :478

    # advice to coefficient form (:481-499)
    advice_polys: List[List[_Poly]] = []
    for adv in advice_values:
        for b in adv:
            dom.lagrange_to_coeff_device(b)
        advice_polys.append([_Poly(ctx, b, n) for b in adv])
    lap("advice_ifft")

    # ---- h(X) (:502-520) ----
    h_ext = pk.ev.evaluate_h(pk, [[p.buf for p in adv] for adv in advice_polys], instance_polys, challenge_list, y,
                             beta, gamma, theta, [[_LookupPolys(lk) for lk in lks] for lks in lookups], permutations)
    for committed in permutations:
        for st in committed.sets:
            st.permutation_product_coset.free()
    lap("evaluate_h")
    # vanishing.construct (vanishing/prover.rs:69-121): divide by t(X), back to coefficients, n-sized pieces
    h_coeff = ctx.alloc(dom.quotient_len * 32)
    dom.extended_to_coeff_device(h_ext, h_coeff, divide_by_vanishing=True)
    h_ext.free()
    n_pieces = dom.quotient_len // n
    for _ in range(n_pieces):
        fr_random(rng)  # h_blinds
    lap("h_to_coeff")
    for point in params.g.msm_many([(h_coeff, n, 0, i * n) for i in range(n_pieces)]):
        transcript.write_point(point)
    lap("h_commit")

    x = transcript.squeeze_challenge_scalar()  # :525
    xn = pow(x, n, R_MOD)

    # ---- evaluations (:548-581) ----
    for adv in advice_polys:
        for c, at in cs.advice_queries:
            transcript.write_scalar(adv[c.index].eval(rot(x, at)))
    fixed_polys = [_Poly(ctx, b, n) for b in pk.fixed_polys]
    for c, at in cs.fixed_queries:
        transcript.write_scalar(fixed_polys[c.index].eval(rot(x, at)))
    # vanishing.evaluate (vanishing/prover.rs:124-152): h_poly = fold(pieces.rev(), acc * xn + piece)
    h_buf = ctx.alloc(n * 32)
    ctx.memset(h_buf, 0)
    xn_l, one_l = fr_encode([xn]), fr_encode([1])
    for i in reversed(range(n_pieces)):
        ctx._check(ctx.lib.h2b_poly_fma(ctx.h, h_buf.ptr, C.c_void_p(xn_l.ctypes.data), h_coeff.at(i * n * 32),
                                        C.c_void_p(one_l.ctypes.data), n))
    h_coeff.free()
    h_poly = _Poly(ctx, h_buf, n)
    transcript.write_scalar(random_poly.eval(x))
    # pk.permutation.evaluate (permutation/prover.rs:208-219)
    sigma_polys = [_Poly(ctx, b, n) for b in pk.permutation_polys]
    for p in sigma_polys:
        transcript.write_scalar(p.eval(x))
    # permutation product evaluations (permutation/prover.rs:222-266)
    x_next, x_last = rot(x, 1), rot(x, -(bf + 1))
    for committed in permutations:
        for si, st in enumerate(committed.sets):
            transcript.write_scalar(st.poly.eval(x))
            transcript.write_scalar(st.poly.eval(x_next))
            if si + 1 < len(committed.sets):
                transcript.write_scalar(st.poly.eval(x_last))
    # lookup evaluations (:588-595, lookup/prover.rs:253-283)
    x_inv = rot(x, -1)
    for lks in lookups:
        for lk in lks:
            for poly, pt in ((lk.product_poly, x), (lk.product_poly, x_next), (lk.permuted_input_poly, x),
                             (lk.permuted_input_poly, x_inv), (lk.permuted_table_poly, x)):
                transcript.write_scalar(poly.eval(pt))
    lap("evals")

    # ---- the opening queries in the reference's order (:596-645) ----
    queries: List[Tuple[int, _Poly]] = []
    for ci in range(len(instances)):
        for c, at in cs.advice_queries:
            queries.append((rot(x, at), advice_polys[ci][c.index]))
        sets = permutations[ci].sets
        for st in sets:
            queries.append((x, st.poly))
            queries.append((x_next, st.poly))
        for st in list(reversed(sets))[1:]:
            queries.append((x_last, st.poly))
        for lk in lookups[ci]:  # lookup/prover.rs:286-323
            queries.append((x, lk.product_poly))
            queries.append((x, lk.permuted_input_poly))
            queries.append((x, lk.permuted_table_poly))
            queries.append((x_inv, lk.permuted_input_poly))
            queries.append((x_next, lk.product_poly))
    for c, at in cs.fixed_queries:
        queries.append((rot(x, at), fixed_polys[c.index]))
    for p in sigma_polys:
        queries.append((x, p))
    queries.append((x, h_poly))
    queries.append((x, random_poly))
    (prover or ProverGWC)(params).create_proof(rng, transcript, queries)
    if replicated is not None:
        replicated("the transcript at the end of create_proof (witness and rng must be identical on every rank)",
                   transcript.state_digest())
    lap("multiopen")

    # per-proof device buffers
    for adv in advice_values:
        for b in adv:
            b.free()
    for lst in instance_values + instance_polys:
        for b in lst:
            b.free()
    for committed in permutations:
        for st in committed.sets:
            st.poly.buf.free()
    for lks in lookups:
        for lk in lks:
            for p in (lk.product_poly, lk.permuted_input_poly, lk.permuted_table_poly):
                p.buf.free()
    random_poly.buf.free()
    h_buf.free()


class ProverGWC:
    """poly/kzg/multiopen/gwc/prover.rs:24-92; QUERY_INSTANCE = false."""

    QUERY_INSTANCE = False

    def __init__(self, params: ParamsKZG):
        self.params = params

    def create_proof(self, rng, transcript: Blake2bWrite, queries: Sequence[Tuple[int, "_Poly"]]) -> None:
        ctx = self.params.ctx
        v = transcript.squeeze_challenge_scalar()
        # construct_intermediate_sets (gwc.rs:36-61): by point, first-occurrence order
        sets: List[Tuple[int, list]] = []
        for q in queries:
            for point, qs in sets:
                if point == q[0]:
                    qs.append(q)
                    break
            else:
                sets.append((q[0], [q]))
        one_l = fr_encode([1])
        witnesses = []
        for z, qs in sets:
            n = qs[0][1].n
            poly_batch = ctx.clone(qs[0][1].buf, n * 32)  # power_of_v = 1
            eval_batch = qs[0][1].eval(z)
            power = 1
            for _, p in qs[1:]:
                power = power * v % R_MOD
                pw = fr_encode([power])
                ctx._check(ctx.lib.h2b_poly_fma(ctx.h, poly_batch.ptr, C.c_void_p(one_l.ctypes.data), p.buf.ptr,
                                                C.c_void_p(pw.ctypes.data), n))
                eval_batch = (eval_batch + p.eval(z) * power) % R_MOD
            # &poly_batch - eval_batch: the constant coefficient (poly.rs:298-305)
            c0 = fr_decode(poly_batch.download(1))[0]
            poly_batch.upload(fr_encode([(c0 - eval_batch) % R_MOD]))
            witnesses.append((ctx.kate_division(poly_batch, z, n), n - 1))
            poly_batch.free()
        # the witness commitments depend on no challenge drawn in between: committed together (gwc/prover.rs:80-88)
        for point in self.params.g.msm_many(witnesses):
            transcript.write_point(point)
        for w, _ in witnesses:
            w.free()


def lagrange_interpolate(points: Sequence[int], evals: Sequence[int]) -> List[int]:
    """arithmetic.rs:405-458 on a handful of points (host integers)."""
    if len(points) != len(evals):
        raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(points.len(), evals.len())")
    if len(points) == 1:
        return [evals[0] % R_MOD]
    final = [0] * len(points)
    for j, xj in enumerate(points):
        tmp = [1]
        for k, xk in enumerate(points):
            if k != j:
                denom = pow((xj - xk) % R_MOD, -1, R_MOD)
                tmp = [(a * (-denom * xk) + b * denom) % R_MOD for a, b in zip(tmp + [0], [0] + tmp)]
        final = [(f + c * evals[j]) % R_MOD for f, c in zip(final, tmp)]
    return final


def evaluate_vanishing_polynomial(roots: Sequence[int], z: int) -> int:
    """arithmetic.rs:460-478"""
    acc = 1
    for r in roots:
        acc = acc * (z - r) % R_MOD
    return acc


class ProverSHPLONK:
    """poly/kzg/multiopen/shplonk/prover.rs:94-285 with the rotation sets of shplonk.rs:55-134."""

    QUERY_INSTANCE = False

    def __init__(self, params: ParamsKZG):
        self.params = params

    @staticmethod
    def construct_intermediate_sets(queries):
        """-> ([(points ascending, [(poly, evals at those points)])], super_point_set ascending).
        Polynomials are told apart by identity, as the reference's PolynomialPointer equality does."""
        super_points = sorted({pt for pt, _ in queries})
        by_commitment: List[Tuple["_Poly", set]] = []
        for pt, p in queries:
            for poly, rots in by_commitment:
                if poly is p:
                    rots.add(pt)
                    break
            else:
                by_commitment.append((p, {pt}))
        by_set: List[Tuple[set, list]] = []
        for poly, rots in by_commitment:
            for rs, polys in by_set:
                if rs == rots:
                    polys.append(poly)
                    break
            else:
                by_set.append((rots, [poly]))
        out = []
        for rots, polys in by_set:
            pts = sorted(rots)
            out.append((pts, [(p, [p.eval(pt) for pt in pts]) for p in polys]))
        return out, super_points

    def create_proof(self, rng, transcript: Blake2bWrite, queries: Sequence[Tuple[int, "_Poly"]]) -> None:
        ctx, n = self.params.ctx, self.params.n
        lib = ctx.lib
        one_l = fr_encode([1])

        def fma(acc: DeviceBuffer, a: int, p: DeviceBuffer, b: int, count: int) -> None:
            al, bl = fr_encode([a]), fr_encode([b])
            ctx._check(lib.h2b_poly_fma(ctx.h, acc.ptr, C.c_void_p(al.ctypes.data), p.ptr, C.c_void_p(bl.ctypes.data),
                                        count))

        def sub_low(buf: DeviceBuffer, low: Sequence[int]) -> None:
            """buf[0..len(low)) -= low (the first coefficients only)."""
            head = fr_decode(buf.download(len(low)))
            buf.upload(fr_encode([(a - b) % R_MOD for a, b in zip(head, low)]))

        y = transcript.squeeze_challenge_scalar()
        rotation_sets, super_points = self.construct_intermediate_sets(queries)
        extended = [(pts, [(p, lagrange_interpolate(pts, evals)) for p, evals in coms]) for pts, coms in rotation_sets]
        v = transcript.squeeze_challenge_scalar()

        # h(X) = sum_i v^i * (sum_j y^j (p_j - r_j)) / Z_i   (:120-177)
        h_x = ctx.alloc(n * 32)
        ctx.memset(h_x, 0)
        pv = 1
        for pts, coms in extended:
            n_x = ctx.alloc(n * 32)
            ctx.memset(n_x, 0)
            py = 1
            for p, low in coms:
                num = ctx.clone(p.buf, n * 32)
                sub_low(num, low)
                fma(n_x, 1, num, py, n)
                num.free()
                py = py * y % R_MOD
            length = n
            for pt in pts:  # div_by_vanishing: one Kate division per point
                q_ = ctx.kate_division(n_x, pt, length)
                n_x.free()
                n_x = q_
                length -= 1
            fma(h_x, 1, n_x, pv, length)  # poly.resize(n, zero) * v^i, accumulated
            n_x.free()
            pv = pv * v % R_MOD
        transcript.write_point(self.params.g.msm(h_x, n))
        u = transcript.squeeze_challenge_scalar()

        # l(X) = sum_i v^i z_i(u) sum_j y^j (p_j - r_j(u)) - Z_T(u) h(X)   (:183-232)
        l_x = ctx.alloc(n * 32)
        ctx.memset(l_x, 0)
        z_diffs, pv = [], 1
        for pts, coms in extended:
            z_i = evaluate_vanishing_polynomial([p for p in super_points if p not in pts], u)
            py = 1
            const = 0
            for p, low in coms:
                r_eval = 0
                for c in reversed(low):  # eval_polynomial of the low-degree equivalent at u
                    r_eval = (r_eval * u + c) % R_MOD
                w = py * z_i % R_MOD * pv % R_MOD
                fma(l_x, 1, p.buf, w, n)
                const = (const + r_eval * w) % R_MOD
                py = py * y % R_MOD
            c0 = fr_decode(l_x.download(1))[0]
            l_x.upload(fr_encode([(c0 - const) % R_MOD]))
            z_diffs.append(z_i)
            pv = pv * v % R_MOD
        zt_eval = evaluate_vanishing_polynomial(super_points, u)
        fma(l_x, 1, h_x, -zt_eval % R_MOD, n)
        h_x.free()
        h2 = ctx.kate_division(l_x, u, n)  # div_by_vanishing(l_x, [u])   (:240)
        l_x.free()
        ctx.poly_scale(h2, pow(z_diffs[0], -1, R_MOD), n - 1)
        transcript.write_point(self.params.g.msm(h2, n - 1))
        h2.free()
