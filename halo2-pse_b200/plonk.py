"""Host-side mirror of the reference's constraint-system and quotient-evaluation interface.

Names, argument meaning and behaviour follow halo2_proofs (paths relative to
/root/reference/halo2_proofs/src):

* ``Expression``, ``ConstraintSystem``   -- plonk/circuit.rs:780-1100, 1330-2060 (the part a prover needs:
  columns, queries, gates, lookups, the permutation argument, degree(), blinding_factors())
* ``GraphEvaluator``, ``Evaluator``      -- plonk/evaluation.rs:183-746
* ``Evaluator.evaluate_h``               -- plonk/evaluation.rs:280-522

This is bookkeeping only: expression trees are compiled into the word stream of include/halo2_b200.h
and every field operation over rows runs in libhalo2b200 on the GPU (no CPU fallback).  Field
constants are canonical Python integers mod r.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from ._ffi import H2B_DEVICE, H2BError
from .api import R_MOD, Context, DeviceBuffer, EvaluationDomain, fr_encode

# Any::{Advice, Fixed, Instance} in declaration order (plonk/circuit.rs `Any`)
ADVICE, FIXED, INSTANCE = 0, 1, 2


class Column(tuple):
    """Column<Any>: (column_type, index)."""

    def __new__(cls, column_type: int, index: int):
        return super().__new__(cls, (column_type, index))

    @property
    def column_type(self) -> int:
        return self[0]

    @property
    def index(self) -> int:
        return self[1]


class Expression:
    """plonk/circuit.rs:780-1100 without `Selector` (the prover works from the constraint system of
    the verifying key, where selectors are already fixed columns: plonk/prover.rs:69-71).
    node: ("constant", v) | ("fixed"|"advice"|"instance", column_index, rotation, query_index)
          | ("challenge", index, phase)
          | ("negated", e) | ("sum", a, b) | ("product", a, b) | ("scaled", e, f)"""

    __slots__ = ("node",)

    def __init__(self, *node):
        self.node = node

    @staticmethod
    def constant(v: int) -> "Expression":
        return Expression("constant", v % R_MOD)

    def __neg__(self):
        return Expression("negated", self)

    def __add__(self, rhs: "Expression"):
        return Expression("sum", self, rhs)

    def __sub__(self, rhs: "Expression"):  # impl Sub: self + (-rhs)
        return Expression("sum", self, Expression("negated", rhs))

    def __mul__(self, rhs):
        if isinstance(rhs, Expression):
            return Expression("product", self, rhs)
        return Expression("scaled", self, int(rhs) % R_MOD)  # impl Mul<F>

    def square(self):
        return Expression("product", self, self)

    def degree(self) -> int:  # plonk/circuit.rs:1002-1015
        k = self.node[0]
        if k in ("constant", "challenge"):
            return 0
        if k in ("fixed", "advice", "instance"):
            return 1
        if k in ("negated", "scaled"):
            return self.node[1].degree()
        if k == "sum":
            return max(self.node[1].degree(), self.node[2].degree())
        return self.node[1].degree() + self.node[2].degree()

    def to_tuple(self):
        """Nested tuples (plain data, for tests and serialisation)."""
        k = self.node[0]
        if k in ("negated",):
            return (k, self.node[1].to_tuple())
        if k in ("sum", "product"):
            return (k, self.node[1].to_tuple(), self.node[2].to_tuple())
        if k == "scaled":
            return (k, self.node[1].to_tuple(), self.node[2])
        return tuple(self.node)


class LookupArgument:
    """plonk/lookup.rs:10-60"""

    def __init__(self, name: str, table_map: Sequence[Tuple[Expression, Expression]]):
        self.name = name
        self.input_expressions = [a for a, _ in table_map]
        self.table_expressions = [b for _, b in table_map]

    def required_degree(self) -> int:
        input_degree = max([1] + [e.degree() for e in self.input_expressions])
        table_degree = max([1] + [e.degree() for e in self.table_expressions])
        return max(4, 2 + input_degree + table_degree)


class PermutationArgument:
    """plonk/permutation.rs:19-75"""

    def __init__(self):
        self.columns: List[Column] = []

    def required_degree(self) -> int:
        return 3

    def add_column(self, column: Column) -> None:
        if column not in self.columns:
            self.columns.append(column)


class ConstraintSystem:
    """The fields of plonk/circuit.rs:1330-1400 a prover reads, and the `configure`-time methods that fill
    them (plonk/circuit.rs:1516-1640, 1760-1800, 1974-2031).  No selectors, regions or floor planner:
    fixed columns are assigned directly."""

    def __init__(self):
        self.num_fixed_columns = 0
        self.num_advice_columns = 0
        self.num_instance_columns = 0
        self.num_challenges = 0
        self.advice_column_phase: List[int] = []
        self.challenge_phase: List[int] = []
        self.gates: List[Tuple[str, List[Expression]]] = []
        self.advice_queries: List[Tuple[Column, int]] = []
        self.num_advice_queries: List[int] = []
        self.instance_queries: List[Tuple[Column, int]] = []
        self.fixed_queries: List[Tuple[Column, int]] = []
        self.permutation = PermutationArgument()
        self.lookups: List[LookupArgument] = []
        self.minimum_degree: Optional[int] = None

    # columns
    def advice_column(self, phase: int = 0) -> Column:
        c = Column(ADVICE, self.num_advice_columns)
        self.num_advice_columns += 1
        self.num_advice_queries.append(0)
        self.advice_column_phase.append(phase)
        return c

    def fixed_column(self) -> Column:
        c = Column(FIXED, self.num_fixed_columns)
        self.num_fixed_columns += 1
        return c

    def instance_column(self) -> Column:
        c = Column(INSTANCE, self.num_instance_columns)
        self.num_instance_columns += 1
        return c

    def challenge_usable_after(self, phase: int) -> Expression:
        idx = self.num_challenges
        self.num_challenges += 1
        self.challenge_phase.append(phase)
        return Expression("challenge", idx, phase)

    def phases(self) -> List[int]:
        return list(range(max([0] + self.advice_column_phase) + 1))

    # queries (plonk/circuit.rs:1571-1627)
    def _query_index(self, column: Column, at: int) -> int:
        qs = {ADVICE: self.advice_queries, FIXED: self.fixed_queries, INSTANCE: self.instance_queries}[column.column_type]
        for i, q in enumerate(qs):
            if q == (column, at):
                return i
        qs.append((column, at))
        if column.column_type == ADVICE:
            self.num_advice_queries[column.index] += 1
        return len(qs) - 1

    def query_advice(self, column: Column, at: int = 0) -> Expression:
        return Expression("advice", column.index, at, self._query_index(column, at))

    def query_fixed(self, column: Column, at: int = 0) -> Expression:
        return Expression("fixed", column.index, at, self._query_index(column, at))

    def query_instance(self, column: Column, at: int = 0) -> Expression:
        return Expression("instance", column.index, at, self._query_index(column, at))

    def enable_equality(self, column: Column) -> None:  # :1516-1520
        self._query_index(column, 0)
        self.permutation.add_column(column)

    def create_gate(self, name: str, polys: Sequence[Expression]) -> None:
        assert polys, "Gates must contain at least one constraint."
        self.gates.append((name, list(polys)))

    def lookup(self, name: str, table_map: Sequence[Tuple[Expression, Expression]]) -> int:
        self.lookups.append(LookupArgument(name, table_map))
        return len(self.lookups) - 1

    def set_minimum_degree(self, degree: int) -> None:
        self.minimum_degree = degree

    def degree(self) -> int:  # :1974-2002
        degree = self.permutation.required_degree()
        degree = max(degree, max([1] + [l.required_degree() for l in self.lookups]))
        degree = max(degree, max([0] + [p.degree() for _, polys in self.gates for p in polys]))
        return max(degree, self.minimum_degree or 1)

    def blinding_factors(self) -> int:  # :2006-2031
        factors = max(self.num_advice_queries) if self.num_advice_queries else 1
        return max(3, factors) + 2

    def minimum_rows(self) -> int:
        return self.blinding_factors() + 3


# --------------------------------------------------------------------------
# GraphEvaluator (plonk/evaluation.rs:183-746)
# --------------------------------------------------------------------------
# ValueSource as (variant, a, b): tuple order == the derived PartialOrd of the reference's enum (:37-61)
VS_CONSTANT, VS_INTERMEDIATE, VS_FIXED, VS_ADVICE, VS_INSTANCE, VS_CHALLENGE, VS_BETA, VS_GAMMA, VS_THETA, \
    VS_Y, VS_PREVIOUS = range(11)
# Calculation variants (:110-127)
C_ADD, C_SUB, C_MUL, C_SQUARE, C_DOUBLE, C_NEGATE, C_HORNER, C_STORE = range(8)


def _vs(kind: int, a: int = 0, b: int = 0) -> tuple:
    return (kind, a, b)


class GraphEvaluator:
    def __init__(self):
        self.constants: List[int] = [0, 1, 2]  # fixed positions (:525-538)
        self.rotations: List[int] = []
        self.calculations: List[Tuple[tuple, int]] = []  # (calculation, target)
        self.num_intermediates = 0
        self._graph = None  # compiled h2b_graph handle

    def add_rotation(self, rotation: int) -> int:
        if rotation in self.rotations:
            return self.rotations.index(rotation)
        self.rotations.append(rotation)
        return len(self.rotations) - 1

    def add_constant(self, constant: int) -> tuple:
        constant %= R_MOD
        if constant in self.constants:
            return _vs(VS_CONSTANT, self.constants.index(constant))
        self.constants.append(constant)
        return _vs(VS_CONSTANT, len(self.constants) - 1)

    def add_calculation(self, calculation: tuple) -> tuple:
        for calc, target in self.calculations:
            if calc == calculation:
                return _vs(VS_INTERMEDIATE, target)
        target = self.num_intermediates
        self.calculations.append((calculation, target))
        self.num_intermediates += 1
        return _vs(VS_INTERMEDIATE, target)

    def add_expression(self, expr: Expression) -> tuple:  # :590-690
        k = expr.node[0]
        zero, one, two = _vs(VS_CONSTANT, 0), _vs(VS_CONSTANT, 1), _vs(VS_CONSTANT, 2)
        if k == "constant":
            return self.add_constant(expr.node[1])
        if k in ("fixed", "advice", "instance"):
            rot_idx = self.add_rotation(expr.node[2])
            kind = {"fixed": VS_FIXED, "advice": VS_ADVICE, "instance": VS_INSTANCE}[k]
            return self.add_calculation((C_STORE, _vs(kind, expr.node[1], rot_idx)))
        if k == "challenge":
            return self.add_calculation((C_STORE, _vs(VS_CHALLENGE, expr.node[1])))
        if k == "negated":
            a = expr.node[1]
            if a.node[0] == "constant":
                return self.add_constant(-a.node[1])
            result_a = self.add_expression(a)
            if result_a == zero:
                return result_a
            return self.add_calculation((C_NEGATE, result_a))
        if k == "sum":
            a, b = expr.node[1], expr.node[2]
            if b.node[0] == "negated":  # undo subtraction stored as a + (-b)
                result_a = self.add_expression(a)
                result_b = self.add_expression(b.node[1])
                if result_a == zero:
                    return self.add_calculation((C_NEGATE, result_b))
                if result_b == zero:
                    return result_a
                return self.add_calculation((C_SUB, result_a, result_b))
            result_a = self.add_expression(a)
            result_b = self.add_expression(b)
            if result_a == zero:
                return result_b
            if result_b == zero:
                return result_a
            if result_a <= result_b:
                return self.add_calculation((C_ADD, result_a, result_b))
            return self.add_calculation((C_ADD, result_b, result_a))
        if k == "product":
            result_a = self.add_expression(expr.node[1])
            result_b = self.add_expression(expr.node[2])
            if result_a == zero or result_b == zero:
                return zero
            if result_a == one:
                return result_b
            if result_b == one:
                return result_a
            if result_a == two:
                return self.add_calculation((C_DOUBLE, result_b))
            if result_b == two:
                return self.add_calculation((C_DOUBLE, result_a))
            if result_a == result_b:
                return self.add_calculation((C_SQUARE, result_a))
            if result_a <= result_b:
                return self.add_calculation((C_MUL, result_a, result_b))
            return self.add_calculation((C_MUL, result_b, result_a))
        if k == "scaled":
            f = expr.node[2]
            if f == 0:
                return zero
            if f == 1:
                return self.add_expression(expr.node[1])
            cst = self.add_constant(f)
            result_a = self.add_expression(expr.node[1])
            return self.add_calculation((C_MUL, result_a, cst))
        raise ValueError(k)

    # ---- the boundary: flatten for h2b_graph_new (include/halo2_b200.h) ----
    def encode(self) -> np.ndarray:
        words: List[int] = []
        for calc, target in self.calculations:
            op = calc[0]
            words += [op, target]
            if op == C_HORNER:
                start, parts, factor = calc[1], calc[2], calc[3]
                words += list(start) + list(factor) + [len(parts)]
                for p in parts:
                    words += list(p)
            else:
                for src in calc[1:]:
                    words += list(src)
        return np.asarray(words, dtype=np.uint32)

    def compile(self, ctx: Context):
        if self._graph is not None and self._graph[0] is ctx:
            return self._graph[1]
        code = self.encode()
        consts = fr_encode(self.constants)
        rots = np.asarray(self.rotations, dtype=np.int32)
        h = C.c_void_p()
        ctx._check(ctx.lib.h2b_graph_new(ctx.h, C.c_void_p(code.ctypes.data), code.size,
                                         C.c_void_p(consts.ctypes.data), len(self.constants),
                                         C.c_void_p(rots.ctypes.data), len(self.rotations),
                                         self.num_intermediates, C.byref(h)))
        self._graph = (ctx, h)
        return h

    def free(self) -> None:
        if self._graph is not None:
            ctx, h = self._graph
            if ctx.h is not None:
                ctx.lib.h2b_graph_free(h)
            self._graph = None


class _EvalColumns(C.Structure):
    """h2b_eval_columns (include/halo2_b200.h)"""
    _fields_ = [("fixed", C.POINTER(C.c_void_p)), ("n_fixed", C.c_uint32),
                ("advice", C.POINTER(C.c_void_p)), ("n_advice", C.c_uint32),
                ("instance", C.POINTER(C.c_void_p)), ("n_instance", C.c_uint32),
                ("challenges", C.c_void_p), ("n_challenges", C.c_uint32),
                ("beta", C.c_uint64 * 4), ("gamma", C.c_uint64 * 4), ("theta", C.c_uint64 * 4),
                ("y", C.c_uint64 * 4)]


def _ptr_array(bufs: Sequence) -> "C.Array":
    arr = (C.c_void_p * max(len(bufs), 1))()
    for i, b in enumerate(bufs):
        arr[i] = b.ptr.value if isinstance(b, DeviceBuffer) else int(b)
    return arr


def make_eval_columns(fixed, advice, instance, challenges: Sequence[int], beta: int, gamma: int, theta: int,
                      y: int):
    """-> (struct, keepalive).  Columns are DeviceBuffers (or raw device addresses) of 2^extended_k elements."""
    keep = [_ptr_array(fixed), _ptr_array(advice), _ptr_array(instance), fr_encode(list(challenges) or [0])]
    s = _EvalColumns()
    s.fixed, s.n_fixed = C.cast(keep[0], C.POINTER(C.c_void_p)), len(fixed)
    s.advice, s.n_advice = C.cast(keep[1], C.POINTER(C.c_void_p)), len(advice)
    s.instance, s.n_instance = C.cast(keep[2], C.POINTER(C.c_void_p)), len(instance)
    s.challenges, s.n_challenges = keep[3].ctypes.data, len(challenges)
    for name, v in (("beta", beta), ("gamma", gamma), ("theta", theta), ("y", y)):
        limbs = fr_encode([v])[0]
        setattr(s, name, (C.c_uint64 * 4)(*[int(x) for x in limbs]))
    return s, keep


class Evaluator:
    """plonk/evaluation.rs:183-277: custom_gates + one graph per lookup."""

    def __init__(self, cs: ConstraintSystem):
        self.custom_gates = GraphEvaluator()
        self.lookups: List[GraphEvaluator] = []
        parts = []
        for _, polys in cs.gates:
            parts += [self.custom_gates.add_expression(p) for p in polys]
        self.custom_gates.add_calculation((C_HORNER, _vs(VS_PREVIOUS), tuple(parts), _vs(VS_Y)))
        for lookup in cs.lookups:
            graph = GraphEvaluator()

            def evaluate_lc(expressions):
                ps = tuple(graph.add_expression(e) for e in expressions)
                return graph.add_calculation((C_HORNER, _vs(VS_CONSTANT, 0), ps, _vs(VS_THETA)))

            compressed_input_coset = evaluate_lc(lookup.input_expressions)
            compressed_table_coset = evaluate_lc(lookup.table_expressions)
            right_gamma = graph.add_calculation((C_ADD, compressed_table_coset, _vs(VS_GAMMA)))
            lc = graph.add_calculation((C_ADD, compressed_input_coset, _vs(VS_BETA)))
            graph.add_calculation((C_MUL, lc, right_gamma))
            self.lookups.append(graph)

    def evaluate_h(self, pk, advice_polys, instance_polys, challenges, y, beta, gamma, theta, lookups,
                   permutations) -> DeviceBuffer:
        """plonk/evaluation.rs:280-522.  `pk` carries domain, cs, fixed_cosets, l0, l_last, l_active_row and
        permutation cosets as DeviceBuffers; advice_polys / instance_polys are per-circuit lists of
        coefficient-form DeviceBuffers; lookups[i][n] has product_poly / permuted_input_poly /
        permuted_table_poly; permutations[i].sets[s].permutation_product_coset.  Returns the extended-domain
        values as a DeviceBuffer."""
        domain: EvaluationDomain = pk.domain
        ctx: Context = domain.ctx
        cs: ConstraintSystem = pk.cs
        ext = domain.extended_len()
        values = ctx.alloc(ext * 32)
        ctx.memset(values, 0)
        for advice_c, instance_c, lookups_c, permutation in zip(advice_polys, instance_polys, lookups, permutations):
            # the advice and instance cosets (:305-323)
            advice = [_to_extended(domain, p) for p in advice_c]
            instance = [_to_extended(domain, p) for p in instance_c]
            cols, keep = make_eval_columns(pk.fixed_cosets, advice, instance, challenges, beta, gamma, theta, y)
            # custom gates (:335-362)
            ctx._check(ctx.lib.h2b_evaluate_h_gates(domain.h, self.custom_gates.compile(ctx), C.byref(cols),
                                                    values.ptr))
            # permutations (:364-444)
            sets = permutation.sets
            if sets:
                pc = cs.permutation.columns
                ctype = np.asarray([c.column_type for c in pc], dtype=np.uint32)
                cidx = np.asarray([c.index for c in pc], dtype=np.uint32)
                sigma = _ptr_array(pk.permutation_cosets)
                zs = _ptr_array([s.permutation_product_coset for s in sets])
                ctx._check(ctx.lib.h2b_evaluate_h_permutation(
                    domain.h, C.byref(cols), C.c_void_p(ctype.ctypes.data), C.c_void_p(cidx.ctypes.data), len(pc),
                    sigma, zs, len(sets), cs.degree() - 2, cs.blinding_factors(), pk.l0.ptr, pk.l_last.ptr,
                    pk.l_active_row.ptr, values.ptr))
            # lookups (:446-519): the three cosets live only while their lookup is evaluated
            for n, lookup in enumerate(lookups_c):
                product_coset = _to_extended(domain, lookup.product_poly)
                permuted_input_coset = _to_extended(domain, lookup.permuted_input_poly)
                permuted_table_coset = _to_extended(domain, lookup.permuted_table_poly)
                ctx._check(ctx.lib.h2b_evaluate_h_lookup(
                    domain.h, self.lookups[n].compile(ctx), C.byref(cols), product_coset.ptr,
                    permuted_input_coset.ptr, permuted_table_coset.ptr, pk.l0.ptr, pk.l_last.ptr,
                    pk.l_active_row.ptr, values.ptr))
                for b in (product_coset, permuted_input_coset, permuted_table_coset):
                    b.free()
            for b in advice + instance:
                b.free()
            del keep
        return values

    def free(self) -> None:
        for g in [self.custom_gates] + self.lookups:
            g.free()


def _to_extended(domain: EvaluationDomain, coeff: DeviceBuffer) -> DeviceBuffer:
    """domain.coeff_to_extended(poly.clone()) on the device."""
    out = domain.ctx.alloc(domain.extended_len() * 32)
    domain.coeff_to_extended_device(coeff, out)
    return out
