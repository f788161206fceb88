"""ORACLE (test infrastructure, never shipped, never on the product path).

Big-integer restatement of the quotient evaluation of halo2_proofs
(/root/reference/halo2_proofs/src/plonk/evaluation.rs:280-522), written the slow, direct way:
every constraint expression is evaluated from its tree at every row (the reference's own
`evaluate`, evaluation.rs:749-787) and folded with y, instead of through the GraphEvaluator
the product compiles.  The two therefore only agree if the graph compiler, the slot renaming
and the device interpreter are all right.

Expressions are nested tuples:
  ("constant", v) | ("fixed"|"advice"|"instance", column_index, rotation) | ("challenge", i)
  | ("negated", e) | ("sum", a, b) | ("product", a, b) | ("scaled", e, f)

PARITY: the field / curve layer underneath is pinned against the reference's golden verifying key through its
Vesta instance (see oracle/bn256.py, oracle/pasta.py).  The constraint folding of this file has no reference-held
output (a proof needs the rng); it is pinned by the mathematical definition of the constraints it folds
(tests/test_oracle.py checks h vanishes on the domain for a satisfied circuit, i.e. that it is divisible by X^n - 1).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

from typing import List, Sequence

from .bn256 import MULT_GEN, R_MOD, S, ZETA

DELTA = pow(MULT_GEN, 1 << S, R_MOD)  # Fr::DELTA = g^(2^S): 7^(2^28) for bn256 (halo2curves bn256/fr.rs)
ADVICE, FIXED, INSTANCE = 0, 1, 2  # plonk/circuit.rs `Any`


def get_rotation_idx(idx: int, rot: int, rot_scale: int, isize: int) -> int:
    """evaluation.rs:32-34"""
    return (idx + rot * rot_scale) % isize


def evaluate_expression(expr, idx, rot_scale, isize, fixed, advice, instance, challenges) -> int:
    """evaluation.rs:749-787 for one row."""
    k = expr[0]
    if k == "constant":
        return expr[1] % R_MOD
    if k == "fixed":
        return fixed[expr[1]][get_rotation_idx(idx, expr[2], rot_scale, isize)]
    if k == "advice":
        return advice[expr[1]][get_rotation_idx(idx, expr[2], rot_scale, isize)]
    if k == "instance":
        return instance[expr[1]][get_rotation_idx(idx, expr[2], rot_scale, isize)]
    if k == "challenge":
        return challenges[expr[1]]
    ev = lambda e: evaluate_expression(e, idx, rot_scale, isize, fixed, advice, instance, challenges)  # noqa: E731
    if k == "negated":
        return -ev(expr[1]) % R_MOD
    if k == "sum":
        return (ev(expr[1]) + ev(expr[2])) % R_MOD
    if k == "product":
        return ev(expr[1]) * ev(expr[2]) % R_MOD
    if k == "scaled":
        return ev(expr[1]) * expr[2] % R_MOD
    raise ValueError(k)


def evaluate_h(*, k: int, extended_k: int, extended_omega: int, gates: Sequence[Sequence], lookups: Sequence,
               perm_columns: Sequence, chunk_len: int, blinding_factors: int, fixed: Sequence[List[int]],
               l0: List[int], l_last: List[int], l_active_row: List[int], sigma_cosets: Sequence[List[int]],
               circuits: Sequence[dict], challenges: Sequence[int], y: int, beta: int, gamma: int,
               theta: int) -> List[int]:
    """evaluation.rs:280-522.  Everything is already in extended-Lagrange form (lists of 2^extended_k ints).
    gates: list of lists of expressions (gate.polynomials()); lookups: list of (input_exprs, table_exprs);
    perm_columns: [(column_type, index)]; circuits: per circuit instance a dict with `advice`, `instance`
    (cosets), `perm_sets` (z cosets) and `lookups` (dicts product / permuted_input / permuted_table)."""
    size = 1 << extended_k
    rot_scale = 1 << (extended_k - k)
    values = [0] * size
    for circ in circuits:
        advice, instance = circ["advice"], circ["instance"]
        # custom gates (:335-362): Horner over all gate polynomials with y, seeded with the previous value
        for idx in range(size):
            v = values[idx]
            for polys in gates:
                for poly in polys:
                    v = (v * y + evaluate_expression(poly, idx, rot_scale, size, fixed, advice, instance,
                                                     challenges)) % R_MOD
            values[idx] = v
        # permutations (:364-444)
        sets = circ["perm_sets"]
        if sets:
            last_rotation = -(blinding_factors + 1)
            delta_start = beta * ZETA % R_MOD

            def column(c):
                return {ADVICE: advice, FIXED: fixed, INSTANCE: instance}[c[0]][c[1]]

            beta_term = 1
            for idx in range(size):
                r_next = get_rotation_idx(idx, 1, rot_scale, size)
                r_last = get_rotation_idx(idx, last_rotation, rot_scale, size)
                v = values[idx]
                v = (v * y + (1 - sets[0][idx]) * l0[idx]) % R_MOD
                zl = sets[-1][idx]
                v = (v * y + (zl * zl - zl) * l_last[idx]) % R_MOD
                for s in range(1, len(sets)):
                    v = (v * y + (sets[s][idx] - sets[s - 1][r_last]) * l0[idx]) % R_MOD
                current_delta = delta_start * beta_term % R_MOD
                for s, z in enumerate(sets):
                    cols = perm_columns[s * chunk_len:(s + 1) * chunk_len]
                    sig = sigma_cosets[s * chunk_len:(s + 1) * chunk_len]
                    left = z[r_next]
                    for c, perm in zip(cols, sig):
                        left = left * (column(c)[idx] + beta * perm[idx] + gamma) % R_MOD
                    right = z[idx]
                    for c in cols:
                        right = right * (column(c)[idx] + current_delta + gamma) % R_MOD
                        current_delta = current_delta * DELTA % R_MOD
                    v = (v * y + (left - right) * l_active_row[idx]) % R_MOD
                values[idx] = v
                beta_term = beta_term * extended_omega % R_MOD
        # lookups (:446-519)
        for (input_exprs, table_exprs), lk in zip(lookups, circ["lookups"]):
            product, a_, s_ = lk["product"], lk["permuted_input"], lk["permuted_table"]
            for idx in range(size):
                def compress(exprs):
                    acc = 0
                    for e in exprs:
                        acc = (acc * theta + evaluate_expression(e, idx, rot_scale, size, fixed, advice, instance,
                                                                 challenges)) % R_MOD
                    return acc
                table_value = (compress(input_exprs) + beta) * (compress(table_exprs) + gamma) % R_MOD
                r_next = get_rotation_idx(idx, 1, rot_scale, size)
                r_prev = get_rotation_idx(idx, -1, rot_scale, size)
                a_minus_s = (a_[idx] - s_[idx]) % R_MOD
                v = values[idx]
                v = (v * y + (1 - product[idx]) * l0[idx]) % R_MOD
                v = (v * y + (product[idx] * product[idx] - product[idx]) * l_last[idx]) % R_MOD
                v = (v * y + (product[r_next] * (a_[idx] + beta) * (s_[idx] + gamma) - product[idx] * table_value)
                     * l_active_row[idx]) % R_MOD
                v = (v * y + a_minus_s * l0[idx]) % R_MOD
                v = (v * y + a_minus_s * (a_[idx] - a_[r_prev]) * l_active_row[idx]) % R_MOD
                values[idx] = v
    return values
