"""The reference's benchmark circuit (benches/plonk.rs:29-270, `MyCircuit` over `StandardPlonk`), as the
constraint system `configure` builds and the cells `synthesize` assigns under SimpleFloorPlanner: iteration i
puts raw_multiply (a, a, a^2) on row 2i and raw_add (a, a^2, a^2 + a) on row 2i + 1, then constrains
a0 = a1 and b1 = c0.  Pure data: columns come back as Montgomery limb arrays for keygen / create_proof."""
from __future__ import annotations

import numpy as np

from .api import R_MOD, fr_encode
from .plonk import ADVICE, ConstraintSystem


def standard_plonk_cs() -> ConstraintSystem:
    """MyCircuit::configure (benches/plonk.rs:203-241)."""
    cs = ConstraintSystem()
    cs.set_minimum_degree(5)
    a, b, c = cs.advice_column(), cs.advice_column(), cs.advice_column()
    for col in (a, b, c):
        cs.enable_equality(col)
    sm, sa, sb, sc = (cs.fixed_column() for _ in range(4))
    qa, qb, qc = cs.query_advice(a), cs.query_advice(b), cs.query_advice(c)
    qsa, qsb, qsc, qsm = cs.query_fixed(sa), cs.query_fixed(sb), cs.query_fixed(sc), cs.query_fixed(sm)
    cs.create_gate("Combined add-mult", [qa * qsa + qb * qsb + qa * qb * qsm - (qc * qsc)])
    return cs


def my_circuit(k: int, a: int):
    """MyCircuit::synthesize (benches/plonk.rs:243-270) -> (fixed [sm, sa, sb, sc], advice [a, b, c], copies)."""
    iters = (1 << (k - 1)) - 3
    a %= R_MOD
    a2, fin = a * a % R_MOD, (a * a + a) % R_MOD
    tile = lambda even, odd: np.tile(fr_encode([even, odd]), (iters, 1))  # noqa: E731
    fixed = [tile(1, 0), tile(0, 1), tile(0, 1), tile(1, 1)]
    advice = [tile(a, a), tile(a, a2), tile(a2, fin)]
    A, B, C_ = (ADVICE, 0), (ADVICE, 1), (ADVICE, 2)
    copies = []
    for i in range(iters):
        copies.append((A, 2 * i, A, 2 * i + 1))
        copies.append((B, 2 * i + 1, C_, 2 * i))
    return fixed, advice, copies
