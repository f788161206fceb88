"""Shared bodies of the quotient-evaluation parity tests: the same function runs against the CPU kernel
emulator (tests/test_emulator_plonk.py) and on the GPU (tests/test_gpu_plonk.py)."""
from __future__ import annotations

import random
from types import SimpleNamespace

import halo2_pse_b200 as h
from oracle import bn256 as O
from oracle import plonk as OP
from tests import helpers as H


def build_cs(variant: str) -> h.ConstraintSystem:
    cs = h.ConstraintSystem()
    if variant == "bench":  # benches/plonk.rs:203-241
        cs.set_minimum_degree(5)
        a, b, c = cs.advice_column(), cs.advice_column(), cs.advice_column()
        for col in (a, b, c):
            cs.enable_equality(col)
        sm, sa, sb, sc = (cs.fixed_column() for _ in range(4))
        qa, qb, qc = cs.query_advice(a), cs.query_advice(b), cs.query_advice(c)
        qsa, qsb, qsc, qsm = cs.query_fixed(sa), cs.query_fixed(sb), cs.query_fixed(sc), cs.query_fixed(sm)
        cs.create_gate("Combined add-mult", [qa * qsa + qb * qsb + qa * qb * qsm - (qc * qsc)])
        return cs
    # "rich": rotations, constants, scaling, negation, challenges, instance, repeated sub-expressions,
    # several gates, two lookups, a permutation over advice + fixed + instance columns in several sets
    a, b, c, d = (cs.advice_column() for _ in range(4))
    f0, f1, f2 = (cs.fixed_column() for _ in range(3))
    inst = cs.instance_column()
    ch = cs.challenge_usable_after(0)
    for col in (a, b, f1, inst, d):
        cs.enable_equality(col)
    qa, qb, qc, qd = cs.query_advice(a), cs.query_advice(b), cs.query_advice(c), cs.query_advice(d)
    qa_next, qb_prev, qc_far = cs.query_advice(a, 1), cs.query_advice(b, -1), cs.query_advice(c, 3)
    q0, q1, q2 = cs.query_fixed(f0), cs.query_fixed(f1), cs.query_fixed(f2, -2)
    qi = cs.query_instance(inst)
    one, two = h.Expression.constant(1), h.Expression.constant(2)
    cs.create_gate("mul", [q0 * (qa * qb - qc), q0 * (qa_next - qa - one)])
    cs.create_gate("misc", [
        q1 * (qa * qa * two - qd * 7 + (-qb_prev)) + h.Expression.constant(0) * qa,
        (qa + qb) * (qa + qb) * q2 - qc_far * ch + qi * (O.R_MOD - 5),
        -(h.Expression.constant(3)) + two * qd - (qa * qb - qc) * one,
        (qa - h.Expression.constant(0)) * (qb * 1) * (qc * 0 + q1),
    ])
    if variant == "rich":
        cs.lookup("l0", [(qa * q0, q1), (qb + qd, q2 * two)])
        cs.lookup("l1", [(qc_far * ch, q0 + one)])
    return cs


def random_case(cs: h.ConstraintSystem, k: int, seed: int, n_circuits: int = 1):
    """Random coefficient-form polynomials for everything evaluate_h reads (it is a function of arbitrary
    inputs: no satisfying witness is needed to compare values)."""
    rng = random.Random(seed)
    n = 1 << k
    rp = lambda: H.rand_fr(rng, n)  # noqa: E731
    case = SimpleNamespace(k=k, n=n)
    case.fixed_polys = [rp() for _ in range(cs.num_fixed_columns)]
    case.l0, case.l_last, case.l_active = rp(), rp(), rp()
    case.sigma_polys = [rp() for _ in cs.permutation.columns]
    chunk = cs.degree() - 2
    n_sets = (len(cs.permutation.columns) + chunk - 1) // chunk
    case.circuits = []
    for _ in range(n_circuits):
        case.circuits.append(SimpleNamespace(
            advice=[rp() for _ in range(cs.num_advice_columns)],
            instance=[rp() for _ in range(cs.num_instance_columns)],
            z=[rp() for _ in range(n_sets)],
            lookups=[SimpleNamespace(product=rp(), permuted_input=rp(), permuted_table=rp()) for _ in cs.lookups]))
    case.challenges = H.rand_fr(rng, cs.num_challenges)
    case.y, case.beta, case.gamma, case.theta = H.rand_fr(rng, 4)
    return case


def oracle_h(cs: h.ConstraintSystem, case) -> list:
    dom = O.EvaluationDomain(cs.degree(), case.k)
    ext = dom.coeff_to_extended
    circuits = []
    for c in case.circuits:
        circuits.append(dict(advice=[ext(p) for p in c.advice], instance=[ext(p) for p in c.instance],
                             perm_sets=[ext(p) for p in c.z],
                             lookups=[dict(product=ext(l.product), permuted_input=ext(l.permuted_input),
                                           permuted_table=ext(l.permuted_table)) for l in c.lookups]))
    return OP.evaluate_h(
        k=case.k, extended_k=dom.extended_k, extended_omega=dom.extended_omega,
        gates=[[p.to_tuple() for p in polys] for _, polys in cs.gates],
        lookups=[([e.to_tuple() for e in l.input_expressions], [e.to_tuple() for e in l.table_expressions])
                 for l in cs.lookups],
        perm_columns=[tuple(c) for c in cs.permutation.columns], chunk_len=cs.degree() - 2,
        blinding_factors=cs.blinding_factors(), fixed=[ext(p) for p in case.fixed_polys], l0=ext(case.l0),
        l_last=ext(case.l_last), l_active_row=ext(case.l_active), sigma_cosets=[ext(p) for p in case.sigma_polys],
        circuits=circuits, challenges=case.challenges, y=case.y, beta=case.beta, gamma=case.gamma, theta=case.theta)


def device_h(ctx: h.Context, cs: h.ConstraintSystem, case) -> list:
    """Evaluator.evaluate_h through the C ABI, everything device-resident."""
    dom = h.EvaluationDomain(ctx, cs.degree(), case.k)
    up = lambda vals: ctx.upload_fr(h.fr_encode(vals))  # noqa: E731

    def coset(vals):
        src = up(vals)
        out = ctx.alloc(dom.extended_len() * 32)
        dom.coeff_to_extended_device(src, out)
        src.free()
        return out

    pk = SimpleNamespace(domain=dom, cs=cs, fixed_cosets=[coset(p) for p in case.fixed_polys], l0=coset(case.l0),
                         l_last=coset(case.l_last), l_active_row=coset(case.l_active),
                         permutation_cosets=[coset(p) for p in case.sigma_polys])
    ev = h.Evaluator(cs)
    perms = [SimpleNamespace(sets=[SimpleNamespace(permutation_product_coset=coset(z)) for z in c.z])
             for c in case.circuits]
    lookups = [[SimpleNamespace(product_poly=up(l.product), permuted_input_poly=up(l.permuted_input),
                                permuted_table_poly=up(l.permuted_table)) for l in c.lookups]
               for c in case.circuits]
    values = ev.evaluate_h(pk, [[up(p) for p in c.advice] for c in case.circuits],
                           [[up(p) for p in c.instance] for c in case.circuits], case.challenges, case.y, case.beta,
                           case.gamma, case.theta, lookups, perms)
    out = h.fr_decode(values.download(dom.extended_len()))
    ev.free()
    dom.free()
    return out


def check_evaluate_h(ctx: h.Context, variant: str, k: int, seed: int, n_circuits: int = 1):
    cs = build_cs(variant)
    case = random_case(cs, k, seed, n_circuits)
    want = oracle_h(cs, case)
    got = device_h(ctx, cs, case)
    assert len(got) == len(want)
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, (variant, k, len(bad), bad[:4])


# ---------------------------------------------------------------------------
# prover cases
# ---------------------------------------------------------------------------
def oracle_cs(cs: h.ConstraintSystem, OV=None):
    """The product's ConstraintSystem as the oracle's plain-data CS (data copying only).  OV: the oracle prover
    module to build it for (default: the bn256 instance; oracle/pasta.py loads the same source over Vesta)."""
    if OV is None:
        from oracle import prover as OV
    return OV.CS(num_fixed_columns=cs.num_fixed_columns, num_advice_columns=cs.num_advice_columns,
                 num_instance_columns=cs.num_instance_columns,
                 gates=[[p.to_tuple() for p in polys] for _, polys in cs.gates],
                 advice_queries=[(c.index, r) for c, r in cs.advice_queries],
                 instance_queries=[(c.index, r) for c, r in cs.instance_queries],
                 fixed_queries=[(c.index, r) for c, r in cs.fixed_queries],
                 perm_columns=[tuple(c) for c in cs.permutation.columns],
                 num_advice_queries=cs.num_advice_queries, minimum_degree=cs.minimum_degree,
                 num_challenges=cs.num_challenges, advice_column_phase=cs.advice_column_phase,
                 challenge_phase=cs.challenge_phase,
                 lookups=[([e.to_tuple() for e in l.input_expressions], [e.to_tuple() for e in l.table_expressions])
                          for l in cs.lookups])


def bench_circuit(k: int, a: int):
    """MyCircuit of benches/plonk.rs:246-270 laid out by SimpleFloorPlanner: iteration i puts raw_multiply on
    row 2i and raw_add on row 2i+1, then copies a0 = a1 and b1 = c0.
    -> (fixed columns [sm, sa, sb, sc], advice columns [a, b, c], copy constraints)"""
    iters = (1 << (k - 1)) - 3
    r = O.R_MOD
    a2 = a * a % r
    fin = (a2 + a) % r
    sm, sa, sb, sc, ca, cb, cc, copies = [], [], [], [], [], [], [], []
    A, B, C_ = (h.ADVICE, 0), (h.ADVICE, 1), (h.ADVICE, 2)
    for i in range(iters):
        # raw_multiply: (a, a, a^2), sa = sb = 0, sc = sm = 1
        ca.append(a), cb.append(a), cc.append(a2)
        sa.append(0), sb.append(0), sc.append(1), sm.append(1)
        # raw_add: (a, a^2, a^2 + a), sa = sb = sc = 1, sm = 0
        ca.append(a), cb.append(a2), cc.append(fin)
        sa.append(1), sb.append(1), sc.append(1), sm.append(0)
        copies.append((A, 2 * i, A, 2 * i + 1))
        copies.append((B, 2 * i + 1, C_, 2 * i))
    return [sm, sa, sb, sc], [ca, cb, cc], copies


S_TOXIC = 0x1234567890ABCDEF1234567890ABCDEF


def oracle_bench_proof(k: int, a: int, seed: bytes):
    """Proof bytes of the bench circuit from the big-integer oracle (+ what is needed to verify them)."""
    from oracle import prover as OV
    params = O.ParamsKZG.setup(k, S_TOXIC)
    cs = oracle_cs(build_cs("bench"))
    fixed, advice, copies = bench_circuit(k, a)
    pk = OV.keygen(params, cs, fixed, copies)
    t = OV.Blake2bWrite()
    OV.create_proof(params, pk, [lambda phase, ch: dict(enumerate(advice))], [[]], OV.XorShiftRng(seed), t)
    return params, pk, t.finalize()


def device_bench_proof(ctx: h.Context, k: int, a: int, seed: bytes, precompute: bool = False, timings=None,
                       params_hook=None):
    """The same proof through keygen / create_proof of the product on `ctx`; params_hook (e.g.
    dist.shard_params) may replace the ParamsKZG before keygen."""
    params = h.ParamsKZG.setup(ctx, k, S_TOXIC, precompute=precompute)
    if params_hook is not None:
        params = params_hook(params)
    cs = build_cs("bench")
    fixed, advice, copies = bench_circuit(k, a)
    pk = h.keygen(params, cs, fixed, copies)
    t = h.Blake2bWrite()
    h.create_proof(params, pk, [lambda phase, ch: dict(enumerate(advice))], [[]], h.XorShiftRng(seed), t,
                   timings=timings)
    return params, pk, t.finalize()


def check_bench_proof_bytes(ctx: h.Context, k: int, seed: bytes = b"\x07" * 16):
    from oracle import prover as OV
    a = 0xDEADBEEF
    oparams, opk, want = oracle_bench_proof(k, a, seed)
    _, pk, got = device_bench_proof(ctx, k, a, seed)
    assert pk.pinned == opk.debug            # identical verifying key (commitments included)
    assert pk.transcript_repr == opk.transcript_repr
    assert got == want, [i for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]][:4]
    assert OV.verify_proof(oparams, S_TOXIC, opk, [[]], got)
    pk.free()


def bench_circuit_limbs(k: int, a: int):
    """bench_circuit() as numpy limb arrays (the columns have period 2: tiled, no per-row Python)."""
    import numpy as np
    iters = (1 << (k - 1)) - 3
    r = O.R_MOD
    a2, fin = a * a % r, (a * a + a) % r
    tile = lambda even, odd: np.tile(h.fr_encode([even, odd]), (iters, 1))  # noqa: E731
    fixed = [tile(1, 0), tile(0, 1), tile(0, 1), tile(1, 1)]  # sm, sa, sb, sc
    advice = [tile(a, a), tile(a, a2), tile(a2, fin)]
    A, B, C_ = (h.ADVICE, 0), (h.ADVICE, 1), (h.ADVICE, 2)
    copies = []
    for i in range(iters):
        copies.append((A, 2 * i, A, 2 * i + 1))
        copies.append((B, 2 * i + 1, C_, 2 * i))
    return fixed, advice, copies


def oracle_vk_of(pk):
    """The verifying-key half of a device ProvingKey as the object oracle.prover.verify_proof reads
    (constants and commitments only: cheap at any k)."""
    from types import SimpleNamespace
    cs = oracle_cs(pk.cs)
    return SimpleNamespace(cs=cs, domain=O.EvaluationDomain(cs.degree(), pk.k), n=pk.n, k=pk.k,
                           fixed_commitments=pk.fixed_commitments, perm_commitments=pk.perm_commitments,
                           transcript_repr=pk.transcript_repr)


def build_lookup_cs() -> h.ConstraintSystem:
    """A small circuit with a gate, copy constraints and one lookup argument (degree 5):
    q_mul * (a * a - b) = 0,  (q_lk * a) in t."""
    cs = h.ConstraintSystem()
    a, b = cs.advice_column(), cs.advice_column()
    q_mul, q_lk, t = cs.fixed_column(), cs.fixed_column(), cs.fixed_column()
    cs.enable_equality(a)
    cs.enable_equality(b)
    qa, qb = cs.query_advice(a), cs.query_advice(b)
    qm, ql, qt = cs.query_fixed(q_mul), cs.query_fixed(q_lk), cs.query_fixed(t)
    cs.create_gate("square", [qm * (qa * qa - qb)])
    cs.lookup("range", [(ql * qa, qt)])
    return cs


def lookup_circuit(k: int, seed: int = 5):
    """-> (fixed [q_mul, q_lk, t], advice [a, b], copies) satisfying build_lookup_cs() on 2^k rows."""
    cs = build_lookup_cs()
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    rng = random.Random(seed)
    table = list(range(usable))
    a = [rng.randrange(min(usable, 11)) for _ in range(usable)]  # many repeats, all inside the table
    a[3] = a[2]
    q_lk = [1 if i < usable - 4 else 0 for i in range(usable)]
    q_mul = [1 if i % 3 else 0 for i in range(usable)]
    b = [x * x % O.R_MOD if q else rng.randrange(O.R_MOD) for x, q in zip(a, q_mul)]
    copies = [((h.ADVICE, 0), 3, (h.ADVICE, 0), 2)]
    if q_mul[4] and q_mul[7]:
        a[7] = a[4]
        b[7] = b[4]
        copies.append(((h.ADVICE, 1), 7, (h.ADVICE, 1), 4))
    return [q_mul, q_lk, table], [a, b], copies


def check_lookup_proof_bytes(ctx: h.Context, k: int = 5, seed: bytes = b"\x21" * 16):
    """keygen + create_proof of the lookup circuit: same vk, same proof bytes as the big-integer oracle."""
    from oracle import prover as OV
    fixed, advice, copies = lookup_circuit(k)
    oparams = O.ParamsKZG.setup(k, S_TOXIC)
    opk = OV.keygen(oparams, oracle_cs(build_lookup_cs()), fixed, copies)
    t = OV.Blake2bWrite()
    OV.create_proof(oparams, opk, [lambda phase, ch: dict(enumerate(advice))], [[]], OV.XorShiftRng(seed), t)
    want = t.finalize()
    params = h.ParamsKZG.setup(ctx, k, S_TOXIC)
    pk = h.keygen(params, build_lookup_cs(), fixed, copies)
    t = h.Blake2bWrite()
    h.create_proof(params, pk, [lambda phase, ch: dict(enumerate(advice))], [[]], h.XorShiftRng(seed), t)
    got = t.finalize()
    assert pk.pinned == opk.debug
    assert got == want, [i // 32 for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]][:6]
    assert OV.verify_proof(oparams, S_TOXIC, opk, [[]], got)
    # an input value outside the table: Error::ConstraintSystemFailure
    bad = [list(c) for c in advice]
    bad[0][0] = 1 << 40
    try:
        h.create_proof(params, pk, [lambda phase, ch: dict(enumerate(bad))], [[]], h.XorShiftRng(seed), h.Blake2bWrite())
        raise AssertionError("expected H2B_ERR_CONSTRAINT")
    except h.H2BError as e:
        assert e.code == h.H2B_ERR_CONSTRAINT
    pk.free()


def check_lookup_permute(ctx: h.Context, n: int, seed: int, distinct: int):
    """h2b_lookup_permute against the oracle's permute_expression_pair on random columns with many repeats."""
    from oracle import prover as OV
    rng = random.Random(seed)
    table = [rng.randrange(O.R_MOD) for _ in range(distinct)]
    table = (table + [rng.choice(table) for _ in range(n - distinct)])[:n]
    rng.shuffle(table)
    inp = [rng.choice(table) for _ in range(n)]

    class Zero:
        def next_u64(self):
            return 0
    want_i, want_t = OV.permute_expression_pair(inp, table, n, -1, Zero())
    di, dt = ctx.upload_fr(h.fr_encode(inp)), ctx.upload_fr(h.fr_encode(table))
    oi, ot = ctx.alloc(n * 32), ctx.alloc(n * 32)
    ctx._check(ctx.lib.h2b_lookup_permute(ctx.h, di.ptr, dt.ptr, n, oi.ptr, ot.ptr))
    assert h.fr_decode(oi.download(n)) == want_i
    assert h.fr_decode(ot.download(n)) == want_t


def check_shplonk_proof_bytes(ctx: h.Context, which: str, k: int = 5, seed: bytes = b"\x33" * 16):
    """create_proof with ProverSHPLONK: same proof bytes as the big-integer oracle, accepted by its verifier."""
    from oracle import prover as OV
    if which == "bench":
        cs_fn, (fixed, advice, copies) = (lambda: build_cs("bench")), bench_circuit(k, 0xFACE)
    else:
        cs_fn, (fixed, advice, copies) = build_lookup_cs, lookup_circuit(k)
    witness = lambda phase, ch: dict(enumerate(advice))  # noqa: E731
    oparams = O.ParamsKZG.setup(k, S_TOXIC)
    opk = OV.keygen(oparams, oracle_cs(cs_fn()), fixed, copies)
    t = OV.Blake2bWrite()
    OV.create_proof(oparams, opk, [witness], [[]], OV.XorShiftRng(seed), t, multiopen="shplonk")
    want = t.finalize()
    params = h.ParamsKZG.setup(ctx, k, S_TOXIC)
    pk = h.keygen(params, cs_fn(), fixed, copies)
    t = h.Blake2bWrite()
    h.create_proof(params, pk, [witness], [[]], h.XorShiftRng(seed), t, prover=h.ProverSHPLONK)
    got = t.finalize()
    assert got == want, [i // 32 for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]][:6]
    assert OV.verify_proof(oparams, S_TOXIC, opk, [[]], got, multiopen="shplonk")
    pk.free()


# ---------------------------------------------------------------------------
# the C++ restatement of evaluate_h (oracle/ref_cpu.cpp): checker at sizes the big-integer oracle cannot reach
# ---------------------------------------------------------------------------
def oracle_c_h(oc, cs: h.ConstraintSystem, case, threads: int = 0):
    """h over the extended domain from oracle_evaluate_h, with cosets from the C++ domain transforms."""
    import numpy as np
    dom = oc.domain(cs.degree(), case.k, threads)
    ext = lambda vals: dom.coeff_to_extended(vals if isinstance(vals, np.ndarray) else H.fr_enc(vals))  # noqa: E731
    ev = h.Evaluator(cs)  # the calculation lists (plain data); interpreted by the C++ restatement

    def graph(g):
        return oc.graph(g.encode(), H.fr_enc(g.constants), g.rotations, g.num_intermediates)

    gates, lgs = graph(ev.custom_gates), [graph(g) for g in ev.lookups]
    fixed = [ext(p) for p in case.fixed_polys]
    l0, l_last, l_act = ext(case.l0), ext(case.l_last), ext(case.l_active)
    sigma = [ext(p) for p in case.sigma_polys]
    values = np.zeros((1 << dom.extended_k, 4), dtype=np.uint64)
    for c in case.circuits:
        lcos = []
        for l in c.lookups:
            lcos += [ext(l.product), ext(l.permuted_input), ext(l.permuted_table)]
        oc.evaluate_h(dom, gates, fixed, [ext(p) for p in c.advice], [ext(p) for p in c.instance],
                      H.fr_enc(case.challenges), case.beta, case.gamma, case.theta, case.y,
                      [tuple(col) for col in cs.permutation.columns], sigma, [ext(p) for p in c.z], cs.degree() - 2,
                      cs.blinding_factors(), l0, l_last, l_act, lgs, lcos, values, threads)
    for g in [gates] + lgs:
        oc.lib.oracle_graph_free(g)
    dom.free()
    return values


def random_case_limbs(cs: h.ConstraintSystem, k: int, seed: int, n_circuits: int = 1):
    """random_case() with numpy limb arrays instead of Python integers (large k)."""
    n = 1 << k
    ctr = [seed * 1000]

    def rp():
        ctr[0] += 1
        return H.rand_fr_limbs(ctr[0], n)
    case = SimpleNamespace(k=k, n=n)
    case.fixed_polys = [rp() for _ in range(cs.num_fixed_columns)]
    case.l0, case.l_last, case.l_active = rp(), rp(), rp()
    case.sigma_polys = [rp() for _ in cs.permutation.columns]
    chunk = cs.degree() - 2
    n_sets = (len(cs.permutation.columns) + chunk - 1) // chunk
    case.circuits = [SimpleNamespace(advice=[rp() for _ in range(cs.num_advice_columns)],
                                     instance=[rp() for _ in range(cs.num_instance_columns)],
                                     z=[rp() for _ in range(n_sets)],
                                     lookups=[SimpleNamespace(product=rp(), permuted_input=rp(), permuted_table=rp())
                                              for _ in cs.lookups]) for _ in range(n_circuits)]
    rng = random.Random(seed)
    case.challenges = H.rand_fr(rng, cs.num_challenges)
    case.y, case.beta, case.gamma, case.theta = H.rand_fr(rng, 4)
    return case


def device_h_limbs(ctx: h.Context, cs: h.ConstraintSystem, case):
    """device_h() for limb-array cases; returns (2^ek, 4) limbs."""
    dom = h.EvaluationDomain(ctx, cs.degree(), case.k)
    up = ctx.upload_fr

    def coset(limbs):
        src = up(limbs)
        out = ctx.alloc(dom.extended_len() * 32)
        dom.coeff_to_extended_device(src, out)
        src.free()
        return out
    pk = SimpleNamespace(domain=dom, cs=cs, fixed_cosets=[coset(p) for p in case.fixed_polys], l0=coset(case.l0),
                         l_last=coset(case.l_last), l_active_row=coset(case.l_active),
                         permutation_cosets=[coset(p) for p in case.sigma_polys])
    ev = h.Evaluator(cs)
    perms = [SimpleNamespace(sets=[SimpleNamespace(permutation_product_coset=coset(z)) for z in c.z])
             for c in case.circuits]
    lookups = [[SimpleNamespace(product_poly=up(l.product), permuted_input_poly=up(l.permuted_input),
                                permuted_table_poly=up(l.permuted_table)) for l in c.lookups] for c in case.circuits]
    values = ev.evaluate_h(pk, [[up(p) for p in c.advice] for c in case.circuits],
                           [[up(p) for p in c.instance] for c in case.circuits], case.challenges, case.y, case.beta,
                           case.gamma, case.theta, lookups, perms)
    out = values.download(dom.extended_len())
    ev.free()
    dom.free()
    return out


def build_phases_cs() -> h.ConstraintSystem:
    """Two phases: a (phase 0), a challenge usable after phase 0, b (phase 1); gate q * (b - a * challenge)."""
    cs = h.ConstraintSystem()
    a, b = cs.advice_column(0), cs.advice_column(1)
    q = cs.fixed_column()
    ch = cs.challenge_usable_after(0)
    cs.enable_equality(a)
    cs.enable_equality(b)
    qa, qb, qq = cs.query_advice(a), cs.query_advice(b), cs.query_fixed(q)
    cs.create_gate("prod", [qq * (qb - qa * ch)])
    return cs


def phases_circuit(k: int):
    """-> (fixed [q], witness(phase, challenges) -> columns, copies) satisfying build_phases_cs() on 2^k rows."""
    cs = build_phases_cs()
    usable = (1 << k) - (cs.blinding_factors() + 1)
    a = [i + 2 for i in range(usable)]
    a[2] = a[1]

    def witness(phase, challenges):
        if phase == 0:
            return {0: a}
        return {1: [x * challenges[0] % O.R_MOD for x in a]}

    copies = [((h.ADVICE, 0), 1, (h.ADVICE, 0), 2), ((h.ADVICE, 1), 1, (h.ADVICE, 1), 2)]
    return [[1] * usable], witness, copies


def check_phases_proof_bytes(ctx: h.Context, k: int = 5, seed: bytes = b"\x42" * 16):
    """A circuit with two advice phases and a challenge: same vk and proof bytes from the product's prover as from
    the big-integer oracle, and the restated reference verifier accepts them.  -> (proof bytes, pinned vk)"""
    from oracle import prover as OV
    fixed, witness, copies = phases_circuit(k)
    oparams = O.ParamsKZG.setup(k, S_TOXIC)
    opk = OV.keygen(oparams, oracle_cs(build_phases_cs()), fixed, copies)
    t = OV.Blake2bWrite()
    OV.create_proof(oparams, opk, [witness], [[]], OV.XorShiftRng(seed), t)
    want = t.finalize()
    assert OV.verify_proof(oparams, S_TOXIC, opk, [[]], want)
    params = h.ParamsKZG.setup(ctx, k, S_TOXIC)
    pk = h.keygen(params, build_phases_cs(), fixed, copies)
    t = h.Blake2bWrite()
    h.create_proof(params, pk, [witness], [[]], h.XorShiftRng(seed), t)
    got = t.finalize()
    assert pk.pinned == opk.debug
    assert got == want, [i // 32 for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]][:6]
    pk.free()
    params.g.free()
    params.g_lagrange.free()
    return want, opk.debug


def plonk_api_configure():
    """MyCircuit::configure of the reference's tests/plonk_api.rs:389-470, statement by statement, on the host
    mirror's ConstraintSystem (the circuit behind the reference's pinned verifying key)."""
    cs = h.ConstraintSystem()
    e, a, b = cs.advice_column(), cs.advice_column(), cs.advice_column()
    sf = cs.fixed_column()
    c, d = cs.advice_column(), cs.advice_column()
    p = cs.instance_column()
    for col in (a, b, c):
        cs.enable_equality(col)
    sm, sa, sb, sc, sp = (cs.fixed_column() for _ in range(5))
    sl = cs.fixed_column()  # lookup_table_column
    a_ = cs.query_advice(a)
    cs.lookup("lookup", [(a_, cs.query_fixed(sl))])  # the table column is queried after the closure ran
    qd, qa, qsf = cs.query_advice(d, 1), cs.query_advice(a), cs.query_fixed(sf)
    qe, qb, qc = cs.query_advice(e, -1), cs.query_advice(b), cs.query_advice(c)
    qsa, qsb, qsc, qsm = cs.query_fixed(sa), cs.query_fixed(sb), cs.query_fixed(sc), cs.query_fixed(sm)
    cs.create_gate("Combined add-mult", [qa * qsa + qb * qsb + qa * qb * qsm - (qc * qsc) + qsf * (qd * qe)])
    qa, qp, qsp = cs.query_advice(a), cs.query_instance(p), cs.query_fixed(sp)
    cs.create_gate("Public input", [qsp * (qa - qp)])
    for col in (sf, e, d, p, sm, sa, sb, sc, sp):
        cs.enable_equality(col)
    return cs


def plonk_api_keygen_inputs(k: int, a_value: int, instance: int, blinding_factors: int):
    """What MyCircuit::synthesize (tests/plonk_api.rs:472-500) assigns at keygen, laid out the way
    SimpleFloorPlanner places it (circuit/floor_planner/single_pass.rs): region `public_input` on row 0 (columns a,
    sp), then ten times raw_multiply and raw_add on the next free rows of their columns (1, 2, ..., 20) and two
    column-less `copy` regions; the lookup table in rows 0..3 of `sl`, the rest of the usable rows filled with its
    first value (single_pass.rs:192-198, keygen.rs:154-175).
    -> (fixed columns in index order [sf, sm, sa, sb, sc, sp, sl], copy constraints)"""
    n = 1 << k
    sf, sm, sa, sb, sc, sp, sl = ([0] * n for _ in range(7))
    A, B, C_ = (h.ADVICE, 1), (h.ADVICE, 2), (h.ADVICE, 3)  # advice columns are created in the order e, a, b, c, d
    copies = []
    sp[0] = 1
    row = 1
    for _ in range(10):
        mul, add = row, row + 1
        sa[mul], sb[mul], sc[mul], sm[mul] = 0, 0, 1, 1  # raw_multiply, plonk_api.rs:141-148
        sa[add], sb[add], sc[add], sm[add] = 1, 1, 1, 0  # raw_add, :199-206
        copies += [(A, mul, A, add)] * 2                 # copy(a0, a1): constrain_equal twice, :223-226
        copies += [(B, add, C_, mul)] * 2                # copy(b1, c0)
        row += 2
    table = [instance, a_value, a_value, 0]              # common!(): lookup_table, :511-516
    usable = n - (blinding_factors + 1)
    for i in range(usable):
        sl[i] = table[i] if i < len(table) else table[0]
    return [sf, sm, sa, sb, sc, sp, sl], copies
