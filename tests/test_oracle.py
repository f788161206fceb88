"""Pins the oracles (oracle/bn256.py, oracle/ref_cpu.cpp) -- CPU only.

The reference holds no golden bytes for bn256, so the oracle is
pinned by (i) the definitional known answers of tests/golden/kat_bn256.json,
(ii) the reference's own test identities (kzg/commitment.rs:361-384,
domain.rs:488-557 restated for Fr), (iii) agreement with the mathematical
definition (naive DFT, naive sum), and (iv) agreement between the two
independently written restatements (Python big-int vs C++ limbs).
"""
import random

import numpy as np
import pytest

from oracle import bn256 as O
from tests import helpers as H

KAT = H.load_golden("kat_bn256.json")
DEF = KAT["definitional"]
VEC = KAT["vectors"]


def ih(s):
    return int(s, 16)


def test_constants_match_definition():
    r, q = ih(DEF["fr_modulus"]), ih(DEF["fq_modulus"])
    assert (O.R_MOD, O.Q_MOD) == (r, q)
    u = DEF["bn_u"]
    assert q == 36 * u**4 + 36 * u**3 + 24 * u**2 + 6 * u + 1
    assert r == 36 * u**4 + 36 * u**3 + 18 * u**2 + 6 * u + 1
    assert (r - 1) % (1 << 28) == 0 and (r - 1) % (1 << 29) != 0 and O.S == DEF["fr_S"] == 28
    assert O.ROOT_OF_UNITY == ih(DEF["root_of_unity"]) == pow(7, (r - 1) >> 28, r)
    assert O.ROOT_OF_UNITY_INV == ih(DEF["root_of_unity_inv"])
    assert O.ZETA == ih(DEF["zeta"]) and pow(O.ZETA, 3, r) == 1 and O.ZETA != 1
    assert O.ZETA * O.ZETA % r == ih(DEF["zeta_sq"])
    assert pow(7, 1 << 28, r) == ih(DEF["delta"])
    assert O.TWO_INV == ih(DEF["two_inv"])
    assert O.MONT_R_FR == ih(DEF["fr_R"]) and O.MONT_R_FR**2 % r == ih(DEF["fr_R2"])
    assert O.MONT_R_FQ == ih(DEF["fq_R"]) and O.MONT_R_FQ**2 % q == ih(DEF["fq_R2"])
    assert (-pow(r, -1, 1 << 64)) % (1 << 64) == ih(DEF["fr_inv64"])
    assert (-pow(q, -1, 1 << 64)) % (1 << 64) == ih(DEF["fq_inv64"])
    for k, w in DEF["omega"].items():
        assert O.omega_for(int(k)) == ih(w)
        assert pow(ih(w), 1 << (int(k) - 1), r) == r - 1  # exact order 2^k


def test_definitional_known_answers():
    kat = DEF["ntt_k2"]
    a = list(kat["in"])
    O.best_fft(a, ih(kat["omega"]), 2)
    assert a == [ih(x) for x in kat["out"]]
    G = tuple(DEF["g1_generator"])
    assert O.g1_is_on_curve(G)
    for m, (x, y) in DEF["g1_multiples"].items():
        assert O.g1_mul(G, int(m)) == (ih(x), ih(y))
    for case in DEF["msm"]:
        sc = [O.R_MOD - 1 if s == "r-1" else s for s in case["scalars"]]
        bases = [O.g1_mul(G, m) for m in case["bases_multiples_of_G"]]
        want = O.g1_mul(G, case["result_multiple_of_G"]) if case["result_multiple_of_G"] else None
        assert O.best_multiexp(sc, bases) == want
        assert O.msm_naive(sc, bases) == want


def test_group_law_exceptional_cases():
    G = O.G1_GEN
    P = O.g1_mul(G, 12345)
    assert O.g1_add(P, None) == P and O.g1_add(None, P) == P
    assert O.g1_add(P, O.g1_neg(P)) is None
    assert O.g1_add(P, P) == O.g1_double(P) == O.g1_mul(G, 24690)
    assert O.g1_mul(G, O.R_MOD) is None and O.g1_mul(G, O.R_MOD - 1) == O.g1_neg(G)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 7])
def test_best_fft_is_the_dft(k):
    rng = random.Random(k)
    a = H.rand_fr(rng, 1 << k)
    w = O.omega_for(k)
    b = list(a)
    O.best_fft(b, w, k)
    assert b == O.dft_naive(a, w)
    O.best_fft(b, pow(w, -1, O.R_MOD), k)
    ninv = pow(1 << k, -1, O.R_MOD)
    assert [x * ninv % O.R_MOD for x in b] == a


@pytest.mark.parametrize("n,threads", [(1, 1), (3, 1), (5, 2), (31, 1), (32, 1), (70, 3), (130, 8)])
def test_best_multiexp_is_the_sum(n, threads):
    rng = random.Random(n * 31 + threads)
    bases = [O.g1_mul(O.G1_GEN, rng.randrange(1, 1 << 40)) for _ in range(n)]
    sc = H.rand_fr(rng, n)
    assert O.best_multiexp(sc, bases, threads) == O.msm_naive(sc, bases)


@pytest.mark.parametrize("j,k", [(2, 3), (3, 3), (4, 4), (5, 4), (9, 2)])
def test_domain_identities(j, k):
    rng = random.Random(j * 100 + k)
    D = O.EvaluationDomain(j, k)
    r = O.R_MOD
    n = 1 << k
    assert (1 << D.extended_k) >= n * (j - 1) and (D.extended_k == k or (1 << (D.extended_k - 1)) < n * (j - 1))
    assert pow(D.omega, n, r) == 1 and pow(D.omega, n // 2, r) == r - 1
    coeffs = H.rand_fr(rng, n)
    # lagrange_to_coeff inverts evaluation on the domain (domain.rs:488-557 restated)
    evals = [O.eval_polynomial(coeffs, pow(D.omega, i, r)) for i in range(n)]
    assert D.lagrange_to_coeff(evals) == coeffs
    # coeff_to_extended evaluates on the zeta-coset of the extended domain
    ext = D.coeff_to_extended(coeffs)
    for i in (0, 1, 2, len(ext) // 2, len(ext) - 1):
        # the reference "coset" multiplies coefficient i by zeta^(i mod 3): p'(X) with p'_i = p_i zeta^(i%3)
        pc = [c * pow(O.ZETA, idx % 3, r) % r for idx, c in enumerate(coeffs)]
        assert ext[i] == O.eval_polynomial(pc, pow(D.extended_omega, i, r))
    # extended_to_coeff undoes it (truncated to n*(j-1) >= n coefficients)
    back = D.extended_to_coeff(ext)
    assert len(back) == n * (j - 1)
    assert back[:n] == coeffs and all(x == 0 for x in back[n:])
    # t_evaluations: 1 / ((zeta w_ext^i)^n - 1)
    for i, t in enumerate(D.t_evaluations):
        x = pow(O.ZETA * pow(D.extended_omega, i, r) % r, n, r)
        assert t * (x - 1) % r == 1


def test_kzg_commit_identity():
    """kzg/commitment.rs:361-384: commit(lagrange_to_coeff(a)) == commit_lagrange(a)."""
    rng = random.Random(99)
    k = 3
    P = O.ParamsKZG.setup(k, rng.randrange(O.R_MOD))
    D = O.EvaluationDomain(1, k) if False else O.EvaluationDomain(2, k)
    a = H.rand_fr(rng, 1 << k)
    assert P.commit(D.lagrange_to_coeff(a)) == P.commit_lagrange(a)


def test_golden_vectors_python_oracle():
    for v in VEC["best_fft"]:
        a = O.frs_from_bytes(bytes.fromhex(v["in"]))
        O.best_fft(a, O.fr_from_mont_bytes(bytes.fromhex(v["omega"])), v["log_n"])
        assert O.frs_to_bytes(a).hex() == v["out"]
    for v in VEC["best_multiexp"]:
        sc = O.frs_from_bytes(bytes.fromhex(v["scalars"]))
        raw = bytes.fromhex(v["bases"])
        bases = [O.g1_from_bytes(raw[i:i + 64]) for i in range(0, len(raw), 64)]
        assert O.g1_to_bytes(O.best_multiexp(sc, bases, 2)).hex() == v["result"], v["name"]


# ---------------------------------------------------------------------------
# C++ restatement (the timed CPU baseline) against the golden vectors and Python
# ---------------------------------------------------------------------------
def _np(hexs, width):
    return np.frombuffer(bytes.fromhex(hexs), dtype=np.uint64).reshape(-1, width)


def test_c_oracle_field_ops(oracle_c):
    rng = random.Random(5)
    for field, mod in ((0, O.R_MOD), (1, O.Q_MOD)):
        a = [rng.randrange(mod) for _ in range(300)] + [0, 1, mod - 1, mod - 1]
        b = [rng.randrange(mod) for _ in range(300)] + [mod - 1, mod - 1, mod - 1, 1]
        A, B = H.to_limbs(a, mod), H.to_limbs(b, mod)
        assert H.from_limbs(oracle_c.field_op(field, 0, A, B), mod) == [x * y % mod for x, y in zip(a, b)]
        assert H.from_limbs(oracle_c.field_op(field, 1, A, B), mod) == [(x + y) % mod for x, y in zip(a, b)]
        assert H.from_limbs(oracle_c.field_op(field, 2, A, B), mod) == [(x - y) % mod for x, y in zip(a, b)]
        nz = [x or 1 for x in a[:20]]
        assert H.from_limbs(oracle_c.field_op(field, 7, H.to_limbs(nz, mod), B[:20]), mod) == \
            [pow(x, -1, mod) for x in nz]


@pytest.mark.parametrize("threads", [1, 2, 8])
def test_c_oracle_golden_fft(oracle_c, threads):
    for v in VEC["best_fft"]:
        out = oracle_c.best_fft(_np(v["in"], 4), _np(v["omega"], 4)[0], v["log_n"], threads)
        assert out.tobytes().hex() == v["out"]


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_c_oracle_golden_msm(oracle_c, threads):
    for v in VEC["best_multiexp"]:
        out = oracle_c.best_multiexp(_np(v["scalars"], 4), _np(v["bases"], 8), threads)
        assert out.tobytes().hex() == v["result"], v["name"]


def test_c_oracle_golden_domain(oracle_c):
    for v in VEC["domain"]:
        d = oracle_c.domain(v["j"], v["k"], 2)
        assert d.extended_k == v["extended_k"]
        assert d.constant(0).tobytes().hex() == v["omega"]
        assert d.constant(2).tobytes().hex() == v["extended_omega"]
        te = b"".join(d.constant(8 + i).tobytes() for i in range(1 << (d.extended_k - d.k)))
        assert te.hex() == v["t_evaluations"]
        assert d.lagrange_to_coeff(_np(v["a"], 4)).tobytes().hex() == v["lagrange_to_coeff"]
        assert d.coeff_to_extended(_np(v["a"], 4)).tobytes().hex() == v["coeff_to_extended"]
        assert d.divide_by_vanishing_poly(_np(v["ext"], 4)).tobytes().hex() == v["divide_by_vanishing_poly"]
        assert d.extended_to_coeff(_np(v["ext"], 4)).tobytes().hex() == v["extended_to_coeff"]
        d.free()


def test_c_oracle_vs_python_medium(oracle_c):
    """Sizes where the recursive/threaded branches of the restatement are taken."""
    rng = random.Random(11)
    k = 12
    a = H.rand_fr(rng, 1 << k)
    w = O.omega_for(k)
    want = list(a)
    O.best_fft(want, w, k)
    for threads in (1, 4, 16):  # log_n <= log_threads never holds here: recursive branch
        assert H.fr_dec(oracle_c.best_fft(H.fr_enc(a), H.fr_enc([w])[0], k, threads)) == want
    # iterative branch: log_n <= log2(threads)
    a3 = a[:8]
    want = list(a3)
    O.best_fft(want, O.omega_for(3), 3)
    assert H.fr_dec(oracle_c.best_fft(H.fr_enc(a3), H.fr_enc([O.omega_for(3)])[0], 3, 8)) == want
    # MSM with known discrete logs: sum c_i [h_i]G = [sum c_i h_i]G
    n = 3000
    hs = [rng.randrange(1, 1 << 64) for _ in range(n)]
    bases = oracle_c.g1_mul_gen(hs)
    sc = H.rand_fr(rng, n)
    want = O.g1_mul(O.G1_GEN, sum(c * h for c, h in zip(sc, hs)) % O.R_MOD)
    for threads in (1, 7):
        assert H.g1_dec(oracle_c.best_multiexp(H.fr_enc(sc), bases, threads))[0] == want


def test_poly_helpers_oracles(oracle_c):
    """eval_polynomial / kate_division / compute_inner_product (arithmetic.rs:304-367): Python vs C++ vs identities."""
    rng = random.Random(21)
    for n in (1, 2, 7, 33, 500):
        a = H.rand_fr(rng, n)
        b = H.rand_fr(rng, n)
        x = rng.randrange(O.R_MOD)
        ev = O.eval_polynomial(a, x)
        assert ev == sum(c * pow(x, i, O.R_MOD) for i, c in enumerate(a)) % O.R_MOD
        for threads in (1, 3, 64):
            assert H.fr_dec(oracle_c.eval_polynomial(H.fr_enc(a), H.fr_enc([x])[0], threads))[0] == ev
        assert H.fr_dec(oracle_c.inner_product(H.fr_enc(a), H.fr_enc(b)))[0] == O.compute_inner_product(a, b)
        q = O.kate_division(a, x)
        assert H.fr_dec(oracle_c.kate_division(H.fr_enc(a), H.fr_enc([x])[0])) == q
        # a(X) = q(X) (X - x) + a(x): check at a random point
        z = rng.randrange(O.R_MOD)
        assert (O.eval_polynomial(q, z) * (z - x) + ev) % O.R_MOD == O.eval_polynomial(a, z)


def test_pinned_vk_of_the_reference():
    """The reference's one golden artefact: the pinned verifying key of tests/plonk_api.rs:626-1019 (k = 5, IPA
    over Vesta; fixture tests/golden/pinned_vk_plonk_api.json, made by make_pinned_vk_fixture.py).  It pins,
    against bytes the reference itself asserts on:
      * EvaluationDomain::new's derivation of extended_k and omega (domain.rs:39-73) -- same oracle code path as
        bn256, with Vesta's scalar field (modulus, S = 32, generator 5) as the FieldExt;
      * the ConstraintSystem bookkeeping of the host mirror (column counts, query indices and their order,
        enable_equality's queries, lookup table queries, permutation column order, degree -> extended_k);
      * `format!("{:?}", vk.pinned())`, the string every transcript is seeded with (plonk.rs:192-203), in the
        oracle's and in the product mirror's formatter, character for character.
    The commitments in the key are Vesta points from IPA parameters (hash-to-curve generators): they are passed
    through as given; bn256 curve arithmetic has no golden vector in the reference."""
    import halo2_pse_b200 as h
    from halo2_pse_b200.prover import pinned_debug
    from oracle import prover as OV
    from tests import plonk_cases as PC

    fx = H.load_golden("pinned_vk_plonk_api.json")
    cs = PC.plonk_api_configure()  # plonk_api.rs:389-470 -- MyCircuit::configure, statement by statement

    # the domain, through the oracle's EvaluationDomain::new restatement with Vesta's scalar field
    modulus = int(fx["scalar_modulus"], 16)
    root = pow(5, (modulus - 1) >> 32, modulus)  # pasta Fp: S = 32, multiplicative generator 5
    ek, _, omega = O.domain_roots(cs.degree(), fx["k"], modulus, root, 32)
    assert ek == fx["extended_k"] and omega == int(fx["omega"], 16)

    pts = lambda key: [(int(x, 16), int(y, 16)) for x, y in fx[key]]  # noqa: E731
    args = (fx["k"], ek, omega, pts("fixed_commitments"), pts("permutation_commitments"))
    mod = dict(base_modulus=int(fx["base_modulus"], 16), scalar_modulus=modulus)
    assert pinned_debug(cs, *args, **mod) == fx["debug"]
    assert OV.pinned_vk_debug(PC.oracle_cs(cs), *args, **mod) == fx["debug"]


def test_reference_golden_commitments():
    """THE PIN of the MSM / NTT restatements against outputs the reference itself holds: the 19 commitment points of
    its pinned verifying key (tests/plonk_api.rs:994-1017; k = 5, IPA over Vesta).  They are produced by
    keygen_vk -> `commit_lagrange` -> `best_multiexp` over `g_lagrange = g_to_lagrange(g)` (best_fft over curve
    points) with hash-to-curve generators (poly/ipa/commitment.rs:158-207).  The reference is generic over the curve
    and so is the oracle: oracle/pasta.py executes the SAME source files (oracle/bn256.py, plonk.py, prover.py) a
    second time with Vesta's constants, adds pasta's hash-to-curve, and this test reproduces

      * w = hash_to_curve("Halo2-Parameters")(&[1]) = the commitment of the never-assigned fixed column `sf` (1 * w),
      * the other 6 fixed-column commitments (multiexp_serial's windows and buckets on 33 points, arithmetic.rs:13-159;
        the group FFT and 1/n scaling of g_to_lagrange, :277-301),
      * the 12 permutation commitments (permutation/keygen.rs: cycle merging, delta^i * omega^j),
      * and with them the whole `{:?}` string of the pinned key, character for character."""
    from oracle import pasta
    from tests import plonk_cases as PC
    V, _, VV = pasta.load_vesta()
    fx = H.load_golden("pinned_vk_plonk_api.json")
    pts = lambda key: [(int(x, 16), int(y, 16)) for x, y in fx[key]]  # noqa: E731
    assert (V.R_MOD, V.Q_MOD) == (int(fx["scalar_modulus"], 16), int(fx["base_modulus"], 16))
    assert O.R_MOD != V.R_MOD and O.S == 28 and V.S == 32  # the bn256 instance is untouched

    # 1. hash-to-curve, against the one golden point that is a bare generator
    w = pasta.hash_to_curve("Halo2-Parameters")(b"\x01")
    assert w == pts("fixed_commitments")[0]

    # 2. ParamsIPA::new(5): g by hash-to-curve, g_lagrange by the group FFT (consistency: the Lagrange commitment of
    #    evaluations equals the monomial commitment of the interpolated coefficients, ipa/commitment.rs:255-301)
    params = pasta.ParamsIPA(V, fx["k"])
    dom = V.EvaluationDomain(1, fx["k"])
    vals = [(i * i + 7) % V.R_MOD for i in range(params.n)]
    assert params.commit(dom.lagrange_to_coeff(list(vals)), 5) == params.commit_lagrange(vals, 5)

    # 3. keygen of the plonk_api circuit over those parameters
    cs = PC.oracle_cs(PC.plonk_api_configure(), VV)
    # pasta Fp::ZETA (pasta_curves fields/fp.rs) -- of the two primitive cube roots of unity it is 5^(2(p-1)/3); the
    # commitment of the lookup-table column below only comes out right with this one
    zeta = 0x12CCCA834ACDBA712CAAD5DC57AAB1B01D1F8BD237AD31491DAD5EBDFDFE4AB9
    assert zeta == pow(5, 2 * (V.R_MOD - 1) // 3, V.R_MOD) and pow(zeta, 3, V.R_MOD) == 1
    a_value = 2834758237 * zeta % V.R_MOD        # common!(): Scalar::from(2834758237) * Scalar::ZETA
    fixed, copies = PC.plonk_api_keygen_inputs(fx["k"], a_value, 2, cs.blinding_factors())
    pk = VV.keygen(params, cs, fixed, copies)
    assert pk.domain.extended_k == fx["extended_k"] and pk.domain.omega == int(fx["omega"], 16)
    assert pk.fixed_commitments == pts("fixed_commitments")
    assert pk.perm_commitments == pts("permutation_commitments")
    assert pk.debug == fx["debug"]
    # the same points through every thread count of best_multiexp's chunking (arithmetic.rs:135-153) and through
    # small_multiexp (:105-125): the window rule changes with the chunk length, the sum must not
    col = fixed[4] + [1]
    bases = params.g_lagrange + [params.w]
    for threads in (1, 2, 3, 8, 33, 64):
        assert V.best_multiexp(col, bases, threads) == pts("fixed_commitments")[4], threads
    assert V.small_multiexp(col, bases) == pts("fixed_commitments")[4]
