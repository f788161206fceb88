"""The prover/verifier oracle (oracle/prover.py) against itself and against known answers: transcript
hashing vs hashlib, the vk Debug form vs the shape of the reference's pinned fixture
(tests/plonk_api.rs:624-1020), proofs accepted by the restated verifier, tampering rejected."""
import hashlib

import pytest

from oracle import bn256 as O
from oracle import prover as OV
from tests import plonk_cases as PC

S_TOXIC = 0x1234567890ABCDEF1234567890ABCDEF


@pytest.fixture(scope="module")
def bench_k5():
    k = 5
    params = O.ParamsKZG.setup(k, S_TOXIC)
    cs = PC.oracle_cs(PC.build_cs("bench"))
    fixed, advice, copies = PC.bench_circuit(k, a=0xDEADBEEF)
    pk = OV.keygen(params, cs, fixed, copies)
    return params, cs, pk, advice


def _prove(params, pk, advice, seed=b"\x07" * 16, instances=([],)):
    t = OV.Blake2bWrite()
    witness = lambda phase, ch: {i: col for i, col in enumerate(advice)}  # noqa: E731
    OV.create_proof(params, pk, [witness], [list(i) for i in instances], OV.XorShiftRng(seed), t)
    return t.finalize()


def test_xorshift_known_answers():
    # rand_xorshift's documented behaviour: an all-zero seed is replaced by 0x0BAD5EED x 4
    r = OV.XorShiftRng(bytes(16))
    assert (r.x, r.y, r.z, r.w) == (0x0BAD5EED,) * 4
    # the xorshift128 recurrence on the canonical seed of Marsaglia's paper
    r = OV.XorShiftRng(b"".join(v.to_bytes(4, "little") for v in (123456789, 362436069, 521288629, 88675123)))
    assert [r.next_u32() for _ in range(3)] == [3701687786, 458299110, 2500872618]


def test_transcript_challenge_is_blake2b_wide_reduction():
    t = OV.Blake2bWrite()
    t.common_scalar(5)
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
    h.update(b"\x02" + (5).to_bytes(32, "little") + b"\x00")
    assert t.squeeze_challenge_scalar() == int.from_bytes(h.digest(), "little") % O.R_MOD
    # the challenge prefix stays absorbed (transcript.rs:372-377)
    h.update(b"\x00")
    assert t.squeeze_challenge_scalar() == int.from_bytes(h.digest(), "little") % O.R_MOD


def test_point_compression_round_trip():
    for kk in (1, 2, 3, 12345, O.R_MOD - 1):
        p = O.g1_mul(O.G1_GEN, kk)
        assert OV.g1_from_bytes(OV.g1_to_bytes(p)) == p
        assert OV.g1_from_bytes(OV.g1_to_bytes(O.g1_neg(p))) == O.g1_neg(p)
    assert OV.g1_to_bytes(None) == bytes(32) and OV.g1_from_bytes(bytes(32)) is None


def test_pinned_vk_debug_shape(bench_k5):
    _, _, pk, _ = bench_k5
    d = pk.debug
    assert d.startswith('PinnedVerificationKey { base_modulus: "0x30644e72e131a029b85045b68181585d97816a916871ca8d'
                        '3c208c16d87cfd47", scalar_modulus: "0x30644e72e131a029b85045b68181585d2833e84879b9709143e1'
                        'f593f0000001", domain: PinnedEvaluationDomain { k: 5, extended_k: 7, omega: 0x')
    assert "cs: PinnedConstraintSystem { num_fixed_columns: 4, num_advice_columns: 3, num_instance_columns: 0, " \
           "num_selectors: 0, gates: [Sum(Sum(Sum(Product(Advice { query_index: 0, column_index: 0, rotation: " \
           "Rotation(0) }, Fixed { query_index: 0, column_index: 1, rotation: Rotation(0) })," in d
    assert "permutation: Argument { columns: [Column { index: 0, column_type: Advice }, Column { index: 1, " \
           "column_type: Advice }, Column { index: 2, column_type: Advice }] }, lookups: [], constants: [], " \
           "minimum_degree: Some(5) }" in d
    assert d.count("(0x") == 4 + 3  # fixed + permutation commitments
    assert d.endswith(")] } }")


def test_permutation_assembly_cycles():
    asm = OV.PermutationAssembly(4, [(0, 0), (0, 1)])
    asm.copy((0, 0), 0, (0, 1), 2)
    asm.copy((0, 1), 2, (0, 0), 3)
    asm.copy((0, 0), 0, (0, 0), 3)  # already in the same cycle: no-op
    # the three cells form one cycle under the mapping
    seen, cur = [], (0, 0)
    for _ in range(3):
        seen.append(cur)
        cur = asm.mapping[cur[0]][cur[1]]
    assert cur == (0, 0) and sorted(seen) == [(0, 0), (0, 3), (1, 2)]


def test_bench_circuit_proof_verifies(bench_k5):
    params, cs, pk, advice = bench_k5
    proof = _prove(params, pk, advice)
    # 3 advice + 1 permutation product + 1 random + 4 h pieces + 2 GWC witnesses (points x, x*omega)
    # + 3 advice + 4 fixed + 1 random + 3 sigma + 2 product evals
    assert len(proof) == 32 * (3 + 1 + 1 + 4 + 2) + 32 * (3 + 4 + 1 + 3 + 2)
    assert OV.verify_proof(params, S_TOXIC, pk, [[]], proof)
    assert proof == _prove(params, pk, advice)  # deterministic in the rng seed
    assert proof != _prove(params, pk, advice, seed=b"\x08" * 16)


def test_tampered_proofs_are_rejected(bench_k5):
    params, cs, pk, advice = bench_k5
    proof = bytearray(_prove(params, pk, advice))
    for pos in (40, 32 * 11 + 3, len(proof) - 5):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert not OV.verify_proof(params, S_TOXIC, pk, [[]], bytes(bad))
    assert not OV.verify_proof(params, S_TOXIC + 1, pk, [[]], bytes(proof))
    assert not OV.verify_proof(params, S_TOXIC, pk, [[]], bytes(proof[:-32]))


def test_unsatisfied_witness_does_not_verify(bench_k5):
    params, cs, pk, advice = bench_k5
    bad = [list(c) for c in advice]
    bad[2][4] = (bad[2][4] + 1) % O.R_MOD  # break one gate and one copy constraint
    assert not OV.verify_proof(params, S_TOXIC, pk, [[]], _prove(params, pk, bad))


@pytest.fixture(scope="module")
def lookup_k5():
    k = 5
    params = O.ParamsKZG.setup(k, S_TOXIC)
    cs = PC.oracle_cs(PC.build_lookup_cs())
    fixed, advice, copies = PC.lookup_circuit(k)
    return params, cs, OV.keygen(params, cs, fixed, copies), advice


def test_permute_expression_pair_known_case():
    class NoRng:
        def next_u64(self):
            return 1
    inp = [3, 1, 3, 3, 2, 1]
    tab = [1, 2, 3, 4, 5, 6]
    pi, pt = OV.permute_expression_pair(inp, tab, 6, 0, NoRng())
    assert pi[:6] == [1, 1, 2, 3, 3, 3]
    # first occurrences carry their own value; leftovers {4, 5, 6} ascending go to the repeated rows 5, 4, 1
    assert pt[:6] == [1, 6, 2, 3, 5, 4]
    assert len(pi) == len(pt) == 7
    with pytest.raises(AssertionError):
        OV.permute_expression_pair([9], [1], 1, 0, NoRng())


def test_lookup_circuit_proof_verifies(lookup_k5):
    params, cs, pk, advice = lookup_k5
    assert cs.degree() == 5 and len(cs.lookups) == 1
    proof = _prove(params, pk, advice)
    assert OV.verify_proof(params, S_TOXIC, pk, [[]], proof)
    bad = bytearray(proof)
    bad[32 * 3 + 1] ^= 2  # inside the permuted-table commitment
    assert not OV.verify_proof(params, S_TOXIC, pk, [[]], bytes(bad))


def test_lookup_outside_the_table_fails(lookup_k5):
    params, cs, pk, advice = lookup_k5
    bad = [list(c) for c in advice]
    bad[0][0] = 1 << 40  # q_lk = 1 on row 0 and 2^40 is not in the table
    with pytest.raises(AssertionError):
        _prove(params, pk, bad)


def test_lagrange_interpolate_and_vanishing():
    pts = [3, 7, 11, 12345]
    poly = [5, 0, 9, 2]
    evals = [O.eval_polynomial(poly, p) for p in pts]
    assert OV.lagrange_interpolate(pts, evals) == poly
    assert OV.lagrange_interpolate([4], [9]) == [9]
    assert OV.evaluate_vanishing_polynomial(pts, 7) == 0
    assert OV.evaluate_vanishing_polynomial([1, 2], 5) == 12


@pytest.mark.parametrize("which", ["bench", "lookup"])
def test_shplonk_proofs_verify(bench_k5, lookup_k5, which):
    params, cs, pk, advice = bench_k5 if which == "bench" else lookup_k5
    t = OV.Blake2bWrite()
    OV.create_proof(params, pk, [lambda phase, ch: dict(enumerate(advice))], [[]], OV.XorShiftRng(b"\x05" * 16), t,
                    multiopen="shplonk")
    proof = t.finalize()
    assert OV.verify_proof(params, S_TOXIC, pk, [[]], proof, multiopen="shplonk")
    assert not OV.verify_proof(params, S_TOXIC, pk, [[]], proof, multiopen="gwc")
    for pos in (10, len(proof) - 40, len(proof) - 3):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert not OV.verify_proof(params, S_TOXIC, pk, [[]], bytes(bad), multiopen="shplonk")
    # always two opening points whatever the number of rotation sets
    gwc = _prove(params, pk, advice)
    assert len(proof) <= len(gwc)


@pytest.mark.parametrize("variant,ncirc", [("bench", 1), ("rich", 2)])
def test_cpp_evaluate_h_matches_the_big_integer_oracle(oracle_c, variant, ncirc):
    """oracle/ref_cpu.cpp's restatement of evaluate_h (GraphEvaluator interpreter, permutation and lookup
    loops on threads) against the direct expression-tree evaluation of oracle/plonk.py."""
    from tests import helpers as H
    cs = PC.build_cs(variant)
    case = PC.random_case(cs, 4, seed=77, n_circuits=ncirc)
    want = PC.oracle_h(cs, case)
    got = H.fr_dec(PC.oracle_c_h(oracle_c, cs, case, threads=3))
    assert got == want
