"""GPU parity of the quotient evaluation through the C ABI against the direct
big-integer oracle, plus size-independent properties at the bench size."""
import numpy as np
import pytest

import halo2_pse_b200 as h
from tests import plonk_cases as PC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,k,ncirc", [("bench", 4, 1), ("bench", 7, 1), ("rich", 4, 1), ("rich", 6, 2),
                                             ("gates_only", 5, 1)])
def test_evaluate_h_vs_oracle(gpu_ctx, variant, k, ncirc):
    PC.check_evaluate_h(gpu_ctx, variant, k, seed=200 + k, n_circuits=ncirc)


def test_evaluate_h_without_prefetch_hoisting(gpu_ctx, monkeypatch):
    monkeypatch.setenv("H2B_EVALH_NO_PREFETCH", "1")
    PC.check_evaluate_h(gpu_ctx, "rich", 5, seed=8)


def test_evaluate_h_slot_overflow_path(gpu_ctx, monkeypatch):
    monkeypatch.setenv("H2B_EVALH_SMEM_CAP", "4096")
    PC.check_evaluate_h(gpu_ctx, "rich", 5, seed=9)


@pytest.mark.parametrize("k", [5, 6])
def test_create_proof_bytes_equal_the_oracle(gpu_ctx, k):
    """Identical verifying key and identical proof bytes for the same circuit and rng seed."""
    PC.check_bench_proof_bytes(gpu_ctx, k, seed=bytes(range(16)))


def test_create_proof_k14_is_accepted_by_the_reference_verifier(gpu_ctx):
    """Size-independent property at a size the big-integer prover cannot reach: the proof of a 2^14-row
    circuit verifies under the restated reference verifier; a flipped bit or another witness does not."""
    from types import SimpleNamespace
    from oracle import prover as OV
    k = 14
    params = h.ParamsKZG.setup(gpu_ctx, k, PC.S_TOXIC, precompute=True)
    cs = PC.build_cs("bench")
    fixed, advice, copies = PC.bench_circuit_limbs(k, 0xC0FFEE)
    pk = h.keygen(params, cs, fixed, copies)
    vk = PC.oracle_vk_of(pk)
    vparams = SimpleNamespace(g=[h.g1_decode(params.g.download()[:1])[0]])

    def prove(adv, seed):
        t = h.Blake2bWrite()
        h.create_proof(params, pk, [lambda phase, ch: dict(enumerate(adv))], [[]], h.CounterRng(seed), t)
        return t.finalize()

    proof = prove(advice, 1)
    assert OV.verify_proof(vparams, PC.S_TOXIC, vk, [[]], proof)
    assert proof == prove(advice, 1) and proof != prove(advice, 2)
    bad = bytearray(proof)
    bad[100] ^= 4
    assert not OV.verify_proof(vparams, PC.S_TOXIC, vk, [[]], bytes(bad))
    wrong = [a.copy() for a in advice]
    wrong[2][10] = wrong[2][11]
    assert not OV.verify_proof(vparams, PC.S_TOXIC, vk, [[]], prove(wrong, 1))
    pk.free()


@pytest.mark.parametrize("n,distinct", [(1, 1), (7, 3), (1000, 37), (5000, 5000), (1 << 15, 1000)])
def test_lookup_permute_vs_oracle(gpu_ctx, n, distinct):
    PC.check_lookup_permute(gpu_ctx, n, seed=n, distinct=distinct)


@pytest.mark.parametrize("k", [5, 6])
def test_create_proof_with_a_lookup_equals_the_oracle(gpu_ctx, k):
    PC.check_lookup_proof_bytes(gpu_ctx, k)


@pytest.mark.parametrize("fmt", ["Processed", "RawBytes", "RawBytesUnchecked"])
def test_params_serde_round_trip(gpu_ctx, fmt):
    """ParamsKZG::write_custom / read_custom in the three SerdeFormats (k = 10); compressed points equal the
    oracle's to_bytes; an off-curve point is rejected by the checked formats."""
    import io
    from halo2_pse_b200 import serde
    from oracle import bn256 as O
    from oracle import prover as OV
    params = h.ParamsKZG.setup(gpu_ctx, 10, PC.S_TOXIC)
    blob = serde.params_to_bytes(params, fmt)
    back = serde.read_params(gpu_ctx, io.BytesIO(blob), fmt)
    assert (back.g.download() == params.g.download()).all()
    assert (back.g_lagrange.download() == params.g_lagrange.download()).all()
    assert serde.params_to_bytes(back, fmt) == blob and back.s_g2 == params.s_g2
    if fmt == "Processed":
        pts = h.g1_decode(params.g_lagrange.download()[:64])
        off = 4 + 32 * params.n
        assert all(blob[off + 32 * i:off + 32 * i + 32] == OV.g1_to_bytes(p) for i, p in enumerate(pts))
    else:
        bad = bytearray(blob)
        bad[4 + 64 * 100 + 40] ^= 8
        if fmt == "RawBytes":
            with pytest.raises(h.H2BError):
                serde.read_params(gpu_ctx, io.BytesIO(bytes(bad)), fmt)


@pytest.mark.parametrize("which", ["bench", "lookup"])
def test_shplonk_proof_bytes_equal_the_oracle(gpu_ctx, which):
    PC.check_shplonk_proof_bytes(gpu_ctx, which, k=6)


@pytest.mark.parametrize("variant,k,ncirc", [("bench", 12, 1), ("rich", 11, 2)])
def test_evaluate_h_vs_cpp_restatement_at_larger_k(gpu_ctx, oracle_c, variant, k, ncirc):
    """2^13 .. 2^14 extended rows: every row of h equal to the C++ restatement of evaluate_h."""
    cs = PC.build_cs(variant)
    case = PC.random_case_limbs(cs, k, seed=k, n_circuits=ncirc)
    want = PC.oracle_c_h(oracle_c, cs, case)
    got = PC.device_h_limbs(gpu_ctx, cs, case)
    assert got.shape == want.shape and (got == want).all()


def test_create_proof_with_two_phases_equals_the_oracle(gpu_ctx):
    PC.check_phases_proof_bytes(gpu_ctx, 6)
