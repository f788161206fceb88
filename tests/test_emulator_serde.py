"""SerdeFormat wire formats of ParamsKZG (helpers.rs:8-52, poly/kzg/commitment.rs:142-244) on the CPU kernel
emulator: round trips in all three formats, point validation, the host-side G2 arithmetic."""
import io

import numpy as np
import pytest

import halo2_pse_b200 as h
from halo2_pse_b200 import serde
from oracle import bn256 as O
from tests import helpers as H

S_TOXIC = 0xABCDEF0123456789


@pytest.fixture(scope="module")
def params(emu_ctx):
    return h.ParamsKZG.setup(emu_ctx, 4, S_TOXIC)


def test_g2_generator_and_arithmetic():
    G = serde.G2_GENERATOR
    assert serde.g2_is_on_curve(G)
    assert serde.g2_mul(G, O.R_MOD) is None                      # the generator has order r
    p, q = serde.g2_mul(G, 5), serde.g2_mul(G, 7)
    assert serde.g2_is_on_curve(p) and serde.g2_add(p, q) == serde.g2_mul(G, 12)
    for fmt in (serde.PROCESSED, serde.RAW_BYTES, serde.RAW_BYTES_UNCHECKED):
        for pt in (G, p, serde.g2_add(q, q), None):
            assert serde.g2_read(serde.g2_write(pt, fmt), fmt) == pt
    with pytest.raises(h.H2BError):
        serde.g2_read(serde.g2_write(((1, 2), (3, 4)), serde.RAW_BYTES), serde.RAW_BYTES)


@pytest.mark.parametrize("fmt", [serde.PROCESSED, serde.RAW_BYTES, serde.RAW_BYTES_UNCHECKED])
def test_params_round_trip(emu_ctx, params, fmt):
    blob = serde.params_to_bytes(params, fmt)
    n = params.n
    point = 32 if fmt == serde.PROCESSED else 64
    assert len(blob) == 4 + 2 * n * point + 2 * serde.g2_len(fmt)
    assert blob[:4] == (4).to_bytes(4, "little")
    back = serde.read_params(emu_ctx, io.BytesIO(blob), fmt)
    assert back.k == params.k
    assert (back.g.download() == params.g.download()).all()
    assert (back.g_lagrange.download() == params.g_lagrange.download()).all()
    assert back.g2 == serde.G2_GENERATOR and back.s_g2 == serde.g2_mul(serde.G2_GENERATOR, S_TOXIC)
    assert serde.params_to_bytes(back, fmt) == blob
    # the loaded parameters commit like the originals
    a = H.rand_fr_limbs(3, n)
    assert back.commit(a) == params.commit(a) and back.commit_lagrange(a) == params.commit_lagrange(a)


def test_processed_points_are_to_bytes_of_the_oracle_points(emu_ctx, params):
    from oracle import prover as OV
    blob = serde.params_to_bytes(params, serde.PROCESSED)
    pts = h.g1_decode(params.g.download())
    for i, p in enumerate(pts):
        assert blob[4 + 32 * i:4 + 32 * i + 32] == OV.g1_to_bytes(p)
    assert pts[0] == (1, 2) and pts[1] == O.g1_mul(O.G1_GEN, S_TOXIC)


def test_invalid_points_are_rejected(emu_ctx, params):
    raw = bytearray(serde.params_to_bytes(params, serde.RAW_BYTES))
    raw[4 + 64 * 3 + 5] ^= 1                                     # g[3] leaves the curve
    with pytest.raises(h.H2BError):
        serde.read_params(emu_ctx, io.BytesIO(bytes(raw)), serde.RAW_BYTES)
    serde.read_params(emu_ctx, io.BytesIO(bytes(raw)), serde.RAW_BYTES_UNCHECKED)  # unchecked by definition
    # non-canonical limbs: x + q satisfies the Montgomery curve equation but SerdeObject::read_raw rejects
    # coordinates >= q before it looks at the curve (the group law's == / is_zero need reduced limbs)
    raw = bytearray(serde.params_to_bytes(params, serde.RAW_BYTES))
    for coord in (0, 32):
        bad = bytearray(raw)
        off = 4 + 64 * 2 + coord
        v = int.from_bytes(bad[off:off + 32], "little") + O.Q_MOD
        assert v < 1 << 256
        bad[off:off + 32] = v.to_bytes(32, "little")
        with pytest.raises(h.H2BError):
            serde.read_params(emu_ctx, io.BytesIO(bytes(bad)), serde.RAW_BYTES)
        serde.read_params(emu_ctx, io.BytesIO(bytes(bad)), serde.RAW_BYTES_UNCHECKED)
    comp = bytearray(serde.params_to_bytes(params, serde.PROCESSED))
    # x = 4 has x^3 + 3 = 67, a non-residue mod q?  find a non-point abscissa by search
    x = next(v for v in range(2, 50) if pow((v ** 3 + 3) % O.Q_MOD, (O.Q_MOD - 1) // 2, O.Q_MOD) != 1)
    comp[4 + 32 * 2:4 + 32 * 3] = x.to_bytes(32, "little")
    with pytest.raises(h.H2BError):
        serde.read_params(emu_ctx, io.BytesIO(bytes(comp)), serde.PROCESSED)
    with pytest.raises(h.H2BError):
        serde.read_params(emu_ctx, io.BytesIO(bytes(comp[:100])), serde.PROCESSED)  # truncated


@pytest.mark.parametrize("fmt", [serde.PROCESSED, serde.RAW_BYTES, serde.RAW_BYTES_UNCHECKED])
def test_proving_key_round_trip_proves_the_same_bytes(emu_ctx, fmt):
    """ProvingKey::write / read (plonk.rs:307-354): a key read back from bytes yields the same proof."""
    from tests import plonk_cases as PC
    k = 5
    params = h.ParamsKZG.setup(emu_ctx, k, PC.S_TOXIC)
    fixed, advice, copies = PC.lookup_circuit(k)
    pk = h.keygen(params, PC.build_lookup_cs(), fixed, copies)
    w = io.BytesIO()
    serde.write_pk(pk, w, fmt)
    blob = w.getvalue()
    n, ext, point = 1 << k, pk.domain.extended_len(), (32 if fmt == serde.PROCESSED else 64)
    assert len(blob) == 8 + (3 + 2) * point + 3 * (4 + 32 * ext) + 2 * (4 + 3 * (4 + 32 * n)) + (4 + 3 * (4 + 32 * ext)) \
        + 2 * (4 + 2 * (4 + 32 * n)) + (4 + 2 * (4 + 32 * ext))
    back = serde.read_pk(params, PC.build_lookup_cs(), io.BytesIO(blob), fmt)
    assert back.pinned == pk.pinned and back.transcript_repr == pk.transcript_repr
    w2 = io.BytesIO()
    serde.write_pk(back, w2, fmt)
    assert w2.getvalue() == blob

    def prove(key):
        t = h.Blake2bWrite()
        h.create_proof(params, key, [lambda phase, ch: dict(enumerate(advice))], [[]], h.XorShiftRng(b"\x01" * 16), t)
        return t.finalize()

    assert prove(back) == prove(pk)
    with pytest.raises(h.H2BError):
        serde.read_pk(params, PC.build_cs("bench"), io.BytesIO(blob), fmt)  # another circuit's constraint system
    with pytest.raises(h.H2BError):
        serde.read_pk(params, PC.build_lookup_cs(), io.BytesIO(blob[:-7]), fmt)
    if fmt != serde.RAW_BYTES_UNCHECKED:
        bad = bytearray(blob)
        off = 8 + 5 * point + 4  # first element of l0: make it >= r
        bad[off:off + 32] = b"\xff" * 32
        with pytest.raises(h.H2BError):
            serde.read_pk(params, PC.build_lookup_cs(), io.BytesIO(bytes(bad)), fmt)
    pk.free()
    back.free()
