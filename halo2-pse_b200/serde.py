"""Wire formats of the KZG parameters: ``SerdeFormat`` (helpers.rs:8-52) and ``ParamsKZG::write_custom`` /
``read_custom`` (poly/kzg/commitment.rs:142-244), paths relative to /root/reference/halo2_proofs/src.

File layout: k as u32 little-endian, 2^k points of ``g``, 2^k points of ``g_lagrange``, then ``g2`` and
``s_g2``.  In the RawBytes formats a G1 point is its 64 in-memory bytes (Montgomery limbs), which is exactly
the layout ``h2b_bases_upload`` takes: the 2^(k+1) points go from the file to the device without conversion;
``RawBytes`` additionally checks the curve equation on the GPU (``SerdeObject::read_raw``).  ``Processed`` is
``G1Affine::to_bytes`` (32 B, assumption A2 of DESIGN.md); compression and decompression (one square root in
Fq per point) run on the GPU.

The prover never touches G2: ``g2`` and ``s_g2`` are two points carried through as host integers.  Their
arithmetic (Fq2, the twist y^2 = x^3 + 3/(9 + u)) is a few lines of host code used once per setup;
encodings follow halo2curves 0.3.1 (c0 then c1; compressed form = x with the parity of y.c0 in the top bit of
the last byte -- assumption A5, only exercised by the Processed format).
"""
from __future__ import annotations

import ctypes as C
import io
from typing import BinaryIO, Optional, Tuple

import numpy as np

from . import _ffi
from ._ffi import H2B_DEVICE, H2BError
from .api import Q_MOD, R_MOD, Bases, Context, ParamsKZG

PROCESSED, RAW_BYTES, RAW_BYTES_UNCHECKED = "Processed", "RawBytes", "RawBytesUnchecked"
SIGN_BIT = 7
_R = (1 << 256) % Q_MOD
_RINV = pow(_R, -1, Q_MOD)

Fq2 = Tuple[int, int]  # c0 + c1 u, u^2 = -1


# ---- Fq2 / G2 on the host (two points per parameter set) --------------------------------------------
def _f2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)


def _f2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)


def _f2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)


def _f2_inv(a: Fq2) -> Fq2:
    d = pow((a[0] * a[0] + a[1] * a[1]) % Q_MOD, -1, Q_MOD)
    return (a[0] * d % Q_MOD, -a[1] * d % Q_MOD)


def _f2_pow(a: Fq2, e: int) -> Fq2:
    r: Fq2 = (1, 0)
    while e:
        if e & 1:
            r = _f2_mul(r, a)
        a = _f2_mul(a, a)
        e >>= 1
    return r


G2_B: Fq2 = _f2_mul((3, 0), _f2_inv((9, 1)))  # the sextic twist's constant 3 / (9 + u)
G2_GENERATOR = ((0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED,
                 0x198E9393920D483A7260BFB731FB5D25F1AA493335A9E71297E485B7AEF312C2),
                (0x12C85EA5DB8C6DEB4AAB71808DCB408FE3D1E7690C43D37B4CE6CC0166FA7DAA,
                 0x090689D0585FF075EC9E99AD690C3395BC4B313370B38EF355ACDADCD122975B))


def g2_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return _f2_mul(y, y) == _f2_add(_f2_mul(_f2_mul(x, x), x), G2_B)


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if _f2_add(y1, y2) == (0, 0):
            return None
        lam = _f2_mul(_f2_mul((3, 0), _f2_mul(x1, x1)), _f2_inv(_f2_add(y1, y1)))
    else:
        lam = _f2_mul(_f2_sub(y2, y1), _f2_inv(_f2_sub(x2, x1)))
    x3 = _f2_sub(_f2_sub(_f2_mul(lam, lam), x1), x2)
    return (x3, _f2_sub(_f2_mul(lam, _f2_sub(x1, x3)), y1))


def g2_mul(p, k: int):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc


def _f2_sqrt(a: Fq2) -> Optional[Fq2]:
    """Square root in Fq2 for q = 3 mod 4 (Adj--Rodriguez-Henriquez, algorithm 9)."""
    if a == (0, 0):
        return (0, 0)
    a1 = _f2_pow(a, (Q_MOD - 3) // 4)
    alpha = _f2_mul(_f2_mul(a1, a1), a)
    a0 = _f2_mul((alpha[0], -alpha[1] % Q_MOD), alpha)  # alpha^q * alpha
    if a0 == (Q_MOD - 1, 0):
        return None
    x0 = _f2_mul(a1, a)
    if alpha == (Q_MOD - 1, 0):
        return _f2_mul((0, 1), x0)
    b = _f2_pow(_f2_add((1, 0), alpha), (Q_MOD - 1) // 2)
    return _f2_mul(b, x0)


def _fq_raw(x: int) -> bytes:
    return (x * _R % Q_MOD).to_bytes(32, "little")


def _fq_from_raw(b: bytes) -> int:
    return int.from_bytes(b, "little") * _RINV % Q_MOD


def g2_write(p, fmt: str) -> bytes:
    if fmt == PROCESSED:
        if p is None:
            return bytes(64)
        (x0, x1), (y0, _) = p
        b = bytearray(x0.to_bytes(32, "little") + x1.to_bytes(32, "little"))
        b[63] |= (y0 & 1) << SIGN_BIT
        return bytes(b)
    if p is None:
        return bytes(128)
    (x0, x1), (y0, y1) = p
    return _fq_raw(x0) + _fq_raw(x1) + _fq_raw(y0) + _fq_raw(y1)


def g2_read(b: bytes, fmt: str):
    if fmt == PROCESSED:
        if b == bytes(64):
            return None
        bb = bytearray(b)
        sign = (bb[63] >> SIGN_BIT) & 1
        bb[63] &= 0xFF ^ (1 << SIGN_BIT)
        x = (int.from_bytes(bb[:32], "little"), int.from_bytes(bb[32:], "little"))
        y = _f2_sqrt(_f2_add(_f2_mul(_f2_mul(x, x), x), G2_B))
        if y is None or x[0] >= Q_MOD or x[1] >= Q_MOD:
            raise H2BError(_ffi.H2B_ERR_ARG, "Invalid point encoding")
        if (y[0] & 1) != sign:
            y = (-y[0] % Q_MOD, -y[1] % Q_MOD)
        return (x, y)
    if b == bytes(128):
        return None
    v = [_fq_from_raw(b[32 * i:32 * i + 32]) for i in range(4)]
    p = ((v[0], v[1]), (v[2], v[3]))
    if fmt == RAW_BYTES and not g2_is_on_curve(p):
        raise H2BError(_ffi.H2B_ERR_ARG, "Invalid point encoding")
    return p


def g2_len(fmt: str) -> int:
    return 64 if fmt == PROCESSED else 128


# ---- ParamsKZG::write_custom / read_custom -----------------------------------------------------------
def _write_g1(ctx: Context, bases: Bases, writer: BinaryIO, fmt: str) -> None:
    n = len(bases)
    if fmt == PROCESSED:
        out = np.empty(n * 32, dtype=np.uint8)
        ctx._check(ctx.lib.h2b_g1_compress(ctx.h, ctx.lib.h2b_bases_device_ptr(bases.h), n, SIGN_BIT,
                                           C.c_void_p(out.ctypes.data)))
        writer.write(out.tobytes())
    else:
        writer.write(bases.download().tobytes())  # write_raw: the limbs as they are


def _read_g1(ctx: Context, reader: BinaryIO, n: int, fmt: str) -> Bases:
    size = 32 if fmt == PROCESSED else 64
    raw = reader.read(n * size)
    if len(raw) != n * size:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "unexpected end of file")
    ok = C.c_int(1)
    if fmt == PROCESSED:
        buf = ctx.alloc(n * 64)
        arr = np.frombuffer(raw, dtype=np.uint8)
        ctx._check(ctx.lib.h2b_g1_decompress(ctx.h, C.c_void_p(arr.ctypes.data), n, SIGN_BIT, buf.ptr, C.byref(ok)))
        b = Bases(ctx, buf.ptr, n, H2B_DEVICE)
        buf.free()
    else:
        b = Bases(ctx, np.frombuffer(raw, dtype=np.uint64).reshape(n, 8), n)
        if fmt == RAW_BYTES:
            ctx._check(ctx.lib.h2b_g1_check_on_curve(ctx.h, ctx.lib.h2b_bases_device_ptr(b.h), n, C.byref(ok)))
    if not ok.value:
        b.free()
        raise H2BError(_ffi.H2B_ERR_ARG, "invalid point encoding")
    return b


def write_params(params: ParamsKZG, writer: BinaryIO, fmt: str = RAW_BYTES) -> None:
    """ParamsKZG::write_custom (poly/kzg/commitment.rs:142-158)."""
    if getattr(params, "g2", None) is None or getattr(params, "s_g2", None) is None:
        raise H2BError(_ffi.H2B_ERR_ARG, "these parameters carry no G2 points (not read from a file nor set up with s)")
    writer.write(int(params.k).to_bytes(4, "little"))
    _write_g1(params.ctx, params.g, writer, fmt)
    _write_g1(params.ctx, params.g_lagrange, writer, fmt)
    writer.write(g2_write(params.g2, fmt))
    writer.write(g2_write(params.s_g2, fmt))


def read_params(ctx: Context, reader: BinaryIO, fmt: str = RAW_BYTES, precompute: bool = False) -> ParamsKZG:
    """ParamsKZG::read_custom (poly/kzg/commitment.rs:161-244): the bases end up device-resident."""
    head = reader.read(4)
    if len(head) != 4:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "unexpected end of file")
    k = int.from_bytes(head, "little")
    if k > 28:
        raise H2BError(_ffi.H2B_ERR_ARG, "k exceeds the two-adicity of Fr")
    n = 1 << k
    g = _read_g1(ctx, reader, n, fmt)
    g_lagrange = _read_g1(ctx, reader, n, fmt)
    tail = reader.read(2 * g2_len(fmt))
    if len(tail) != 2 * g2_len(fmt):
        raise H2BError(_ffi.H2B_ERR_LENGTH, "unexpected end of file")
    params = ParamsKZG(ctx, k, g, g_lagrange)
    params.g2 = g2_read(tail[:g2_len(fmt)], fmt)
    params.s_g2 = g2_read(tail[g2_len(fmt):], fmt)
    if precompute:
        params.g.precompute()
        params.g_lagrange.precompute()
    return params


def params_to_bytes(params: ParamsKZG, fmt: str = RAW_BYTES) -> bytes:
    w = io.BytesIO()
    write_params(params, w, fmt)
    return w.getvalue()


# ---- VerifyingKey / ProvingKey write and read (plonk.rs:60-160, 300-354; poly.rs:152-177; helpers.rs:112-146) ----
def _g1_point_write(p, fmt: str) -> bytes:
    from .api import g1_encode
    from .prover import g1_to_bytes
    return g1_to_bytes(p) if fmt == PROCESSED else g1_encode([p]).tobytes()


def _g1_point_read(ctx: Context, reader: BinaryIO, fmt: str):
    from .api import g1_decode
    size = 32 if fmt == PROCESSED else 64
    raw = reader.read(size)
    if len(raw) != size:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "unexpected end of file")
    b = _read_g1(ctx, io.BytesIO(raw), 1, fmt)  # validation / decompression on the device, like the bases
    p = g1_decode(b.download())[0]
    b.free()
    return p


def _poly_write(ctx: Context, buf, count: int, writer: BinaryIO, fmt: str) -> None:
    """Polynomial::write: u32 big-endian length, then the elements (Processed: canonical repr; else raw limbs)."""
    writer.write(count.to_bytes(4, "big"))
    if fmt == PROCESSED:  # Montgomery -> canonical (to_repr) on a device copy
        tmp = ctx.clone(buf, max(count, 1) * 32)
        ctx._check(ctx.lib.h2b_fr_repr(ctx.h, tmp.ptr, H2B_DEVICE, count, 0, None))
        limbs = tmp.download(count)
        tmp.free()
    else:
        limbs = buf.download(count)
    writer.write(limbs.tobytes())


def _poly_read(ctx: Context, reader: BinaryIO, fmt: str, expect: int):
    head = reader.read(4)
    if len(head) != 4:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "unexpected end of file")
    count = int.from_bytes(head, "big")
    if count != expect:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "polynomial length does not match the domain")
    raw = reader.read(count * 32)
    if len(raw) != count * 32:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "unexpected end of file")
    buf = ctx.upload_fr(np.frombuffer(raw, dtype=np.uint64).reshape(count, 4))
    if fmt != RAW_BYTES_UNCHECKED:  # from_repr / read_raw reject values >= r: checked (and converted) on the device
        ok = C.c_int(1)
        ctx._check(ctx.lib.h2b_fr_repr(ctx.h, buf.ptr, H2B_DEVICE, count, 1 if fmt == PROCESSED else 2, C.byref(ok)))
        if not ok.value:
            buf.free()
            raise H2BError(_ffi.H2B_ERR_ARG, "Invalid prime field point encoding")
    return buf


def _poly_vec_write(ctx, bufs, count, writer, fmt) -> None:
    writer.write(len(bufs).to_bytes(4, "big"))
    for b in bufs:
        _poly_write(ctx, b, count, writer, fmt)


def _poly_vec_read(ctx, reader, fmt, expect, expect_len):
    head = reader.read(4)
    if len(head) != 4 or int.from_bytes(head, "big") != expect_len:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "polynomial vector does not match the constraint system")
    return [_poly_read(ctx, reader, fmt, expect) for _ in range(expect_len)]


def write_vk(pk, writer: BinaryIO, fmt: str = RAW_BYTES) -> None:
    """VerifyingKey::write (plonk.rs:73-91); the mirror has no selectors."""
    writer.write(int(pk.k).to_bytes(4, "big"))
    writer.write(len(pk.fixed_commitments).to_bytes(4, "big"))
    for c in pk.fixed_commitments:
        writer.write(_g1_point_write(c, fmt))
    for c in pk.perm_commitments:  # permutation::VerifyingKey::write
        writer.write(_g1_point_write(c, fmt))


def write_pk(pk, writer: BinaryIO, fmt: str = RAW_BYTES) -> None:
    """ProvingKey::write (plonk.rs:307-318)."""
    ctx, n, ext = pk.domain.ctx, pk.n, pk.domain.extended_len()
    write_vk(pk, writer, fmt)
    for b in (pk.l0, pk.l_last, pk.l_active_row):
        _poly_write(ctx, b, ext, writer, fmt)
    _poly_vec_write(ctx, pk.fixed_values, n, writer, fmt)
    _poly_vec_write(ctx, pk.fixed_polys, n, writer, fmt)
    _poly_vec_write(ctx, pk.fixed_cosets, ext, writer, fmt)
    _poly_vec_write(ctx, pk.permutations, n, writer, fmt)        # permutation::ProvingKey::write
    _poly_vec_write(ctx, pk.permutation_polys, n, writer, fmt)
    _poly_vec_write(ctx, pk.permutation_cosets, ext, writer, fmt)


def read_pk(params: ParamsKZG, cs, reader: BinaryIO, fmt: str = RAW_BYTES):
    """ProvingKey::read::<_, ConcreteCircuit> (plonk.rs:331-354): `cs` is what ConcreteCircuit::configure builds.
    The polynomials go straight to the device; the vk hash is recomputed as from_parts does."""
    import hashlib
    from .api import EvaluationDomain
    from .plonk import Evaluator
    from .prover import ProvingKey, _lookup_compress_graphs, pinned_debug
    ctx = params.ctx
    head = reader.read(8)
    if len(head) != 8:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "unexpected end of file")
    k, nfixed = int.from_bytes(head[:4], "big"), int.from_bytes(head[4:], "big")
    if k != params.k or nfixed != cs.num_fixed_columns:
        raise H2BError(_ffi.H2B_ERR_LENGTH, "the key does not match the parameters / constraint system")
    pk = ProvingKey()
    pk.params, pk.cs, pk.k, pk.n = params, cs, k, 1 << k
    pk.domain = EvaluationDomain(ctx, cs.degree(), k)
    n, ext, ncols = pk.n, pk.domain.extended_len(), len(cs.permutation.columns)
    pk.fixed_commitments = [_g1_point_read(ctx, reader, fmt) for _ in range(nfixed)]
    pk.perm_commitments = [_g1_point_read(ctx, reader, fmt) for _ in range(ncols)]
    pk.l0, pk.l_last, pk.l_active_row = (_poly_read(ctx, reader, fmt, ext) for _ in range(3))
    pk.fixed_values = _poly_vec_read(ctx, reader, fmt, n, nfixed)
    pk.fixed_polys = _poly_vec_read(ctx, reader, fmt, n, nfixed)
    pk.fixed_cosets = _poly_vec_read(ctx, reader, fmt, ext, nfixed)
    pk.permutations = _poly_vec_read(ctx, reader, fmt, n, ncols)
    pk.permutation_polys = _poly_vec_read(ctx, reader, fmt, n, ncols)
    pk.permutation_cosets = _poly_vec_read(ctx, reader, fmt, ext, ncols)
    pk.ev = Evaluator(cs)
    pk.lookup_compress = _lookup_compress_graphs(cs)
    pk.pinned = pinned_debug(cs, k, pk.domain.extended_k, pk.domain.constant("omega"), pk.fixed_commitments,
                             pk.perm_commitments)
    hsh = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    hsh.update(len(pk.pinned).to_bytes(8, "little") + pk.pinned.encode())
    pk.transcript_repr = int.from_bytes(hsh.digest(), "little") % R_MOD
    return pk
