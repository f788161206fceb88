"""Parity tests proper: the CUDA path (libhalo2b200.so through its C ABI) against
the oracle on the same seeded inputs, against the committed golden fixtures, and
-- at BASELINE.json's full sizes -- through size-independent properties.
Bit-exact everywhere: this is integer arithmetic with canonical outputs.

Run on the B200 box:  python -m pytest tests -m gpu
"""
import random

import numpy as np
import pytest

import halo2_pse_b200 as h
from oracle import bn256 as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
KAT = H.load_golden("kat_bn256.json")["vectors"]


def _np(hexs, width):
    return np.frombuffer(bytes.fromhex(hexs), dtype=np.uint64).reshape(-1, width).copy()


# ---------------------------------------------------------------------------
# field and group law (device code) vs big integers
# ---------------------------------------------------------------------------
def test_device_field_ops(gpu_ctx):
    rng = random.Random(1)
    for field, mod in ((0, O.R_MOD), (1, O.Q_MOD)):
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, (1 << 253) % mod, (mod + 1) // 2]
        a = [rng.randrange(mod) for _ in range(4000)] + edge + edge
        b = [rng.randrange(mod) for _ in range(4000)] + edge + edge[::-1]
        A, B = H.to_limbs(a, mod), H.to_limbs(b, mod)
        for op, f in ((0, lambda x, y: x * y % mod), (1, lambda x, y: (x + y) % mod),
                      (2, lambda x, y: (x - y) % mod), (3, lambda x, y: x * x % mod),
                      (6, lambda x, y: -x % mod),
                      (8, lambda x, y: (x * y - (x + y) * (x - y)) % mod),   # mul_sub: two products, one reduction
                      (9, lambda x, y: x * y % mod)):                         # mul_shoup: y as a fixed multiplier
            out = np.zeros_like(A)
            gpu_ctx._check(gpu_ctx.lib.h2b_test_field_op(gpu_ctx.h, field, op, A.ctypes.data, B.ctypes.data,
                                                         out.ctypes.data, len(a)))
            assert H.from_limbs(out, mod) == [f(x, y) for x, y in zip(a, b)], (field, op)
        # Montgomery conversions and inversion
        nz = [x or 1 for x in a[:64]]
        NZ = H.to_limbs(nz, mod)
        out = np.zeros_like(NZ)
        gpu_ctx._check(gpu_ctx.lib.h2b_test_field_op(gpu_ctx.h, field, 7, NZ.ctypes.data, NZ.ctypes.data,
                                                     out.ctypes.data, len(nz)))
        assert H.from_limbs(out, mod) == [pow(x, -1, mod) for x in nz]


def test_device_group_law(gpu_ctx, oracle_c):
    rng = random.Random(2)
    n = 512
    ka = [rng.randrange(1, 1 << 64) for _ in range(n)]
    kb = [rng.randrange(1, 1 << 64) for _ in range(n)]
    # exceptional cases: P + P, P + (-P), identity operands
    kb[0] = ka[0]
    A, B = oracle_c.g1_mul_gen(ka), oracle_c.g1_mul_gen(kb)
    pa, pb = H.g1_dec(A), H.g1_dec(B)
    pb[1] = O.g1_neg(pa[1])
    pa[2] = None
    pb[3] = None
    pa[4] = pb[4] = None
    A, B = H.g1_enc(pa), H.g1_enc(pb)
    for op, f in ((0, lambda x, y: O.g1_add(x, y)), (1, lambda x, y: O.g1_double(x)),
                  (2, lambda x, y: O.g1_add(x, y))):
        out = np.zeros_like(A)
        gpu_ctx._check(gpu_ctx.lib.h2b_test_g1_op(gpu_ctx.h, op, A.ctypes.data, B.ctypes.data, out.ctypes.data, n))
        assert H.g1_dec(out) == [f(x, y) for x, y in zip(pa, pb)], op


# ---------------------------------------------------------------------------
# best_fft                                                    arithmetic.rs:171
# ---------------------------------------------------------------------------
def test_golden_best_fft(gpu_ctx):
    for v in KAT["best_fft"]:
        a = _np(v["in"], 4)
        gpu_ctx.best_fft(a, _np(v["omega"], 4), v["log_n"])
        assert a.tobytes().hex() == v["out"], v["log_n"]


@pytest.mark.parametrize("k", list(range(0, 21)))
def test_best_fft_vs_oracle_every_k(gpu_ctx, oracle_c, k):
    a = H.rand_fr_limbs(k, 1 << k)
    w = H.fr_enc([O.omega_for(k)])[0]
    want = oracle_c.best_fft(a, w, k)
    got = a.copy()
    gpu_ctx.best_fft(got, w.reshape(1, 4), k)
    assert (got == want).all()
    # the inverse root as well (every iFFT of the prover)
    wi = H.fr_enc([pow(O.omega_for(k), -1, O.R_MOD)])[0]
    want = oracle_c.best_fft(a, wi, k)
    got = a.copy()
    gpu_ctx.best_fft(got, wi.reshape(1, 4), k)
    assert (got == want).all()


def test_best_fft_edge_values(gpu_ctx, oracle_c):
    k = 10
    w = H.fr_enc([O.omega_for(k)])[0]
    for vals in ([0] * 1024, [O.R_MOD - 1] * 1024, [1] + [0] * 1023, [0] * 1023 + [1]):
        a = H.fr_enc(vals)
        want = oracle_c.best_fft(a, w, k)
        gpu_ctx.best_fft(a, w.reshape(1, 4), k)
        assert (a == want).all()


def test_best_fft_rejects_bad_input(gpu_ctx):
    a = H.rand_fr_limbs(0, 8)
    with pytest.raises(h.H2BError) as e:  # arithmetic.rs:184
        gpu_ctx.best_fft(a, O.omega_for(4), 4)
    assert e.value.code == h.H2B_ERR_LENGTH
    with pytest.raises(h.H2BError) as e:
        gpu_ctx.best_fft(a, O.omega_for(4), 3)
    assert e.value.code == h.H2B_ERR_BAD_OMEGA


def test_best_fft_batched_columns(gpu_ctx, oracle_c):
    k, ncols = 13, 5
    n = 1 << k
    stride = n + 32
    w = H.fr_enc([O.omega_for(k)])[0]
    cols = [H.rand_fr_limbs(c, n) for c in range(ncols)]
    buf = gpu_ctx.alloc(ncols * stride * 32)
    for c in range(ncols):
        buf.upload(cols[c], c * stride * 32)
    gpu_ctx.best_fft_device(buf, w.reshape(1, 4), k, ncols, stride)
    for c in range(ncols):
        assert (buf.download(n, c * stride * 32) == oracle_c.best_fft(cols[c], w, k)).all()
    buf.free()


@pytest.mark.parametrize("k", [22, 24])
def test_best_fft_large_properties(gpu_ctx, oracle_c, k):
    """Full benchmark size: (i) inverse(forward(a)) == n*a everywhere, checked on device-downloaded
    slices; (ii) linearity / spot values: output K of the forward transform equals the Horner
    evaluation of the input at omega^K for a few K (computed by the oracle on the host)."""
    n = 1 << k
    buf = gpu_ctx.synth_scalars(n, 7, 0)
    a = buf.download(n)
    w, wi = O.omega_for(k), pow(O.omega_for(k), -1, O.R_MOD)
    gpu_ctx.best_fft_device(buf, w, k)
    fwd = buf.download(n)
    # spot check 3 outputs against sum_j a_j w^(jK) computed with the C oracle: evaluate via a
    # size-n transform of the oracle would take too long at k=24, so use the decimated identity
    # X[K] = sum_{r<R} w^(rK) * (sum_q a[qR+r] (w^R)^(qK)) with R = 2^(k-16): 2^(k-16) transforms of 2^16.
    if k <= 22:
        R = 1 << (k - 16)
        wR = pow(w, R, O.R_MOD)
        subs = [oracle_c.best_fft(a[r::R], H.fr_enc([wR])[0], 16) for r in range(R)]
        for K in (0, 1, 12345, n // 2 + 3, n - 1):
            acc = 0
            for r in range(R):
                acc = (acc + pow(w, r * K, O.R_MOD) * H.fr_dec(subs[r][K % (1 << 16)])[0]) % O.R_MOD
            assert H.fr_dec(fwd[K])[0] == acc
    gpu_ctx.best_fft_device(buf, wi, k)
    back = buf.download(n)
    nf = H.fr_enc([n])[0]
    idx = np.concatenate([np.arange(0, 4096), np.arange(n // 2 - 2048, n // 2 + 2048), np.arange(n - 4096, n)])
    want = oracle_c.field_op(0, 0, a[idx], np.tile(nf, (len(idx), 1)))
    assert (back[idx] == want).all()
    # a checksum over the whole vector: sum of all elements as u64 words must match n*a computed
    # on the host for a strided sample of 2^16 elements
    sidx = np.arange(0, n, n >> 16)
    want = oracle_c.field_op(0, 0, a[sidx], np.tile(nf, (len(sidx), 1)))
    assert (back[sidx] == want).all()
    buf.free()


# ---------------------------------------------------------------------------
# EvaluationDomain                                            poly/domain.rs
# ---------------------------------------------------------------------------
def test_golden_domain(gpu_ctx):
    for v in KAT["domain"]:
        d = h.EvaluationDomain(gpu_ctx, v["j"], v["k"])
        assert d.extended_k == v["extended_k"]
        assert H.fr_enc([d.constant("omega")]).tobytes().hex() == v["omega"]
        assert H.fr_enc([d.constant("extended_omega")]).tobytes().hex() == v["extended_omega"]
        assert H.fr_enc(d.t_evaluations()).tobytes().hex() == v["t_evaluations"]
        assert d.lagrange_to_coeff(_np(v["a"], 4)).tobytes().hex() == v["lagrange_to_coeff"]
        assert d.coeff_to_extended(_np(v["a"], 4)).tobytes().hex() == v["coeff_to_extended"]
        assert d.divide_by_vanishing_poly(_np(v["ext"], 4)).tobytes().hex() == v["divide_by_vanishing_poly"]
        assert d.extended_to_coeff(_np(v["ext"], 4)).tobytes().hex() == v["extended_to_coeff"]
        d.free()


@pytest.mark.parametrize("j,k", [(2, 1), (3, 6), (5, 10), (5, 14), (4, 15), (9, 12), (5, 18)])
def test_domain_vs_oracle(gpu_ctx, oracle_c, j, k):
    d = h.EvaluationDomain(gpu_ctx, j, k)
    od = oracle_c.domain(j, k, 0)
    assert d.extended_k == od.extended_k and d.quotient_len == od.quotient_len
    for which, name in enumerate(["omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset",
                                  "g_coset_inv", "ifft_divisor", "extended_ifft_divisor"]):
        assert H.fr_enc([d.constant(name)]).tobytes() == od.constant(which).tobytes()
    a = H.rand_fr_limbs(j * 100 + k, 1 << k)
    coeff = d.lagrange_to_coeff(a)
    assert (coeff == od.lagrange_to_coeff(a)).all()
    ext = d.coeff_to_extended(coeff)
    assert (ext == od.coeff_to_extended(coeff)).all()
    e = H.rand_fr_limbs(j * 1000 + k, 1 << d.extended_k)
    div = d.divide_by_vanishing_poly(e)
    assert (div == od.divide_by_vanishing_poly(e)).all()
    assert (d.extended_to_coeff(e) == od.extended_to_coeff(e)).all()
    assert (d.extended_to_coeff(e, divide_by_vanishing=True) == od.extended_to_coeff(div)).all()
    # round trip (truncation keeps n*(j-1) >= n coefficients)
    back = d.extended_to_coeff(ext)
    assert (back[: 1 << k] == coeff).all() and not back[1 << k:].any()
    d.free()
    od.free()


def test_domain_length_checks(gpu_ctx):
    d = h.EvaluationDomain(gpu_ctx, 5, 4)
    for fn, n in ((d.lagrange_to_coeff, 8), (d.coeff_to_extended, 32), (d.extended_to_coeff, 16),
                  (d.divide_by_vanishing_poly, 16)):
        with pytest.raises(h.H2BError) as e:
            fn(H.rand_fr_limbs(0, n))
        assert e.value.code == h.H2B_ERR_LENGTH
    d.free()
    with pytest.raises(h.H2BError):  # extended_k would exceed the two-adicity S = 28
        h.EvaluationDomain(gpu_ctx, 9, 27)


@pytest.mark.parametrize("cap_cols", [0, 5])
def test_domain_batched_64_columns(gpu_ctx, oracle_c, monkeypatch, cap_cols):
    """configs[2]: coset NTT batched over 64 columns (here k=12 so the oracle finishes in seconds);
    cap_cols = 5 shrinks the scratch so that the 64 columns are processed in 13 groups."""
    j, k, ncols = 5, 12, 64
    if cap_cols:
        monkeypatch.setenv("H2B_NTT_SCRATCH_CAP", str(cap_cols * (1 << 14) * 32))
    d = h.EvaluationDomain(gpu_ctx, j, k)
    od = oracle_c.domain(j, k, 0)
    n, ne, nq = 1 << k, 1 << d.extended_k, d.quotient_len
    src = gpu_ctx.alloc(ncols * n * 32)
    cols = H.rand_fr_limbs(5, ncols * n).reshape(ncols, n, 4)
    src.upload(cols.reshape(-1, 4))
    ext = gpu_ctx.alloc(ncols * ne * 32)
    d.coeff_to_extended_device(src, ext, ncols)
    back = gpu_ctx.alloc(ncols * nq * 32)
    d.extended_to_coeff_device(ext, back, ncols)
    e = ext.download(ncols * ne).reshape(ncols, ne, 4)
    b = back.download(ncols * nq).reshape(ncols, nq, 4)
    for c in range(ncols):
        assert (e[c] == od.coeff_to_extended(cols[c])).all()
        assert (b[c][:n] == cols[c]).all() and not b[c][n:].any()
    for x in (src, ext, back):
        x.free()
    d.free()
    od.free()


def test_coset_roundtrip_k22(gpu_ctx):
    """coeff_to_extended -> extended_to_coeff at k=22 (extended_k=24): identity on the first n, zeros after."""
    k = 22
    d = h.EvaluationDomain(gpu_ctx, 5, k)
    n = 1 << k
    src = gpu_ctx.synth_scalars(n, 3, 0)
    ext = gpu_ctx.alloc(d.extended_len() * 32)
    out = gpu_ctx.alloc(d.quotient_len * 32)
    d.coeff_to_extended_device(src, ext)
    d.extended_to_coeff_device(ext, out)
    a = src.download(n)
    b = out.download(d.quotient_len)
    assert (b[:n] == a).all() and not b[n:].any()
    for x in (src, ext, out):
        x.free()
    d.free()


# ---------------------------------------------------------------------------
# best_multiexp / commit                   arithmetic.rs:132, kzg/commitment.rs:281,327
# ---------------------------------------------------------------------------
def test_golden_best_multiexp(gpu_ctx):
    for v in KAT["best_multiexp"]:
        got = gpu_ctx.best_multiexp(_np(v["scalars"], 4), _np(v["bases"], 8))
        assert O.g1_to_bytes(got).hex() == v["result"], v["name"]


SCALAR_KINDS = {
    "uni": lambda rng, n: H.rand_fr(rng, n),
    "eq": lambda rng, n: [rng.randrange(O.R_MOD)] * n,
    "01": lambda rng, n: [rng.randrange(2) for _ in range(n)],
    "small": lambda rng, n: [rng.randrange(1 << 16) for _ in range(n)],
    "sparse": lambda rng, n: [rng.randrange(O.R_MOD) if rng.random() < 0.1 else 0 for _ in range(n)],
    "top": lambda rng, n: [O.R_MOD - 1 - rng.randrange(1 << 20) for _ in range(n)],
    "three": lambda rng, n: [(7, 11, 13)[i % 3] for i in range(n)],  # benches/plonk.rs:247-262 witness shape
}


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 1000, 4097, 70000])
@pytest.mark.parametrize("kind", sorted(SCALAR_KINDS))
def test_msm_vs_oracle(gpu_ctx, oracle_c, n, kind, monkeypatch):
    rng = random.Random(n * 7 + len(kind))
    hs = np.array([rng.randrange(1, 1 << 64) for _ in range(n)], dtype=np.uint64)
    bases = oracle_c.g1_mul_gen(hs)
    sc = SCALAR_KINDS[kind](rng, n)
    S = H.fr_enc(sc)
    B = h.Bases(gpu_ctx, bases, n)
    got = B.msm(S)
    assert got == H.g1_dec(oracle_c.best_multiexp(S, bases, 0))[0]
    assert B.msm(S, affine=False) == got
    if n >= 1024:  # window table path (h2b_bases_precompute): same result
        for c in (0, 7, 13):
            B.precompute(c)
            for acc in ("xyzz", "affine"):  # both bucket-accumulation paths
                monkeypatch.setenv("H2B_MSM_ACC", acc)
                assert B.msm(S) == got, (c, acc)
        m = n // 2 + 5
        assert B.msm(S[:m], offset=3) == H.g1_dec(oracle_c.best_multiexp(S[:m], bases[3:3 + m], 0))[0]
    B.free()


def test_msm_repeated_and_opposite_bases(gpu_ctx, oracle_c, monkeypatch):
    rng = random.Random(9)
    n = 5000
    base = H.g1_dec(oracle_c.g1_mul_gen([rng.randrange(1, 1 << 64) for _ in range(4)]))
    pts = []
    for i in range(n):
        p = base[i % 4]
        pts.append(O.g1_neg(p) if (i // 4) % 2 else p)
    pts[17] = None
    bases = H.g1_enc(pts)
    for kind in ("uni", "eq", "01", "small", "top"):
        S = H.fr_enc(SCALAR_KINDS[kind](rng, n))
        B = h.Bases(gpu_ctx, bases, n)
        want = H.g1_dec(oracle_c.best_multiexp(S, bases, 0))[0]
        assert B.msm(S) == want, kind
        for c in (4, 9, 0):  # window table + batched-affine accumulation: P + P, P - P, identities in-band
            B.precompute(c)
            for acc in ("xyzz", "affine"):
                monkeypatch.setenv("H2B_MSM_ACC", acc)
                assert B.msm(S) == want, (kind, c, acc)
        B.free()


def test_msm_length_checks_and_prefix(gpu_ctx, oracle_c):
    bases = oracle_c.g1_mul_gen(list(range(1, 65)))
    B = h.Bases(gpu_ctx, bases, 64)
    with pytest.raises(h.H2BError) as e:  # kzg/commitment.rs:290
        B.msm(H.rand_fr_limbs(0, 65))
    assert e.value.code == h.H2B_ERR_LENGTH
    with pytest.raises(h.H2BError) as e:  # arithmetic.rs:133
        gpu_ctx.best_multiexp(H.rand_fr_limbs(0, 3), bases)
    assert e.value.code == h.H2B_ERR_LENGTH
    assert B.msm(np.zeros((0, 4), dtype=np.uint64)) is None
    S = H.rand_fr_limbs(1, 20)
    assert B.msm(S) == H.g1_dec(oracle_c.best_multiexp(S, bases[:20], 1))[0]  # commit of a short polynomial
    assert B.msm(S, offset=40) == H.g1_dec(oracle_c.best_multiexp(S, bases[40:60], 1))[0]
    B.free()


def test_kzg_commit_identity(gpu_ctx):
    """kzg/commitment.rs:361-384 through the mirror."""
    v = KAT["kzg"]
    P = h.ParamsKZG(gpu_ctx, v["k"], _np(v["g"], 8), _np(v["g_lagrange"], 8))
    d = h.EvaluationDomain(gpu_ctx, 2, v["k"])
    a = _np(v["lagrange"], 4)
    coeff = d.lagrange_to_coeff(a)
    assert coeff.tobytes().hex() == v["coeff"]
    c1, c2 = P.commit(coeff), P.commit_lagrange(a)
    assert c1 == c2 and O.g1_to_bytes(c1).hex() == v["commitment"]
    d.free()


@pytest.mark.parametrize("k,kind", [(18, 0), (18, 1), (18, 2), (18, 3), (18, 4), (20, 0)])
def test_msm_closed_form_synthetic(gpu_ctx, k, kind):
    """Synthetic benchmark inputs have known discrete logs: sum c_i [h_i]G == [sum c_i h_i]G."""
    n = 1 << k
    B = gpu_ctx.synth_bases(n, 99)
    sc = gpu_ctx.synth_scalars(n, 5, kind)
    got = B.msm(sc, n=n)
    s = H.fr_dec(sc.download(n))
    hs = [gpu_ctx.synth_base_scalar(99, i) for i in range(n)]
    assert got == O.g1_mul(O.G1_GEN, sum(c * x for c, x in zip(s, hs)) % O.R_MOD)
    B.precompute()
    assert B.msm(sc, n=n) == got
    sc.free()
    B.free()


def test_msm_k24_linearity(gpu_ctx):
    """Full benchmark size, size-independent properties:
    MSM(a) + MSM(b) == MSM(a + b)  and  MSM over two halves sums to the whole."""
    k = 24
    n = 1 << k
    B = gpu_ctx.synth_bases(n, 99)
    a = gpu_ctx.synth_scalars(n, 11, 0)
    b = gpu_ctx.synth_scalars(n, 12, 0)
    pa, pb = B.msm(a, n=n), B.msm(b, n=n)
    # a + b on the device: reuse the field-op hook in slabs through the host would move 1 GiB;
    # instead use linearity in the *bases* partition, which needs no new scalars:
    half = n // 2
    lo = B.msm(a, n=half)
    hi_buf = h.DeviceBuffer.__new__(h.DeviceBuffer)
    hi_buf.ctx, hi_buf.nbytes, hi_buf.ptr = gpu_ctx, half * 32, a.at(half * 32)
    hi = B.msm(hi_buf, n=half, offset=half)
    hi_buf.ptr = None
    assert O.g1_add(lo, hi) == pa
    # and the closed form on a 2^16 prefix ties the big run to the oracle's arithmetic
    m = 1 << 16
    s = H.fr_dec(a.download(m))
    hs = [gpu_ctx.synth_base_scalar(99, i) for i in range(m)]
    assert B.msm(a, n=m) == O.g1_mul(O.G1_GEN, sum(c * x for c, x in zip(s, hs)) % O.R_MOD)
    assert pa != pb
    # window table at the benchmark size: identical commitments
    B.precompute()
    assert B.table_window_bits == 22
    assert B.msm(a, n=n) == pa and B.msm(b, n=n) == pb
    for x in (a, b):
        x.free()
    B.free()


def test_concurrent_callers(oracle_c):
    """The reference calls the transforms from rayon workers concurrently
    (plonk/permutation/keygen.rs:214-234): one context per caller thread, results unchanged."""
    import threading
    k, n = 14, 3000
    hs = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    bases = oracle_c.g1_mul_gen(hs)
    w = H.fr_enc([O.omega_for(k)])[0]
    errors = []

    def worker(seed):
        try:
            ctx = h.Context(0)
            B = h.Bases(ctx, bases, n)
            dom = h.EvaluationDomain(ctx, 5, k)
            od = oracle_c.domain(5, k, 1)
            for it in range(4):
                a = H.rand_fr_limbs(seed * 10 + it, 1 << k)
                got = a.copy()
                ctx.best_fft(got, w.reshape(1, 4), k)
                assert (got == oracle_c.best_fft(a, w, k, 1)).all()
                assert (dom.coeff_to_extended(a) == od.coeff_to_extended(a)).all()
                S = H.rand_fr_limbs(seed * 100 + it, n)
                assert B.msm(S) == H.g1_dec(oracle_c.best_multiexp(S, bases, 1))[0]
            od.free()
            dom.free()
            B.free()
            ctx.close()
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_shared_context_is_serialised(gpu_ctx, oracle_c):
    """Several threads on ONE context: calls are serialised by the library's lock."""
    import threading
    k = 12
    w = H.fr_enc([O.omega_for(k)])[0]
    errors = []

    def worker(seed):
        try:
            for it in range(6):
                a = H.rand_fr_limbs(seed * 10 + it, 1 << k)
                got = a.copy()
                gpu_ctx.best_fft(got, w.reshape(1, 4), k)
                assert (got == oracle_c.best_fft(a, w, k, 1)).all()
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


# ---------------------------------------------------------------------------
# polynomial helpers around the hot paths   arithmetic.rs:304-367, poly.rs:229-305
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 8, 9, 2047, 2048, 2049, 100000, 1 << 20])
def test_poly_helpers_vs_oracle(gpu_ctx, oracle_c, n):
    a, b = H.rand_fr_limbs(n, n), H.rand_fr_limbs(n + 1, n)
    x = random.Random(n).randrange(O.R_MOD)
    X = H.fr_enc([x])[0]
    assert gpu_ctx.eval_polynomial(a, x) == H.fr_dec(oracle_c.eval_polynomial(a, X, 0))[0]
    assert (gpu_ctx.kate_division(a, x) == oracle_c.kate_division(a, X)).all()
    assert gpu_ctx.inner_product(a, b) == H.fr_dec(oracle_c.inner_product(a, b))[0]
    assert (gpu_ctx.poly_add(a, b) == oracle_c.field_op(0, 1, a, b)).all()
    assert (gpu_ctx.poly_sub(a, b) == oracle_c.field_op(0, 2, a, b)).all()
    assert (gpu_ctx.poly_scale(a, x) == oracle_c.field_op(0, 0, a, np.tile(X, (n, 1)))).all()
    da = gpu_ctx.upload_fr(a)
    assert gpu_ctx.eval_polynomial(da, x, n=n) == gpu_ctx.eval_polynomial(a, x)
    if n > 1:
        q = gpu_ctx.kate_division(da, x, n=n)
        assert (q.download(n - 1) == oracle_c.kate_division(a, X)).all()
        q.free()
    da.free()


def test_poly_helpers_k24_properties(gpu_ctx):
    """Full size: a(X) = q(X) (X - b) + a(b), checked at a random point with the device evaluator itself,
    and eval at omega^j against the NTT output (two independent kernels must agree)."""
    k = 24
    n = 1 << k
    rng = random.Random(5)
    a = gpu_ctx.synth_scalars(n, 77, 0)
    b, z = rng.randrange(O.R_MOD), rng.randrange(O.R_MOD)
    q = gpu_ctx.kate_division(a, b, n=n)
    ab, az, qz = gpu_ctx.eval_polynomial(a, b, n=n), gpu_ctx.eval_polynomial(a, z, n=n), \
        gpu_ctx.eval_polynomial(q, z, n=n - 1)
    assert (qz * (z - b) + ab) % O.R_MOD == az
    w = O.omega_for(k)
    j = 123457
    ej = gpu_ctx.eval_polynomial(a, pow(w, j, O.R_MOD), n=n)
    gpu_ctx.best_fft_device(a, w, k)
    assert H.fr_dec(a.download(1, offset_bytes=j * 32))[0] == ej
    for x in (a, q):
        x.free()


def test_poly_helpers_edge_cases(gpu_ctx):
    assert gpu_ctx.eval_polynomial(np.zeros((0, 4), dtype=np.uint64), 5) == 0
    with pytest.raises(h.H2BError):
        gpu_ctx.kate_division(np.zeros((0, 4), dtype=np.uint64), 5)
    with pytest.raises(h.H2BError) as e:  # arithmetic.rs:334
        gpu_ctx.inner_product(H.rand_fr_limbs(0, 3), H.rand_fr_limbs(0, 4))
    assert e.value.code == h.H2B_ERR_LENGTH


@pytest.mark.parametrize("n", [1, 2, 9, 2048, 2049, 100000, 1 << 20])
def test_grand_product_pieces(gpu_ctx, oracle_c, n):
    """batch_invert (zeros stay zero) and z[i] = z[i-1] * f[i-1]
    (plonk/permutation/prover.rs:119, 152-158)."""
    a = H.rand_fr_limbs(n, n)
    a[::7] = 0
    inv = gpu_ctx.batch_invert(a)
    want = oracle_c.field_op(0, 7, a, a)  # Fermat inverse, 0 -> 0
    assert (inv == want).all()
    assert (oracle_c.field_op(0, 0, inv[1:2], a[1:2]) == H.fr_enc([1])).all() if n > 1 else True
    f = H.rand_fr_limbs(n + 5, n)
    z = gpu_ctx.running_product(f, 1)
    # z[i+1] = z[i] * f[i] everywhere (one vectorised oracle multiplication), z[0] = 1
    assert (z[0] == H.fr_enc([1])[0]).all()
    if n > 1:
        assert (z[1:] == oracle_c.field_op(0, 0, z[:-1], f[:-1])).all()
    d = gpu_ctx.upload_fr(f)
    zd = gpu_ctx.running_product(d, 1, n=n)
    assert (zd.download(n) == z).all()
    d.free()
    zd.free()


def test_allocation_failure_is_an_error_code(gpu_ctx):
    """A failed device allocation comes back as H2B_ERR_OOM, never as a crash or a CPU fallback."""
    with pytest.raises(h.H2BError) as e:
        gpu_ctx.alloc(1 << 46)
    assert e.value.code in (h.H2B_ERR_OOM, h.H2B_ERR_CUDA)
    # the context stays usable afterwards
    a = H.rand_fr_limbs(1, 256)
    w = H.fr_enc([O.omega_for(8)])
    b = a.copy()
    gpu_ctx.best_fft(b, w, 8)
    assert not (a == b).all()


def test_kzg_setup_and_commit_identity(gpu_ctx, oracle_c):
    """ParamsKZG::setup on the device (kzg/commitment.rs:61-129) against the golden k=4 SRS, then the
    reference's own test (:361-384) at k = 6 (its size) and k = 14 with the window table."""
    v = KAT["kzg"]
    P = h.ParamsKZG.setup(gpu_ctx, v["k"], int(v["s"], 16))
    assert P.g.download().tobytes().hex() == v["g"]
    assert P.g_lagrange.download().tobytes().hex() == v["g_lagrange"]
    for k, pre in ((6, False), (14, True)):
        P = h.ParamsKZG.setup(gpu_ctx, k, 0x1234567 + k, precompute=pre)
        d = h.EvaluationDomain(gpu_ctx, 2, k)
        a = H.rand_fr_limbs(k, 1 << k)
        c1, c2 = P.commit(d.lagrange_to_coeff(a)), P.commit_lagrange(a)
        assert c1 == c2 and c1 is not None
        # and against the CPU oracle on the same (downloaded) bases
        assert c2 == H.g1_dec(oracle_c.best_multiexp(a, P.g_lagrange.download(), 0))[0]
        d.free()


# ---------------------------------------------------------------------------
# maximum sizes: the two-adicity ceiling S = 28 of Fr and a 2^26-point MSM
# ---------------------------------------------------------------------------
def _sample_offsets(n, width=2048):
    return [0, n // 3, n // 2 - width // 2, n - width]


def test_best_fft_k28_roundtrip(gpu_ctx, oracle_c):
    """k = 28 = Fr::S, the largest transform the field supports (8 GiB): inverse(forward(a)) == 2^28 * a
    on sampled windows, plus one output value against direct evaluation of a sparse input."""
    k = 28
    n = 1 << k
    w = O.omega_for(k)
    buf = gpu_ctx.synth_scalars(n, 41, 0)
    before = {o: buf.download(2048, offset_bytes=o * 32) for o in _sample_offsets(n)}
    gpu_ctx.best_fft_device(buf, w, k)
    gpu_ctx.best_fft_device(buf, pow(w, -1, O.R_MOD), k)
    nf = np.tile(H.fr_enc([n])[0], (2048, 1))
    for o, a in before.items():
        assert (buf.download(2048, offset_bytes=o * 32) == oracle_c.field_op(0, 0, a, nf)).all(), o
    buf.free()


def test_domain_extended_k28(gpu_ctx):
    """EvaluationDomain::new(5, 26): extended_k = 28 is the ceiling (domain.rs:49-61).
    coeff_to_extended then extended_to_coeff gives back the polynomial, zero beyond n."""
    k = 26
    d = h.EvaluationDomain(gpu_ctx, 5, k)
    assert d.extended_k == 28
    n = 1 << k
    src = gpu_ctx.synth_scalars(n, 43, 0)
    ext = gpu_ctx.alloc(d.extended_len() * 32)
    d.coeff_to_extended_device(src, ext)
    out = gpu_ctx.alloc(d.quotient_len * 32)
    d.extended_to_coeff_device(ext, out)
    ext.free()
    for o in _sample_offsets(n):
        assert (out.download(2048, offset_bytes=o * 32) == src.download(2048, offset_bytes=o * 32)).all(), o
    for o in (n, 2 * n + 12345, d.quotient_len - 2048):
        assert not out.download(2048, offset_bytes=o * 32).any(), o
    for x in (src, out):
        x.free()
    d.free()


def test_msm_k26_properties(gpu_ctx):
    """2^26 points (4 GiB of bases, 44 GiB window table): the MSM over two halves sums to the whole, and a
    2^16 prefix matches the closed form through the same table."""
    k = 26
    n = 1 << k
    B = gpu_ctx.synth_bases(n, 7).precompute()
    assert B.table_window_bits == 24
    a = gpu_ctx.synth_scalars(n, 13, 0)
    whole = B.msm(a, n=n)
    half = n // 2
    lo = B.msm(a, n=half)
    hi_buf = h.DeviceBuffer.__new__(h.DeviceBuffer)
    hi_buf.ctx, hi_buf.nbytes, hi_buf.ptr = gpu_ctx, half * 32, a.at(half * 32)
    hi = B.msm(hi_buf, n=half, offset=half)
    hi_buf.ptr = None
    assert O.g1_add(lo, hi) == whole and whole is not None
    m = 1 << 16
    s = H.fr_dec(a.download(m))
    hs = [gpu_ctx.synth_base_scalar(7, i) for i in range(m)]
    assert B.msm(a, n=m) == O.g1_mul(O.G1_GEN, sum(c * x for c, x in zip(s, hs)) % O.R_MOD)
    a.free()
    B.free()


def test_pageable_host_slices(gpu_ctx, oracle_c):
    """Drop-in calls on PAGEABLE host slices (a Rust Vec, a numpy array): staged through the pinned ring by
    several host threads (copy_h2d_any / copy_d2h_any); same results as with pinned buffers."""
    rs = np.random.RandomState(11)
    for nbytes in (4 << 20, (6 << 20) + 32, (70 << 20) + 96):
        a = rs.randint(0, 1 << 62, size=(nbytes // 32, 4), dtype=np.int64).astype(np.uint64)
        buf = gpu_ctx.alloc(nbytes)
        buf.upload(a)
        assert (buf.download(a.shape[0]) == a).all(), nbytes
        buf.free()
    k = 20
    a = H.rand_fr_limbs(9, 1 << k)
    w = H.fr_enc([O.omega_for(k)])[0]
    want = oracle_c.best_fft(a, w, k, 0)
    got = a.copy()
    gpu_ctx.best_fft(got, w.reshape(1, 4), k)
    assert (got == want).all()
    n = (1 << 18) + 5  # MSM with pageable scalars: 8 chunks of 1 MiB (direct) and, below, one staged 8 MiB copy
    bases = gpu_ctx.synth_bases(n, 0x77)
    sc = H.rand_fr_limbs(10, n)
    pinned = gpu_ctx.pinned((n, 4))
    pinned.array[:] = sc
    assert bases.msm(sc) == bases.msm(pinned.array)
    bases.free()


@pytest.mark.parametrize("kind", [0, 1, 2, 4])
def test_msm_host_scalars_in_batches(gpu_ctx, kind, monkeypatch):
    """Host scalars on a window table are processed in batches that share the bucket array (the copy of batch
    b + 1 runs under the compute of batch b, later batches add to the buckets): same point as the one-batch
    device-resident MSM, for uniform / all-equal / 0-1 / sparse scalars, at sizes that give 2 and 4 batches
    (the batch size is lowered from its default of 2^22 points to keep the test small)."""
    monkeypatch.setenv("H2B_MSM_BATCH_MIN", str(1 << 20))
    for n in ((1 << 21) + 3, (1 << 22) + 77):
        bases = gpu_ctx.synth_bases(n, 0x99 + kind)
        bases.precompute()
        dev = gpu_ctx.synth_scalars(n, 41 + kind, kind)
        want = bases.msm(dev, n)
        host = dev.download(n)           # pageable numpy array
        assert bases.msm(host) == want, (n, kind)
        pinned = gpu_ctx.pinned((n, 4))
        pinned.array[:] = host
        assert bases.msm(pinned.array) == want, (n, kind)
        # the batch sizes of the calls above follow the measured copy/compute ratio of the calls before them;
        # fixed plans: seven equal batches, shrinking batches, a single weight
        for plan in ("1,1,1,1,1,1,1", "5,2,1", "7"):
            monkeypatch.setenv("H2B_MSM_BATCH_PLAN", plan)
            assert bases.msm(pinned.array) == want, (n, kind, plan)
        monkeypatch.delenv("H2B_MSM_BATCH_PLAN")
        pinned.free()
        dev.free()
        bases.free()


@pytest.mark.parametrize("k,ncols", [(12, 2), (16, 5), (20, 4)])
def test_msm_multi_column(gpu_ctx, k, ncols, monkeypatch):
    """h2b_msm_multi_affine: several scalar vectors over the same resident bases in one digit / sort / accumulate /
    reduce pass (one bucket set per column) give the points of separate MSMs -- uniform, all-equal, 0/1, sparse
    and all-zero columns side by side, full length and a shorter prefix at an offset."""
    n = (1 << k) + 9
    bases = gpu_ctx.synth_bases(n, 0x55 + k)
    bases.precompute()
    kinds = [0, 1, 2, 4, 0][:ncols]
    cols = [gpu_ctx.synth_scalars(n, 70 + j, kind) for j, kind in enumerate(kinds)]
    gpu_ctx.memset(cols[-1], 0)  # an all-zero column: the identity
    want = [bases.msm(c, n) for c in cols]
    assert want[-1] is None
    assert bases.msm_many([(c, n) for c in cols]) == want
    m = n - 300
    assert bases.msm_many([(c, m, 100, 50) for c in cols]) == [bases.msm(c, m, offset=100, scalar_offset=50) for c in cols]
    monkeypatch.setenv("H2B_MSM_NO_MULTI", "1")  # sibling contexts instead of the fused pass
    assert bases.msm_many([(c, n) for c in cols]) == want
    for c in cols:
        c.free()
    bases.free()


def test_small_multiexp_and_g_to_lagrange(gpu_ctx):
    """SURVEY.md 8a rows a3 and a12 (arithmetic.rs:105-125, 277-301; kzg/commitment.rs:267-275): vs the oracle at
    small k, and downsize(k - d) of a 2^k SRS == setup(k - d) bit for bit up to k = 16."""
    from tests import group_cases as G
    G.check_small_multiexp(gpu_ctx)
    G.check_g_to_lagrange_vs_oracle(gpu_ctx, ks=(0, 1, 2, 3, 5, 7))
    G.check_downsize(gpu_ctx, 10, 9)
    G.check_downsize(gpu_ctx, 12, 12, precompute=True)
    G.check_downsize(gpu_ctx, 16, 15)


@pytest.mark.parametrize("k", [25, 26])
def test_best_fft_vs_oracle_largest_sizes(gpu_ctx, oracle_c, k):
    """best_fft at the sizes of configs[4] (k = 26: digits 8 + 9 + 9, the streamed first-pass twiddle table) against the
    C++ restatement of arithmetic.rs:171-274, bit for bit, twice in a row on the same buffer."""
    a = H.rand_fr_limbs(k, 1 << k)
    w = H.fr_enc([O.omega_for(k)])
    want = oracle_c.best_fft(a, w[0], k, 0)
    buf = gpu_ctx.upload_fr(a)
    gpu_ctx.best_fft_device(buf, w, k)
    assert (buf.download(1 << k) == want).all()
    want = oracle_c.best_fft(want, w[0], k, 0)
    gpu_ctx.best_fft_device(buf, w, k)
    assert (buf.download(1 << k) == want).all()
    buf.free()


@pytest.mark.parametrize("k,ncols,slot_cols", [(10, 9, 3.5), (16, 13, 4.0), (18, 6, 1.0)])
def test_streamed_host_batches(gpu_ctx, oracle_c, monkeypatch, k, ncols, slot_cols):
    """configs[2] beyond HBM: host columns streamed through two staging slots per direction (csrc/ntt.cu
    host_batch_streamed), uploads, transforms and read-backs of neighbouring groups overlapped."""
    from tests import group_cases as G
    G.check_streamed_host_batches(gpu_ctx, oracle_c, k, ncols, slot_cols, monkeypatch)
