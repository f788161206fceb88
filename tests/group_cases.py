"""Shared bodies of the a3 / a12 tests: small_multiexp (arithmetic.rs:105-125),
g_to_lagrange (arithmetic.rs:277-301) and ParamsKZG::downsize (poly/kzg/commitment.rs:267-275),
run against the emulator on the CPU and against libhalo2b200.so on the GPU."""
import random

import numpy as np

import halo2_pse_b200 as h
from oracle import bn256 as O
from tests import helpers as H


def check_small_multiexp(ctx):
    rng = random.Random(5)
    for n in (0, 1, 3, 5, 12):
        cs = [rng.randrange(O.R_MOD) for _ in range(n)]
        pts = [O.g1_mul(O.G1_GEN, rng.randrange(1, 1 << 64)) for _ in range(n)]
        if n >= 3:
            cs[1], cs[2] = 0, O.R_MOD - 1  # zero and the largest scalar
        if n >= 5:
            pts[4], pts[0] = pts[3], None  # a repeated base and the identity
        want = O.small_multiexp(cs, pts)
        assert want == O.msm_naive(cs, pts)  # the restatement against the definition
        c = H.fr_enc(cs) if n else np.zeros((0, 4), dtype=np.uint64)
        b = H.g1_enc(pts) if n else np.zeros((0, 8), dtype=np.uint64)
        assert ctx.small_multiexp(c, b) == want, n
        if n:
            assert ctx.best_multiexp(c, b) == want, n  # and the two entry points agree
    try:  # more coefficients than bases: the reference indexes out of bounds (:117)
        ctx.small_multiexp(H.fr_enc([1, 2]), H.g1_enc([O.G1_GEN]))
    except h.H2BError:
        pass
    else:
        raise AssertionError("small_multiexp accepted more coefficients than bases")


def check_g_to_lagrange_vs_oracle(ctx, ks=(0, 1, 2, 3, 5)):
    rng = random.Random(11)
    for k in ks:
        P = O.ParamsKZG.setup(k, rng.randrange(O.R_MOD))
        want = O.g_to_lagrange(P.g, k)
        assert want == P.g_lagrange, k  # restatement: the Lagrange SRS of setup() is the iFFT of the monomial one
        assert H.g1_dec(ctx.g_to_lagrange(H.g1_enc(P.g), k)) == want, k
    # exceptional inputs of the butterflies: identity points, P + P, P - P
    g = [None, O.G1_GEN, O.G1_GEN, O.g1_neg(O.G1_GEN)] + [O.g1_mul(O.G1_GEN, i + 2) for i in range(4)]
    assert H.g1_dec(ctx.g_to_lagrange(H.g1_enc(g), 3)) == O.g_to_lagrange(g, 3)
    assert H.g1_dec(ctx.g_to_lagrange(H.g1_enc([None] * 4), 2)) == [None] * 4
    try:
        ctx.g_to_lagrange(H.g1_enc(g), 2)  # best_fft: assert_eq!(a.len(), 1 << log_n), :184
    except h.H2BError:
        pass
    else:
        raise AssertionError("length mismatch accepted")


def check_downsize(ctx, k_from: int, k_to: int, s: int = 0x1234567, precompute: bool = False):
    """downsize(k_to) of a 2^k_from SRS == setup(k_to) with the same secret, bit for bit (both vectors)."""
    big = h.ParamsKZG.setup(ctx, k_from, s, precompute=precompute)
    small = h.ParamsKZG.setup(ctx, k_to, s)
    big.downsize(k_to)
    assert (big.k, big.n) == (k_to, 1 << k_to)
    assert (big.g.download() == small.g.download()).all()
    assert (big.g_lagrange.download() == small.g_lagrange.download()).all()
    # the reference's commit identity on the downsized parameters (kzg/commitment.rs:361-384)
    d = h.EvaluationDomain(ctx, 1, k_to)
    a = H.rand_fr_limbs(k_to, 1 << k_to)
    assert big.commit(d.lagrange_to_coeff(a)) == big.commit_lagrange(a)
    try:
        big.downsize(k_to + 1)  # assert!(k <= self.k), :268
    except h.H2BError:
        pass
    else:
        raise AssertionError("downsize to a larger k accepted")
    d.free()
    for p in (big, small):
        p.g.free()
        p.g_lagrange.free()


def check_streamed_host_batches(ctx, oc, k: int, ncols: int, slot_cols: float, monkeypatch):
    """The *_batch entry points on HOST columns that exceed one staging slot: column groups, double-buffered
    (H2B_STREAM_SLOT_BYTES sets the slot; ragged last group, padded strides).  Every column equals the oracle's
    single-column transform (poly/domain.rs:226-303, arithmetic.rs:171)."""
    import ctypes as C
    d = h.EvaluationDomain(ctx, 5, k)
    od = oc.domain(5, k, 0)
    n, ne, nq = 1 << k, d.extended_len(), d.quotient_len
    monkeypatch.setenv("H2B_STREAM_SLOT_BYTES", str(int(slot_cols * ne * 32)))
    P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    sin, sout = n + 3, ne + 5
    cols = H.rand_fr_limbs(k * 7 + ncols, ncols * sin)
    out = np.zeros((ncols * sout, 4), dtype=np.uint64)
    ctx._check(ctx.lib.h2b_coeff_to_extended_batch(d.h, P(cols), sin, P(out), sout, h.H2B_HOST, ncols))
    for c in range(ncols):
        assert (out[c * sout:c * sout + ne] == od.coeff_to_extended(cols[c * sin:c * sin + n])).all(), ("c2e", c)
        assert not out[c * sout + ne:(c + 1) * sout].any()  # the padding between columns is not touched
    # extended_to_coeff (with the fused vanishing division), host in -> host out, in place on the same array
    q = np.zeros((ncols * nq, 4), dtype=np.uint64)
    ctx._check(ctx.lib.h2b_extended_to_coeff_batch(d.h, P(out), sout, P(q), nq, h.H2B_HOST, ncols, 1))
    for c in range(ncols):
        want = od.extended_to_coeff(od.divide_by_vanishing_poly(out[c * sout:c * sout + ne]))
        assert (q[c * nq:(c + 1) * nq] == want).all(), ("e2c", c)
    # lagrange_to_coeff and best_fft, in place
    a = cols.copy()
    ctx._check(ctx.lib.h2b_lagrange_to_coeff_batch(d.h, P(a), h.H2B_HOST, ncols, sin))
    w = H.fr_enc([O.omega_for(k)])
    b = cols.copy()
    ctx._check(ctx.lib.h2b_best_fft_batch(ctx.h, P(b), h.H2B_HOST, P(w), k, ncols, sin))
    for c in range(ncols):
        assert (a[c * sin:c * sin + n] == od.lagrange_to_coeff(cols[c * sin:c * sin + n])).all(), ("l2c", c)
        assert (b[c * sin:c * sin + n] == oc.best_fft(cols[c * sin:c * sin + n], w[0], k, 0)).all(), ("fft", c)
        assert (a[c * sin + n:(c + 1) * sin] == cols[c * sin + n:(c + 1) * sin]).all()
    d.free()
    od.free()
