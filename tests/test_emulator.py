"""The product's .cu sources, compiled as C++ on the CPU fiber emulator
(tests/emu/cuda_runtime.h, TEST INFRASTRUCTURE), checked against the oracle
through the same C ABI and Python mirror the GPU tests use.  This validates the
kernels' index arithmetic, tiling, carry-chain host twins and the host
orchestration without a GPU; it is not a product code path.
"""
import random

import numpy as np
import pytest

import halo2_pse_b200 as h
from oracle import bn256 as O
from tests import helpers as H

KAT = H.load_golden("kat_bn256.json")["vectors"]


def _np(hexs, width):
    return np.frombuffer(bytes.fromhex(hexs), dtype=np.uint64).reshape(-1, width).copy()


def test_host_field_ops_match_big_integers(emu_ctx):
    rng = random.Random(1)
    for field, mod in ((0, O.R_MOD), (1, O.Q_MOD)):
        a = [rng.randrange(mod) for _ in range(200)] + [0, 1, mod - 1, mod - 1]
        b = [rng.randrange(mod) for _ in range(200)] + [mod - 1, 0, mod - 1, 1]
        A, B = H.to_limbs(a, mod), H.to_limbs(b, mod)
        for op, f in ((0, lambda x, y: x * y % mod), (1, lambda x, y: (x + y) % mod),
                      (2, lambda x, y: (x - y) % mod), (3, lambda x, y: x * x % mod),
                      (6, lambda x, y: -x % mod),
                      (8, lambda x, y: (x * y - (x + y) * (x - y)) % mod),   # mul_sub: two products, one reduction
                      (9, lambda x, y: x * y % mod)):                         # mul_shoup: y as a fixed multiplier
            out = np.zeros_like(A)
            assert emu_ctx.lib.h2b_host_field_op(field, op, A.ctypes.data, B.ctypes.data, out.ctypes.data, len(a)) == 0
            assert H.from_limbs(out, mod) == [f(x, y) for x, y in zip(a, b)], (field, op)
            out2 = np.zeros_like(A)
            emu_ctx._check(emu_ctx.lib.h2b_test_field_op(emu_ctx.h, field, op, A.ctypes.data, B.ctypes.data,
                                                         out2.ctypes.data, len(a)))
            assert (out2 == out).all()


def test_field_product_variants_host(tmp_path):
    """csrc/field.cuh's product variants against the plain Montgomery product, on the host emulation of the same carry
    chains the kernels compile (tests/cpp/test_field_variants.cpp): dedicated squaring, two products on one reduction,
    the Shoup product for fixed multipliers, lazy residues ([0, 2p), and below 4p where a product follows) with their
    stated ranges.  Random operands and edge values in both fields; the GPU suite checks the device paths
    (test_device_field_ops ops 3, 8, 9; every NTT / MSM parity test)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_field_variants")
    subprocess.run(["g++", "-std=c++17", "-O2", "-DH2B_EMU", os.path.join(root, "tests", "cpp", "test_field_variants.cpp"),
                    "-o", exe], check=True, capture_output=True, text=True)
    r = subprocess.run([exe, "60000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Fr: 60000 iterations, 0 mismatches" in r.stdout and "Fq: 60000 iterations, 0 mismatches" in r.stdout


def test_golden_best_fft(emu_ctx):
    for v in KAT["best_fft"]:
        a = _np(v["in"], 4)
        emu_ctx.best_fft(a, _np(v["omega"], 4), v["log_n"])
        assert a.tobytes().hex() == v["out"], v["log_n"]


@pytest.mark.parametrize("k", [2, 5, 8, 9, 10, 12, 13, 14, 17, 18])
def test_best_fft_vs_oracle(emu_ctx, oracle_c, k):
    a = H.rand_fr_limbs(k, 1 << k)
    w = H.fr_enc([O.omega_for(k)])[0]
    want = oracle_c.best_fft(a, w, k)
    got = a.copy()
    emu_ctx.best_fft(got, w.reshape(1, 4), k)
    assert (got == want).all()


def test_best_fft_rejects_bad_input(emu_ctx):
    a = H.rand_fr_limbs(0, 8)
    with pytest.raises(h.H2BError) as e:  # assert_eq!(n, 1 << log_n)   arithmetic.rs:184
        emu_ctx.best_fft(a, O.omega_for(4), 4)
    assert e.value.code == h.H2B_ERR_LENGTH
    with pytest.raises(h.H2BError) as e:  # omega of the wrong order
        emu_ctx.best_fft(a, O.omega_for(4), 3)
    assert e.value.code == h.H2B_ERR_BAD_OMEGA


def test_golden_domain(emu_ctx):
    for v in KAT["domain"]:
        d = h.EvaluationDomain(emu_ctx, v["j"], v["k"])
        assert d.extended_k == v["extended_k"]
        assert H.fr_enc([d.constant("omega")]).tobytes().hex() == v["omega"]
        assert H.fr_enc(d.t_evaluations()).tobytes().hex() == v["t_evaluations"]
        assert d.lagrange_to_coeff(_np(v["a"], 4)).tobytes().hex() == v["lagrange_to_coeff"]
        assert d.coeff_to_extended(_np(v["a"], 4)).tobytes().hex() == v["coeff_to_extended"]
        assert d.divide_by_vanishing_poly(_np(v["ext"], 4)).tobytes().hex() == v["divide_by_vanishing_poly"]
        assert d.extended_to_coeff(_np(v["ext"], 4)).tobytes().hex() == v["extended_to_coeff"]
        fused = d.extended_to_coeff(_np(v["ext"], 4), divide_by_vanishing=True)
        want = d.extended_to_coeff(_np(v["divide_by_vanishing_poly"], 4))
        assert (fused == want).all()
        d.free()


def test_domain_length_checks(emu_ctx):
    d = h.EvaluationDomain(emu_ctx, 5, 4)
    for fn, n in ((d.lagrange_to_coeff, 8), (d.coeff_to_extended, 32), (d.extended_to_coeff, 16),
                  (d.divide_by_vanishing_poly, 16)):
        with pytest.raises(h.H2BError) as e:  # domain.rs:227,244,282,311
            fn(H.rand_fr_limbs(0, n))
        assert e.value.code == h.H2B_ERR_LENGTH
    d.free()


def test_batched_domain_transforms(emu_ctx, oracle_c, monkeypatch):
    monkeypatch.setenv("H2B_NTT_SCRATCH_CAP", str(2 * (1 << 9) * 32))  # two columns per group: 3 columns -> 2 groups
    j, k, ncols = 5, 7, 3
    d = h.EvaluationDomain(emu_ctx, j, k)
    od = oracle_c.domain(j, k, 2)
    n, ne, nq = 1 << k, 1 << d.extended_k, d.quotient_len
    cols = [H.rand_fr_limbs(100 + c, n) for c in range(ncols)]
    stride = n + 5
    buf = emu_ctx.alloc(ncols * stride * 32)
    for c in range(ncols):
        buf.upload(cols[c], c * stride * 32)
    d.lagrange_to_coeff_device(buf, ncols, stride)
    coeffs = [buf.download(n, c * stride * 32) for c in range(ncols)]
    for c in range(ncols):
        assert (coeffs[c] == od.lagrange_to_coeff(cols[c])).all()
    ext = emu_ctx.alloc(ncols * ne * 32)
    d.coeff_to_extended_device(buf, ext, ncols, stride, ne)
    exts = [ext.download(ne, c * ne * 32) for c in range(ncols)]
    for c in range(ncols):
        assert (exts[c] == od.coeff_to_extended(coeffs[c])).all()
    back = emu_ctx.alloc(ncols * nq * 32)
    d.extended_to_coeff_device(ext, back, ncols, ne, nq, divide_by_vanishing=True)
    for c in range(ncols):
        want = od.extended_to_coeff(od.divide_by_vanishing_poly(exts[c]))
        assert (back.download(nq, c * nq * 32) == want).all()
    for b in (buf, ext, back):
        b.free()
    d.free()
    od.free()


def test_golden_best_multiexp(emu_ctx):
    for v in KAT["best_multiexp"]:
        got = emu_ctx.best_multiexp(_np(v["scalars"], 4), _np(v["bases"], 8))
        assert O.g1_to_bytes(got).hex() == v["result"], v["name"]


@pytest.mark.parametrize("n,kind", [(1, "uni"), (2, "uni"), (257, "uni"), (1500, "uni"), (1500, "eq"),
                                    (1500, "01"), (1500, "small"), (700, "sparse")])
def test_msm_vs_oracle(emu_ctx, oracle_c, n, kind, monkeypatch):
    rng = random.Random(n)
    hs = [rng.randrange(1, 1 << 64) for _ in range(n)]
    bases = oracle_c.g1_mul_gen(hs)
    sc = {"uni": lambda: H.rand_fr(rng, n), "eq": lambda: [rng.randrange(O.R_MOD)] * n,
          "01": lambda: [rng.randrange(2) for _ in range(n)],
          "small": lambda: [rng.randrange(1 << 16) for _ in range(n)],
          "sparse": lambda: [rng.randrange(O.R_MOD) if rng.random() < 0.1 else 0 for _ in range(n)]}[kind]()
    S = H.fr_enc(sc)
    B = h.Bases(emu_ctx, bases, n)
    got = B.msm(S)
    assert got == H.g1_dec(oracle_c.best_multiexp(S, bases, 4))[0]
    assert got == O.g1_mul(O.G1_GEN, sum(c * x for c, x in zip(sc, hs)) % O.R_MOD)
    assert B.msm(S, affine=False) == got  # Jacobian output of h2b_msm
    if n > 10:  # prefix / offset forms used by commit on shorter polynomials
        assert B.msm(S[:10], offset=3) == H.g1_dec(oracle_c.best_multiexp(S[:10], bases[3:13], 1))[0]
    if n >= 1024:  # window table: all windows share one bucket set
        for c in (0, 5, 13):  # 13: two-dimensional bucket reduction (>= 2^12 buckets)
            B.precompute(c)
            auto = min(range(8, 25), key=lambda cc: (n * ((255 + cc - 1) // cc) + 2.8 * (1 << (cc - 1)), cc))
            assert B.table_window_bits == (c or auto)
            monkeypatch.setenv("H2B_MSM_ACC", "affine" if c else "xyzz")
            assert B.msm(S) == got
            assert B.msm(S[:1100], offset=200) == H.g1_dec(oracle_c.best_multiexp(S[:1100], bases[200:1300], 2))[0]
        # several columns over the same base slice in ONE pass (h2b_msm_multi_affine: one bucket set per column)
        cols = [S, S[::-1].copy(), np.zeros_like(S), H.fr_enc([1] * n)]
        dev = [emu_ctx.upload_fr(cc) for cc in cols]
        want_cols = [B.msm(cc) for cc in cols]
        assert B.msm_many([(d, n) for d in dev]) == want_cols
        assert B.msm_many([(d, 1100, 200, 7) for d in dev]) == [B.msm(cc[7:1107], offset=200) for cc in cols]
        monkeypatch.setenv("H2B_MSM_NO_MULTI", "1")   # the same jobs dealt to sibling contexts instead
        assert B.msm_many([(d, n) for d in dev]) == want_cols
        monkeypatch.delenv("H2B_MSM_NO_MULTI")
        for d in dev:
            d.free()
        # host scalars in batches that share the bucket array (copy of batch b + 1 under the compute of batch b;
        # later batches ADD to the buckets): 2 batches, then 4, then ragged
        monkeypatch.setenv("H2B_MSM_ACC", "xyzz")
        for batch_min in (700, 300, 375):
            monkeypatch.setenv("H2B_MSM_BATCH_MIN", str(batch_min))
            assert B.msm(S) == got, batch_min
            assert B.msm(S[:1101], offset=200) == H.g1_dec(oracle_c.best_multiexp(S[:1101], bases[200:1301], 2))[0]
        # batch sizes: the default geometric plan above, the plan adapted to the measured copy/compute ratio (every call
        # after the first), and fixed plans
        monkeypatch.setenv("H2B_MSM_BATCH_MIN", "200")
        for plan in (None, "1,1,1,1,1,1,1", "5,2,1", "7"):
            if plan:
                monkeypatch.setenv("H2B_MSM_BATCH_PLAN", plan)
            assert B.msm(S) == got, plan
        monkeypatch.delenv("H2B_MSM_BATCH_PLAN")
        monkeypatch.delenv("H2B_MSM_BATCH_MIN")
    B.free()


def test_msm_exceptional_cases_batched_affine(emu_ctx, oracle_c, monkeypatch):
    """Window-table / batched-affine path with P + P, P + (-P) and identity operands everywhere."""
    rng = random.Random(77)
    n = 1300
    base = H.g1_dec(oracle_c.g1_mul_gen([rng.randrange(1, 1 << 64) for _ in range(3)]))
    pts = []
    for i in range(n):
        q = base[i % 3]
        pts.append(O.g1_neg(q) if (i // 3) % 2 else q)
    for i in (0, 17, 18, 500, n - 1):
        pts[i] = None
    bases = H.g1_enc(pts)
    B = h.Bases(emu_ctx, bases, n).precompute(6)
    monkeypatch.setenv("H2B_MSM_TILE", "1000")  # several round-0 tiles
    for acc, sc in [(a, x) for a in ("affine", "xyzz") for x in ([5] * n, [rng.randrange(2) for _ in range(n)], [rng.randrange(1 << 12) for _ in range(n)],
               H.rand_fr(rng, n), [O.R_MOD - 1] * n)]:
        monkeypatch.setenv("H2B_MSM_ACC", acc)
        S = H.fr_enc(sc)
        assert B.msm(S) == H.g1_dec(oracle_c.best_multiexp(S, bases, 3))[0], acc
    B.free()


def test_msm_length_checks(emu_ctx, oracle_c):
    bases = oracle_c.g1_mul_gen([1, 2, 3, 4])
    B = h.Bases(emu_ctx, bases, 4)
    with pytest.raises(h.H2BError) as e:  # assert!(bases.len() >= size)  kzg/commitment.rs:290
        B.msm(H.rand_fr_limbs(0, 5))
    assert e.value.code == h.H2B_ERR_LENGTH
    with pytest.raises(h.H2BError) as e:  # assert_eq!(coeffs.len(), bases.len())  arithmetic.rs:133
        emu_ctx.best_multiexp(H.rand_fr_limbs(0, 3), bases)
    assert e.value.code == h.H2B_ERR_LENGTH
    assert B.msm(np.zeros((0, 4), dtype=np.uint64)) is None  # empty input -> identity
    B.free()


def test_kzg_commit_identity(emu_ctx):
    """kzg/commitment.rs:361-384 through the mirror: commit(lagrange_to_coeff(a)) == commit_lagrange(a)."""
    v = KAT["kzg"]
    P = h.ParamsKZG(emu_ctx, v["k"], _np(v["g"], 8), _np(v["g_lagrange"], 8))
    d = h.EvaluationDomain(emu_ctx, 2, v["k"])
    a = _np(v["lagrange"], 4)
    coeff = d.lagrange_to_coeff(a)
    assert coeff.tobytes().hex() == v["coeff"]
    c1, c2 = P.commit(coeff), P.commit_lagrange(a)
    assert c1 == c2 and O.g1_to_bytes(c1).hex() == v["commitment"]
    d.free()


@pytest.mark.parametrize("n", [1, 2, 8, 9, 2047, 2048, 2049, 5000, 70000])
def test_poly_helpers(emu_ctx, oracle_c, n):
    """eval_polynomial, kate_division, inner product, + / - / * scalar."""
    a, b = H.rand_fr_limbs(n, n), H.rand_fr_limbs(n + 1, n)
    x = random.Random(n).randrange(O.R_MOD)
    X = H.fr_enc([x])[0]
    assert emu_ctx.eval_polynomial(a, x) == H.fr_dec(oracle_c.eval_polynomial(a, X, 2))[0]
    assert (emu_ctx.kate_division(a, x) == oracle_c.kate_division(a, X)).all()
    assert emu_ctx.inner_product(a, b) == H.fr_dec(oracle_c.inner_product(a, b))[0]
    assert (emu_ctx.poly_add(a, b) == oracle_c.field_op(0, 1, a, b)).all()
    assert (emu_ctx.poly_sub(a, b) == oracle_c.field_op(0, 2, a, b)).all()
    assert (emu_ctx.poly_scale(a, x) == oracle_c.field_op(0, 0, a, np.tile(X, (n, 1)))).all()
    # device-resident forms
    da, db = emu_ctx.upload_fr(a), emu_ctx.upload_fr(b)
    assert emu_ctx.eval_polynomial(da, x, n=n) == emu_ctx.eval_polynomial(a, x)
    if n > 1:
        q = emu_ctx.kate_division(da, x, n=n)
        assert (q.download(n - 1) == oracle_c.kate_division(a, X)).all()
        q.free()
    emu_ctx.poly_add(da, db, n=n)
    assert (da.download(n) == oracle_c.field_op(0, 1, a, b)).all()
    da.free()
    db.free()


def test_poly_helpers_edge_cases(emu_ctx):
    assert emu_ctx.eval_polynomial(np.zeros((0, 4), dtype=np.uint64), 5) == 0
    with pytest.raises(h.H2BError):
        emu_ctx.kate_division(np.zeros((0, 4), dtype=np.uint64), 5)
    with pytest.raises(h.H2BError) as e:  # assert_eq!(a.len(), b.len())  arithmetic.rs:334
        emu_ctx.inner_product(H.rand_fr_limbs(0, 3), H.rand_fr_limbs(0, 4))
    assert e.value.code == h.H2B_ERR_LENGTH
    one = H.fr_enc([7])
    assert emu_ctx.eval_polynomial(one, 12345) == 7
    assert emu_ctx.kate_division(one, 3).shape == (0, 4)


@pytest.mark.parametrize("n", [1, 2, 9, 2048, 2049, 5000, 40000])
def test_grand_product_pieces(emu_ctx, oracle_c, n):
    """batch_invert and the running product of the permutation argument."""
    rng = random.Random(n)
    vals = H.rand_fr(rng, n)
    for i in range(0, n, 7):
        vals[i] = 0  # zeros stay zero (ff::BatchInvert skips them)
    a = H.fr_enc(vals)
    inv = emu_ctx.batch_invert(a)
    assert H.fr_dec(inv) == [pow(v, -1, O.R_MOD) if v else 0 for v in vals]
    f = H.rand_fr(rng, n)
    init = rng.randrange(O.R_MOD)
    z = emu_ctx.running_product(H.fr_enc(f), init)
    want, cur = [], init
    for i in range(n):
        want.append(cur)
        cur = cur * f[i] % O.R_MOD
    assert H.fr_dec(z) == want
    d = emu_ctx.upload_fr(H.fr_enc(f))
    zd = emu_ctx.running_product(d, init, n=n)
    assert (zd.download(n) == z).all()
    d.free()
    zd.free()


def test_kzg_setup_on_device(emu_ctx):
    """ParamsKZG::setup (kzg/commitment.rs:61-129): the golden k=4 SRS, and the reference's own test
    identity commit(lagrange_to_coeff(a)) == commit_lagrange(a) (:361-384) on it."""
    v = KAT["kzg"]
    P = h.ParamsKZG.setup(emu_ctx, v["k"], int(v["s"], 16))
    assert P.g.download().tobytes().hex() == v["g"]
    assert P.g_lagrange.download().tobytes().hex() == v["g_lagrange"]
    d = h.EvaluationDomain(emu_ctx, 2, v["k"])
    a = _np(v["lagrange"], 4)
    assert P.commit(d.lagrange_to_coeff(a)) == P.commit_lagrange(a)
    assert O.g1_to_bytes(P.commit_lagrange(a)).hex() == v["commitment"]
    d.free()
    # scalar edge cases of the fixed-base multiplication: 0, 1, r - 1
    sc = H.fr_enc([0, 1, 2, O.R_MOD - 1])
    out = np.zeros((4, 8), dtype=np.uint64)
    emu_ctx._check(emu_ctx.lib.h2b_g1_mul_generator(emu_ctx.h, sc.ctypes.data, h.H2B_HOST, 4, out.ctypes.data, h.H2B_HOST))
    assert H.g1_dec(out) == [None, O.G1_GEN, O.g1_mul(O.G1_GEN, 2), O.g1_neg(O.G1_GEN)]


def test_pageable_copies_through_the_pinned_ring(emu_ctx, oracle_c, monkeypatch):
    """copy_h2d_any / copy_d2h_any (csrc/ctx.cu): host slices that are not pinned go through the context's
    pinned ring on several host threads.  The emulator has no pinned memory, so the staged path is opted
    in; sizes cover one short chunk, a ragged tail and more chunks than slots (2 MiB chunks, 16 slots)."""
    monkeypatch.setenv("H2B_EMU_STAGED_COPY", "1")
    monkeypatch.setenv("H2B_COPY_THREADS", "3")
    rs = np.random.RandomState(7)
    for nbytes in (4 << 20, (6 << 20) + 32, (38 << 20) + 96):
        a = rs.randint(0, 1 << 62, size=(nbytes // 32, 4), dtype=np.int64).astype(np.uint64)
        buf = emu_ctx.alloc(nbytes)
        buf.upload(a)
        assert (buf.download(a.shape[0]) == a).all(), nbytes
        buf.free()
    # a drop-in call on host slices above the staging threshold: best_fft at k = 17 (4 MiB in and out)
    k = 17
    a = H.rand_fr_limbs(5, 1 << k)
    w = H.fr_enc([O.omega_for(k)])[0]
    want = oracle_c.best_fft(a, w, k, 0)
    got = a.copy()
    emu_ctx.best_fft(got, w.reshape(1, 4), k)
    assert (got == want).all()


def test_device_block_cache(emu_ctx):
    """h2b_device_alloc / h2b_device_free keep released blocks per size and hand them to the next request of that
    size (the prover asks for the same sizes dozens of times per proof); other sizes and small blocks go to the
    allocator as before."""
    a = emu_ctx.alloc(1 << 20)
    pa = a.ptr.value
    a.free()
    b = emu_ctx.alloc(1 << 20)        # same size: the released block comes back
    assert b.ptr.value == pa
    c = emu_ctx.alloc(1 << 20)        # the cache is empty again: a new block
    assert c.ptr.value != pa
    d = emu_ctx.alloc((1 << 20) + 32)  # another size never takes a cached block of a different size
    b.free()
    e = emu_ctx.alloc((1 << 20) + 32)
    assert e.ptr.value not in (pa, d.ptr.value)
    small = emu_ctx.alloc(64)          # below the caching threshold
    small.upload(np.arange(8, dtype=np.uint64))
    assert (small.download(2).reshape(-1) == np.arange(8, dtype=np.uint64)).all()
    for x in (c, d, e, small):
        x.free()
    # contents written through a recycled block are what the next kernel reads
    v = H.rand_fr_limbs(3, 1 << 15)   # 1 MiB
    f = emu_ctx.alloc(1 << 20)
    f.upload(v)
    assert (f.download(1 << 15) == v).all()
    f.free()


def test_msm_many_forms(emu_ctx, oracle_c):
    """Bases.msm_many: no jobs, one job, jobs of different lengths (dealt to sibling contexts), jobs with their
    own host->device copies (`pre`), host-array jobs."""
    n = 1200
    rng = random.Random(5)
    bases = oracle_c.g1_mul_gen([rng.randrange(1, 1 << 64) for _ in range(n)])
    B = h.Bases(emu_ctx, bases, n)
    B.precompute(9)
    cols = [H.rand_fr_limbs(40 + j, n) for j in range(3)]
    want = [B.msm(c) for c in cols]
    assert B.msm_many([]) == []
    dev = [emu_ctx.upload_fr(c) for c in cols]
    assert B.msm_many([(dev[0], n)]) == want[:1]
    assert B.msm_many([(dev[0], n), (dev[1], 1100), (dev[2], n, 0, 0)]) == [want[0], B.msm(cols[1][:1100]), want[2]]
    assert B.msm_many([(c,) for c in cols]) == want                       # host arrays
    empty = [emu_ctx.alloc(n * 32) for _ in cols]                         # filled by the job's own copy
    pre = [(lambda ctx, b=b, c=c: b.upload(c, ctx=ctx)) for b, c in zip(empty, cols)]
    assert B.msm_many([(b, n) for b in empty], pre=pre) == want
    for b in dev + empty:
        b.free()
    B.free()


def test_small_multiexp_and_g_to_lagrange(emu_ctx):
    """SURVEY.md 8a rows a3 and a12 on the emulator (arithmetic.rs:105-125, 277-301; kzg/commitment.rs:267-275)."""
    from tests import group_cases as G
    G.check_small_multiexp(emu_ctx)
    G.check_g_to_lagrange_vs_oracle(emu_ctx)
    G.check_downsize(emu_ctx, 6, 4)
    G.check_downsize(emu_ctx, 5, 5, precompute=True)


@pytest.mark.parametrize("log_n,ncols", [(6, 32), (6, 64), (7, 16), (8, 24), (9, 4), (9, 12), (9, 5)])
def test_single_pass_batches_on_the_register_kernel(emu_ctx, oracle_c, log_n, ncols, monkeypatch):
    """Batches of 64..512-point transforms run the register kernel with batch members as tile columns (ncols a
    multiple of the tile width) or the generic kernel (otherwise, or H2B_NTT_NO_SB): same results as best_fft per
    column (arithmetic.rs:171), also through the scaled domain transforms (poly/domain.rs:226-303)."""
    n = 1 << log_n
    stride = n + 8
    w = H.fr_enc([O.omega_for(log_n)])
    a = H.rand_fr_limbs(log_n * 100 + ncols, ncols * stride)
    want = a.copy()
    for c in range(ncols):
        want[c * stride:c * stride + n] = oracle_c.best_fft(a[c * stride:c * stride + n], w[0], log_n, 1)
    for no_sb in (False, True):
        if no_sb:
            monkeypatch.setenv("H2B_NTT_NO_SB", "1")
        buf = emu_ctx.upload_fr(a)
        emu_ctx.best_fft_device(buf, w, log_n, ncols, stride)
        assert (buf.download(ncols * stride) == want).all(), no_sb
        buf.free()
    monkeypatch.delenv("H2B_NTT_NO_SB")
    if log_n - 2 >= 4:  # coeff_to_extended / extended_to_coeff in batches: extended_k = log_n, one pass with pre / post scaling
        k = log_n - 2
        d = h.EvaluationDomain(emu_ctx, 5, k)
        od = oracle_c.domain(5, k, 1)
        assert d.extended_k == log_n
        m = 1 << k
        cols = H.rand_fr_limbs(7 + log_n, ncols * m)
        src, dst = emu_ctx.upload_fr(cols), emu_ctx.alloc(ncols * n * 32)
        d.coeff_to_extended_device(src, dst, ncols)
        got = dst.download(ncols * n)
        for c in range(ncols):
            assert (got[c * n:(c + 1) * n] == od.coeff_to_extended(cols[c * m:(c + 1) * m])).all(), c
        back = emu_ctx.alloc(ncols * d.quotient_len * 32)
        d.extended_to_coeff_device(dst, back, ncols, divide_by_vanishing=True)
        gb = back.download(ncols * d.quotient_len)
        for c in range(ncols):
            ext = od.divide_by_vanishing_poly(got[c * n:(c + 1) * n])
            assert (gb[c * d.quotient_len:(c + 1) * d.quotient_len] == od.extended_to_coeff(ext)).all(), c
        for b in (src, dst, back):
            b.free()
        d.free()
        od.free()


@pytest.mark.parametrize("k,ncols,slot_cols", [(6, 7, 2.0), (9, 5, 1.0), (10, 9, 3.5)])
def test_streamed_host_batches(emu_ctx, oracle_c, monkeypatch, k, ncols, slot_cols):
    from tests import group_cases as G
    G.check_streamed_host_batches(emu_ctx, oracle_c, k, ncols, slot_cols, monkeypatch)
