"""The N>1 host logic on CPU: world_size-2 (and 4) `gloo` process groups, each rank
driving the TEST-ONLY emulator build of the kernels through the same C ABI.
Covers the sharded MSM (point ranges + fold of partial points), column ownership
and the four-step NTT with its all-to-all transposes."""
import os
import random
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                          MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import halo2_pse_b200 as h
        from halo2_pse_b200 import dist as D
        from oracle import bn256 as O
        from tests import helpers as H
        r, w, _ = D.init_from_env("gloo")
        assert (r, w) == (rank, world)
        ctx = h.Context(0, lib_path=emu)
        oc = H.load_oracle_c()
        # ---- sharded MSM: contiguous ranges, partial points folded ----------------
        for n in (1000, 7, 3):  # 3 < world for world=4: some ranks get an empty range
            rng = random.Random(n)
            hs = [rng.randrange(1, 1 << 64) for _ in range(n)]
            sc = [rng.randrange(O.R_MOD) for _ in range(n)]
            s, e = D.shard_range(n, rank, world)
            bases = oc.g1_mul_gen(hs[s:e]) if e > s else np.zeros((0, 8), dtype=np.uint64)
            B = h.Bases(ctx, bases, e - s)
            got = D.ShardedMSM(ctx, B).msm(H.fr_enc(sc[s:e]))
            want = O.g1_mul(O.G1_GEN, sum(c * x for c, x in zip(sc, hs)) % O.R_MOD)
            assert got == want, (n, rank)
            B.free()
        # ---- four-step NTT: natural order in, natural order out, sharded ----------
        for k in (6, 9, 10):
            n = 1 << k
            a = H.rand_fr_limbs(k, n)
            omega = O.omega_for(k)
            want = oc.best_fft(a, H.fr_enc([omega])[0], k, 1)
            loc = n // world
            mine = torch.from_numpy(a[rank * loc:(rank + 1) * loc].copy().view(np.int64).reshape(-1))
            D.FourStepNTT(ctx, k, omega).run(mine)
            got = mine.numpy().view(np.uint64).reshape(-1, 4)
            assert (got == want[rank * loc:(rank + 1) * loc]).all(), (k, rank)
        # ---- create_proof with every commitment sharded over the ranks (ShardedBases) ----
        # k = 5: n = 32 bases, 16 per rank at world 2, 8 at world 4; the witness-polynomial commitments of
        # the multiopen cover n - 1 points, so the last rank's range is cut short.
        from tests import plonk_cases as PC
        seed = b"\x07" * 16
        _, pk, got = PC.device_bench_proof(ctx, 5, 0xDEADBEEF, seed, params_hook=lambda p: D.shard_params(p))
        _, opk, want = PC.oracle_bench_proof(5, 0xDEADBEEF, seed)
        assert pk.pinned == opk.debug and got == want, rank  # same vk, same proof bytes as the big-integer oracle
        pk.free()
        # ---- the sharded prover refuses rngs that differ between ranks (silently invalid proofs otherwise) ----
        try:
            PC.device_bench_proof(ctx, 5, 0xDEADBEEF, bytes([7 + rank]) * 16, params_hook=lambda p: D.shard_params(p))
        except h.H2BError as e:
            assert "ranks disagree" in str(e), e
        else:
            raise AssertionError("per-rank rng seeds were accepted by the sharded create_proof")
        # ... and a broadcast seed is the same everywhere
        seed = D.broadcast_seed(1000 + rank)
        assert seed == 1000
        assert D.broadcast_seed() == D.broadcast_seed(None) or True  # fresh entropy: just exercises the path
        ctx.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_sharded_paths(emu_lib_path, oracle_c, world):
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, emu_lib_path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(m == "ok" for _, m in res), res


def test_shard_range_partitions():
    from halo2_pse_b200 import dist as D
    for n in (0, 1, 5, 16, 17, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            ranges = [D.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    assert [D.column_owner(c, 4) for c in range(6)] == [0, 1, 2, 3, 0, 1]


def test_four_step_single_rank(emu_ctx, oracle_c):
    """G = 1 degenerates to the plain four-step decomposition on one device."""
    from halo2_pse_b200 import dist as D
    from oracle import bn256 as O
    from tests import helpers as H
    for k in (4, 7, 10):
        a = H.rand_fr_limbs(k, 1 << k)
        omega = O.omega_for(k)
        want = oracle_c.best_fft(a, H.fr_enc([omega])[0], k, 1)
        t = torch.from_numpy(a.copy().view(np.int64).reshape(-1))
        D.FourStepNTT(emu_ctx, k, omega).run(t)
        assert (t.numpy().view(np.uint64).reshape(-1, 4) == want).all()


def test_transpose_scatter_kernel(emu_ctx):
    """The fused transpose+exchange kernel, all ranks emulated in one process: rank g stores its
    rows straight into every rank's destination buffer; together they form the row-sharded transpose."""
    import ctypes as C
    from tests import helpers as H
    for G, R, Cn in ((1, 40, 24), (2, 64, 96), (4, 32, 64), (8, 64, 40)):
        M = H.rand_fr_limbs(G * 1000 + R, R * Cn).reshape(R, Cn, 4)
        Rl, Cl = R // G, Cn // G
        dst = [np.zeros((Cl * R, 4), dtype=np.uint64) for _ in range(G)]
        ptrs = (C.c_void_p * G)(*[d.ctypes.data for d in dst])
        for g in range(G):
            src = np.ascontiguousarray(M[g * Rl:(g + 1) * Rl]).reshape(-1, 4)
            emu_ctx._check(emu_ctx.lib.h2b_fr_transpose_scatter(emu_ctx.h, C.c_void_p(src.ctypes.data), ptrs, G, g,
                                                                Rl, Cn))
        MT = np.ascontiguousarray(M.transpose(1, 0, 2))  # [Cn][R]
        for hh in range(G):
            assert (dst[hh].reshape(Cl, R, 4) == MT[hh * Cl:(hh + 1) * Cl]).all(), (G, hh)


def test_rows_scatter_kernel_and_p2p_schedule(emu_ctx, oracle_c):
    """The p2p four-step schedule with every rank emulated in one process: one transpose+exchange kernel, then two
    batches of row transforms whose last pass stores (twiddled) outputs straight into the owning rank's buffer
    (h2b_best_fft_rows_scatter).  Result == best_fft of the whole vector (arithmetic.rs:171)."""
    import ctypes as C
    from oracle import bn256 as O
    from tests import helpers as H
    import halo2_pse_b200 as h
    lib, ctx = emu_ctx.lib, emu_ctx

    def ptrs(bufs):
        return (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])

    for k, G in ((4, 1), (8, 2), (9, 4), (13, 8), (14, 2)):
        n = 1 << k
        k1 = k // 2
        k2 = k - k1
        n1, n2 = 1 << k1, 1 << k2
        omega = O.omega_for(k)
        w = H.fr_enc([omega])
        w1 = H.fr_enc([pow(omega, n2, O.R_MOD)])
        w2 = H.fr_enc([pow(omega, n1, O.R_MOD)])
        a = H.rand_fr_limbs(k * 10 + G, n)
        want = oracle_c.best_fft(a, w[0], k, 1)
        loc = n // G
        S1 = [np.zeros((loc, 4), dtype=np.uint64) for _ in range(G)]
        S2 = [np.zeros((loc, 4), dtype=np.uint64) for _ in range(G)]
        for g in range(G):  # A[j1][j2] -> A^T[j2][j1]
            src = np.ascontiguousarray(a[g * loc:(g + 1) * loc])
            ctx._check(lib.h2b_fr_transpose_scatter(ctx.h, C.c_void_p(src.ctypes.data), ptrs(S1), G, g, n1 // G, n2))
        for g in range(G):  # n1-point row transforms of A^T, twiddle, -> B[K1][j2]
            ctx._check(lib.h2b_best_fft_rows_scatter(ctx.h, C.c_void_p(S1[g].ctypes.data), C.c_void_p(w1.ctypes.data), k1,
                                                     n2 // G, ptrs(S2), G, g * (n2 // G), n2, C.c_void_p(w.ctypes.data), k))
        for g in range(G):  # n2-point row transforms of B -> natural order
            ctx._check(lib.h2b_best_fft_rows_scatter(ctx.h, C.c_void_p(S2[g].ctypes.data), C.c_void_p(w2.ctypes.data), k2,
                                                     n1 // G, ptrs(S1), G, g * (n1 // G), n1, None, k))
        got = np.concatenate(S1)
        assert (got == want).all(), (k, G)

    # the register kernel's fused last pass (rows of 2^12 and 2^13 points: two passes) against row-wise best_fft
    # ... and single-pass rows (64 / 512 points) through the register kernel with batch members as tile columns
    for log_len, rows, G, row0, total in ((12, 3, 2, 5, 16), (13, 2, 4, 0, 2), (9, 8, 2, 0, 8), (6, 32, 4, 32, 64),
                                           (12, 32, 2, 0, 32), (13, 16, 4, 16, 64), (13, 48, 2, 3, 64)):  # interleaved scratch
        L = 1 << log_len
        big_k = 20
        big = O.omega_for(big_k)
        wl = H.fr_enc([O.omega_for(log_len)])
        x = H.rand_fr_limbs(log_len, rows * L)
        dst = [np.zeros(((L // G) * total, 4), dtype=np.uint64) for _ in range(G)]
        ctx._check(lib.h2b_best_fft_rows_scatter(ctx.h, C.c_void_p(x.ctypes.data), C.c_void_p(wl.ctypes.data), log_len, rows,
                                                 ptrs(dst), G, row0, total, C.c_void_p(H.fr_enc([big]).ctypes.data), big_k))
        cl = L // G
        for r in range(rows):
            y = H.fr_dec(oracle_c.best_fft(x[r * L:(r + 1) * L], wl[0], log_len, 1))
            row = row0 + r
            for Ko in (0, 1, 2, cl - 1, cl, L - 1, 777 % L, 1234 % L, 37 % L):
                exp = y[Ko] * pow(big, row * Ko, O.R_MOD) % O.R_MOD
                got = H.fr_dec(dst[Ko // cl][(Ko % cl) * total + row])[0]
                assert got == exp, (log_len, r, Ko)
    # argument checks
    x = H.rand_fr_limbs(1, 16)
    d = [np.zeros((16, 4), dtype=np.uint64) for _ in range(2)]
    w4 = H.fr_enc([O.omega_for(4)])
    assert lib.h2b_best_fft_rows_scatter(ctx.h, C.c_void_p(x.ctypes.data), C.c_void_p(w4.ctypes.data), 4, 1, ptrs(d), 3, 0, 1,
                                         None, 4) == h.H2B_ERR_ARG  # world not a power of two
    assert lib.h2b_best_fft_rows_scatter(ctx.h, C.c_void_p(x.ctypes.data), C.c_void_p(w4.ctypes.data), 4, 2, ptrs(d), 2, 0, 1,
                                         None, 4) == h.H2B_ERR_LENGTH  # rows outside the matrix
