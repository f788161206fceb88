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
