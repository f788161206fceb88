"""Multi-GPU parity tests proper (`-m gpu`, skipped below 2 devices): one process per GPU over NCCL,
every rank drives libhalo2b200.so through the C ABI.

  * ShardedBases / ShardedMSM commitments == the single-GPU commitment == the oracle
    (best_multiexp's per-chunk fold, arithmetic.rs:132-159);
  * FourStepNTT over NVLink peer stores and over the NCCL all-to-all == h2b_best_fft of the
    same vector on one GPU == the oracle (best_fft, arithmetic.rs:171), k = 12 ... 22;
  * ONE create_proof sharded over the ranks: proof bytes == the single-GPU prover's == the oracle's.
"""
import os
import random
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                          MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import ctypes as C

        import halo2_pse_b200 as h
        from halo2_pse_b200 import dist as D
        from oracle import bn256 as O
        from tests import helpers as H
        torch.cuda.set_device(rank)
        r, w, _ = D.init_from_env("nccl")
        assert (r, w) == (rank, world)
        dev = torch.device("cuda", rank)
        ctx = h.Context(rank)
        oc = H.load_oracle_c()
        notes = []

        # ---- sharded MSM against the single-GPU MSM and the oracle -----------------------------
        for n in (5000, 1 << 14, 3):
            rng = random.Random(n)
            hs = [rng.randrange(1, 1 << 64) for _ in range(n)]
            bases = oc.g1_mul_gen(hs)
            sc = H.rand_fr_limbs(n + 1, n)
            want = H.g1_dec(oc.best_multiexp(sc, bases, 0))[0]
            full = h.Bases(ctx, bases, n)
            if n >= 1024:
                full.precompute()
            assert full.msm(sc) == want, ("single", n, rank)
            s, e = D.shard_range(n, rank, world)
            part = h.Bases(ctx, np.ascontiguousarray(bases[s:e]), e - s)
            if e - s >= 1024:
                part.precompute()
            assert D.ShardedMSM(ctx, part).msm(np.ascontiguousarray(sc[s:e])) == want, ("sharded", n, rank)
            # ShardedBases: the replicated scalar vector, prefix / offset forms of ParamsKZG::commit
            sb = D.ShardedBases.from_full(ctx, full, precompute=(n >= 4096))  # frees `full`
            assert sb.msm(sc) == want, ("ShardedBases", n, rank)
            if n > 100:
                m, off = n - 37, 11
                w2 = H.g1_dec(oc.best_multiexp(sc[5:5 + m], bases[off:off + m], 0))[0]
                assert sb.msm(sc, m, off, scalar_offset=5) == w2, ("ShardedBases slice", n, rank)
            sb.free()
            part.free()
        notes.append("msm")

        # ---- four-step NTT: both exchange methods against h2b_best_fft and the oracle -----------
        for k in (12, 16, 20, 22):
            n = 1 << k
            omega = O.omega_for(k)
            a = H.rand_fr_limbs(100 + k, n)
            fullb = ctx.upload_fr(a)
            ctx.best_fft_device(fullb, h.fr_encode([omega]), k)
            single = fullb.download(n)
            fullb.free()
            if k <= 20:
                assert (single == oc.best_fft(a, H.fr_enc([omega])[0], k, 0)).all(), ("single vs oracle", k)
            loc = n // world
            for p2p in (None, False):
                fs = D.FourStepNTT(ctx, k, omega, p2p=p2p)
                if p2p is None and not fs.p2p:
                    notes.append(f"p2p unavailable: {fs.p2p_error}")
                    continue
                mine = torch.from_numpy(a[rank * loc:(rank + 1) * loc].copy().view(np.int64).reshape(-1)).to(dev)
                out = fs.run(mine)
                torch.cuda.synchronize()
                got = out.cpu().numpy().view(np.uint64).reshape(-1, 4)
                assert (got == single[rank * loc:(rank + 1) * loc]).all(), ("four-step", k, p2p, rank)
                # a second run on the same object (buffers and barriers are reused)
                mine2 = torch.from_numpy(a[rank * loc:(rank + 1) * loc].copy().view(np.int64).reshape(-1)).to(dev)
                out2 = fs.run(mine2)
                torch.cuda.synchronize()
                assert torch.equal(out2.cpu(), out.cpu()), ("four-step rerun", k, p2p, rank)
        notes.append("four-step")

        # ---- ONE create_proof sharded over the ranks --------------------------------------------
        from tests import plonk_cases as PC
        seed = b"\x07" * 16
        for k in (5, 10):
            _, pk1, single = PC.device_bench_proof(ctx, k, 0xDEADBEEF, seed)
            pk1.free()
            _, pk, got = PC.device_bench_proof(ctx, k, 0xDEADBEEF, seed, params_hook=lambda p: D.shard_params(p))
            assert got == single, ("sharded proof != single-GPU proof", k, rank)
            if k == 5:
                _, opk, want = PC.oracle_bench_proof(k, 0xDEADBEEF, seed)
                assert pk.pinned == opk.debug and got == want, ("sharded proof != oracle", rank)
            pk.free()
        notes.append("create_proof")
        assert ctx.launches > 0
        ctx.close()
        torch.distributed.destroy_process_group()
        q.put((rank, "ok", notes))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc(), []))


def _run(world):
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(m == "ok" for _, m, _ in res), res
    return res


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_sharded_paths(oracle_c, world):
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: there is no CPU fallback")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    _run(world)
