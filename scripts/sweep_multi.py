"""Multi-GPU sweep (BASELINE.json configs[1], [2]) under torchrun, one rank per GPU:
  * ONE MSM of 2^k points, k = 16..26, sharded by contiguous point range over the ranks (strong scaling):
    local Pippenger on each rank's resident bases + window table, 64-byte partial points all-gathered and folded;
  * 64 independent columns of coeff_to_extended / extended_to_coeff (j = 5) at k = 16..24, distributed by column
    (column c on rank c % world, no communication), where they fit in memory.
Times are wall clock around synchronous calls bracketed by barriers, best of 3, MAX over ranks.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep_multi.py [k,k,...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import halo2_pse_b200 as h  # noqa: E402
from halo2_pse_b200 import dist as D  # noqa: E402

rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
ctx = h.Context(local)
ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(16, 27, 2))
NCOLS = 64


def barrier():
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()


def max_over_ranks(v):
    t = torch.tensor([v], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        barrier()
        t = time.perf_counter()
        fn()
        ctx.sync()
        ts.append(time.perf_counter() - t)
    return max_over_ranks(min(ts))


out = {"n_gpus": world, "msm": {}, "ntt_x64": {}}
for k in ks:
    n = 1 << k
    s, e = D.shard_range(n, rank, world)
    bases = ctx.synth_bases(e - s, 0x6B7A67 + rank).precompute()
    sc = ctx.synth_scalars(e - s, 5 + rank, 0)
    msm = D.ShardedMSM(ctx, bases)
    t = best(lambda: msm.msm(sc, e - s))
    cw = bases.table_window_bits
    Ww = (255 + cw - 1) // cw
    out["msm"][k] = {"points_per_gpu": e - s, "window_bits": cw, "ms": t * 1e3, "mpts_s": n / t / 1e6,
                     # whole sharded MSM against the IMAD.WIDE peak of ALL ranks (algorithmic n*W*11*136)
                     "roofline_frac": n * Ww * 11 * 136 / t / (ctx.pipe_peak("imad_wide")[0] * world)}
    sc.free()
    bases.free()
    if rank == 0:
        print("msm", k, json.dumps(out["msm"][k]), flush=True)
imad = ctx.pipe_peak("imad_wide")[0]
for k in ks:
    n = 1 << k
    dom = h.EvaluationDomain(ctx, 5, k)
    ne, ek = dom.extended_len(), dom.extended_k
    mine = len([c for c in range(NCOLS) if D.column_owner(c, world) == rank])
    # this rank's columns, inputs resident; the extended outputs go through a ring of column groups (<= 16 GiB)
    group = max(1, min(mine, (16 << 30) // (ne * 32)))
    src, dst = ctx.alloc(mine * n * 32), ctx.alloc(group * ne * 32)
    ctx._check(ctx.lib.h2b_synth_scalars(ctx.h, src.ptr, mine * n, 9 + rank, 0))
    lib = ctx.lib

    def c2e():
        for c0 in range(0, mine, group):
            nc = min(group, mine - c0)
            ctx._check(lib.h2b_coeff_to_extended_batch(dom.h, src.at(c0 * n * 32), n, dst.ptr, ne, h.H2B_DEVICE, nc))

    def e2c():
        for c0 in range(0, mine, group):
            nc = min(group, mine - c0)
            ctx._check(lib.h2b_extended_to_coeff_batch(dom.h, dst.ptr, ne, dst.ptr, ne, h.H2B_DEVICE, nc, 1))

    t1 = best(c2e, reps=2)
    t2 = best(e2c, reps=2)
    out["ntt_x64"][k] = {"columns_per_gpu": mine, "columns_per_group": group, "extended_k": ek,
                         "coeff_to_extended_ms": t1 * 1e3,
                         "coeff_to_extended_melem_s_out": NCOLS * ne / t1 / 1e6,
                         "coeff_to_extended_int_frac_per_gpu": mine * (ne / 2) * ek * 136 / t1 / imad,
                         "extended_to_coeff_ms": t2 * 1e3, "extended_to_coeff_melem_s_in": NCOLS * ne / t2 / 1e6,
                         "extended_to_coeff_int_frac_per_gpu": mine * (ne / 2) * ek * 136 / t2 / imad}
    src.free()
    dst.free()
    dom.free()
    if rank == 0:
        print("ntt_x64", k, json.dumps(out["ntt_x64"][k]), flush=True)
barrier()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/r2_sweep_multi_n{world}.json", "w"), indent=1)
ctx.close()
if world > 1:
    torch.distributed.destroy_process_group()
