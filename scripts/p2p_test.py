"""torchrun script: four-step NTT with NVLink peer stores vs the single-GPU transform (k = 20, 24) + timing at k=26."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_pse_b200 as h
from halo2_pse_b200 import dist as D

rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
ctx = h.Context(local)
dev = torch.device("cuda", local)
for kv in (20, 24):
    wv = h.EvaluationDomain(ctx, 2, kv).constant("omega")
    full = ctx.synth_scalars(1 << kv, 12345, 0)
    lv = (1 << kv) // world
    mine = np.ascontiguousarray(full.download(1 << kv)[rank * lv:(rank + 1) * lv])
    ctx.best_fft_device(full, h.fr_encode([wv]), kv)
    want = full.download(1 << kv)[rank * lv:(rank + 1) * lv]
    for p2p in (True, False):
        part = torch.from_numpy(mine.copy().view(np.int64).reshape(-1)).to(dev)
        fs = D.FourStepNTT(ctx, kv, wv, p2p=p2p)
        fs.run(part)
        got = part.cpu().numpy().view(np.uint64).reshape(-1, 4)
        print(f"rank {rank} k={kv} p2p={fs.p2p} ok={bool((got == want).all())}", flush=True)
        # run twice more: buffers are reused across transforms
        part2 = torch.from_numpy(mine.copy().view(np.int64).reshape(-1)).to(dev)
        fs.run(part2)
        assert (part2.cpu().numpy().view(np.uint64).reshape(-1, 4) == want).all()
    full.free()
k4 = 26
w4 = h.EvaluationDomain(ctx, 2, k4).constant("omega")
loc = (1 << k4) // world
buf = torch.empty(loc * 4, dtype=torch.int64, device=dev)
ctx._check(ctx.lib.h2b_synth_scalars(ctx.h, C.c_void_p(buf.data_ptr()), loc, 77 + rank, 0))
for p2p in (True, False):
    fs = D.FourStepNTT(ctx, k4, w4, p2p=p2p)
    fs.run(buf); torch.distributed.barrier()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); torch.distributed.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fs.run(buf); torch.cuda.synchronize(); torch.distributed.barrier(); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        print(f"k={k4} world={world} p2p={p2p}: {min(ts):.3f} ms -> {(1 << k4) / min(ts) / 1e3:.0f} Melem/s", flush=True)
ctx.close()
torch.distributed.destroy_process_group()
