import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_pse_b200 as h
from halo2_pse_b200 import dist as D
rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
ctx = h.Context(local)
dev = torch.device("cuda", local)
for k in [int(x) for x in (sys.argv[1:] or ["22", "24", "26"])]:
    w = h.EvaluationDomain(ctx, 2, k).constant("omega")
    loc = (1 << k) // world
    full = ctx.synth_scalars(1 << k, 1234 + k, 0)
    ctx.sync()
    ft = D._as_tensor(full, (1 << k) * 4, dev)
    orig = ft[rank * loc * 4:(rank + 1) * loc * 4].clone()
    torch.cuda.synchronize()
    ctx.best_fft_device(full, h.fr_encode([w]), k)
    ctx.sync()
    want = ft[rank * loc * 4:(rank + 1) * loc * 4].clone()
    torch.cuda.synchronize()
    for label, p2p in (("p2p", None), ("nccl", False)):
        fs = D.FourStepNTT(ctx, k, w, p2p=p2p)
        buf = orig.clone()
        torch.cuda.synchronize()
        res = fs.run(buf)
        torch.cuda.synchronize()
        diff = (res.view(-1, 4) != want.view(-1, 4)).any(dim=1)
        nbad = int(diff.sum().item())
        first = diff.nonzero()[:6].flatten().tolist()
        print(f"rank {rank} k={k} {label} k1={fs.k1} p2p={fs.p2p}: mismatching elements {nbad} of {loc} first {first}", flush=True)
        if fs.p2p and os.environ.get("H2B_FOURSTEP_TIMING"):
            for _ in range(3):
                fs.run(buf)
            torch.cuda.synchronize()
            if rank in (0, world - 1):
                print(f"rank {rank} k={k} stages ms:", {kk: round(v, 3) for kk, v in fs.last_stage_ms.items()}, flush=True)
        del fs
    full.free()
torch.distributed.barrier()
ctx.close()
torch.distributed.destroy_process_group()
