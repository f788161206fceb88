"""ONE create_proof of the bench circuit on all ranks (bases sharded by point range), under torchrun:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/proof_multi.py [k]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import halo2_pse_b200 as h  # noqa: E402
from halo2_pse_b200 import dist as D  # noqa: E402

if len(sys.argv) > 1:
    bench.PROOF_K = int(sys.argv[1])
rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
ctx = h.Context(local)


def barrier():
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()


def max_over_ranks(*vals):
    t = torch.tensor(vals, dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return t.tolist()


r = bench.gpu_create_proof(ctx, h, world, barrier, max_over_ranks) if world > 1 else bench.gpu_create_proof(ctx, h)
if rank == 0:
    keep = {k: r[k] for k in ("seconds", "stages_seconds", "gpu_launches", "keygen_seconds") if k in r}
    keep.update(n_gpus=world, k=bench.PROOF_K, same_bytes=r.get("proof_bytes_equal_single_gpu_prover_on_every_rank"))
    print("PROOF " + json.dumps(keep), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(keep, open(f"gpurun_out/proof_multi_n{world}_k{bench.PROOF_K}.json", "w"), indent=1)
ctx.close()
if world > 1:
    torch.distributed.destroy_process_group()
