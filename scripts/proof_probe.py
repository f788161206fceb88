"""create_proof of the bench circuit at k (default 20) on cuda:0: seconds and stage times.  env: H2B_COMMIT_WAYS"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import halo2_pse_b200 as h  # noqa: E402

if len(sys.argv) > 1:
    bench.PROOF_K = int(sys.argv[1])
ctx = h.Context(0)
r = bench.gpu_create_proof(ctx, h)
print(json.dumps({"ways": os.environ.get("H2B_COMMIT_WAYS", "default"), "k": bench.PROOF_K, "seconds": r["seconds"],
                  "stages": r["stages_seconds"], "launches": r["gpu_launches"]}))
ctx.close()
