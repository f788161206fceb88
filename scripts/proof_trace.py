"""Per-call wall times of the host-mirror calls made during the `evals` stage of create_proof (debugging aid)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import halo2_pse_b200 as h  # noqa: E402

log = []


def wrap(cls, name):
    f = getattr(cls, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        log.append((name, t0, time.perf_counter() - t0))
        return r
    setattr(cls, name, g)


for nm in ("eval_polynomial", "alloc", "memset", "sync", "kate_division", "clone"):
    wrap(h.Context, nm)
wrap(h.DeviceBuffer, "free")
wrap(h.Blake2bWrite, "squeeze_challenge_scalar")
wrap(h.Blake2bWrite, "write_scalar")
ctx = h.Context(0)
r = bench.gpu_create_proof(ctx, h)
print(json.dumps({"seconds": r["seconds"], "evals": r["stages_seconds"]["evals"]}))
# the slowest calls overall, and the 40 calls around the slowest eval_polynomial
slow = sorted(log, key=lambda x: -x[2])[:12]
print("slowest:", [(n, round(d * 1e3, 3)) for n, _, d in slow])
ev = [i for i, x in enumerate(log) if x[0] == "eval_polynomial"]
worst = max(ev, key=lambda i: log[i][2])
print("around worst eval:", [(n, round(d * 1e3, 3)) for n, _, d in log[max(0, worst - 6):worst + 6]])
tail = log[-260:]
print("last proof, calls over 0.5 ms:", [(i, n, round(d * 1e3, 2)) for i, (n, _, d) in enumerate(tail) if d > 5e-4])
# gaps between consecutive calls (time spent outside the wrapped calls)
gaps = [(i, tail[i][0], round((tail[i][1] - (tail[i - 1][1] + tail[i - 1][2])) * 1e3, 2)) for i in range(1, len(tail))]
print("gaps over 0.5 ms:", [g for g in gaps if g[2] > 0.5])
ctx.close()
