"""Chunk-length sweep of the level-wise bucket accumulation (H2B_MSM_L0 / H2B_MSM_LN) on a window table.
usage: msm_l_sweep.py k [k ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

ctx = h.Context(0)
for k in [int(a) for a in sys.argv[1:]] or [16, 20]:
    n = 1 << k
    bases = ctx.synth_bases(n, 0x6B7A67).precompute()
    dev = ctx.synth_scalars(n, 1, 0)
    want = bases.msm(dev, n)
    res = []
    for ln in [int(x) for x in os.environ.get("LNS", "8,16,32").split(",")]:
        for l0 in [int(x) for x in os.environ.get("L0S", "0,16,24,32,48,64,96,128,192").split(",")]:
            os.environ["H2B_MSM_LN"] = str(ln)
            if l0:
                os.environ["H2B_MSM_L0"] = str(l0)
            else:
                os.environ.pop("H2B_MSM_L0", None)
            for _ in range(2):
                assert bases.msm(dev, n) == want
            reps = 20 if k <= 20 else 5
            t0 = time.perf_counter()
            for _ in range(reps):
                bases.msm(dev, n)
            res.append((ln, l0, (time.perf_counter() - t0) / reps * 1e3))
    print(f"k={k}: " + "  ".join(f"LN{ln}/L0={l0 or 'auto'}:{t:.3f}" for ln, l0, t in res), flush=True)
    bases.free()
    dev.free()
ctx.close()
