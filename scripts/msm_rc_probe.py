"""Device-resident MSM on a window table: the two-dimensional bucket reduction against the segmented running sums
(H2B_MSM_RC=0), same result required.  usage: msm_rc_probe.py [k ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

ks = [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22, 24]
ctx = h.Context(0)
for k in ks:
    n = 1 << k
    bases = ctx.synth_bases(n, 0x6B7A67)
    bases.precompute()
    dev = ctx.synth_scalars(n, 1, 0)
    res = {}
    for rc in ("1", "0"):
        os.environ["H2B_MSM_RC"] = rc
        for _ in range(3):
            got = bases.msm(dev, n)
        reps = 20 if k <= 20 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            bases.msm(dev, n)
        res[rc] = ((time.perf_counter() - t0) / reps, got)
    assert res["1"][1] == res["0"][1], k
    print(f"k={k}: 2D reduction {res['1'][0] * 1e3:.3f} ms, running sums {res['0'][0] * 1e3:.3f} ms "
          f"({n / res['1'][0] / 1e6:.1f} vs {n / res['0'][0] / 1e6:.1f} Mpts/s)", flush=True)
    bases.free()
    dev.free()
ctx.close()
