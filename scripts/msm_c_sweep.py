"""Device-resident MSM on a window table: wall time per call for a range of window widths c around the library's
automatic choice.  usage: msm_c_sweep.py k [k ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

ctx = h.Context(0)
for k in [int(a) for a in sys.argv[1:]] or [16, 18, 20]:
    n = 1 << k
    bases = ctx.synth_bases(n, 0x6B7A67)
    dev = ctx.synth_scalars(n, 1, 0)
    bases.precompute()
    auto = bases.table_window_bits
    out = []
    for c in range(max(8, auto - 4), min(24, auto + 2) + 1):
        if ((255 + c - 1) // c) * n >= 1 << 31:
            continue
        bases.precompute(c)
        for _ in range(3):
            bases.msm(dev, n)
        reps = 30 if k <= 20 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            bases.msm(dev, n)
        out.append((c, (time.perf_counter() - t0) / reps * 1e3))
    print(f"k={k} auto c={auto}: " + "  ".join(f"c={c}: {t:.3f}" for c, t in out), flush=True)
    bases.free()
    dev.free()
ctx.close()
