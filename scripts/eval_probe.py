"""eval_polynomial / poly_fma wall times vs k on cuda:0 (device-resident)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

ctx = h.Context(0)
for k in (16, 18, 20, 21, 22, 23, 24):
    n = 1 << k
    a = ctx.synth_scalars(n, 3, 0)
    x = 0x1234567890ABCDEF1234567890ABCDEF1234567
    ctx.eval_polynomial(a, x, n)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        ctx.eval_polynomial(a, x + _, n)
        ts.append(time.perf_counter() - t0)
    print(f"k={k}: eval_polynomial min {min(ts) * 1e3:.3f} ms, median {sorted(ts)[5] * 1e3:.3f} ms", flush=True)
    a.free()
ctx.close()
