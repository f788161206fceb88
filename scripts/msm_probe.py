"""Wall time of one device-resident MSM on a window table vs the sum of its kernels (run once plainly for
the wall time, once under `ncu --metrics gpu__time_duration.sum` for the kernel list)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 1 << k
ctx = h.Context(0)
bases = ctx.synth_bases(n, 0x6B7A67)
bases.precompute()
sc = ctx.synth_scalars(n, 1, 0)
for _ in range(3):
    bases.msm(sc, n)
l0 = ctx.launches
t0 = time.perf_counter()
for _ in range(reps):
    bases.msm(sc, n)
dt = (time.perf_counter() - t0) / reps
print(f"k={k} c={bases.table_window_bits}: {dt * 1e3:.3f} ms per MSM wall, {(ctx.launches - l0) // reps} launches, "
      f"{n / dt / 1e6:.1f} Mpts/s")
ctx.close()
