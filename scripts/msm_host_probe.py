"""MSM with HOST (pinned) scalars on resident bases + window table: wall time per call.  env H2B_MSM_BATCH_MIN"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << k
ctx = h.Context(0)
bases = ctx.synth_bases(n, 0x6B7A67)
bases.precompute()
dev = ctx.synth_scalars(n, 1, 0)
pin = ctx.pinned((n, 4))
pin.array[:] = dev.download(n)
want = bases.msm(dev, n)
for _ in range(2):
    assert bases.msm(pin.array) == want
t0 = time.perf_counter()
for _ in range(reps):
    bases.msm(pin.array)
dt = (time.perf_counter() - t0) / reps
t0 = time.perf_counter()
for _ in range(reps):
    bases.msm(dev, n)
dd = (time.perf_counter() - t0) / reps
print(f"k={k} batch_min={os.environ.get('H2B_MSM_BATCH_MIN', 'default')}: host scalars {dt * 1e3:.2f} ms "
      f"({n / dt / 1e6:.1f} Mpts/s), device scalars {dd * 1e3:.2f} ms ({n / dd / 1e6:.1f} Mpts/s)")
ctx.close()
