"""MSM with HOST (pinned) scalars on resident bases + window table: wall time per call, for a list of batch plans.
usage: msm_host_probe.py [k] [reps] [plan ...]   (plan = H2B_MSM_BATCH_PLAN weights, e.g. 1,3,9; "default" = library rule)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
plans = sys.argv[3:] or ["default"]
n = 1 << k
ctx = h.Context(0)
bases = ctx.synth_bases(n, 0x6B7A67)
bases.precompute()
dev = ctx.synth_scalars(n, 1, 0)
pin = ctx.pinned((n, 4))
pin.array[:] = dev.download(n)
want = bases.msm(dev, n)
t0 = time.perf_counter()
for _ in range(reps):
    bases.msm(dev, n)
dd = (time.perf_counter() - t0) / reps
print(f"k={k}: device scalars {dd * 1e3:.2f} ms ({n / dd / 1e6:.1f} Mpts/s)", flush=True)
for plan in plans:
    if plan == "default":
        os.environ.pop("H2B_MSM_BATCH_PLAN", None)
    else:
        os.environ["H2B_MSM_BATCH_PLAN"] = plan
    for _ in range(2):
        assert bases.msm(pin.array) == want, plan
    t0 = time.perf_counter()
    for _ in range(reps):
        bases.msm(pin.array)
    dt = (time.perf_counter() - t0) / reps
    print(f"k={k} plan={plan}: host scalars {dt * 1e3:.2f} ms ({n / dt / 1e6:.1f} Mpts/s)", flush=True)
ctx.close()
