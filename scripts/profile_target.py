"""Short workload for ncu captures: 2 MSMs and 2 NTTs at k (default 24), device-resident inputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
ctx = h.Context(0)
bases = ctx.synth_bases(n, 0x6B7A67)
if os.environ.get("H2B_NO_TABLE") != "1":
    bases.precompute()
sc = ctx.synth_scalars(n, 0x68616C6F32, 0)
poly = ctx.synth_scalars(n, 7, 0)
omega = h.fr_encode([h.EvaluationDomain(ctx, 2, k).constant("omega")])
for _ in range(2):
    bases.msm(sc, n=n)
    ctx.best_fft_device(poly, omega, k)
ctx.sync()
print("launches", ctx.launches)
ctx.close()
