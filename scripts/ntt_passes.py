import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h
k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
ctx = h.Context(0)
dom = h.EvaluationDomain(ctx, 2, k)
a = ctx.synth_scalars(n, 7, 0)
ctx.set_profile(True)
for name in ("omega", "omega_inv"):
    w = h.fr_encode([dom.constant(name)])
    for _ in range(3):
        ctx.best_fft_device(a, w, k)
    res = []
    for _ in range(5):
        t = time.perf_counter(); ctx.best_fft_device(a, w, k); ctx.sync(); dt = time.perf_counter() - t
        res.append((dt * 1e3, ctx.last_ntt_pass_ms()))
    res.sort()
    print(name, "total %.3f ms" % res[0][0], ["%.3f" % x for x in res[0][1]], flush=True)
ctx.close()
