"""MSM parameter sweep at k (default 24): window-table width c and level-0 chunk length L0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [22]
l0s = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64]
kinds = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
n = 1 << k
ctx = h.Context(0)
bases = ctx.synth_bases(n, 0x6B7A67)
ctx.set_profile(True)
for kind in kinds:
    sc = ctx.synth_scalars(n, 0x68616C6F32, kind)
    ref = None
    for c in cs:
        t = time.perf_counter()
        if c:
            bases.precompute(c)
        ctx.sync()
        tp = time.perf_counter() - t
        for l0 in l0s:
            os.environ["H2B_MSM_L0"] = str(l0)
            r = bases.msm(sc, n=n)
            ref = ref or r
            assert r == ref
            ts = []
            for _ in range(3):
                t = time.perf_counter(); bases.msm(sc, n=n); ts.append(time.perf_counter() - t)
            print(f"k={k} kind={kind} c={c} L0={l0}: msm {min(ts)*1e3:7.2f} ms  accum0 {ctx.last_kernel_ms():7.2f} ms  (table build {tp:.2f}s)", flush=True)
    sc.free()
ctx.close()
