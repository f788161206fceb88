"""Exploration probe run on the GPU box (not a test, not the bench)."""
import os, sys, time, random, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import halo2_pse_b200 as h
from oracle import bn256 as O

ctx = h.Context(0)
rng = random.Random(1)
res = {}

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    ctx.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ctx.sync(); ts.append(time.perf_counter() - t)
    return min(ts), sorted(ts)[len(ts)//2]

# field parity
for field, mod in ((0, O.R_MOD), (1, O.Q_MOD)):
    enc = h.fr_encode if field == 0 else h.fq_encode
    dec = h.fr_decode if field == 0 else h.fq_decode
    a = [rng.randrange(mod) for _ in range(500)] + [0, 1, mod - 1]
    b = [rng.randrange(mod) for _ in range(500)] + [mod - 1] * 3
    A, B = enc(a), enc(b)
    for op, f in ((0, lambda x, y: x * y % mod), (1, lambda x, y: (x + y) % mod), (2, lambda x, y: (x - y) % mod)):
        out = np.zeros_like(A)
        ctx._check(ctx.lib.h2b_test_field_op(ctx.h, field, op, A.ctypes.data, B.ctypes.data, out.ctypes.data, len(a)))
        assert dec(out) == [f(x, y) for x, y in zip(a, b)], (field, op)
print("field parity ok", flush=True)
print("imad peak (mad.wide.u32/s):", ctx.imad_peak(), flush=True)
res["imad_peak"] = ctx.imad_peak()

# NTT parity small
for k in [3, 8, 10, 12, 13, 14, 16, 17, 18]:
    n = 1 << k
    a = [rng.randrange(O.R_MOD) for _ in range(n)]
    w = O.omega_for(k)
    exp = list(a); O.best_fft(exp, w, k)
    arr = h.fr_encode(a); ctx.best_fft(arr, w, k)
    ok = h.fr_decode(arr) == exp
    print("ntt parity", k, ok, flush=True)
    assert ok
# NTT timing + roundtrip at large k
for k in [16, 18, 20, 22, 24, 26]:
    n = 1 << k
    buf = ctx.synth_scalars(n, 7, 0)
    ref = buf.download(min(n, 4096))
    w, wi = O.omega_for(k), pow(O.omega_for(k), -1, O.R_MOD)
    ctx.best_fft_device(buf, w, k); ctx.best_fft_device(buf, wi, k)
    back = h.fr_decode(buf.download(min(n, 4096)))
    exp = [x * n % O.R_MOD for x in h.fr_decode(ref)]
    print("ntt roundtrip", k, back == exp, flush=True)
    best, med = timed(lambda: ctx.best_fft_device(buf, w, k))
    print(f"ntt k={k}: best {best*1e3:.3f} ms med {med*1e3:.3f} ms -> {n/best/1e6:.1f} Melem/s", flush=True)
    res[f"ntt_{k}_ms"] = best * 1e3
    buf.free()

# MSM parity small
G = O.G1_GEN
for n in [1, 5, 100, 1000, 5000]:
    hs = [rng.randrange(1, 1 << 64) for _ in range(n)]
    bases = O.batch_to_affine([O._jac_mul(O._to_jac(G), x) for x in hs])
    Bs = h.Bases(ctx, h.g1_encode(bases), n)
    for tag, sc in (("uni", [rng.randrange(O.R_MOD) for _ in range(n)]), ("eq", [12345] * n), ("01", [rng.randrange(2) for _ in range(n)])):
        got = Bs.msm(h.fr_encode(sc))
        exp = O.g1_mul(G, sum(s * x for s, x in zip(sc, hs)) % O.R_MOD)
        print("msm parity", n, tag, got == exp, flush=True)
        assert got == exp
    Bs.free()
# MSM timing with closed-form check
for k in [16, 18, 20, 22, 24]:
    n = 1 << k
    t = time.perf_counter(); Bs = ctx.synth_bases(n, 99); ctx.sync()
    print(f"synth_bases k={k}: {time.perf_counter()-t:.2f}s", flush=True)
    for kind in ([0, 1, 2, 4] if k <= 20 else [0]):
        sc = ctx.synth_scalars(n, 5, kind)
        got = Bs.msm(sc, n=n)
        if k <= 18:
            s = h.fr_decode(sc.download(n))
            tt = sum(x * ctx.synth_base_scalar(99, i) for i, x in enumerate(s)) % O.R_MOD
            ok = got == O.g1_mul(G, tt)
            print("msm closed-form", k, kind, ok, flush=True)
            assert ok
        for c in ([None] if k < 20 else [None, 16, 18]):
            if c: os.environ["H2B_MSM_C"] = str(c)
            else: os.environ.pop("H2B_MSM_C", None)
            best, med = timed(lambda: Bs.msm(sc, n=n), reps=3, warm=1)
            print(f"msm k={k} kind={kind} c={c}: best {best*1e3:.2f} ms -> {n/best/1e6:.1f} Mpts/s", flush=True)
            res[f"msm_{k}_{kind}_{c}_ms"] = best * 1e3
        os.environ.pop("H2B_MSM_C", None)
        sc.free()
    Bs.free()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
print("DONE")
