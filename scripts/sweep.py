"""Single-GPU sweep k = 16..26 (BASELINE.json configs[1], [2]): MSM Mpts/s (uniform and witness-like
scalars) and NTT Melem/s (best_fft, coeff_to_extended / extended_to_coeff with j = 5, 64-column batches
where they fit), device-resident inputs, CUDA-event-free wall timing around synchronous calls."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h

ctx = h.Context(0)
ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(16, 27, 2))
out = {"msm": {}, "ntt": {}}


def best(fn, reps=3):
    fn(); ctx.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ctx.sync(); ts.append(time.perf_counter() - t)
    return min(ts)


for k in ks:
    n = 1 << k
    bases = ctx.synth_bases(n, 0x6B7A67).precompute()
    row = {"window_bits": bases.table_window_bits}
    for kind, name in ((0, "uniform"), (1, "all_equal"), (2, "zero_one"), (3, "16bit"), (4, "90pct_zero")):
        sc = ctx.synth_scalars(n, 5, kind)
        t = best(lambda: bases.msm(sc, n=n))
        row[name] = {"ms": t * 1e3, "mpts_s": n / t / 1e6}
        sc.free()
    bases.free()
    out["msm"][k] = row
    print("msm", k, json.dumps(row), flush=True)
    dom = h.EvaluationDomain(ctx, 5, k) if k <= 26 else None
    omega = h.fr_encode([dom.constant("omega")])
    a = ctx.synth_scalars(n, 7, 0)
    r = {}
    t = best(lambda: ctx.best_fft_device(a, omega, k)); r["best_fft"] = {"ms": t * 1e3, "melem_s": n / t / 1e6}
    t = best(lambda: dom.lagrange_to_coeff_device(a)); r["lagrange_to_coeff"] = {"ms": t * 1e3, "melem_s": n / t / 1e6}
    ne = dom.extended_len()
    ext = ctx.alloc(ne * 32)
    t = best(lambda: dom.coeff_to_extended_device(a, ext)); r["coeff_to_extended"] = {"ms": t * 1e3, "melem_s_out": ne / t / 1e6}
    q = ctx.alloc(dom.quotient_len * 32)
    t = best(lambda: dom.extended_to_coeff_device(ext, q, divide_by_vanishing=True))
    r["extended_to_coeff"] = {"ms": t * 1e3, "melem_s_in": ne / t / 1e6}
    for x in (a, ext, q):
        x.free()
    # 64 columns at once where 64 * (n + 2^ek) * 32 B fits comfortably (<= 40 GiB)
    if 64 * (n + ne) * 32 <= (40 << 30):
        src = ctx.alloc(64 * n * 32); dst = ctx.alloc(64 * ne * 32)
        ctx._check(ctx.lib.h2b_synth_scalars(ctx.h, src.ptr, 64 * n, 9, 0))
        t = best(lambda: dom.coeff_to_extended_device(src, dst, 64), reps=2)
        r["coeff_to_extended_x64"] = {"ms": t * 1e3, "melem_s_out": 64 * ne / t / 1e6}
        t = best(lambda: dom.extended_to_coeff_device(dst, dst, 64, out_stride=ne), reps=2)
        r["extended_to_coeff_x64"] = {"ms": t * 1e3, "melem_s_in": 64 * ne / t / 1e6}
        src.free(); dst.free()
    dom.free()
    out["ntt"][k] = r
    print("ntt", k, json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_r1.json", "w"), indent=1)
ctx.close()
