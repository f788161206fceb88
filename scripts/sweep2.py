"""Round-2 sweep k = 16..26 on ONE GPU (BASELINE.json configs[1], [2]), every number with its roofline fraction.

  MSM   uniform + witness-like scalars: ms, Mpts/s, the accumulation kernel's share, fraction of the IMAD.WIDE peak
        (algorithmic n*W*11*136 as SURVEY.md 8d defines it, and the 1160 wide multiplies per mixed addition actually issued).
  NTT   best_fft / lagrange_to_coeff / coeff_to_extended / extended_to_coeff (j = 5): ms, Melem/s, per-pass ms, HBM GB/s
        per pass against the measured peak, integer fraction ((n/2) log2 n * 136 multiplies).
  x64   the 64-column batches of configs[2] at EVERY k: inputs device-resident (64 * n * 32 B, 128 GiB at k = 26), the
        extended outputs written through a bounded ring of column groups (a prover consumes the cosets group by
        group; 64 * 2^ek * 32 B would be 512 GiB at k = 26) -- and, where host memory allows, streamed from and to
        pinned host memory through the C ABI's double-buffered host path.

usage: python scripts/sweep2.py [k,k,...] [out.json]
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(16, 27, 2))
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/r2b_sweep_1gpu.json"
HOST_K_MAX = int(os.environ.get("H2B_SWEEP_HOST_K", "20"))
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]
out = {"msm": {}, "ntt": {}, "x64": {}, "peaks": {"hbm_gbs": HBM}}
# H2B_SWEEP_PARTS=msm re-measures the MSM rows only, on top of an earlier sweep file (H2B_SWEEP_BASE): the transform
# rows of that file stay valid as long as csrc/ntt.cu has not changed since
PARTS = os.environ.get("H2B_SWEEP_PARTS", "msm,ntt,x64").split(",")
if os.environ.get("H2B_SWEEP_BASE"):
    out = json.load(open(os.environ["H2B_SWEEP_BASE"]))
    out.setdefault("notes", []).append("rows of " + ",".join(PARTS) + " at k = " + ",".join(str(x) for x in ks) + " re-measured on top of " + os.environ["H2B_SWEEP_BASE"])


def best(ctx, fn, reps=3):
    fn(); ctx.sync(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ctx.sync(); ts.append(time.perf_counter() - t)
    return min(ts)


def flush():
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(out, open(out_path, "w"), indent=1)


ctx = h.Context(0)
imad = ctx.pipe_peak("imad_wide")[0]
out["peaks"]["imad_wide_mults_s"] = imad
out["peaks"]["mulmod_s"] = ctx.pipe_peak("fr_mul")[1]
ctx.close()

for k in ks:
    n = 1 << k
    if "msm" in PARTS:
        # ---------------- MSM (its own context: the workspace is released before the transforms) ----------------
        ctx = h.Context(0)
        bases = ctx.synth_bases(n, 0x6B7A67).precompute()
        c = bases.table_window_bits
        W = (255 + c - 1) // c
        row = {"window_bits": c, "windows": W}
        for kind, name in ((0, "uniform"), (1, "all_equal"), (2, "zero_one"), (3, "16bit"), (4, "90pct_zero")):
            sc = ctx.synth_scalars(n, 5, kind)
            t = best(ctx, lambda: bases.msm(sc, n=n))
            r = {"ms": t * 1e3, "mpts_s": n / t / 1e6}
            if kind == 0:
                ctx.set_profile(True)
                acc = []
                for _ in range(3):
                    bases.msm(sc, n=n)
                    acc.append(ctx.last_kernel_ms())
                ctx.set_profile(False)
                a_ms = sorted(acc)[1]
                r.update({"accum_kernel_ms": a_ms, "accum_share": a_ms / (t * 1e3),
                          "roofline_frac": n * W * 11 * 136 / (t) / imad,                     # whole MSM, algorithmic
                          "roofline_frac_accum_kernel": n * W * 11 * 136 / (a_ms * 1e-3) / imad,
                          "roofline_frac_executed_accum_kernel": n * W * 1160 / (a_ms * 1e-3) / imad,
                          "ec_adds_per_s": n * W / t})
            row[name] = r
            sc.free()
        bases.free()
        ctx.close()
        out["msm"][str(k)] = row
        out["msm"].pop(k, None)
        print("msm", k, json.dumps(row), flush=True)
        flush()
    if "ntt" not in PARTS:
        continue

    # ---------------- single transforms ----------------
    ctx = h.Context(0)
    dom = h.EvaluationDomain(ctx, 5, k)
    ek = dom.extended_k
    ne, nq = dom.extended_len(), dom.quotient_len
    omega = h.fr_encode([dom.constant("omega")])
    a = ctx.synth_scalars(n, 7, 0)
    ext = ctx.alloc(ne * 32)
    q = ctx.alloc(max(nq, 1) * 32)
    r = {"extended_k": ek}

    def timed(name, fn, size_log, elems, key):
        fn(); ctx.sync()
        t = best(ctx, fn)
        ctx.set_profile(True)
        ps = []
        for _ in range(3):
            fn()
            ps.append(ctx.last_ntt_pass_ms())
        ctx.set_profile(False)
        passes = [sorted(p[i] for p in ps)[1] for i in range(len(ps[0]))]
        m = 1 << size_log
        r[name] = {"ms": t * 1e3, key: elems / t / 1e6, "pass_ms": passes,
                   "hbm_GBps_per_pass": [64.0 * m / (x * 1e-3) / 1e9 for x in passes],
                   "hbm_frac_slowest_pass": 64.0 * m / (max(passes) * 1e-3) / 1e9 / HBM,
                   "int_frac": (m / 2) * size_log * 136 / (sum(passes) * 1e-3) / imad}

    timed("best_fft", lambda: ctx.best_fft_device(a, omega, k), k, n, "melem_s")
    timed("lagrange_to_coeff", lambda: dom.lagrange_to_coeff_device(a), k, n, "melem_s")
    timed("coeff_to_extended", lambda: dom.coeff_to_extended_device(a, ext), ek, ne, "melem_s_out")
    timed("extended_to_coeff", lambda: dom.extended_to_coeff_device(ext, q, divide_by_vanishing=True), ek, ne, "melem_s_in")
    for x in (a, ext, q):
        x.free()
    out["ntt"][str(k)] = r
    print("ntt", k, json.dumps(r), flush=True)
    flush()

    # ---------------- 64 columns, device-resident inputs, outputs through a ring of column groups ----------------
    x = {}
    try:
        ring_bytes = 16 << 30
        group = max(1, min(64, ring_bytes // (ne * 32)))
        src = ctx.alloc(64 * n * 32)
        ctx._check(ctx.lib.h2b_synth_scalars(ctx.h, src.ptr, 64 * n, 9, 0))
        dst = ctx.alloc(group * ne * 32)
        lib = ctx.lib

        def c2e():
            for c0 in range(0, 64, group):
                nc = min(group, 64 - c0)
                ctx._check(lib.h2b_coeff_to_extended_batch(dom.h, src.at(c0 * n * 32), n, dst.ptr, ne, h.H2B_DEVICE, nc))

        def e2c():  # the group's extended columns back to coefficients (division by the vanishing polynomial fused)
            for c0 in range(0, 64, group):
                nc = min(group, 64 - c0)
                ctx._check(lib.h2b_extended_to_coeff_batch(dom.h, dst.ptr, ne, dst.ptr, ne, h.H2B_DEVICE, nc, 1))

        t = best(ctx, c2e, reps=2)
        x["coeff_to_extended_x64"] = {"ms": t * 1e3, "melem_s_out": 64 * ne / t / 1e6, "columns_per_group": group,
                                      "int_frac": 64 * (ne / 2) * ek * 136 / t / imad,
                                      "hbm_GBps_algorithmic": 64 * (n + ne) * 32 / t / 1e9}
        t = best(ctx, e2c, reps=2)
        x["extended_to_coeff_x64"] = {"ms": t * 1e3, "melem_s_in": 64 * ne / t / 1e6, "columns_per_group": group,
                                      "int_frac": 64 * (ne / 2) * ek * 136 / t / imad}
        src.free(); dst.free()
    except Exception as e:  # noqa: BLE001
        x["device_error"] = str(e)[:300]
    # ---------------- 64 columns from and to pinned host memory through the C ABI (streamed groups) ----------------
    if k <= HOST_K_MAX:
        try:
            hin, hout = ctx.pinned((64 * n, 4)), ctx.pinned((64 * ne, 4))
            tmp = ctx.synth_scalars(64 * n, 11, 0)
            ctx._check(ctx.lib.h2b_copy_d2h(ctx.h, C.c_void_p(hin.ptr.value), tmp.ptr, 64 * n * 32))
            tmp.free()
            fn = lambda: ctx._check(ctx.lib.h2b_coeff_to_extended_batch(  # noqa: E731
                dom.h, C.c_void_p(hin.ptr.value), n, C.c_void_p(hout.ptr.value), ne, h.H2B_HOST, 64))
            t = best(ctx, fn, reps=2)
            x["coeff_to_extended_x64_host"] = {"ms": t * 1e3, "melem_s_out": 64 * ne / t / 1e6,
                                               "pcie_GBps": 64 * (n + ne) * 32 / t / 1e9,
                                               "h2d_bytes": 64 * n * 32, "d2h_bytes": 64 * ne * 32}
            hin.free(); hout.free()
        except Exception as e:  # noqa: BLE001
            x["host_error"] = str(e)[:300]
    dom.free()
    ctx.close()
    out["x64"][str(k)] = x
    print("x64", k, json.dumps(x), flush=True)
    flush()
print("wrote", out_path)
