#!/usr/bin/env python
"""Benchmark of the two hot paths at BASELINE.json's headline configuration:
bn256 MSM (Mpts/s) and NTT (Melem/s) at k = 24, on 1/2/4/8 B200s of one node.

A "step" is one pass of the hot path over one batch of synthetic input:
one best_multiexp over 2^24 uniformly random scalars (bases device-resident, as
ParamsKZG keeps them) followed by one best_fft of 2^24 elements.  With N > 1 the
MSM is the N*2^24-point MSM sharded by point range (per-rank partial points
all-gathered and folded, arithmetic.rs:139-153) and the NTTs are independent
columns, one per rank (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU
algorithm (the C++ restatement under oracle/: the reference is Rust and cannot
be compiled in this image) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_LOG = 24


def workload_config():
    """The workload both arms run, key for key (implementation details of the GPU arm go to `impl_config`)."""
    return {"workload": f"msm_k{K_LOG}+ntt_k{K_LOG}", "points_per_gpu": 1 << K_LOG, "scalars": "uniform in [0, r)",
            "l2": "inputs (1.5 GiB MSM, 0.5 GiB NTT per step) exceed the 126 MB L2 and every host cache; no flush needed"}


METRIC = "bn256 MSM Mpts/s (and NTT Melem/s, key `ntt`) at k=24"
SEED = 0x68616C6F32
# algorithmic work of one bucket addition = 11 modular multiplications
# (mixed Jacobian add, 7M+4S) of 136 32x32-bit multiplies each.
MULMODS_PER_ADD = 11
MULTS_PER_MULMOD = 136


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle/ref_cpu.cpp)
# ---------------------------------------------------------------------------
# Same configuration as the GPU arm (BASELINE.json: k = 24): one pass is ~10-25 s of host time.
CPU_MSM_LOG, CPU_NTT_LOG = K_LOG, K_LOG
CPU_ARM_BUDGET_S = 150.0  # the reference arm stops starting new passes after this much timed work


def cpu_pass(oc, H, O, scalars, bases, ntt_in, omega, threads):
    t0 = time.perf_counter()
    oc.lib.oracle_best_multiexp(scalars.ctypes.data, bases.ctypes.data, scalars.shape[0], threads,
                                _OUT.ctypes.data)
    t1 = time.perf_counter()
    oc.lib.oracle_best_fft(ntt_in.ctypes.data, omega.ctypes.data, CPU_NTT_LOG, threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


_OUT = None


def cpu_setup():
    global _OUT
    import numpy as np
    from oracle import bn256 as O
    from tests import helpers as H
    oc = H.load_oracle_c()
    threads = host_threads()
    _OUT = np.zeros(8, dtype=np.uint64)
    bases = oc.synth_bases(1 << CPU_MSM_LOG, threads=threads)
    scalars = H.rand_fr_limbs(SEED & 0xFFFF, 1 << CPU_MSM_LOG)
    ntt_in = H.rand_fr_limbs(7, 1 << CPU_NTT_LOG)
    omega = H.fr_enc([O.omega_for(CPU_NTT_LOG)])[0]
    return oc, H, O, scalars, bases, ntt_in, omega, threads


def cpu_baseline(steps: int = 1, warmup: int = 0):
    """The reference's CPU algorithm at the SAME size as the GPU arm (k = 24 MSM + k = 24 NTT per pass).  A pass
    is ~15 s on 16 cores, so the number of passes is bounded by CPU_ARM_BUDGET_S of timed work (at least one
    full pass, never a smaller problem); `passes` says how many ran."""
    oc, H, O, scalars, bases, ntt_in, omega, threads = cpu_setup()
    for _ in range(min(warmup, 1)):  # one pass touches every page; more warm-up changes nothing on a CPU
        cpu_pass(oc, H, O, scalars, bases, ntt_in.copy(), omega, threads)
    tm = tn = 0.0
    done = 0
    for _ in range(max(steps, 1)):
        a, b = cpu_pass(oc, H, O, scalars, bases, ntt_in, omega, threads)
        tm += a
        tn += b
        done += 1
        if tm + tn > CPU_ARM_BUDGET_S:
            break
    sample = (f"oracle/ref_cpu.cpp (C++ restatement of arithmetic.rs, std::thread for rayon): best_multiexp on "
              f"2^{CPU_MSM_LOG} points + best_fft at k={CPU_NTT_LOG} -- the GPU arm's full size -- {done} pass(es), "
              f"{threads} threads (RAYON_NUM_THREADS equivalent), nproc={os.cpu_count()}")
    return {"value": (1 << CPU_MSM_LOG) * done / tm / 1e6, "unit": "Mpts/s", "cores": threads, "kind": "port",
            "sample": sample, "ntt": {"value": (1 << CPU_NTT_LOG) * done / tn / 1e6, "unit": "Melem/s"},
            "ms_per_pass": (tm + tn) / done * 1e3, "msm_ms_per_pass": tm / done * 1e3, "passes": done}


OPMIX_K = 20
OPMIX = ("create_proof op-mix replay of benches/plonk.rs at k=20 (SURVEY.md 3.1): 11 commits (3 constant advice "
         "columns + 8 uniform), 4 lagrange_to_coeff (2^20), 4 coeff_to_extended (2^22), 1 divide_by_vanishing + "
         "extended_to_coeff (2^22); witness synthesis, evaluate_h, eval_polynomial and multiopen algebra NOT included")


def cpu_opmix():
    """The same op mix on the host cores with the C++ restatement of the reference algorithms."""
    import numpy as np
    from tests import helpers as H
    oc = H.load_oracle_c()
    threads = host_threads()
    n = 1 << OPMIX_K
    bases = oc.synth_bases(n, threads=threads)
    uni = H.rand_fr_limbs(3, n)
    const = np.tile(H.rand_fr_limbs(4, 1), (n, 1))
    d = oc.domain(5, OPMIX_K, threads)
    out = np.zeros(8, dtype=np.uint64)
    ext = np.zeros((1 << d.extended_k, 4), dtype=np.uint64)
    t0 = time.perf_counter()
    for i in range(11):
        sc = const if i < 3 else uni
        oc.lib.oracle_best_multiexp(sc.ctypes.data, bases.ctypes.data, n, threads, out.ctypes.data)
    t1 = time.perf_counter()
    a = uni.copy()
    for _ in range(4):
        oc.lib.oracle_lagrange_to_coeff(d.h, a.ctypes.data)
    for _ in range(4):
        oc.lib.oracle_coeff_to_extended(d.h, a.ctypes.data, ext.ctypes.data)
    oc.lib.oracle_divide_by_vanishing_poly(d.h, ext.ctypes.data)
    oc.lib.oracle_extended_to_coeff(d.h, ext.ctypes.data)
    t2 = time.perf_counter()
    d.free()
    return {"seconds": t2 - t0, "msm_seconds": t1 - t0, "ntt_seconds": t2 - t1, "cores": threads, "kind": "port"}


def gpu_opmix(ctx, h):
    """The op mix through the drop-in entry points with HOST (pinned) polynomials, as create_proof holds them."""
    import ctypes as C
    k, n = OPMIX_K, 1 << OPMIX_K
    bases = ctx.synth_bases(n, 0xABCD)
    bases.precompute()
    dom = h.EvaluationDomain(ctx, 5, k)
    ne, nq = dom.extended_len(), dom.quotient_len
    uni_d = ctx.synth_scalars(n, 31, 0)
    const_d = ctx.synth_scalars(n, 32, 1)
    ext_d = ctx.synth_scalars(ne, 33, 0)
    uni, const, ext, extout = ctx.pinned((n, 4)), ctx.pinned((n, 4)), ctx.pinned((ne, 4)), ctx.pinned((ne, 4))
    for src, dst, cnt in ((uni_d, uni, n), (const_d, const, n), (ext_d, ext, ne)):
        ctx._check(ctx.lib.h2b_copy_d2h(ctx.h, C.c_void_p(dst.ptr.value), src.ptr, cnt * 32))

    def one(host):
        t0 = time.perf_counter()
        for i in range(11):
            if host:
                bases.msm(const.array if i < 3 else uni.array)
            else:
                bases.msm(const_d if i < 3 else uni_d, n=n)
        ctx.sync()
        t1 = time.perf_counter()
        for _ in range(4):
            if host:
                ctx._check(ctx.lib.h2b_lagrange_to_coeff(dom.h, C.c_void_p(uni.ptr.value), h.H2B_HOST))
            else:
                dom.lagrange_to_coeff_device(uni_d)
        for _ in range(4):
            if host:
                ctx._check(ctx.lib.h2b_coeff_to_extended(dom.h, C.c_void_p(uni.ptr.value), C.c_void_p(extout.ptr.value),
                                                         h.H2B_HOST))
            else:
                dom.coeff_to_extended_device(uni_d, ext_d)
        if host:
            ctx._check(ctx.lib.h2b_extended_to_coeff(dom.h, C.c_void_p(ext.ptr.value), C.c_void_p(extout.ptr.value),
                                                     h.H2B_HOST, 1))
        else:
            dom.extended_to_coeff_device(ext_d, ext_d, divide_by_vanishing=True)
        ctx.sync()
        t2 = time.perf_counter()
        return t2 - t0, t1 - t0, t2 - t1

    one(True)
    res = {}
    for host in (True, False):
        best = min(one(host) for _ in range(2))
        res["host_buffers" if host else "device_resident"] = {"seconds": best[0], "msm_seconds": best[1],
                                                              "ntt_seconds": best[2]}
    for x in (uni, const, ext, extout):
        x.free()
    for x in (uni_d, const_d, ext_d):
        x.free()
    dom.free()
    bases.free()
    return res


PROOF_K = 20
EVALH_K, EVALH_CPU_K = 20, 20


def evalh_case(h, ctx_or_none, k, np):
    """Random extended-domain columns for the bench circuit's evaluate_h at k (device buffers or host arrays)."""
    from halo2_pse_b200 import circuits
    cs = circuits.standard_plonk_cs()
    ek = k + 2  # degree 5
    names = ["fixed"] * cs.num_fixed_columns + ["advice"] * cs.num_advice_columns + ["l0", "l_last", "l_act"] + \
            ["sigma"] * len(cs.permutation.columns) + ["z"]
    return cs, ek, names


def gpu_evaluate_h(ctx, h):
    """Evaluator::evaluate_h kernels (custom gates + permutation argument) of the bench circuit on 2^22 rows."""
    import ctypes as C
    import numpy as np
    from halo2_pse_b200.plonk import _ptr_array
    cs, ek, names = evalh_case(h, ctx, EVALH_K, np)
    dom = h.EvaluationDomain(ctx, cs.degree(), EVALH_K)
    ext = dom.extended_len()
    bufs = [ctx.synth_scalars(ext, seed=4000 + i) for i in range(len(names))]
    fixed, advice = bufs[:4], bufs[4:7]
    l0, l_last, l_act = bufs[7:10]
    sigma, zs = bufs[10:13], bufs[13:14]
    values = ctx.alloc(ext * 32)
    ctx.memset(values, 0)
    ev = h.Evaluator(cs)
    cols, keep = h.make_eval_columns(fixed, advice, [], [], 11, 13, 17, 19)
    g = ev.custom_gates.compile(ctx)
    pc = cs.permutation.columns
    ctype = np.asarray([c.column_type for c in pc], dtype=np.uint32)
    cidx = np.asarray([c.index for c in pc], dtype=np.uint32)
    sg, zz = _ptr_array(sigma), _ptr_array(zs)
    ctx.set_profile(True)
    tg, tp = [], []
    for _ in range(5):
        ctx._check(ctx.lib.h2b_evaluate_h_gates(dom.h, g, C.byref(cols), values.ptr))
        tg.append(ctx.last_kernel_ms())
        ctx._check(ctx.lib.h2b_evaluate_h_permutation(dom.h, C.byref(cols), C.c_void_p(ctype.ctypes.data),
                                                      C.c_void_p(cidx.ctypes.data), len(pc), sg, zz, 1, cs.degree() - 2,
                                                      cs.blinding_factors(), l0.ptr, l_last.ptr, l_act.ptr, values.ptr))
        tp.append(ctx.last_kernel_ms())
    ctx.set_profile(False)
    gm, pm = sorted(tg)[2], sorted(tp)[2]
    graph_info = {"instructions": int(ctx.lib.h2b_graph_num_instructions(g)), "slots": int(ctx.lib.h2b_graph_num_slots(g))}
    for b in bufs + [values]:
        b.free()
    ev.free()
    dom.free()
    mulmods = 6 * ext, 21 * ext  # field multiplications per row: gate graph, permutation constraints
    return {"what": f"evaluate_h of benches/plonk.rs MyCircuit at k={EVALH_K}: 2^{ek} extended rows, device-resident cosets",
            "gates_ms": gm, "permutation_ms": pm, "mrows_per_s": ext / ((gm + pm) * 1e-3) / 1e6,
            "gates_GBps_algorithmic": 9 * 32 * ext / (gm * 1e-3) / 1e9,
            "permutation_Gmulmod_s": mulmods[1] / (pm * 1e-3) / 1e9,
            "graph": graph_info}


def cpu_evaluate_h():
    """The same on the host cores: oracle/ref_cpu.cpp's restatement of evaluation.rs:280-444 (bounded sample)."""
    import numpy as np
    import halo2_pse_b200 as h
    from tests import helpers as H
    oc = H.load_oracle_c()
    threads = host_threads()
    cs, ek, names = evalh_case(h, None, EVALH_CPU_K, np)
    dom = oc.domain(cs.degree(), EVALH_CPU_K, threads)
    ext = 1 << ek
    arrs = [H.rand_fr_limbs(900 + i, ext) for i in range(len(names))]
    ev = h.Evaluator(cs)  # the calculation list (plain data)
    gg = ev.custom_gates
    graph = oc.graph(gg.encode(), H.fr_enc(gg.constants), gg.rotations, gg.num_intermediates)
    values = np.zeros((ext, 4), dtype=np.uint64)
    t0 = time.perf_counter()
    oc.evaluate_h(dom, graph, arrs[:4], arrs[4:7], [], [], 11, 13, 17, 19, [tuple(c) for c in cs.permutation.columns],
                  arrs[10:13], arrs[13:14], cs.degree() - 2, cs.blinding_factors(), arrs[7], arrs[8], arrs[9], [], [],
                  values, threads)
    dt = time.perf_counter() - t0
    oc.lib.oracle_graph_free(graph)
    dom.free()
    return {"seconds": dt, "mrows_per_s": ext / dt / 1e6, "cores": threads, "kind": "port",
            "sample": f"bench circuit at k={EVALH_CPU_K} (2^{ek} rows), gates + permutation argument"}


def gpu_create_proof(ctx, h, world=1, barrier=None, max_over_ranks=None):
    """keygen + create_proof (KZG/bn256, GWC, Blake2b) of the reference's bench circuit at k = 20 with every
    commitment, transform, quotient evaluation and opening on the GPU; the witness starts in host memory.
    world > 1: ONE proof on all GPUs -- the base vectors are sharded by point range (dist.ShardedBases), every
    commitment runs on every GPU at once, the rest of the prover is replicated (the Fiat-Shamir chain does not
    shard); proof bytes are compared with the single-GPU prover's on every rank."""
    import hashlib
    from halo2_pse_b200 import circuits
    k = PROOF_K
    t0 = time.perf_counter()
    params = h.ParamsKZG.setup(ctx, k, 0x1234567890ABCDEF1234567890ABCDEF, precompute=(world == 1))
    t_setup = time.perf_counter() - t0
    cs = circuits.standard_plonk_cs()
    fixed, advice, copies = circuits.my_circuit(k, 0xDEADBEEF)
    witness = lambda phase, ch: dict(enumerate(advice))  # noqa: E731
    single_digest = None
    if world > 1:
        from halo2_pse_b200 import dist as D
        # the single-GPU proof of the same circuit and rng seed, on this rank alone, as the check value
        pk1 = h.keygen(params, cs, fixed, copies)
        tr = h.Blake2bWrite()
        h.create_proof(params, pk1, [witness], [[]], h.CounterRng(1234 + 3), tr)
        single_digest = hashlib.sha256(tr.finalize()).hexdigest()
        pk1.free()
        params = D.shard_params(params)  # keeps this rank's range (+ window table), frees the full vectors
        barrier()
    t0 = time.perf_counter()
    pk = h.keygen(params, cs, fixed, copies)
    ctx.sync()
    t_keygen = time.perf_counter() - t0
    # the circuit description holds millions of Python objects (2^20 copy constraints as tuples): park them in the
    # permanent generation so that a generational collection inside the timed region does not walk them
    import gc
    gc.collect()
    gc.freeze()
    best = None
    digest = None
    for rep in range(4):
        timings = {}
        tr = h.Blake2bWrite()
        l0 = ctx.launches
        if barrier:
            barrier()
        t0 = time.perf_counter()
        h.create_proof(params, pk, [witness], [[]], h.CounterRng(1234 + rep), tr, timings=timings)
        ctx.sync()
        dt = time.perf_counter() - t0
        if max_over_ranks:
            (dt,) = max_over_ranks(dt)
        proof_bytes = tr.finalize()
        digest = hashlib.sha256(proof_bytes).hexdigest()
        if rep and (best is None or dt < best[0]):
            best = (dt, timings, ctx.launches - l0, len(proof_bytes))
    pk.free()
    params.g.free()
    params.g_lagrange.free()
    out = {"what": f"create_proof of benches/plonk.rs MyCircuit at k={k} over KZG/bn256 (ProverGWC, Blake2bWrite, "
                   "Challenge255): 11 MSMs, 7 iNTT 2^20, 4 coset NTT 2^22, evaluate_h on 2^22 rows, 1 inverse coset "
                   "NTT, 17 Horner evaluations, 2 Kate divisions; witness columns start in host memory; proof bytes "
                   "are checked against the big-integer oracle and the restated verifier in tests/ (k = 5, 6, 14)",
           "seconds": best[0], "stages_seconds": {kk: round(v, 5) for kk, v in best[1].items()},
           "gpu_launches": best[2], "proof_bytes": best[3], "keygen_seconds": t_keygen,
           "setup_seconds_untimed_host_bookkeeping": t_setup, "wall_clock": "host perf_counter, best of 3 after 1 warm-up"}
    if world > 1:
        same = 0.0 if digest == single_digest else 1.0
        (bad,) = max_over_ranks(same)
        out["n_gpus"] = world
        out["sharding"] = ("ONE proof on all GPUs: g / g_lagrange sharded by contiguous point range (each rank holds "
                           "n/N bases + its window table), every commitment = N local MSMs + all-gather of 64-byte "
                           "partial points; transforms, quotient evaluation and openings replicated; seconds = max over ranks")
        out["proof_bytes_equal_single_gpu_prover_on_every_rank"] = bad == 0.0
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "Mpts/s", "n_gpus": args.gpus,
        "steps": base["passes"], "steps_requested": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": base["ms_per_pass"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
        "config": workload_config(),
        "impl_config": {"note": f"same size as the GPU arm: every step is one full 2^{CPU_MSM_LOG}-point best_multiexp + one "
                                f"2^{CPU_NTT_LOG}-point best_fft on the host cores; passes bounded by "
                                f"{CPU_ARM_BUDGET_S:.0f} s of timed work, never a smaller problem"},
        "ntt": base["ntt"],
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import numpy as np
    import torch

    import halo2_pse_b200 as h
    from halo2_pse_b200 import dist as D

    # stdout carries exactly ONE JSON line: route everything else (NCCL prints its version banner
    # to fd 1 from C) to stderr and keep the real stdout for the final line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    ctx = h.Context(local)
    k, n = K_LOG, 1 << K_LOG
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    # ---- synthetic inputs, generated on the device ------------------------------
    bases = ctx.synth_bases(n, 0x6B7A67 + rank)   # this rank's point range
    if os.environ.get("H2B_NO_TABLE") != "1":
        bases.precompute()  # one-time, like the upload: ParamsKZG's bases are immutable (untimed)
    scalars = ctx.synth_scalars(n, SEED + rank, 0)
    poly = ctx.synth_scalars(n, SEED + 1000 + rank, 0)
    dom = h.EvaluationDomain(ctx, 2, k)
    omega = h.fr_encode([dom.constant("omega")])  # the primitive 2^k-th root every caller passes
    msm = D.ShardedMSM(ctx, bases)
    ctx.sync()

    # sanity of the sharded path on this very process group (cheap, untimed): the first 2^12 points of
    # every rank against the closed form  sum_i c_i [h_i] G = [sum_i c_i h_i] G  over ALL ranks' ranges
    m = 1 << 12
    got = msm.msm(scalars, m)
    cs = h.fr_decode(scalars.download(m))
    part = sum(c * ctx.synth_base_scalar(0x6B7A67 + rank, i) for i, c in enumerate(cs)) % h.R_MOD
    if world > 1:
        parts = [None] * world
        torch.distributed.all_gather_object(parts, part)
        part = sum(parts) % h.R_MOD
    gen = h.Bases(ctx, h.g1_encode([(1, 2)]), 1)
    want = gen.msm(h.fr_encode([part]))  # [t] G through the library's own scalar multiplication
    gen.free()
    sharded_ok = got == want

    def step_device():
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e2 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        msm.msm(scalars, n)
        e1.record(stream)
        ctx.best_fft_device(poly, omega, k)
        e2.record(stream)
        return e0, e1, e2

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launches
    evs = []
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        evs.append(step_device())
    ctx.sync()
    torch.cuda.synchronize()
    t_msm = sum(a.elapsed_time(b) for a, b, _ in evs)
    t_ntt = sum(b.elapsed_time(c) for _, b, c in evs)
    t_all = evs[0][0].elapsed_time(evs[-1][2])
    barrier()
    t_wall = (time.perf_counter() - t_wall) * 1e3
    launches = ctx.launches - launches0
    clocks = sampler.stop() if sampler else None

    def max_over_ranks(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device=torch.device("cuda", local))
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return t.tolist()

    t_msm, t_ntt, t_all = max_over_ranks(t_msm, t_ntt, t_all)

    # ---- roofline of the dominant kernels (live, CUDA events on the library stream) ----
    ctx.set_profile(True)
    acc_ms, pass_ms = [], []
    for _ in range(3):
        msm.msm(scalars, n)
        acc_ms.append(ctx.last_kernel_ms())
        ctx.best_fft_device(poly, omega, k)
        pass_ms.append(ctx.last_ntt_pass_ms())
    ctx.set_profile(False)
    acc_ms = sorted(acc_ms)[1]
    pass_ms = [sorted(p[i] for p in pass_ms)[1] for i in range(len(pass_ms[0]))]
    peaks, peak_src = measured_peaks()
    c_win = bases.table_window_bits or int(ctx.lib.h2b_msm_window_bits(n))
    windows = (255 + c_win - 1) // c_win
    adds = n * windows
    mults = adds * MULMODS_PER_ADD * MULTS_PER_MULMOD
    pipe = {name: ctx.pipe_peak(name) for name in ("imad", "imad_wide", "carry_chain", "fr_mul")}
    # The integer roofline: 32x32->64 multiplies (IMAD.WIDE.U32), measured live.  Round 1 quoted 18.5 T/s here, but
    # that benchmark multiplied two loop-invariant registers, ptxas hoisted the product and the loop timed 64-bit
    # ADDS.  With operands that change every iteration (SASS checked: profiles/r2_pipe_rates_sass.txt) a wide
    # multiply issues every 4 cycles per SM sub-partition -- 32 lanes/clk/SM = 9.3 T/s at 1965 MHz, with or
    # without an addend or a carry; only the 32-bit IMAD runs at 64 lanes/clk/SM (scripts/mulbench/pipes.cu).
    mult_peak = pipe["imad_wide"][0]
    # what the kernel issues per mixed XYZZ addition (8M + 2S): 6 products x 128, 2 dedicated squarings x 100, and the two
    # products of y3 on ONE shared Montgomery reduction (64 + 64 + 64) = 1160 IMAD.WIDE (round 1 / early round 2: 10 x 130)
    executed = adds * 1160
    roofline = {
        "kernel": "msm_accum0_kernel (bucket accumulation, level 0)", "bound": "int32-multiply (IMAD.WIDE pipe)",
        "achieved": mults / (acc_ms * 1e-3) / 1e12, "peak": mult_peak / 1e12, "unit": "Tmul/s",
        "frac": mults / (acc_ms * 1e-3) / mult_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at k=24, c=22 (ncu, profiles/r2b_accum0_dram_k24.csv:
        # 15.91 GB + 0.52 GB with 64-byte L2 fills for the table gathers; 28.65 + 0.52 GB before)
        "traffic": 16.43e9 if (k == 24 and c_win == 22) else None,
        "frac_executed": executed / (acc_ms * 1e-3) / mult_peak,
        "executed": "n*W*1160 IMAD.WIDE actually issued per XYZZ mixed addition (6 products x 128, 2 squarings x 100, "
                    "y3's two products on one shared reduction: 192); `frac` uses the survey's algorithmic 11*136 as "
                    "the contract asks, so it can exceed 1",
        "mulmod_ceiling_frac": (adds * 10 / (acc_ms * 1e-3)) / (pipe["fr_mul"][1] or 1.0),
        "peak_source": "live register-only microbenchmark h2b_pipe_peak: IMAD.WIDE.U32 (32x32->64) with operands "
                       "that change every iteration, 32 lanes/clk/SM; round 1's 18.5 T/s was a 64-bit-add loop",
        "algorithmic": f"n*W*{MULMODS_PER_ADD}*{MULTS_PER_MULMOD} 32x32 multiplies, n=2^{k}, c={c_win}, W={windows}",
        "kernel_ms": acc_ms, "ec_adds_per_s": adds / (acc_ms * 1e-3),
        "fr_mul_microbench_Tmul_s": pipe["fr_mul"][0] / 1e12,
        "imad_wide_Tmul_s": pipe["imad_wide"][0] / 1e12,
        "imad_wide_carry_chain_Tmul_s": pipe["carry_chain"][0] / 1e12,
        "imad_32bit_Tmul_s": pipe["imad"][0] / 1e12,
    }
    slow = max(pass_ms)
    roofline_ntt = {
        "kernel": "ntt_pass_fast<8> (one radix-256 pass over HBM)", "bound": "hbm",
        "achieved": 64.0 * n / (slow * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": 64.0 * n / (slow * 1e-3) / 1e9 / peaks["hbm_gbs"],
        # same capture, slowest pass (the first): 1 074 MB read (512 MiB of input + the 512 MiB twiddle stream) +
        # 505 MB written at k=24; the other passes move the algorithmic 2 x 512 MiB (538 + 479 MB)
        "traffic": 1.579e9 if k == 24 else None, "peak_source": peak_src,
        "algorithmic": f"64 B per element per pass (one read + one write), n=2^{k}", "pass_ms": pass_ms,
        "int_Tmul_s": (n / 2) * k * MULTS_PER_MULMOD / (sum(pass_ms) * 1e-3) / 1e12,
        # the binding roofline of a 256-bit NTT (max of the two): (n/2) log2 n butterflies x 136 multiplies
        "int_frac": (n / 2) * k * MULTS_PER_MULMOD / (sum(pass_ms) * 1e-3) / mult_peak,
        "hbm_GBps_per_pass": [64.0 * n / (t * 1e-3) / 1e9 for t in pass_ms],
        "transform_ms": sum(pass_ms),
    }

    # ---- end to end through the C ABI with HOST buffers ---------------------------
    h_sc = ctx.pinned((n, 4))
    h_poly = ctx.pinned((n, 4))
    ctx._check(ctx.lib.h2b_copy_d2h(ctx.h, C.c_void_p(h_sc.ptr.value), scalars.ptr, n * 32))
    ctx._check(ctx.lib.h2b_copy_d2h(ctx.h, C.c_void_p(h_poly.ptr.value), poly.ptr, n * 32))
    out = np.zeros(8, dtype=np.uint64)

    def step_host():
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e2 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        msm.msm(h_sc.array)                    # H2D of the scalars inside, 64 B result back
        e1.record(stream)
        ctx.best_fft(h_poly.array, omega, k)   # H2D + in-place transform + D2H inside
        e2.record(stream)
        return e0, e1, e2

    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    evs = [step_host() for _ in range(e2e_steps)]
    ctx.sync()
    torch.cuda.synchronize()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e_msm = sum(a.elapsed_time(b) for a, b, _ in evs)
    e_ntt = sum(b.elapsed_time(c) for _, b, c in evs)
    e_msm, e_ntt, e2e_wall = max_over_ranks(e_msm, e_ntt, e2e_wall)

    # ---- what the host side can deliver: pinned copies on EVERY rank at the same time ----------------------
    # (the end-to-end number above is bounded by these; on this pool the 8-GPU box is one VM with a single
    # NUMA node and 32 vCPUs, so there is no per-socket placement to do -- the probe shows whether the PCIe /
    # host-memory path scales with the number of GPUs)
    def copy_probe():
        nbytes = n * 32
        res = {}
        for name, fn in (("h2d", lambda: ctx.lib.h2b_copy_h2d(ctx.h, poly.ptr, C.c_void_p(h_poly.ptr.value), nbytes)),
                         ("d2h", lambda: ctx.lib.h2b_copy_d2h(ctx.h, C.c_void_p(h_poly.ptr.value), poly.ptr, nbytes))):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                ctx._check(fn())
            dt = (time.perf_counter() - t0) / 3
            barrier()
            mine = nbytes / dt / 1e9
            (slowest,) = max_over_ranks(-mine)
            (fastest,) = max_over_ranks(mine)
            res[name] = {"per_rank_GBps_min": -slowest, "per_rank_GBps_max": fastest,
                         "aggregate_GBps_at_the_slowest_rank": -slowest * world}
        res["what"] = f"{nbytes >> 20} MiB pinned copies issued by all {world} rank(s) at once, 3 repetitions"
        return res

    pcie = copy_probe()

    # ---- four-step NTT with all-to-all (configs[4]) when sharded -------------------
    four = None
    if world > 1:
        k4 = 26
        loc = (1 << k4) // world
        dev = torch.device("cuda", local)
        # Correctness at the size that is timed: every rank generates the same full 2^26 vector (2 GiB),
        # transforms it locally with the single-GPU kernel and compares its own slice of the four-step
        # result, for each exchange method, bit for bit on the device.
        w4 = h.EvaluationDomain(ctx, 2, k4).constant("omega")
        full = ctx.synth_scalars(1 << k4, SEED + 77, 0)
        ctx.sync()  # generated on the library's stream; the clone below runs on torch's
        full_t = D._as_tensor(full, (1 << k4) * 4, dev)
        orig = full_t[rank * loc * 4:(rank + 1) * loc * 4].clone()
        torch.cuda.synchronize()
        ctx.best_fft_device(full, h.fr_encode([w4]), k4)  # also builds the k = 26 twiddle tables
        ctx.sync()
        want = full_t[rank * loc * 4:(rank + 1) * loc * 4].clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.best_fft_device(full, h.fr_encode([w4]), k4)  # timed: same kernels on the (already transformed) vector
        ctx.sync()
        single_ms = (time.perf_counter() - t0) * 1e3
        del full_t
        full.free()
        buf = torch.empty(loc * 4, dtype=torch.int64, device=dev)
        four = {"k": k4, "single_gpu_ms_same_run": single_ms}
        for label, p2p in (("p2p", None), ("nccl_all_to_all", False)):
            fs = D.FourStepNTT(ctx, k4, w4, p2p=p2p)
            if label == "p2p" and not fs.p2p:
                four["p2p_unavailable"] = fs.p2p_error
                continue
            buf.copy_(orig)
            torch.cuda.synchronize()
            res = fs.run(buf)
            torch.cuda.synchronize()
            (bad,) = max_over_ranks(0.0 if torch.equal(res, want) else 1.0)
            barrier()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                fs.run(buf)
                barrier()
                ts.append((time.perf_counter() - t0) * 1e3)
            (best,) = max_over_ranks(min(ts))
            four[label] = {"ms": best, "melem_s": (1 << k4) / (best * 1e-3) / 1e6,
                           "verified_vs_single_gpu_k26_every_rank": bad == 0.0,
                           "nvlink_bytes_per_transform": getattr(fs, "nvlink_bytes", None),
                           "exchanges": getattr(fs, "exchanges", 3)}
            del fs
        del want, orig
        four["method"] = ("four-step; p2p = transposes fused with the exchange as direct stores into NVLink-mapped peer "
                          "buffers (symmetric memory), device-side barriers; nccl_all_to_all = tile transposes + "
                          "all_to_all_single + permute; wall clock, max over ranks; every method's k=26 result is "
                          "compared on every rank with the single-GPU h2b_best_fft of the same vector")
        best = four.get("p2p", four.get("nccl_all_to_all"))
        four["ms"], four["melem_s"] = best["ms"], best["melem_s"]

    # ---- ONE k = 26 MSM sharded by point range over all ranks (configs[1], strong scaling) ----
    strong = None
    if True:
        ks = 26
        ns = (1 << ks) // world
        sb = ctx.synth_bases(ns, 0x51 + rank)
        sb.precompute()
        ssc = ctx.synth_scalars(ns, SEED + 900 + rank, 0)
        smsm = D.ShardedMSM(ctx, sb)
        smsm.msm(ssc, ns)
        barrier()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            smsm.msm(ssc, ns)
            e1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            barrier()
        (best,) = max_over_ranks(min(ts))
        # no hard-coded single-GPU time: the N = 1 run of this same bench reports this key for one GPU
        strong = {"k": ks, "points_per_gpu": ns, "ms": best, "mpts_s": (1 << ks) / (best * 1e-3) / 1e6,
                  "window_bits": sb.table_window_bits}
        ssc.free()
        sb.free()

    opmix = gpu_opmix(ctx, h) if (rank == 0 and world == 1) else None
    proof = evalh = None
    if rank == 0 and world == 1:
        pctx = h.Context(local)  # its own context: scratch and MSM workspace sized for k = 20, not for the k = 24 runs
        proof = gpu_create_proof(pctx, h)
        evalh = gpu_evaluate_h(pctx, h)
        pctx.close()
    elif world > 1 and not os.environ.get("H2B_BENCH_NO_PROOF"):
        pctx = h.Context(local)
        proof = gpu_create_proof(pctx, h, world, barrier, max_over_ranks)
        pctx.close()
    if rank == 0:
        total_pts = world * n * args.steps
        line = {
            "metric": METRIC, "value": total_pts / (t_msm * 1e-3) / 1e6, "unit": "Mpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
            "config": workload_config(),
            "impl_config": {"msm_window_bits": c_win, "msm_window_table": bool(bases.table_window_bits),
                            "sharding": "MSM by point range, NTT by column" if world > 1 else "none",
                            "sanitizer": "compute-sanitizer is closed on this pool: memory safety rests on the "
                                         "bit-exact parity tests at every shape and the CPU emulator suite"},
            "ntt": {"value": world * n * args.steps / (t_ntt * 1e-3) / 1e6, "unit": "Melem/s",
                    "ms": t_ntt / args.steps},
            "msm_ms": t_msm / args.steps, "wall_ms_per_step": t_wall / args.steps,
            "roofline": roofline, "roofline_ntt": roofline_ntt,
            "e2e": {"value": world * n * e2e_steps / (e_msm * 1e-3) / 1e6, "unit": "Mpts/s",
                    "h2d_bytes_per_step": 2 * n * 32, "d2h_bytes_per_step": n * 32 + 64,
                    "ntt": {"value": world * n * e2e_steps / (e_ntt * 1e-3) / 1e6, "unit": "Melem/s"},
                    "ms_per_step": e2e_wall / e2e_steps, "msm_ms": e_msm / e2e_steps, "ntt_ms": e_ntt / e2e_steps,
                    "concurrent_pinned_copy_probe": pcie,
                    "path": "h2b_msm_affine + h2b_best_fft with H2B_HOST pointers (pinned), copies inside the timed region"},
            "gpu_launches": launches, "clocks": clocks, "sharded_msm_closed_form_check": bool(sharded_ok),
        }
        if four:
            line["four_step_ntt"] = four
        if strong:
            line["msm_k26_sharded"] = strong
        if world > 1 and proof:
            line["create_proof"] = proof
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(steps=1)
            line["create_proof_opmix"] = {"what": OPMIX, "gpu": opmix, "cpu": cpu_opmix()}
            line["create_proof"] = proof
            line["evaluate_h"] = {"gpu": evalh, "cpu": cpu_evaluate_h()}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    barrier()
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
