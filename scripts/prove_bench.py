"""keygen + create_proof of the reference's bench circuit (benches/plonk.rs) over KZG/bn256 on the GPU, with
per-stage timings, the proof checked by the restated reference verifier (oracle, test infrastructure).
Usage: python scripts/prove_bench.py [k] [reps]"""
import json
import os
import sys
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h
from tests import plonk_cases as PC

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scheme = sys.argv[3] if len(sys.argv) > 3 else "gwc"
lib = os.environ.get("H2B_LIB")
ctx = h.Context(0, lib_path=lib)
res = {"k": k, "multiopen": scheme}
t0 = time.perf_counter()
params = h.ParamsKZG.setup(ctx, k, PC.S_TOXIC, precompute=True)
res["setup_s"] = time.perf_counter() - t0
cs = PC.build_cs("bench")
fixed, advice, copies = PC.bench_circuit_limbs(k, 0xDEADBEEF)
t0 = time.perf_counter()
pk = h.keygen(params, cs, fixed, copies)
ctx.sync()
res["keygen_s"] = time.perf_counter() - t0
witness = lambda phase, ch: dict(enumerate(advice))  # noqa: E731
best = None
for rep in range(reps):
    timings = {}
    tr = h.Blake2bWrite()
    l0 = ctx.launches
    t0 = time.perf_counter()
    h.create_proof(params, pk, [witness], [[]], h.CounterRng(1234 + rep), tr, timings=timings,
                   prover=h.ProverSHPLONK if scheme == "shplonk" else h.ProverGWC)
    ctx.sync()
    dt = time.perf_counter() - t0
    if best is None or dt < best[0]:
        best = (dt, timings, ctx.launches - l0)
    proof = tr.finalize()
res["create_proof_s"] = best[0]
res["stages_s"] = {kk: round(v, 5) for kk, v in best[1].items()}
res["gpu_launches"] = best[2]
res["proof_bytes"] = len(proof)
from oracle import prover as OV
from oracle import bn256 as O
g0 = h.g1_decode(params.g.download()[:1])[0]
t0 = time.perf_counter()
res["verified"] = bool(OV.verify_proof(SimpleNamespace(g=[g0]), PC.S_TOXIC, PC.oracle_vk_of(pk), [[]], proof, multiopen=scheme))
res["verify_s"] = time.perf_counter() - t0
print(json.dumps(res))
