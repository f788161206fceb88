"""Times the quotient-evaluation kernels on device-resident random cosets (k given, bench circuit of
benches/plonk.rs): gates graph and permutation constraints.  Usage: python scripts/evalh_bench.py [k]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import halo2_pse_b200 as h
from tests import plonk_cases as PC

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
variant = sys.argv[2] if len(sys.argv) > 2 else "bench"
ctx = h.Context(0)
cs = PC.build_cs(variant)
dom = h.EvaluationDomain(ctx, cs.degree(), k)
ext = dom.extended_len()
ev = h.Evaluator(cs)
rnd = lambda i: ctx.synth_scalars(ext, seed=1000 + i)  # noqa: E731
fixed = [rnd(i) for i in range(cs.num_fixed_columns)]
advice = [rnd(10 + i) for i in range(cs.num_advice_columns)]
inst = [rnd(20 + i) for i in range(cs.num_instance_columns)]
l0, l_last, l_act = rnd(30), rnd(31), rnd(32)
pc = cs.permutation.columns
sigma = [rnd(40 + i) for i in range(len(pc))]
chunk = cs.degree() - 2
zs = [rnd(50 + i) for i in range((len(pc) + chunk - 1) // chunk)]
values = ctx.alloc(ext * 32)
ctx.memset(values, 0)
cols, keep = h.make_eval_columns(fixed, advice, inst, [5] * cs.num_challenges, 11, 13, 17, 19)
g = ev.custom_gates.compile(ctx)
ctx.set_profile(True)
res = {"k": k, "extended_k": dom.extended_k, "variant": variant,
       "graph_instructions": int(ctx.lib.h2b_graph_num_instructions(g)),
       "graph_slots": int(ctx.lib.h2b_graph_num_slots(g))}
ts = []
for _ in range(5):
    ctx._check(ctx.lib.h2b_evaluate_h_gates(dom.h, g, C.byref(cols), values.ptr))
    ts.append(ctx.last_kernel_ms())
res["gates_ms"] = min(ts)
ctype = np.asarray([c.column_type for c in pc], dtype=np.uint32)
cidx = np.asarray([c.index for c in pc], dtype=np.uint32)
from halo2_pse_b200.plonk import _ptr_array
sg, zz = _ptr_array(sigma), _ptr_array(zs)
ts = []
for _ in range(5):
    ctx._check(ctx.lib.h2b_evaluate_h_permutation(dom.h, C.byref(cols), C.c_void_p(ctype.ctypes.data),
                                                  C.c_void_p(cidx.ctypes.data), len(pc), sg, zz, len(zs), chunk,
                                                  cs.blinding_factors(), l0.ptr, l_last.ptr, l_act.ptr, values.ptr))
    ts.append(ctx.last_kernel_ms())
res["permutation_ms"] = min(ts)
n_in = cs.num_fixed_columns + cs.num_advice_columns + cs.num_instance_columns
res["gates_GBps"] = (n_in + 2) * ext * 32 / res["gates_ms"] / 1e6
res["rows_per_s_total"] = ext / ((res["gates_ms"] + res["permutation_ms"]) * 1e-3)
print(json.dumps(res))
