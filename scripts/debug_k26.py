import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import halo2_pse_b200 as h
from tests import helpers as H
from oracle import bn256 as O
ctx = h.Context(0)
oc = H.load_oracle_c()
for k in (25, 26):
    a = H.rand_fr_limbs(k, 1 << k)
    w = H.fr_enc([O.omega_for(k)])
    t = time.time(); want = oc.best_fft(a, w[0], k, 0); print("oracle", k, time.time() - t, flush=True)
    buf = ctx.upload_fr(a)
    ctx.best_fft_device(buf, w, k)
    got = buf.download(1 << k)
    bad = np.nonzero((got != want).any(axis=1))[0]
    print("k", k, "mismatching elements:", bad.size, bad[:8], flush=True)
    # second transform of the same (already transformed) buffer vs oracle of `want`
    want2 = oc.best_fft(want, w[0], k, 0)
    ctx.best_fft_device(buf, w, k)
    got2 = buf.download(1 << k)
    bad = np.nonzero((got2 != want2).any(axis=1))[0]
    print("k", k, "2nd transform mismatching:", bad.size, bad[:8], flush=True)
    buf.free()
