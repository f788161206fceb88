"""Host <-> device copy rates of pageable vs pinned slices through h2b_copy_h2d / h2b_copy_d2h (GPU box).
usage: copy_bench.py [MiB ...]   env: H2B_COPY_THREADS, H2B_COPY_CHUNK_KB"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [32, 512]
ctx = h.Context(0)
tag = f"threads={os.environ.get('H2B_COPY_THREADS', 'auto')} chunk_kb={os.environ.get('H2B_COPY_CHUNK_KB', 'auto')}"
for mib in sizes:
    n = mib << 20
    a = np.random.RandomState(1).randint(0, 1 << 62, size=(n // 32, 4), dtype=np.int64).astype(np.uint64)
    pin = ctx.pinned((n // 32, 4))
    pin.array[:] = a
    buf = ctx.alloc(n)
    for name, src in (("pageable", a), ("pinned", pin.array)):
        out = np.empty_like(a) if name == "pageable" else pin.array
        p = C.c_void_p(out.ctypes.data)
        ups, downs = [], []
        for i in range(8):
            t0 = time.perf_counter()
            buf.upload(src)
            ups.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            ctx._check(ctx.lib.h2b_copy_d2h(ctx.h, p, buf.ptr, n))
            downs.append(time.perf_counter() - t0)
        assert (out == a).all()
        up, down = min(ups[2:]), min(downs[2:])
        print(f"{mib} MiB {name}: h2d {n / up / 1e9:.1f} GB/s ({up * 1e3:.2f} ms), d2h {n / down / 1e9:.1f} GB/s "
              f"({down * 1e3:.2f} ms) {tag}", flush=True)
    buf.free()
ctx.close()
