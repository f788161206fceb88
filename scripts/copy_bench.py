"""Host <-> device copy rates of pageable vs pinned slices through h2b_copy_h2d / h2b_copy_d2h (GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_pse_b200 as h  # noqa: E402

ctx = h.Context(0)
for mib in (32, 512):
    n = mib << 20
    a = np.random.RandomState(1).randint(0, 1 << 62, size=(n // 32, 4), dtype=np.int64).astype(np.uint64)
    pin = ctx.pinned((n // 32, 4))
    pin.array[:] = a
    buf = ctx.alloc(n)
    for name, src in (("pageable", a), ("pinned", pin.array)):
        for _ in range(2):
            buf.upload(src)
        t0 = time.perf_counter()
        for _ in range(5):
            buf.upload(src)
        up = (time.perf_counter() - t0) / 5
        out = np.empty_like(a) if name == "pageable" else pin.array
        p = out.ctypes.data
        import ctypes as C
        for _ in range(2):
            ctx._check(ctx.lib.h2b_copy_d2h(ctx.h, C.c_void_p(p), buf.ptr, n))
        t0 = time.perf_counter()
        for _ in range(5):
            ctx._check(ctx.lib.h2b_copy_d2h(ctx.h, C.c_void_p(p), buf.ptr, n))
        down = (time.perf_counter() - t0) / 5
        assert (out == a).all()
        print(f"{mib} MiB {name}: h2d {n / up / 1e9:.1f} GB/s ({up * 1e3:.2f} ms), d2h {n / down / 1e9:.1f} GB/s "
              f"({down * 1e3:.2f} ms), threads={os.environ.get('H2B_COPY_THREADS', 'auto')}", flush=True)
    buf.free()
ctx.close()
