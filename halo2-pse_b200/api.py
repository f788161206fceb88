"""Host-side mirror of the reference interface for the two hot paths.

Names, argument meaning and error behaviour follow halo2_proofs:

* ``best_multiexp``, ``best_fft``            -- src/arithmetic.rs:132, :171
* ``EvaluationDomain``                       -- src/poly/domain.rs:19-361
* ``ParamsKZG.commit / commit_lagrange``     -- src/poly/kzg/commitment.rs:281-292, 327-334

Field elements cross the boundary exactly as the reference stores them:
``numpy.uint64`` arrays of shape (n, 4) holding little-endian Montgomery limbs
(halo2curves ``Fr([u64; 4])``); affine points are (n, 8) (x limbs, y limbs),
the identity is all zeros.  Where the reference panics, these raise
``H2BError``.  Everything runs in libhalo2b200 on the GPU; nothing here
computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from ._ffi import H2B_DEVICE, H2B_HOST, H2BError

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------
# encodings (pure data conversion, no arithmetic on the hot path)
# --------------------------------------------------------------------------
def _to_limbs(vals: Sequence[int], mod: int) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % mod) * (1 << 256) % mod
        out[i, 0] = m & _MASK64
        out[i, 1] = (m >> 64) & _MASK64
        out[i, 2] = (m >> 128) & _MASK64
        out[i, 3] = m >> 192
    return out


def _from_limbs(arr: np.ndarray, mod: int) -> list:
    rinv = pow(1 << 256, -1, mod)
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in a.tolist():
        m = row[0] | (row[1] << 64) | (row[2] << 128) | (row[3] << 192)
        out.append(m * rinv % mod)
    return out


def fr_encode(vals: Sequence[int]) -> np.ndarray:
    """Canonical integers -> (n, 4) Montgomery limbs of Fr."""
    return _to_limbs(vals, R_MOD)


def fr_decode(arr: np.ndarray) -> list:
    return _from_limbs(arr, R_MOD)


def fq_encode(vals: Sequence[int]) -> np.ndarray:
    return _to_limbs(vals, Q_MOD)


def fq_decode(arr: np.ndarray) -> list:
    return _from_limbs(arr, Q_MOD)


def g1_encode(points) -> np.ndarray:
    """[(x, y) | None] -> (n, 8) limbs; None (identity) -> zeros."""
    out = np.zeros((len(points), 8), dtype=np.uint64)
    idx = [i for i, p in enumerate(points) if p is not None]
    if idx:
        out[idx, :4] = fq_encode([points[i][0] for i in idx])
        out[idx, 4:] = fq_encode([points[i][1] for i in idx])
    return out


def g1_decode(arr: np.ndarray) -> list:
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 8)
    xs = fq_decode(a[:, :4])
    ys = fq_decode(a[:, 4:])
    out = []
    for i in range(a.shape[0]):
        out.append(None if not a[i].any() else (xs[i], ys[i]))
    return out


def g1_jacobian_to_affine(arr: np.ndarray):
    """(12,) Jacobian limbs -> affine (x, y) or None.  Normalisation of a single
    returned point, as the reference does with `.to_affine()` before hashing."""
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(3, 4)
    x, y, z = fq_decode(a)
    if z == 0:
        return None
    zi = pow(z, -1, Q_MOD)
    return (x * zi * zi % Q_MOD, y * zi * zi * zi % Q_MOD)


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _fr_array(a, n: Optional[int] = None) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.uint64)
    if arr.ndim == 1:
        arr = arr.reshape(-1, 4)
    if arr.ndim != 2 or arr.shape[1] != 4:
        raise H2BError(_ffi.H2B_ERR_ARG, "expected (n, 4) uint64 limbs")
    return arr


# --------------------------------------------------------------------------
# context and device buffers
# --------------------------------------------------------------------------
class DeviceBuffer:
    """A device-resident array of Fr (or raw bytes) owned by a Context."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx = ctx
        self.nbytes = nbytes
        p = C.c_void_p()
        ctx._check(ctx.lib.h2b_device_alloc(ctx.h, nbytes, C.byref(p)))
        self.ptr = p

    def upload(self, arr: np.ndarray, offset_bytes: int = 0, ctx: Optional["Context"] = None) -> "DeviceBuffer":
        """Host -> device copy, complete on return.  `ctx`: a sibling context whose stream (and pinned ring)
        carries the copy instead of the owner's -- the buffer must not be in use on the owner's stream."""
        a = np.ascontiguousarray(arr)
        if offset_bytes + a.nbytes > self.nbytes:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "upload larger than the buffer")
        ctx = self.ctx if ctx is None else ctx
        ctx._check(ctx.lib.h2b_copy_h2d(ctx.h, C.c_void_p(self.ptr.value + offset_bytes), _ptr(a), a.nbytes))
        return self

    def download(self, count: int, offset_bytes: int = 0, width: int = 4) -> np.ndarray:
        out = np.empty((count, width), dtype=np.uint64)
        if offset_bytes + out.nbytes > self.nbytes:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "download larger than the buffer")
        self.ctx._check(self.ctx.lib.h2b_copy_d2h(self.ctx.h, _ptr(out),
                                                  C.c_void_p(self.ptr.value + offset_bytes), out.nbytes))
        return out

    def at(self, offset_bytes: int) -> C.c_void_p:
        return C.c_void_p(self.ptr.value + offset_bytes)

    def free(self) -> None:
        if self.ptr is not None and self.ctx.h is not None:
            self.ctx.lib.h2b_device_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host memory viewed as a numpy uint64 array (pinned staging)."""

    def __init__(self, lib, shape: Tuple[int, ...]):
        self.lib = lib
        n = int(np.prod(shape)) * 8
        p = C.c_void_p()
        rc = lib.h2b_host_alloc(n, C.byref(p))
        if rc != 0:
            raise H2BError(rc, "pinned allocation failed")
        self.ptr = p
        buf = (C.c_uint64 * (n // 8)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.uint64).reshape(shape)

    def free(self):
        if self.ptr is not None:
            self.array = None
            self.lib.h2b_host_free(self.ptr)
            self.ptr = None


class Context:
    """One GPU + stream + scratch.  `lib_path` is for the test-suite's emulator
    build only; the product always loads halo2-pse_b200/lib/libhalo2b200.so."""

    def __init__(self, device: int = 0, lib_path: Optional[str] = None):
        self.lib = _ffi.load(lib_path)
        h = C.c_void_p()
        rc = self.lib.h2b_ctx_create(device, C.byref(h))
        if rc != 0:
            raise H2BError(rc, "h2b_ctx_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self.device = device
        self._lib_path = lib_path
        self._aux: List["Context"] = []

    def aux(self, i: int = 0) -> "Context":
        """A sibling context on the same device (own stream, scratch and MSM workspace): independent
        commitments run on both at once, the way the reference runs rayon par_iter around (not inside)
        its commits (poly/kzg/multiopen/shplonk/prover.rs:179-196)."""
        while len(self._aux) <= i:
            self._aux.append(Context(self.device, self._lib_path))
        return self._aux[i]

    def close(self) -> None:
        for a in self._aux:
            a.close()
        self._aux = []
        if self.h is not None:
            self.lib.h2b_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        pass  # explicit close(); buffers may outlive GC ordering otherwise

    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self.lib.h2b_last_error(self.h)
            raise H2BError(rc, msg.decode() if msg else "")

    def sync(self) -> None:
        self._check(self.lib.h2b_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.h2b_ctx_launches(self.h)) + sum(a.launches for a in self._aux)

    @property
    def stream(self) -> int:
        return int(self.lib.h2b_ctx_stream(self.h) or 0)

    def set_profile(self, on: bool) -> None:
        self.lib.h2b_ctx_set_profile(self.h, 1 if on else 0)

    def last_kernel_ms(self) -> float:
        """Duration of the dominant kernel of the last MSM call made in profile mode."""
        return float(self.lib.h2b_ctx_last_kernel_ms(self.h))

    def last_ntt_pass_ms(self) -> list:
        buf = (C.c_float * 5)()
        n = self.lib.h2b_ctx_last_ntt_passes(self.h, buf, 5)
        return [float(buf[i]) for i in range(n)]

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def pinned(self, shape) -> PinnedArray:
        return PinnedArray(self.lib, tuple(shape))

    def memset(self, buf: DeviceBuffer, value: int = 0, nbytes: Optional[int] = None, offset_bytes: int = 0) -> None:
        nbytes = buf.nbytes - offset_bytes if nbytes is None else nbytes
        self._check(self.lib.h2b_device_memset(self.h, buf.at(offset_bytes), value, nbytes))

    def clone(self, buf: DeviceBuffer, nbytes: Optional[int] = None) -> DeviceBuffer:
        nbytes = buf.nbytes if nbytes is None else nbytes
        out = self.alloc(nbytes)
        self._check(self.lib.h2b_copy_d2d(self.h, out.ptr, buf.ptr, nbytes))
        return out

    def upload_fr(self, arr) -> DeviceBuffer:
        a = _fr_array(arr)
        return self.alloc(max(a.nbytes, 32)).upload(a)

    # ---- synthetic inputs ----
    def synth_scalars(self, n: int, seed: int = 0x68616C6F32, kind: int = 0) -> DeviceBuffer:
        buf = self.alloc(max(n, 1) * 32)
        self._check(self.lib.h2b_synth_scalars(self.h, buf.ptr, n, seed, kind))
        return buf

    def synth_bases(self, n: int, seed: int = 0x6B7A67) -> "Bases":
        buf = self.alloc(max(n, 1) * 64)
        self._check(self.lib.h2b_synth_bases(self.h, buf.ptr, n, seed))
        b = Bases(self, buf.ptr, n, H2B_DEVICE)
        buf.free()
        return b

    def synth_base_scalar(self, seed: int, i: int) -> int:
        return int(self.lib.h2b_synth_base_scalar(seed, i))

    PIPE_BENCH = {"imad": 0, "imad_hi": 1, "imad_wide": 2, "carry_chain": 3, "fr_mul": 4}

    def pipe_peak(self, which: str = "imad_wide"):
        """(32x32 multiplies/s, instructions-or-mulmods/s) of a register-only microbenchmark."""
        v = C.c_double()
        m = C.c_double()
        self._check(self.lib.h2b_pipe_peak(self.h, self.PIPE_BENCH[which], C.byref(v), C.byref(m)))
        return v.value, m.value

    # ---- the two reference entry points ----
    def best_multiexp(self, coeffs, bases):
        """arithmetic.rs:132 -- host slices in, affine point (x, y) | None out."""
        c = _fr_array(coeffs)
        b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        if c.shape[0] != b.shape[0]:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(coeffs.len(), bases.len())")  # :133
        out = np.zeros(12, dtype=np.uint64)
        self._check(self.lib.h2b_best_multiexp(self.h, _ptr(c), _ptr(b), c.shape[0], _ptr(out)))
        return g1_jacobian_to_affine(out)

    def small_multiexp(self, coeffs, bases):
        """arithmetic.rs:105-125 -- shared-doubling double-and-add over a few points (host-side, as in the
        reference); affine point (x, y) | None out."""
        c = _fr_array(coeffs)
        b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        if c.shape[0] > b.shape[0]:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "index out of bounds: bases[coeff_idx]")  # :117
        out = np.zeros(12, dtype=np.uint64)
        self._check(self.lib.h2b_small_multiexp(_ptr(c), _ptr(b), c.shape[0], _ptr(out)))
        return g1_jacobian_to_affine(out)

    def g_to_lagrange(self, g, k: int) -> np.ndarray:
        """arithmetic.rs:277-301 -- (2^k, 8) affine limbs in, the Lagrange-basis points out (host arrays)."""
        b = np.ascontiguousarray(g, dtype=np.uint64).reshape(-1, 8)
        if b.shape[0] != 1 << k:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(a.len(), 1 << log_n)")  # best_fft, :184
        out = np.zeros_like(b)
        self._check(self.lib.h2b_g_to_lagrange(self.h, _ptr(b), H2B_HOST, k, _ptr(out), H2B_HOST))
        return out

    def best_fft(self, a: np.ndarray, omega, log_n: int) -> np.ndarray:
        """arithmetic.rs:171 -- in place on the (n, 4) limb array `a`."""
        arr = _fr_array(a)
        if arr.shape[0] != 1 << log_n:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(a.len(), 1 << log_n)")  # :184
        w = fr_encode([omega]) if isinstance(omega, int) else _fr_array(omega)
        self._check(self.lib.h2b_best_fft(self.h, _ptr(arr), H2B_HOST, _ptr(w), log_n))
        if arr is not a:
            np.copyto(np.asarray(a).reshape(-1, 4), arr)
        return a

    # ---- polynomial helpers: host arrays or DeviceBuffers ----
    @staticmethod
    def _fr_arg(x, n=None):
        """(pointer, loc, count, keepalive) of a host limb array or a DeviceBuffer."""
        if isinstance(x, DeviceBuffer):
            if n is None:
                raise H2BError(_ffi.H2B_ERR_ARG, "n required for device polynomials")
            return x.ptr, H2B_DEVICE, n, x
        arr = _fr_array(x)
        return _ptr(arr), H2B_HOST, arr.shape[0] if n is None else n, arr

    def eval_polynomial(self, poly, point: int, n: Optional[int] = None) -> int:
        """arithmetic.rs:304 -- sum poly[i] * point^i."""
        p, loc, n, _keep = self._fr_arg(poly, n)
        x = fr_encode([point])
        out = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.h2b_eval_polynomial(self.h, p, loc, n, _ptr(x), _ptr(out)))
        return fr_decode(out)[0]

    def kate_division(self, a, b: int, n: Optional[int] = None, out: Optional[DeviceBuffer] = None):
        """arithmetic.rs:348 -- a(X) / (X - b); host arrays in -> (n-1, 4) array out, DeviceBuffer in -> `out`."""
        p, loc, n, _keep = self._fr_arg(a, n)
        if n == 0:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "kate_division of an empty polynomial")
        x = fr_encode([b])
        if loc == H2B_DEVICE:
            if out is None:
                out = self.alloc(max(n - 1, 1) * 32)
            self._check(self.lib.h2b_kate_division(self.h, p, loc, n, _ptr(x), out.ptr))
            return out
        q = np.zeros((n - 1, 4), dtype=np.uint64)
        self._check(self.lib.h2b_kate_division(self.h, p, loc, n, _ptr(x), _ptr(q) if n > 1 else None))
        return q

    def inner_product(self, a, b, n: Optional[int] = None) -> int:
        """arithmetic.rs:331 -- panics (H2B_ERR_LENGTH) if the lengths differ."""
        pa, la, na, _ka = self._fr_arg(a, n)
        pb, lb, nb, _kb = self._fr_arg(b, n)
        if na != nb or la != lb:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(a.len(), b.len())")  # :334
        out = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.h2b_inner_product(self.h, pa, pb, la, na, _ptr(out)))
        return fr_decode(out)[0]

    def _poly_binop(self, fn, lhs, rhs, n):
        pl, ll, nl, keep = self._fr_arg(lhs, n)
        pr, lr, nr, _k = self._fr_arg(rhs, n)
        if nl != nr or ll != lr:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "polynomials of different length")
        if ll == H2B_HOST:
            keep = keep.copy()
            pl = _ptr(keep)
        self._check(fn(self.h, pl, pr, ll, nl))
        return keep

    def poly_add(self, lhs, rhs, n: Optional[int] = None):
        """poly.rs:229 -- lhs + rhs (host: new array; DeviceBuffer: in place)."""
        return self._poly_binop(self.lib.h2b_poly_add, lhs, rhs, n)

    def poly_sub(self, lhs, rhs, n: Optional[int] = None):
        """poly.rs:243"""
        return self._poly_binop(self.lib.h2b_poly_sub, lhs, rhs, n)

    def poly_scale(self, a, scalar: int, n: Optional[int] = None):
        """poly.rs:278 -- a * scalar."""
        p, loc, n, keep = self._fr_arg(a, n)
        if loc == H2B_HOST:
            keep = keep.copy()
            p = _ptr(keep)
        s = fr_encode([scalar])
        self._check(self.lib.h2b_poly_scale(self.h, p, loc, n, _ptr(s)))
        return keep

    def batch_invert(self, a, n: Optional[int] = None):
        """ff::BatchInvert (plonk/permutation/prover.rs:119): element-wise inverses, zeros stay zero."""
        p, loc, n, keep = self._fr_arg(a, n)
        if loc == H2B_HOST:
            keep = keep.copy()
            p = _ptr(keep)
        self._check(self.lib.h2b_batch_invert(self.h, p, loc, n))
        return keep

    def running_product(self, f, init: int = 1, n: Optional[int] = None, out: Optional[DeviceBuffer] = None):
        """z[0] = init, z[i] = z[i-1] * f[i-1] (plonk/permutation/prover.rs:152-158), n values."""
        p, loc, n, _keep = self._fr_arg(f, n)
        i0 = fr_encode([init])
        if loc == H2B_DEVICE:
            if out is None:
                out = self.alloc(max(n, 1) * 32)
            self._check(self.lib.h2b_running_product(self.h, p, loc, n, _ptr(i0), out.ptr))
            return out
        z = np.zeros((n, 4), dtype=np.uint64)
        self._check(self.lib.h2b_running_product(self.h, p, loc, n, _ptr(i0), _ptr(z)))
        return z

    def best_fft_device(self, buf: DeviceBuffer, omega, log_n: int, ncols: int = 1,
                        stride: Optional[int] = None) -> None:
        w = fr_encode([omega]) if isinstance(omega, int) else _fr_array(omega)
        stride = stride if stride is not None else 1 << log_n
        self._check(self.lib.h2b_best_fft_batch(self.h, buf.ptr, H2B_DEVICE, _ptr(w), log_n, ncols, stride))


def msm_many_mixed_raw(jobs: Sequence[tuple], concurrent: bool = True, pre: Optional[Sequence] = None) -> List[np.ndarray]:
    """Independent commitments: jobs[i] = (bases, scalars, n, offset, scalar_offset), any base sets of one
    device; pre[i](ctx), if given, runs right before job i on the context that will compute it (the upload of
    a witness column: copies of later columns then overlap the commitments of earlier ones).
    With more than one job they are dealt round-robin to the first base set's context and its siblings (one
    host thread and one stream each, H2B_COMMIT_WAYS of them, default 2), so that the latency-bound tail of
    one MSM (bucket reduction, short sort passes) runs under the throughput-bound accumulation of the
    others -- the reference's rayon par_iter around (not inside) its commits
    (poly/kzg/multiopen/shplonk/prover.rs:179-196)."""
    if not jobs:
        return []
    jobs = [tuple(j) + (None, 0, 0)[len(j) - 2:] for j in jobs]
    main = jobs[0][0].ctx
    b0, _, n0, off0, _ = jobs[0]
    if (len(jobs) > 1 and concurrent and not os.environ.get("H2B_MSM_NO_MULTI") and b0.table_window_bits
            and all(j[0] is b0 and isinstance(j[1], DeviceBuffer) and j[2] == n0 and j[3] == off0 for j in jobs)
            and (pre is None or all(f is None for f in pre))):
        # the same base slice under every job, device-resident scalars: ONE multi-column MSM
        return b0.msm_multi_raw(main, [(j[1], j[4]) for j in jobs], n0, off0)
    # 2 ways by default (more brings nothing once the accumulation kernels saturate the GPU); 3 when the jobs
    # carry their own host->device copies, which keep a worker off the GPU for a while
    has_copies = pre is not None and any(f is not None for f in pre)
    ways = min(len(jobs), max(1, int(os.environ.get("H2B_COMMIT_WAYS", "3" if has_copies else "2"))))
    if ways < 2 or not concurrent:
        outs = []
        for i, j in enumerate(jobs):
            if pre is not None and pre[i] is not None:
                pre[i](main)
            outs.append(j[0].msm_raw(main, *j[1:]))
        return outs
    import threading
    ctxs = [main] + [main.aux(i) for i in range(ways - 1)]
    main.sync()  # the scalars were produced on the main stream
    outs: List[Optional[np.ndarray]] = [None] * len(jobs)
    errs: list = []

    def run(ctx, idxs):
        try:
            for i in idxs:
                if pre is not None and pre[i] is not None:
                    pre[i](ctx)
                outs[i] = jobs[i][0].msm_raw(ctx, *jobs[i][1:])
        except Exception as e:  # re-raised on the calling thread
            errs.append(e)

    threads = [threading.Thread(target=run, args=(ctxs[w], range(w, len(jobs), ways))) for w in range(1, ways)]
    for t in threads:
        t.start()
    run(ctxs[0], range(0, len(jobs), ways))
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return outs


class Bases:
    """Device-resident, immutable affine bases: ParamsKZG.g or .g_lagrange
    (poly/kzg/commitment.rs:23-31)."""

    def __init__(self, ctx: Context, src, n: int, loc: int = H2B_HOST):
        self.ctx = ctx
        h = C.c_void_p()
        if loc == H2B_HOST:
            arr = np.ascontiguousarray(src, dtype=np.uint64).reshape(-1, 8)
            n = arr.shape[0]
            ctx._check(ctx.lib.h2b_bases_upload(ctx.h, _ptr(arr), n, H2B_HOST, C.byref(h)))
        else:
            ctx._check(ctx.lib.h2b_bases_upload(ctx.h, src, n, H2B_DEVICE, C.byref(h)))
        self.h = h
        self.n = n

    def __len__(self) -> int:
        return self.n

    def precompute(self, window_bits: int = 0) -> "Bases":
        """Build the window table (one-time; the bases of a ParamsKZG never change)."""
        self.ctx._check(self.ctx.lib.h2b_bases_precompute(self.ctx.h, self.h, window_bits))
        return self

    @property
    def table_window_bits(self) -> int:
        return int(self.ctx.lib.h2b_bases_table_window_bits(self.h))

    def download(self) -> np.ndarray:
        out = np.empty((self.n, 8), dtype=np.uint64)
        p = self.ctx.lib.h2b_bases_device_ptr(self.h)
        self.ctx._check(self.ctx.lib.h2b_copy_d2h(self.ctx.h, _ptr(out), p, out.nbytes))
        return out

    def msm_raw(self, ctx: "Context", scalars, n: Optional[int] = None, offset: int = 0,
                scalar_offset: int = 0) -> np.ndarray:
        """The affine sum as its 8 limbs, computed on `ctx` (this base set's context or a sibling on the
        same device: the bases are immutable, any stream may read them)."""
        if isinstance(scalars, DeviceBuffer):
            if n is None:
                raise H2BError(_ffi.H2B_ERR_ARG, "n required for device scalars")
            if (scalar_offset + n) * 32 > scalars.nbytes:
                raise H2BError(_ffi.H2B_ERR_LENGTH, "scalar slice outside the buffer")
            sp, loc = scalars.at(scalar_offset * 32), H2B_DEVICE
        else:
            arr = _fr_array(scalars)[scalar_offset:]
            n = arr.shape[0] if n is None else n
            sp, loc = _ptr(arr), H2B_HOST
        out = np.zeros(8, dtype=np.uint64)
        ctx._check(ctx.lib.h2b_msm_affine(ctx.h, self.h, offset, sp, loc, n, _ptr(out)))
        return out

    def msm_multi_raw(self, ctx: "Context", cols: Sequence[tuple], n: int, offset: int = 0) -> List[np.ndarray]:
        """cols[j] = (DeviceBuffer, scalar_offset): the commitments of all columns to bases[offset..offset + n] in one
        pass (h2b_msm_multi_affine: shared digit / sort / accumulate / reduce pipeline, one bucket set per column)."""
        for buf, so in cols:
            if (so + n) * 32 > buf.nbytes:
                raise H2BError(_ffi.H2B_ERR_LENGTH, "scalar slice outside the buffer")
        ptrs = (C.c_void_p * len(cols))(*[buf.ptr.value + so * 32 for buf, so in cols])
        out = np.zeros((len(cols), 8), dtype=np.uint64)
        ctx._check(ctx.lib.h2b_msm_multi_affine(ctx.h, self.h, offset, ptrs, len(cols), n, _ptr(out)))
        return [out[j].copy() for j in range(len(cols))]

    def msm_many_raw(self, jobs: Sequence[tuple], concurrent: bool = True, pre: Optional[Sequence] = None) -> List[np.ndarray]:
        """Independent commitments on this base set: jobs[i] = (scalars, n, offset, scalar_offset); see
        msm_many_mixed_raw."""
        return msm_many_mixed_raw([(self,) + tuple(j) for j in jobs], concurrent, pre)

    @staticmethod
    def msm_many_mixed(jobs: Sequence[tuple], concurrent: bool = True, pre: Optional[Sequence] = None) -> list:
        """jobs[i] = (bases, scalars, n, offset, scalar_offset) over any base sets of one device."""
        return [g1_decode(o)[0] for o in msm_many_mixed_raw(jobs, concurrent, pre)]

    def msm_many(self, jobs: Sequence[tuple], concurrent: bool = True, pre: Optional[Sequence] = None) -> list:
        return [g1_decode(o)[0] for o in self.msm_many_raw(jobs, concurrent, pre)]

    def msm(self, scalars, n: Optional[int] = None, offset: int = 0, affine: bool = True, scalar_offset: int = 0):
        """best_multiexp(scalars[scalar_offset..][..n], &bases[offset..offset+n]); scalars are host limbs
        or a DeviceBuffer.  Returns the affine point (x, y) | None."""
        if isinstance(scalars, DeviceBuffer):
            if n is None:
                raise H2BError(_ffi.H2B_ERR_ARG, "n required for device scalars")
            if (scalar_offset + n) * 32 > scalars.nbytes:
                raise H2BError(_ffi.H2B_ERR_LENGTH, "scalar slice outside the buffer")
            sp, loc = scalars.at(scalar_offset * 32), H2B_DEVICE
        else:
            arr = _fr_array(scalars)[scalar_offset:]
            n = arr.shape[0] if n is None else n
            sp, loc = _ptr(arr), H2B_HOST
        if affine:
            out = np.zeros(8, dtype=np.uint64)
            self.ctx._check(self.ctx.lib.h2b_msm_affine(self.ctx.h, self.h, offset, sp, loc, n, _ptr(out)))
            return g1_decode(out)[0]
        out = np.zeros(12, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.h2b_msm(self.ctx.h, self.h, offset, sp, loc, n, _ptr(out)))
        return g1_jacobian_to_affine(out)

    def free(self) -> None:
        if self.h is not None and self.ctx.h is not None:
            self.ctx.lib.h2b_bases_free(self.h)
        self.h = None


class ParamsKZG:
    """The commit half of poly/kzg/commitment.rs: k, n, g, g_lagrange."""

    def __init__(self, ctx: Context, k: int, g, g_lagrange):
        self.ctx = ctx
        self.k = k
        self.n = 1 << k
        # anything with the Bases interface passes through (dist.ShardedBases: one range per GPU)
        self.g = g if hasattr(g, "msm") else Bases(ctx, g, self.n)
        self.g_lagrange = g_lagrange if hasattr(g_lagrange, "msm") else Bases(ctx, g_lagrange, self.n)
        self.g2 = self.s_g2 = None  # the two G2 points of the verifier half (host integers; see serde.py)

    @classmethod
    def setup(cls, ctx: Context, k: int, s: int, precompute: bool = False) -> "ParamsKZG":
        """ParamsKZG::setup (poly/kzg/commitment.rs:61-129) with the toxic secret `s` given instead of
        drawn from an rng (MUST NOT be used in production, as the reference says): g[i] = [s^i] G and
        g_lagrange[i] = [(s^n - 1)/n * w^i / (s - w^i)] G.  The O(n) scalar bookkeeping is host integers;
        the 2n scalar multiplications run on the GPU."""
        if k > 28:
            raise H2BError(_ffi.H2B_ERR_ARG, "assert!(k <= E::Scalar::S)")  # :64
        r = R_MOD
        n = 1 << k
        root = pow(pow(7, (r - 1) >> 28, r), 1 << (28 - k), r)  # :90-93
        powers, roots = [1] * n, [1] * n
        for i in range(1, n):
            powers[i] = powers[i - 1] * s % r
            roots[i] = roots[i - 1] * root % r
        multiplier = (pow(s, n, r) - 1) * pow(n, -1, r) % r  # :96
        dens = [(s - w) % r for w in roots]
        # one inversion for all denominators (Montgomery's trick)
        pre, run = [1] * n, 1
        for i in range(n):
            pre[i] = run
            run = run * dens[i] % r
        inv = pow(run, -1, r)
        lag = [0] * n
        for i in range(n - 1, -1, -1):
            lag[i] = multiplier * roots[i] % r * (inv * pre[i] % r) % r  # :101-102
            inv = inv * dens[i] % r
        out = []
        for scalars in (powers, lag):
            sc = fr_encode(scalars)
            buf = ctx.alloc(n * 64)
            ctx._check(ctx.lib.h2b_g1_mul_generator(ctx.h, _ptr(sc), H2B_HOST, n, buf.ptr, H2B_DEVICE))
            b = Bases(ctx, buf.ptr, n, H2B_DEVICE)
            buf.free()
            if precompute:
                b.precompute()
            out.append(b)
        params = cls(ctx, k, out[0], out[1])
        from . import serde  # :118-119  g2 = generator of G2, s_g2 = [s] g2
        params.g2, params.s_g2 = serde.G2_GENERATOR, serde.g2_mul(serde.G2_GENERATOR, s % R_MOD)
        return params

    def downsize(self, k: int) -> None:
        """ParamsKZG::downsize (poly/kzg/commitment.rs:267-275): g.truncate(1 << k) and
        g_lagrange = g_to_lagrange(g, k), both on the device (the group NTT of arithmetic.rs:277-301)."""
        if k > self.k:
            raise H2BError(_ffi.H2B_ERR_ARG, "assert!(k <= self.k)")  # :268
        if not isinstance(self.g, Bases):
            raise H2BError(_ffi.H2B_ERR_ARG, "downsize needs the replicated base vectors (shard afterwards)")
        ctx, n = self.ctx, 1 << k
        had_table = bool(self.g.table_window_bits)
        gp = C.c_void_p(int(ctx.lib.h2b_bases_device_ptr(self.g.h)))
        new_g = Bases(ctx, gp, n, H2B_DEVICE)  # device-to-device copy of the first n points
        buf = ctx.alloc(n * 64)
        ctx._check(ctx.lib.h2b_g_to_lagrange(ctx.h, gp, H2B_DEVICE, k, buf.ptr, H2B_DEVICE))
        new_l = Bases(ctx, buf.ptr, n, H2B_DEVICE)
        buf.free()
        self.g.free()
        self.g_lagrange.free()
        self.g, self.g_lagrange, self.k, self.n = new_g, new_l, k, n
        if had_table:
            self.g.precompute()
            self.g_lagrange.precompute()

    def commit(self, poly, blind=None):
        """:327-334 -- blind is accepted and ignored, as in the reference."""
        a = _fr_array(poly)
        if a.shape[0] > len(self.g):
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert!(bases.len() >= size)")  # :332
        return self.g.msm(a)

    def commit_lagrange(self, poly, blind=None):
        """:281-292"""
        a = _fr_array(poly)
        if a.shape[0] > len(self.g_lagrange):
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert!(bases.len() >= size)")  # :290
        return self.g_lagrange.msm(a)


class EvaluationDomain:
    """poly/domain.rs:19-361 for G = Fr."""

    _CONST = {"omega": 0, "omega_inv": 1, "extended_omega": 2, "extended_omega_inv": 3, "g_coset": 4,
              "g_coset_inv": 5, "ifft_divisor": 6, "extended_ifft_divisor": 7}

    def __init__(self, ctx: Context, j: int, k: int):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._check(ctx.lib.h2b_domain_new(ctx.h, j, k, C.byref(h)))
        self.h = h
        self.k = int(ctx.lib.h2b_domain_k(h))
        self.extended_k = int(ctx.lib.h2b_domain_extended_k(h))
        self.n = 1 << self.k
        self.quotient_len = int(ctx.lib.h2b_domain_quotient_len(h))

    def constant(self, name: str) -> int:
        out = np.zeros(4, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.h2b_domain_constant(self.h, self._CONST[name], _ptr(out)))
        return fr_decode(out)[0]

    def t_evaluations(self) -> list:
        out = []
        for i in range(1 << (self.extended_k - self.k)):
            v = np.zeros(4, dtype=np.uint64)
            self.ctx._check(self.ctx.lib.h2b_domain_constant(self.h, 8 + i, _ptr(v)))
            out.append(fr_decode(v)[0])
        return out

    def extended_len(self) -> int:
        return 1 << self.extended_k

    def lagrange_to_coeff(self, a) -> np.ndarray:
        arr = _fr_array(a).copy()
        if arr.shape[0] != self.n:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(a.values.len(), 1 << self.k)")  # :227
        self.ctx._check(self.ctx.lib.h2b_lagrange_to_coeff(self.h, _ptr(arr), H2B_HOST))
        return arr

    def coeff_to_extended(self, a) -> np.ndarray:
        arr = _fr_array(a)
        if arr.shape[0] != self.n:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(a.values.len(), 1 << self.k)")  # :244
        out = np.empty((self.extended_len(), 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.h2b_coeff_to_extended(self.h, _ptr(arr), _ptr(out), H2B_HOST))
        return out

    def extended_to_coeff(self, a, divide_by_vanishing: bool = False) -> np.ndarray:
        arr = _fr_array(a)
        if arr.shape[0] != self.extended_len():
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(a.values.len(), extended_len())")  # :282
        out = np.empty((self.quotient_len, 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.h2b_extended_to_coeff(self.h, _ptr(arr), _ptr(out), H2B_HOST,
                                                           1 if divide_by_vanishing else 0))
        return out

    def divide_by_vanishing_poly(self, a) -> np.ndarray:
        arr = _fr_array(a).copy()
        if arr.shape[0] != self.extended_len():
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert_eq!(a.values.len(), extended_len())")  # :311
        self.ctx._check(self.ctx.lib.h2b_divide_by_vanishing_poly(self.h, _ptr(arr), H2B_HOST))
        return arr

    # device-resident, batched (column-parallel) forms
    def lagrange_to_coeff_device(self, buf: DeviceBuffer, ncols: int = 1, stride: Optional[int] = None):
        stride = self.n if stride is None else stride
        self.ctx._check(self.ctx.lib.h2b_lagrange_to_coeff_batch(self.h, buf.ptr, H2B_DEVICE, ncols, stride))

    def coeff_to_extended_device(self, src: DeviceBuffer, dst: DeviceBuffer, ncols: int = 1,
                                 in_stride: Optional[int] = None, out_stride: Optional[int] = None):
        in_stride = self.n if in_stride is None else in_stride
        out_stride = self.extended_len() if out_stride is None else out_stride
        self.ctx._check(self.ctx.lib.h2b_coeff_to_extended_batch(self.h, src.ptr, in_stride, dst.ptr,
                                                                 out_stride, H2B_DEVICE, ncols))

    def extended_to_coeff_device(self, src: DeviceBuffer, dst: DeviceBuffer, ncols: int = 1,
                                 in_stride: Optional[int] = None, out_stride: Optional[int] = None,
                                 divide_by_vanishing: bool = False):
        in_stride = self.extended_len() if in_stride is None else in_stride
        out_stride = self.quotient_len if out_stride is None else out_stride
        self.ctx._check(self.ctx.lib.h2b_extended_to_coeff_batch(
            self.h, src.ptr, in_stride, dst.ptr, out_stride, H2B_DEVICE, ncols,
            1 if divide_by_vanishing else 0))

    def free(self) -> None:
        if self.h is not None and self.ctx.h is not None:
            self.ctx.lib.h2b_domain_free(self.h)
        self.h = None
