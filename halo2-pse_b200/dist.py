"""Multi-GPU sharding of the two hot paths (one process per GPU, torch.distributed).

The reference is single-process (rayon threads only); the
analogue of its per-thread partitioning is kept where the path shards naturally:

* MSM            -- contiguous point ranges per rank, exactly the chunking of
                    `best_multiexp` (arithmetic.rs:135-153): each rank runs a full
                    local Pippenger on its range, the per-rank partial points
                    (64 B each) are all-gathered and folded.  No data-path collective.
* column batches -- column c of a batch of independent transforms/commitments
                    belongs to rank c % world.  No communication at all.
* one giant NTT  -- four-step decomposition n = n1 * n2 over row-sharded
                    matrices; the only real exchange step of the path is the
                    transpose, an all-to-all (NCCL over NVLink on GPUs).

Buffers that take part in a collective are torch tensors (int64 views of the Fr
limbs) on the process group's device; the kernels run on them through the C ABI
with H2B_DEVICE pointers.  With the test-suite's CPU emulator build and the
gloo backend the same code runs on CPU tensors (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from ._ffi import H2B_DEVICE, H2BError
from .api import Bases, Context, DeviceBuffer, fr_encode, g1_decode


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the
    default process group when world_size > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous point range [start, end) of `rank`: ceil(n / world) points per rank,
    the last ranks may be short or empty (the reference chunks the same way,
    arithmetic.rs:135-145, with the remainder in a trailing chunk)."""
    per = -(-n // world) if n else 0
    start = min(n, rank * per)
    return start, min(n, start + per)


def column_owner(col: int, world: int) -> int:
    return col % world


def _group_device(group=None):
    import torch
    import torch.distributed as dist
    be = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if be == "nccl" else torch.device("cpu")


class ShardedMSM:
    """best_multiexp over a point range sharded across ranks.

    Each rank holds `bases` = its contiguous slice of the global base vector
    (device-resident) and passes the matching slice of the scalars."""

    def __init__(self, ctx: Context, bases: Bases, group=None):
        self.ctx = ctx
        self.bases = bases
        self.group = group

    def msm(self, scalars, n_local: Optional[int] = None):
        """Returns the affine point sum over ALL ranks' ranges (same value on every rank)."""
        import torch
        import torch.distributed as dist
        out = np.zeros(8, dtype=np.uint64)
        if isinstance(scalars, DeviceBuffer):
            sp, loc, n = scalars.ptr, H2B_DEVICE, n_local
        else:
            arr = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
            sp, loc, n = C.c_void_p(arr.ctypes.data), _ffi.H2B_HOST, arr.shape[0]
        self.ctx._check(self.ctx.lib.h2b_msm_affine(self.ctx.h, self.bases.h, 0, sp, loc, n,
                                                    C.c_void_p(out.ctypes.data)))
        return _gather_fold(self.ctx, out, self.group)


class _CudaView:
    """A device buffer of the library seen through __cuda_array_interface__ (torch.as_tensor shares the memory)."""

    def __init__(self, ptr: int, nwords: int):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def _as_tensor(buf: DeviceBuffer, nwords: int, dev):
    """int64 tensor over the first `nwords` words of a DeviceBuffer, on the process group's device (no copy)."""
    import torch
    if dev.type == "cuda":
        return torch.as_tensor(_CudaView(int(buf.ptr.value), nwords), device=dev)
    return torch.frombuffer((C.c_int64 * nwords).from_address(int(buf.ptr.value)), dtype=torch.int64)


def _gather_fold_many(ctx: Context, outs: np.ndarray, group=None) -> list:
    """outs: (m, 8) limbs, this rank's partial points of m independent MSMs.  ONE all-gather of m * 64
    bytes per rank, each column folded in rank order with h2b_g1_sum (`.fold` of the per-chunk results,
    arithmetic.rs:153); the same affine points on every rank."""
    import torch
    import torch.distributed as dist
    outs = np.ascontiguousarray(outs, dtype=np.uint64).reshape(-1, 8)
    m = outs.shape[0]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return g1_decode(outs)
    world = dist.get_world_size(group)
    dev = _group_device(group)
    mine = torch.from_numpy(outs.view(np.int64).reshape(-1)).to(dev)
    parts = torch.empty(world * m * 8, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(parts, mine, group=group)
    allp = parts.cpu().numpy().view(np.uint64).reshape(world, m, 8)
    res = []
    for i in range(m):
        col = np.ascontiguousarray(allp[:, i, :])
        total = np.zeros(8, dtype=np.uint64)
        rc = ctx.lib.h2b_g1_sum(C.c_void_p(col.ctypes.data), world, C.c_void_p(total.ctypes.data))
        if rc != 0:
            raise H2BError(rc, "h2b_g1_sum")
        res.append(g1_decode(total)[0])
    return res


def _gather_fold(ctx: Context, out: np.ndarray, group=None):
    return _gather_fold_many(ctx, out, group)[0]


class ShardedBases:
    """A base vector of a ParamsKZG (`g` or `g_lagrange`) sharded by contiguous point range over the
    ranks of `group`, behind the interface of `api.Bases`: `msm(scalars, n, offset, scalar_offset)`
    is best_multiexp(scalars[scalar_offset..][..n], bases[offset..offset + n]) and returns the same
    affine point on every rank.  Rank g keeps only bases[start_g, end_g) (and its window table) on its
    device; every rank passes the WHOLE scalar vector (the prover's polynomials are replicated, the
    Fiat-Shamir chain is not shardable) and multiplies its own range of it.  No data-path collective:
    the partial points (64 B per rank) are all-gathered and folded.

    This is what makes `create_proof` a multi-GPU call: `ParamsKZG(ctx, k, ShardedBases, ShardedBases)`
    goes through `keygen` / `create_proof` unchanged and every commitment runs on all GPUs at once."""

    def __init__(self, ctx: Context, local: Bases, n_total: int, start: int, group=None):
        self.ctx, self.local, self.n, self.start, self.group = ctx, local, n_total, start, group
        self.end = start + len(local)

    @classmethod
    def from_full(cls, ctx: Context, full: Bases, group=None, precompute: bool = True,
                  free_full: bool = True) -> "ShardedBases":
        """Keep this rank's range of a replicated device-resident base vector (device-to-device copy)."""
        import torch.distributed as dist
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = len(full)
        s, e = shard_range(n, rank, world)
        base_ptr = int(ctx.lib.h2b_bases_device_ptr(full.h) or 0)
        local = Bases(ctx, C.c_void_p(base_ptr + s * 64), e - s, H2B_DEVICE)
        if precompute and e > s:
            local.precompute()
        if free_full:
            full.free()
        return cls(ctx, local, n, s, group)

    def __len__(self) -> int:
        return self.n

    def check_replicated(self, what: str, digest: Optional[bytes]) -> None:
        """Raise unless `digest` is the same on every rank of the group.

        The sharded prover multiplies, on every rank, that rank's point range of a polynomial that is supposed
        to be REPLICATED (same witness, same blinding rows): with per-rank entropy (an OsRng, as the reference
        passes) the all-gathered commitment would be a sum of partial MSMs over different polynomials and the
        proof silently invalid.  `create_proof` calls this with the rng's position before it draws anything and
        with the transcript state after the openings.  digest = None (an rng that cannot report its state) is
        only accepted if every rank passes None; the end-of-proof check still catches a divergence."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        dev = _group_device(self.group)
        raw = (b"\x01" + digest[:31]) if digest is not None else bytes(32)
        mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        parts = torch.empty(world * 32, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(parts, mine, group=self.group)
        rows = parts.cpu().view(world, 32)
        if not bool((rows == rows[0]).all()):
            raise H2BError(_ffi.H2B_ERR_ARG, "sharded create_proof: ranks disagree on " + what)

    @property
    def table_window_bits(self) -> int:
        return self.local.table_window_bits

    def precompute(self, window_bits: int = 0) -> "ShardedBases":
        if len(self.local):
            self.local.precompute(window_bits)
        return self

    def upload_columns(self, bufs: Sequence[DeviceBuffer], cols: Sequence[np.ndarray]) -> None:
        """Host columns (the witness) -> the replicated device buffers of every rank, without every rank pulling
        every byte over its own PCIe link and through its own host threads: rank g copies rows [start_g, end_g)
        of each column from the host, then ONE in-place all-gather per column spreads them over NVLink.  Rows
        past the end of a column (padding and blinding rows) were already written identically on every rank and
        travel along unchanged."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        per = self.n // world if world else self.n
        if world == 1 or self.n % world or per * world != self.n:
            for b, col in zip(bufs, cols):
                if col.shape[0]:
                    b.upload(col)
            return
        self.ctx.sync()  # padding / blinding rows are in place before anything is gathered
        dev = _group_device(self.group)
        # this rank's rows of every column, the columns side by side on sibling contexts (one host thread, one
        # stream and one pinned ring each): pageable copies are host-bound, not PCIe-bound
        import threading
        errs: list = []

        def put(ctx, b, col):
            try:
                lo, hi = self.start, min(self.end, col.shape[0])
                if hi > lo:
                    b.upload(col[lo:hi], lo * 32, ctx=ctx)  # complete on return
            except Exception as e:
                errs.append(e)

        threads = [threading.Thread(target=put, args=(self.ctx.aux(i - 1), b, col))
                   for i, (b, col) in enumerate(zip(bufs, cols)) if i > 0]
        for t in threads:
            t.start()
        if bufs:
            put(self.ctx, bufs[0], cols[0])
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        pending = []
        for b in bufs:
            full = _as_tensor(b, self.n * 4, dev)
            mine = full[self.start * 4:(self.start + per) * 4]
            if dev.type == "cuda":
                # in place (`mine` is its own slot of `full`), asynchronous
                pending.append(dist.all_gather_into_tensor(full, mine, group=self.group, async_op=True))
            else:
                out = torch.empty_like(full)
                dist.all_gather_into_tensor(out, mine.clone(), group=self.group)
                full.copy_(out)
        for w in pending:
            w.wait()
        if dev.type == "cuda":
            torch.cuda.current_stream().synchronize()

    def _local_job(self, scalars, n, offset, scalar_offset):
        """This rank's part of best_multiexp(scalars[scalar_offset..][..n], bases[offset..offset + n]) as a job
        for the local base set, or None if its range misses the slice."""
        if not isinstance(scalars, DeviceBuffer):
            scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
            n = scalars.shape[0] - scalar_offset if n is None else n
        elif n is None:
            raise H2BError(_ffi.H2B_ERR_ARG, "n required for device scalars")
        if offset + n > self.n:
            raise H2BError(_ffi.H2B_ERR_LENGTH, "assert!(bases.len() >= size)")  # poly/kzg/commitment.rs:290,332
        lo, hi = max(self.start, offset), min(self.end, offset + n)
        if hi <= lo:
            return None
        return (scalars, hi - lo, lo - self.start, scalar_offset + (lo - offset))

    def msm_many(self, jobs: Sequence[tuple], concurrent: bool = True, pre: Optional[Sequence] = None) -> list:
        return ShardedBases.msm_many_mixed([(self,) + tuple(j) for j in jobs], concurrent, pre)

    @staticmethod
    def msm_many_mixed(jobs: Sequence[tuple], concurrent: bool = True, pre: Optional[Sequence] = None) -> list:
        """Independent commitments, jobs[i] = (sharded bases, scalars, n, offset, scalar_offset): the local parts
        run on this GPU (several at a time, api.msm_many_mixed_raw), then ONE all-gather carries every job's
        partial point.  pre[i](ctx) as in api.msm_many_mixed_raw."""
        if not jobs:
            return []
        jobs = [tuple(j) + (None, 0, 0)[len(j) - 2:] for j in jobs]
        first = jobs[0][0]
        local = [j[0]._local_job(*j[1:]) for j in jobs]
        outs = np.zeros((len(jobs), 8), dtype=np.uint64)  # identity where this rank's range misses the slice
        live = [i for i, j in enumerate(local) if j is not None]
        if pre is not None:
            for i, j in enumerate(local):
                if j is None and pre[i] is not None:
                    pre[i](first.ctx)
        lpre = None if pre is None else [pre[i] for i in live]
        from .api import msm_many_mixed_raw
        for i, o in zip(live, msm_many_mixed_raw([(jobs[i][0].local,) + local[i] for i in live], concurrent, lpre)):
            outs[i] = o
        return _gather_fold_many(first.ctx, outs, first.group)

    def msm(self, scalars, n: Optional[int] = None, offset: int = 0, affine: bool = True, scalar_offset: int = 0):
        return self.msm_many([(scalars, n, offset, scalar_offset)])[0]

    def free(self) -> None:
        self.local.free()


def broadcast_seed(seed: Optional[int] = None, group=None, src: int = 0) -> int:
    """A 64-bit rng seed that is the same on every rank: rank `src`'s `seed` (or fresh OS entropy there) is
    broadcast.  The sharded prover needs identical blinding factors on every rank (ShardedBases.check_replicated):
    build the rng passed to create_proof from this, e.g. CounterRng(broadcast_seed())."""
    import torch
    import torch.distributed as dist
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return seed & ((1 << 64) - 1)
    t = torch.tensor([(seed & ((1 << 64) - 1)) - (1 << 63)], dtype=torch.int64, device=_group_device(group))
    dist.broadcast(t, src=src, group=group)
    return int(t.item()) + (1 << 63)


def shard_params(params, group=None, precompute: bool = True):
    """ParamsKZG with `g` and `g_lagrange` replaced by their ShardedBases (the replicated full vectors
    are released): keygen and create_proof on the result commit on every GPU of the group.

    REQUIREMENT (checked by create_proof through ShardedBases.check_replicated, H2BError otherwise): every rank
    passes the same witness and an rng in the same state -- e.g. CounterRng(broadcast_seed()) -- because the
    polynomials are replicated and only the base vectors are sharded."""
    from .api import ParamsKZG
    out = ParamsKZG(params.ctx, params.k,
                    ShardedBases.from_full(params.ctx, params.g, group, precompute),
                    ShardedBases.from_full(params.ctx, params.g_lagrange, group, precompute))
    out.g2, out.s_g2 = params.g2, params.s_g2
    return out


class FourStepNTT:
    """One 2^k-point best_fft whose vector is sharded in natural order across the
    ranks of `group` (rank g holds elements [g*n/G, (g+1)*n/G)), natural order out.

    n = n1 * n2 viewed as an n1 x n2 row-major matrix A[j1][j2]; with
    K = K1 + n1*K2:
        X[K] = sum_j2 w^(j2*K1) * ( sum_j1 A[j1][j2] * (w^n2)^(j1*K1) ) * (w^n1)^(j2*K2)
    Steps (T = distributed transpose = local tile transposes + all-to-all + local permute):
        T ; n1-point row NTTs ; twiddle w^(j2*K1) ; T ; n2-point row NTTs ; T
    """

    def __init__(self, ctx: Context, log_n: int, omega: int, group=None, p2p: Optional[bool] = None):
        """p2p: None = use NVLink peer stores (symmetric memory) when the group is NCCL and it can be
        set up, else the all-to-all collective; True = require it; False = collective only."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.k, self.group = ctx, log_n, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        G = self.world
        if G & (G - 1):
            raise H2BError(_ffi.H2B_ERR_ARG, "world size must be a power of two")
        lg = G.bit_length() - 1
        # n = n1 * n2.  On the p2p path n1 is at most 2^9, ONE pass of the register kernel, so that a transform makes
        # 1 + 2 passes like the single-GPU schedule (8 + 9 + 9 bits at k = 26) instead of 2 + 2; H2B_FOURSTEP_SQUARE=1
        # restores the square split.
        self.k1 = log_n // 2
        if p2p is not False and self.k1 > 9 and not os.environ.get("H2B_FOURSTEP_SQUARE"):
            self.k1 = 9
        self.k2 = log_n - self.k1
        if self.k1 < lg or self.k2 < lg:
            raise H2BError(_ffi.H2B_ERR_ARG, "transform too small to shard over this many ranks")
        self.n1, self.n2 = 1 << self.k1, 1 << self.k2
        self.omega = omega
        from .api import R_MOD
        self.w = fr_encode([omega])
        self.w1 = fr_encode([pow(omega, self.n2, R_MOD)])  # order n1
        self.w2 = fr_encode([pow(omega, self.n1, R_MOD)])  # order n2
        self.dev = _group_device(group) if self.world > 1 else None
        self.local = (1 << log_n) // G
        self.p2p = False
        self.p2p_error = None
        if self.world > 1 and self.dev is not None and self.dev.type == "cuda" and p2p is not False:
            try:
                self._init_p2p()
            except Exception as e:  # symmetric memory unavailable: keep the collective path
                if p2p:
                    raise
                self.p2p_error = f"{type(e).__name__}: {e}"

    def _init_p2p(self) -> None:
        import torch.distributed._symmetric_memory as sm
        torch, dist = self.torch, self.dist
        grp = self.group if self.group is not None else dist.group.WORLD
        self.S = [sm.empty(self.local * 4, dtype=torch.int64, device=self.dev) for _ in range(2)]
        self.H = [sm.rendezvous(t, grp) for t in self.S]
        G = self.world
        self.peer = [(C.c_void_p * G)(*[int(p) for p in hd.buffer_ptrs]) for hd in self.H]
        self.xstream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        self.p2p = True

    def _scatter(self, src_ptr, dst: int, rows_local: int, cols: int) -> None:
        ctx = self.ctx
        ctx._check(ctx.lib.h2b_fr_transpose_scatter(ctx.h, src_ptr, self.peer[dst], self.world, self.rank,
                                                    rows_local, cols))

    def _rows_scatter(self, src, dst: int, omega_l, log_len: int, rows: int, total_rows: int, twiddle: bool) -> None:
        ctx = self.ctx
        ctx._check(ctx.lib.h2b_best_fft_rows_scatter(
            ctx.h, self._p(src), C.c_void_p(omega_l.ctypes.data), log_len, rows, self.peer[dst], self.world,
            self.rank * rows, total_rows, C.c_void_p(self.w.ctypes.data) if twiddle else None, self.k))

    # ONE fused transpose+exchange kernel, then two batches of row transforms whose LAST pass stores every output
    # (times the four-step twiddle for the first batch) straight into the rank that owns it in the transposed
    # matrix: 5 passes over the local slab (1 + 2 + 2 at k = 26) where the three-transpose schedule made 9, no
    # twiddle kernel, no final copy.  NVLink carries 3 x (G-1)/G of the vector per transform.
    exchanges = 3
    passes_note = "transpose+exchange, row NTTs (+twiddle) with the 2nd exchange in their last pass, row NTTs with the 3rd"

    @property
    def nvlink_bytes(self) -> Optional[int]:
        """Bytes this rank sends over NVLink per transform (p2p path)."""
        if not self.p2p:
            return None
        return 3 * self.local * 32 * (self.world - 1) // self.world

    def _run_p2p(self, a):
        """Everything is enqueued on the library's stream; device-side barriers between the steps.  The result is
        returned in the symmetric buffer S1 (valid until the next run() on this object)."""
        ctx, G, torch = self.ctx, self.world, self.torch
        n1, n2 = self.n1, self.n2
        S1, S2 = self.S
        H1, H2 = self.H
        torch.cuda.current_stream().synchronize()  # `a` may have been produced on torch's stream
        timing = os.environ.get("H2B_FOURSTEP_TIMING")
        ev = []

        def mark():
            if timing:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.xstream)
                ev.append(e)

        with torch.cuda.stream(self.xstream):
            H1.barrier(0)  # every peer is done with S1 of the previous transform
            mark()
            self._scatter(self._p(a), 0, n1 // G, n2)           # A[j1][j2] -> A^T[j2][j1] in S1
            mark()
            H1.barrier(0)
            mark()
            # n1-point transforms of the rows j2 of A^T, times w^(j2*K1), stored as B[K1][j2] in the peers' S2
            self._rows_scatter(S1, 1, self.w1, self.k1, n2 // G, n2, True)
            mark()
            H2.barrier(0)
            mark()
            # n2-point transforms of the rows K1 of B; C[K1][K2] = X[K1 + n1*K2] stored in natural order in the peers' S1
            self._rows_scatter(S2, 0, self.w2, self.k2, n1 // G, n1, False)
            mark()
            H1.barrier(0)
            mark()
        ctx.sync()
        if timing:
            names = ["transpose+exchange", "barrier", "rows n1 (+twiddle, exchange)", "barrier", "rows n2 (+exchange)", "barrier"]
            self.last_stage_ms = {f"{i}:{n}": ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
        return S1

    # -- helpers ---------------------------------------------------------
    def _p(self, t) -> C.c_void_p:
        return C.c_void_p(t.data_ptr())

    def _alloc(self, like):
        return self.torch.empty_like(like)

    def _sync_torch(self) -> None:
        if self.dev is not None and self.dev.type == "cuda":
            self.torch.cuda.current_stream().synchronize()

    def _transpose(self, src, dst, tmp, R: int, Cn: int) -> None:
        """Global R x Cn matrix sharded by rows (src: [R/G][Cn]) -> its transpose
        sharded by rows (dst: [Cn/G][R]).  src is clobbered (it is the receive
        buffer of the all-to-all); tmp is scratch of the same size."""
        G, ctx = self.world, self.ctx
        Rl, Cl = R // G, Cn // G
        if G == 1:
            ctx._check(ctx.lib.h2b_fr_transpose_batch(ctx.h, self._p(src), self._p(dst), Rl, Cl, Cn, 1, Cl,
                                                      Cl * Rl))
            return
        # send[h][c][r] = src[r][h*Cl + c]: G tile transposes, one launch
        ctx._check(ctx.lib.h2b_fr_transpose_batch(ctx.h, self._p(src), self._p(tmp), Rl, Cl, Cn, G, Cl,
                                                  Cl * Rl))
        ctx.sync()  # the library's stream is not the collective's stream
        self.dist.all_to_all_single(src, tmp, group=self.group)
        self._sync_torch()
        # src now holds [g][c][r_g]; wanted [c][g][r_g]
        ctx._check(ctx.lib.h2b_fr_permute3(ctx.h, self._p(src), self._p(dst), G, Cl, Rl))

    def run(self, a):
        """a: int64 tensor of 4 * n/G limb words (this rank's natural-order slice).  Returns the tensor that holds
        this rank's slice of the transform: `a` itself (transformed in place) on the collective path, the
        symmetric buffer on the p2p path (valid until the next run(); `a` is left untouched there)."""
        if self.p2p:
            return self._run_p2p(a)
        ctx, G = self.ctx, self.world
        n1, n2 = self.n1, self.n2
        self.dev = a.device
        t1, t2 = self._alloc(a), self._alloc(a)
        self._sync_torch()
        # A[j1][j2] rows sharded  ->  A^T[j2][j1]
        self._transpose(a, t1, t2, n1, n2)
        rows = n2 // G
        ctx._check(ctx.lib.h2b_best_fft_batch(ctx.h, self._p(t1), H2B_DEVICE, C.c_void_p(self.w1.ctypes.data),
                                              self.k1, rows, n1))
        # B^T[j2][K1] *= w^(j2*K1)
        ctx._check(ctx.lib.h2b_fr_twiddle_rows(ctx.h, self._p(t1), C.c_void_p(self.w.ctypes.data), self.k,
                                               self.rank * rows, rows, n1))
        # -> B[K1][j2]
        self._transpose(t1, a, t2, n2, n1)
        rows = n1 // G
        ctx._check(ctx.lib.h2b_best_fft_batch(ctx.h, self._p(a), H2B_DEVICE, C.c_void_p(self.w2.ctypes.data),
                                              self.k2, rows, n2))
        # C[K1][K2] = X[K1 + n1*K2]  ->  natural order is the transpose [K2][K1]
        self._transpose(a, t1, t2, n1, n2)
        ctx.sync()
        a.copy_(t1)
        self._sync_torch()
        return a
