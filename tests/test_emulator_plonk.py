"""Quotient evaluation on the CPU kernel emulator (TEST INFRASTRUCTURE) against the
direct big-integer oracle: graph compiler, slot renaming, interpreter, permutation and lookup kernels."""
import numpy as np
import pytest

import halo2_pse_b200 as h
from tests import plonk_cases as PC


@pytest.mark.parametrize("variant,k,ncirc", [("bench", 4, 1), ("rich", 4, 1), ("rich", 5, 2), ("gates_only", 3, 1)])
def test_evaluate_h_vs_oracle(emu_ctx, variant, k, ncirc):
    PC.check_evaluate_h(emu_ctx, variant, k, seed=100 + k, n_circuits=ncirc)


def test_graph_compiler_renames_slots(emu_ctx):
    cs = PC.build_cs("rich")
    ev = h.Evaluator(cs)
    g = ev.custom_gates
    hnd = g.compile(emu_ctx)
    slots = emu_ctx.lib.h2b_graph_num_slots(hnd)
    instrs = emu_ctx.lib.h2b_graph_num_instructions(hnd)
    # Horner over the six gate polynomials is unrolled, one wait follows the hoisted column prefetches;
    # live slots are far fewer than intermediates
    assert instrs == len(g.calculations) - 1 + 6 + 1
    assert 0 < slots < g.num_intermediates
    ev.free()


def test_graph_rejects_malformed_streams(emu_ctx):
    import ctypes as C
    consts = h.fr_encode([0, 1, 2])
    rots = np.asarray([0], dtype=np.int32)

    def new(words, n_inter=4):
        code = np.asarray(words, dtype=np.uint32)
        out = C.c_void_p()
        return emu_ctx.lib.h2b_graph_new(emu_ctx.h, C.c_void_p(code.ctypes.data), code.size,
                                         C.c_void_p(consts.ctypes.data), 3, C.c_void_p(rots.ctypes.data), 1, n_inter,
                                         C.byref(out))

    assert new([7, 0, 0, 1, 0]) == 0                    # Store(Constant(1))
    assert new([7, 0, 0, 9, 0]) == h.H2B_ERR_ARG        # constant index out of range
    assert new([0, 0, 1, 1, 0, 0, 1, 0]) == h.H2B_ERR_ARG  # reads intermediate 1 before it exists
    assert new([7, 9, 0, 1, 0]) == h.H2B_ERR_ARG        # target out of range
    assert new([7, 0, 2, 0, 5]) == h.H2B_ERR_ARG        # rotation index out of range
    assert new([6, 0, 0, 1, 0, 9, 0, 0, 3, 0, 0, 0]) == h.H2B_ERR_ARG  # Horner with a truncated part list
    assert new([42, 0, 0, 1, 0]) == h.H2B_ERR_ARG       # unknown opcode


def test_evaluate_h_without_prefetch_hoisting(emu_ctx, monkeypatch):
    """The original instruction order (kept when hoisted live ranges would not fit shared memory)."""
    monkeypatch.setenv("H2B_EVALH_NO_PREFETCH", "1")
    PC.check_evaluate_h(emu_ctx, "rich", 4, seed=8)


def test_evaluate_h_slot_overflow_path(emu_ctx, monkeypatch):
    """More live slots than shared memory holds: the slot file moves to a global overflow area."""
    monkeypatch.setenv("H2B_EVALH_SMEM_CAP", "4096")
    PC.check_evaluate_h(emu_ctx, "rich", 4, seed=7)


def test_create_proof_bytes_equal_the_oracle(emu_ctx):
    """keygen + create_proof of the bench circuit (benches/plonk.rs) at k = 5: identical verifying key,
    identical proof bytes for the same rng seed, accepted by the restated verifier."""
    PC.check_bench_proof_bytes(emu_ctx, 5)


def test_counter_rng_host_and_device_streams_agree(emu_ctx):
    """CounterRng draws the same Fr::random values one u64 at a time, in bulk on the host
    (h2b_fr_from_u512) and generated on the device (h2b_fr_random_counter)."""
    n = 37
    a, b, c = h.CounterRng(99), h.CounterRng(99), h.CounterRng(99)
    a.next_u64(), b.next_u64(), c.next_u64()  # a non-zero starting counter
    want = [h.fr_random(a) for _ in range(n)]
    dev = h.fr_random_device(emu_ctx, b, n)                      # device generator
    assert h.fr_decode(dev.download(n)) == want

    class HostOnly:  # no fill_fr_device: words from the host, reduction on the device
        def __init__(self, r):
            self.r = r

        def fill_u64(self, m):
            return self.r.fill_u64(m)

    dev2 = h.fr_random_device(emu_ctx, HostOnly(c), n)
    assert h.fr_decode(dev2.download(n)) == want
    assert a.next_u64() == b.next_u64() == c.next_u64()          # the counters moved identically
    # from_u512 edge cases: all-ones words, values just above r
    import numpy as np
    wide = np.array([[2**64 - 1] * 8, [0] * 8, [h.R_MOD & (2**64 - 1), (h.R_MOD >> 64) & (2**64 - 1),
                     (h.R_MOD >> 128) & (2**64 - 1), h.R_MOD >> 192, 1, 0, 0, 0]], dtype=np.uint64)
    out = emu_ctx.alloc(3 * 32)
    import ctypes as C
    emu_ctx._check(emu_ctx.lib.h2b_fr_from_u512(emu_ctx.h, C.c_void_p(wide.ctypes.data), 0, 3, out.ptr))
    assert h.fr_decode(out.download(3)) == [(2**512 - 1) % h.R_MOD, 0, (h.R_MOD + 2**256) % h.R_MOD]


@pytest.mark.parametrize("n,distinct", [(1, 1), (7, 3), (64, 5), (300, 300), (1000, 37)])
def test_lookup_permute_vs_oracle(emu_ctx, n, distinct):
    PC.check_lookup_permute(emu_ctx, n, seed=n, distinct=distinct)


def test_create_proof_with_a_lookup_equals_the_oracle(emu_ctx):
    PC.check_lookup_proof_bytes(emu_ctx, 5)


@pytest.mark.parametrize("which", ["bench", "lookup"])
def test_shplonk_proof_bytes_equal_the_oracle(emu_ctx, which):
    PC.check_shplonk_proof_bytes(emu_ctx, which)


def test_create_proof_with_two_phases_equals_the_oracle(emu_ctx):
    """Advice columns in two phases with a challenge squeezed in between (prover.rs:287-405): proof bytes equal
    the oracle's, the restated verifier accepts them."""
    PC.check_phases_proof_bytes(emu_ctx, 5)
