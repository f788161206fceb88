"""Shared test helpers: ctypes wrapper of the C++ oracle, encodings, seeded inputs."""
from __future__ import annotations

import ctypes as C
import json
import os
import random
import subprocess

import numpy as np

from oracle import bn256 as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libh2b_oracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")
MASK64 = (1 << 64) - 1


def build_oracle_c() -> str:
    src = os.path.join(ROOT, "oracle", "ref_cpu.cpp")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return ORACLE_SO


class OracleC:
    """oracle/ref_cpu.cpp through ctypes.  numpy (n, 4) uint64 Montgomery limbs in and out."""

    def __init__(self, path: str):
        lib = C.CDLL(path)
        P, SZ, I, U32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32
        lib.oracle_best_multiexp.argtypes = [P, P, SZ, I, P]
        lib.oracle_best_fft.argtypes = [P, P, U32, I]
        lib.oracle_domain_new.argtypes = [U32, U32, I]
        lib.oracle_domain_new.restype = P
        lib.oracle_domain_free.argtypes = [P]
        lib.oracle_domain_extended_k.argtypes = [P]
        lib.oracle_domain_extended_k.restype = U32
        lib.oracle_domain_quotient_len.argtypes = [P]
        lib.oracle_domain_quotient_len.restype = SZ
        lib.oracle_domain_constant.argtypes = [P, U32, P]
        lib.oracle_lagrange_to_coeff.argtypes = [P, P]
        lib.oracle_coeff_to_extended.argtypes = [P, P, P]
        lib.oracle_divide_by_vanishing_poly.argtypes = [P, P]
        lib.oracle_extended_to_coeff.argtypes = [P, P]
        lib.oracle_field_op.argtypes = [I, I, P, P, P, SZ]
        lib.oracle_g1_mul_gen_u64.argtypes = [P, SZ, I, P]
        lib.oracle_synth_bases.argtypes = [C.c_uint64, C.c_uint64, SZ, I, P]
        lib.oracle_eval_polynomial.argtypes = [P, SZ, P, I, P]
        lib.oracle_kate_division.argtypes = [P, SZ, P, P]
        lib.oracle_inner_product.argtypes = [P, P, SZ, P]
        lib.oracle_graph_new.argtypes = [P, SZ, P, U32, P, U32, U32]
        lib.oracle_graph_new.restype = P
        lib.oracle_graph_free.argtypes = [P]
        lib.oracle_evaluate_h.argtypes = [P, P, P, P, P, P, P, P, P, U32, P, P, U32, U32, U32, P, P, P, P, P, U32, P, I]
        self.lib = lib

    @staticmethod
    def _p(a):
        return C.c_void_p(a.ctypes.data)

    def best_multiexp(self, coeffs: np.ndarray, bases: np.ndarray, threads: int = 0) -> np.ndarray:
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        assert c.shape[0] == b.shape[0]
        out = np.zeros(8, dtype=np.uint64)
        self.lib.oracle_best_multiexp(self._p(c), self._p(b), c.shape[0], threads, self._p(out))
        return out

    def best_fft(self, a: np.ndarray, omega: np.ndarray, log_n: int, threads: int = 0) -> np.ndarray:
        arr = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4).copy()
        assert arr.shape[0] == 1 << log_n
        w = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
        assert self.lib.oracle_best_fft(self._p(arr), self._p(w), log_n, threads) == 0
        return arr

    def field_op(self, field: int, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(a)
        self.lib.oracle_field_op(field, op, self._p(a), self._p(b), self._p(out), a.shape[0])
        return out

    def g1_mul_gen(self, ks, threads: int = 0) -> np.ndarray:
        k = np.ascontiguousarray(ks, dtype=np.uint64)
        out = np.zeros((k.shape[0], 8), dtype=np.uint64)
        self.lib.oracle_g1_mul_gen_u64(self._p(k), k.shape[0], threads, self._p(out))
        return out

    def eval_polynomial(self, poly: np.ndarray, point: np.ndarray, threads: int = 0) -> np.ndarray:
        p = np.ascontiguousarray(poly, dtype=np.uint64).reshape(-1, 4)
        x = np.ascontiguousarray(point, dtype=np.uint64).reshape(4)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.oracle_eval_polynomial(self._p(p), p.shape[0], self._p(x), threads, self._p(out))
        return out

    def kate_division(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        x = np.ascontiguousarray(b, dtype=np.uint64).reshape(4)
        q = np.zeros((a.shape[0] - 1, 4), dtype=np.uint64)
        assert self.lib.oracle_kate_division(self._p(a), a.shape[0], self._p(x), self._p(q)) == 0
        return q

    def inner_product(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.oracle_inner_product(self._p(a), self._p(b), a.shape[0], self._p(out))
        return out

    def synth_bases(self, n: int, a: int = 0x1234567, d: int = 0x9E3779B9, threads: int = 0) -> np.ndarray:
        """out[i] = [a + i*d] G"""
        out = np.zeros((n, 8), dtype=np.uint64)
        self.lib.oracle_synth_bases(a, d, n, threads, self._p(out))
        return out

    def graph(self, words: np.ndarray, constants: np.ndarray, rotations, num_intermediates: int):
        """A GraphEvaluator's calculation stream (the reference's enums, flattened) for oracle_evaluate_h."""
        w = np.ascontiguousarray(words, dtype=np.uint32)
        c = np.ascontiguousarray(constants, dtype=np.uint64).reshape(-1, 4)
        r = np.ascontiguousarray(rotations, dtype=np.int32)
        g = self.lib.oracle_graph_new(self._p(w), w.size, self._p(c), c.shape[0], self._p(r), r.size, num_intermediates)
        assert g, "malformed calculation stream"
        return g

    def evaluate_h(self, dom: "OracleCDomain", gates_graph, fixed, advice, instance, challenges, beta, gamma, theta, y,
                   perm_columns, sigma_cosets, z_cosets, chunk_len, blinding_factors, l0, l_last, l_active,
                   lookup_graphs, lookup_cosets, values: np.ndarray, threads: int = 0) -> np.ndarray:
        """Evaluator::evaluate_h for one circuit instance, in place on `values` ((2^ek, 4) uint64)."""
        def ptrs(arrs):
            a = (C.c_void_p * max(len(arrs), 1))()
            for i, x in enumerate(arrs):
                a[i] = x.ctypes.data
            return a
        ch = np.ascontiguousarray(challenges if len(challenges) else np.zeros((1, 4), dtype=np.uint64), dtype=np.uint64)
        sc = np.ascontiguousarray(np.concatenate([fr_enc([v]) for v in (beta, gamma, theta, y)]), dtype=np.uint64)
        ct = np.asarray([c[0] for c in perm_columns] or [0], dtype=np.uint32)
        ci = np.asarray([c[1] for c in perm_columns] or [0], dtype=np.uint32)
        lg = (C.c_void_p * max(len(lookup_graphs), 1))(*lookup_graphs)
        rc = self.lib.oracle_evaluate_h(dom.h, gates_graph, ptrs(fixed), ptrs(advice), ptrs(instance), self._p(ch),
                                        self._p(sc), self._p(ct), self._p(ci), len(perm_columns), ptrs(sigma_cosets),
                                        ptrs(z_cosets), len(z_cosets), chunk_len, blinding_factors, self._p(l0),
                                        self._p(l_last), self._p(l_active), lg, ptrs(lookup_cosets), len(lookup_graphs),
                                        self._p(values), threads)
        assert rc == 0
        return values

    def domain(self, j: int, k: int, threads: int = 0) -> "OracleCDomain":
        return OracleCDomain(self, j, k, threads)


class OracleCDomain:
    def __init__(self, oc: OracleC, j: int, k: int, threads: int):
        self.oc = oc
        self.h = oc.lib.oracle_domain_new(j, k, threads)
        assert self.h, "oracle_domain_new failed"
        self.k = k
        self.n = 1 << k
        self.extended_k = int(oc.lib.oracle_domain_extended_k(self.h))
        self.quotient_len = int(oc.lib.oracle_domain_quotient_len(self.h))

    def constant(self, which: int) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        assert self.oc.lib.oracle_domain_constant(self.h, which, OracleC._p(out)) == 0
        return out

    def lagrange_to_coeff(self, a):
        arr = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4).copy()
        assert arr.shape[0] == self.n
        self.oc.lib.oracle_lagrange_to_coeff(self.h, OracleC._p(arr))
        return arr

    def coeff_to_extended(self, a):
        arr = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        assert arr.shape[0] == self.n
        out = np.zeros((1 << self.extended_k, 4), dtype=np.uint64)
        self.oc.lib.oracle_coeff_to_extended(self.h, OracleC._p(arr), OracleC._p(out))
        return out

    def divide_by_vanishing_poly(self, a):
        arr = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4).copy()
        assert arr.shape[0] == 1 << self.extended_k
        self.oc.lib.oracle_divide_by_vanishing_poly(self.h, OracleC._p(arr))
        return arr

    def extended_to_coeff(self, a):
        arr = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4).copy()
        assert arr.shape[0] == 1 << self.extended_k
        self.oc.lib.oracle_extended_to_coeff(self.h, OracleC._p(arr))
        return arr[: self.quotient_len]

    def free(self):
        if self.h:
            self.oc.lib.oracle_domain_free(self.h)
            self.h = None


def load_oracle_c() -> OracleC:
    return OracleC(build_oracle_c())


# ---- encodings (independent of the package under test) -----------------------
def to_limbs(vals, mod) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % mod) * (1 << 256) % mod
        out[i] = [(m >> (64 * j)) & MASK64 for j in range(4)]
    return out


def from_limbs(arr, mod) -> list:
    rinv = pow(1 << 256, -1, mod)
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [(r[0] | (r[1] << 64) | (r[2] << 128) | (r[3] << 192)) * rinv % mod for r in a.tolist()]


def fr_enc(vals):
    return to_limbs(vals, O.R_MOD)


def fr_dec(arr):
    return from_limbs(arr, O.R_MOD)


def g1_enc(points) -> np.ndarray:
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is not None:
            out[i, :4] = to_limbs([p[0]], O.Q_MOD)[0]
            out[i, 4:] = to_limbs([p[1]], O.Q_MOD)[0]
    return out


def g1_dec(arr) -> list:
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 8)
    out = []
    for row in a:
        if not row.any():
            out.append(None)
        else:
            out.append((from_limbs(row[:4], O.Q_MOD)[0], from_limbs(row[4:], O.Q_MOD)[0]))
    return out


def rand_fr(rng: random.Random, n: int) -> list:
    return [rng.randrange(O.R_MOD) for _ in range(n)]


def rand_fr_limbs(seed: int, n: int) -> np.ndarray:
    """n pseudo-random reduced Montgomery residues, generated with numpy (fast at 2^20+)."""
    g = np.random.default_rng(seed)
    a = g.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + \
        g.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)  # < 2^252 < r: every value is a valid reduced residue
    return a


def load_golden(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)
