"""ctypes binding of libhalo2b200's C ABI (include/halo2_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared
object is missing or no CUDA device is usable, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libhalo2b200.so")

H2B_OK = 0
H2B_ERR_ARG = -1
H2B_ERR_LENGTH = -2
H2B_ERR_CUDA = -3
H2B_ERR_OOM = -4
H2B_ERR_BAD_OMEGA = -5
H2B_ERR_CONSTRAINT = -6
H2B_HOST = 0
H2B_DEVICE = 1

_ERR_NAMES = {
    H2B_ERR_ARG: "H2B_ERR_ARG",
    H2B_ERR_LENGTH: "H2B_ERR_LENGTH",
    H2B_ERR_CUDA: "H2B_ERR_CUDA",
    H2B_ERR_OOM: "H2B_ERR_OOM",
    H2B_ERR_BAD_OMEGA: "H2B_ERR_BAD_OMEGA",
    H2B_ERR_CONSTRAINT: "H2B_ERR_CONSTRAINT",
}


class H2BError(RuntimeError):
    """Raised where the reference would panic (assert_eq!/assert!) or CUDA failed."""

    def __init__(self, code: int, msg: str = ""):
        self.code = code
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")


# every symbol include/halo2_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SZ = C.c_size_t
_U32 = C.c_uint32
_U64 = C.c_uint64
_I = C.c_int
SYMBOLS = {
    "h2b_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "h2b_ctx_destroy": (None, [_P]),
    "h2b_last_error": (C.c_char_p, [_P]),
    "h2b_ctx_sync": (_I, [_P]),
    "h2b_ctx_stream": (_P, [_P]),
    "h2b_ctx_launches": (_U64, [_P]),
    "h2b_ctx_set_profile": (None, [_P, _I]),
    "h2b_ctx_last_kernel_ms": (C.c_float, [_P]),
    "h2b_ctx_last_ntt_passes": (_I, [_P, C.POINTER(C.c_float), _I]),
    "h2b_bases_upload": (_I, [_P, _P, _SZ, _I, C.POINTER(_P)]),
    "h2b_bases_free": (None, [_P]),
    "h2b_bases_precompute": (_I, [_P, _P, _U32]),
    "h2b_bases_table_window_bits": (_U32, [_P]),
    "h2b_bases_len": (_SZ, [_P]),
    "h2b_bases_device_ptr": (_P, [_P]),
    "h2b_msm": (_I, [_P, _P, _SZ, _P, _I, _SZ, _P]),
    "h2b_msm_affine": (_I, [_P, _P, _SZ, _P, _I, _SZ, _P]),
    "h2b_msm_multi_affine": (_I, [_P, _P, _SZ, _P, _U32, _SZ, _P]),
    "h2b_best_multiexp": (_I, [_P, _P, _P, _SZ, _P]),
    "h2b_msm_window_bits": (_U32, [_SZ]),
    "h2b_g1_mul_generator": (_I, [_P, _P, _I, _SZ, _P, _I]),
    "h2b_g1_sum": (_I, [_P, _SZ, _P]),
    "h2b_fr_repr": (_I, [_P, _P, _I, _SZ, _I, _P]),
    "h2b_small_multiexp": (_I, [_P, _P, _SZ, _P]),
    "h2b_g_to_lagrange": (_I, [_P, _P, _I, _U32, _P, _I]),
    "h2b_best_fft": (_I, [_P, _P, _I, _P, _U32]),
    "h2b_best_fft_batch": (_I, [_P, _P, _I, _P, _U32, _U32, _SZ]),
    "h2b_domain_new": (_I, [_P, _U32, _U32, C.POINTER(_P)]),
    "h2b_domain_free": (None, [_P]),
    "h2b_domain_k": (_U32, [_P]),
    "h2b_domain_extended_k": (_U32, [_P]),
    "h2b_domain_quotient_len": (_SZ, [_P]),
    "h2b_domain_constant": (_I, [_P, _U32, _P]),
    "h2b_lagrange_to_coeff": (_I, [_P, _P, _I]),
    "h2b_coeff_to_extended": (_I, [_P, _P, _P, _I]),
    "h2b_extended_to_coeff": (_I, [_P, _P, _P, _I, _I]),
    "h2b_divide_by_vanishing_poly": (_I, [_P, _P, _I]),
    "h2b_lagrange_to_coeff_batch": (_I, [_P, _P, _I, _U32, _SZ]),
    "h2b_coeff_to_extended_batch": (_I, [_P, _P, _SZ, _P, _SZ, _I, _U32]),
    "h2b_extended_to_coeff_batch": (_I, [_P, _P, _SZ, _P, _SZ, _I, _U32, _I]),
    "h2b_eval_polynomial": (_I, [_P, _P, _I, _SZ, _P, _P]),
    "h2b_kate_division": (_I, [_P, _P, _I, _SZ, _P, _P]),
    "h2b_inner_product": (_I, [_P, _P, _P, _I, _SZ, _P]),
    "h2b_poly_add": (_I, [_P, _P, _P, _I, _SZ]),
    "h2b_poly_sub": (_I, [_P, _P, _P, _I, _SZ]),
    "h2b_poly_scale": (_I, [_P, _P, _I, _SZ, _P]),
    "h2b_batch_invert": (_I, [_P, _P, _I, _SZ]),
    "h2b_running_product": (_I, [_P, _P, _I, _SZ, _P, _P]),
    "h2b_fr_transpose_batch": (_I, [_P, _P, _P, _U32, _U32, _SZ, _U32, _SZ, _SZ]),
    "h2b_fr_transpose_scatter": (_I, [_P, _P, C.POINTER(_P), _U32, _U32, _U32, _U32]),
    "h2b_best_fft_rows_scatter": (_I, [_P, _P, _P, _U32, _U32, C.POINTER(_P), _U32, _U64, _U64, _P, _U32]),
    "h2b_fr_permute3": (_I, [_P, _P, _P, _U32, _U32, _U32]),
    "h2b_fr_twiddle_rows": (_I, [_P, _P, _P, _U32, _U64, _U32, _U32]),
    "h2b_device_alloc": (_I, [_P, _SZ, C.POINTER(_P)]),
    "h2b_device_free": (None, [_P, _P]),
    "h2b_host_alloc": (_I, [_SZ, C.POINTER(_P)]),
    "h2b_host_free": (None, [_P]),
    "h2b_copy_h2d": (_I, [_P, _P, _P, _SZ]),
    "h2b_copy_d2h": (_I, [_P, _P, _P, _SZ]),
    "h2b_copy_d2d": (_I, [_P, _P, _P, _SZ]),
    "h2b_device_memset": (_I, [_P, _P, _I, _SZ]),
    "h2b_graph_new": (_I, [_P, _P, _SZ, _P, _U32, _P, _U32, _U32, C.POINTER(_P)]),
    "h2b_graph_free": (None, [_P]),
    "h2b_graph_num_slots": (_U32, [_P]),
    "h2b_graph_num_instructions": (_U32, [_P]),
    "h2b_evaluate_h_gates": (_I, [_P, _P, _P, _P]),
    "h2b_evaluate_h_permutation": (_I, [_P, _P, _P, _P, _U32, _P, _P, _U32, _U32, _U32, _P, _P, _P, _P]),
    "h2b_evaluate_h_lookup": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "h2b_fr_from_u512": (_I, [_P, _P, _I, _SZ, _P]),
    "h2b_fr_random_counter": (_I, [_P, _U64, _U64, _SZ, _P]),
    "h2b_permutation_fractions": (_I, [_P, _P, _P, _U32, _U32, _P, _P, _P]),
    "h2b_lookup_permute": (_I, [_P, _P, _P, _SZ, _P, _P]),
    "h2b_lookup_product_fractions": (_I, [_P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "h2b_graph_evaluate_lagrange": (_I, [_P, _P, _P, _P]),
    "h2b_g1_check_on_curve": (_I, [_P, _P, _SZ, C.POINTER(_I)]),
    "h2b_g1_compress": (_I, [_P, _P, _SZ, _U32, _P]),
    "h2b_g1_decompress": (_I, [_P, _P, _SZ, _U32, _P, C.POINTER(_I)]),
    "h2b_poly_fma": (_I, [_P, _P, _P, _P, _P, _SZ]),
    "h2b_synth_scalars": (_I, [_P, _P, _SZ, _U64, _U32]),
    "h2b_synth_bases": (_I, [_P, _P, _SZ, _U64]),
    "h2b_synth_base_scalar": (_U64, [_U64, _U64]),
    "h2b_pipe_peak": (_I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "h2b_test_field_op": (_I, [_P, _I, _I, _P, _P, _P, _SZ]),
    "h2b_host_field_op": (_I, [_I, _I, _P, _P, _P, _SZ]),
    "h2b_test_g1_op": (_I, [_P, _I, _P, _P, _P, _SZ]),
    "h2b_host_g1_op": (_I, [_I, _P, _P, _P, _SZ]),
}


def load(path: str | None = None) -> C.CDLL:
    path = path or os.environ.get("H2B_LIB") or DEFAULT_LIB
    if not os.path.exists(path):
        raise H2BError(H2B_ERR_CUDA, f"{path} not found: build it with __graft_entry__.build() "
                       "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    return lib
