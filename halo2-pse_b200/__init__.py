"""halo2-pse_b200: B200-native backend for halo2_proofs' MSM and NTT hot paths.

The directory name carries a hyphen (the repo's naming contract); import it as
``halo2_pse_b200`` through the shim at the repository root.
"""
from ._ffi import (H2B_DEVICE, H2B_ERR_ARG, H2B_ERR_BAD_OMEGA, H2B_ERR_CUDA, H2B_ERR_LENGTH,
                   H2B_ERR_OOM, H2B_ERR_CONSTRAINT, H2B_HOST, H2BError, SYMBOLS, DEFAULT_LIB)
from .api import (Bases, Context, DeviceBuffer, EvaluationDomain, ParamsKZG, PinnedArray, Q_MOD, R_MOD,
                  fq_decode, fq_encode, fr_decode, fr_encode, g1_decode, g1_encode,
                  g1_jacobian_to_affine)
from .plonk import (ADVICE, FIXED, INSTANCE, Column, ConstraintSystem, Evaluator, Expression, GraphEvaluator,
                    LookupArgument, PermutationArgument, make_eval_columns)
from .prover import (Blake2bWrite, CounterRng, PermutationAssembly, ProverGWC, ProverSHPLONK, ProvingKey, XorShiftRng, create_proof,
                     fr_random, fr_random_device, keygen, pinned_debug)
from . import circuits
from . import serde
