"""GPU parity of the quotient evaluation through the C ABI against the direct
big-integer oracle, plus size-independent properties at the bench size."""
import numpy as np
import pytest

import halo2_pse_b200 as h
from tests import plonk_cases as PC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,k,ncirc", [("bench", 4, 1), ("bench", 7, 1), ("rich", 4, 1), ("rich", 6, 2),
                                             ("gates_only", 5, 1)])
def test_evaluate_h_vs_oracle(gpu_ctx, variant, k, ncirc):
    PC.check_evaluate_h(gpu_ctx, variant, k, seed=200 + k, n_circuits=ncirc)


def test_evaluate_h_slot_overflow_path(gpu_ctx, monkeypatch):
    monkeypatch.setenv("H2B_EVALH_SMEM_CAP", "4096")
    PC.check_evaluate_h(gpu_ctx, "rich", 5, seed=9)
