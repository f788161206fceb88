"""Generates tests/golden/kat_bn256.json.

Two kinds of entries:

* ``definitional`` -- literal bn256 constants and known answers, typed in from
  SURVEY.md section 8c (derived there from the mathematical definition of the
  curve / field, independently of oracle/).  The oracle is CHECKED against
  these (tests/test_oracle.py); they are not produced by it.
* ``vectors`` -- seeded inputs and outputs of the Python big-integer oracle
  (oracle/bn256.py) in the C-ABI boundary encoding (hex of the little-endian
  Montgomery limbs).  The CUDA path and the C++ oracle are checked against them.

The reference itself (Rust, no toolchain in this image) cannot be run to
produce vectors, and holds no bn256 golden bytes of its own.

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bn256 as O  # noqa: E402

DEFINITIONAL = {
    "fr_modulus": "0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001",
    "fq_modulus": "0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47",
    "bn_u": 4965661367192848881,
    "fr_S": 28,
    "fr_generator": 7,
    "root_of_unity": "0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c",
    "root_of_unity_inv": "0x048127174daabc261bbe587180f34361b22625f59115aba70ed3e50a414e6dba",
    "zeta": "0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd",
    "zeta_sq": "0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23",
    "delta": "0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2",
    "two_inv": "0x183227397098d014dc2822db40c0ac2e9419f4243cdcb848a1f0fac9f8000001",
    "fr_R": "0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb",
    "fr_R2": "0x0216d0b17f4e44a58c49833d53bb808553fe3ab1e35c59e31bb8e645ae216da7",
    "fr_inv64": "0xc2e1f593efffffff",
    "fq_R": "0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d",
    "fq_R2": "0x06d89f71cab8351f47ab1eff0a417ff6b5e71911d44501fbf32cfc5b538afa89",
    "fq_inv64": "0x87d20782e4866389",
    "omega": {
        "16": "0x09d2cc4b5782fbe923e49ace3f647643a5f5d8fb89091c3ababd582133584b29",
        "20": "0x2a14464f1ff42de3856402b62520e670745e39fada049d5b2f0e1e3182673378",
        "24": "0x1951441010b2b95a6e47a6075066a50a036f5ba978c050f2821df86636c0facb",
        "26": "0x1dba8b5bdd64ef6ce29a9039aca3c0e524395c43b9227b96c75090cc6cc7ec97",
    },
    "ntt_k2": {
        "omega": "0x30644e72e131a029048b6e193fd841045cea24f6fd736bec231204708f703636",
        "in": [1, 2, 3, 4],
        "out": ["0x0a",
                "0x16789af3a83522eb1969386a2f88c094a419fe246c11f9394",
                "0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593efffffff",
                "0x30644e72e131a02850c6967bfe2f29ab91a061a5812d67470242134d2ee06c69"],
    },
    "g1_generator": [1, 2],
    "g1_multiples": {
        "2": ["0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3",
              "0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4"],
        "3": ["0x0769bf9ac56bea3ff40232bcb1b6bd159315d84715b8e679f2d355961915abf0",
              "0x2ab799bee0489429554fdb7c8d086475319e63b40b9c5b57cdf1ff3dd9fe2261"],
        "8": ["0x08b1d51d23480c10f472f5e93b9cfea88238c121fe155af7043937882c306a63",
              "0x299836713dad3fa34e337aa412466015c366af8ec50b9d7bd05aa74642822021"],
    },
    "msm": [
        {"scalars": [2, 3], "bases_multiples_of_G": [1, 2], "result_multiple_of_G": 8},
        {"scalars": ["r-1", 1], "bases_multiples_of_G": [1, 1], "result_multiple_of_G": 0},
    ],
}


def hx(b: bytes) -> str:
    return b.hex()


def vectors():
    rng = random.Random(0x68616C6F32)
    out = {"encoding": "hex of little-endian Montgomery limbs: Fr 32 B, G1Affine 64 B (identity = zeros)"}
    # NTT
    ntt = []
    for k in (0, 1, 3, 6, 9, 11):
        n = 1 << k
        a = [rng.randrange(O.R_MOD) for _ in range(n)]
        w = O.omega_for(k)
        b = list(a)
        O.best_fft(b, w, k)
        ntt.append({"log_n": k, "omega": hx(O.fr_to_mont_bytes(w)), "in": hx(O.frs_to_bytes(a)),
                    "out": hx(O.frs_to_bytes(b))})
    out["best_fft"] = ntt
    # domain transforms
    dom = []
    for (j, k) in ((5, 4), (3, 5), (2, 3), (4, 6), (9, 3)):
        D = O.EvaluationDomain(j, k)
        a = [rng.randrange(O.R_MOD) for _ in range(1 << k)]
        ext_in = [rng.randrange(O.R_MOD) for _ in range(D.extended_len())]
        dom.append({
            "j": j, "k": k, "extended_k": D.extended_k,
            "omega": hx(O.fr_to_mont_bytes(D.omega)),
            "extended_omega": hx(O.fr_to_mont_bytes(D.extended_omega)),
            "t_evaluations": hx(O.frs_to_bytes(D.t_evaluations)),
            "a": hx(O.frs_to_bytes(a)),
            "lagrange_to_coeff": hx(O.frs_to_bytes(D.lagrange_to_coeff(a))),
            "coeff_to_extended": hx(O.frs_to_bytes(D.coeff_to_extended(a))),
            "ext": hx(O.frs_to_bytes(ext_in)),
            "divide_by_vanishing_poly": hx(O.frs_to_bytes(D.divide_by_vanishing_poly(ext_in))),
            "extended_to_coeff": hx(O.frs_to_bytes(D.extended_to_coeff(ext_in))),
        })
    out["domain"] = dom
    # MSM, including the exceptional cases of the group law
    G = O.G1_GEN
    msm = []

    def case(name, scalars, bases):
        msm.append({"name": name, "scalars": hx(O.frs_to_bytes(scalars)), "bases": hx(O.g1s_to_bytes(bases)),
                    "result": hx(O.g1_to_bytes(O.msm_naive(scalars, bases)))})

    pts = [O.g1_mul(G, rng.randrange(1, O.R_MOD)) for _ in range(48)]
    case("uniform_48", [rng.randrange(O.R_MOD) for _ in range(48)], pts)
    case("single", [rng.randrange(O.R_MOD)], pts[:1])
    case("all_zero_scalars", [0] * 8, pts[:8])
    case("all_equal_scalars", [0x1234567] * 33, pts[:33])
    case("zero_one_scalars", [rng.randrange(2) for _ in range(40)], pts[:40])
    case("r_minus_1", [O.R_MOD - 1] * 5 + [1] * 5, pts[:10])
    case("repeated_base", [rng.randrange(O.R_MOD) for _ in range(16)], [pts[3]] * 16)
    case("p_and_minus_p", [7, 7, 9, 9], [pts[0], O.g1_neg(pts[0]), pts[1], O.g1_neg(pts[1])])
    case("identity_bases", [rng.randrange(O.R_MOD) for _ in range(6)], [None, pts[0], None, pts[1], None, None])
    case("cancels_to_identity", [5, O.R_MOD - 5], [pts[2], pts[2]])
    case("top_bits", [O.R_MOD - 1 - i for i in range(20)], pts[:20])
    out["best_multiexp"] = msm
    # KZG: commit(ifft(a)) == commit_lagrange(a)   (kzg/commitment.rs:361-384), k = 4, seeded s
    P = O.ParamsKZG.setup(4, 0xDEADBEEFCAFE)
    a = [rng.randrange(O.R_MOD) for _ in range(16)]
    D = O.EvaluationDomain(1 + 1, 4)
    coeff = D.lagrange_to_coeff(a)
    cl = P.commit_lagrange(a)
    assert P.commit(coeff) == cl
    out["kzg"] = {"k": 4, "s": hex(0xDEADBEEFCAFE), "g": hx(O.g1s_to_bytes(P.g)),
                  "g_lagrange": hx(O.g1s_to_bytes(P.g_lagrange)), "lagrange": hx(O.frs_to_bytes(a)),
                  "coeff": hx(O.frs_to_bytes(coeff)), "commitment": hx(O.g1_to_bytes(cl))}
    return out


if __name__ == "__main__":
    doc = {"definitional": DEFINITIONAL, "vectors": vectors()}
    path = os.path.join(ROOT, "tests", "golden", "kat_bn256.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    print(path, os.path.getsize(path), "bytes")
