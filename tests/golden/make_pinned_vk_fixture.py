"""Generates tests/golden/pinned_vk_plonk_api.json from the reference's ONLY golden artefact: the pinned
verifying key of `plonk_api` (halo2_proofs/tests/plonk_api.rs:622-1020, IPA over Vesta).  The test prints the
key with `{:#?}`; the transcript hashes `{:?}` (plonk.rs:192-203), so the fixture stores the compact form: the
pretty text with line breaks, indentation and trailing commas removed (std::fmt's two renderings of the same
Debug tree differ in nothing else).  Run in the build container, where /root/reference exists:
    python tests/golden/make_pinned_vk_fixture.py
"""
import json
import os
import re

SRC = "/root/reference/halo2_proofs/tests/plonk_api.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pinned_vk_plonk_api.json")


def compact(pretty: str) -> str:
    t = " ".join(line.strip() for line in pretty.splitlines() if line.strip())
    t = re.sub(r",\s*\}", " }", t)
    t = re.sub(r",\s*\)", ")", t)
    t = re.sub(r",\s*\]", "]", t)
    t = re.sub(r"\(\s+", "(", t)
    t = re.sub(r"\[\s+", "[", t)
    return t


def main():
    text = open(SRC).read()
    m = re.search(r'r#####"(PinnedVerificationKey \{.*?\})"#####', text, re.S)
    pretty = m.group(1)
    first_line = text[:m.start(1)].count("\n") + 1
    c = compact(pretty)
    pts = re.findall(r"\((0x[0-9a-f]{64}), (0x[0-9a-f]{64})\)", c)
    n_fixed = c[c.index("fixed_commitments:"):c.index("permutation: VerifyingKey")].count("(0x")
    fixture = {
        "source": f"halo2_proofs/tests/plonk_api.rs:{first_line}-{first_line + pretty.count(chr(10))}",
        "what": "format!(\"{:?}\", pk.get_vk().pinned()) of the plonk_api test circuit (k = 5, IPA over Vesta): the "
                "compact rendering of the pinned `{:#?}` text",
        "base_modulus": re.search(r'base_modulus: "(0x[0-9a-f]+)"', c).group(1),
        "scalar_modulus": re.search(r'scalar_modulus: "(0x[0-9a-f]+)"', c).group(1),
        "k": int(re.search(r"\bk: (\d+)", c).group(1)),
        "extended_k": int(re.search(r"extended_k: (\d+)", c).group(1)),
        "omega": re.search(r"omega: (0x[0-9a-f]+)", c).group(1),
        "fixed_commitments": [list(p) for p in pts[:n_fixed]],
        "permutation_commitments": [list(p) for p in pts[n_fixed:]],
        "debug": c,
    }
    with open(OUT, "w") as f:
        json.dump(fixture, f, indent=1)
    print(OUT, len(c), "chars,", len(pts), "points")


if __name__ == "__main__":
    main()
