"""Build recipes: the product library (nvcc, sm_100a only) and, for the CPU
test-suite only, the emulator build of the same sources (g++ -DH2B_EMU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libhalo2b200.so")
SOURCES = ["ctx.cu", "ntt.cu", "msm.cu", "poly.cu", "evalh.cu", "prover.cu", "lookup.cu", "serde.cu"]
HEADERS = ["common.cuh", "field.cuh", "shoup_chains.cuh", "ec.cuh", os.path.join(ROOT, "include", "halo2_b200.h")]
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libhalo2b200_emu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "550,177,128"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    return [os.path.join(CSRC, s) for s in SOURCES] + \
           [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]


def build_product(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -> halo2-pse_b200/lib/libhalo2b200.so"""
    if not force and not _stale(LIB, _deps()):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)

    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        # objects are rebuilt per source: a change to one .cu file recompiles that file only
        if not force and not verbose and not os.environ.get("H2B_NVCC_EXTRA") and \
                not _stale(obj, [os.path.join(CSRC, src)] + hdrs):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              os.environ.get("H2B_NVCC_EXTRA", "").split() + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_emulator(force: bool = False) -> str:
    """TEST-ONLY: the same sources as plain C++ on the fiber emulator
    (tests/emu/cuda_runtime.h).  Never loaded by the product."""
    deps = _deps() + [os.path.join(EMU_DIR, "cuda_runtime.h")]
    if not force and not _stale(EMU_LIB, deps):
        return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-DH2B_EMU", "-I" + EMU_DIR, "-fPIC", "-shared",
           "-pthread", "-o", EMU_LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"emulator build failed:\n{r.stderr}")
    return EMU_LIB


if __name__ == "__main__":
    print(build_product(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--emu" in sys.argv:
        print(build_emulator(force=True))
