"""The C++ host side above the C ABI (include/halo2_b200.hpp: the reference's own names -- best_multiexp,
best_fft, EvaluationDomain, Polynomial, ParamsKZG::commit / commit_lagrange): (1) its self-checking tests,
written after the reference's own unit tests (tests/cpp/test_mirror.cpp); (2) bit-exact parity with the
oracle on seeded inputs through a file-driven CLI (tests/cpp/mirror_cli.cpp).  The CPU suite links the
TEST-ONLY emulator build of the kernels, the `-m gpu` suite the product library on cuda:0."""
import os
import random
import subprocess

import numpy as np
import pytest

import halo2_pse_b200  # noqa: F401
from halo2_pse_b200 import build
from oracle import bn256 as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
OUT = os.path.join(CPP, "_build")


def _build(name: str, lib: str, tag: str) -> str:
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, f"{name}_{tag}")
    src = os.path.join(CPP, name + ".cpp")
    deps = [src, os.path.join(ROOT, "include", "halo2_b200.hpp"), os.path.join(ROOT, "include", "halo2_b200.h"), lib]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        libdir, libname = os.path.dirname(lib), os.path.basename(lib)[3:-3]
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                        "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    return exe


def _run(exe, *args, timeout=600):
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout)


def _self_tests(lib, tag, big_k):
    r = _run(_build("test_mirror", lib, tag), big_k)
    assert r.returncode == 0 and "all mirror tests passed" in r.stdout, r.stdout + r.stderr


def _parity(lib, tag, tmp_path, oc, ks, msm_ns, commit_k):
    cli = _build("mirror_cli", lib, tag)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")

    def call(op, blobs, *args, expect=0):
        with open(fin, "wb") as f:
            for b in blobs:
                f.write(np.ascontiguousarray(b, dtype=np.uint64).tobytes())
        r = _run(cli, op, fin, fout, *args)
        assert r.returncode == expect, (op, args, r.stdout, r.stderr)
        return np.fromfile(fout, dtype=np.uint64) if expect == 0 else r.stdout

    for k in ks:  # best_fft and the domain transforms against the C++ restatement of arithmetic.rs / domain.rs
        n = 1 << k
        a = H.rand_fr_limbs(100 + k, n)
        w = H.fr_enc([O.omega_for(k)])
        assert (call("best_fft", [a, w], k).reshape(-1, 4) == oc.best_fft(a, w[0], k, 0)).all(), k
        for j in (1, 3, 5):
            od = oc.domain(j, k, 0)
            assert (call("lagrange_to_coeff", [a], j, k).reshape(-1, 4) == od.lagrange_to_coeff(a)).all(), (j, k)
            ext = od.coeff_to_extended(a)
            assert (call("coeff_to_extended", [a], j, k).reshape(-1, 4) == ext).all(), (j, k)
            e = H.rand_fr_limbs(200 + k + j, ext.shape[0])
            assert (call("extended_to_coeff", [e], j, k, 0).reshape(-1, 4) == od.extended_to_coeff(e)).all(), (j, k)
            want = od.extended_to_coeff(od.divide_by_vanishing_poly(e))
            for mode in (1, 2):  # two calls, and the fused form
                assert (call("extended_to_coeff", [e], j, k, mode).reshape(-1, 4) == want).all(), (j, k, mode)
            od.free()
    rng = random.Random(77)
    for n in msm_ns:  # best_multiexp against multiexp_serial / best_multiexp of the restatement
        hs = [rng.randrange(1, 1 << 64) for _ in range(n)]
        bases = oc.g1_mul_gen(hs) if n else np.zeros((0, 8), dtype=np.uint64)
        sc = H.rand_fr_limbs(300 + n, n) if n else np.zeros((0, 4), dtype=np.uint64)
        if n > 4:
            sc[1] = 0  # a zero scalar contributes nothing (arithmetic.rs:86)
        got = call("best_multiexp", [sc, bases], n).reshape(1, 8)
        want = oc.best_multiexp(sc, bases, 0) if n else np.zeros((1, 8), dtype=np.uint64)
        assert H.g1_dec(got) == H.g1_dec(want), n
    # ParamsKZG::setup + commit / commit_lagrange against the big-integer restatement of kzg/commitment.rs
    s = 0x1234567890ABCDEF1234567890ABCDEF
    op = O.ParamsKZG.setup(commit_k, s)
    poly = H.rand_fr_limbs(400, 1 << commit_k)
    pts = H.g1_dec(call("commit", [H.fr_enc([s]), poly], commit_k).reshape(4, 8))
    vals = H.fr_dec(poly)
    assert pts[0] == op.commit(vals) and pts[1] == op.commit_lagrange(vals)
    assert pts[2] == op.g[1] and pts[3] == op.g_lagrange[0]
    # contract violations exit like a Rust panic (status 101) with the reference's file:line
    out = call("best_fft", [H.rand_fr_limbs(1, 8), H.fr_enc([3])], 3, expect=101)
    assert "panic" in out
    out = call("lagrange_to_coeff", [H.rand_fr_limbs(1, 8)], 6, 27, expect=101)  # extended_k = 30 > Fr::S
    assert "extended_k" in out
    out = call("lagrange_to_coeff", [H.rand_fr_limbs(1, 8)], 3, 4, expect=2)  # 8 elements for a 2^4 domain: the CLI's own I/O error
    assert "too short" in out


def _transcript_and_gwc(lib, tag, tmp_path, ctx, k):
    """Blake2bWrite / Challenge255 and ProverGWC of the C++ mirror against the Python mirror's (whose proofs are
    byte-equal to the big-integer oracle's, tests/test_emulator_plonk.py / test_gpu_plonk.py)."""
    import halo2_pse_b200 as h
    from halo2_pse_b200.prover import _Poly
    cli = _build("mirror_cli", lib, tag)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")

    def call(op, blobs, *args):
        with open(fin, "wb") as f:
            for b in blobs:
                f.write(np.ascontiguousarray(b, dtype=np.uint64).tobytes())
        r = _run(cli, op, fin, fout, *args)
        assert r.returncode == 0, (op, r.stdout, r.stderr)
        return open(fout, "rb").read()

    # transcript: the same schedule of absorbs and squeezes on both sides
    rng = random.Random(3)
    sc = [rng.randrange(O.R_MOD) for _ in range(2)]
    pts = [O.g1_mul(O.G1_GEN, rng.randrange(1, O.R_MOD)) for _ in range(2)]
    t = h.Blake2bWrite()
    t.common_scalar(sc[0]), t.write_point(pts[0]), t.write_scalar(t.squeeze_challenge_scalar()), t.write_scalar(sc[1])
    t.common_point(pts[1]), t.write_point(pts[1]), t.write_scalar(t.squeeze_challenge_scalar())
    t.write_scalar(t.squeeze_challenge_scalar())
    for i in range(9):
        t.write_scalar(sc[i & 1])
    t.write_scalar(t.squeeze_challenge_scalar())
    assert call("transcript", [H.fr_enc(sc), H.g1_enc(pts)]) == t.finalize()

    # multi-opening: 3 polynomials, 5 queries at 3 distinct points (first-occurrence order matters)
    n, s = 1 << k, 0x1234567890ABCDEF1234567890ABCDEF
    polys = [H.rand_fr_limbs(900 + i, n) for i in range(3)]
    zs = [rng.randrange(O.R_MOD) for _ in range(3)]
    q_idx, q_pt = [0, 1, 0, 2, 1], [zs[0], zs[0], zs[1], zs[2], zs[1]]
    idx_limbs = np.zeros((5, 4), dtype=np.uint64)
    idx_limbs[:, 0] = q_idx
    got = call("gwc", [H.fr_enc([s])] + polys + [H.fr_enc(q_pt), idx_limbs], k, 3, 5)
    params = h.ParamsKZG.setup(ctx, k, s)
    dev = [_Poly(ctx, ctx.upload_fr(p), n) for p in polys]
    t = h.Blake2bWrite()
    t.common_scalar(7)
    queries = [(z, dev[i]) for i, z in zip(q_idx, q_pt)]
    for z, p in queries:
        t.write_scalar(p.eval(z))
    h.ProverGWC(params).create_proof(None, t, queries)
    assert got == t.finalize()
    # the same queries through SHPLONK (rotation sets {z0, z1}, {z0, z1} again for polynomial 1, {z2})
    got = call("shplonk", [H.fr_enc([s])] + polys + [H.fr_enc(q_pt), idx_limbs], k, 3, 5)
    t = h.Blake2bWrite()
    t.common_scalar(7)
    for z, p in queries:
        t.write_scalar(p.eval(z))
    h.ProverSHPLONK(params).create_proof(None, t, queries)
    assert got == t.finalize()
    for p in dev:
        p.buf.free()
    params.g.free()
    params.g_lagrange.free()


def test_cpp_mirror_pinned_vk_of_the_reference(emu_lib_path, tmp_path):
    """include/halo2_b200_plonk.hpp: `configure()` of the reference's tests/plonk_api.rs rebuilt with the C++
    ConstraintSystem (host-only: no kernel runs) gives, character for character, the reference's golden pinned
    verifying key (tests/golden/pinned_vk_plonk_api.json); its transcript_repr hash equals the oracle's."""
    from oracle import prover as OV
    fx = H.load_golden("pinned_vk_plonk_api.json")
    cli = _build("mirror_cli", emu_lib_path, "emu")
    fin, fout = str(tmp_path / "in.txt"), str(tmp_path / "out.bin")
    pt = lambda xy: "(%s, %s)" % tuple(xy)  # noqa: E731
    lines = [str(fx["k"]), str(fx["extended_k"]), fx["omega"], fx["base_modulus"], fx["scalar_modulus"],
             str(len(fx["fixed_commitments"]))] + [pt(p) for p in fx["fixed_commitments"] + fx["permutation_commitments"]]
    open(fin, "w").write("\n".join(lines) + "\n")
    r = _run(cli, "pinned_vk", fin, fout, 0)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(fout, "rb").read()
    assert out[32:].decode() == fx["debug"]
    assert H.fr_dec(np.frombuffer(out[:32], dtype=np.uint64).reshape(1, 4))[0] == OV.vk_transcript_repr(fx["debug"])


def _plonk_api_cs(extra: bool):
    """MyCircuit::configure of tests/plonk_api.rs:389-470 through the Python mirror (the same statements as
    tests/cpp/mirror_cli.cpp::plonk_api_circuit), optionally with the extra gates of the CLI's `graph 1`."""
    import halo2_pse_b200 as h
    cs = h.ConstraintSystem()
    e, a, b = cs.advice_column(), cs.advice_column(), cs.advice_column()
    sf = cs.fixed_column()
    c, d = cs.advice_column(), cs.advice_column()
    p = cs.instance_column()
    for col in (a, b, c):
        cs.enable_equality(col)
    sm, sa, sb, sc, sp = (cs.fixed_column() for _ in range(5))
    sl = cs.fixed_column()
    a_ = cs.query_advice(a)
    cs.lookup("lookup", [(a_, cs.query_fixed(sl))])
    qd, qa, qsf = cs.query_advice(d, 1), cs.query_advice(a), cs.query_fixed(sf)
    qe, qb, qc = cs.query_advice(e, -1), cs.query_advice(b), cs.query_advice(c)
    qsa, qsb, qsc, qsm = cs.query_fixed(sa), cs.query_fixed(sb), cs.query_fixed(sc), cs.query_fixed(sm)
    cs.create_gate("Combined add-mult", [qa * qsa + qb * qsb + qa * qb * qsm - (qc * qsc) + qsf * (qd * qe)])
    qa, qp, qsp = cs.query_advice(a), cs.query_instance(p), cs.query_fixed(sp)
    cs.create_gate("Public input", [qsp * (qa - qp)])
    for col in (sf, e, d, p, sm, sa, sb, sc, sp):
        cs.enable_equality(col)
    if extra:
        E = h.Expression
        qa, qb, qf = cs.query_advice(a), cs.query_advice(b, 2), cs.query_fixed(sf, -1)
        ch = cs.challenge_usable_after(0)
        one, two, zero = E.constant(1), E.constant(2), E.constant(0)
        cs.create_gate("extra", [qf * (qa * qa * two - qb * 7 + (-qa)) + zero * qa,
                                 (qa + qb) * (qa + qb) * qf - qb * ch + qa * (O.R_MOD - 5),
                                 -(E.constant(3)) + two * qb - (qa * qb - qf) * one,
                                 (qa - zero) * (qb * 1) * (qf * 0 + qa)])
        cs.lookup("l1", [(qb * ch + one, cs.query_fixed(sl))])
    return cs


@pytest.mark.parametrize("extra", [0, 1])
def test_cpp_mirror_graph_evaluator(emu_lib_path, tmp_path, extra):
    """GraphEvaluator / Evaluator::new of include/halo2_b200_plonk.hpp (plonk/evaluation.rs:224-277, 590-690): the
    calculation list, constants, rotations and intermediate count of the custom-gates graph and of every lookup
    graph equal the Python mirror's for the reference's plonk_api circuit (and with extra gates that reach the
    constant folding, doubling, squaring, scaling and challenge branches); the library compiles each list."""
    import struct

    import halo2_pse_b200 as h
    cli = _build("mirror_cli", emu_lib_path, "emu")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    open(fin, "wb").write(b"")
    r = _run(cli, "graph", fin, fout, extra)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(fout, "rb").read()
    ev = h.Evaluator(_plonk_api_cs(bool(extra)))
    off = 0
    for g in [ev.custom_gates] + ev.lookups:
        (nw,) = struct.unpack_from("<I", out, off)
        words = np.frombuffer(out, dtype=np.uint32, count=nw, offset=off + 4)
        off += 4 + 4 * nw
        (nc,) = struct.unpack_from("<I", out, off)
        consts = H.fr_dec(np.frombuffer(out, dtype=np.uint64, count=4 * nc, offset=off + 4).reshape(nc, 4))
        off += 4 + 32 * nc
        (nr,) = struct.unpack_from("<I", out, off)
        rots = list(np.frombuffer(out, dtype=np.int32, count=nr, offset=off + 4))
        off += 4 + 4 * nr
        ni, ninstr = struct.unpack_from("<II", out, off)
        off += 8
        assert list(words) == list(g.encode())
        assert consts == [c % O.R_MOD for c in g.constants] and rots == g.rotations and ni == g.num_intermediates
        assert ninstr > 0
    assert off == len(out)


def test_cpp_mirror_keygen(emu_lib_path, tmp_path):
    """keygen_pk of include/halo2_b200_plonk.hpp on the reference's bench circuit (benches/plonk.rs MyCircuit, built in
    C++ with its copy constraints) at k = 5: the pinned verifying-key string -- fixed and permutation commitments,
    omega, constraint system -- and its hash equal the big-integer oracle's keygen."""
    from tests import plonk_cases as PC
    cli = _build("mirror_cli", emu_lib_path, "emu")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    open(fin, "wb").write(H.fr_enc([PC.S_TOXIC]).tobytes())
    r = _run(cli, "keygen", fin, fout, 5)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(fout, "rb").read()
    _, opk, _ = PC.oracle_bench_proof(5, 0xDEADBEEF, b"\x07" * 16)
    assert out[32:].decode() == opk.debug
    assert H.fr_dec(np.frombuffer(out[:32], dtype=np.uint64).reshape(1, 4))[0] == opk.transcript_repr


def _cpp_proof(lib, tag, tmp_path, k, scheme, seed):
    from tests import plonk_cases as PC
    cli = _build("mirror_cli", lib, tag)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    open(fin, "wb").write(H.fr_enc([PC.S_TOXIC]).tobytes() + seed)
    r = _run(cli, "prove", fin, fout, k, scheme)
    assert r.returncode == 0, r.stdout + r.stderr
    return open(fout, "rb").read()


@pytest.mark.gpu
def test_cpp_mirror_create_proof_gpu(tmp_path):
    """The C++ keygen_pk + create_proof against libhalo2b200.so on cuda:0: proof bytes of the bench circuit equal the
    big-integer oracle's at k = 5 (GWC) and k = 8 (SHPLONK), and the restated reference verifier accepts them."""
    from oracle import prover as OV
    from tests import plonk_cases as PC
    seed = b"\x07" * 16
    lib = build.build_product()
    oparams, opk, want = PC.oracle_bench_proof(5, 0xDEADBEEF, seed)
    assert _cpp_proof(lib, "gpu", tmp_path, 5, 0, seed) == want
    oparams, opk, _ = PC.oracle_bench_proof(8, 0xDEADBEEF, seed)
    _, advice, _ = PC.bench_circuit(8, 0xDEADBEEF)
    t = OV.Blake2bWrite()
    OV.create_proof(oparams, opk, [lambda phase, ch: dict(enumerate(advice))], [[]], OV.XorShiftRng(seed), t,
                    multiopen="shplonk")
    got = _cpp_proof(lib, "gpu", tmp_path, 8, 1, seed)
    assert got == t.finalize() and OV.verify_proof(oparams, PC.S_TOXIC, opk, [[]], got, multiopen="shplonk")


@pytest.mark.parametrize("k", [5, 6])
def test_cpp_mirror_create_proof_with_a_lookup(emu_lib_path, tmp_path, k):
    """The C++ create_proof on the lookup circuit of tests/plonk_cases.py (a gate, copy constraints, one lookup
    argument): permute_expression_pair, the lookup grand product and its five constraints in h(X), evaluations and
    openings in the reference's order -- same verifying key and same proof bytes as the big-integer oracle; an
    input outside the table is Error::ConstraintSystemFailure (exit status 101)."""
    import struct

    from oracle import prover as OV
    from tests import plonk_cases as PC
    seed = b"\x21" * 16
    fixed, advice, copies = PC.lookup_circuit(k)
    oparams = O.ParamsKZG.setup(k, PC.S_TOXIC)
    opk = OV.keygen(oparams, PC.oracle_cs(PC.build_lookup_cs()), fixed, copies)
    t = OV.Blake2bWrite()
    OV.create_proof(oparams, opk, [lambda phase, ch: dict(enumerate(advice))], [[]], OV.XorShiftRng(seed), t)
    want = t.finalize()
    cli = _build("mirror_cli", emu_lib_path, "emu")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")

    def run(adv):
        usable = len(fixed[0])
        blob = H.fr_enc([PC.S_TOXIC]).tobytes() + seed + struct.pack("<Q", usable)
        for col in list(fixed) + list(adv):
            blob += H.fr_enc(col).tobytes()
        blob += struct.pack("<Q", len(copies))
        for (lc, lr, rc, rr) in copies:
            blob += struct.pack("<4Q", lc[1], lr, rc[1], rr)
        open(fin, "wb").write(blob)
        return _run(cli, "prove_lookup", fin, fout, k)

    r = run(advice)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(fout, "rb").read()
    assert out[len(want):].decode() == opk.debug
    got = out[:len(want)]
    assert got == want, [i // 32 for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]][:6]
    assert OV.verify_proof(oparams, PC.S_TOXIC, opk, [[]], got)
    bad = [list(c) for c in advice]
    bad[0][0] = 1 << 40
    r = run(bad)
    assert r.returncode == 101 and "ConstraintSystemFailure" in r.stdout


def test_cpp_mirror_create_proof_with_two_phases(emu_lib_path, emu_ctx, tmp_path):
    """The C++ create_proof with a Witness callback over two advice phases and a challenge in between: same
    verifying key and proof bytes as the oracle (and as the Python mirror)."""
    from tests import plonk_cases as PC
    seed = b"\x42" * 16
    want, debug = PC.check_phases_proof_bytes(emu_ctx, 5, seed)
    cli = _build("mirror_cli", emu_lib_path, "emu")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    open(fin, "wb").write(H.fr_enc([PC.S_TOXIC]).tobytes() + seed)
    r = _run(cli, "prove_phases", fin, fout, 5)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(fout, "rb").read()
    assert out[len(want):].decode() == debug
    assert out[:len(want)] == want


def test_cpp_example_program(emu_lib_path):
    """examples/prove_bench_circuit.cpp builds against the headers and proves the bench circuit (emulator, k = 5);
    the verifying-key hash it prints is the oracle's."""
    from tests import plonk_cases as PC
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "prove_bench_circuit_emu")
    libdir = os.path.dirname(emu_lib_path)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "prove_bench_circuit.cpp"), "-o", exe, "-L" + libdir,
                    "-lhalo2b200_emu", "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    r = _run(exe, 5)
    assert r.returncode == 0 and "proof 768 bytes" in r.stdout, r.stdout + r.stderr
    # the example's SRS secret differs from the tests' S_TOXIC: its key hash is that of the same circuit under its own SRS
    s = (0x1234567890ABCDEF << 64) | 0x1234567890ABCDEF
    from oracle import prover as OV
    fixed, _, copies = PC.bench_circuit(5, 0xDEADBEEF)
    opk = OV.keygen(O.ParamsKZG.setup(5, s), PC.oracle_cs(PC.build_cs("bench")), fixed, copies)
    assert ("vk transcript_repr = 0x%064x" % opk.transcript_repr) in r.stdout


def test_cpp_mirror_fr_random_stream(emu_lib_path, tmp_path):
    """XorShiftRng + Fr::random (from_bytes_wide) of the C++ mirror: 5000 draws equal the Python mirror's -- a raw
    256-bit half of the wide integer may exceed r five times over and must not go through from_raw as it is."""
    from halo2_pse_b200.prover import XorShiftRng, fr_random
    cli = _build("mirror_cli", emu_lib_path, "emu")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    seed = bytes(range(1, 17))
    open(fin, "wb").write(seed)
    r = _run(cli, "rng", fin, fout, 5000)
    assert r.returncode == 0, r.stdout + r.stderr
    rng = XorShiftRng(seed)
    assert H.fr_dec(np.fromfile(fout, dtype=np.uint64).reshape(-1, 4)) == [fr_random(rng) for _ in range(5000)]


@pytest.mark.parametrize("k", [5, 7])
@pytest.mark.parametrize("scheme", [0, 1])
def test_cpp_mirror_create_proof_bytes_equal_the_oracle(emu_lib_path, tmp_path, scheme, k):
    """keygen_pk + create_proof of include/halo2_b200_plonk.hpp on the reference's bench circuit at k = 5 with a
    seeded XorShiftRng: the proof bytes equal the big-integer oracle's (GWC), resp. the Python mirror's SHPLONK
    proof (itself equal to the oracle's, tests/test_emulator_plonk.py), and the oracle's verifier accepts them."""
    from oracle import prover as OV
    from tests import plonk_cases as PC
    seed = b"\x07" * 16
    got = _cpp_proof(emu_lib_path, "emu", tmp_path, k, scheme, seed)
    oparams, opk, want = PC.oracle_bench_proof(k, 0xDEADBEEF, seed)
    if scheme == 0:
        assert got == want, [i for i in range(0, len(want), 32) if got[i:i + 32] != want[i:i + 32]][:4]
        assert OV.verify_proof(oparams, PC.S_TOXIC, opk, [[]], got)
    else:
        t = OV.Blake2bWrite()
        _, advice, _ = PC.bench_circuit(k, 0xDEADBEEF)
        OV.create_proof(oparams, opk, [lambda phase, ch: dict(enumerate(advice))], [[]], OV.XorShiftRng(seed), t,
                        multiopen="shplonk")
        assert got == t.finalize()
        assert OV.verify_proof(oparams, PC.S_TOXIC, opk, [[]], got, multiopen="shplonk")


def _params_files(lib, tag, tmp_path, ctx, k):
    """ParamsKZG::read_custom / write_custom of the C++ mirror in the three SerdeFormats against the Python
    mirror's files (kzg/commitment.rs:142-244): the file written back is the file read, the commitments through
    the loaded SRS are the Python mirror's, a corrupted point is rejected where the format checks points."""
    import struct

    import halo2_pse_b200 as h
    from halo2_pse_b200 import serde
    cli = _build("mirror_cli", lib, tag)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    params = h.ParamsKZG.setup(ctx, k, 0x1234567890ABCDEF1234567890ABCDEF)
    poly = H.rand_fr_limbs(77, 1 << k)
    want = H.g1_enc([params.commit(poly), params.commit_lagrange(poly)]).tobytes()
    fmts = [serde.PROCESSED, serde.RAW_BYTES, serde.RAW_BYTES_UNCHECKED]
    for i, fmt in enumerate(fmts):
        blob = serde.params_to_bytes(params, fmt)
        for j in ([i] if i == 0 else [1, 2]):  # the two raw formats share their bytes
            with open(fin, "wb") as f:
                f.write(struct.pack("<Q", len(blob)) + blob + poly.tobytes())
            r = _run(cli, "params", fin, fout, i, j, k)
            assert r.returncode == 0, (fmt, r.stdout, r.stderr)
            out = open(fout, "rb").read()
            assert out[:-128] == blob and out[-128:] == want, (fmt, j)
        # a point that is not on the curve / not canonical
        bad = bytearray(blob)
        if i == 0:
            bad[4 + 5 * 32 + 31] |= 0x3f   # x >= q: not canonical, whatever the square root would say
        else:
            bad[4 + 5 * 64 + 3] ^= 0x40    # off the curve
        with open(fin, "wb") as f:
            f.write(struct.pack("<Q", len(bad)) + bytes(bad) + poly.tobytes())
        r = _run(cli, "params", fin, fout, i, i, k)
        if fmt == serde.RAW_BYTES_UNCHECKED:
            assert r.returncode == 0  # unchecked means unchecked (helpers.rs:46-50)
        else:
            assert r.returncode == 2 and "invalid point encoding" in r.stdout, (fmt, r.stdout)
    # writing the G2 points in the other encoding would need G2 arithmetic the commit half does not carry
    blob = serde.params_to_bytes(params, serde.RAW_BYTES)
    with open(fin, "wb") as f:
        f.write(struct.pack("<Q", len(blob)) + blob + poly.tobytes())
    assert _run(cli, "params", fin, fout, 1, 0, k).returncode == 101
    params.g.free()
    params.g_lagrange.free()


def test_cpp_mirror_params_files_emulator(emu_lib_path, emu_ctx, tmp_path):
    _params_files(emu_lib_path, "emu", tmp_path, emu_ctx, 5)


@pytest.mark.gpu
def test_cpp_mirror_params_files_gpu(gpu_ctx, tmp_path):
    _params_files(build.build_product(), "gpu", tmp_path, gpu_ctx, 10)


def test_cpp_mirror_transcript_and_gwc_emulator(emu_lib_path, emu_ctx, tmp_path):
    _transcript_and_gwc(emu_lib_path, "emu", tmp_path, emu_ctx, 5)


@pytest.mark.gpu
def test_cpp_mirror_transcript_and_gwc_gpu(gpu_ctx, tmp_path):
    _transcript_and_gwc(build.build_product(), "gpu", tmp_path, gpu_ctx, 12)


def test_cpp_mirror_self_tests_emulator(emu_lib_path):
    _self_tests(emu_lib_path, "emu", 6)


def test_cpp_mirror_parity_emulator(emu_lib_path, oracle_c, tmp_path):
    _parity(emu_lib_path, "emu", tmp_path, oracle_c, ks=(0, 1, 5, 8), msm_ns=(0, 1, 33, 700), commit_k=4)


@pytest.mark.gpu
def test_cpp_mirror_self_tests_gpu():
    _self_tests(build.build_product(), "gpu", 12)


@pytest.mark.gpu
def test_cpp_mirror_parity_gpu(oracle_c, tmp_path):
    # every CLI call is a process with its own CUDA context (~1.5 s): a handful of sizes is plenty here, the
    # ABI itself is swept over every size by tests/test_gpu_parity.py
    _parity(build.build_product(), "gpu", tmp_path, oracle_c, ks=(0, 10, 16), msm_ns=(0, 1000, 70001), commit_k=6)
