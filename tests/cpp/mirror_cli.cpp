// File-driven entry to the C++ host mirror (include/halo2_b200.hpp) so the Python test-suite can hand it
// seeded inputs and compare its outputs bit for bit with the oracle:  mirror_cli <op> <in.bin> <out.bin> [args]
#include <cstdio>
#include <fstream>
#include <string>

#include "halo2_b200_plonk.hpp"

using namespace halo2_proofs;

static std::vector<uint8_t> slurp(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) throw std::runtime_error(std::string("cannot read ") + path);
  std::fseek(f, 0, SEEK_END);
  std::vector<uint8_t> b(static_cast<size_t>(std::ftell(f)));
  std::fseek(f, 0, SEEK_SET);
  if (!b.empty() && std::fread(b.data(), 1, b.size(), f) != b.size()) throw std::runtime_error("short read");
  std::fclose(f);
  return b;
}
static void spit(const char* path, const void* p, size_t n) {
  FILE* f = std::fopen(path, "wb");
  if (!f || (n && std::fwrite(p, 1, n, f) != n)) throw std::runtime_error(std::string("cannot write ") + path);
  std::fclose(f);
}
template <class T>
static std::vector<T> take(const std::vector<uint8_t>& b, size_t off_bytes, size_t count) {
  if (off_bytes + count * sizeof(T) > b.size()) throw std::runtime_error("input file too short");
  std::vector<T> v(count);
  if (count) std::memcpy(static_cast<void*>(v.data()), b.data() + off_bytes, count * sizeof(T));
  return v;
}

// tests/plonk_api.rs:389-470 -- MyCircuit::configure of the reference's own test, statement by statement
static plonk::ConstraintSystem plonk_api_circuit() {
  using namespace plonk;
  ConstraintSystem meta;
  const Column e = meta.advice_column(), a = meta.advice_column(), b = meta.advice_column();
  const Column sf = meta.fixed_column();
  const Column c = meta.advice_column(), d = meta.advice_column();
  const Column p = meta.instance_column();
  meta.enable_equality(a), meta.enable_equality(b), meta.enable_equality(c);
  const Column sm = meta.fixed_column(), sa = meta.fixed_column(), sb = meta.fixed_column(), sc = meta.fixed_column(),
               sp = meta.fixed_column();
  const Column sl = meta.lookup_table_column();
  {
    const Expression a_ = meta.query_any(a, 0);
    meta.lookup("lookup", {{a_, sl}});
  }
  {
    const Expression qd = meta.query_advice(d, 1), qa = meta.query_advice(a, 0), qsf = meta.query_fixed(sf, 0);
    const Expression qe = meta.query_advice(e, -1), qb = meta.query_advice(b, 0), qc = meta.query_advice(c, 0);
    const Expression qsa = meta.query_fixed(sa, 0), qsb = meta.query_fixed(sb, 0), qsc = meta.query_fixed(sc, 0),
                     qsm = meta.query_fixed(sm, 0);
    meta.create_gate("Combined add-mult", {qa * qsa + qb * qsb + qa * qb * qsm - (qc * qsc) + qsf * (qd * qe)});
  }
  {
    const Expression qa = meta.query_advice(a, 0), qp = meta.query_instance(p, 0), qsp = meta.query_fixed(sp, 0);
    meta.create_gate("Public input", {qsp * (qa - qp)});
  }
  for (const Column& col : {sf, e, d, p, sm, sa, sb, sc, sp}) meta.enable_equality(col);
  // a second gate set that exercises the rest of the graph compiler: constants, scaling, negation, doubling,
  // squares, repeated sub-expressions, a challenge (not part of the reference's circuit: appended by `extra`)
  return meta;
}

// benches/plonk.rs:203-270 -- MyCircuit (StandardPlonk: a, b, c; sm, sa, sb, sc) laid out by SimpleFloorPlanner:
// iteration i puts raw_multiply on row 2i and raw_add on row 2i + 1, then copies a0 = a1 and b1 = c0
struct BenchCircuit {
  plonk::ConstraintSystem cs;
  std::vector<std::vector<Fr>> fixed, advice;  // [sm, sa, sb, sc], [a, b, c]
  std::vector<plonk::CopyConstraint> copies;
};
static BenchCircuit bench_circuit(uint32_t k, const Fr& a) {
  using namespace plonk;
  BenchCircuit bc;
  ConstraintSystem& meta = bc.cs;
  meta.set_minimum_degree(5);
  const Column ca = meta.advice_column(), cb = meta.advice_column(), cc = meta.advice_column();
  meta.enable_equality(ca), meta.enable_equality(cb), meta.enable_equality(cc);
  const Column sm = meta.fixed_column(), sa = meta.fixed_column(), sb = meta.fixed_column(), sc = meta.fixed_column();
  const Expression qa = meta.query_advice(ca), qb = meta.query_advice(cb), qc = meta.query_advice(cc);
  const Expression qsa = meta.query_fixed(sa), qsb = meta.query_fixed(sb), qsc = meta.query_fixed(sc), qsm = meta.query_fixed(sm);
  meta.create_gate("Combined add-mult", {qa * qsa + qb * qsb + qa * qb * qsm - (qc * qsc)});
  const size_t iters = (size_t(1) << (k - 1)) - 3;
  const Fr a2 = a * a, fin = a2 + a, one = Fr::one(), zero = Fr::zero();
  bc.fixed.assign(4, {}), bc.advice.assign(3, {});
  for (size_t i = 0; i < iters; ++i) {
    // raw_multiply: (a, a, a^2), sa = sb = 0, sc = sm = 1;  raw_add: (a, a^2, a^2 + a), sa = sb = sc = 1, sm = 0
    bc.advice[0].push_back(a), bc.advice[1].push_back(a), bc.advice[2].push_back(a2);
    bc.fixed[0].push_back(one), bc.fixed[1].push_back(zero), bc.fixed[2].push_back(zero), bc.fixed[3].push_back(one);
    bc.advice[0].push_back(a), bc.advice[1].push_back(a2), bc.advice[2].push_back(fin);
    bc.fixed[0].push_back(zero), bc.fixed[1].push_back(one), bc.fixed[2].push_back(one), bc.fixed[3].push_back(one);
    bc.copies.push_back({ca, 2 * i, ca, 2 * i + 1});
    bc.copies.push_back({cb, 2 * i + 1, cc, 2 * i});
  }
  return bc;
}

static void dump_graph(std::ofstream& o, const plonk::GraphEvaluator& g) {
  auto u32 = [&](uint32_t v) { o.write(reinterpret_cast<const char*>(&v), 4); };
  const auto w = g.encode();
  u32(uint32_t(w.size()));
  o.write(reinterpret_cast<const char*>(w.data()), std::streamsize(w.size() * 4));
  u32(uint32_t(g.constants.size()));
  o.write(reinterpret_cast<const char*>(g.constants.data()), std::streamsize(g.constants.size() * 32));
  u32(uint32_t(g.rotations.size()));
  o.write(reinterpret_cast<const char*>(g.rotations.data()), std::streamsize(g.rotations.size() * 4));
  u32(g.num_intermediates);
  const auto compiled = g.compile();  // the library accepts the list
  u32(h2b_graph_num_instructions(compiled.get()));
}

int main(int argc, char** argv) {
  if (argc < 4) return 64;
  const std::string op = argv[1];
  auto arg = [&](int i) { return static_cast<uint32_t>(std::atoi(argv[4 + i])); };
  try {
    const auto in = slurp(argv[2]);
    if (op == "best_fft") {  // in: a[2^k], omega
      const uint32_t k = arg(0);
      auto a = take<Fr>(in, 0, size_t(1) << k);
      const Fr omega = take<Fr>(in, a.size() * 32, 1)[0];
      arithmetic::best_fft(a, omega, k);
      spit(argv[3], a.data(), a.size() * 32);
    } else if (op == "best_multiexp") {  // in: coeffs[n], bases[n]; out: affine sum
      const size_t n = arg(0);
      const auto c = take<Fr>(in, 0, n);
      const auto b = take<G1Affine>(in, n * 32, n);
      const G1Affine r = arithmetic::best_multiexp(c, b).to_affine();
      spit(argv[3], &r, 64);
    } else if (op == "lagrange_to_coeff" || op == "coeff_to_extended") {
      poly::EvaluationDomain d(arg(0), arg(1));
      auto v = take<Fr>(in, 0, size_t(1) << arg(1));
      if (op == "lagrange_to_coeff") {
        const auto r = d.lagrange_to_coeff(d.lagrange_from_vec(v));
        spit(argv[3], r.values.data(), r.len() * 32);
      } else {
        const auto r = d.coeff_to_extended(d.coeff_from_vec(v));
        spit(argv[3], r.values.data(), r.len() * 32);
      }
    } else if (op == "extended_to_coeff") {  // args: j k divide_by_vanishing(0 | 1 two calls | 2 fused)
      poly::EvaluationDomain d(arg(0), arg(1));
      poly::Polynomial<poly::ExtendedLagrangeCoeff> a{take<Fr>(in, 0, d.extended_len())};
      const auto r = arg(2) == 0 ? d.extended_to_coeff(a)
                     : arg(2) == 1 ? d.extended_to_coeff(d.divide_by_vanishing_poly(a))
                                   : d.divide_by_vanishing_poly_then_extended_to_coeff(a);
      spit(argv[3], r.data(), r.size() * 32);
    } else if (op == "commit") {  // in: s, poly[2^k]; out: commit(poly), commit_lagrange(poly), g[1], g_lagrange[0]
      const uint32_t k = arg(0);
      const Fr s = take<Fr>(in, 0, 1)[0];
      const auto v = take<Fr>(in, 32, size_t(1) << k);
      const auto params = poly::kzg::ParamsKZG::setup(k, s);
      const G1Affine out[4] = {params.commit(poly::Polynomial<poly::Coeff>{v}).to_affine(),
                               params.commit_lagrange(poly::Polynomial<poly::LagrangeCoeff>{v}).to_affine(),
                               params.get_g()[1], params.get_g_lagrange()[0]};
      spit(argv[3], out, sizeof out);
    } else if (op == "pinned_vk") {  // in.bin: text lines: k, extended_k, omega, base modulus, scalar modulus, #fixed, points...
      using namespace plonk;
      ConstraintSystem meta = plonk_api_circuit();
      std::ifstream f(argv[2]);
      std::vector<std::string> lines;
      for (std::string line; std::getline(f, line);) lines.push_back(line);
      const size_t nfixed = std::stoul(lines.at(5));
      std::vector<std::string> fixed(lines.begin() + 6, lines.begin() + 6 + nfixed), perm(lines.begin() + 6 + nfixed, lines.end());
      if (meta.degree() != 4 || meta.blinding_factors() != 5) throw std::runtime_error("degree / blinding_factors of the plonk_api circuit");
      const std::string s = pinned_debug(meta, std::stoul(lines.at(0)), std::stoul(lines.at(1)), lines.at(2), fixed, perm, lines.at(3), lines.at(4));
      const Fr repr = vk_transcript_repr(s);
      std::ofstream o(argv[3], std::ios::binary);
      o.write(reinterpret_cast<const char*>(repr.l), 32);
      o << s;
    } else if (op == "keygen") {  // args: k; in: s (Fr); out: transcript_repr (32 B) then the pinned verifying-key string
      const Fr s = take<Fr>(in, 0, 1)[0];
      const auto params = poly::kzg::ParamsKZG::setup(arg(0), s, false);
      const BenchCircuit bc = bench_circuit(arg(0), Fr::from_raw(0xDEADBEEF));
      const plonk::ProvingKey pk = plonk::keygen_pk(params, bc.cs, bc.fixed, bc.copies);
      std::ofstream o(argv[3], std::ios::binary);
      o.write(reinterpret_cast<const char*>(pk.transcript_repr.l), 32);
      o << pk.pinned;
    } else if (op == "prove") {  // args: k scheme(0 GWC, 1 SHPLONK); in: s (Fr), rng seed (16 B); out: proof bytes
      const Fr s = take<Fr>(in, 0, 1)[0];
      const auto seed = take<uint8_t>(in, 32, 16);
      const auto params = poly::kzg::ParamsKZG::setup(arg(0), s, false);
      const BenchCircuit bc = bench_circuit(arg(0), Fr::from_raw(0xDEADBEEF));
      const plonk::ProvingKey pk = plonk::keygen_pk(params, bc.cs, bc.fixed, bc.copies);
      plonk::XorShiftRng rng(seed.data());
      transcript::Blake2bWrite t;
      plonk::create_proof(params, pk, {bc.advice}, {{}}, rng, t, arg(1) ? plonk::Multiopen::SHPLONK : plonk::Multiopen::GWC);
      spit(argv[3], t.finalize().data(), t.finalize().size());
    } else if (op == "prove_lookup") {  // args: k; in: s, seed (16 B), usable (u64), q_mul q_lk t a b [usable each], ncopies (u64), copies (4 x u64: column, row, column, row; advice)
      // tests/plonk_cases.py::build_lookup_cs -- q_mul * (a * a - b) = 0, (q_lk * a) in t
      using namespace plonk;
      const Fr s = take<Fr>(in, 0, 1)[0];
      const auto seed = take<uint8_t>(in, 32, 16);
      const uint64_t usable = take<uint64_t>(in, 48, 1)[0];
      ConstraintSystem meta;
      const Column a = meta.advice_column(), b = meta.advice_column();
      const Column q_mul = meta.fixed_column(), q_lk = meta.fixed_column(), t = meta.fixed_column();
      meta.enable_equality(a), meta.enable_equality(b);
      const Expression qa = meta.query_advice(a), qb = meta.query_advice(b);
      const Expression qm = meta.query_fixed(q_mul), ql = meta.query_fixed(q_lk);
      const Expression qt = meta.query_fixed(t);
      (void)qt;
      meta.create_gate("square", {qm * (qa * qa - qb)});
      meta.lookup("range", {{ql * qa, t}});
      std::vector<std::vector<Fr>> cols;
      for (int c = 0; c < 5; ++c) cols.push_back(take<Fr>(in, 56 + size_t(c) * usable * 32, usable));
      const size_t off = 56 + 5 * usable * 32;
      const uint64_t ncopies = take<uint64_t>(in, off, 1)[0];
      const auto raw = take<uint64_t>(in, off + 8, 4 * ncopies);
      std::vector<CopyConstraint> copies;
      for (uint64_t i = 0; i < ncopies; ++i)
        copies.push_back({Column{Any::Advice, uint32_t(raw[4 * i])}, size_t(raw[4 * i + 1]), Column{Any::Advice, uint32_t(raw[4 * i + 2])}, size_t(raw[4 * i + 3])});
      const auto params = poly::kzg::ParamsKZG::setup(arg(0), s, false);
      const ProvingKey pk = keygen_pk(params, meta, {cols[0], cols[1], cols[2]}, copies);
      XorShiftRng rng(seed.data());
      transcript::Blake2bWrite tr;
      create_proof(params, pk, {{cols[3], cols[4]}}, {{}}, rng, tr);
      std::ofstream o(argv[3], std::ios::binary);
      o.write(reinterpret_cast<const char*>(tr.finalize().data()), std::streamsize(tr.finalize().size()));
      o << pk.pinned;
    } else if (op == "prove_phases") {  // args: k; in: s, seed (16 B).  Two phases: b (phase 1) = a (phase 0) * challenge
      using namespace plonk;
      const Fr s = take<Fr>(in, 0, 1)[0];
      const auto seed = take<uint8_t>(in, 32, 16);
      ConstraintSystem meta;
      const Column a = meta.advice_column(0), b = meta.advice_column(1);
      const Column q = meta.fixed_column();
      const Expression ch = meta.challenge_usable_after(0);
      meta.enable_equality(a), meta.enable_equality(b);
      const Expression qa = meta.query_advice(a), qb = meta.query_advice(b), qq = meta.query_fixed(q);
      meta.create_gate("prod", {qq * (qb - qa * ch)});
      const auto params = poly::kzg::ParamsKZG::setup(arg(0), s, false);
      const size_t usable = params.n() - (meta.blinding_factors() + 1);
      std::vector<Fr> av(usable), qv(usable, Fr::one());
      for (size_t i = 0; i < usable; ++i) av[i] = Fr::from(i + 2);
      av[2] = av[1];
      const ProvingKey pk = keygen_pk(params, meta, {qv}, {{a, 1, a, 2}, {b, 1, b, 2}});
      const Witness witness = [&](size_t, uint32_t phase, const std::vector<Fr>& challenges) {
        std::vector<std::vector<Fr>> cols(2);
        cols[0] = av;
        if (phase == 1) {
          cols[1] = av;
          for (auto& x : cols[1]) x *= challenges.at(0);
        }
        return cols;
      };
      XorShiftRng rng(seed.data());
      transcript::Blake2bWrite tr;
      create_proof(params, pk, witness, {{}}, rng, tr);
      std::ofstream o(argv[3], std::ios::binary);
      o.write(reinterpret_cast<const char*>(tr.finalize().data()), std::streamsize(tr.finalize().size()));
      o << pk.pinned;
    } else if (op == "rng") {  // in: 16-byte seed; out: arg(0) draws of Fr::random from XorShiftRng
      const auto seed = take<uint8_t>(in, 0, 16);
      plonk::XorShiftRng rng(seed.data());
      std::vector<Fr> v(arg(0));
      for (auto& x : v) x = plonk::fr_random(rng);
      spit(argv[3], v.data(), v.size() * 32);
    } else if (op == "graph") {  // arg: 0 = the plonk_api circuit, 1 = with extra gates; out: custom_gates, then every lookup graph
      using namespace plonk;
      ConstraintSystem meta = plonk_api_circuit();
      if (arg(0) == 1) {
        const Column a{Any::Advice, 1}, b{Any::Advice, 2}, f{Any::Fixed, 0};
        const Expression qa = meta.query_advice(a, 0), qb = meta.query_advice(b, 2), qf = meta.query_fixed(f, -1);
        const Expression ch = meta.challenge_usable_after(0);
        const Expression one = Expression::constant(Fr::one()), two = Expression::constant(Fr::from(2)), zero = Expression::constant(Fr::zero());
        meta.create_gate("extra", {qf * (qa * qa * two - qb * Fr::from(7) + (-qa)) + zero * qa,
                                   (qa + qb) * (qa + qb) * qf - qb * ch + qa * (-Fr::from(5)),
                                   -(Expression::constant(Fr::from(3))) + two * qb - (qa * qb - qf) * one,
                                   (qa - zero) * (qb * Fr::one()) * (qf * Fr::zero() + qa)});
        meta.lookup("l1", {{qb * ch + one, Column{Any::Fixed, 6}}});
      }
      const Evaluator ev(meta);
      std::ofstream o(argv[3], std::ios::binary);
      dump_graph(o, ev.custom_gates);
      for (const auto& g : ev.lookups) dump_graph(o, g);
    } else if (op == "params") {  // args: in_format out_format (0 Processed, 1 RawBytes, 2 RawBytesUnchecked) k_poly
      // in.bin: [file length u64][params file][poly 2^k]; out.bin: [the file written back][commit][commit_lagrange]
      const SerdeFormat fmts[3] = {SerdeFormat::Processed, SerdeFormat::RawBytes, SerdeFormat::RawBytesUnchecked};
      uint64_t flen;
      std::memcpy(&flen, in.data(), 8);
      std::ifstream f(argv[2], std::ios::binary);
      f.seekg(8);
      const auto params = poly::kzg::ParamsKZG::read_custom(f, fmts[arg(0)]);
      const auto v = take<Fr>(in, 8 + flen, size_t(1) << params.k());
      std::ofstream o(argv[3], std::ios::binary);
      params.write_custom(o, fmts[arg(1)]);
      const G1Affine c[2] = {params.commit(poly::Polynomial<poly::Coeff>{v}).to_affine(),
                             params.commit_lagrange(poly::Polynomial<poly::LagrangeCoeff>{v}).to_affine()};
      o.write(reinterpret_cast<const char*>(c), sizeof c);
    } else if (op == "transcript") {  // in: s0, s1 (Fr), p0, p1 (G1Affine); a fixed schedule of absorbs and squeezes
      const auto sc = take<Fr>(in, 0, 2);
      const auto pt = take<G1Affine>(in, 64, 2);
      transcript::Blake2bWrite t;
      t.common_scalar(sc[0]);
      t.write_point(pt[0]);
      t.write_scalar(t.squeeze_challenge_scalar());
      t.write_scalar(sc[1]);
      t.common_point(pt[1]);
      t.write_point(pt[1]);
      t.write_scalar(t.squeeze_challenge_scalar());
      t.write_scalar(t.squeeze_challenge_scalar());
      for (int i = 0; i < 9; ++i) t.write_scalar(sc[i & 1]);  // past one 128-byte block between squeezes
      t.write_scalar(t.squeeze_challenge_scalar());
      spit(argv[3], t.finalize().data(), t.finalize().size());
    } else if (op == "gwc" || op == "shplonk") {  // args: k npolys nqueries; in: s, polys, query points, query poly indices (limb 0)
      const uint32_t k = arg(0), np = arg(1), nq = arg(2);
      const size_t n = size_t(1) << k;
      const Fr s = take<Fr>(in, 0, 1)[0];
      std::vector<poly::Polynomial<poly::Coeff>> polys;
      for (uint32_t i = 0; i < np; ++i) polys.push_back({take<Fr>(in, 32 + size_t(i) * n * 32, n)});
      const auto points = take<Fr>(in, 32 + size_t(np) * n * 32, nq);
      const auto idx = take<Fr>(in, 32 + size_t(np) * n * 32 + size_t(nq) * 32, nq);
      const auto params = poly::kzg::ParamsKZG::setup(k, s);
      transcript::Blake2bWrite t;
      t.common_scalar(Fr::from(7));
      std::vector<poly::ProverQuery> queries;
      for (uint32_t q = 0; q < nq; ++q) {
        const auto& pl = polys.at(idx[q].l[0]);
        t.write_scalar(arithmetic::eval_polynomial(pl.values, points[q]));  // the evaluations go first (plonk/prover.rs:548-595)
        queries.push_back({points[q], &pl});
      }
      if (op == "gwc")
        poly::kzg::multiopen::ProverGWC(params).create_proof(t, queries);
      else
        poly::kzg::multiopen::ProverSHPLONK(params).create_proof(t, queries);
      spit(argv[3], t.finalize().data(), t.finalize().size());
    } else {
      return 64;
    }
  } catch (const Panic& e) {
    std::printf("panic: %s\n", e.what());
    return 101;  // the exit status of a Rust panic
  } catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
    return 2;
  }
  return 0;
}
