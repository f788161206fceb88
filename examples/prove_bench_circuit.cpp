// The reference's benchmark circuit (halo2_proofs/benches/plonk.rs: StandardPlonk MyCircuit) proved end to end from
// C++ through include/halo2_b200.hpp + include/halo2_b200_plonk.hpp -- what halo2_proofs/examples/serialization.rs
// does in Rust: parameters, keygen, an SRS round trip through RawBytes, create_proof.
//
//   g++ -std=c++17 -O2 -Iinclude examples/prove_bench_circuit.cpp -o prove_bench_circuit
//       -Lhalo2-pse_b200/lib -lhalo2b200 -Wl,-rpath,$PWD/halo2-pse_b200/lib        (one command line)
//   ./prove_bench_circuit 16            # k; needs a B200 (there is no CPU fallback)
#include <chrono>
#include <cstdio>
#include <sstream>

#include "halo2_b200_plonk.hpp"

using namespace halo2_proofs;
using namespace halo2_proofs::plonk;

int main(int argc, char** argv) {
  const uint32_t k = argc > 1 ? std::atoi(argv[1]) : 10;
  try {
    // benches/plonk.rs:203-241 -- StandardPlonk::configure
    ConstraintSystem meta;
    meta.set_minimum_degree(5);
    const Column a = meta.advice_column(), b = meta.advice_column(), c = meta.advice_column();
    meta.enable_equality(a), meta.enable_equality(b), meta.enable_equality(c);
    const Column sm = meta.fixed_column(), sa = meta.fixed_column(), sb = meta.fixed_column(), sc = meta.fixed_column();
    const Expression qa = meta.query_advice(a), qb = meta.query_advice(b), qc = meta.query_advice(c);
    const Expression qsa = meta.query_fixed(sa), qsb = meta.query_fixed(sb), qsc = meta.query_fixed(sc), qsm = meta.query_fixed(sm);
    meta.create_gate("Combined add-mult", {qa * qsa + qb * qsb + qa * qb * qsm - (qc * qsc)});
    // benches/plonk.rs:246-270 -- MyCircuit::synthesize: raw_multiply on row 2i, raw_add on row 2i + 1, two copies
    const size_t iters = (size_t(1) << (k - 1)) - 3;
    const Fr x = Fr::from(0xDEADBEEF), x2 = x * x, fin = x2 + x, one = Fr::one(), zero = Fr::zero();
    std::vector<std::vector<Fr>> fixed(4), advice(3);
    std::vector<CopyConstraint> copies;
    for (size_t i = 0; i < iters; ++i) {
      advice[0].push_back(x), advice[1].push_back(x), advice[2].push_back(x2);
      fixed[0].push_back(one), fixed[1].push_back(zero), fixed[2].push_back(zero), fixed[3].push_back(one);
      advice[0].push_back(x), advice[1].push_back(x2), advice[2].push_back(fin);
      fixed[0].push_back(zero), fixed[1].push_back(one), fixed[2].push_back(one), fixed[3].push_back(one);
      copies.push_back({a, 2 * i, a, 2 * i + 1});
      copies.push_back({b, 2 * i + 1, c, 2 * i});
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto t0, auto t1) { return std::chrono::duration<double, std::milli>(t1 - t0).count(); };

    auto t0 = now();
    const auto params = poly::kzg::ParamsKZG::setup(k, Fr::from_raw(0x1234567890ABCDEFull, 0x1234567890ABCDEFull));  // test SRS: never in production
    auto t1 = now();
    const ProvingKey pk = keygen_pk(params, meta, fixed, copies);
    auto t2 = now();
    const uint8_t seed[16] = {7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7};
    XorShiftRng rng(seed);
    transcript::Blake2bWrite transcript;
    create_proof(params, pk, {advice}, {{}}, rng, transcript);
    auto t3 = now();
    std::printf("k = %u: setup %.1f ms, keygen %.1f ms, create_proof %.1f ms, proof %zu bytes\n", k, ms(t0, t1), ms(t1, t2), ms(t2, t3),
                transcript.finalize().size());
    std::printf("vk transcript_repr = %s\n", hex(pk.transcript_repr).c_str());
  } catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
    return 1;
  }
  return 0;
}
