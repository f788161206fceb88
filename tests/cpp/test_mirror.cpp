// Self-checking tests of the C++ host mirror (include/halo2_b200.hpp), written the way the reference's
// own unit tests are: halo2_proofs/src/poly/domain.rs:488-557 (test_rotate, test_l_i),
// halo2_proofs/src/poly/kzg/commitment.rs:361-384 (test_commit_lagrange), plus the panics the
// reference raises on contract violations (arithmetic.rs:133,184; domain.rs:227,244,282,311;
// kzg/commitment.rs:290,332).  Linked against the product library (GPU) or, for the CPU test-suite,
// the emulator build of the same sources -- the header cannot tell the difference.
#include <cstdio>
#include <functional>

#include "halo2_b200.hpp"

using namespace halo2_proofs;
using poly::Rotation;

static int failures = 0;
#define CHECK(cond)                                                    \
  do {                                                                 \
    if (!(cond)) {                                                     \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);      \
      ++failures;                                                      \
    }                                                                  \
  } while (0)

static uint64_t sm64(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static Fr random_fr(uint64_t& s) {  // 128 random bits squared up: enough spread for identities
  Fr a = Fr::from_raw(sm64(s), sm64(s)), b = Fr::from_raw(sm64(s), sm64(s), sm64(s));
  return a * b + b;
}
static bool panics(const std::function<void()>& f) {
  try {
    f();
  } catch (const Panic&) {
    return true;
  }
  return false;
}

static void test_rotate() {  // domain.rs:488-527
  uint64_t s = 1;
  poly::EvaluationDomain domain(1, 3);
  auto poly = domain.empty_lagrange();
  CHECK(poly.len() == 8);
  for (auto& v : poly) v = random_fr(s);
  auto cur = poly.rotate(Rotation::cur()), next = poly.rotate(Rotation::next()), prev = poly.rotate(Rotation::prev());
  auto c = domain.lagrange_to_coeff(poly), ccur = domain.lagrange_to_coeff(cur), cnext = domain.lagrange_to_coeff(next),
       cprev = domain.lagrange_to_coeff(prev);
  const Fr x = random_fr(s);
  CHECK(arithmetic::eval_polynomial(c.values, x) == arithmetic::eval_polynomial(ccur.values, x));
  CHECK(arithmetic::eval_polynomial(c.values, x * domain.omega) == arithmetic::eval_polynomial(cnext.values, x));
  CHECK(arithmetic::eval_polynomial(c.values, x * domain.omega_inv) == arithmetic::eval_polynomial(cprev.values, x));
}

static void test_l_i() {  // domain.rs:529-557, with l_i = lagrange_to_coeff(e_i) for lagrange_interpolate
  uint64_t s = 2;
  poly::EvaluationDomain domain(1, 3);
  std::vector<std::vector<Fr>> l;
  for (int i = 0; i < 8; ++i) {
    auto e = domain.empty_lagrange();
    e[i] = Fr::one();
    l.push_back(domain.lagrange_to_coeff(e).values);
  }
  const Fr x = random_fr(s), xn = x.pow_vartime(8);
  std::vector<int32_t> rot;
  for (int r = -7; r <= 7; ++r) rot.push_back(r);
  const auto ev = domain.l_i_range(x, xn, rot);
  for (int i = 0; i < 8; ++i) {
    CHECK(arithmetic::eval_polynomial(l[i], x) == ev[7 + i]);
    CHECK(arithmetic::eval_polynomial(l[(8 - i) % 8], x) == ev[7 - i]);
  }
}

static void test_commit_lagrange(uint32_t K) {  // kzg/commitment.rs:361-384
  uint64_t s = 3;
  const auto params = poly::kzg::ParamsKZG::setup(K, random_fr(s));
  poly::EvaluationDomain domain(1, K);
  auto a = domain.empty_lagrange();
  for (auto& v : a) v = random_fr(s);
  const auto b = domain.lagrange_to_coeff(a);
  const poly::Blind alpha;
  CHECK(params.commit(b, alpha) == params.commit_lagrange(a, alpha));
  // the same commitment through the one-shot best_multiexp on the downloaded bases
  CHECK(arithmetic::best_multiexp(b.values, params.get_g()) == params.commit(b, alpha));
}

static void test_fft_and_multiexp_definitions() {
  uint64_t s = 4;
  for (uint32_t k : {0u, 1u, 4u, 9u}) {  // best_fft against the O(n^2) definition  sum_j a_j w^(ij)
    const size_t n = size_t(1) << k;
    poly::EvaluationDomain domain(1, k);
    std::vector<Fr> a(n);
    for (auto& v : a) v = random_fr(s);
    std::vector<Fr> got = a;
    arithmetic::best_fft(got, domain.get_omega(), k);
    for (size_t i = 0; i < n; i += (n > 16 ? n / 8 + 1 : 1)) {
      const Fr wi = domain.get_omega().pow_vartime(i);
      CHECK(arithmetic::eval_polynomial(a, wi) == got[i]);  // X[i] = a(w^i)
    }
  }
  // best_multiexp: sum c_i [h_i]G == [sum c_i h_i]G, bases from the setup scalar multiplications
  const auto params = poly::kzg::ParamsKZG::setup(6, Fr::from(5));  // g[i] = [5^i] G
  const auto g = params.get_g();
  std::vector<Fr> c(64);
  Fr acc = Fr::zero(), p = Fr::one();
  for (size_t i = 0; i < 64; ++i) c[i] = random_fr(s), acc += c[i] * p, p *= Fr::from(5);
  const G1 lhs = arithmetic::best_multiexp(c, g);
  const G1 rhs = arithmetic::best_multiexp(std::vector<Fr>{acc}, std::vector<G1Affine>{G1Affine::generator()});
  CHECK(lhs == rhs);
  // zero scalars contribute nothing (arithmetic.rs:86); an all-zero vector gives the identity
  CHECK(arithmetic::best_multiexp(std::vector<Fr>(64, Fr::zero()), g).to_affine().is_identity());
  CHECK(arithmetic::best_multiexp(std::vector<Fr>{}, std::vector<G1Affine>{}).to_affine().is_identity());
  // group law on the host path: g[1] + g[1] == [2 * 5] G
  CHECK(g[1] + g[1] == arithmetic::best_multiexp(std::vector<Fr>{Fr::from(10)}, std::vector<G1Affine>{G1Affine::generator()}).to_affine());
}

static void test_downsize_and_small_multiexp() {  // kzg/commitment.rs:267-275, arithmetic.rs:105-125, 277-301
  const Fr s5 = Fr::from(0x1234567);
  auto big = poly::kzg::ParamsKZG::setup(5, s5, false);
  const auto small = poly::kzg::ParamsKZG::setup(3, s5, false);
  // g_to_lagrange of the first 8 points of the k = 5 SRS is the k = 3 SRS's g_lagrange
  auto g = big.get_g();
  g.resize(8);
  CHECK(arithmetic::g_to_lagrange(g, 3) == small.get_g_lagrange());
  big.downsize(3);
  CHECK(big.k() == 3 && big.n() == 8);
  CHECK(big.get_g() == small.get_g());
  CHECK(big.get_g_lagrange() == small.get_g_lagrange());
  bool panicked = false;
  try { big.downsize(4); } catch (const Panic&) { panicked = true; }
  CHECK(panicked);
  // small_multiexp == best_multiexp on a handful of points, zero and r - 1 scalars included
  uint64_t st = 9;
  std::vector<Fr> c = {random_fr(st), Fr::zero(), Fr::zero() - Fr::one(), random_fr(st)};
  const std::vector<G1Affine> b(g.begin(), g.begin() + 4);
  CHECK(arithmetic::small_multiexp(c, b) == arithmetic::best_multiexp(c, b));
  CHECK(arithmetic::small_multiexp({}, {}).to_affine().is_identity());
}

static void test_extended_round_trip() {  // coeff_to_extended / extended_to_coeff / divide_by_vanishing_poly
  uint64_t s = 5;
  poly::EvaluationDomain domain(5, 6);  // j = 5: extended_k = k + 2
  CHECK(domain.extended_k() == 8 && domain.get_quotient_poly_degree() == 4);
  auto a = domain.empty_coeff();
  for (auto& v : a) v = random_fr(s);
  const auto ext = domain.coeff_to_extended(a);
  CHECK(ext.len() == domain.extended_len());
  // ext[i] = a(zeta * w_ext^i)
  for (size_t i : {size_t(0), size_t(1), size_t(77), size_t(255)})
    CHECK(ext[i] == arithmetic::eval_polynomial(a.values, domain.g_coset * domain.get_extended_omega().pow_vartime(i)));
  const auto back = domain.extended_to_coeff(ext);
  CHECK(back.size() == 4 * 64);
  bool same = true;
  for (size_t i = 0; i < back.size(); ++i) same = same && back[i] == (i < 64 ? a[i] : Fr::zero());
  CHECK(same);
  // h(X) (X^n - 1) on the coset, divided by the vanishing polynomial, comes back as h
  auto h = domain.empty_coeff();
  for (auto& v : h) v = random_fr(s);
  std::vector<Fr> prod(256, Fr::zero());  // h(X) * (X^64 - 1), degree < 128
  for (size_t i = 0; i < 64; ++i) prod[i + 64] += h[i], prod[i] -= h[i];
  auto num = domain.empty_extended();
  {
    poly::EvaluationDomain d8(1, 8);  // evaluate prod on the zeta coset directly
    std::vector<Fr> scaled = prod;
    Fr z = Fr::one();
    for (auto& v : scaled) v *= z, z *= domain.g_coset;
    arithmetic::best_fft(scaled, domain.get_extended_omega(), 8);
    num.values = scaled;
  }
  const auto q1 = domain.extended_to_coeff(domain.divide_by_vanishing_poly(num));
  const auto q2 = domain.divide_by_vanishing_poly_then_extended_to_coeff(num);
  bool ok = q1 == q2;
  for (size_t i = 0; i < q1.size(); ++i) ok = ok && q1[i] == (i < 64 ? h[i] : Fr::zero());
  CHECK(ok);
  // rotate_extended by one row of the original domain = 4 extended rows
  const auto r1 = domain.rotate_extended(ext, Rotation::next());
  CHECK(r1[0] == ext[4] && r1[255] == ext[3]);
}

static void test_poly_ops() {  // poly.rs:229-305, arithmetic.rs:334-367
  uint64_t s = 6;
  poly::EvaluationDomain domain(1, 5);
  auto a = domain.empty_coeff(), b = domain.empty_coeff();
  for (auto& v : a) v = random_fr(s);
  for (auto& v : b) v = random_fr(s);
  const Fr x = random_fr(s), c = random_fr(s);
  const Fr ax = arithmetic::eval_polynomial(a.values, x), bx = arithmetic::eval_polynomial(b.values, x);
  CHECK(arithmetic::eval_polynomial((a + b).values, x) == ax + bx);
  CHECK(arithmetic::eval_polynomial((a - b).values, x) == ax - bx);
  CHECK(arithmetic::eval_polynomial((a * c).values, x) == ax * c);
  Fr ip = Fr::zero();
  for (size_t i = 0; i < 32; ++i) ip += a[i] * b[i];
  CHECK(arithmetic::compute_inner_product(a.values, b.values) == ip);
  // kate_division: a(X) - a(z) = q(X) (X - z)
  const Fr z = random_fr(s);
  const auto q = arithmetic::kate_division(a.values, z);
  CHECK(q.size() == 31);
  CHECK(arithmetic::eval_polynomial(q, x) * (x - z) == ax - arithmetic::eval_polynomial(a.values, z));
}

static void test_panics() {
  poly::EvaluationDomain domain(3, 4);
  std::vector<Fr> three(3, Fr::one());
  std::vector<G1Affine> two(2, G1Affine::generator());
  CHECK(panics([&] { arithmetic::best_multiexp(three, two); }));                       // arithmetic.rs:133
  CHECK(panics([&] { arithmetic::best_fft(three, domain.get_omega(), 2); }));          // arithmetic.rs:184
  CHECK(panics([&] { std::vector<Fr> a(16, Fr::one()); arithmetic::best_fft(a, Fr::from(3), 4); }));  // not a 2^4-th root
  CHECK(panics([&] { domain.lagrange_to_coeff(poly::Polynomial<poly::LagrangeCoeff>{three}); }));      // domain.rs:227
  CHECK(panics([&] { domain.coeff_to_extended(poly::Polynomial<poly::Coeff>{three}); }));              // domain.rs:244
  CHECK(panics([&] { domain.extended_to_coeff(poly::Polynomial<poly::ExtendedLagrangeCoeff>{three}); }));       // domain.rs:282
  CHECK(panics([&] { domain.divide_by_vanishing_poly(poly::Polynomial<poly::ExtendedLagrangeCoeff>{three}); }));  // domain.rs:311
  CHECK(panics([&] { domain.lagrange_from_vec(three); }));                             // domain.rs:148
  CHECK(panics([&] { poly::EvaluationDomain too_big(5, 27); }));                       // extended_k = 29 > S
  const auto params = poly::kzg::ParamsKZG::setup(3, Fr::from(7));
  CHECK(panics([&] { params.commit(poly::Polynomial<poly::Coeff>{std::vector<Fr>(9, Fr::one())}); }));            // commitment.rs:332
  CHECK(panics([&] { params.commit_lagrange(poly::Polynomial<poly::LagrangeCoeff>{std::vector<Fr>(9, Fr::one())}); }));  // :290
  CHECK(panics([&] { poly::kzg::ParamsKZG::setup(29, Fr::from(7)); }));                // commitment.rs:64
  // shorter polynomials commit against the first len bases (commitment.rs:286-291)
  const auto g = params.get_g();
  CHECK(params.commit(poly::Polynomial<poly::Coeff>{std::vector<Fr>(5, Fr::one())}) ==
        arithmetic::best_multiexp(std::vector<Fr>(5, Fr::one()), std::vector<G1Affine>(g.begin(), g.begin() + 5)));
}

int main(int argc, char** argv) {
  const uint32_t big_k = argc > 1 ? std::atoi(argv[1]) : 6;
  try {
    test_rotate();
    test_l_i();
    test_commit_lagrange(4);  // the reference's K = 6 is `big_k` by default
    test_commit_lagrange(big_k);
    test_fft_and_multiexp_definitions();
    test_extended_round_trip();
    test_downsize_and_small_multiexp();
    test_poly_ops();
    test_panics();
  } catch (const std::exception& e) {
    std::printf("FAIL uncaught %s\n", e.what());
    return 2;
  }
  std::printf(failures ? "%d check(s) failed\n" : "all mirror tests passed\n", failures);
  return failures ? 1 : 0;
}
