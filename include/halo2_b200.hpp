// halo2_b200.hpp -- C++ host side above the C ABI (include/halo2_b200.h): the reference's own
// function and type names for the hot path, so that code written against halo2_proofs reads the
// same here.  The reference is Rust (no cargo/rustc in this image); this header is the compiled-code
// mirror a caller links instead of the Rust crate, and the model for the Rust shim in INTEGRATION.md.
//
//   halo2_proofs::arithmetic::best_multiexp / best_fft / eval_polynomial / kate_division /
//                             compute_inner_product            halo2_proofs/src/arithmetic.rs:132,171,304,348,334
//   halo2_proofs::poly::Polynomial<Basis>, Rotation             halo2_proofs/src/poly.rs:52-72, 229-330
//   halo2_proofs::poly::EvaluationDomain                        halo2_proofs/src/poly/domain.rs:19-480
//   halo2_proofs::poly::kzg::ParamsKZG (commit half)            halo2_proofs/src/poly/kzg/commitment.rs:23-131, 281-334
//   halo2_proofs::transcript::Blake2bWrite (Challenge255)       halo2_proofs/src/transcript.rs:282-514
//   halo2_proofs::poly::kzg::multiopen::ProverGWC               halo2_proofs/src/poly/kzg/multiopen/gwc/prover.rs:24-92
//   halo2_proofs::poly::kzg::multiopen::ProverSHPLONK           halo2_proofs/src/poly/kzg/multiopen/shplonk/prover.rs:94-285
//
// Error behaviour: where the reference panics (assert_eq! on lengths, assert!(bases.len() >= size)),
// these throw halo2_proofs::Panic with the reference's file:line in the message; a CUDA / allocation
// failure throws BackendError.  There is no CPU fallback: every call goes to the CUDA library, and
// without a usable device the first call throws.
//
// Threading: every calling thread gets its own context (stream + scratch) on first use, as the ABI
// asks of concurrent callers (rayon workers call the transforms concurrently in
// plonk/permutation/keygen.rs:214-234); objects that own device state (EvaluationDomain, ParamsKZG)
// are bound to the context of the thread that built them and may be used from any thread (the ABI
// serialises calls per context).
#ifndef HALO2_B200_HPP
#define HALO2_B200_HPP

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "halo2_b200.h"

namespace halo2_proofs {

struct Panic : std::logic_error {  // the reference's panic!/assert!
  using std::logic_error::logic_error;
};
struct BackendError : std::runtime_error {  // no counterpart in the reference (it cannot fail this way)
  using std::runtime_error::runtime_error;
};

namespace detail {
struct Backend {
  h2b_ctx* ctx = nullptr;
  Backend() {
    const char* e = std::getenv("H2B_DEVICE");
    const int rc = h2b_ctx_create(e ? std::atoi(e) : 0, &ctx);
    if (rc != H2B_OK) throw BackendError("h2b_ctx_create failed (no CUDA device? there is no CPU fallback): rc=" + std::to_string(rc));
  }
  ~Backend() {
    if (ctx) h2b_ctx_destroy(ctx);
  }
  Backend(const Backend&) = delete;
  Backend& operator=(const Backend&) = delete;
};
inline Backend& backend() {
  thread_local Backend b;
  return b;
}
inline void check(h2b_ctx* ctx, int rc, const char* panic_msg) {
  if (rc == H2B_OK) return;
  if (rc == H2B_ERR_LENGTH || rc == H2B_ERR_ARG || rc == H2B_ERR_BAD_OMEGA) throw Panic(panic_msg);
  throw BackendError(std::string(ctx ? h2b_last_error(ctx) : "backend error") + " (rc=" + std::to_string(rc) + ")");
}
inline void host_op(int field, int op, const void* a, const void* b, void* out) {
  if (h2b_host_field_op(field, op, static_cast<const h2b_fr*>(a), static_cast<const h2b_fr*>(b ? b : a),
                        static_cast<h2b_fr*>(out), 1) != H2B_OK)
    throw BackendError("h2b_host_field_op");
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------
// bn256::Fr / bn256::G1Affine / bn256::G1 (halo2curves 0.3.1): same bytes as the ABI types.  The
// host-side operators below are for the O(#columns) scalars around the kernels (challenges, omega
// powers); they run the library's host code path, not the GPU.
// ---------------------------------------------------------------------------------------------
template <int FIELD>
struct Field : h2b_fr {
  static Field zero() {
    Field z;
    std::memset(&z, 0, sizeof z);
    return z;
  }
  static Field from_raw(uint64_t l0, uint64_t l1 = 0, uint64_t l2 = 0, uint64_t l3 = 0) {  // canonical limbs (a value < p) -> Montgomery
    Field c, r;
    c.l[0] = l0, c.l[1] = l1, c.l[2] = l2, c.l[3] = l3;
    detail::host_op(FIELD, 4, &c, nullptr, &r);
    return r;
  }
  static Field from(uint64_t v) { return from_raw(v); }
  static Field one() { return from_raw(1); }
  Field to_repr_limbs() const {  // Montgomery -> canonical little-endian limbs (to_repr)
    Field r;
    detail::host_op(FIELD, 5, this, nullptr, &r);
    return r;
  }
  Field operator*(const Field& o) const { return bin(0, o); }
  Field operator+(const Field& o) const { return bin(1, o); }
  Field operator-(const Field& o) const { return bin(2, o); }
  Field operator-() const { return un(6); }
  Field square() const { return un(3); }
  Field invert() const { return un(7); }  // 0 -> 0 (callers check is_zero where CtOption matters)
  Field& operator*=(const Field& o) { return *this = *this * o; }
  Field& operator+=(const Field& o) { return *this = *this + o; }
  Field& operator-=(const Field& o) { return *this = *this - o; }
  bool operator==(const Field& o) const { return std::memcmp(l, o.l, sizeof l) == 0; }
  bool operator!=(const Field& o) const { return !(*this == o); }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  Field pow_vartime(uint64_t e) const {
    Field r = one(), b = *this;
    for (; e; e >>= 1, b = b.square())
      if (e & 1) r *= b;
    return r;
  }

 private:
  Field bin(int op, const Field& o) const {
    Field r;
    detail::host_op(FIELD, op, this, &o, &r);
    return r;
  }
  Field un(int op) const {
    Field r;
    detail::host_op(FIELD, op, this, nullptr, &r);
    return r;
  }
};
using Fr = Field<0>;
using Fq = Field<1>;
static_assert(sizeof(Fr) == 32 && sizeof(Fq) == 32, "Fr/Fq are four u64 limbs");

struct G1Affine {
  Fq x, y;  // identity = (0, 0)
  static G1Affine identity() { return G1Affine{Fq::zero(), Fq::zero()}; }
  static G1Affine generator() { return G1Affine{Fq::from(1), Fq::from(2)}; }
  bool is_identity() const { return x.is_zero() && y.is_zero(); }
  bool operator==(const G1Affine& o) const { return x == o.x && y == o.y; }
  bool operator!=(const G1Affine& o) const { return !(*this == o); }
  G1Affine operator+(const G1Affine& o) const {
    G1Affine r;
    if (h2b_host_g1_op(0, reinterpret_cast<const h2b_g1_affine*>(this), reinterpret_cast<const h2b_g1_affine*>(&o),
                       reinterpret_cast<h2b_g1_affine*>(&r), 1) != H2B_OK)
      throw BackendError("h2b_host_g1_op");
    return r;
  }
};
static_assert(sizeof(G1Affine) == 64, "G1Affine is {x, y}");

struct G1 {  // Jacobian (x/z^2, y/z^3), identity has z = 0
  Fq x, y, z;
  G1Affine to_affine() const {
    if (z.is_zero()) return G1Affine::identity();
    const Fq zi = z.invert(), zi2 = zi.square();
    return G1Affine{x * zi2, y * zi2 * zi};
  }
  bool operator==(const G1& o) const { return to_affine() == o.to_affine(); }
  bool operator!=(const G1& o) const { return !(*this == o); }
};
static_assert(sizeof(G1) == 96, "G1 is {x, y, z}");

// ---------------------------------------------------------------------------------------------
// halo2_proofs::arithmetic
// ---------------------------------------------------------------------------------------------
namespace arithmetic {

/// best_multiexp(coeffs, bases) -> C::Curve                                   arithmetic.rs:132
inline G1 best_multiexp(const Fr* coeffs, size_t n_coeffs, const G1Affine* bases, size_t n_bases) {
  if (n_coeffs != n_bases) throw Panic("assertion failed: `(left == right)` coeffs.len() == bases.len() (arithmetic.rs:133)");
  h2b_ctx* ctx = detail::backend().ctx;
  G1 out;
  detail::check(ctx, h2b_best_multiexp(ctx, coeffs, reinterpret_cast<const h2b_g1_affine*>(bases), n_coeffs,
                                       reinterpret_cast<h2b_g1*>(&out)), "best_multiexp (arithmetic.rs:133)");
  return out;
}
inline G1 best_multiexp(const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
  return best_multiexp(coeffs.data(), coeffs.size(), bases.data(), bases.size());
}

/// small_multiexp(coeffs, bases): shared-doubling double-and-add, host-side            arithmetic.rs:105-125
inline G1 small_multiexp(const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
  if (coeffs.size() > bases.size()) throw Panic("index out of bounds: bases[coeff_idx] (arithmetic.rs:117)");
  G1 out;
  detail::check(nullptr, h2b_small_multiexp(coeffs.data(), reinterpret_cast<const h2b_g1_affine*>(bases.data()), coeffs.size(),
                                            reinterpret_cast<h2b_g1*>(&out)), "small_multiexp");
  return out;
}

/// g_to_lagrange(g_projective, k) -> Vec<C>: inverse FFT over curve points, 1/n, batch_normalize; takes the
/// affine forms (the reference's caller builds the projective vector from them)    arithmetic.rs:277-301
inline std::vector<G1Affine> g_to_lagrange(const std::vector<G1Affine>& g, uint32_t k) {
  if (k >= 64 || g.size() != (size_t(1) << k)) throw Panic("assertion failed: `(left == right)` a.len() == 1 << log_n (arithmetic.rs:184)");
  h2b_ctx* ctx = detail::backend().ctx;
  std::vector<G1Affine> out(g.size());
  detail::check(ctx, h2b_g_to_lagrange(ctx, reinterpret_cast<const h2b_g1_affine*>(g.data()), H2B_HOST, k,
                                       reinterpret_cast<h2b_g1_affine*>(out.data()), H2B_HOST), "g_to_lagrange");
  return out;
}

/// best_fft(a, omega, log_n): in place, natural order                        arithmetic.rs:171
/// omega must be a primitive 2^log_n-th root of unity (every non-bench caller passes one).
inline void best_fft(Fr* a, size_t len, const Fr& omega, uint32_t log_n) {
  if (log_n >= 64 || len != (size_t(1) << log_n)) throw Panic("assertion failed: `(left == right)` a.len() == 1 << log_n (arithmetic.rs:184)");
  h2b_ctx* ctx = detail::backend().ctx;
  detail::check(ctx, h2b_best_fft(ctx, a, H2B_HOST, &omega, log_n), "best_fft: omega is not a primitive 2^log_n-th root of unity");
}
inline void best_fft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n) { best_fft(a.data(), a.size(), omega, log_n); }

/// eval_polynomial(poly, point)                                              arithmetic.rs:304
inline Fr eval_polynomial(const Fr* poly, size_t n, const Fr& point) {
  h2b_ctx* ctx = detail::backend().ctx;
  Fr out = Fr::zero();
  if (n) detail::check(ctx, h2b_eval_polynomial(ctx, poly, H2B_HOST, n, &point, &out), "eval_polynomial");
  return out;
}
inline Fr eval_polynomial(const std::vector<Fr>& poly, const Fr& point) { return eval_polynomial(poly.data(), poly.size(), point); }

/// compute_inner_product(a, b)                                               arithmetic.rs:334
inline Fr compute_inner_product(const std::vector<Fr>& a, const std::vector<Fr>& b) {
  if (a.size() != b.size()) throw Panic("assertion failed: `(left == right)` a.len() == b.len() (arithmetic.rs:336)");
  h2b_ctx* ctx = detail::backend().ctx;
  Fr out = Fr::zero();
  if (!a.empty()) detail::check(ctx, h2b_inner_product(ctx, a.data(), b.data(), H2B_HOST, a.size(), &out), "compute_inner_product");
  return out;
}

/// kate_division(a, b): a(X) / (X - b), remainder dropped                    arithmetic.rs:348
inline std::vector<Fr> kate_division(const std::vector<Fr>& a, const Fr& b) {
  if (a.empty()) throw Panic("attempt to subtract with overflow: a.len() - 1 (arithmetic.rs:354)");
  std::vector<Fr> q(a.size() - 1);
  h2b_ctx* ctx = detail::backend().ctx;
  if (!q.empty()) detail::check(ctx, h2b_kate_division(ctx, a.data(), H2B_HOST, a.size(), &b, q.data()), "kate_division");
  return q;
}
}  // namespace arithmetic

// ---------------------------------------------------------------------------------------------
// halo2_proofs::poly
// ---------------------------------------------------------------------------------------------
namespace poly {

struct Coeff {};                  // poly.rs:52
struct LagrangeCoeff {};          // poly.rs:57
struct ExtendedLagrangeCoeff {};  // poly.rs:63

struct Rotation {  // poly.rs:311
  int32_t v;
  static Rotation cur() { return Rotation{0}; }
  static Rotation prev() { return Rotation{-1}; }
  static Rotation next() { return Rotation{1}; }
};

struct Blind {  // commitment.rs:198 -- drawn for every commitment, ignored by ParamsKZG::commit*
  Fr v = Fr::one();
};

/// Polynomial<F, B>: a Vec<F> tagged with its basis                          poly.rs:69-72
template <class Basis>
struct Polynomial {
  std::vector<Fr> values;
  size_t len() const { return values.size(); }
  size_t num_coeffs() const { return values.size(); }
  Fr& operator[](size_t i) { return values[i]; }
  const Fr& operator[](size_t i) const { return values[i]; }
  std::vector<Fr>::iterator begin() { return values.begin(); }
  std::vector<Fr>::iterator end() { return values.end(); }
  std::vector<Fr>::const_iterator begin() const { return values.begin(); }
  std::vector<Fr>::const_iterator end() const { return values.end(); }

  /// poly + &poly / poly - &poly / poly * scalar                             poly.rs:229, 243, 278
  Polynomial operator+(const Polynomial& rhs) const { return zip(rhs, 0); }
  Polynomial operator-(const Polynomial& rhs) const { return zip(rhs, 1); }
  Polynomial operator*(const Fr& scalar) const {
    Polynomial out = *this;
    h2b_ctx* ctx = detail::backend().ctx;
    if (!out.values.empty()) detail::check(ctx, h2b_poly_scale(ctx, out.values.data(), H2B_HOST, out.values.size(), &scalar), "Polynomial * scalar");
    return out;
  }
  /// rotate (Lagrange basis only in the reference)                           poly.rs:259
  Polynomial rotate(Rotation r) const {
    Polynomial out = *this;
    const size_t n = values.size();
    if (n == 0) return out;
    const size_t s = static_cast<size_t>(r.v < 0 ? -static_cast<int64_t>(r.v) : r.v) % n;
    for (size_t i = 0; i < n; ++i) out.values[i] = values[r.v >= 0 ? (i + s) % n : (i + n - s) % n];
    return out;
  }

 private:
  Polynomial zip(const Polynomial& rhs, int sub) const {
    if (rhs.values.size() != values.size()) throw Panic("Polynomial +/-: lengths differ (poly.rs:229-256)");
    Polynomial out = *this;
    h2b_ctx* ctx = detail::backend().ctx;
    if (!values.empty())
      detail::check(ctx, (sub ? h2b_poly_sub : h2b_poly_add)(ctx, out.values.data(), rhs.values.data(), H2B_HOST, values.size()), "Polynomial +/-");
    return out;
  }
};

/// EvaluationDomain<G> for G = bn256::Fr                                     poly/domain.rs:19-480
class EvaluationDomain {
 public:
  /// EvaluationDomain::new(j, k)                                             domain.rs:39
  EvaluationDomain(uint32_t j, uint32_t k) : ctx_(detail::backend().ctx) {
    h2b_domain* d = nullptr;
    detail::check(ctx_, h2b_domain_new(ctx_, j, k, &d), "EvaluationDomain::new: extended_k exceeds Fr::S = 28 (domain.rs:54-61)");
    dom_.reset(d, h2b_domain_free);
    n_ = uint64_t(1) << h2b_domain_k(d);
    omega = constant(0), omega_inv = constant(1), extended_omega = constant(2), extended_omega_inv = constant(3);
    g_coset = constant(4), g_coset_inv = constant(5), ifft_divisor = constant(6), extended_ifft_divisor = constant(7);
    barycentric_weight = ifft_divisor;  // 1/n (domain.rs:115)
  }
  Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor,
      barycentric_weight;

  uint32_t k() const { return h2b_domain_k(dom_.get()); }                                  // :364
  uint32_t extended_k() const { return h2b_domain_extended_k(dom_.get()); }                // :369
  size_t extended_len() const { return size_t(1) << extended_k(); }                        // :374
  Fr get_omega() const { return omega; }                                                   // :379
  Fr get_omega_inv() const { return omega_inv; }                                           // :385
  Fr get_extended_omega() const { return extended_omega; }                                 // :390
  size_t get_quotient_poly_degree() const { return h2b_domain_quotient_len(dom_.get()) / n_; }  // :463

  Polynomial<LagrangeCoeff> lagrange_from_vec(std::vector<Fr> values) const {              // :147
    if (values.size() != n_) throw Panic("assertion failed: `(left == right)` values.len() == self.n (domain.rs:148)");
    return Polynomial<LagrangeCoeff>{std::move(values)};
  }
  Polynomial<Coeff> coeff_from_vec(std::vector<Fr> values) const {                         // :159
    if (values.size() != n_) throw Panic("assertion failed: `(left == right)` values.len() == self.n (domain.rs:160)");
    return Polynomial<Coeff>{std::move(values)};
  }
  Polynomial<Coeff> empty_coeff() const { return {std::vector<Fr>(n_, Fr::zero())}; }       // :169
  Polynomial<LagrangeCoeff> empty_lagrange() const { return {std::vector<Fr>(n_, Fr::zero())}; }  // :177
  Polynomial<LagrangeCoeff> constant_lagrange(const Fr& s) const { return {std::vector<Fr>(n_, s)}; }  // :197
  Polynomial<ExtendedLagrangeCoeff> empty_extended() const { return {std::vector<Fr>(extended_len(), Fr::zero())}; }  // :206
  Polynomial<ExtendedLagrangeCoeff> constant_extended(const Fr& s) const { return {std::vector<Fr>(extended_len(), s)}; }  // :215

  /// lagrange_to_coeff(a): consumes the Lagrange vector, returns the coefficients        domain.rs:226
  Polynomial<Coeff> lagrange_to_coeff(Polynomial<LagrangeCoeff> a) const {
    if (a.values.size() != n_) throw Panic("assertion failed: `(left == right)` a.values.len() == 1 << self.k (domain.rs:227)");
    detail::check(ctx_, h2b_lagrange_to_coeff(dom_.get(), a.values.data(), H2B_HOST), "lagrange_to_coeff");
    return Polynomial<Coeff>{std::move(a.values)};
  }
  /// coeff_to_extended(a): zeta-coset evaluations over the extended domain               domain.rs:240
  Polynomial<ExtendedLagrangeCoeff> coeff_to_extended(const Polynomial<Coeff>& a) const {
    if (a.values.size() != n_) throw Panic("assertion failed: `(left == right)` a.values.len() == 1 << self.k (domain.rs:244)");
    Polynomial<ExtendedLagrangeCoeff> out{std::vector<Fr>(extended_len())};
    detail::check(ctx_, h2b_coeff_to_extended(dom_.get(), a.values.data(), out.values.data(), H2B_HOST), "coeff_to_extended");
    return out;
  }
  /// rotate_extended(poly, rotation)                                                     domain.rs:257
  Polynomial<ExtendedLagrangeCoeff> rotate_extended(const Polynomial<ExtendedLagrangeCoeff>& p, Rotation r) const {
    const int64_t step = int64_t(1) << (extended_k() - k());
    Polynomial<ExtendedLagrangeCoeff> out{std::vector<Fr>(p.values.size())};
    const size_t n = p.values.size();
    const size_t s = static_cast<size_t>((r.v < 0 ? -int64_t(r.v) : int64_t(r.v)) * step) % (n ? n : 1);
    for (size_t i = 0; i < n; ++i) out.values[i] = p.values[r.v >= 0 ? (i + s) % n : (i + n - s) % n];
    return out;
  }
  /// extended_to_coeff(a) -> Vec<G> of n * quotient_poly_degree coefficients             domain.rs:281
  std::vector<Fr> extended_to_coeff(const Polynomial<ExtendedLagrangeCoeff>& a) const {
    if (a.values.size() != extended_len()) throw Panic("assertion failed: `(left == right)` a.values.len() == self.extended_len() (domain.rs:282)");
    std::vector<Fr> out(h2b_domain_quotient_len(dom_.get()));
    detail::check(ctx_, h2b_extended_to_coeff(dom_.get(), a.values.data(), out.data(), H2B_HOST, 0), "extended_to_coeff");
    return out;
  }
  /// divide_by_vanishing_poly(a)                                                         domain.rs:307
  Polynomial<ExtendedLagrangeCoeff> divide_by_vanishing_poly(Polynomial<ExtendedLagrangeCoeff> a) const {
    if (a.values.size() != extended_len()) throw Panic("assertion failed: `(left == right)` a.values.len() == 1 << self.extended_k (domain.rs:311)");
    detail::check(ctx_, h2b_divide_by_vanishing_poly(dom_.get(), a.values.data(), H2B_HOST), "divide_by_vanishing_poly");
    return a;
  }
  /// the only call order in the reference (plonk/vanishing/prover.rs:84-87), fused into one transform
  std::vector<Fr> divide_by_vanishing_poly_then_extended_to_coeff(const Polynomial<ExtendedLagrangeCoeff>& a) const {
    if (a.values.size() != extended_len()) throw Panic("assertion failed: `(left == right)` a.values.len() == self.extended_len() (domain.rs:282,311)");
    std::vector<Fr> out(h2b_domain_quotient_len(dom_.get()));
    detail::check(ctx_, h2b_extended_to_coeff(dom_.get(), a.values.data(), out.data(), H2B_HOST, 1), "extended_to_coeff");
    return out;
  }
  /// rotate_omega(value, rotation)                                                       domain.rs:396
  Fr rotate_omega(const Fr& value, Rotation r) const {
    return r.v >= 0 ? value * omega.pow_vartime(uint64_t(r.v)) : value * omega_inv.pow_vartime(uint64_t(-int64_t(r.v)));
  }
  /// l_i_range(x, xn, rotations)                                                         domain.rs:435
  std::vector<Fr> l_i_range(const Fr& x, const Fr& xn, const std::vector<int32_t>& rotations) const {
    std::vector<Fr> results;
    results.reserve(rotations.size());
    for (int32_t r : rotations) results.push_back((x - rotate_omega(Fr::one(), Rotation{r})).invert());
    const Fr common = (xn - Fr::one()) * barycentric_weight;
    for (size_t i = 0; i < rotations.size(); ++i) results[i] = rotate_omega(results[i] * common, Rotation{rotations[i]});
    return results;
  }

  h2b_domain* raw() const { return dom_.get(); }

 private:
  Fr constant(uint32_t which) const {
    Fr out;
    detail::check(ctx_, h2b_domain_constant(dom_.get(), which, &out), "h2b_domain_constant");
    return out;
  }
  h2b_ctx* ctx_;
  std::shared_ptr<h2b_domain> dom_;
  uint64_t n_ = 0;
};

}  // namespace poly

/// SerdeFormat                                                              halo2_proofs/src/helpers.rs:8-52
enum class SerdeFormat { Processed, RawBytes, RawBytesUnchecked };

namespace poly {
namespace kzg {
/// ParamsKZG<Bn256>: the commit half (k, n, g, g_lagrange); the bases are uploaded once and stay on
/// the device with their window table                                       poly/kzg/commitment.rs:23-31
class ParamsKZG {
 public:
  /// from the two base vectors of an existing SRS (what read_custom yields, commitment.rs:160-244)
  static ParamsKZG from_parts(uint32_t k, const std::vector<G1Affine>& g, const std::vector<G1Affine>& g_lagrange,
                              bool precompute = true) {
    const size_t n = size_t(1) << k;
    if (g.size() != n || g_lagrange.size() != n) throw Panic("ParamsKZG: g and g_lagrange must hold 2^k points (commitment.rs:108-116)");
    ParamsKZG p;
    p.ctx_ = detail::backend().ctx;
    p.k_ = k;
    p.n_ = n;
    p.g_ = upload(p.ctx_, g.data(), n, H2B_HOST, precompute);
    p.g_lagrange_ = upload(p.ctx_, g_lagrange.data(), n, H2B_HOST, precompute);
    return p;
  }
  /// ParamsKZG::setup(k, rng) with `s = Fr::random(rng)` handed in (MUST NOT be used in production, as
  /// the reference says): g[i] = [s^i] G, g_lagrange[i] = [(s^n - 1)/n * w^i / (s - w^i)] G; the 2n scalar
  /// multiplications run on the GPU                                           commitment.rs:61-129
  static ParamsKZG setup(uint32_t k, const Fr& s, bool precompute = true) {
    if (k > 28) throw Panic("assertion failed: k <= E::Scalar::S (commitment.rs:64)");
    const size_t n = size_t(1) << k;
    Fr root = Fr::from_raw(0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull, 0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull);  // ROOT_OF_UNITY
    for (uint32_t i = k; i < 28; ++i) root = root.square();  // :90-93
    std::vector<Fr> powers(n), lag(n), dens(n), pre(n);
    Fr cur = Fr::one(), w = Fr::one(), run = Fr::one();
    for (size_t i = 0; i < n; ++i) {
      powers[i] = cur, cur *= s;
      dens[i] = s - w, lag[i] = w, w *= root;  // lag holds w^i for now
      pre[i] = run, run *= dens[i];
    }
    const Fr multiplier = (s.pow_vartime(n) - Fr::one()) * Fr::from(n).invert();  // :96
    Fr inv = run.invert();
    for (size_t i = n; i-- > 0;) {
      lag[i] = multiplier * lag[i] * (inv * pre[i]);  // :101-102
      inv *= dens[i];
    }
    ParamsKZG p;
    p.ctx_ = detail::backend().ctx;
    p.k_ = k;
    p.n_ = n;
    for (int which = 0; which < 2; ++which) {
      void* dev = nullptr;
      detail::check(p.ctx_, h2b_device_alloc(p.ctx_, n * sizeof(G1Affine), &dev), "h2b_device_alloc");
      const int rc = h2b_g1_mul_generator(p.ctx_, (which ? lag : powers).data(), H2B_HOST, n, static_cast<h2b_g1_affine*>(dev), H2B_DEVICE);
      std::shared_ptr<h2b_bases> b;
      if (rc == H2B_OK) b = upload(p.ctx_, static_cast<const G1Affine*>(dev), n, H2B_DEVICE, precompute);
      h2b_device_free(p.ctx_, dev);
      detail::check(p.ctx_, rc, "h2b_g1_mul_generator");
      (which ? p.g_lagrange_ : p.g_) = b;
    }
    return p;
  }

  /// ParamsKZG::read_custom(reader, format)                                   commitment.rs:160-244
  /// k (u32 LE), g[2^k], g_lagrange[2^k], g2, s_g2.  G1: Processed = 32-byte compressed points, decompressed
  /// and checked on the GPU; RawBytes = the Montgomery limbs as they are, curve equation checked on the GPU;
  /// RawBytesUnchecked = uploaded as they are.  The two G2 points belong to the verifier: they are kept as the
  /// bytes of the file (64 B each when Processed, 128 B raw) and written back unchanged by write_custom in the
  /// same format.
  static ParamsKZG read_custom(std::istream& reader, SerdeFormat format, bool precompute = true) {
    uint8_t kb[4];
    if (!reader.read(reinterpret_cast<char*>(kb), 4)) throw std::runtime_error("io::Error: unexpected end of file");
    const uint32_t k = uint32_t(kb[0]) | uint32_t(kb[1]) << 8 | uint32_t(kb[2]) << 16 | uint32_t(kb[3]) << 24;
    if (k > 28) throw Panic("ParamsKZG::read: k exceeds Fr::S (commitment.rs:64)");
    ParamsKZG p;
    p.ctx_ = detail::backend().ctx;
    p.k_ = k;
    p.n_ = uint64_t(1) << k;
    p.g_ = read_g1(p.ctx_, reader, p.n_, format, precompute);
    p.g_lagrange_ = read_g1(p.ctx_, reader, p.n_, format, precompute);
    const size_t g2len = format == SerdeFormat::Processed ? 64 : 128;
    p.g2_bytes_.resize(2 * g2len);
    if (!reader.read(reinterpret_cast<char*>(p.g2_bytes_.data()), std::streamsize(2 * g2len)))
      throw std::runtime_error("io::Error: unexpected end of file");
    p.g2_format_ = format;
    return p;
  }
  /// ParamsKZG::write_custom(writer, format)                                  commitment.rs:142-158
  void write_custom(std::ostream& writer, SerdeFormat format) const {
    const bool raw = format != SerdeFormat::Processed;
    if (g2_bytes_.empty() || (g2_format_ != SerdeFormat::Processed) != raw)
      throw Panic("ParamsKZG::write: the G2 points are only held as the bytes of the file they were read from (same point encoding required)");
    const uint8_t kb[4] = {uint8_t(k_), uint8_t(k_ >> 8), uint8_t(k_ >> 16), uint8_t(k_ >> 24)};
    writer.write(reinterpret_cast<const char*>(kb), 4);
    for (const h2b_bases* b : {g_.get(), g_lagrange_.get()}) {
      if (raw) {
        const auto pts = download(b);
        writer.write(reinterpret_cast<const char*>(pts.data()), std::streamsize(pts.size() * sizeof(G1Affine)));
      } else {
        std::vector<uint8_t> out(n_ * 32);
        detail::check(ctx_, h2b_g1_compress(ctx_, static_cast<const h2b_g1_affine*>(h2b_bases_device_ptr(b)), n_, 7, out.data()), "h2b_g1_compress");
        writer.write(reinterpret_cast<const char*>(out.data()), std::streamsize(out.size()));
      }
    }
    writer.write(reinterpret_cast<const char*>(g2_bytes_.data()), std::streamsize(g2_bytes_.size()));
  }

  /// downsize(k): g.truncate(1 << k); g_lagrange = g_to_lagrange(g, k) -- on the device    commitment.rs:267-275
  void downsize(uint32_t k) {
    if (k > k_) throw Panic("assertion failed: k <= self.k (commitment.rs:268)");
    const size_t n = size_t(1) << k;
    const bool table = h2b_bases_table_window_bits(g_.get()) != 0;
    const auto* gp = static_cast<const G1Affine*>(h2b_bases_device_ptr(g_.get()));
    auto new_g = upload(ctx_, gp, n, H2B_DEVICE, table);
    void* dev = nullptr;
    detail::check(ctx_, h2b_device_alloc(ctx_, n * sizeof(G1Affine), &dev), "h2b_device_alloc");
    const int rc = h2b_g_to_lagrange(ctx_, reinterpret_cast<const h2b_g1_affine*>(gp), H2B_DEVICE, k, static_cast<h2b_g1_affine*>(dev), H2B_DEVICE);
    std::shared_ptr<h2b_bases> new_l;
    if (rc == H2B_OK) new_l = upload(ctx_, static_cast<const G1Affine*>(dev), n, H2B_DEVICE, table);
    h2b_device_free(ctx_, dev);
    detail::check(ctx_, rc, "h2b_g_to_lagrange");
    g_ = new_g, g_lagrange_ = new_l, k_ = k, n_ = n;
  }

  uint32_t k() const { return k_; }       // commitment.rs:254
  uint64_t n() const { return n_; }       // commitment.rs:258
  std::vector<G1Affine> get_g() const { return download(g_.get()); }  // commitment.rs:316
  std::vector<G1Affine> get_g_lagrange() const { return download(g_lagrange_.get()); }

  /// commit_lagrange(poly, _)                                                 commitment.rs:281-292
  G1 commit_lagrange(const Polynomial<LagrangeCoeff>& poly, const Blind& = Blind{}) const { return msm(g_lagrange_.get(), poly.values, "assertion failed: bases.len() >= size (commitment.rs:290)"); }
  /// commit(poly, _)                                                          commitment.rs:327-334
  G1 commit(const Polynomial<Coeff>& poly, const Blind& = Blind{}) const { return msm(g_.get(), poly.values, "assertion failed: bases.len() >= size (commitment.rs:332)"); }

 private:
  static std::shared_ptr<h2b_bases> read_g1(h2b_ctx* ctx, std::istream& reader, size_t n, SerdeFormat format, bool precompute) {
    const size_t size = format == SerdeFormat::Processed ? 32 : 64;
    std::vector<uint8_t> raw(n * size);
    if (!reader.read(reinterpret_cast<char*>(raw.data()), std::streamsize(raw.size()))) throw std::runtime_error("io::Error: unexpected end of file");
    int ok = 1;
    std::shared_ptr<h2b_bases> b;
    if (format == SerdeFormat::Processed) {
      void* dev = nullptr;
      detail::check(ctx, h2b_device_alloc(ctx, n * sizeof(G1Affine), &dev), "h2b_device_alloc");
      const int rc = h2b_g1_decompress(ctx, raw.data(), n, 7, static_cast<h2b_g1_affine*>(dev), &ok);
      if (rc == H2B_OK && ok) b = upload(ctx, static_cast<const G1Affine*>(dev), n, H2B_DEVICE, precompute);
      h2b_device_free(ctx, dev);
      detail::check(ctx, rc, "h2b_g1_decompress");
    } else {
      b = upload(ctx, reinterpret_cast<const G1Affine*>(raw.data()), n, H2B_HOST, false);
      if (format == SerdeFormat::RawBytes)
        detail::check(ctx, h2b_g1_check_on_curve(ctx, static_cast<const h2b_g1_affine*>(h2b_bases_device_ptr(b.get())), n, &ok), "h2b_g1_check_on_curve");
      if (ok && precompute) detail::check(ctx, h2b_bases_precompute(ctx, b.get(), 0), "h2b_bases_precompute");
    }
    if (!ok) throw std::runtime_error("io::Error: invalid point encoding (helpers.rs:27-44)");
    return b;
  }
  static std::shared_ptr<h2b_bases> upload(h2b_ctx* ctx, const G1Affine* pts, size_t n, int loc, bool precompute) {
    h2b_bases* b = nullptr;
    detail::check(ctx, h2b_bases_upload(ctx, reinterpret_cast<const h2b_g1_affine*>(pts), n, loc, &b), "h2b_bases_upload");
    std::shared_ptr<h2b_bases> sp(b, h2b_bases_free);
    if (precompute) detail::check(ctx, h2b_bases_precompute(ctx, b, 0), "h2b_bases_precompute");
    return sp;
  }
  std::vector<G1Affine> download(const h2b_bases* b) const {
    std::vector<G1Affine> out(h2b_bases_len(b));
    detail::check(ctx_, h2b_copy_d2h(ctx_, out.data(), h2b_bases_device_ptr(b), out.size() * sizeof(G1Affine)), "h2b_copy_d2h");
    return out;
  }
  G1 msm(const h2b_bases* b, const std::vector<Fr>& scalars, const char* panic_msg) const {
    if (scalars.size() > n_) throw Panic(panic_msg);
    G1 out;
    detail::check(ctx_, h2b_msm(ctx_, b, 0, scalars.data(), H2B_HOST, scalars.size(), reinterpret_cast<h2b_g1*>(&out)), panic_msg);
    return out;
  }
  h2b_ctx* ctx_ = nullptr;
  uint32_t k_ = 0;
  uint64_t n_ = 0;
  std::shared_ptr<h2b_bases> g_, g_lagrange_;
  std::vector<uint8_t> g2_bytes_;  // g2 then s_g2, as read (verifier half)
  SerdeFormat g2_format_ = SerdeFormat::RawBytes;
};
}  // namespace kzg
}  // namespace poly

// ---------------------------------------------------------------------------------------------
// halo2_proofs::transcript -- Blake2bWrite<_, G1Affine, Challenge255<_>>      halo2_proofs/src/transcript.rs:282-514
// Host-only (hashing is inherently sequential); BLAKE2b per RFC 7693 with the "Halo2-Transcript" personalisation.
// ---------------------------------------------------------------------------------------------
namespace transcript {
class Blake2b {  // digest_size 64, no key, personalisation = 16 bytes
 public:
  explicit Blake2b(const char person[16]) {
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    for (int i = 0; i < 8; ++i) h_[i] = iv[i];
    h_[0] ^= 0x01010000ull ^ 64ull;  // depth 1, fanout 1, digest length 64
    uint64_t p0, p1;
    std::memcpy(&p0, person, 8);
    std::memcpy(&p1, person + 8, 8);
    h_[6] ^= p0;
    h_[7] ^= p1;
  }
  void update(const void* data, size_t len) {
    const uint8_t* in = static_cast<const uint8_t*>(data);
    while (len) {
      if (fill_ == 128) {  // a full buffer is compressed only when more input follows (the last block is special)
        t_ += 128;
        compress(false);
        fill_ = 0;
      }
      const size_t take = len < 128 - fill_ ? len : 128 - fill_;
      std::memcpy(buf_ + fill_, in, take);
      fill_ += take, in += take, len -= take;
    }
  }
  void digest(uint8_t out[64]) const {  // of a copy: the state keeps absorbing (hashlib's state.copy().digest())
    Blake2b c = *this;
    c.t_ += c.fill_;
    std::memset(c.buf_ + c.fill_, 0, 128 - c.fill_);
    c.compress(true);
    std::memcpy(out, c.h_, 64);
  }

 private:
  static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  void compress(bool last) {
    static const uint8_t sigma[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    uint64_t m[16], v[16];
    std::memcpy(m, buf_, 128);
    for (int i = 0; i < 8; ++i) v[i] = h_[i], v[i + 8] = iv[i];
    v[12] ^= t_;  // messages are far below 2^64 bytes: the high counter word stays 0
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
      v[a] = v[a] + v[b] + x, v[d] = rotr(v[d] ^ v[a], 32), v[c] = v[c] + v[d], v[b] = rotr(v[b] ^ v[c], 24);
      v[a] = v[a] + v[b] + y, v[d] = rotr(v[d] ^ v[a], 16), v[c] = v[c] + v[d], v[b] = rotr(v[b] ^ v[c], 63);
    };
    for (int r = 0; r < 12; ++r) {
      const uint8_t* s = sigma[r];
      G(0, 4, 8, 12, m[s[0]], m[s[1]]), G(1, 5, 9, 13, m[s[2]], m[s[3]]), G(2, 6, 10, 14, m[s[4]], m[s[5]]), G(3, 7, 11, 15, m[s[6]], m[s[7]]);
      G(0, 5, 10, 15, m[s[8]], m[s[9]]), G(1, 6, 11, 12, m[s[10]], m[s[11]]), G(2, 7, 8, 13, m[s[12]], m[s[13]]), G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; ++i) h_[i] ^= v[i] ^ v[i + 8];
  }
  uint64_t h_[8];
  uint64_t t_ = 0;
  uint8_t buf_[128] = {0};
  size_t fill_ = 0;
};

/// Fr::from_bytes_wide: the 512-bit little-endian integer mod r (halo2curves 0.3.1)    transcript.rs:501
inline Fr fr_from_bytes_wide(const uint8_t b[64]) {
  // four 128-bit digits: each is a canonical value (< r), which is what from_raw takes -- a raw 256-bit half can
  // exceed r five times over, outside the range the Montgomery conversion is built for
  uint64_t w[8];
  std::memcpy(w, b, 64);
  const Fr t128 = Fr::from_raw(0, 0, 1, 0), t256 = t128.square(), t384 = t256 * t128;
  return Fr::from_raw(w[0], w[1]) + Fr::from_raw(w[2], w[3]) * t128 + Fr::from_raw(w[4], w[5]) * t256 + Fr::from_raw(w[6], w[7]) * t384;
}

class Blake2bWrite {
 public:
  Blake2bWrite() : state_("Halo2-Transcript") {}                                         // transcript.rs:296-303
  /// squeeze_challenge_scalar through Challenge255                                      :320-332, :486-514
  Fr squeeze_challenge_scalar() {
    const uint8_t prefix = 0;  // BLAKE2B_PREFIX_CHALLENGE: stays absorbed
    state_.update(&prefix, 1);
    uint8_t d[64];
    state_.digest(d);
    return fr_from_bytes_wide(d);
  }
  void common_point(const G1Affine& p) {                                                   // :334-347
    if (p.is_identity()) throw Panic("cannot write points at infinity to the transcript (transcript.rs:338)");
    const uint8_t prefix = 1;
    const Fq x = p.x.to_repr_limbs(), y = p.y.to_repr_limbs();
    state_.update(&prefix, 1), state_.update(x.l, 32), state_.update(y.l, 32);
  }
  void common_scalar(const Fr& s) {                                                        // :349-355
    const uint8_t prefix = 2;
    const Fr c = s.to_repr_limbs();
    state_.update(&prefix, 1), state_.update(c.l, 32);
  }
  void write_point(const G1Affine& p) {                                                    // :376-383, G1Affine::to_bytes
    common_point(p);
    const Fq x = p.x.to_repr_limbs(), y = p.y.to_repr_limbs();
    uint8_t b[32];
    std::memcpy(b, x.l, 32);
    b[31] |= static_cast<uint8_t>((y.l[0] & 1) << 7);
    writer_.insert(writer_.end(), b, b + 32);
  }
  void write_scalar(const Fr& s) {                                                         // :384-390
    common_scalar(s);
    const Fr c = s.to_repr_limbs();
    const uint8_t* b = reinterpret_cast<const uint8_t*>(c.l);
    writer_.insert(writer_.end(), b, b + 32);
  }
  const std::vector<uint8_t>& finalize() const { return writer_; }                        // :407-410

 private:
  Blake2b state_;
  std::vector<uint8_t> writer_;
};
}  // namespace transcript

// ---------------------------------------------------------------------------------------------
// halo2_proofs::poly::kzg::multiopen::ProverGWC              halo2_proofs/src/poly/kzg/multiopen/gwc/prover.rs:24-92
// ---------------------------------------------------------------------------------------------
namespace poly {
struct ProverQuery {  // poly/query.rs:10-19 (blind omitted: KZG ignores it)
  Fr point;
  const Polynomial<Coeff>* poly;
};
namespace kzg {
namespace multiopen {
class ProverGWC {
 public:
  explicit ProverGWC(const ParamsKZG& params) : params_(params) {}
  /// create_proof(rng, transcript, queries): one witness commitment per distinct point, in first-occurrence
  /// order; the evaluations themselves were written by the caller (plonk/prover.rs:548-595)
  void create_proof(transcript::Blake2bWrite& t, const std::vector<ProverQuery>& queries) const {
    const Fr v = t.squeeze_challenge_scalar();                                             // :58
    std::vector<std::pair<Fr, std::vector<const ProverQuery*>>> sets;                      // construct_intermediate_sets, gwc.rs:36-61
    for (const auto& q : queries) {
      bool found = false;
      for (auto& s : sets)
        if (s.first == q.point) s.second.push_back(&q), found = true;
      if (!found) sets.push_back({q.point, {&q}});
    }
    for (const auto& s : sets) {                                                           // :61-89
      const Fr& z = s.first;
      Polynomial<Coeff> poly_batch = *s.second[0]->poly;
      Fr eval_batch = arithmetic::eval_polynomial(poly_batch.values, z), power = Fr::one();
      for (size_t i = 1; i < s.second.size(); ++i) {
        power *= v;
        poly_batch = poly_batch + (*s.second[i]->poly * power);  // poly_batch * v + poly, unrolled from the back
        eval_batch += arithmetic::eval_polynomial(s.second[i]->poly->values, z) * power;
      }
      if (poly_batch.values.empty()) throw Panic("empty polynomial in a multi-opening");
      poly_batch.values[0] -= eval_batch;                                                  // &poly_batch - eval_batch, poly.rs:298-305
      Polynomial<Coeff> witness{arithmetic::kate_division(poly_batch.values, z)};          // :79
      t.write_point(params_.commit(witness).to_affine());                                  // :80-88
    }
  }

 private:
  const ParamsKZG& params_;
};
}  // namespace multiopen
}  // namespace kzg
}  // namespace poly

/// arithmetic::lagrange_interpolate on a handful of points (host)            halo2_proofs/src/arithmetic.rs:405-458
namespace arithmetic {
inline std::vector<Fr> lagrange_interpolate(const std::vector<Fr>& points, const std::vector<Fr>& evals) {
  if (points.size() != evals.size()) throw Panic("assertion failed: `(left == right)` points.len() == evals.len() (arithmetic.rs:406)");
  if (points.size() == 1) return {evals[0]};
  std::vector<Fr> fin(points.size(), Fr::zero());
  for (size_t j = 0; j < points.size(); ++j) {
    std::vector<Fr> tmp{Fr::one()};
    for (size_t k = 0; k < points.size(); ++k) {
      if (k == j) continue;
      const Fr denom = (points[j] - points[k]).invert();
      std::vector<Fr> next(tmp.size() + 1, Fr::zero());
      for (size_t i = 0; i < tmp.size(); ++i) {  // tmp(X) * (X - x_k) / (x_j - x_k)
        next[i] += tmp[i] * (-(denom * points[k]));
        next[i + 1] += tmp[i] * denom;
      }
      tmp.swap(next);
    }
    for (size_t i = 0; i < fin.size(); ++i) fin[i] += tmp[i] * evals[j];
  }
  return fin;
}
/// evaluate_vanishing_polynomial(roots, z)                                  arithmetic.rs:460-478
inline Fr evaluate_vanishing_polynomial(const std::vector<Fr>& roots, const Fr& z) {
  Fr acc = Fr::one();
  for (const Fr& r : roots) acc *= z - r;
  return acc;
}
}  // namespace arithmetic

// ---------------------------------------------------------------------------------------------
// halo2_proofs::poly::kzg::multiopen::ProverSHPLONK
//            halo2_proofs/src/poly/kzg/multiopen/shplonk.rs:55-134, shplonk/prover.rs:94-285
// ---------------------------------------------------------------------------------------------
namespace poly {
namespace kzg {
namespace multiopen {
class ProverSHPLONK {
 public:
  explicit ProverSHPLONK(const ParamsKZG& params) : params_(params) {}
  void create_proof(transcript::Blake2bWrite& t, const std::vector<ProverQuery>& queries) const {
    using Poly = Polynomial<Coeff>;
    const size_t n = params_.n();
    auto less = [](const Fr& a, const Fr& b) {  // Ord for Fr: the canonical integers (BTreeSet order of the rotation sets)
      const Fr x = a.to_repr_limbs(), y = b.to_repr_limbs();
      for (int i = 3; i >= 0; --i)
        if (x.l[i] != y.l[i]) return x.l[i] < y.l[i];
      return false;
    };
    auto sorted_unique = [&](std::vector<Fr> v) {
      for (size_t i = 1; i < v.size(); ++i)
        for (size_t j = i; j > 0 && less(v[j], v[j - 1]); --j) std::swap(v[j], v[j - 1]);
      std::vector<Fr> u;
      for (const Fr& x : v)
        if (u.empty() || u.back() != x) u.push_back(x);
      return u;
    };
    // construct_intermediate_sets (shplonk.rs:55-134): points per polynomial, polynomials per point set
    std::vector<Fr> all_points;
    std::vector<std::pair<const Poly*, std::vector<Fr>>> by_commitment;  // first-occurrence order, told apart by identity
    for (const auto& q : queries) {
      all_points.push_back(q.point);
      bool found = false;
      for (auto& c : by_commitment)
        if (c.first == q.poly) c.second.push_back(q.point), found = true;
      if (!found) by_commitment.push_back({q.poly, {q.point}});
    }
    const std::vector<Fr> super_points = sorted_unique(all_points);
    struct RotationSet {
      std::vector<Fr> points;                                        // ascending
      std::vector<std::pair<const Poly*, std::vector<Fr>>> coms;     // polynomial, low-degree equivalent r(X)
    };
    std::vector<RotationSet> sets;
    for (auto& c : by_commitment) {
      const std::vector<Fr> pts = sorted_unique(c.second);
      RotationSet* rs = nullptr;
      for (auto& s : sets)
        if (s.points == pts) rs = &s;
      if (!rs) sets.push_back({pts, {}}), rs = &sets.back();
      std::vector<Fr> evals;
      for (const Fr& pt : pts) evals.push_back(arithmetic::eval_polynomial(c.first->values, pt));
      rs->coms.push_back({c.first, arithmetic::lagrange_interpolate(pts, evals)});
    }
    const Fr y = t.squeeze_challenge_scalar();                                             // prover.rs:113
    const Fr v = t.squeeze_challenge_scalar();                                             // :118
    auto fma = [](Poly& acc, const std::vector<Fr>& p, const Fr& w) {  // acc[..p.len()] += p * w
      Poly scaled = Poly{p} * w;
      Poly head{std::vector<Fr>(acc.values.begin(), acc.values.begin() + p.size())};
      head = head + scaled;
      std::copy(head.values.begin(), head.values.end(), acc.values.begin());
    };
    // h(X) = sum_i v^i (sum_j y^j (p_j - r_j)) / Z_i                                      :120-177
    Poly h_x{std::vector<Fr>(n, Fr::zero())};
    Fr pv = Fr::one();
    for (const auto& s : sets) {
      Poly n_x{std::vector<Fr>(n, Fr::zero())};
      Fr py = Fr::one();
      for (const auto& c : s.coms) {
        std::vector<Fr> num = c.first->values;
        for (size_t i = 0; i < c.second.size(); ++i) num[i] -= c.second[i];
        fma(n_x, num, py);
        py *= y;
      }
      std::vector<Fr> q = n_x.values;
      for (const Fr& pt : s.points) q = arithmetic::kate_division(q, pt);                  // div_by_vanishing
      fma(h_x, q, pv);
      pv *= v;
    }
    t.write_point(params_.commit(h_x).to_affine());                                        // :178-180
    const Fr u = t.squeeze_challenge_scalar();                                             // :181
    // l(X) = sum_i v^i z_i(u) sum_j y^j (p_j - r_j(u)) - Z_T(u) h(X)                      :183-232
    Poly l_x{std::vector<Fr>(n, Fr::zero())};
    std::vector<Fr> z_diffs;
    pv = Fr::one();
    for (const auto& s : sets) {
      std::vector<Fr> others;
      for (const Fr& p : super_points) {
        bool in = false;
        for (const Fr& q : s.points) in = in || q == p;
        if (!in) others.push_back(p);
      }
      const Fr z_i = arithmetic::evaluate_vanishing_polynomial(others, u);
      Fr py = Fr::one(), konst = Fr::zero();
      for (const auto& c : s.coms) {
        Fr r_eval = Fr::zero();
        for (size_t i = c.second.size(); i-- > 0;) r_eval = r_eval * u + c.second[i];
        const Fr w = py * z_i * pv;
        fma(l_x, c.first->values, w);
        konst += r_eval * w;
        py *= y;
      }
      l_x.values[0] -= konst;
      z_diffs.push_back(z_i);
      pv *= v;
    }
    fma(l_x, h_x.values, -arithmetic::evaluate_vanishing_polynomial(super_points, u));
    Poly h2{arithmetic::kate_division(l_x.values, u)};                                     // :240
    h2 = h2 * z_diffs.at(0).invert();                                                      // :243-246
    t.write_point(params_.commit(h2).to_affine());                                         // :248-250
  }

 private:
  const ParamsKZG& params_;
};
}  // namespace multiopen
}  // namespace kzg
}  // namespace poly
}  // namespace halo2_proofs

#endif  // HALO2_B200_HPP
