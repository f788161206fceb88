// halo2_b200_plonk.hpp -- C++ host side, second part: the constraint-system description a prover works from
// and the verifying-key string that seeds every transcript.  Same names as the reference:
//
//   halo2_proofs::plonk::{Any, Column, Expression, ConstraintSystem}      halo2_proofs/src/plonk/circuit.rs:780-2060
//   halo2_proofs::plonk::lookup::Argument, permutation::Argument         plonk/lookup.rs:10-60, plonk/permutation.rs:19-75
//   halo2_proofs::plonk::pinned_debug  = format!("{:?}", vk.pinned())     plonk.rs:192-230, circuit.rs:1399-1448
//   halo2_proofs::plonk::vk_transcript_repr                               plonk.rs:192-203
//
// Bookkeeping only (no selectors, regions or floor planner: the prover reads the constraint system of the
// verifying key, where selectors are already fixed columns, plonk/prover.rs:69-71).  tests/test_cpp_mirror.py
// rebuilds `configure()` of the reference's tests/plonk_api.rs with it and compares the string, character for
// character, with the reference's golden verifying key.
#ifndef HALO2_B200_PLONK_HPP
#define HALO2_B200_PLONK_HPP

#include <algorithm>
#include <cstdio>

#include "halo2_b200.hpp"

namespace halo2_proofs {
namespace plonk {

enum class Any { Advice = 0, Fixed = 1, Instance = 2 };  // declaration order of plonk/circuit.rs `Any`

struct Column {
  Any column_type;
  uint32_t index;
  bool operator==(const Column& o) const { return column_type == o.column_type && index == o.index; }
};

inline std::string hex(const Fr& x) {  // {:?} of a field element: 0x + 64 hex digits, big-endian canonical
  const Fr c = x.to_repr_limbs();
  char buf[67];
  std::snprintf(buf, sizeof buf, "0x%016llx%016llx%016llx%016llx", (unsigned long long)c.l[3], (unsigned long long)c.l[2],
                (unsigned long long)c.l[1], (unsigned long long)c.l[0]);
  return buf;
}

/// Expression<F>                                                          plonk/circuit.rs:780-1100
class Expression {
 public:
  enum Kind { Constant, Fixed, Advice, Instance, Challenge, Negated, Sum, Product, Scaled };
  static Expression constant(const Fr& v) {
    Expression e(Constant);
    e.n_->value = v;
    return e;
  }
  static Expression query(Kind kind, uint32_t query_index, uint32_t column_index, int32_t rotation) {
    Expression e(kind);
    e.n_->a = query_index, e.n_->b = column_index, e.n_->rot = rotation;
    return e;
  }
  static Expression challenge(uint32_t index, uint32_t phase) {
    Expression e(Challenge);
    e.n_->a = index, e.n_->b = phase;
    return e;
  }
  Expression operator-() const { return unary(Negated); }
  Expression operator+(const Expression& r) const { return binary(Sum, r); }
  Expression operator-(const Expression& r) const { return binary(Sum, -r); }  // impl Sub: self + (-rhs)
  Expression operator*(const Expression& r) const { return binary(Product, r); }
  Expression operator*(const Fr& f) const {  // impl Mul<F>
    Expression e = unary(Scaled);
    e.n_->value = f;
    return e;
  }
  Expression square() const { return *this * *this; }
  size_t degree() const {  // :1002-1015
    switch (n_->kind) {
      case Constant: case Challenge: return 0;
      case Fixed: case Advice: case Instance: return 1;
      case Negated: case Scaled: return n_->l->degree();
      case Sum: return std::max(n_->l->degree(), n_->r->degree());
      default: return n_->l->degree() + n_->r->degree();
    }
  }
  std::string debug() const {  // #[derive(Debug)]-shaped: circuit.rs:1017-1100
    auto q = [&](const char* name) {
      return std::string(name) + " { query_index: " + std::to_string(n_->a) + ", column_index: " + std::to_string(n_->b) +
             ", rotation: Rotation(" + std::to_string(n_->rot) + ") }";
    };
    switch (n_->kind) {
      case Constant: return "Constant(" + hex(n_->value) + ")";
      case Fixed: return q("Fixed");
      case Advice: return q("Advice");
      case Instance: return q("Instance");
      case Challenge: return "Challenge(Challenge { index: " + std::to_string(n_->a) + ", phase: Phase(" + std::to_string(n_->b) + ") })";
      case Negated: return "Negated(" + n_->l->debug() + ")";
      case Scaled: return "Scaled(" + n_->l->debug() + ", " + hex(n_->value) + ")";
      case Sum: return "Sum(" + n_->l->debug() + ", " + n_->r->debug() + ")";
      default: return "Product(" + n_->l->debug() + ", " + n_->r->debug() + ")";
    }
  }
  Kind kind() const { return n_->kind; }

 private:
  struct Node {
    Kind kind;
    Fr value = Fr::zero();
    uint32_t a = 0, b = 0;
    int32_t rot = 0;
    std::shared_ptr<const Expression> l, r;
  };
  explicit Expression(Kind k) : n_(std::make_shared<Node>()) { n_->kind = k; }
  Expression unary(Kind k) const {
    Expression e(k);
    e.n_->l = std::make_shared<const Expression>(*this);
    return e;
  }
  Expression binary(Kind k, const Expression& r) const {
    Expression e = unary(k);
    e.n_->r = std::make_shared<const Expression>(r);
    return e;
  }
  std::shared_ptr<Node> n_;
};

namespace lookup {
struct Argument {  // plonk/lookup.rs:10-60
  std::string name;
  std::vector<Expression> input_expressions, table_expressions;
  size_t required_degree() const {
    size_t in = 1, tab = 1;
    for (const auto& e : input_expressions) in = std::max(in, e.degree());
    for (const auto& e : table_expressions) tab = std::max(tab, e.degree());
    return std::max<size_t>(4, 2 + in + tab);
  }
};
}  // namespace lookup

namespace permutation {
struct Argument {  // plonk/permutation.rs:19-75
  std::vector<Column> columns;
  size_t required_degree() const { return 3; }
  void add_column(const Column& c) {
    if (std::find(columns.begin(), columns.end(), c) == columns.end()) columns.push_back(c);
  }
};
}  // namespace permutation

/// ConstraintSystem<F>: the fields a prover reads and the `configure`-time methods that fill them
///                                                                         plonk/circuit.rs:1330-1400, 1516-1640, 1974-2031
class ConstraintSystem {
 public:
  using Query = std::pair<Column, int32_t>;
  size_t num_fixed_columns = 0, num_advice_columns = 0, num_instance_columns = 0, num_challenges = 0;
  std::vector<uint32_t> advice_column_phase, challenge_phase, num_advice_queries;
  std::vector<std::pair<std::string, std::vector<Expression>>> gates;
  std::vector<Query> advice_queries, instance_queries, fixed_queries;
  permutation::Argument permutation;
  std::vector<lookup::Argument> lookups;
  int64_t minimum_degree = -1;  // None

  Column advice_column(uint32_t phase = 0) {
    num_advice_queries.push_back(0), advice_column_phase.push_back(phase);
    return Column{Any::Advice, uint32_t(num_advice_columns++)};
  }
  Column fixed_column() { return Column{Any::Fixed, uint32_t(num_fixed_columns++)}; }
  Column instance_column() { return Column{Any::Instance, uint32_t(num_instance_columns++)}; }
  Column lookup_table_column() { return fixed_column(); }  // :1495-1502
  Expression challenge_usable_after(uint32_t phase) {
    challenge_phase.push_back(phase);
    return Expression::challenge(uint32_t(num_challenges++), phase);
  }
  Expression query_advice(const Column& c, int32_t at = 0) { return Expression::query(Expression::Advice, query_index(c, at), c.index, at); }
  Expression query_fixed(const Column& c, int32_t at = 0) { return Expression::query(Expression::Fixed, query_index(c, at), c.index, at); }
  Expression query_instance(const Column& c, int32_t at = 0) { return Expression::query(Expression::Instance, query_index(c, at), c.index, at); }
  Expression query_any(const Column& c, int32_t at = 0) {
    return c.column_type == Any::Advice ? query_advice(c, at) : c.column_type == Any::Fixed ? query_fixed(c, at) : query_instance(c, at);
  }
  void enable_equality(const Column& c) {  // :1516-1520
    query_index(c, 0);
    permutation.add_column(c);
  }
  void create_gate(const std::string& name, std::vector<Expression> polys) {
    if (polys.empty()) throw Panic("Gates must contain at least one constraint. (circuit.rs:1749)");
    gates.push_back({name, std::move(polys)});
  }
  /// meta.lookup(name, |meta| vec![(input, table_column)]): the closure's queries come first, then the table
  /// column is queried at Rotation::cur (:1640-1664)
  size_t lookup(const std::string& name, const std::vector<std::pair<Expression, Column>>& table_map) {
    lookup::Argument a{name, {}, {}};
    for (const auto& m : table_map) a.input_expressions.push_back(m.first), a.table_expressions.push_back(query_fixed(m.second));
    lookups.push_back(std::move(a));
    return lookups.size() - 1;
  }
  void set_minimum_degree(size_t d) { minimum_degree = int64_t(d); }
  size_t degree() const {  // :1974-2002
    size_t d = permutation.required_degree();
    for (const auto& l : lookups) d = std::max(d, l.required_degree());
    for (const auto& g : gates)
      for (const auto& p : g.second) d = std::max(d, p.degree());
    return std::max<size_t>(d, minimum_degree < 0 ? 1 : size_t(minimum_degree));
  }
  size_t blinding_factors() const {  // :2006-2031
    size_t factors = 1;
    if (!num_advice_queries.empty()) factors = *std::max_element(num_advice_queries.begin(), num_advice_queries.end());
    return std::max<size_t>(3, factors) + 2;
  }
  size_t minimum_rows() const { return blinding_factors() + 3; }

  std::string pinned_debug() const {  // PinnedConstraintSystem, circuit.rs:1399-1448
    static const char* type_name[3] = {"Advice", "Fixed", "Instance"};
    auto list = [](const std::vector<std::string>& v) {
      std::string s = "[";
      for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + v[i];
      return s + "]";
    };
    auto col = [&](const Column& c) { return "Column { index: " + std::to_string(c.index) + ", column_type: " + type_name[int(c.column_type)] + " }"; };
    auto queries = [&](const std::vector<Query>& qs) {
      std::vector<std::string> v;
      for (const auto& q : qs) v.push_back("(" + col(q.first) + ", Rotation(" + std::to_string(q.second) + "))");
      return list(v);
    };
    auto exprs = [&](const std::vector<Expression>& es) {
      std::vector<std::string> v;
      for (const auto& e : es) v.push_back(e.debug());
      return list(v);
    };
    auto phases = [&](const std::vector<uint32_t>& ps) {
      std::vector<std::string> v;
      for (uint32_t p : ps) v.push_back("Phase(" + std::to_string(p) + ")");
      return list(v);
    };
    std::string s = "num_fixed_columns: " + std::to_string(num_fixed_columns) + ", num_advice_columns: " + std::to_string(num_advice_columns) +
                    ", num_instance_columns: " + std::to_string(num_instance_columns) + ", num_selectors: 0";
    if (num_challenges > 0)  // multi-phase fields only when used (circuit.rs:1424-1430)
      s += ", num_challenges: " + std::to_string(num_challenges) + ", advice_column_phase: " + phases(advice_column_phase) +
           ", challenge_phase: " + phases(challenge_phase);
    std::vector<std::string> g, cols, lks;
    for (const auto& gate : gates)
      for (const auto& p : gate.second) g.push_back(p.debug());
    for (const auto& c : permutation.columns) cols.push_back(col(c));
    for (const auto& l : lookups)
      lks.push_back("Argument { input_expressions: " + exprs(l.input_expressions) + ", table_expressions: " + exprs(l.table_expressions) + " }");
    s += ", gates: " + list(g) + ", advice_queries: " + queries(advice_queries) + ", instance_queries: " + queries(instance_queries) +
         ", fixed_queries: " + queries(fixed_queries) + ", permutation: Argument { columns: " + list(cols) + " }, lookups: " + list(lks) +
         ", constants: [], minimum_degree: " + (minimum_degree < 0 ? std::string("None") : "Some(" + std::to_string(minimum_degree) + ")");
    return s;
  }

 private:
  uint32_t query_index(const Column& c, int32_t at) {  // :1571-1627
    auto& qs = c.column_type == Any::Advice ? advice_queries : c.column_type == Any::Fixed ? fixed_queries : instance_queries;
    for (size_t i = 0; i < qs.size(); ++i)
      if (qs[i].first == c && qs[i].second == at) return uint32_t(i);
    qs.push_back({c, at});
    if (c.column_type == Any::Advice) num_advice_queries[c.index] += 1;
    return uint32_t(qs.size() - 1);
  }
};

/// format!("{:?}", vk.pinned()): moduli and points are given as their `{:?}` strings (0x + 64 hex digits; a point
/// is "(x, y)"), so the same formatter serves bn256 and the reference's Vesta fixture           plonk.rs:220-230
inline std::string pinned_debug(const ConstraintSystem& cs, uint32_t k, uint32_t extended_k, const std::string& omega_hex,
                                const std::vector<std::string>& fixed_commitments,
                                const std::vector<std::string>& permutation_commitments,
                                const std::string& base_modulus_hex, const std::string& scalar_modulus_hex) {
  auto list = [](const std::vector<std::string>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + v[i];
    return s + "]";
  };
  return "PinnedVerificationKey { base_modulus: \"" + base_modulus_hex + "\", scalar_modulus: \"" + scalar_modulus_hex +
         "\", domain: PinnedEvaluationDomain { k: " + std::to_string(k) + ", extended_k: " + std::to_string(extended_k) +
         ", omega: " + omega_hex + " }, cs: PinnedConstraintSystem { " + cs.pinned_debug() + " }, fixed_commitments: " +
         list(fixed_commitments) + ", permutation: VerifyingKey { commitments: " + list(permutation_commitments) + " } }";
}
inline std::string debug_point(const G1Affine& p) {
  const Fq x = p.x.to_repr_limbs(), y = p.y.to_repr_limbs();
  char buf[140];
  std::snprintf(buf, sizeof buf, "(0x%016llx%016llx%016llx%016llx, 0x%016llx%016llx%016llx%016llx)", (unsigned long long)x.l[3],
                (unsigned long long)x.l[2], (unsigned long long)x.l[1], (unsigned long long)x.l[0], (unsigned long long)y.l[3],
                (unsigned long long)y.l[2], (unsigned long long)y.l[1], (unsigned long long)y.l[0]);
  return p.is_identity() ? "Infinity" : buf;
}

/// VerifyingKey::from_parts: transcript_repr = from_bytes_wide(Blake2b-512("Halo2-Verify-Key"; len as u64 LE ‖ s))
///                                                                                          plonk.rs:192-203
inline Fr vk_transcript_repr(const std::string& pinned) {
  transcript::Blake2b h("Halo2-Verify-Key");
  const uint64_t len = pinned.size();
  h.update(&len, 8);
  h.update(pinned.data(), pinned.size());
  uint8_t d[64];
  h.digest(d);
  return transcript::fr_from_bytes_wide(d);
}

}  // namespace plonk
}  // namespace halo2_proofs

#endif  // HALO2_B200_PLONK_HPP
