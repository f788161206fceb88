// halo2_b200_plonk.hpp -- C++ host side, second part: the constraint-system description a prover works from
// and the verifying-key string that seeds every transcript.  Same names as the reference:
//
//   halo2_proofs::plonk::{Any, Column, Expression, ConstraintSystem}      halo2_proofs/src/plonk/circuit.rs:780-2060
//   halo2_proofs::plonk::lookup::Argument, permutation::Argument         plonk/lookup.rs:10-60, plonk/permutation.rs:19-75
//   halo2_proofs::plonk::{ValueSource, Calculation, GraphEvaluator, Evaluator::new}   plonk/evaluation.rs:38-277, 525-690
//   halo2_proofs::plonk::pinned_debug  = format!("{:?}", vk.pinned())     plonk.rs:192-230, circuit.rs:1399-1448
//   halo2_proofs::plonk::vk_transcript_repr                               plonk.rs:192-203
//   halo2_proofs::plonk::{keygen_pk, ProvingKey, permutation::Assembly}   plonk/keygen.rs:203-367, plonk/permutation/keygen.rs
//   halo2_proofs::plonk::create_proof (gates + permutation + instances)    plonk/prover.rs:37-651
//
// Bookkeeping only (no selectors, regions or floor planner: the prover reads the constraint system of the
// verifying key, where selectors are already fixed columns, plonk/prover.rs:69-71).  tests/test_cpp_mirror.py
// rebuilds `configure()` of the reference's tests/plonk_api.rs with it and compares the string, character for
// character, with the reference's golden verifying key.
#ifndef HALO2_B200_PLONK_HPP
#define HALO2_B200_PLONK_HPP

#include <algorithm>
#include <cstdio>
#include <functional>

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
  const Fr& value() const { return n_->value; }            // Constant / Scaled
  uint32_t column_index() const { return n_->b; }          // queries
  uint32_t challenge_index() const { return n_->a; }       // Challenge
  int32_t rotation() const { return n_->rot; }
  const Expression& lhs() const { return *n_->l; }
  const Expression& rhs() const { return *n_->r; }

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

// ---------------------------------------------------------------------------------------------
// GraphEvaluator / Evaluator::new: expressions -> the calculation list the library compiles (h2b_graph_new)
//                                                                 halo2_proofs/src/plonk/evaluation.rs:38-127, 183-277, 525-690
// ---------------------------------------------------------------------------------------------
struct ValueSource {  // variant order == the reference's derived PartialOrd (evaluation.rs:38-61) == the ABI's operand kinds
  enum Kind : uint32_t { Constant, Intermediate, Fixed, Advice, Instance, Challenge, Beta, Gamma, Theta, Y, PreviousValue };
  uint32_t kind = Constant, a = 0, b = 0;
  bool operator==(const ValueSource& o) const { return kind == o.kind && a == o.a && b == o.b; }
  bool operator<=(const ValueSource& o) const {
    return kind != o.kind ? kind < o.kind : a != o.a ? a < o.a : b <= o.b;
  }
};

struct Calculation {  // evaluation.rs:110-127; op numbering of include/halo2_b200.h
  enum Op : uint32_t { Add, Sub, Mul, Square, Double, Negate, Horner, Store };
  uint32_t op = Store;
  std::vector<ValueSource> src;  // Horner: start, factor, then the parts
  bool operator==(const Calculation& o) const { return op == o.op && src == o.src; }
};

class GraphEvaluator {
 public:
  std::vector<Fr> constants{Fr::zero(), Fr::one(), Fr::from(2)};  // fixed positions (:525-538)
  std::vector<int32_t> rotations;
  std::vector<std::pair<Calculation, uint32_t>> calculations;     // (calculation, target)
  uint32_t num_intermediates = 0;

  uint32_t add_rotation(int32_t r) {
    for (size_t i = 0; i < rotations.size(); ++i)
      if (rotations[i] == r) return uint32_t(i);
    rotations.push_back(r);
    return uint32_t(rotations.size() - 1);
  }
  ValueSource add_constant(const Fr& c) {
    for (size_t i = 0; i < constants.size(); ++i)
      if (constants[i] == c) return {ValueSource::Constant, uint32_t(i), 0};
    constants.push_back(c);
    return {ValueSource::Constant, uint32_t(constants.size() - 1), 0};
  }
  ValueSource add_calculation(const Calculation& c) {
    for (const auto& e : calculations)
      if (e.first == c) return {ValueSource::Intermediate, e.second, 0};
    calculations.push_back({c, num_intermediates});
    return {ValueSource::Intermediate, num_intermediates++, 0};
  }
  ValueSource add_expression(const Expression& e) {  // :590-690
    const ValueSource zero{ValueSource::Constant, 0, 0}, one{ValueSource::Constant, 1, 0}, two{ValueSource::Constant, 2, 0};
    auto calc = [&](uint32_t op, std::initializer_list<ValueSource> s) { return add_calculation(Calculation{op, s}); };
    switch (e.kind()) {
      case Expression::Constant: return add_constant(e.value());
      case Expression::Fixed: return calc(Calculation::Store, {{ValueSource::Fixed, e.column_index(), add_rotation(e.rotation())}});
      case Expression::Advice: return calc(Calculation::Store, {{ValueSource::Advice, e.column_index(), add_rotation(e.rotation())}});
      case Expression::Instance: return calc(Calculation::Store, {{ValueSource::Instance, e.column_index(), add_rotation(e.rotation())}});
      case Expression::Challenge: return calc(Calculation::Store, {{ValueSource::Challenge, e.challenge_index(), 0}});
      case Expression::Negated: {
        if (e.lhs().kind() == Expression::Constant) return add_constant(-e.lhs().value());
        const ValueSource a = add_expression(e.lhs());
        return a == zero ? a : calc(Calculation::Negate, {a});
      }
      case Expression::Sum: {
        if (e.rhs().kind() == Expression::Negated) {  // undo subtraction stored as a + (-b)
          const ValueSource a = add_expression(e.lhs()), b = add_expression(e.rhs().lhs());
          if (a == zero) return calc(Calculation::Negate, {b});
          if (b == zero) return a;
          return calc(Calculation::Sub, {a, b});
        }
        const ValueSource a = add_expression(e.lhs()), b = add_expression(e.rhs());
        if (a == zero) return b;
        if (b == zero) return a;
        return a <= b ? calc(Calculation::Add, {a, b}) : calc(Calculation::Add, {b, a});
      }
      case Expression::Product: {
        const ValueSource a = add_expression(e.lhs()), b = add_expression(e.rhs());
        if (a == zero || b == zero) return zero;
        if (a == one) return b;
        if (b == one) return a;
        if (a == two) return calc(Calculation::Double, {b});
        if (b == two) return calc(Calculation::Double, {a});
        if (a == b) return calc(Calculation::Square, {a});
        return a <= b ? calc(Calculation::Mul, {a, b}) : calc(Calculation::Mul, {b, a});
      }
      default: {  // Scaled
        if (e.value().is_zero()) return zero;
        if (e.value() == Fr::one()) return add_expression(e.lhs());
        const ValueSource c = add_constant(e.value());
        const ValueSource a = add_expression(e.lhs());
        return calc(Calculation::Mul, {a, c});
      }
    }
  }
  /// the word stream of include/halo2_b200.h: op, target, operands (three words each); Horner: start, factor, nparts, parts
  std::vector<uint32_t> encode() const {
    std::vector<uint32_t> w;
    auto put = [&](const ValueSource& s) { w.push_back(s.kind), w.push_back(s.a), w.push_back(s.b); };
    for (const auto& e : calculations) {
      w.push_back(e.first.op), w.push_back(e.second);
      if (e.first.op == Calculation::Horner) {
        put(e.first.src[0]), put(e.first.src[1]);
        w.push_back(uint32_t(e.first.src.size() - 2));
        for (size_t i = 2; i < e.first.src.size(); ++i) put(e.first.src[i]);
      } else {
        for (const auto& s : e.first.src) put(s);
      }
    }
    return w;
  }
  /// h2b_graph_new on the calling thread's context (the library compiles the list once, as Evaluator::new does at keygen)
  std::shared_ptr<h2b_graph> compile() const {
    h2b_ctx* ctx = detail::backend().ctx;
    const auto code = encode();
    h2b_graph* g = nullptr;
    detail::check(ctx, h2b_graph_new(ctx, code.data(), code.size(), constants.data(), uint32_t(constants.size()), rotations.data(),
                                     uint32_t(rotations.size()), num_intermediates, &g), "h2b_graph_new: malformed calculation list");
    return std::shared_ptr<h2b_graph>(g, h2b_graph_free);
  }
};

/// Evaluator::new(cs): custom_gates (all gate polynomials folded with y onto the previous value) and one graph
/// per lookup ((compressed input + beta) * (compressed table + gamma))                    evaluation.rs:224-277
struct Evaluator {
  GraphEvaluator custom_gates;
  std::vector<GraphEvaluator> lookups;
  explicit Evaluator(const ConstraintSystem& cs) {
    Calculation horner{Calculation::Horner, {{ValueSource::PreviousValue, 0, 0}, {ValueSource::Y, 0, 0}}};
    for (const auto& gate : cs.gates)
      for (const auto& poly : gate.second) horner.src.push_back(custom_gates.add_expression(poly));
    custom_gates.add_calculation(horner);
    for (const auto& lk : cs.lookups) {
      GraphEvaluator graph;
      auto evaluate_lc = [&](const std::vector<Expression>& exprs) {
        Calculation h{Calculation::Horner, {{ValueSource::Constant, 0, 0}, {ValueSource::Theta, 0, 0}}};
        for (const auto& e : exprs) h.src.push_back(graph.add_expression(e));
        return graph.add_calculation(h);
      };
      const ValueSource in = evaluate_lc(lk.input_expressions), tab = evaluate_lc(lk.table_expressions);
      const ValueSource right_gamma = graph.add_calculation({Calculation::Add, {tab, {ValueSource::Gamma, 0, 0}}});
      const ValueSource lc = graph.add_calculation({Calculation::Add, {in, {ValueSource::Beta, 0, 0}}});
      graph.add_calculation({Calculation::Mul, {lc, right_gamma}});
      lookups.push_back(std::move(graph));
    }
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

// ---------------------------------------------------------------------------------------------
// keygen_vk + keygen_pk for a circuit handed over as its constraint system, its assigned fixed columns (what
// Assembly::assign_fixed collects) and its copy constraints (Assembly::copy)
//                       halo2_proofs/src/plonk/keygen.rs:203-367, plonk/permutation/keygen.rs:16-241, plonk.rs:171-305
// Polynomials are host vectors, as in the reference's ProvingKey; every transform and commitment runs on the GPU
// through the entry points of halo2_b200.hpp.
// ---------------------------------------------------------------------------------------------
struct CopyConstraint {
  Column left_column;
  size_t left_row;
  Column right_column;
  size_t right_row;
};

namespace permutation {
/// Assembly: cycles of equal cells, merged smaller-into-larger                   plonk/permutation/keygen.rs:16-107
class Assembly {
 public:
  Assembly(size_t n, const std::vector<Column>& columns)
      : mapping(n * columns.size()), n_(n), columns_(columns), aux_(mapping.size()), sizes_(mapping.size(), 1) {
    for (size_t i = 0; i < mapping.size(); ++i) mapping[i] = aux_[i] = i;
  }
  void copy(const Column& lc, size_t lrow, const Column& rc, size_t rrow) {
    const auto li = std::find(columns_.begin(), columns_.end(), lc), ri = std::find(columns_.begin(), columns_.end(), rc);
    if (li == columns_.end() || ri == columns_.end()) throw Panic("Error::ColumnNotInPermutation (permutation/keygen.rs:55-62)");
    if (lrow >= n_ || rrow >= n_) throw Panic("Error::BoundsFailure (permutation/keygen.rs:65-67)");
    const size_t left = size_t(li - columns_.begin()) * n_ + lrow, right = size_t(ri - columns_.begin()) * n_ + rrow;
    size_t left_cycle = aux_[left], right_cycle = aux_[right];
    if (left_cycle == right_cycle) return;
    if (sizes_[left_cycle] < sizes_[right_cycle]) std::swap(left_cycle, right_cycle);
    sizes_[left_cycle] += sizes_[right_cycle];
    for (size_t i = right_cycle;;) {
      aux_[i] = left_cycle;
      i = mapping[i];
      if (i == right_cycle) break;
    }
    std::swap(mapping[left], mapping[right]);
  }
  std::vector<size_t> mapping;  // cell (column i, row j) as i * n + j  ->  the next cell of its cycle

 private:
  size_t n_;
  std::vector<Column> columns_;
  std::vector<size_t> aux_, sizes_;
};
}  // namespace permutation

struct ProvingKey {  // plonk.rs:258-305 with the VerifyingKey inside it (:45-58)
  ConstraintSystem cs;
  std::shared_ptr<poly::EvaluationDomain> domain;
  uint32_t k = 0;
  size_t n = 0;
  std::vector<G1Affine> fixed_commitments, permutation_commitments;
  std::vector<poly::Polynomial<poly::LagrangeCoeff>> fixed_values, permutations;
  std::vector<poly::Polynomial<poly::Coeff>> fixed_polys, permutation_polys;
  std::vector<poly::Polynomial<poly::ExtendedLagrangeCoeff>> fixed_cosets, permutation_cosets;
  poly::Polynomial<poly::ExtendedLagrangeCoeff> l0, l_last, l_active_row;
  std::shared_ptr<Evaluator> ev;
  std::vector<std::pair<GraphEvaluator, GraphEvaluator>> lookup_compress;  // theta-compression of (input, table) expressions
  std::string pinned;   // format!("{:?}", vk.pinned())
  Fr transcript_repr;   // its hash, the first thing every transcript absorbs (plonk.rs:192-203)
};

inline ProvingKey keygen_pk(const poly::kzg::ParamsKZG& params, const ConstraintSystem& cs,
                            const std::vector<std::vector<Fr>>& fixed_values, const std::vector<CopyConstraint>& copies) {
  using namespace poly;
  const size_t n = params.n();
  if (n < cs.minimum_rows()) throw Panic("Error::not_enough_rows_available (keygen.rs:219-221)");
  if (fixed_values.size() != cs.num_fixed_columns) throw Panic("one assignment per fixed column expected");
  ProvingKey pk;
  pk.cs = cs, pk.k = params.k(), pk.n = n;
  pk.domain = std::make_shared<EvaluationDomain>(uint32_t(cs.degree()), params.k());
  const EvaluationDomain& dom = *pk.domain;
  // fixed columns (keygen.rs:237-258, 300-316)
  for (const auto& col : fixed_values) {
    if (col.size() > n) throw Panic("Error::not_enough_rows_available");
    std::vector<Fr> v = col;
    v.resize(n, Fr::zero());
    pk.fixed_values.push_back(dom.lagrange_from_vec(std::move(v)));
    pk.fixed_commitments.push_back(params.commit_lagrange(pk.fixed_values.back()).to_affine());
    pk.fixed_polys.push_back(dom.lagrange_to_coeff(pk.fixed_values.back()));
    pk.fixed_cosets.push_back(dom.coeff_to_extended(pk.fixed_polys.back()));
  }
  // permutation (permutation/keygen.rs:109-241): sigma_i[j] = delta^i' omega^j' for mapping[i][j] = (i', j')
  const auto& pc = cs.permutation.columns;
  permutation::Assembly assembly(n, pc);
  for (const auto& c : copies) assembly.copy(c.left_column, c.left_row, c.right_column, c.right_row);
  const Fr delta = Fr::from_raw(0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull, 0x64ec26aad4c86e71ull, 0x09226b6e22c6f0caull);  // Fr::DELTA = 7^(2^28)
  std::vector<Fr> omega_powers(n), deltaomega(pc.size() * n);
  {
    Fr cur = Fr::one();
    for (size_t j = 0; j < n; ++j) omega_powers[j] = cur, cur *= dom.get_omega();
    Fr d = Fr::one();
    for (size_t i = 0; i < pc.size(); ++i, d *= delta)
      for (size_t j = 0; j < n; ++j) deltaomega[i * n + j] = omega_powers[j] * d;
  }
  for (size_t i = 0; i < pc.size(); ++i) {
    std::vector<Fr> sigma(n);
    for (size_t j = 0; j < n; ++j) sigma[j] = deltaomega[assembly.mapping[i * n + j]];
    pk.permutations.push_back(dom.lagrange_from_vec(std::move(sigma)));
    pk.permutation_commitments.push_back(params.commit_lagrange(pk.permutations.back()).to_affine());
    pk.permutation_polys.push_back(dom.lagrange_to_coeff(pk.permutations.back()));
    pk.permutation_cosets.push_back(dom.coeff_to_extended(pk.permutation_polys.back()));
  }
  // l_0, l_blind, l_last, l_active_row (keygen.rs:322-350)
  const size_t bf = cs.blinding_factors();
  auto indicator = [&](size_t first, size_t last) {
    auto v = dom.empty_lagrange();
    for (size_t r = first; r < last; ++r) v[r] = Fr::one();
    return dom.coeff_to_extended(dom.lagrange_to_coeff(v));
  };
  pk.l0 = indicator(0, 1);
  const auto l_blind = indicator(n - bf, n);
  pk.l_last = indicator(n - bf - 1, n - bf);
  pk.l_active_row = dom.constant_extended(Fr::one()) - pk.l_last - l_blind;
  pk.ev = std::make_shared<Evaluator>(cs);  // keygen.rs:353
  for (const auto& lk : cs.lookups) {       // compression over the Lagrange rows (lookup/prover.rs:82-104), as graphs
    auto compress = [](const std::vector<Expression>& exprs) {
      GraphEvaluator g;
      Calculation h{Calculation::Horner, {{ValueSource::Constant, 0, 0}, {ValueSource::Theta, 0, 0}}};
      for (const auto& e : exprs) h.src.push_back(g.add_expression(e));
      g.add_calculation(h);
      return g;
    };
    pk.lookup_compress.push_back({compress(lk.input_expressions), compress(lk.table_expressions)});
  }
  std::vector<std::string> fc, pcm;
  for (const auto& p : pk.fixed_commitments) fc.push_back(debug_point(p));
  for (const auto& p : pk.permutation_commitments) pcm.push_back(debug_point(p));
  pk.pinned = pinned_debug(cs, pk.k, dom.extended_k(), hex(dom.get_omega()), fc, pcm,
                           "0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47",
                           "0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001");
  pk.transcript_repr = vk_transcript_repr(pk.pinned);
  return pk;
}

// ---------------------------------------------------------------------------------------------
// create_proof<KZGCommitmentScheme<Bn256>, ProverGWC | ProverSHPLONK, Challenge255, R, Blake2bWrite>
//                                                                          halo2_proofs/src/plonk/prover.rs:37-651
// for constraint systems with gates, lookups, a permutation argument, instance columns and challenge phases.  The witness is handed over as assigned advice
// columns (the role of Circuit::synthesize through WitnessCollection, :143-285).  Host vectors between the steps,
// as in the reference; every transform, commitment, grand product and the whole quotient evaluation run on the GPU.
// ---------------------------------------------------------------------------------------------
/// rand_xorshift::XorShiftRng::from_seed([u8; 16]); next_u64 = two next_u32, low word first
class XorShiftRng {
 public:
  explicit XorShiftRng(const uint8_t seed[16]) {
    std::memcpy(s_, seed, 16);
    if (!(s_[0] | s_[1] | s_[2] | s_[3])) s_[0] = s_[1] = s_[2] = s_[3] = 0x0BAD5EED;
  }
  uint32_t next_u32() {
    const uint32_t t = s_[0] ^ (s_[0] << 11);
    s_[0] = s_[1], s_[1] = s_[2], s_[2] = s_[3];
    s_[3] = s_[3] ^ (s_[3] >> 19) ^ (t ^ (t >> 8));
    return s_[3];
  }
  uint64_t next_u64() {
    const uint64_t lo = next_u32();
    return (uint64_t(next_u32()) << 32) | lo;
  }

 private:
  uint32_t s_[4];
};

/// Fr::random(rng): the 512-bit little-endian integer of eight next_u64 draws, mod r (halo2curves 0.3.1)
template <class Rng>
inline Fr fr_random(Rng& rng) {
  uint8_t b[64];
  for (int i = 0; i < 8; ++i) {
    const uint64_t v = rng.next_u64();
    std::memcpy(b + 8 * i, &v, 8);
  }
  return transcript::fr_from_bytes_wide(b);
}

namespace detail {
class DeviceVec {  // a device-resident array of Fr on the calling thread's context
 public:
  explicit DeviceVec(size_t n) : ctx_(halo2_proofs::detail::backend().ctx), n_(n) {
    halo2_proofs::detail::check(ctx_, h2b_device_alloc(ctx_, (n ? n : 1) * sizeof(Fr), &p_), "h2b_device_alloc");
  }
  explicit DeviceVec(const std::vector<Fr>& v) : DeviceVec(v.size()) {
    if (!v.empty()) halo2_proofs::detail::check(ctx_, h2b_copy_h2d(ctx_, p_, v.data(), v.size() * sizeof(Fr)), "h2b_copy_h2d");
  }
  DeviceVec(const DeviceVec&) = delete;
  DeviceVec& operator=(const DeviceVec&) = delete;
  ~DeviceVec() { h2b_device_free(ctx_, p_); }
  h2b_fr* ptr() const { return static_cast<h2b_fr*>(p_); }
  void zero() { halo2_proofs::detail::check(ctx_, h2b_device_memset(ctx_, p_, 0, n_ * sizeof(Fr)), "h2b_device_memset"); }
  std::vector<Fr> download() const {
    std::vector<Fr> v(n_);
    if (n_) halo2_proofs::detail::check(ctx_, h2b_copy_d2h(ctx_, v.data(), p_, n_ * sizeof(Fr)), "h2b_copy_d2h");
    return v;
  }

 private:
  h2b_ctx* ctx_;
  void* p_ = nullptr;
  size_t n_;
};
}  // namespace detail

enum class Multiopen { GWC, SHPLONK };

/// witness(circuit, phase, challenges) -> the assigned advice columns of that circuit (one vector per advice
/// column; only the columns of `phase` are read; shorter than n: zero-padded): the role of Circuit::synthesize through
/// WitnessCollection, called once per phase with the challenges squeezed so far (prover.rs:143-285, 287-405).
using Witness = std::function<std::vector<std::vector<Fr>>(size_t circuit, uint32_t phase, const std::vector<Fr>& challenges)>;

/// instances[i][c]: the values of instance column c of circuit i
template <class Rng>
inline void create_proof(const poly::kzg::ParamsKZG& params, const ProvingKey& pk, const Witness& witness,
                         const std::vector<std::vector<std::vector<Fr>>>& instances, Rng& rng,
                         transcript::Blake2bWrite& transcript, Multiopen scheme = Multiopen::GWC) {
  const size_t n_circuits = instances.size();  // one witness per instance list
  using namespace poly;
  using detail::DeviceVec;
  const ConstraintSystem& cs = pk.cs;
  const EvaluationDomain& dom = *pk.domain;
  h2b_ctx* ctx = halo2_proofs::detail::backend().ctx;
  const size_t n = pk.n, bf = cs.blinding_factors(), ext = dom.extended_len();
  for (const auto& inst : instances)
    if (inst.size() != cs.num_instance_columns) throw Panic("Error::InvalidInstances (prover.rs:55-59)");
  auto rot = [&](const Fr& v, int32_t r) { return dom.rotate_omega(v, Rotation{r}); };
  auto eval = [](const Polynomial<Coeff>& p, const Fr& x) { return arithmetic::eval_polynomial(p.values, x); };

  transcript.common_scalar(pk.transcript_repr);  // :62

  // ---- instances (:79-138; QUERY_INSTANCE = false) ----
  std::vector<std::vector<Polynomial<LagrangeCoeff>>> instance_values(instances.size());
  std::vector<std::vector<Polynomial<Coeff>>> instance_polys(instances.size());
  for (size_t ci = 0; ci < instances.size(); ++ci)
    for (const auto& values : instances[ci]) {
      if (values.size() > n - (bf + 1)) throw Panic("Error::InstanceTooLarge (prover.rs:93-95)");
      for (const Fr& v : values) transcript.common_scalar(v);
      std::vector<Fr> padded = values;
      padded.resize(n, Fr::zero());
      instance_values[ci].push_back(dom.lagrange_from_vec(padded));
      instance_polys[ci].push_back(dom.lagrange_to_coeff(instance_values[ci].back()));
    }

  // ---- advice (:287-405) ----
  const size_t unusable_rows_start = n - (bf + 1);
  std::vector<std::vector<Polynomial<LagrangeCoeff>>> advice_values(n_circuits, std::vector<Polynomial<LagrangeCoeff>>(cs.num_advice_columns));
  std::vector<Fr> challenges(cs.num_challenges, Fr::zero());
  uint32_t last_phase = 0;
  for (uint32_t p : cs.advice_column_phase) last_phase = std::max(last_phase, p);
  for (uint32_t phase = 0; phase <= last_phase; ++phase) {
    std::vector<size_t> column_indices;
    for (size_t c = 0; c < cs.num_advice_columns; ++c)
      if (cs.advice_column_phase[c] == phase) column_indices.push_back(c);
    for (size_t ci = 0; ci < n_circuits; ++ci) {
      const auto assigned = witness(ci, phase, challenges);
      if (assigned.size() != cs.num_advice_columns) throw Panic("one assignment per advice column expected");
      for (size_t c : column_indices) {  // assigned rows, zero padding, then the blinding factors (:364-368)
        if (assigned[c].size() > unusable_rows_start) throw Panic("Error::not_enough_rows_available (prover.rs:228-230)");
        std::vector<Fr> v = assigned[c];
        v.resize(unusable_rows_start, Fr::zero());
        for (size_t r = unusable_rows_start; r < n; ++r) v.push_back(fr_random(rng));
        advice_values[ci][c] = dom.lagrange_from_vec(std::move(v));
      }
      for (size_t j = 0; j < column_indices.size(); ++j) fr_random(rng);  // Blind(Scalar::random(rng)), ignored by KZG (:371-374)
      for (size_t c : column_indices) transcript.write_point(params.commit_lagrange(advice_values[ci][c]).to_affine());  // :375-392
    }
    for (size_t i = 0; i < cs.challenge_phase.size(); ++i)  // :394-403
      if (cs.challenge_phase[i] == phase) challenges[i] = transcript.squeeze_challenge_scalar();
  }
  const Fr theta = transcript.squeeze_challenge_scalar();  // :410

  // ---- lookups: permuted columns (:412-437, lookup/prover.rs:55-140) ----
  struct Lookup {
    std::vector<Fr> compressed_input, compressed_table, permuted_input, permuted_table;  // Lagrange
    Polynomial<Coeff> permuted_input_poly, permuted_table_poly, product_poly;
  };
  std::vector<std::vector<Lookup>> lookups(n_circuits);
  for (size_t ci = 0; ci < n_circuits && !cs.lookups.empty(); ++ci) {
    std::vector<std::unique_ptr<DeviceVec>> keep;
    std::vector<const h2b_fr*> fixed_ptrs, adv_ptrs, inst_ptrs;
    for (const auto& p : pk.fixed_values) keep.emplace_back(new DeviceVec(p.values)), fixed_ptrs.push_back(keep.back()->ptr());
    for (const auto& p : advice_values[ci]) keep.emplace_back(new DeviceVec(p.values)), adv_ptrs.push_back(keep.back()->ptr());
    for (const auto& p : instance_values[ci]) keep.emplace_back(new DeviceVec(p.values)), inst_ptrs.push_back(keep.back()->ptr());
    h2b_eval_columns cols;
    std::memset(&cols, 0, sizeof cols);
    cols.fixed = fixed_ptrs.data(), cols.n_fixed = uint32_t(fixed_ptrs.size());
    cols.advice = adv_ptrs.data(), cols.n_advice = uint32_t(adv_ptrs.size());
    cols.instance = inst_ptrs.data(), cols.n_instance = uint32_t(inst_ptrs.size());
    cols.challenges = challenges.data(), cols.n_challenges = uint32_t(challenges.size());
    cols.theta = theta;
    for (const auto& graphs : pk.lookup_compress) {
      Lookup lk;
      DeviceVec cin(n), ctab(n), pin(n), ptab(n);
      cin.zero(), ctab.zero();
      halo2_proofs::detail::check(ctx, h2b_graph_evaluate_lagrange(dom.raw(), graphs.first.compile().get(), &cols, cin.ptr()), "h2b_graph_evaluate_lagrange");
      halo2_proofs::detail::check(ctx, h2b_graph_evaluate_lagrange(dom.raw(), graphs.second.compile().get(), &cols, ctab.ptr()), "h2b_graph_evaluate_lagrange");
      const int rc = h2b_lookup_permute(ctx, cin.ptr(), ctab.ptr(), unusable_rows_start, pin.ptr(), ptab.ptr());
      if (rc == H2B_ERR_CONSTRAINT) throw Panic("Error::ConstraintSystemFailure: a lookup input is not in the table (lookup/prover.rs:425-433)");
      halo2_proofs::detail::check(ctx, rc, "h2b_lookup_permute");
      lk.compressed_input = cin.download(), lk.compressed_table = ctab.download();
      lk.permuted_input = pin.download(), lk.permuted_table = ptab.download();
      for (auto* v : {&lk.permuted_input, &lk.permuted_table})  // blinding rows (:447-449)
        for (size_t r = unusable_rows_start; r < n; ++r) (*v)[r] = fr_random(rng);
      lk.permuted_input_poly = dom.lagrange_to_coeff(dom.lagrange_from_vec(lk.permuted_input));  // commit_values (:114-125)
      fr_random(rng);                                                                          // Blind
      lk.permuted_table_poly = dom.lagrange_to_coeff(dom.lagrange_from_vec(lk.permuted_table));
      fr_random(rng);
      transcript.write_point(params.commit_lagrange(dom.lagrange_from_vec(lk.permuted_input)).to_affine());
      transcript.write_point(params.commit_lagrange(dom.lagrange_from_vec(lk.permuted_table)).to_affine());
      lookups[ci].push_back(std::move(lk));
    }
  }

  const Fr beta = transcript.squeeze_challenge_scalar();   // :440
  const Fr gamma = transcript.squeeze_challenge_scalar();  // :443

  // ---- permutation argument (:446-463, permutation/prover.rs:40-190) ----
  struct Set {
    Polynomial<Coeff> poly;
    Polynomial<ExtendedLagrangeCoeff> coset;
  };
  std::vector<std::vector<Set>> permutations(n_circuits);
  const auto& pcols = cs.permutation.columns;
  const size_t chunk_len = cs.degree() - 2;
  for (size_t ci = 0; ci < n_circuits; ++ci) {
    Fr last_z = Fr::one();
    for (size_t s0 = 0; s0 < pcols.size(); s0 += chunk_len) {
      const size_t m = std::min(chunk_len, pcols.size() - s0);
      std::vector<std::unique_ptr<DeviceVec>> keep;
      std::vector<const h2b_fr*> vals, sigmas;
      for (size_t j = 0; j < m; ++j) {
        const Column& c = pcols[s0 + j];
        const std::vector<Fr>& column = c.column_type == Any::Advice ? advice_values[ci][c.index].values
                                        : c.column_type == Any::Fixed ? pk.fixed_values[c.index].values
                                                                      : instance_values[ci][c.index].values;
        keep.emplace_back(new DeviceVec(column)), vals.push_back(keep.back()->ptr());
        keep.emplace_back(new DeviceVec(pk.permutations[s0 + j].values)), sigmas.push_back(keep.back()->ptr());
      }
      DeviceVec frac(n), zdev(n);
      halo2_proofs::detail::check(ctx, h2b_permutation_fractions(dom.raw(), vals.data(), sigmas.data(), uint32_t(m), uint32_t(s0), &beta, &gamma, frac.ptr()),
                                  "h2b_permutation_fractions");
      halo2_proofs::detail::check(ctx, h2b_running_product(ctx, frac.ptr(), H2B_DEVICE, n, &last_z, zdev.ptr()), "h2b_running_product");  // :150-158
      std::vector<Fr> z = zdev.download();
      for (size_t r = n - bf; r < n; ++r) z[r] = fr_random(rng);  // :161-163
      last_z = z[n - (bf + 1)];                                   // :165
      fr_random(rng);                                             // Blind (:167)
      const auto zl = dom.lagrange_from_vec(std::move(z));
      transcript.write_point(params.commit_lagrange(zl).to_affine());
      Set st;
      st.poly = dom.lagrange_to_coeff(zl);
      st.coset = dom.coeff_to_extended(st.poly);
      permutations[ci].push_back(std::move(st));
    }
  }

  // ---- lookups: grand products (:466-475, lookup/prover.rs:146-250) ----
  for (auto& lks : lookups)
    for (auto& lk : lks) {
      const DeviceVec pin(lk.permuted_input), ptab(lk.permuted_table), cin(lk.compressed_input), ctab(lk.compressed_table);
      DeviceVec frac(n), zdev(n);
      halo2_proofs::detail::check(ctx, h2b_lookup_product_fractions(ctx, pin.ptr(), ptab.ptr(), cin.ptr(), ctab.ptr(), &beta, &gamma, n, frac.ptr()),
                                  "h2b_lookup_product_fractions");
      const Fr one = Fr::one();
      halo2_proofs::detail::check(ctx, h2b_running_product(ctx, frac.ptr(), H2B_DEVICE, n, &one, zdev.ptr()), "h2b_running_product");  // :201-209
      std::vector<Fr> z = zdev.download();
      for (size_t r = n - bf; r < n; ++r) z[r] = fr_random(rng);
      fr_random(rng);  // product_blind
      const auto zl = dom.lagrange_from_vec(std::move(z));
      transcript.write_point(params.commit_lagrange(zl).to_affine());
      lk.product_poly = dom.lagrange_to_coeff(zl);
    }

  // ---- vanishing argument: random polynomial (vanishing/prover.rs:36-66) ----
  Polynomial<Coeff> random_poly = dom.empty_coeff();
  for (auto& c : random_poly) c = fr_random(rng);
  fr_random(rng);  // random_blind
  transcript.write_point(params.commit(random_poly).to_affine());

  const Fr y = transcript.squeeze_challenge_scalar();  // :478

  std::vector<std::vector<Polynomial<Coeff>>> advice_polys(n_circuits);  // :481-499
  for (size_t ci = 0; ci < n_circuits; ++ci)
    for (const auto& v : advice_values[ci]) advice_polys[ci].push_back(dom.lagrange_to_coeff(v));

  // ---- h(X): Evaluator::evaluate_h on the device (:502-520, evaluation.rs:280-522) ----
  std::vector<Fr> h_ext;
  {
    DeviceVec values(ext);
    values.zero();
    auto upload_all = [](const std::vector<Polynomial<ExtendedLagrangeCoeff>>& polys, std::vector<std::unique_ptr<DeviceVec>>& keep,
                         std::vector<const h2b_fr*>& ptrs) {
      for (const auto& p : polys) keep.emplace_back(new DeviceVec(p.values)), ptrs.push_back(keep.back()->ptr());
    };
    std::vector<std::unique_ptr<DeviceVec>> fixed_keep, sigma_keep;
    std::vector<const h2b_fr*> fixed_ptrs, sigma_ptrs;
    upload_all(pk.fixed_cosets, fixed_keep, fixed_ptrs);
    upload_all(pk.permutation_cosets, sigma_keep, sigma_ptrs);
    const DeviceVec l0(pk.l0.values), l_last(pk.l_last.values), l_active_row(pk.l_active_row.values);
    const auto graph = pk.ev->custom_gates.compile();
    std::vector<uint32_t> ctype, cidx;
    for (const auto& c : pcols) ctype.push_back(uint32_t(c.column_type)), cidx.push_back(c.index);
    for (size_t ci = 0; ci < n_circuits; ++ci) {
      std::vector<Polynomial<ExtendedLagrangeCoeff>> adv, inst, zs;
      for (const auto& p : advice_polys[ci]) adv.push_back(dom.coeff_to_extended(p));    // :305-323
      for (const auto& p : instance_polys[ci]) inst.push_back(dom.coeff_to_extended(p));
      for (const auto& st : permutations[ci]) zs.push_back(st.coset);
      std::vector<std::unique_ptr<DeviceVec>> keep;
      std::vector<const h2b_fr*> adv_ptrs, inst_ptrs, z_ptrs;
      upload_all(adv, keep, adv_ptrs), upload_all(inst, keep, inst_ptrs), upload_all(zs, keep, z_ptrs);
      h2b_eval_columns cols;
      std::memset(&cols, 0, sizeof cols);
      cols.fixed = fixed_ptrs.data(), cols.n_fixed = uint32_t(fixed_ptrs.size());
      cols.advice = adv_ptrs.data(), cols.n_advice = uint32_t(adv_ptrs.size());
      cols.instance = inst_ptrs.data(), cols.n_instance = uint32_t(inst_ptrs.size());
      cols.challenges = challenges.data(), cols.n_challenges = uint32_t(challenges.size());
      cols.beta = beta, cols.gamma = gamma, cols.theta = theta, cols.y = y;
      halo2_proofs::detail::check(ctx, h2b_evaluate_h_gates(dom.raw(), graph.get(), &cols, values.ptr()), "h2b_evaluate_h_gates");
      if (!z_ptrs.empty())
        halo2_proofs::detail::check(ctx, h2b_evaluate_h_permutation(dom.raw(), &cols, ctype.data(), cidx.data(), uint32_t(pcols.size()), sigma_ptrs.data(),
                                                                    z_ptrs.data(), uint32_t(z_ptrs.size()), uint32_t(chunk_len), uint32_t(bf), l0.ptr(),
                                                                    l_last.ptr(), l_active_row.ptr(), values.ptr()),
                                    "h2b_evaluate_h_permutation");
      for (size_t li = 0; li < lookups[ci].size(); ++li) {  // evaluation.rs:446-519: the three cosets live only here
        const auto& lk = lookups[ci][li];
        const DeviceVec product(dom.coeff_to_extended(lk.product_poly).values), pin(dom.coeff_to_extended(lk.permuted_input_poly).values),
            ptab(dom.coeff_to_extended(lk.permuted_table_poly).values);
        halo2_proofs::detail::check(ctx, h2b_evaluate_h_lookup(dom.raw(), pk.ev->lookups[li].compile().get(), &cols, product.ptr(), pin.ptr(), ptab.ptr(),
                                                               l0.ptr(), l_last.ptr(), l_active_row.ptr(), values.ptr()),
                                    "h2b_evaluate_h_lookup");
      }
    }
    h_ext = values.download();
  }
  // vanishing.construct (vanishing/prover.rs:69-121): divide by t(X), back to coefficients, n-sized pieces
  const std::vector<Fr> h_coeff = dom.divide_by_vanishing_poly_then_extended_to_coeff(Polynomial<ExtendedLagrangeCoeff>{std::move(h_ext)});
  const size_t n_pieces = h_coeff.size() / n;
  for (size_t i = 0; i < n_pieces; ++i) fr_random(rng);  // h_blinds
  std::vector<Polynomial<Coeff>> pieces;
  for (size_t i = 0; i < n_pieces; ++i) {
    pieces.push_back({std::vector<Fr>(h_coeff.begin() + i * n, h_coeff.begin() + (i + 1) * n)});
    transcript.write_point(params.commit(pieces.back()).to_affine());
  }

  const Fr x = transcript.squeeze_challenge_scalar();  // :525
  const Fr xn = x.pow_vartime(n);

  // ---- evaluations (:548-581) ----
  for (size_t ci = 0; ci < n_circuits; ++ci)
    for (const auto& q : cs.advice_queries) transcript.write_scalar(eval(advice_polys[ci][q.first.index], rot(x, q.second)));
  for (const auto& q : cs.fixed_queries) transcript.write_scalar(eval(pk.fixed_polys[q.first.index], rot(x, q.second)));
  // vanishing.evaluate (vanishing/prover.rs:124-152): h_poly = fold(pieces.rev(), acc * xn + piece)
  Polynomial<Coeff> h_poly = dom.empty_coeff();
  for (size_t i = n_pieces; i-- > 0;) h_poly = (h_poly * xn) + pieces[i];
  transcript.write_scalar(eval(random_poly, x));
  for (const auto& p : pk.permutation_polys) transcript.write_scalar(eval(p, x));  // permutation/prover.rs:208-219
  const Fr x_next = rot(x, 1), x_last = rot(x, -int32_t(bf + 1));
  for (const auto& sets : permutations)                                           // permutation/prover.rs:222-266
    for (size_t si = 0; si < sets.size(); ++si) {
      transcript.write_scalar(eval(sets[si].poly, x));
      transcript.write_scalar(eval(sets[si].poly, x_next));
      if (si + 1 < sets.size()) transcript.write_scalar(eval(sets[si].poly, x_last));
    }
  const Fr x_inv = rot(x, -1);
  for (const auto& lks : lookups)  // :588-595, lookup/prover.rs:253-283
    for (const auto& lk : lks) {
      transcript.write_scalar(eval(lk.product_poly, x)), transcript.write_scalar(eval(lk.product_poly, x_next));
      transcript.write_scalar(eval(lk.permuted_input_poly, x)), transcript.write_scalar(eval(lk.permuted_input_poly, x_inv));
      transcript.write_scalar(eval(lk.permuted_table_poly, x));
    }

  // ---- the opening queries in the reference's order (:596-645) ----
  std::vector<ProverQuery> queries;
  for (size_t ci = 0; ci < n_circuits; ++ci) {
    for (const auto& q : cs.advice_queries) queries.push_back({rot(x, q.second), &advice_polys[ci][q.first.index]});
    const auto& sets = permutations[ci];
    for (const auto& st : sets) queries.push_back({x, &st.poly}), queries.push_back({x_next, &st.poly});
    for (size_t si = sets.size(); si-- > 1;) queries.push_back({x_last, &sets[si - 1].poly});  // .rev().skip(1)
    for (const auto& lk : lookups[ci]) {  // lookup/prover.rs:286-323
      queries.push_back({x, &lk.product_poly}), queries.push_back({x, &lk.permuted_input_poly}), queries.push_back({x, &lk.permuted_table_poly});
      queries.push_back({x_inv, &lk.permuted_input_poly}), queries.push_back({x_next, &lk.product_poly});
    }
  }
  for (const auto& q : cs.fixed_queries) queries.push_back({rot(x, q.second), &pk.fixed_polys[q.first.index]});
  for (const auto& p : pk.permutation_polys) queries.push_back({x, &p});
  queries.push_back({x, &h_poly});
  queries.push_back({x, &random_poly});
  if (scheme == Multiopen::GWC)
    kzg::multiopen::ProverGWC(params).create_proof(transcript, queries);
  else
    kzg::multiopen::ProverSHPLONK(params).create_proof(transcript, queries);
}

/// single-phase form: advice[i][c] = the assigned values of advice column c of circuit i
template <class Rng>
inline void create_proof(const poly::kzg::ParamsKZG& params, const ProvingKey& pk,
                         const std::vector<std::vector<std::vector<Fr>>>& advice,
                         const std::vector<std::vector<std::vector<Fr>>>& instances, Rng& rng,
                         transcript::Blake2bWrite& transcript, Multiopen scheme = Multiopen::GWC) {
  if (advice.size() != instances.size()) throw Panic("one instance list per circuit");
  const Witness w = [&](size_t ci, uint32_t, const std::vector<Fr>&) { return advice[ci]; };
  create_proof(params, pk, w, instances, rng, transcript, scheme);
}

}  // namespace plonk
}  // namespace halo2_proofs

#endif  // HALO2_B200_PLONK_HPP
