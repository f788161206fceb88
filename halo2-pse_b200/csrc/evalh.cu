// Quotient-polynomial evaluation on the device:
//   Evaluator::evaluate_h                 halo2_proofs/src/plonk/evaluation.rs:280-522
//   GraphEvaluator::evaluate              halo2_proofs/src/plonk/evaluation.rs:700-746
//   Calculation::evaluate / ValueSource   halo2_proofs/src/plonk/evaluation.rs:36-180
//
// The reference interprets a list of `Calculation`s per row of the extended domain on rayon
// threads, then runs two hand-written loops for the permutation and lookup constraints.  Its
// inputs are exactly the cosets the NTT has just produced, so on the device they never leave HBM.
//
// Device schedule: one thread per row (consecutive threads = consecutive rows, every column read
// is a coalesced 32-byte access).  The calculation list is compiled once (h2b_graph_new) into
// three-address instructions:
//   * Horner(start, parts, factor) is unrolled into MULADD chains,
//   * constants, challenges, beta, gamma, theta and y are folded into one table of uniforms,
//   * the intermediates (one per calculation in the reference) are renamed onto the smallest
//     number of live slots by a linear scan, so that a thread's working set fits in shared
//     memory: slot s of thread t lives at [(2 s + half) * blockDim + t] as two 16-byte halves,
//     conflict-free for every warp because the slot index is uniform across the block.
// Every thread executes the same instruction stream, so the interpreter has no divergence; the
// instruction words are broadcast loads.  Field results are canonical, hence bit-identical to the
// reference's whatever the evaluation order.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>

namespace h2b {

// ---- compiled instruction format ---------------------------------------------------------
// word x: op | dst_slot << 8;  words y, z, w: sources.
// source: kind << 30 | payload;  kind 0 uniform(idx)  1 slot(idx)  2 column(rot << 22 | col)  3 previous value
enum : uint32_t { EV_ADD = 0, EV_SUB = 1, EV_MUL = 2, EV_DBL = 3, EV_NEG = 4, EV_MOV = 5, EV_MULADD = 6,
                  EV_PREFETCH = 7,  // slot <- column value, asynchronously (cp.async straight into the slot file)
                  EV_WAIT = 8 };    // all prefetches of this thread have landed
enum : uint32_t { SRC_UNIFORM = 0, SRC_SLOT = 1, SRC_COLUMN = 2, SRC_PREV = 3 };
static const uint32_t kMaxRot = 255, kMaxCol = (1u << 22) - 1;
static const size_t kEvalSmemCap = 200 * 1024;  // slot file per block (128 threads x 32 B per slot)

struct EvalProgram {
  const uint4* code;     // n_instr instructions
  uint32_t n_instr;
  uint32_t result_src;   // source word of the result (slot of the last calculation, or uniform 0)
  const Fr* uniforms;
  const Fr* const* cols;  // fixed ++ advice ++ instance device pointers
  const int32_t* rot;     // rotation * rot_scale, per rotation index
  int async_ok;           // the slot file is in shared memory: EV_PREFETCH may use cp.async
};

struct SlotFile {
  uint4* base;  // shared (or global overflow) storage of this block
  uint32_t stride;
  H2B_D Fr get(uint32_t s) const {
    const uint4 lo = base[(2 * s) * stride + threadIdx.x], hi = base[(2 * s + 1) * stride + threadIdx.x];
    Fr r;
    r.v[0] = lo.x, r.v[1] = lo.y, r.v[2] = lo.z, r.v[3] = lo.w;
    r.v[4] = hi.x, r.v[5] = hi.y, r.v[6] = hi.z, r.v[7] = hi.w;
    return r;
  }
  // slot s <- the 32 bytes at g, without passing through registers (shared slot files only)
  H2B_D void put_async(uint32_t s, const Fr* g) const {
#if defined(__CUDA_ARCH__)
    const uint32_t lo = (uint32_t)__cvta_generic_to_shared(base + (2 * s) * stride + threadIdx.x);
    const uint32_t hi = (uint32_t)__cvta_generic_to_shared(base + (2 * s + 1) * stride + threadIdx.x);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lo), "l"(g) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(hi), "l"(reinterpret_cast<const uint4*>(g) + 1)
                 : "memory");
#else
    put(s, *g);
#endif
  }
  static H2B_D void wait_async() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
  }
  H2B_D void put(uint32_t s, const Fr& r) const {
    base[(2 * s) * stride + threadIdx.x] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    base[(2 * s + 1) * stride + threadIdx.x] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
};

H2B_D Fr ev_fetch(uint32_t src, const EvalProgram& p, const SlotFile& sf, uint64_t row, uint64_t mask,
                  const Fr& prev) {
  const uint32_t kind = src >> 30, payload = src & 0x3fffffffu;
  if (kind == SRC_SLOT) return sf.get(payload);
  if (kind == SRC_UNIFORM) return ld_fp_nc(p.uniforms + payload);
  if (kind == SRC_COLUMN) {
    // get_rotation_idx (evaluation.rs:32-34): rem_euclid by a power of two is a mask
    const uint64_t r = (uint64_t)((int64_t)row + (int64_t)p.rot[payload >> 22]) & mask;
    return ld_fp(p.cols[payload & kMaxCol] + r);
  }
  return prev;
}

// GraphEvaluator::evaluate for one row (evaluation.rs:700-746)
H2B_D Fr ev_run(const EvalProgram& p, const SlotFile& sf, uint64_t row, uint64_t mask, const Fr& prev) {
  for (uint32_t pc = 0; pc < p.n_instr; ++pc) {
    const uint4 ins = __ldg(p.code + pc);
    const uint32_t op = ins.x & 0xffu, dst = ins.x >> 8;
    if (op == EV_WAIT) {
      SlotFile::wait_async();
      continue;
    }
    if (op == EV_PREFETCH && p.async_ok) {
      const uint32_t payload = ins.y & 0x3fffffffu;
      const uint64_t r = (uint64_t)((int64_t)row + (int64_t)p.rot[payload >> 22]) & mask;
      sf.put_async(dst, p.cols[payload & kMaxCol] + r);
      continue;
    }
    const Fr a = ev_fetch(ins.y, p, sf, row, mask, prev);
    Fr r;
    if (op == EV_MUL || op == EV_MULADD) {
      r = mul(a, ev_fetch(ins.z, p, sf, row, mask, prev));
      if (op == EV_MULADD) r = add(r, ev_fetch(ins.w, p, sf, row, mask, prev));
    } else if (op == EV_ADD) {
      r = add(a, ev_fetch(ins.z, p, sf, row, mask, prev));
    } else if (op == EV_SUB) {
      r = sub(a, ev_fetch(ins.z, p, sf, row, mask, prev));
    } else if (op == EV_DBL) {
      r = add(a, a);
    } else if (op == EV_NEG) {
      r = neg(a);
    } else {
      r = a;
    }
    sf.put(dst, r);
  }
  return ev_fetch(p.result_src, p, sf, row, mask, prev);
}

struct LookupTail {
  const Fr *product, *permuted_input, *permuted_table, *l0, *l_last, *l_active_row;
  Fr beta, gamma, y;
  int32_t rot_next, rot_prev;  // +-rot_scale
};

// mode 0 (custom gates, evaluation.rs:336-362): values[row] = graph(previous_value = values[row])
// mode 1 (one lookup, evaluation.rs:462-518):   table_value = graph(previous_value = 0), then the five
//                                              lookup constraints folded into values[row] with y
__global__ void __launch_bounds__(128)
    evalh_graph_kernel(EvalProgram p, Fr* values, uint64_t size, uint4* overflow, uint32_t n_slots, int mode,
                       LookupTail lt) {
  H2B_DYN_SMEM(smem);
  SlotFile sf;
  sf.stride = blockDim.x;
  sf.base = overflow ? overflow + (size_t)blockIdx.x * 2 * n_slots * blockDim.x : reinterpret_cast<uint4*>(smem);
  const uint64_t mask = size - 1;
  for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < size;
       row += (uint64_t)gridDim.x * blockDim.x) {
    const Fr v = ld_fp(values + row);
    if (mode == 0) {
      st_fp(values + row, ev_run(p, sf, row, mask, v));
      continue;
    }
    const Fr table_value = ev_run(p, sf, row, mask, Fr::zero());
    const uint64_t r_next = (uint64_t)((int64_t)row + lt.rot_next) & mask;
    const uint64_t r_prev = (uint64_t)((int64_t)row + lt.rot_prev) & mask;
    const Fr z = ld_fp(lt.product + row), a = ld_fp(lt.permuted_input + row), s = ld_fp(lt.permuted_table + row);
    const Fr l0 = ld_fp(lt.l0 + row), l_last = ld_fp(lt.l_last + row), l_act = ld_fp(lt.l_active_row + row);
    const Fr a_minus_s = sub(a, s);
    Fr acc = v;
    // l_0(X) * (1 - z(X)) = 0
    acc = add(mul(acc, lt.y), mul(sub(Fr::one(), z), l0));
    // l_last(X) * (z(X)^2 - z(X)) = 0
    acc = add(mul(acc, lt.y), mul(sub(mul(z, z), z), l_last));
    // (1 - (l_last + l_blind)) * (z(wX) (a'(X) + beta) (s'(X) + gamma) - z(X) * table_value) = 0
    const Fr left = mul(mul(ld_fp(lt.product + r_next), add(a, lt.beta)), add(s, lt.gamma));
    acc = add(mul(acc, lt.y), mul(sub(left, mul(z, table_value)), l_act));
    // l_0(X) * (a'(X) - s'(X)) = 0
    acc = add(mul(acc, lt.y), mul(a_minus_s, l0));
    // (1 - (l_last + l_blind)) * (a'(X) - s'(X)) * (a'(X) - a'(w^-1 X)) = 0
    acc = add(mul(acc, lt.y), mul(mul(a_minus_s, sub(a, ld_fp(lt.permuted_input + r_prev))), l_act));
    st_fp(values + row, acc);
  }
}

struct PermArgs {
  const Fr* const* col_values;  // n_cols: the permutation's columns, as cosets
  const Fr* const* sigma;       // n_cols: pk.permutation.cosets
  const Fr* const* z;           // n_sets: permutation_product_coset
  uint32_t n_cols, n_sets, chunk_len;
  const Fr *l0, *l_last, *l_active_row;
  const Fr *tw_lo, *tw_hi;  // extended_omega^i two-level table
  uint32_t tw_h;
  Fr y, beta, gamma, delta_start, delta;  // delta_start = beta * ZETA, delta = Fr::DELTA
  int32_t rot_next, rot_last;
};

// Permutation constraints, evaluation.rs:364-444.
__global__ void __launch_bounds__(128) evalh_permutation_kernel(PermArgs q, Fr* values, uint64_t size) {
  const uint64_t mask = size - 1;
  for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < size;
       row += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r_next = (uint64_t)((int64_t)row + q.rot_next) & mask;
    const uint64_t r_last = (uint64_t)((int64_t)row + q.rot_last) & mask;
    const Fr l0 = ld_fp(q.l0 + row);
    Fr acc = ld_fp(values + row);
    // l_0(X) * (1 - z_0(X)) = 0
    acc = add(mul(acc, q.y), mul(sub(Fr::one(), ld_fp(q.z[0] + row)), l0));
    // l_last(X) * (z_l(X)^2 - z_l(X)) = 0
    {
      const Fr zl = ld_fp(q.z[q.n_sets - 1] + row);
      acc = add(mul(acc, q.y), mul(sub(mul(zl, zl), zl), ld_fp(q.l_last + row)));
    }
    // l_0(X) * (z_i(X) - z_{i-1}(w^(last) X)) = 0
    for (uint32_t s = 1; s < q.n_sets; ++s)
      acc = add(mul(acc, q.y), mul(sub(ld_fp(q.z[s] + row), ld_fp(q.z[s - 1] + r_last)), l0));
    // (1 - (l_last + l_blind)) * (z_i(wX) prod (p + beta s_j + gamma) - z_i(X) prod (p + delta^j beta X + gamma))
    const Fr beta_term = mul(ld_fp_nc(q.tw_lo + (row & ((1ull << q.tw_h) - 1))), ld_fp_nc(q.tw_hi + (row >> q.tw_h)));
    Fr current_delta = mul(q.delta_start, beta_term);
    const Fr l_act = ld_fp(q.l_active_row + row);
    for (uint32_t s = 0; s < q.n_sets; ++s) {
      const uint32_t c0 = s * q.chunk_len, c1 = c0 + q.chunk_len < q.n_cols ? c0 + q.chunk_len : q.n_cols;
      Fr left = ld_fp(q.z[s] + r_next), right = ld_fp(q.z[s] + row);
      for (uint32_t c = c0; c < c1; ++c) {
        const Fr v = ld_fp(q.col_values[c] + row);
        left = mul(left, add(add(v, mul(q.beta, ld_fp(q.sigma[c] + row))), q.gamma));
        right = mul(right, add(add(v, current_delta), q.gamma));
        current_delta = mul(current_delta, q.delta);
      }
      acc = add(mul(acc, q.y), mul(sub(left, right), l_act));
    }
    st_fp(values + row, acc);
  }
}

}  // namespace h2b

using namespace h2b;



// A compiled GraphEvaluator (evaluation.rs:193-202): three-address code over renamed slots.  Uniform and
// column sources stay symbolic (kind + index) until an evaluation fixes the table layouts.
struct h2b_graph {
  h2b_ctx* ctx = nullptr;
  struct Instr {
    uint32_t op, dst_slot;
    uint32_t kind[3], a[3], b[3];  // ValueSource per operand (kind 0xff = unused); intermediates already as slots
  };
  std::vector<Instr> code;
  std::vector<Fr> constants;
  std::vector<int32_t> rotations;
  uint32_t n_slots = 1;
  int64_t result_slot = -1;  // -1: no calculations, the result is zero (evaluation.rs:740-744)
  uint32_t need_fixed = 0, need_advice = 0, need_instance = 0, need_challenges = 0;  // 1 + largest index used
  // device copy of the code for the last table layout it was evaluated with
  uint4* d_code = nullptr;
  uint32_t lay_fixed = ~0u, lay_advice = ~0u, lay_instance = ~0u, lay_challenges = ~0u;
};

namespace {

// ValueSource as (kind, a, b) in the order of the reference's enum (evaluation.rs:38-61)
enum : uint32_t { VS_CONSTANT, VS_INTERMEDIATE, VS_FIXED, VS_ADVICE, VS_INSTANCE, VS_CHALLENGE, VS_BETA, VS_GAMMA,
                  VS_THETA, VS_Y, VS_PREVIOUS, VS_NONE = 0xff };
// Calculation in the order of the reference's enum (evaluation.rs:110-127)
enum : uint32_t { CALC_ADD, CALC_SUB, CALC_MUL, CALC_SQUARE, CALC_DOUBLE, CALC_NEGATE, CALC_HORNER, CALC_STORE };

struct VSrc {
  uint32_t kind, a, b;
};
struct PreInstr {  // before slot renaming: intermediates by their reference index
  uint32_t op, target;
  VSrc s[3];
  int ns;
};

const Fr* as_fr(const h2b_fr* p) { return reinterpret_cast<const Fr*>(p); }
Fr* as_fr(h2b_fr* p) { return reinterpret_cast<Fr*>(p); }
Fr load_fr(const h2b_fr& x) {
  Fr r;
  memcpy(&r, &x, sizeof(Fr));
  return r;
}
Fr fr_from_canonical(const uint64_t l[4]) {
  Fr a;
  for (int i = 0; i < 4; ++i) {
    a.v[2 * i] = (uint32_t)l[i];
    a.v[2 * i + 1] = (uint32_t)(l[i] >> 32);
  }
  return to_mont(a);
}
// Fr::ZETA and Fr::DELTA = 7^(2^28), canonical
const uint64_t kZeta[4] = {0x8b17ea66b99c90ddull, 0x5bfc41088d8daaa7ull, 0xb3c4d79d41a91758ull, 0x0ull};
const uint64_t kDelta[4] = {0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull, 0x64ec26aad4c86e71ull, 0x09226b6e22c6f0caull};

// scratch layout of one evaluation call inside ctx->scratch (all 32-byte aligned)
struct CallTables {
  Fr* d_uniforms;
  const Fr** d_cols;
  int32_t* d_rot;
  const Fr** d_aux;  // permutation pointer lists
};

size_t align32(size_t x) { return (x + 31) & ~(size_t)31; }

}  // namespace

extern "C" int h2b_graph_new(h2b_ctx* ctx, const uint32_t* calc, size_t n_words, const h2b_fr* constants,
                             uint32_t n_constants, const int32_t* rotations, uint32_t n_rotations,
                             uint32_t num_intermediates, h2b_graph** out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!out || (n_words && !calc) || (n_constants && !constants) || (n_rotations && !rotations))
    return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n_rotations > kMaxRot + 1) return fail(ctx, H2B_ERR_ARG, "more than 256 distinct rotations");
  std::vector<PreInstr> pre;
  std::vector<int64_t> last_use(num_intermediates, -1);
  std::vector<char> defined(num_intermediates, 0);
  h2b_graph* g = new h2b_graph();
  g->ctx = ctx;
  auto bad = [&](const char* msg) {
    delete g;
    return fail(ctx, H2B_ERR_ARG, msg);
  };
  size_t pos = 0;
  bool ok = true;
  auto rd = [&]() -> uint32_t {
    if (pos >= n_words) {
      ok = false;
      return 0;
    }
    return calc[pos++];
  };
  auto rd_src = [&]() -> VSrc {
    VSrc s;
    s.kind = rd(), s.a = rd(), s.b = rd();
    return s;
  };
  auto check_src = [&](const VSrc& s, size_t at) -> bool {
    switch (s.kind) {
      case VS_CONSTANT: return s.a < n_constants;
      case VS_INTERMEDIATE:
        if (s.a >= num_intermediates || !defined[s.a]) return false;
        last_use[s.a] = (int64_t)at;
        return true;
      case VS_FIXED: g->need_fixed = std::max(g->need_fixed, s.a + 1); return s.b < n_rotations && s.a < kMaxCol;
      case VS_ADVICE: g->need_advice = std::max(g->need_advice, s.a + 1); return s.b < n_rotations && s.a < kMaxCol;
      case VS_INSTANCE:
        g->need_instance = std::max(g->need_instance, s.a + 1);
        return s.b < n_rotations && s.a < kMaxCol;
      case VS_CHALLENGE: g->need_challenges = std::max(g->need_challenges, s.a + 1); return s.a < (1u << 20);
      case VS_BETA: case VS_GAMMA: case VS_THETA: case VS_Y: case VS_PREVIOUS: return true;
      default: return false;
    }
  };
  int64_t last_target = -1;
  bool single_assignment = true;  // every intermediate is the target of exactly one calculation
  while (pos < n_words && ok) {
    const uint32_t op = rd(), target = rd();
    if (!ok || target >= num_intermediates) return bad("calculation target out of range");
    std::vector<PreInstr> emit;
    auto push = [&](uint32_t dop, std::initializer_list<VSrc> srcs) {
      PreInstr pi;
      pi.op = dop, pi.target = target, pi.ns = 0;
      for (const VSrc& s : srcs) pi.s[pi.ns++] = s;
      emit.push_back(pi);
    };
    const VSrc self = {VS_INTERMEDIATE, target, 0};
    switch (op) {
      case CALC_ADD: { VSrc a = rd_src(), b = rd_src(); push(EV_ADD, {a, b}); break; }
      case CALC_SUB: { VSrc a = rd_src(), b = rd_src(); push(EV_SUB, {a, b}); break; }
      case CALC_MUL: { VSrc a = rd_src(), b = rd_src(); push(EV_MUL, {a, b}); break; }
      case CALC_SQUARE: { VSrc a = rd_src(); push(EV_MUL, {a, a}); break; }
      case CALC_DOUBLE: { VSrc a = rd_src(); push(EV_DBL, {a}); break; }
      case CALC_NEGATE: { VSrc a = rd_src(); push(EV_NEG, {a}); break; }
      case CALC_STORE: { VSrc a = rd_src(); push(EV_MOV, {a}); break; }
      case CALC_HORNER: {  // value = start; for part: value = value * factor + part   (evaluation.rs:169-176)
        const VSrc start = rd_src(), factor = rd_src();
        const uint32_t nparts = rd();
        if (!ok) break;
        if (nparts == 0) push(EV_MOV, {start});
        for (uint32_t i = 0; i < nparts && ok; ++i) {
          const VSrc part = rd_src();
          push(EV_MULADD, {i == 0 ? start : self, factor, part});
        }
        break;
      }
      default: return bad("unknown calculation opcode");
    }
    if (!ok) break;
    if (defined[target]) single_assignment = false;
    for (PreInstr& pi : emit) {
      for (int i = 0; i < pi.ns; ++i)
        if (!check_src(pi.s[i], pre.size())) return bad("value source out of range or read before it is computed");
      defined[target] = 1;
      pre.push_back(pi);
    }
    last_target = target;
  }
  if (!ok) return bad("truncated calculation stream");

  // ---- slot renaming: linear scan over the live ranges of the intermediates ----
  auto allocate = [&](const std::vector<PreInstr>& order, std::vector<h2b_graph::Instr>& code, uint32_t& n_slots,
                      int64_t& result_slot) {
    std::vector<int64_t> last(num_intermediates, -1);
    for (size_t at = 0; at < order.size(); ++at)
      for (int i = 0; i < order[at].ns; ++i)
        if (order[at].s[i].kind == VS_INTERMEDIATE) last[order[at].s[i].a] = (int64_t)at;
    if (last_target >= 0) last[last_target] = (int64_t)order.size();  // the result is read after the program
    std::vector<int64_t> slot_of(num_intermediates, -1);
    std::vector<uint32_t> free_slots;
    std::multimap<int64_t, uint32_t> expiring;  // last use -> intermediate
    n_slots = 0;
    code.clear();
    for (size_t at = 0; at < order.size(); ++at) {
      const PreInstr& pi = order[at];
      h2b_graph::Instr in;
      in.op = pi.op;
      in.dst_slot = 0;
      for (int i = 0; i < 3; ++i) {
        in.kind[i] = VS_NONE, in.a[i] = in.b[i] = 0;
        if (i >= pi.ns) continue;
        in.kind[i] = pi.s[i].kind, in.a[i] = pi.s[i].a, in.b[i] = pi.s[i].b;
        if (pi.s[i].kind == VS_INTERMEDIATE) in.a[i] = (uint32_t)slot_of[pi.s[i].a];
      }
      if (pi.op == EV_WAIT) {
        code.push_back(in);
        continue;
      }
      // Slots whose last reader is this instruction are free for its destination: every source is read
      // before the destination is written.
      while (!expiring.empty() && expiring.begin()->first <= (int64_t)at) {
        free_slots.push_back((uint32_t)slot_of[expiring.begin()->second]);
        expiring.erase(expiring.begin());
      }
      if (slot_of[pi.target] < 0) {
        uint32_t sl;
        if (!free_slots.empty()) {
          sl = free_slots.back();
          free_slots.pop_back();
        } else {
          sl = n_slots++;
        }
        slot_of[pi.target] = sl;
        expiring.emplace(last[pi.target] >= 0 ? last[pi.target] : (int64_t)at + 1, pi.target);
      }
      in.dst_slot = (uint32_t)slot_of[pi.target];
      code.push_back(in);
    }
    result_slot = last_target >= 0 ? slot_of[last_target] : -1;
  };
  // Preferred order: every column read hoisted to the top as an asynchronous copy straight into its slot
  // (all of a row's loads in flight at once instead of one round trip to HBM per Store), one wait, then the
  // arithmetic.  Kept only while the longer live ranges still fit a shared-memory slot file.
  uint32_t n_slots = 0;
  bool hoisted = false;
  if (single_assignment && getenv("H2B_EVALH_NO_PREFETCH") == nullptr) {
    std::vector<PreInstr> order, rest;
    for (const PreInstr& pi : pre) {
      const bool col = pi.op == EV_MOV && (pi.s[0].kind == VS_FIXED || pi.s[0].kind == VS_ADVICE ||
                                           pi.s[0].kind == VS_INSTANCE);
      if (col) {
        PreInstr q = pi;
        q.op = EV_PREFETCH;
        order.push_back(q);
      } else {
        rest.push_back(pi);
      }
    }
    if (!order.empty()) {
      PreInstr w;
      w.op = EV_WAIT, w.target = 0, w.ns = 0;
      order.push_back(w);
      order.insert(order.end(), rest.begin(), rest.end());
      allocate(order, g->code, n_slots, g->result_slot);
      hoisted = (size_t)n_slots * 32 * 128 <= kEvalSmemCap;
    }
  }
  if (!hoisted) allocate(pre, g->code, n_slots, g->result_slot);
  if (n_slots >= (1u << 22)) return bad("too many live intermediates");
  g->n_slots = std::max(n_slots, 1u);
  g->constants.resize(n_constants);
  for (uint32_t i = 0; i < n_constants; ++i) g->constants[i] = load_fr(constants[i]);
  g->rotations.assign(rotations, rotations + n_rotations);
  *out = g;
  return H2B_OK;
}

extern "C" void h2b_graph_free(h2b_graph* g) {
  if (!g) return;
  if (g->d_code) {
    cudaSetDevice(g->ctx->device);
    cudaFree(g->d_code);
  }
  delete g;
}
extern "C" uint32_t h2b_graph_num_slots(const h2b_graph* g) { return g ? g->n_slots : 0; }
extern "C" uint32_t h2b_graph_num_instructions(const h2b_graph* g) { return g ? (uint32_t)g->code.size() : 0; }

namespace {

int check_columns(h2b_ctx* ctx, const h2b_eval_columns* c) {
  if (!c) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if ((c->n_fixed && !c->fixed) || (c->n_advice && !c->advice) || (c->n_instance && !c->instance) ||
      (c->n_challenges && !c->challenges))
    return fail(ctx, H2B_ERR_ARG, "null column list");
  return H2B_OK;
}

// uniforms = [constants][challenges][beta, gamma, theta, y][0]
int run_graph(h2b_domain* dom, h2b_graph* g, const h2b_eval_columns* c, Fr* values, int mode, LookupTail lt,
              bool lagrange_basis = false) {
  h2b_ctx* ctx = dom->ctx;
  if (g->need_fixed > c->n_fixed || g->need_advice > c->n_advice || g->need_instance > c->n_instance ||
      g->need_challenges > c->n_challenges)
    return fail(ctx, H2B_ERR_LENGTH, "the graph queries a column or challenge that was not supplied");
  const uint32_t n_const = (uint32_t)g->constants.size();
  const uint32_t n_cols = c->n_fixed + c->n_advice + c->n_instance;
  const uint32_t n_uni = n_const + c->n_challenges + 5;
  const uint64_t size = 1ull << (lagrange_basis ? dom->k : dom->extended_k);
  const int32_t rot_scale = lagrange_basis ? 1 : 1 << (dom->extended_k - dom->k);

  // code for this table layout
  if (!g->d_code || g->lay_fixed != c->n_fixed || g->lay_advice != c->n_advice || g->lay_instance != c->n_instance ||
      g->lay_challenges != c->n_challenges) {
    std::vector<uint4> code(g->code.size());
    auto enc = [&](uint32_t kind, uint32_t a, uint32_t b) -> uint32_t {
      switch (kind) {
        case VS_CONSTANT: return (SRC_UNIFORM << 30) | a;
        case VS_INTERMEDIATE: return (SRC_SLOT << 30) | a;
        case VS_FIXED: return (SRC_COLUMN << 30) | (b << 22) | a;
        case VS_ADVICE: return (SRC_COLUMN << 30) | (b << 22) | (c->n_fixed + a);
        case VS_INSTANCE: return (SRC_COLUMN << 30) | (b << 22) | (c->n_fixed + c->n_advice + a);
        case VS_CHALLENGE: return (SRC_UNIFORM << 30) | (n_const + a);
        case VS_BETA: case VS_GAMMA: case VS_THETA: case VS_Y:
          return (SRC_UNIFORM << 30) | (n_const + c->n_challenges + (kind - VS_BETA));
        case VS_PREVIOUS: return SRC_PREV << 30;
        default: return SRC_UNIFORM << 30;
      }
    };
    for (size_t i = 0; i < g->code.size(); ++i) {
      const h2b_graph::Instr& in = g->code[i];
      code[i] = make_uint4(in.op | (in.dst_slot << 8), enc(in.kind[0], in.a[0], in.b[0]),
                           enc(in.kind[1], in.a[1], in.b[1]), enc(in.kind[2], in.a[2], in.b[2]));
    }
    if (n_cols > kMaxCol) return fail(ctx, H2B_ERR_ARG, "too many columns");
    if (g->d_code) H2B_CUDA(ctx, cudaFree(g->d_code));
    g->d_code = nullptr;
    H2B_CUDA(ctx, cudaMalloc((void**)&g->d_code, std::max<size_t>(code.size(), 1) * sizeof(uint4)));
    if (!code.empty())
      H2B_CUDA(ctx, cudaMemcpyAsync(g->d_code, code.data(), code.size() * sizeof(uint4), cudaMemcpyHostToDevice,
                                    ctx->stream));
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->lay_fixed = c->n_fixed, g->lay_advice = c->n_advice, g->lay_instance = c->n_instance;
    g->lay_challenges = c->n_challenges;
  }

  // per-call tables in one host block -> one copy into the scratch
  const size_t off_cols = align32(n_uni * sizeof(Fr));
  const size_t off_rot = align32(off_cols + (n_cols + 1) * sizeof(void*));
  const size_t tab_bytes = align32(off_rot + (g->rotations.size() + 1) * sizeof(int32_t));
  std::vector<unsigned char> host(tab_bytes, 0);
  Fr* uni = reinterpret_cast<Fr*>(host.data());
  for (uint32_t i = 0; i < n_const; ++i) uni[i] = g->constants[i];
  for (uint32_t i = 0; i < c->n_challenges; ++i) uni[n_const + i] = load_fr(c->challenges[i]);
  uni[n_const + c->n_challenges + 0] = load_fr(c->beta);
  uni[n_const + c->n_challenges + 1] = load_fr(c->gamma);
  uni[n_const + c->n_challenges + 2] = load_fr(c->theta);
  uni[n_const + c->n_challenges + 3] = load_fr(c->y);
  uni[n_const + c->n_challenges + 4] = Fr::zero();
  const void** cols = reinterpret_cast<const void**>(host.data() + off_cols);
  for (uint32_t i = 0; i < c->n_fixed; ++i) cols[i] = c->fixed[i];
  for (uint32_t i = 0; i < c->n_advice; ++i) cols[c->n_fixed + i] = c->advice[i];
  for (uint32_t i = 0; i < c->n_instance; ++i) cols[c->n_fixed + c->n_advice + i] = c->instance[i];
  for (uint32_t i = 0; i < n_cols; ++i)
    if (!cols[i]) return fail(ctx, H2B_ERR_ARG, "null column pointer");
  int32_t* rot = reinterpret_cast<int32_t*>(host.data() + off_rot);
  for (size_t i = 0; i < g->rotations.size(); ++i) {
    const int64_t r = (int64_t)g->rotations[i] * rot_scale;
    if (r > INT32_MAX || r < INT32_MIN) return fail(ctx, H2B_ERR_ARG, "rotation out of range");
    rot[i] = (int32_t)r;
  }

  // launch geometry: slots in shared memory when they fit, else in a global overflow area
  const uint32_t threads = 128;
  const size_t slot_bytes = (size_t)g->n_slots * 32 * threads;
  size_t smem_cap = kEvalSmemCap;
  if (const char* e = getenv("H2B_EVALH_SMEM_CAP")) smem_cap = (size_t)atoll(e);  // tests: force the overflow path
  const bool overflow = slot_bytes > smem_cap;
  uint64_t want = (size + threads - 1) / threads;
  const uint32_t per_sm =
      overflow ? 4 : (uint32_t)std::max<size_t>(1, std::min<size_t>(8, kEvalSmemCap / std::max<size_t>(slot_bytes, 1)));
  const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(want, cap);
  const size_t ovf_bytes = overflow ? (size_t)grid * slot_bytes : 0;
  H2B_TRY(ensure_scratch(ctx, tab_bytes + ovf_bytes + 64));
  unsigned char* d_tab = reinterpret_cast<unsigned char*>(ctx->scratch);
  H2B_CUDA(ctx, cudaMemcpyAsync(d_tab, host.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `host` is freed on return

  EvalProgram p;
  p.code = g->d_code;
  p.n_instr = (uint32_t)g->code.size();
  p.result_src = g->result_slot >= 0 ? ((SRC_SLOT << 30) | (uint32_t)g->result_slot)
                                     : ((SRC_UNIFORM << 30) | (n_const + c->n_challenges + 4));
  p.uniforms = reinterpret_cast<const Fr*>(d_tab);
  p.cols = reinterpret_cast<const Fr* const*>(d_tab + off_cols);
  p.rot = reinterpret_cast<const int32_t*>(d_tab + off_rot);
  uint4* d_ovf = overflow ? reinterpret_cast<uint4*>(d_tab + tab_bytes) : nullptr;
  p.async_ok = overflow ? 0 : 1;
  const size_t smem = overflow ? 0 : slot_bytes;
#ifndef H2B_EMU
  if (smem > 48 * 1024)
    H2B_CUDA(ctx, cudaFuncSetAttribute(evalh_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEvalSmemCap));
#endif
  if (ctx->profile) H2B_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  H2B_TRY(launch(ctx, evalh_graph_kernel, dim3(grid), dim3(threads), smem, p, values, size, d_ovf, g->n_slots, mode,
                 lt));
  if (ctx->profile) {
    H2B_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    H2B_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    H2B_CUDA(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev[0], ctx->ev[1]));
  }
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

}  // namespace

extern "C" int h2b_evaluate_h_gates(h2b_domain* dom, h2b_graph* graph, const h2b_eval_columns* cols,
                                    h2b_fr* values) {
  if (!dom) return H2B_ERR_ARG;
  h2b_ctx* ctx = dom->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!graph || !values) return fail(ctx, H2B_ERR_ARG, "null pointer");
  H2B_TRY(check_columns(ctx, cols));
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  LookupTail lt;
  memset(&lt, 0, sizeof lt);
  return run_graph(dom, graph, cols, as_fr(values), 0, lt);
}

extern "C" int h2b_graph_evaluate_lagrange(h2b_domain* dom, h2b_graph* graph, const h2b_eval_columns* cols,
                                           h2b_fr* values) {
  if (!dom) return H2B_ERR_ARG;
  h2b_ctx* ctx = dom->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!graph || !values) return fail(ctx, H2B_ERR_ARG, "null pointer");
  H2B_TRY(check_columns(ctx, cols));
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  LookupTail lt;
  memset(&lt, 0, sizeof lt);
  return run_graph(dom, graph, cols, as_fr(values), 0, lt, true);
}

extern "C" int h2b_evaluate_h_lookup(h2b_domain* dom, h2b_graph* graph, const h2b_eval_columns* cols,
                                     const h2b_fr* product_coset, const h2b_fr* permuted_input_coset,
                                     const h2b_fr* permuted_table_coset, const h2b_fr* l0, const h2b_fr* l_last,
                                     const h2b_fr* l_active_row, h2b_fr* values) {
  if (!dom) return H2B_ERR_ARG;
  h2b_ctx* ctx = dom->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!graph || !values || !product_coset || !permuted_input_coset || !permuted_table_coset || !l0 || !l_last ||
      !l_active_row)
    return fail(ctx, H2B_ERR_ARG, "null pointer");
  H2B_TRY(check_columns(ctx, cols));
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  LookupTail lt;
  lt.product = as_fr(product_coset), lt.permuted_input = as_fr(permuted_input_coset);
  lt.permuted_table = as_fr(permuted_table_coset);
  lt.l0 = as_fr(l0), lt.l_last = as_fr(l_last), lt.l_active_row = as_fr(l_active_row);
  lt.beta = load_fr(cols->beta), lt.gamma = load_fr(cols->gamma), lt.y = load_fr(cols->y);
  lt.rot_next = 1 << (dom->extended_k - dom->k);
  lt.rot_prev = -lt.rot_next;
  return run_graph(dom, graph, cols, as_fr(values), 1, lt);
}

extern "C" int h2b_evaluate_h_permutation(h2b_domain* dom, const h2b_eval_columns* cols, const uint32_t* column_type,
                                          const uint32_t* column_index, uint32_t n_columns,
                                          const h2b_fr* const* sigma_cosets, const h2b_fr* const* product_cosets,
                                          uint32_t n_sets, uint32_t chunk_len, uint32_t blinding_factors,
                                          const h2b_fr* l0, const h2b_fr* l_last, const h2b_fr* l_active_row,
                                          h2b_fr* values) {
  if (!dom) return H2B_ERR_ARG;
  h2b_ctx* ctx = dom->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  H2B_TRY(check_columns(ctx, cols));
  if (n_sets == 0) return H2B_OK;  // `if !sets.is_empty()`, evaluation.rs:366
  if (!values || !l0 || !l_last || !l_active_row || !product_cosets || (n_columns && (!column_type || !column_index ||
      !sigma_cosets)))
    return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (chunk_len == 0) return fail(ctx, H2B_ERR_ARG, "chunk_len = degree - 2 must be positive");
  if ((uint64_t)n_sets != ((uint64_t)n_columns + chunk_len - 1) / chunk_len)
    return fail(ctx, H2B_ERR_LENGTH, "n_sets does not match columns.chunks(chunk_len)");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t size = 1ull << dom->extended_k;
  const int32_t rot_scale = 1 << (dom->extended_k - dom->k);
  // pointer lists: [column values n_columns][sigma n_columns][z n_sets]
  std::vector<const void*> host(2 * (size_t)n_columns + n_sets);
  for (uint32_t i = 0; i < n_columns; ++i) {
    const void* p = nullptr;
    const uint32_t idx = column_index[i];
    switch (column_type[i]) {  // Any::{Advice, Fixed, Instance}, evaluation.rs:419-423
      case 0: p = idx < cols->n_advice ? cols->advice[idx] : nullptr; break;
      case 1: p = idx < cols->n_fixed ? cols->fixed[idx] : nullptr; break;
      case 2: p = idx < cols->n_instance ? cols->instance[idx] : nullptr; break;
      default: break;
    }
    if (!p || !sigma_cosets[i]) return fail(ctx, H2B_ERR_ARG, "permutation column not supplied");
    host[i] = p;
    host[n_columns + i] = sigma_cosets[i];
  }
  for (uint32_t s = 0; s < n_sets; ++s) {
    if (!product_cosets[s]) return fail(ctx, H2B_ERR_ARG, "null permutation product coset");
    host[2 * (size_t)n_columns + s] = product_cosets[s];
  }
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, dom->extended_omega, dom->extended_k, &tw));
  H2B_TRY(ensure_scratch(ctx, host.size() * sizeof(void*) + 64));
  H2B_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, host.data(), host.size() * sizeof(void*), cudaMemcpyHostToDevice,
                                ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PermArgs q;
  const Fr* const* d_list = reinterpret_cast<const Fr* const*>(ctx->scratch);
  q.col_values = d_list, q.sigma = d_list + n_columns, q.z = d_list + 2 * (size_t)n_columns;
  q.n_cols = n_columns, q.n_sets = n_sets, q.chunk_len = chunk_len;
  q.l0 = as_fr(l0), q.l_last = as_fr(l_last), q.l_active_row = as_fr(l_active_row);
  q.tw_lo = tw->d_lo, q.tw_hi = tw->d_hi, q.tw_h = tw->h;
  q.y = load_fr(cols->y), q.beta = load_fr(cols->beta), q.gamma = load_fr(cols->gamma);
  q.delta_start = mul(q.beta, fr_from_canonical(kZeta));  // evaluation.rs:370
  q.delta = fr_from_canonical(kDelta);
  q.rot_next = rot_scale;
  q.rot_last = -(int32_t)(blinding_factors + 1) * rot_scale;  // evaluation.rs:368
  const uint64_t want = (size + 127) / 128, cap = (uint64_t)ctx->sm_count * 16;
  if (ctx->profile) H2B_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  H2B_TRY(launch(ctx, evalh_permutation_kernel, dim3((uint32_t)std::min(want, cap)), dim3(128), 0, q, as_fr(values),
                 size));
  if (ctx->profile) {
    H2B_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    H2B_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    H2B_CUDA(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev[0], ctx->ev[1]));
  }
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}
