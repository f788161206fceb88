// ORACLE (test infrastructure -- never shipped, never linked by the product).
//
// C++ restatement of the reference's CPU algorithms for the two hot paths, with
// std::thread standing in for rayon.  It follows the reference line by line
// (same window rule, same per-thread chunking, same serial bit-reversal and
// twiddle scan, same recursive butterfly split), so it doubles as the timed CPU
// baseline ("port") of bench.py and as the big-k checker where the Python
// big-integer oracle (oracle/bn256.py) is too slow.
//
//   multiexp_serial / best_multiexp      halo2_proofs/src/arithmetic.rs:13-101, 132-159
//   best_fft / recursive_butterfly_..    halo2_proofs/src/arithmetic.rs:171-274
//   parallelize                          halo2_proofs/src/arithmetic.rs:371-388
//   EvaluationDomain::new                halo2_proofs/src/poly/domain.rs:39-142
//   lagrange_to_coeff / coeff_to_extended / extended_to_coeff /
//   divide_by_vanishing_poly / distribute_powers_zeta / ifft
//                                        halo2_proofs/src/poly/domain.rs:226-361
//   ParamsKZG::commit_lagrange / commit  halo2_proofs/src/poly/kzg/commitment.rs:281-292, 327-334
//   Evaluator::evaluate_h, GraphEvaluator::evaluate, Calculation::evaluate
//                                        halo2_proofs/src/plonk/evaluation.rs:36-180, 280-522, 700-746
//
// The field and curve arithmetic is NOT in the reference tree: it is
// halo2curves 0.3.1 (halo2_proofs/Cargo.toml:51).  It is restated here from the
// published definition of bn256: 4 x 64-bit Montgomery limbs (R = 2^256),
// Jacobian coordinates with the standard a = 0 formulas (dbl-2009-l,
// add-2007-bl, madd-2007-bl).  PARITY UNPINNED by reference golden vectors for
// bn256 (the reference has none); this file is pinned against oracle/bn256.py
// and tests/golden/kat_bn256.json by tests/test_oracle.py.
//
// Build: make -C oracle   ->  oracle/_build/libh2b_oracle.so
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <thread>
#include <array>
#include <vector>

typedef unsigned __int128 u128;

namespace {

// ---------------------------------------------------------------------------
// 256-bit Montgomery field, 4 x u64
// ---------------------------------------------------------------------------
struct FrP {
  static constexpr uint64_t M[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull,
                                    0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull,
                                      0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull,
                                     0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  static constexpr uint64_t INV = 0xc2e1f593efffffffull;
};
struct FqP {
  static constexpr uint64_t M[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull,
                                    0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull,
                                      0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull,
                                     0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
  static constexpr uint64_t INV = 0x87d20782e4866389ull;
};
constexpr uint64_t FrP::M[4], FrP::ONE[4], FrP::R2[4], FqP::M[4], FqP::ONE[4], FqP::R2[4];

template <class P>
struct F {
  uint64_t l[4];
  static F zero() { return F{{0, 0, 0, 0}}; }
  static F one() { return F{{P::ONE[0], P::ONE[1], P::ONE[2], P::ONE[3]}}; }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const F& o) const {
    return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3];
  }
  bool operator!=(const F& o) const { return !(*this == o); }
};

template <class P>
inline void cond_sub(F<P>& a) {  // a < 2p  ->  a mod p
  uint64_t t[4];
  u128 b = 0;
  for (int i = 0; i < 4; ++i) {
    u128 d = (u128)a.l[i] - P::M[i] - (uint64_t)b;
    t[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
  if (!b) memcpy(a.l, t, 32);
}

template <class P>
inline F<P> fadd(const F<P>& a, const F<P>& b) {
  F<P> r;
  u128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (u128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  cond_sub(r);
  return r;
}

template <class P>
inline F<P> fsub(const F<P>& a, const F<P>& b) {
  F<P> r;
  u128 bw = 0;
  for (int i = 0; i < 4; ++i) {
    u128 d = (u128)a.l[i] - b.l[i] - (uint64_t)bw;
    r.l[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
      c += (u128)r.l[i] + P::M[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}

template <class P>
inline F<P> fneg(const F<P>& a) {
  return a.is_zero() ? a : fsub(F<P>::zero(), a);
}

template <class P>
inline F<P> fdbl(const F<P>& a) {
  return fadd(a, a);
}

// CIOS Montgomery product
template <class P>
inline F<P> fmul(const F<P>& a, const F<P>& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    u128 c = 0;
    for (int j = 0; j < 4; ++j) {
      c += (u128)a.l[j] * b.l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    const uint64_t m = t[0] * P::INV;
    c = (u128)m * P::M[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; ++j) {
      c += (u128)m * P::M[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  F<P> r = {{t[0], t[1], t[2], t[3]}};
  cond_sub(r);
  return r;
}

template <class P>
inline F<P> fsqr(const F<P>& a) {
  return fmul(a, a);
}

template <class P>
F<P> fpow(const F<P>& a, const uint64_t e[4]) {
  F<P> r = F<P>::one();
  for (int i = 3; i >= 0; --i)
    for (int b = 63; b >= 0; --b) {
      r = fsqr(r);
      if ((e[i] >> b) & 1) r = fmul(r, a);
    }
  return r;
}

template <class P>
F<P> finv(const F<P>& a) {
  uint64_t e[4] = {P::M[0] - 2, P::M[1], P::M[2], P::M[3]};
  return fpow(a, e);
}

template <class P>
F<P> from_u64(uint64_t v) {
  F<P> x = {{v, 0, 0, 0}};
  F<P> r2 = {{P::R2[0], P::R2[1], P::R2[2], P::R2[3]}};
  return fmul(x, r2);
}

template <class P>
F<P> from_canonical(const uint64_t v[4]) {
  F<P> x = {{v[0], v[1], v[2], v[3]}};
  F<P> r2 = {{P::R2[0], P::R2[1], P::R2[2], P::R2[3]}};
  return fmul(x, r2);
}

// to_repr(): canonical little-endian bytes                       arithmetic.rs:14
template <class P>
void to_repr(const F<P>& a, uint8_t out[32]) {
  F<P> o = {{1, 0, 0, 0}};
  F<P> c = fmul(a, o);
  memcpy(out, c.l, 32);
}

typedef F<FrP> Fr;
typedef F<FqP> Fq;

// ---------------------------------------------------------------------------
// G1: y^2 = x^3 + 3, Jacobian
// ---------------------------------------------------------------------------
struct Aff {
  Fq x, y;
  bool is_identity() const { return x.is_zero() && y.is_zero(); }
};
struct Jac {
  Fq x, y, z;
  bool is_identity() const { return z.is_zero(); }
  static Jac identity() { return Jac{Fq::zero(), Fq::one(), Fq::zero()}; }
};

Jac jdouble(const Jac& p) {  // dbl-2009-l
  if (p.is_identity()) return p;
  Fq a = fsqr(p.x), b = fsqr(p.y), c = fsqr(b);
  Fq d = fsub(fsub(fsqr(fadd(p.x, b)), a), c);
  d = fdbl(d);
  Fq e = fadd(fdbl(a), a), f = fsqr(e);
  Jac r;
  r.z = fdbl(fmul(p.y, p.z));
  r.x = fsub(f, fdbl(d));
  Fq c8 = fdbl(fdbl(fdbl(c)));
  r.y = fsub(fmul(e, fsub(d, r.x)), c8);
  return r;
}

Jac jadd(const Jac& p, const Jac& q) {  // add-2007-bl
  if (p.is_identity()) return q;
  if (q.is_identity()) return p;
  Fq z1z1 = fsqr(p.z), z2z2 = fsqr(q.z);
  Fq u1 = fmul(p.x, z2z2), u2 = fmul(q.x, z1z1);
  Fq s1 = fmul(fmul(p.y, q.z), z2z2), s2 = fmul(fmul(q.y, p.z), z1z1);
  if (u1 == u2) {
    if (s1 == s2) return jdouble(p);
    return Jac::identity();
  }
  Fq h = fsub(u2, u1);
  Fq i = fsqr(fdbl(h));
  Fq j = fmul(h, i);
  Fq r = fdbl(fsub(s2, s1));
  Fq v = fmul(u1, i);
  Jac o;
  o.x = fsub(fsub(fsqr(r), j), fdbl(v));
  o.y = fsub(fmul(r, fsub(v, o.x)), fdbl(fmul(s1, j)));
  o.z = fmul(fsub(fsub(fsqr(fadd(p.z, q.z)), z1z1), z2z2), h);
  return o;
}

Jac jadd_mixed(const Jac& p, const Aff& q) {  // madd-2007-bl
  if (q.is_identity()) return p;
  if (p.is_identity()) return Jac{q.x, q.y, Fq::one()};
  Fq z1z1 = fsqr(p.z);
  Fq u2 = fmul(q.x, z1z1);
  Fq s2 = fmul(fmul(q.y, p.z), z1z1);
  if (p.x == u2) {
    if (p.y == s2) return jdouble(p);
    return Jac::identity();
  }
  Fq h = fsub(u2, p.x);
  Fq hh = fsqr(h);
  Fq i = fdbl(fdbl(hh));
  Fq j = fmul(h, i);
  Fq r = fdbl(fsub(s2, p.y));
  Fq v = fmul(p.x, i);
  Jac o;
  o.x = fsub(fsub(fsqr(r), j), fdbl(v));
  o.y = fsub(fmul(r, fsub(v, o.x)), fdbl(fmul(p.y, j)));
  o.z = fsub(fsub(fsqr(fadd(p.z, h)), z1z1), hh);
  return o;
}

Aff to_affine(const Jac& p) {
  if (p.is_identity()) return Aff{Fq::zero(), Fq::zero()};
  Fq zi = finv(p.z), zi2 = fsqr(zi);
  return Aff{fmul(p.x, zi2), fmul(p.y, fmul(zi2, zi))};
}

// ---------------------------------------------------------------------------
// threads (the role rayon plays in the reference, multicore.rs:5)
// ---------------------------------------------------------------------------
int clamp_threads(int t) {
  if (t <= 0) t = (int)std::thread::hardware_concurrency();
  return t < 1 ? 1 : t;
}

// parallelize(v, f): chunked scope                                arithmetic.rs:371-388
template <class T, class Fn>
void parallelize(T* v, size_t n, int threads, Fn f) {
  size_t chunk = n / (size_t)threads;
  if (chunk < (size_t)threads) {
    // arithmetic.rs:375-377 sets chunk = 1 (one rayon task per element on the
    // pool); OS threads per element would be absurd, so run those inline.
    if (n) f(v, n, 0);
    return;
  }
  std::vector<std::thread> pool;
  for (size_t start = 0; start < n; start += chunk) {
    const size_t len = std::min(chunk, n - start);
    pool.emplace_back([=]() { f(v + start, len, start); });
  }
  for (auto& t : pool) t.join();
}

// ---------------------------------------------------------------------------
// MSM                                                             arithmetic.rs:13-159
// ---------------------------------------------------------------------------
inline size_t get_at(size_t segment, size_t c, const uint8_t* bytes) {  // :24-42
  const size_t skip_bits = segment * c;
  const size_t skip_bytes = skip_bits / 8;
  if (skip_bytes >= 32) return 0;
  uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = bytes[skip_bytes + i];
  uint64_t tmp;
  memcpy(&tmp, v, 8);
  tmp >>= skip_bits - skip_bytes * 8;
  tmp %= (1ull << c);
  return (size_t)tmp;
}

struct Bucket {  // :51-82
  int kind = 0;  // 0 None, 1 Affine, 2 Projective
  Aff a;
  Jac p;
  void add_assign(const Aff& other) {
    if (kind == 0) {
      kind = 1;
      a = other;
    } else if (kind == 1) {
      p = jadd_mixed(Jac{a.x, a.y, a.is_identity() ? Fq::zero() : Fq::one()}, other);
      kind = 2;
    } else {
      p = jadd_mixed(p, other);
    }
  }
  Jac add(const Jac& other) const {
    if (kind == 0) return other;
    if (kind == 1) return jadd_mixed(other, a);
    return jadd(other, p);
  }
};

void multiexp_serial(const Fr* coeffs, const Aff* bases, size_t n, Jac* acc) {
  std::vector<uint8_t> repr(n * 32);
  for (size_t i = 0; i < n; ++i) to_repr(coeffs[i], &repr[i * 32]);  // :14
  size_t c;
  if (n < 4)
    c = 1;
  else if (n < 32)
    c = 3;
  else
    c = (size_t)std::ceil(std::log((double)(uint32_t)n));  // :16-22
  const size_t segments = 256 / c + 1;                    // :44
  std::vector<Bucket> buckets;
  for (size_t seg = segments; seg-- > 0;) {
    for (size_t i = 0; i < c; ++i) *acc = jdouble(*acc);  // :47-49
    buckets.assign(((size_t)1 << c) - 1, Bucket());
    for (size_t i = 0; i < n; ++i) {  // :84-89
      const size_t d = get_at(seg, c, &repr[i * 32]);
      if (d != 0) buckets[d - 1].add_assign(bases[i]);
    }
    Jac running = Jac::identity();  // :95-99
    for (size_t b = buckets.size(); b-- > 0;) {
      running = buckets[b].add(running);
      *acc = jadd(*acc, running);
    }
  }
}

Jac best_multiexp(const Fr* coeffs, const Aff* bases, size_t n, int threads) {  // :132-159
  threads = clamp_threads(threads);
  if (n > (size_t)threads) {
    const size_t chunk = n / (size_t)threads;
    const size_t num_chunks = (n + chunk - 1) / chunk;
    std::vector<Jac> results(num_chunks, Jac::identity());
    std::vector<std::thread> pool;
    for (size_t ci = 0; ci < num_chunks; ++ci) {
      const size_t start = ci * chunk, len = std::min(chunk, n - start);
      pool.emplace_back([=, &results]() {
        multiexp_serial(coeffs + start, bases + start, len, &results[ci]);
      });
    }
    for (auto& t : pool) t.join();
    Jac acc = Jac::identity();
    for (auto& r : results) acc = jadd(acc, r);  // :153
    return acc;
  }
  Jac acc = Jac::identity();
  multiexp_serial(coeffs, bases, n, &acc);
  return acc;
}

// ---------------------------------------------------------------------------
// FFT                                                             arithmetic.rs:171-274
// ---------------------------------------------------------------------------
inline size_t bitreverse(size_t n, size_t l) {  // :172-179
  size_t r = 0;
  for (size_t i = 0; i < l; ++i) {
    r = (r << 1) | (n & 1);
    n >>= 1;
  }
  return r;
}

inline uint32_t log2_floor(size_t x) {
  uint32_t l = 0;
  while ((x >> (l + 1)) != 0) ++l;
  return l;
}

// depth = how many more levels may fork a thread (rayon::join :250-253)
void recursive_butterfly(Fr* a, size_t n, size_t twiddle_chunk, const Fr* tw, int depth) {
  if (n == 2) {  // :243-247
    const Fr t = a[1];
    a[1] = fsub(a[0], t);
    a[0] = fadd(a[0], t);
    return;
  }
  Fr* left = a;
  Fr* right = a + n / 2;
  if (depth > 0) {
    std::thread th([=]() { recursive_butterfly(left, n / 2, twiddle_chunk * 2, tw, depth - 1); });
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, depth - 1);
    th.join();
  } else {
    recursive_butterfly(left, n / 2, twiddle_chunk * 2, tw, 0);
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, 0);
  }
  {  // twiddle factor one, :256-262
    const Fr t = right[0];
    right[0] = fsub(left[0], t);
    left[0] = fadd(left[0], t);
  }
  for (size_t i = 1; i < n / 2; ++i) {  // :264-272
    const Fr t = fmul(right[i], tw[i * twiddle_chunk]);
    right[i] = fsub(left[i], t);
    left[i] = fadd(left[i], t);
  }
}

void best_fft(Fr* a, const Fr& omega, uint32_t log_n, int threads) {
  threads = clamp_threads(threads);
  const uint32_t log_threads = log2_floor((size_t)threads);
  const size_t n = (size_t)1 << log_n;
  for (size_t k = 0; k < n; ++k) {  // :186-191, serial
    const size_t rk = bitreverse(k, log_n);
    if (k < rk) std::swap(a[rk], a[k]);
  }
  std::vector<Fr> tw(n / 2 ? n / 2 : 1);  // :194-200, serial scan
  {
    Fr w = Fr::one();
    for (size_t i = 0; i < n / 2; ++i) {
      tw[i] = w;
      w = fmul(w, omega);
    }
  }
  if (log_n <= log_threads) {  // :202-230
    size_t chunk = 2, twiddle_chunk = n / 2;
    for (uint32_t s = 0; s < log_n; ++s) {
      for (size_t base = 0; base < n; base += chunk) {
        Fr* left = a + base;
        Fr* right = a + base + chunk / 2;
        const Fr t0 = right[0];
        right[0] = fsub(left[0], t0);
        left[0] = fadd(left[0], t0);
        for (size_t i = 1; i < chunk / 2; ++i) {
          const Fr t = fmul(right[i], tw[i * twiddle_chunk]);
          right[i] = fsub(left[i], t);
          left[i] = fadd(left[i], t);
        }
      }
      chunk *= 2;
      twiddle_chunk /= 2;
    }
  } else {
    recursive_butterfly(a, n, 1, tw.data(), (int)log_threads);  // :232
  }
}

// ---------------------------------------------------------------------------
// EvaluationDomain                                                poly/domain.rs
// ---------------------------------------------------------------------------
struct Domain {
  uint32_t j, k, extended_k;
  uint64_t n, quotient_poly_degree;
  Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor,
      extended_ifft_divisor;
  std::vector<Fr> t_evaluations;
  int threads;
};

const uint64_t kRootOfUnity[4] = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull,
                                  0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull};
const uint64_t kZeta[4] = {0x8b17ea66b99c90ddull, 0x5bfc41088d8daaa7ull, 0xb3c4d79d41a91758ull, 0};
const uint32_t kS = 28;

Domain* domain_new(uint32_t j, uint32_t k, int threads) {  // domain.rs:39-142
  if (j < 1 || k > kS) return nullptr;  // j = 1: quotient degree 0, extended_k = k (the reference's own tests, domain.rs:494)
  Domain* d = new Domain();
  d->threads = clamp_threads(threads);
  d->j = j;
  d->k = k;
  d->quotient_poly_degree = j - 1;
  d->n = 1ull << k;
  uint32_t ek = k;
  while ((1ull << ek) < d->n * d->quotient_poly_degree) ++ek;  // :49-52
  if (ek > kS) {
    delete d;
    return nullptr;
  }
  d->extended_k = ek;
  Fr eo = from_canonical<FrP>(kRootOfUnity);
  for (uint32_t i = ek; i < kS; ++i) eo = fsqr(eo);  // :58-60
  d->extended_omega = eo;
  Fr o = eo;
  for (uint32_t i = k; i < ek; ++i) o = fsqr(o);  // :70-73
  d->omega = o;
  d->g_coset = from_canonical<FrP>(kZeta);  // :81
  d->g_coset_inv = fsqr(d->g_coset);        // :82
  {                                          // :84-107
    const uint64_t e[4] = {d->n, 0, 0, 0};
    const Fr orig = fpow(d->g_coset, e), step = fpow(eo, e);
    Fr cur = orig;
    do {
      d->t_evaluations.push_back(cur);
      cur = fmul(cur, step);
    } while (cur != orig);
    for (auto& t : d->t_evaluations) t = finv(fsub(t, Fr::one()));
  }
  d->ifft_divisor = finv(from_u64<FrP>(1ull << k));
  d->extended_ifft_divisor = finv(from_u64<FrP>(1ull << ek));
  d->extended_omega_inv = finv(eo);
  d->omega_inv = finv(o);
  return d;
}

void distribute_powers_zeta(const Domain* d, Fr* a, size_t n, bool into_coset) {  // :335-351
  const Fr p0 = into_coset ? d->g_coset : d->g_coset_inv;
  const Fr p1 = into_coset ? d->g_coset_inv : d->g_coset;
  parallelize(a, n, d->threads, [=](Fr* v, size_t len, size_t index) {
    for (size_t i = 0; i < len; ++i, ++index) {
      const size_t m = index % 3;
      if (m == 1) v[i] = fmul(v[i], p0);
      if (m == 2) v[i] = fmul(v[i], p1);
    }
  });
}

void ifft(const Domain* d, Fr* a, const Fr& omega_inv, uint32_t log_n, const Fr& divisor) {  // :353-361
  best_fft(a, omega_inv, log_n, d->threads);
  parallelize(a, (size_t)1 << log_n, d->threads, [=](Fr* v, size_t len, size_t) {
    for (size_t i = 0; i < len; ++i) v[i] = fmul(v[i], divisor);
  });
}

}  // namespace

// ===========================================================================
// C interface (ctypes from tests/ and bench.py's cpu_baseline leg only)
// Fr / Fq: 4 LE u64 Montgomery limbs; affine: x,y; Jacobian: x,y,z.
// ===========================================================================
extern "C" {

int oracle_best_multiexp(const uint64_t* coeffs, const uint64_t* bases, size_t n, int threads,
                         uint64_t* out_affine) {
  const Jac r = best_multiexp(reinterpret_cast<const Fr*>(coeffs), reinterpret_cast<const Aff*>(bases),
                              n, threads);
  const Aff a = to_affine(r);
  memcpy(out_affine, &a, 64);
  return 0;
}

int oracle_best_fft(uint64_t* a, const uint64_t* omega, uint32_t log_n, int threads) {
  if (log_n > 28) return -1;
  Fr w;
  memcpy(&w, omega, 32);
  best_fft(reinterpret_cast<Fr*>(a), w, log_n, threads);
  return 0;
}

void* oracle_domain_new(uint32_t j, uint32_t k, int threads) { return domain_new(j, k, threads); }
void oracle_domain_free(void* d) { delete static_cast<Domain*>(d); }
uint32_t oracle_domain_extended_k(const void* d) { return static_cast<const Domain*>(d)->extended_k; }
size_t oracle_domain_quotient_len(const void* d) {
  const Domain* D = static_cast<const Domain*>(d);
  return (size_t)(D->n * D->quotient_poly_degree);
}
// which: as h2b_domain_constant (include/halo2_b200.h)
int oracle_domain_constant(const void* d, uint32_t which, uint64_t* out) {
  const Domain* D = static_cast<const Domain*>(d);
  const Fr* src = nullptr;
  switch (which) {
    case 0: src = &D->omega; break;
    case 1: src = &D->omega_inv; break;
    case 2: src = &D->extended_omega; break;
    case 3: src = &D->extended_omega_inv; break;
    case 4: src = &D->g_coset; break;
    case 5: src = &D->g_coset_inv; break;
    case 6: src = &D->ifft_divisor; break;
    case 7: src = &D->extended_ifft_divisor; break;
    default:
      if (which - 8 < D->t_evaluations.size()) src = &D->t_evaluations[which - 8];
  }
  if (!src) return -1;
  memcpy(out, src, 32);
  return 0;
}

// lagrange_to_coeff: a (2^k) in place                              domain.rs:226-236
int oracle_lagrange_to_coeff(const void* d, uint64_t* a) {
  const Domain* D = static_cast<const Domain*>(d);
  ifft(D, reinterpret_cast<Fr*>(a), D->omega_inv, D->k, D->ifft_divisor);
  return 0;
}

// coeff_to_extended: in (2^k) -> out (2^extended_k)                 domain.rs:240-254
int oracle_coeff_to_extended(const void* d, const uint64_t* in, uint64_t* out) {
  const Domain* D = static_cast<const Domain*>(d);
  Fr* o = reinterpret_cast<Fr*>(out);
  memcpy(o, in, (size_t)D->n * 32);
  distribute_powers_zeta(D, o, (size_t)D->n, true);
  memset(o + D->n, 0, (((size_t)1 << D->extended_k) - (size_t)D->n) * 32);  // resize(.., zero) :247
  best_fft(o, D->extended_omega, D->extended_k, D->threads);
  return 0;
}

// divide_by_vanishing_poly: a (2^extended_k) in place               domain.rs:307-326
int oracle_divide_by_vanishing_poly(const void* d, uint64_t* a) {
  const Domain* D = static_cast<const Domain*>(d);
  const Fr* t = D->t_evaluations.data();
  const size_t tl = D->t_evaluations.size();
  parallelize(reinterpret_cast<Fr*>(a), (size_t)1 << D->extended_k, D->threads,
              [=](Fr* v, size_t len, size_t index) {
                for (size_t i = 0; i < len; ++i, ++index) v[i] = fmul(v[i], t[index % tl]);
              });
  return 0;
}

// extended_to_coeff: a (2^extended_k), modified in place; the first
// quotient_len elements are the result                              domain.rs:281-303
int oracle_extended_to_coeff(const void* d, uint64_t* a) {
  const Domain* D = static_cast<const Domain*>(d);
  Fr* v = reinterpret_cast<Fr*>(a);
  ifft(D, v, D->extended_omega_inv, D->extended_k, D->extended_ifft_divisor);
  distribute_powers_zeta(D, v, (size_t)1 << D->extended_k, false);
  return 0;
}

// element-wise field ops for pinning the limb arithmetic against big integers
// field 0 Fr, 1 Fq; op 0 mul, 1 add, 2 sub, 7 inv
int oracle_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    if (field == 0) {
      Fr x, y, r;
      memcpy(&x, a + 4 * i, 32);
      memcpy(&y, b + 4 * i, 32);
      r = op == 0 ? fmul(x, y) : op == 1 ? fadd(x, y) : op == 2 ? fsub(x, y) : finv(x);
      memcpy(out + 4 * i, &r, 32);
    } else {
      Fq x, y, r;
      memcpy(&x, a + 4 * i, 32);
      memcpy(&y, b + 4 * i, 32);
      r = op == 0 ? fmul(x, y) : op == 1 ? fadd(x, y) : op == 2 ? fsub(x, y) : finv(x);
      memcpy(out + 4 * i, &r, 32);
    }
  }
  return 0;
}

// out[i] = [k_i] G for 64-bit k_i (test base generation; G = (1, 2))
int oracle_g1_mul_gen_u64(const uint64_t* ks, size_t n, int threads, uint64_t* out_affine) {
  threads = clamp_threads(threads);
  Aff g{Fq::one(), fadd(Fq::one(), Fq::one())};
  Aff* out = reinterpret_cast<Aff*>(out_affine);
  parallelize(out, n, threads, [=](Aff* v, size_t len, size_t index) {
    for (size_t i = 0; i < len; ++i, ++index) {
      Jac acc = Jac::identity();
      for (int b = 63; b >= 0; --b) {
        acc = jdouble(acc);
        if ((ks[index] >> b) & 1) acc = jadd_mixed(acc, g);
      }
      v[i] = to_affine(acc);
    }
  });
  return 0;
}

// eval_polynomial                                                   arithmetic.rs:304-329
int oracle_eval_polynomial(const uint64_t* poly, size_t n, const uint64_t* point, int threads, uint64_t* out) {
  threads = clamp_threads(threads);
  const Fr* p = reinterpret_cast<const Fr*>(poly);
  Fr x;
  memcpy(&x, point, 32);
  auto evaluate = [&](const Fr* c, size_t len) {
    Fr acc = Fr::zero();
    for (size_t i = len; i-- > 0;) acc = fadd(fmul(acc, x), c[i]);
    return acc;
  };
  Fr res = Fr::zero();
  if (n * 2 < (size_t)threads) {
    res = evaluate(p, n);
  } else {
    const size_t chunk = (n + threads - 1) / threads;
    std::vector<Fr> parts(threads, Fr::zero());
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
      const size_t start = (size_t)t * chunk;
      if (start >= n) break;
      const size_t len = std::min(chunk, n - start);
      pool.emplace_back([&, t, start, len]() {
        const uint64_t e[4] = {start, 0, 0, 0};
        parts[t] = fmul(evaluate(p + start, len), fpow(x, e));
      });
    }
    for (auto& th : pool) th.join();
    for (auto& v : parts) res = fadd(res, v);
  }
  memcpy(out, &res, 32);
  return 0;
}

// kate_division (serial, as in the reference)                        arithmetic.rs:348-367
int oracle_kate_division(const uint64_t* a, size_t n, const uint64_t* b, uint64_t* q) {
  if (n == 0) return -1;
  const Fr* A = reinterpret_cast<const Fr*>(a);
  Fr* Q = reinterpret_cast<Fr*>(q);
  Fr nb;
  memcpy(&nb, b, 32);
  nb = fneg(nb);
  Fr tmp = Fr::zero();
  for (size_t i = n - 1; i-- > 0;) {
    const Fr lead = fsub(A[i + 1], tmp);
    Q[i] = lead;
    tmp = fmul(lead, nb);
  }
  return 0;
}

// compute_inner_product (serial, as in the reference)                arithmetic.rs:331-345
int oracle_inner_product(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  const Fr* A = reinterpret_cast<const Fr*>(a);
  const Fr* B = reinterpret_cast<const Fr*>(b);
  Fr acc = Fr::zero();
  for (size_t i = 0; i < n; ++i) acc = fadd(acc, fmul(A[i], B[i]));
  memcpy(out, &acc, 32);
  return 0;
}

// Synthetic MSM bases for the CPU baseline: out[i] = [a + i*d] G (valid, distinct
// points; an MSM's cost does not depend on the base values).  Each thread walks
// its range with mixed additions and normalises with one batch inversion.
int oracle_synth_bases(uint64_t a, uint64_t d, size_t n, int threads, uint64_t* out_affine) {
  threads = clamp_threads(threads);
  const Aff g{Fq::one(), fadd(Fq::one(), Fq::one())};
  auto mul_u64 = [&](uint64_t k) {
    Jac acc = Jac::identity();
    for (int b = 63; b >= 0; --b) {
      acc = jdouble(acc);
      if ((k >> b) & 1) acc = jadd_mixed(acc, g);
    }
    return acc;
  };
  const Aff D = to_affine(mul_u64(d));
  Aff* out = reinterpret_cast<Aff*>(out_affine);
  parallelize(out, n, threads, [=](Aff* v, size_t len, size_t index) {
    std::vector<Jac> pts(len);
    Jac cur = mul_u64(a + d * index);
    for (size_t i = 0; i < len; ++i) {
      pts[i] = cur;
      cur = jadd_mixed(cur, D);
    }
    // batch inversion of the z coordinates (none is zero: a + i*d < r for the sizes used)
    std::vector<Fq> pre(len);
    Fq run = Fq::one();
    for (size_t i = 0; i < len; ++i) {
      pre[i] = run;
      run = fmul(run, pts[i].z);
    }
    Fq inv = finv(run);
    for (size_t i = len; i-- > 0;) {
      const Fq zi = fmul(inv, pre[i]);
      inv = fmul(inv, pts[i].z);
      const Fq zi2 = fsqr(zi);
      v[i] = Aff{fmul(pts[i].x, zi2), fmul(pts[i].y, fmul(zi2, zi))};
    }
  });
  return 0;
}


}  // extern "C"

namespace {
// ---------------------------------------------------------------------------
// Evaluator::evaluate_h                                   plonk/evaluation.rs:280-522
// The calculation stream is the reference's enums flattened in declaration order (the same words the
// product's h2b_graph_new takes): op, target, operands...; an operand is (kind, a, b).
// ---------------------------------------------------------------------------
struct EvalGraph {
  struct Calc {
    uint32_t op, target;
    std::vector<std::array<uint32_t, 3>> src;  // Horner: [start, factor, parts...]
  };
  std::vector<Calc> calcs;
  std::vector<Fr> constants;
  std::vector<int32_t> rotations;
  uint32_t num_intermediates;
};

struct EvalCols {
  const Fr* const* fixed;
  const Fr* const* advice;
  const Fr* const* instance;
  const Fr* challenges;
  Fr beta, gamma, theta, y;
};

static bool parse_graph(const uint32_t* w, size_t n_words, const uint64_t* constants, uint32_t n_constants,
                        const int32_t* rotations, uint32_t n_rotations, uint32_t n_inter, EvalGraph* g) {
  size_t pos = 0;
  auto rd3 = [&](std::array<uint32_t, 3>& s) {
    if (pos + 3 > n_words) return false;
    s = {w[pos], w[pos + 1], w[pos + 2]};
    pos += 3;
    return true;
  };
  while (pos < n_words) {
    if (pos + 2 > n_words) return false;
    EvalGraph::Calc c;
    c.op = w[pos], c.target = w[pos + 1];
    pos += 2;
    int nsrc = (c.op <= 2) ? 2 : 1;  // Add, Sub, Mul take two operands; Square, Double, Negate, Store one
    if (c.op == 6) {                 // Horner(start, factor, nparts, parts...)
      std::array<uint32_t, 3> s;
      if (!rd3(s)) return false;
      c.src.push_back(s);
      if (!rd3(s)) return false;
      c.src.push_back(s);
      if (pos >= n_words) return false;
      nsrc = (int)w[pos++];
    }
    for (int i = 0; i < nsrc; ++i) {
      std::array<uint32_t, 3> s;
      if (!rd3(s)) return false;
      c.src.push_back(s);
    }
    g->calcs.push_back(c);
  }
  for (uint32_t i = 0; i < n_constants; ++i)
    g->constants.push_back(Fr{{constants[4 * i], constants[4 * i + 1], constants[4 * i + 2], constants[4 * i + 3]}});
  g->rotations.assign(rotations, rotations + n_rotations);
  g->num_intermediates = n_inter;
  return true;
}

// get_rotation_idx                                                    evaluation.rs:32-34
static inline size_t rot_idx(size_t idx, int32_t rot, int32_t rot_scale, int64_t isize) {
  int64_t v = ((int64_t)idx + (int64_t)rot * rot_scale) % isize;
  return (size_t)(v < 0 ? v + isize : v);
}

// GraphEvaluator::evaluate                                            evaluation.rs:700-746
static Fr graph_evaluate(const EvalGraph& g, std::vector<Fr>& inter, std::vector<size_t>& rots, const EvalCols& c,
                         const Fr& previous, size_t idx, int32_t rot_scale, int64_t isize) {
  for (size_t r = 0; r < g.rotations.size(); ++r) rots[r] = rot_idx(idx, g.rotations[r], rot_scale, isize);
  auto get = [&](const std::array<uint32_t, 3>& s) -> Fr {  // ValueSource::get, evaluation.rs:69-105
    switch (s[0]) {
      case 0: return g.constants[s[1]];
      case 1: return inter[s[1]];
      case 2: return c.fixed[s[1]][rots[s[2]]];
      case 3: return c.advice[s[1]][rots[s[2]]];
      case 4: return c.instance[s[1]][rots[s[2]]];
      case 5: return c.challenges[s[1]];
      case 6: return c.beta;
      case 7: return c.gamma;
      case 8: return c.theta;
      case 9: return c.y;
      default: return previous;
    }
  };
  for (const auto& k : g.calcs) {  // Calculation::evaluate, evaluation.rs:131-179
    Fr v;
    switch (k.op) {
      case 0: v = fadd(get(k.src[0]), get(k.src[1])); break;
      case 1: v = fsub(get(k.src[0]), get(k.src[1])); break;
      case 2: v = fmul(get(k.src[0]), get(k.src[1])); break;
      case 3: { Fr a = get(k.src[0]); v = fsqr(a); break; }
      case 4: v = fdbl(get(k.src[0])); break;
      case 5: v = fneg(get(k.src[0])); break;
      case 6: {
        const Fr factor = get(k.src[1]);
        v = get(k.src[0]);
        for (size_t i = 2; i < k.src.size(); ++i) v = fadd(fmul(v, factor), get(k.src[i]));
        break;
      }
      default: v = get(k.src[0]); break;
    }
    inter[k.target] = v;
  }
  return g.calcs.empty() ? Fr::zero() : inter[g.calcs.back().target];
}

// chunks_mut(chunk_size) with chunk_size = ceil(size / num_threads)     evaluation.rs:336-362
template <class Fn>
static void chunked_scope(size_t size, int threads, Fn f) {
  const size_t chunk = (size + (size_t)threads - 1) / (size_t)threads;
  std::vector<std::thread> pool;
  for (size_t start = 0; start < size; start += chunk) {
    const size_t len = std::min(chunk, size - start);
    pool.emplace_back([=]() { f(start, len); });
  }
  for (auto& t : pool) t.join();
}

}  // namespace

extern "C" {
void* oracle_graph_new(const uint32_t* words, size_t n_words, const uint64_t* constants, uint32_t n_constants,
                       const int32_t* rotations, uint32_t n_rotations, uint32_t num_intermediates) {
  EvalGraph* g = new EvalGraph();
  if (!parse_graph(words, n_words, constants, n_constants, rotations, n_rotations, num_intermediates, g)) {
    delete g;
    return nullptr;
  }
  return g;
}
void oracle_graph_free(void* g) { delete reinterpret_cast<EvalGraph*>(g); }

// One circuit instance's contribution to h over the extended domain, in place on `values` (host arrays):
// custom gates, then the permutation argument (n_sets > 0), then every lookup.
//   scalars: beta, gamma, theta, y (4 x 4 u64);  column pointers: arrays of pointers to 2^extended_k elements
//   lookups: per lookup a graph handle and 3 cosets (product, permuted_input, permuted_table)
int oracle_evaluate_h(void* domain, void* gates_graph, const uint64_t* const* fixed, const uint64_t* const* advice,
                      const uint64_t* const* instance, const uint64_t* challenges, const uint64_t* scalars,
                      const uint32_t* perm_col_type, const uint32_t* perm_col_index, uint32_t n_perm_cols,
                      const uint64_t* const* sigma_cosets, const uint64_t* const* z_cosets, uint32_t n_sets,
                      uint32_t chunk_len, uint32_t blinding_factors, const uint64_t* l0_, const uint64_t* l_last_,
                      const uint64_t* l_active_, void* const* lookup_graphs, const uint64_t* const* lookup_cosets,
                      uint32_t n_lookups, uint64_t* values_, int threads) {
  const Domain* d = reinterpret_cast<const Domain*>(domain);
  threads = clamp_threads(threads);
  const size_t size = (size_t)1 << d->extended_k;
  const int32_t rot_scale = 1 << (d->extended_k - d->k);
  const int64_t isize = (int64_t)size;
  auto F4 = [](const uint64_t* p) { return Fr{{p[0], p[1], p[2], p[3]}}; };
  EvalCols c;
  c.fixed = reinterpret_cast<const Fr* const*>(fixed);
  c.advice = reinterpret_cast<const Fr* const*>(advice);
  c.instance = reinterpret_cast<const Fr* const*>(instance);
  c.challenges = reinterpret_cast<const Fr*>(challenges);
  c.beta = F4(scalars), c.gamma = F4(scalars + 4), c.theta = F4(scalars + 8), c.y = F4(scalars + 12);
  Fr* values = reinterpret_cast<Fr*>(values_);
  const Fr* l0 = reinterpret_cast<const Fr*>(l0_);
  const Fr* l_last = reinterpret_cast<const Fr*>(l_last_);
  const Fr* l_active = reinterpret_cast<const Fr*>(l_active_);
  const Fr one = Fr::one();
  const Fr y = c.y, beta = c.beta, gamma = c.gamma;

  // custom gates (:335-362)
  const EvalGraph& gg = *reinterpret_cast<const EvalGraph*>(gates_graph);
  chunked_scope(size, threads, [&](size_t start, size_t len) {
    std::vector<Fr> inter(gg.num_intermediates, Fr::zero());
    std::vector<size_t> rots(gg.rotations.size(), 0);
    for (size_t i = 0; i < len; ++i) {
      const size_t idx = start + i;
      values[idx] = graph_evaluate(gg, inter, rots, c, values[idx], idx, rot_scale, isize);
    }
  });

  // permutations (:364-444)
  if (n_sets) {
    static const uint64_t kZeta[4] = {0x8b17ea66b99c90ddull, 0x5bfc41088d8daaa7ull, 0xb3c4d79d41a91758ull, 0x0ull};
    static const uint64_t kDelta[4] = {0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull, 0x64ec26aad4c86e71ull,
                                       0x09226b6e22c6f0caull};
    const Fr delta_start = fmul(beta, from_canonical<FrP>(kZeta));
    const Fr DELTA = from_canonical<FrP>(kDelta);
    const int32_t last_rotation = -(int32_t)(blinding_factors + 1);
    auto column = [&](uint32_t j) -> const Fr* {
      switch (perm_col_type[j]) {
        case 0: return c.advice[perm_col_index[j]];
        case 1: return c.fixed[perm_col_index[j]];
        default: return c.instance[perm_col_index[j]];
      }
    };
    const Fr* const* z = reinterpret_cast<const Fr* const*>(z_cosets);
    const Fr* const* sg = reinterpret_cast<const Fr* const*>(sigma_cosets);
    parallelize(values, size, threads, [&](Fr* vals, size_t len, size_t start) {
      const uint64_t e[4] = {(uint64_t)start, 0, 0, 0};
      Fr beta_term = fpow(d->extended_omega, e);
      for (size_t i = 0; i < len; ++i) {
        const size_t idx = start + i;
        const size_t r_next = rot_idx(idx, 1, rot_scale, isize);
        const size_t r_last = rot_idx(idx, last_rotation, rot_scale, isize);
        Fr v = vals[i];
        v = fadd(fmul(v, y), fmul(fsub(one, z[0][idx]), l0[idx]));
        const Fr zl = z[n_sets - 1][idx];
        v = fadd(fmul(v, y), fmul(fsub(fmul(zl, zl), zl), l_last[idx]));
        for (uint32_t s2 = 1; s2 < n_sets; ++s2)
          v = fadd(fmul(v, y), fmul(fsub(z[s2][idx], z[s2 - 1][r_last]), l0[idx]));
        Fr current_delta = fmul(delta_start, beta_term);
        for (uint32_t s2 = 0; s2 < n_sets; ++s2) {
          const uint32_t c0 = s2 * chunk_len, c1 = std::min(c0 + chunk_len, n_perm_cols);
          Fr left = z[s2][r_next];
          for (uint32_t j = c0; j < c1; ++j)
            left = fmul(left, fadd(fadd(column(j)[idx], fmul(beta, sg[j][idx])), gamma));
          Fr right = z[s2][idx];
          for (uint32_t j = c0; j < c1; ++j) {
            right = fmul(right, fadd(fadd(column(j)[idx], current_delta), gamma));
            current_delta = fmul(current_delta, DELTA);
          }
          v = fadd(fmul(v, y), fmul(fsub(left, right), l_active[idx]));
        }
        vals[i] = v;
        beta_term = fmul(beta_term, d->extended_omega);
      }
    });
  }

  // lookups (:446-519)
  for (uint32_t n = 0; n < n_lookups; ++n) {
    const EvalGraph& lg = *reinterpret_cast<const EvalGraph*>(lookup_graphs[n]);
    const Fr* product = reinterpret_cast<const Fr*>(lookup_cosets[3 * n]);
    const Fr* pin = reinterpret_cast<const Fr*>(lookup_cosets[3 * n + 1]);
    const Fr* ptab = reinterpret_cast<const Fr*>(lookup_cosets[3 * n + 2]);
    parallelize(values, size, threads, [&](Fr* vals, size_t len, size_t start) {
      std::vector<Fr> inter(lg.num_intermediates, Fr::zero());
      std::vector<size_t> rots(lg.rotations.size(), 0);
      for (size_t i = 0; i < len; ++i) {
        const size_t idx = start + i;
        const Fr table_value = graph_evaluate(lg, inter, rots, c, Fr::zero(), idx, rot_scale, isize);
        const size_t r_next = rot_idx(idx, 1, rot_scale, isize), r_prev = rot_idx(idx, -1, rot_scale, isize);
        const Fr a_minus_s = fsub(pin[idx], ptab[idx]);
        Fr v = vals[i];
        v = fadd(fmul(v, y), fmul(fsub(one, product[idx]), l0[idx]));
        v = fadd(fmul(v, y), fmul(fsub(fmul(product[idx], product[idx]), product[idx]), l_last[idx]));
        v = fadd(fmul(v, y),
                 fmul(fsub(fmul(fmul(product[r_next], fadd(pin[idx], beta)), fadd(ptab[idx], gamma)),
                           fmul(product[idx], table_value)),
                      l_active[idx]));
        v = fadd(fmul(v, y), fmul(a_minus_s, l0[idx]));
        v = fadd(fmul(v, y), fmul(fmul(a_minus_s, fsub(pin[idx], pin[r_prev])), l_active[idx]));
        vals[i] = v;
      }
    });
  }
  return 0;
}

}  // extern "C"
