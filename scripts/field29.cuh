// Carry-free bn256 field arithmetic in 9 x 29-bit limbs for the hot loops.
//
// Why: on B200 an IMAD.WIDE that produces a carry predicate issues at half the
// rate of a plain IMAD.WIDE (30 vs 55-64 per clock per SM, measured with
// scripts/microbench2.cu), and the 8 x 32-bit Montgomery product of field.cuh
// is one long carry chain.  With 29-bit limbs every partial product is < 2^58,
// so 18 of them (9 of a*b, 9 of m*p) fit a 64-bit accumulator without any carry
// out: the whole product is plain `acc += (u64)a * b` (IMAD.WIDE.U32), carries
// are resolved by a handful of shifts on the otherwise idle ALU pipe.
//
// Representation: x = sum l[i] * 2^(29 i).  "Normalised" (N): every limb < 2^29.
// "Fat" (F): limbs may exceed 29 bits (results of lazy add / sub); a product
// accepts ONE fat operand with limbs <= 6 * 2^29.  Values are kept only loosely
// reduced (a few multiples of p); Montgomery radix is R9 = 2^261, so a product
// of values a, b returns a value < p * (1 + a*b / (R9 * p)), R9 / p ~ 169.
//
// The boundary format (field.cuh, halo2curves) is 8 x 32-bit limbs with
// R = 2^256; to_r9 / from_r9 convert (one multiplication by a constant).
// Linear maps (the NTT) need no conversion at all: only the twiddles are put in
// R9 form, since mont29(v, w * R9) = v * w whatever form v is in.
#pragma once
#include "../halo2-pse_b200/csrc/field.cuh"

namespace h2b {

static constexpr uint32_t MASK29 = (1u << 29) - 1u;

struct Fr29Params {
  typedef FrParams Base;
  static H2B_HD constexpr uint32_t mod(int i) {
    constexpr uint32_t t[9] = {0x10000001u, 0x1f0fac9fu, 0x0e5c2450u, 0x07d090f3u, 0x1585d283u,
                               0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu};
    return t[i];
  }
  static H2B_HD constexpr uint32_t one(int i) {  // 2^261 mod r
    constexpr uint32_t t[9] = {0x0fffff57u, 0x1ea70ab4u, 0x052c068bu, 0x17504f49u, 0x0aa8075bu,
                               0x1d4240ceu, 0x11d54c07u, 0x052ac7a8u, 0x000dc836u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t up(int i) {  // 2^266 mod r: R(2^256)-form -> R9-form
    constexpr uint32_t t[9] = {0x0fffead7u, 0x1d5444f4u, 0x04438aa5u, 0x03b4d096u, 0x134c84dau,
                               0x0e92d304u, 0x14cb95b3u, 0x041b9d3du, 0x00058003u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t down(int i) {  // 2^256 mod r: R9-form -> R-form
    constexpr uint32_t t[9] = {0x0ffffffbu, 0x04b1a0e2u, 0x18334a6bu, 0x18ed2b3eu, 0x1462e36fu,
                               0x11b7bc3cu, 0x1cbd99bau, 0x183340fbu, 0x000e0a77u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t pad16(int i) {  // 16 r, limbs 0..7 padded by 2^30
    constexpr uint32_t t[9] = {0x40000010u, 0x50fac9f6u, 0x45c2450du, 0x5d090f35u, 0x585d2831u,
                               0x4db40c08u, 0x4a6e140fu, 0x45c2633eu, 0x030644e5u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t pad8f(int i) {  // 8 r, limbs 0..7 padded by 2^31
    constexpr uint32_t t[9] = {0x80000008u, 0x987d64f8u, 0x92e12283u, 0x9e848797u, 0x8c2e9415u,
                               0x96da0601u, 0x85370a04u, 0x92e1319cu, 0x0183226fu};
    return t[i];
  }
  static constexpr uint32_t INV = 0x0fffffffu;   // -r^-1 mod 2^29
  static constexpr uint32_t PINV = 0x10000001u;  //  r^-1 mod 2^29
};

struct Fq29Params {
  typedef FqParams Base;
  static H2B_HD constexpr uint32_t mod(int i) {
    constexpr uint32_t t[9] = {0x187cfd47u, 0x010460b6u, 0x1c72a34fu, 0x02d522d0u, 0x1585d978u,
                               0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu};
    return t[i];
  }
  static H2B_HD constexpr uint32_t one(int i) {
    constexpr uint32_t t[9] = {0x157ccc21u, 0x141c2758u, 0x185230d3u, 0x014c0419u, 0x0aa36fb9u,
                               0x1d4240ceu, 0x11d54c07u, 0x052ac7a8u, 0x000dc836u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t up(int i) {
    constexpr uint32_t t[9] = {0x13349ca1u, 0x1a5d84a8u, 0x0a3e5cacu, 0x100249e0u, 0x12b951e8u,
                               0x0e92d304u, 0x14cb95b3u, 0x041b9d3du, 0x00058003u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t down(int i) {
    constexpr uint32_t t[9] = {0x058f0d9du, 0x1aea1c6eu, 0x11c2cf74u, 0x11d651ebu, 0x1462c0a7u,
                               0x11b7bc3cu, 0x1cbd99bau, 0x183340fbu, 0x000e0a77u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t pad16(int i) {
    constexpr uint32_t t[9] = {0x47cfd470u, 0x50460b6au, 0x472a34eeu, 0x4d522d0cu, 0x585d977fu,
                               0x4db40c08u, 0x4a6e140fu, 0x45c2633eu, 0x030644e5u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t pad8f(int i) {
    constexpr uint32_t t[9] = {0x83e7ea38u, 0x882305b2u, 0x83951a74u, 0x96a91683u, 0x8c2ecbbcu,
                               0x96da0601u, 0x85370a04u, 0x92e1319cu, 0x0183226fu};
    return t[i];
  }
  static constexpr uint32_t INV = 0x04866389u;
  static constexpr uint32_t PINV = 0x1b799c77u;
};

template <class P>
struct F29 {
  uint32_t l[9];
  static H2B_HD F29 zero() {
    F29 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.l[i] = 0;
    return r;
  }
  static H2B_HD F29 one() {
    F29 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.l[i] = P::one(i);
    return r;
  }
};

// 8 x 32 saturated words  ->  9 x 29 limbs of the same integer (N)
template <class P>
H2B_HD F29<P> unpack29(const Fp<typename P::Base>& a) {
  F29<P> r;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int bit = 29 * i, w = bit >> 5, sh = bit & 31;
    uint32_t v = a.v[w] >> sh;
    if (sh > 3 && w + 1 < 8) v |= a.v[w + 1] << (32 - sh);
    r.l[i] = v & MASK29;
  }
  return r;
}

// N limbs of a value < 2^256  ->  8 x 32 saturated words
template <class P>
H2B_HD Fp<typename P::Base> pack29(const F29<P>& a) {
  Fp<typename P::Base> r;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int bit = 32 * w, i = bit / 29, sh = bit - 29 * i;  // word w starts inside limb i
    uint32_t v = a.l[i] >> sh;
    v |= a.l[i + 1] << (29 - sh);
    if (29 - sh + 29 < 32 && i + 2 < 9) v |= a.l[i + 2] << (58 - sh);
    r.v[w] = v;
  }
  return r;
}

// carry propagation only (no modular reduction): F -> N, same value (< 2^261)
template <class P>
H2B_HD F29<P> norm29(const F29<P>& a) {
  F29<P> r;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t v = a.l[i] + c;  // limbs <= 2^32 - 2^4: no wrap (callers keep limbs < 7 * 2^29)
    r.l[i] = v & MASK29;
    c = v >> 29;
  }
  r.l[8] = a.l[8] + c;
  return r;
}

template <class P>
H2B_HD F29<P> add29(const F29<P>& a, const F29<P>& b) {  // lazy: limbs add, no carry
  F29<P> r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.l[i] = a.l[i] + b.l[i];
  return r;
}

// a - b + 16p, lazy.  Needs b's limbs <= 2^30 - 2 (limb 8: b's value < 16p); result limbs < a + 2^30 + 2^29.
template <class P>
H2B_HD F29<P> sub29(const F29<P>& a, const F29<P>& b) {
  F29<P> r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.l[i] = a.l[i] + P::pad16(i) - b.l[i];
  return r;
}

// a - b + 8p with 2^31 padding: b's limbs < 2^31 (value < 8p); result limbs < a + 2^31 + 2^29.
template <class P>
H2B_HD F29<P> sub29f(const F29<P>& a, const F29<P>& b) {
  F29<P> r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.l[i] = a.l[i] + P::pad8f(i) - b.l[i];
  return r;
}

// Montgomery product a * b / 2^261 mod p.  a: N.  b: N or fat with limbs <= 6 * 2^29.
// Result: N, value < p + a*b / 2^261.
template <class P>
H2B_HD F29<P> mul29(const F29<P>& a, const F29<P>& b) {
  uint64_t t[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) t[j] = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const uint32_t bi = b.l[i];
#pragma unroll
    for (int j = 0; j < 9; ++j) t[j] += (uint64_t)a.l[j] * bi;
    const uint32_t m = ((uint32_t)t[0] * P::INV) & MASK29;
#pragma unroll
    for (int j = 0; j < 9; ++j) t[j] += (uint64_t)m * P::mod(j);
    const uint64_t carry = t[0] >> 29;  // low 29 bits are zero by construction
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = t[j + 1];
    t[0] += carry;
    t[8] = 0;
  }
  F29<P> r;
  uint64_t c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint64_t v = t[j] + c;
    r.l[j] = (uint32_t)v & MASK29;
    c = v >> 29;
  }
  r.l[8] = (uint32_t)c;  // t[8] is always zero after the last shift
  return r;
}

template <class P>
H2B_HD F29<P> sqr29(const F29<P>& a) {
  return mul29(a, a);
}

// N value < 2p  ->  canonical (< p)
template <class P>
H2B_HD F29<P> cond_sub29(const F29<P>& a) {
  F29<P> r;
  uint32_t borrow = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const uint32_t v = a.l[i] - P::mod(i) - borrow;
    borrow = v >> 31;
    r.l[i] = v & MASK29;
  }
  if (borrow) return a;
  return r;
}

// any N value (< 2^257)  ->  canonical representative of the same residue
template <class P>
H2B_HD F29<P> canon29(const F29<P>& a) {
  // a * R9 / R9: the product is < p * (1 + a / p / 169) < 2p
  return cond_sub29(mul29(a, F29<P>::one()));
}

template <class P>
H2B_HD bool is_zero_canon29(const F29<P>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) o |= a.l[i];
  return o == 0;
}

// Cheap necessary condition for  a == 0 (mod p)  when 0 <= a < 32p (a: N):
// a = k*p  =>  a.l[0] * p^-1 = k (mod 2^29), k < 32.  False positives ~ 2^-24.
template <class P>
H2B_HD bool maybe_zero29(const F29<P>& a) {
  return ((a.l[0] * P::PINV) & MASK29) < 32u;
}

// boundary (R = 2^256, canonical, 8 x 32)  ->  R9-form N (value < 1.01 p)
template <class P>
H2B_HD F29<P> to_r9(const Fp<typename P::Base>& a) {
  F29<P> up;
#pragma unroll
  for (int i = 0; i < 9; ++i) up.l[i] = P::up(i);
  return mul29(unpack29<P>(a), up);
}

// R9-form N (value < 2^257)  ->  boundary form, canonical
template <class P>
H2B_HD Fp<typename P::Base> from_r9(const F29<P>& a) {
  F29<P> dn;
#pragma unroll
  for (int i = 0; i < 9; ++i) dn.l[i] = P::down(i);
  return pack29<P>(cond_sub29(mul29(a, dn)));
}

typedef F29<Fr29Params> Fr29;
typedef F29<Fq29Params> Fq29;

}  // namespace h2b
