// FP64-assisted Montgomery product for bn256 Fr / Fq: same inputs, same outputs (bit for bit) as `mul` of field.cuh.
//
// B200 (sm_100a) has a full-rate FP64 pipe (64 DFMA lanes/clk/SM, measured 2.2 cycles per warp instruction per
// sub-partition, scripts/microbench7.cu) that sits idle beside the integer multiplier, and the integer multiplier
// (IMAD.WIDE, 4 cycles per warp instruction) is what bounds the 8 x 32-bit product of field.cuh (130 wide multiplies,
// 94 % of the IMAD.WIDE peak).  This product splits the work between the two pipes:
//   * a * b (the half with no serial dependency) is computed on the FP64 pipe: both operands as five 52-bit limbs held
//     in doubles, each of the 25 limb products split exactly into a high and a low 52-bit half by two round-to-zero
//     fused multiply-adds (hi = fma_rz(x, y, 2^104) is 2^104 + floor(xy / 2^52) * 2^52; lo = fma_rz(x, y, (2^104 + 2^52)
//     - hi) is 2^52 + xy mod 2^52; both exact), the halves accumulated per column as 64-bit integers straight from the
//     doubles' bit patterns (the exponent patterns are pre-subtracted from the accumulators' initial values);
//   * the ten column sums are folded into sixteen 32-bit words (ALU pipe: funnel shifts and carry chains);
//   * the Montgomery reduction (the half that is a serial chain of 32-bit steps) stays on the integer multiplier:
//     64 + 8 multiplies instead of 128 + 8 + 2.
// Per product: ~85 FP64 instructions (187 cycles of that pipe), 72 integer multiplies (288 cycles), ~170 ALU
// instructions, against 556 cycles of the integer multiplier alone for field.cuh's `mul`.
//
// Host path (the CPU emulator build of the test-suite and the CPU unit test): the same algorithm with the C library's
// fma() under FE_TOWARDZERO, so the exactness argument above is exercised with IEEE arithmetic on the CPU too.
#pragma once
#include "../../halo2-pse_b200/csrc/field.cuh"

#ifndef __CUDA_ARCH__
#include <fenv.h>
#include <math.h>
#include <string.h>
#endif

namespace h2b {

H2B_HD double dfma_make(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)hi, (int)lo);
#else
  uint64_t u = ((uint64_t)hi << 32) | lo;
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}
H2B_HD uint64_t dfma_bits(double d) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
#endif
}
H2B_HD double dfma_fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rz(a, b, c);
#else
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  volatile double va = a, vb = b, vc = c;
  volatile double r = fma(va, vb, vc);
  fesetround(old);
  return r;
#endif
}
// exact by construction (both operands multiples of 2^52 below 2^105): the rounding mode does not matter
H2B_HD double dfma_sub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rz(a, b);
#else
  volatile double va = a, vb = b;
  volatile double r = va - vb;
  return r;
#endif
}

// funnel shift right: low 32 bits of ((hi:lo) >> s), 0 < s < 32
H2B_HD uint32_t dfma_shr(uint32_t lo, uint32_t hi, int s) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> s);
#endif
}

// Five 52-bit limbs of a 256-bit value (8 x u32, little endian) as doubles holding the integers exactly.
struct Limbs52 {
  double d[5];
};

H2B_HD Limbs52 dfma_unpack(const uint32_t* w) {
  const uint32_t EXP52 = 0x43300000u;  // high word of 2^52: (EXP52 | hi20, lo32) is the double 2^52 + limb
  const double two52 = 4503599627370496.0;
  Limbs52 r;
  r.d[0] = dfma_make((w[1] & 0xfffffu) | EXP52, w[0]) - two52;
  r.d[1] = dfma_make((dfma_shr(w[2], w[3], 20) & 0xfffffu) | EXP52, dfma_shr(w[1], w[2], 20)) - two52;
  r.d[2] = dfma_make(((w[4] >> 8) & 0xfffffu) | EXP52, dfma_shr(w[3], w[4], 8)) - two52;
  r.d[3] = dfma_make((dfma_shr(w[5], w[6], 28) & 0xfffffu) | EXP52, dfma_shr(w[4], w[5], 28)) - two52;
  r.d[4] = dfma_make((w[7] >> 16) | EXP52, dfma_shr(w[6], w[7], 16)) - two52;
  return r;
}

// t[0..15] = a * b (as integers, a, b < 2^256 with top limbs below 2^48), computed on the FP64 pipe.
template <bool SQUARE>
H2B_HD void dfma_product(uint32_t* t, const Limbs52& a, const Limbs52& b) {
  const double C1 = 20282409603651670423947251286016.0;                        // 2^104
  const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52 (exact)
  const uint64_t HI_BITS = 0x4670000000000000ull;  // bit pattern of 2^104
  const uint64_t LO_BITS = 0x4330000000000000ull;  // bit pattern of 2^52
  uint64_t acc[10];
  // column c receives the low halves of the products with i + j = c and the high halves of those with i + j = c - 1;
  // a doubled product (squaring) is added twice
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    int nlo = 0, nhi = 0;
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 5; ++j) {
        if (i + j == c) ++nlo;
        if (i + j == c - 1) ++nhi;
      }
    acc[c] = 0ull - ((uint64_t)nlo * LO_BITS + (uint64_t)nhi * HI_BITS);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (SQUARE && j < i) continue;
      const double hi = dfma_fma_rz(a.d[j], b.d[i], C1);
      const double lo = dfma_fma_rz(a.d[j], b.d[i], dfma_sub(C2, hi));
      const uint64_t hb = dfma_bits(hi), lb = dfma_bits(lo);
      if (SQUARE && j > i) {
        acc[i + j] += lb + lb;
        acc[i + j + 1] += hb + hb;
      } else {
        acc[i + j] += lb;
        acc[i + j + 1] += hb;
      }
    }
  }
  // fold the columns (each below 2^56, weight 2^(52 c)) into 32-bit words: column c lands at word (52 c) / 32 with a
  // left shift of (52 c) % 32 and spans three words; the partial sum of columns 0..c is below 2^(52 c + 57), inside
  // the window, so the chain never carries out of its third word
#pragma unroll
  for (int k = 0; k < 16; ++k) t[k] = 0;
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    const int w = (52 * c) / 32, s = (52 * c) % 32;
    const uint32_t lo = (uint32_t)acc[c], hi = (uint32_t)(acc[c] >> 32);
    uint32_t x0, x1, x2;
    if (s == 0) {
      x0 = lo; x1 = hi; x2 = 0;
    } else {
      x0 = lo << s;
      x1 = dfma_shr(lo, hi, 32 - s);
      x2 = hi >> (32 - s);
    }
    if (w + 2 < 16) {
#ifdef __CUDA_ARCH__
      asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
          : "+r"(t[w]), "+r"(t[w + 1]), "+r"(t[w + 2]) : "r"(x0), "r"(x1), "r"(x2));
#else
      uint64_t s0 = (uint64_t)t[w] + x0;
      uint64_t s1 = (uint64_t)t[w + 1] + x1 + (s0 >> 32);
      t[w] = (uint32_t)s0; t[w + 1] = (uint32_t)s1; t[w + 2] = t[w + 2] + x2 + (uint32_t)(s1 >> 32);
#endif
    } else {  // last column: words 14, 15 only (a * b < 2^512)
#ifdef __CUDA_ARCH__
      asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(t[w]), "+r"(t[w + 1]) : "r"(x0), "r"(x1));
#else
      uint64_t s0 = (uint64_t)t[w] + x0;
      t[w] = (uint32_t)s0; t[w + 1] = t[w + 1] + x1 + (uint32_t)(s0 >> 32);
#endif
    }
  }
}

// the Montgomery reduction of the 16 words: field.cuh's mont_reduce (integer multiplier)
template <class P>
H2B_HD Fp<P> dfma_reduce(const uint32_t* t) {
  return mont_reduce<P, true>(t);
}

template <class P>
H2B_HD Fp<P> mul_dfma(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[16];
  dfma_product<false>(t, dfma_unpack(a.v), dfma_unpack(b.v));
  return dfma_reduce<P>(t);
}
// the second operand already unpacked (a twiddle, a coordinate used by several products)
template <class P>
H2B_HD Fp<P> mul_dfma(const Fp<P>& a, const Limbs52& b) {
  uint32_t t[16];
  dfma_product<false>(t, dfma_unpack(a.v), b);
  return dfma_reduce<P>(t);
}
template <class P>
H2B_HD Fp<P> mul_dfma(const Limbs52& a, const Limbs52& b) {
  uint32_t t[16];
  dfma_product<false>(t, a, b);
  return dfma_reduce<P>(t);
}
template <class P>
H2B_HD Fp<P> sqr_dfma(const Fp<P>& a) {
  uint32_t t[16];
  const Limbs52 x = dfma_unpack(a.v);
  dfma_product<true>(t, x, x);
  return dfma_reduce<P>(t);
}

}  // namespace h2b
