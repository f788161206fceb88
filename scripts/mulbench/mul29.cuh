// Experimental carry-free Montgomery product for bn256 Fq / Fr in 9 x 29-bit limbs (radix 2^261).
//
// Why another one (scripts/field29.cuh is the round-1 attempt, measured slower than the 8 x 32 carry
// chains): there every row still paid two 64-bit shifts and a 64-bit add on the ALU pipe to move the
// consumed column's carry up.  Here the per-row Montgomery factor m is a FULL 32-bit word chosen so that the
// low 32 bits of the column REGISTER become zero (m = -t0 * p0^-1 mod 2^32, p0 = the 29-bit limb 0 of p):
// the carry t0 >> 29 is then exactly hi32(t0) * 8, one more plain IMAD.WIDE.  The main loop is
// 9 x (9 + 1 + 9 + 1) = 180 multiply-adds without a single carry flag, shift or add.
//
// Bounds: a normalised (limbs < 2^29), b limbs < 2^30.  Column sums stay below 2^64:
// sum_j m*p_j <= (2^32 - 1) * sum_j p_j  (0.43 * 2^64 for Fq, 0.43 for Fr) + 9 * 2^59.
// Result: a*b/2^261 + (< 8.01) * p, limbs normalised (< 2^29, top limb small).
#pragma once
#include <stdint.h>

#include "consts29.h"

#if defined(__CUDACC__)
#define M29_HD __host__ __device__ __forceinline__
#else
#define M29_HD inline
#endif

static constexpr uint32_t MASK29 = (1u << 29) - 1u;

M29_HD void madw(uint64_t& t, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
#ifdef M29_CC
  uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
  t = ((uint64_t)hi << 32) | lo;
#else
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a), "r"(b));
#endif
#else
  t += (uint64_t)a * b;
#endif
}

template <class P>
struct L29 {
  uint32_t l[9];
};

// VARIANT 0: c3 added with a 64-bit add; 1: c3 added with mad.wide(c3, 1)
template <class P, int VARIANT>
M29_HD L29<P> mul29(const L29<P>& a, const L29<P>& b) {
  uint64_t t[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) t[j] = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const uint32_t bi = b.l[i];
#pragma unroll
    for (int j = 0; j < 9; ++j) madw(t[j], a.l[j], bi);
    const uint32_t m = (uint32_t)t[0] * P::INV0;
#pragma unroll
    for (int j = 0; j < 9; ++j) madw(t[j], m, P::P(j));
    madw(t[1], (uint32_t)(t[0] >> 32), 8u);  // low 32 bits of t[0] are zero: t[0] >> 29 == hi * 8
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = t[j + 1];
    t[8] = 0;
  }
  L29<P> r;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t lo = (uint32_t)t[j], hi = (uint32_t)(t[j] >> 32);
    r.l[j] = lo & MASK29;
    madw(t[j + 1], hi, 8u);
    if (VARIANT == 0)
      t[j + 1] += lo >> 29;
    else
      madw(t[j + 1], lo >> 29, 1u);
  }
  r.l[8] = (uint32_t)t[8];
  return r;
}

// 8 x 32 saturated words -> 9 x 29 limbs of the same integer
template <class P>
M29_HD L29<P> unpack29(const uint32_t* v) {
  L29<P> r;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int bit = 29 * i, w = bit >> 5, sh = bit & 31;
    uint32_t x = v[w] >> sh;
    if (sh > 3 && w + 1 < 8) x |= v[w + 1] << (32 - sh);
    r.l[i] = x & MASK29;
  }
  return r;
}
