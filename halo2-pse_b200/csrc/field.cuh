// bn256 Fr / Fq arithmetic, 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces (for the hot path only) halo2curves 0.3.1 `bn256::Fr` / `bn256::Fq`
// (external crate, /root/reference/halo2_proofs/Cargo.toml:51).  The residues
// are bit-identical to halo2curves' 4 x u64 limbs because R is the same.
//
// Device path: PTX carry chains (mad.lo.cc / madc.hi.cc pairs, which ptxas
// fuses into IMAD.WIDE.U32(.X)); the product is accumulated into two
// interleaved column accumulators (columns with (i+j) even / odd) so every
// chain is one unbroken carry chain, with the Montgomery reduction interleaved
// row by row (CIOS).  136 32x32 multiplies per modular multiplication.
// Host path: the same algorithm with 64-bit emulation of each chain, used by
// the host-side finishers and by the CPU unit test of this header.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define H2B_HD __host__ __device__ __forceinline__
#define H2B_D __device__ __forceinline__
#else
#define H2B_HD inline
#define H2B_D inline
#endif

#include "shoup_chains.cuh"  // generated carry chains of mul_shoup (below)

namespace h2b {

// ---------------------------------------------------------------------------
// Parameters
// ---------------------------------------------------------------------------
struct FrParams {
  // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
  static H2B_HD constexpr uint32_t mod(int i) {
    constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return t[i];
  }
  // R mod r
  static H2B_HD constexpr uint32_t one(int i) {
    constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  // R^2 mod r
  static H2B_HD constexpr uint32_t r2(int i) {
    constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return t[i];
  }
  static constexpr uint32_t INV = 0xefffffffu;  // -r^{-1} mod 2^32
};

struct FqParams {
  // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
  static H2B_HD constexpr uint32_t mod(int i) {
    constexpr uint32_t t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t one(int i) {
    constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  static H2B_HD constexpr uint32_t r2(int i) {
    constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return t[i];
  }
  static constexpr uint32_t INV = 0xe4866389u;  // -q^{-1} mod 2^32
};

// ---------------------------------------------------------------------------
// Carry-chain building blocks.  Each has a PTX body and a host emulation.
// ---------------------------------------------------------------------------

// c[0..7] += {lo,hi}(x0*b), {lo,hi}(x1*b), {lo,hi}(x2*b), {lo,hi}(x3*b);  c[8] += carry.
// If SEED, the chain's carry-in is the carry of (u + v).
template <bool SEED>
H2B_HD void chain_mad_top(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                          uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t& c8, uint32_t x0,
                          uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b, uint32_t u,
                          uint32_t v) {
#ifdef __CUDA_ARCH__
  if (SEED) {
    uint32_t tmp;
    asm("add.cc.u32 %9, %15, %16;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7),
          "+r"(c8), "=r"(tmp)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b), "r"(u), "r"(v));
  } else {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7),
          "+r"(c8)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
  }
#else
  uint64_t carry = SEED ? (((uint64_t)u + v) >> 32) : 0;
  uint32_t* c[9] = {&c0, &c1, &c2, &c3, &c4, &c5, &c6, &c7, &c8};
  const uint32_t x[4] = {x0, x1, x2, x3};
  for (int k = 0; k < 4; ++k) {
    uint64_t p = (uint64_t)x[k] * b;
    uint64_t s = (uint64_t)*c[2 * k] + (uint32_t)p + carry;
    *c[2 * k] = (uint32_t)s;
    carry = s >> 32;
    s = (uint64_t)*c[2 * k + 1] + (uint32_t)(p >> 32) + carry;
    *c[2 * k + 1] = (uint32_t)s;
    carry = s >> 32;
  }
  *c[8] += (uint32_t)carry;
#endif
}

// c[0..7] += {lo,hi}(x0*b) ... {lo,hi}(x3*b); the carry out of c[7] is known to be 0.
H2B_HD void chain_mad(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                      uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t x0, uint32_t x1,
                      uint32_t x2, uint32_t x3, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7)
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
  uint64_t carry = 0;
  uint32_t* c[8] = {&c0, &c1, &c2, &c3, &c4, &c5, &c6, &c7};
  const uint32_t x[4] = {x0, x1, x2, x3};
  for (int k = 0; k < 4; ++k) {
    uint64_t p = (uint64_t)x[k] * b;
    uint64_t s = (uint64_t)*c[2 * k] + (uint32_t)p + carry;
    *c[2 * k] = (uint32_t)s;
    carry = s >> 32;
    s = (uint64_t)*c[2 * k + 1] + (uint32_t)(p >> 32) + carry;
    *c[2 * k + 1] = (uint32_t)s;
    carry = s >> 32;
  }
#endif
}

H2B_HD void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

// r = a + b + carry_of(u + v)   (8 limbs, result < 2^256 guaranteed by caller)
H2B_HD void add8_seed(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t u, uint32_t v) {
#ifdef __CUDA_ARCH__
  uint32_t tmp;
  asm("add.cc.u32 %8, %25, %26;\n\t"
      "addc.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.u32 %7, %16, %24;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(tmp)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]),
        "r"(u), "r"(v));
#else
  uint64_t carry = ((uint64_t)u + v) >> 32;
  for (int i = 0; i < 8; ++i) {
    uint64_t s = (uint64_t)a[i] + b[i] + carry;
    r[i] = (uint32_t)s;
    carry = s >> 32;
  }
#endif
}

// r = a + b, returns nothing (caller guarantees no overflow past 2^256)
H2B_HD void add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t carry = 0;
  for (int i = 0; i < 8; ++i) {
    uint64_t s = (uint64_t)a[i] + b[i] + carry;
    r[i] = (uint32_t)s;
    carry = s >> 32;
  }
#endif
}

// r = a - b, returns the borrow as an all-ones / all-zero mask
H2B_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t mask;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(mask)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t borrow = 0;
  for (int i = 0; i < 8; ++i) {
    uint64_t d = (uint64_t)a[i] - b[i] - borrow;
    r[i] = (uint32_t)d;
    borrow = (d >> 32) & 1;
  }
  mask = (uint32_t)(0 - borrow);
#endif
  return mask;
}


// c[0 .. 2K-1] += {lo,hi}(x[0]*b), {lo,hi}(x[2]*b), ... (K products, operands two apart); c[2K] += carry.
template <int K>
H2B_HD void chain_mad_k(uint32_t* c, const uint32_t* x, uint32_t b) {
#ifdef __CUDA_ARCH__
  if (K == 1) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2])
        : "r"(x[0]), "r"(b));
  }
  else if (K == 2) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4])
        : "r"(x[0]), "r"(x[2]), "r"(b));
  }
  else if (K == 3) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6])
        : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(b));
  }
  else if (K == 4) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(c[8])
        : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(b));
  }
#else
  uint64_t carry = 0;
  for (int k = 0; k < K; ++k) {
    uint64_t p = (uint64_t)x[2 * k] * b;
    uint64_t s = (uint64_t)c[2 * k] + (uint32_t)p + carry;
    c[2 * k] = (uint32_t)s;
    carry = s >> 32;
    s = (uint64_t)c[2 * k + 1] + (uint32_t)(p >> 32) + carry;
    c[2 * k + 1] = (uint32_t)s;
    carry = s >> 32;
  }
  c[2 * K] += (uint32_t)carry;
#endif
}

// r = a + b over 16 limbs (caller guarantees no overflow past 2^512)
H2B_HD void add16(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %16, %32;\n\t"
      "addc.cc.u32 %1, %17, %33;\n\t"
      "addc.cc.u32 %2, %18, %34;\n\t"
      "addc.cc.u32 %3, %19, %35;\n\t"
      "addc.cc.u32 %4, %20, %36;\n\t"
      "addc.cc.u32 %5, %21, %37;\n\t"
      "addc.cc.u32 %6, %22, %38;\n\t"
      "addc.cc.u32 %7, %23, %39;\n\t"
      "addc.cc.u32 %8, %24, %40;\n\t"
      "addc.cc.u32 %9, %25, %41;\n\t"
      "addc.cc.u32 %10, %26, %42;\n\t"
      "addc.cc.u32 %11, %27, %43;\n\t"
      "addc.cc.u32 %12, %28, %44;\n\t"
      "addc.cc.u32 %13, %29, %45;\n\t"
      "addc.cc.u32 %14, %30, %46;\n\t"
      "addc.u32 %15, %31, %47;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]), "r"(b[14]), "r"(b[15]));
#else
  uint64_t carry = 0;
  for (int i = 0; i < 16; ++i) {
    uint64_t s = (uint64_t)a[i] + b[i] + carry;
    r[i] = (uint32_t)s;
    carry = s >> 32;
  }
#endif
}

// t[0..15] = 2 * s[0..15] + sum_i a[i]^2 * 2^(64 i)   (the result is known to fit: it is a square below 2^512)
H2B_HD void sqr_finish16(uint32_t* t, const uint32_t* s, const uint32_t* a) {
#ifdef __CUDA_ARCH__
  uint32_t d[16];
  d[0] = s[0] << 1;
#pragma unroll
  for (int k = 1; k < 16; ++k) d[k] = __funnelshift_l(s[k - 1], s[k], 1);
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
      "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
      "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
      "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
      "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
      "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
      "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
      "madc.hi.u32 %15, %23, %23, %15;"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]), "+r"(d[8]),
        "+r"(d[9]), "+r"(d[10]), "+r"(d[11]), "+r"(d[12]), "+r"(d[13]), "+r"(d[14]), "+r"(d[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#pragma unroll
  for (int k = 0; k < 16; ++k) t[k] = d[k];
#else
  uint64_t carry = 0;
  uint32_t prev = 0;
  for (int i = 0; i < 8; ++i) {
    const uint64_t p = (uint64_t)a[i] * a[i];
    const uint32_t d0 = (s[2 * i] << 1) | (prev >> 31), d1 = (s[2 * i + 1] << 1) | (s[2 * i] >> 31);
    prev = s[2 * i + 1];
    uint64_t x = (uint64_t)d0 + (uint32_t)p + carry;
    t[2 * i] = (uint32_t)x;
    carry = x >> 32;
    x = (uint64_t)d1 + (uint32_t)(p >> 32) + carry;
    t[2 * i + 1] = (uint32_t)x;
    carry = x >> 32;
  }
#endif
}

// ---------------------------------------------------------------------------
// Field element
// ---------------------------------------------------------------------------
template <class P>
struct Fp {
  uint32_t v[8];

  static H2B_HD Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0;
    return r;
  }
  static H2B_HD Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = P::one(i);
    return r;
  }
  static H2B_HD Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = P::r2(i);
    return r;
  }
  H2B_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= v[i];
    return o == 0;
  }
  H2B_HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  H2B_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// if x >= p: x -= p      (x < 2p)
template <class P>
H2B_HD void reduce_once(Fp<P>& x) {
  uint32_t m[8], t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = P::mod(i);
  uint32_t borrow = sub8(t, x.v, m);
#pragma unroll
  for (int i = 0; i < 8; ++i) x.v[i] = borrow ? x.v[i] : t[i];
}

template <class P>
H2B_HD Fp<P> add(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  add8(r.v, a.v, b.v);  // a,b < p < 2^254: no overflow
  reduce_once(r);
  return r;
}

template <class P>
H2B_HD Fp<P> sub(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  uint32_t mask = sub8(r.v, a.v, b.v);
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = P::mod(i) & mask;
  add8(r.v, r.v, m);
  return r;
}

template <class P>
H2B_HD Fp<P> neg(const Fp<P>& a) {
  Fp<P> r;
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = P::mod(i);
  sub8(r.v, m, a.v);
  bool z = a.is_zero();
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = z ? 0u : r.v[i];
  return r;
}

template <class P>
H2B_HD Fp<P> dbl(const Fp<P>& a) {
  return add(a, a);
}

// Montgomery product a*b/R mod p.  REDUCE = false leaves the result below 2p (lazy residues, see add_lazy).
template <class P, bool REDUCE = true>
H2B_HD Fp<P> mul(const Fp<P>& a, const Fp<P>& b) {
  // A[s][c]: absolute column c of the accumulator collecting the products
  // a_j*b_i (and m_i*p_j) with (i + j) & 1 == s.
  uint32_t A[2][18];
#pragma unroll
  for (int c = 0; c < 18; ++c) A[0][c] = A[1][c] = 0;

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s0 = i & 1;   // accumulator of the even-j products of this row
    const int s1 = s0 ^ 1;  // accumulator of the odd-j products
    uint32_t* E = &A[s0][i];      // columns i .. i+8
    uint32_t* O = &A[s1][i + 1];  // columns i+1 .. i+8
    const uint32_t bi = b.v[i];
    if (i == 0) {
      // first row: every product owns its two columns, no carries
      mul_wide(E[0], E[1], a.v[0], bi);
      mul_wide(E[2], E[3], a.v[2], bi);
      mul_wide(E[4], E[5], a.v[4], bi);
      mul_wide(E[6], E[7], a.v[6], bi);
      mul_wide(O[0], O[1], a.v[1], bi);
      mul_wide(O[2], O[3], a.v[3], bi);
      mul_wide(O[4], O[5], a.v[5], bi);
      mul_wide(O[6], O[7], a.v[7], bi);
    } else {
      // carry-in: column i-1 of both accumulators sums to 0 or 2^32
      chain_mad_top<true>(E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], E[8], a.v[0], a.v[2],
                          a.v[4], a.v[6], bi, A[0][i - 1], A[1][i - 1]);
      chain_mad(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], a.v[1], a.v[3], a.v[5], a.v[7],
                bi);
    }
    const uint32_t m = (A[0][i] + A[1][i]) * P::INV;
    chain_mad_top<false>(E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], E[8], P::mod(0),
                         P::mod(2), P::mod(4), P::mod(6), m, 0u, 0u);
    chain_mad(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], P::mod(1), P::mod(3), P::mod(5),
              P::mod(7), m);
  }
  Fp<P> r;
  add8_seed(r.v, &A[0][8], &A[1][8], A[0][7], A[1][7]);
  if (REDUCE) reduce_once(r);
  return r;
}


template <class P>
H2B_HD Fp<P> to_mont(const Fp<P>& a) {
  return mul(a, Fp<P>::r2());
}

// r = t / 2^256 mod p for a 16-word t < p * 2^256: the reduction half of `mul` alone (64 + 8 multiplies), on the same
// two interleaved column accumulators.  The high words join at the end: columns >= 8 must start empty, because the
// ends of the chains (E[8], O[7]) do not capture a carry-out.
template <class P, bool HIGH>
H2B_HD Fp<P> mont_reduce(const uint32_t* t) {
  uint32_t A[2][18];
#pragma unroll
  for (int c = 0; c < 18; ++c) {
    A[0][c] = c < 8 ? t[c] : 0u;
    A[1][c] = 0;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s0 = i & 1, s1 = s0 ^ 1;
    uint32_t* E = &A[s0][i];
    uint32_t* O = &A[s1][i + 1];
    uint32_t x;
    if (i == 0) {
      x = A[0][0] + A[1][0];
    } else {
#ifdef __CUDA_ARCH__
      uint32_t tmp;
      asm("add.cc.u32 %1, %2, %3;\n\taddc.u32 %0, %4, %5;"
          : "=r"(x), "=r"(tmp)
          : "r"(A[0][i - 1]), "r"(A[1][i - 1]), "r"(A[0][i]), "r"(A[1][i]));
#else
      x = A[0][i] + A[1][i] + (uint32_t)(((uint64_t)A[0][i - 1] + A[1][i - 1]) >> 32);
#endif
    }
    const uint32_t m = x * P::INV;
    if (i == 0)
      chain_mad_top<false>(E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], E[8], P::mod(0), P::mod(2), P::mod(4),
                           P::mod(6), m, 0u, 0u);
    else
      chain_mad_top<true>(E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], E[8], P::mod(0), P::mod(2), P::mod(4),
                          P::mod(6), m, A[0][i - 1], A[1][i - 1]);
    chain_mad(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
  }
  Fp<P> r;
  add8_seed(r.v, &A[0][8], &A[1][8], A[0][7], A[1][7]);  // <= p
  if (HIGH) add8(r.v, r.v, t + 8);                       // + floor(t / 2^256) < p: below 2p
  reduce_once(r);
  return r;
}

// Montgomery square a*a/R mod p: the 28 products a_i*a_j (i < j) once, doubled, plus the 8 squares a_i^2 (36 wide
// multiplies instead of 64), then the Montgomery reduction alone (64 + 8): 108 wide multiplies against 136 of `mul`.
// Rows i = 0..6 multiply a_i into a_(i+1..7); the products with i + j odd and those with i + j even go to separate
// accumulators (so that consecutive products of a chain own consecutive column pairs), as in `mul`.
template <class P>
H2B_HD Fp<P> sqr(const Fp<P>& a) {
#ifdef H2B_SQR_IS_MUL
  return mul(a, a);
#else
  uint32_t A[2][17];
#pragma unroll
  for (int c = 0; c < 17; ++c) A[0][c] = A[1][c] = 0;
  // odd accumulator: j = i + 1, i + 3, ... at columns 2i + 1 ...;  even accumulator: j = i + 2, i + 4, ... at 2i + 2 ...
  chain_mad_k<4>(&A[1][1], &a.v[1], a.v[0]);
  chain_mad_k<3>(&A[0][2], &a.v[2], a.v[0]);
  chain_mad_k<3>(&A[1][3], &a.v[2], a.v[1]);
  chain_mad_k<3>(&A[0][4], &a.v[3], a.v[1]);
  chain_mad_k<3>(&A[1][5], &a.v[3], a.v[2]);
  chain_mad_k<2>(&A[0][6], &a.v[4], a.v[2]);
  chain_mad_k<2>(&A[1][7], &a.v[4], a.v[3]);
  chain_mad_k<2>(&A[0][8], &a.v[5], a.v[3]);
  chain_mad_k<2>(&A[1][9], &a.v[5], a.v[4]);
  chain_mad_k<1>(&A[0][10], &a.v[6], a.v[4]);
  chain_mad_k<1>(&A[1][11], &a.v[6], a.v[5]);
  chain_mad_k<1>(&A[0][12], &a.v[7], a.v[5]);
  chain_mad_k<1>(&A[1][13], &a.v[7], a.v[6]);
  uint32_t s16[16], t[16];
  add16(s16, &A[0][0], &A[1][0]);  // the off-diagonal sum: below 2^511
  sqr_finish16(t, s16, a.v);
  return mont_reduce<P, true>(t);
#endif
}

// t[0..15] = a * b as integers (64 wide multiplies), on the two column accumulators of `mul`: row i sends its even-j
// products to accumulator i & 1 at columns i .., its odd-j products to the other at columns i + 1 ...  Rows go in
// ascending order, so the column that takes a chain's carry-out is still empty or holds an earlier carry (0 or 1).
template <class P>
H2B_HD void mul_wide16(uint32_t* t, const Fp<P>& a, const Fp<P>& b) {
  uint32_t A[2][18];
#pragma unroll
  for (int c = 0; c < 18; ++c) A[0][c] = A[1][c] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    chain_mad_k<4>(&A[i & 1][i], &a.v[0], b.v[i]);
    chain_mad_k<4>(&A[(i & 1) ^ 1][i + 1], &a.v[1], b.v[i]);
  }
  add16(t, &A[0][0], &A[1][0]);
}

// a*b - c*d (Montgomery form): both products share ONE reduction (2 * 64 + 64 + 8 wide multiplies instead of
// 2 * 136).  a*b + (p - c)*d < 2 p^2 < p * 2^256, the precondition of mont_reduce.
template <class P>
H2B_HD Fp<P> mul_sub(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
  Fp<P> nc;
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = P::mod(i);
  sub8(nc.v, m, c.v);  // p - c in (0, p]: p itself stands for 0, the bound above still holds
  uint32_t t[16], u[16];
  mul_wide16(t, a, b);
  mul_wide16(u, nc, d);
  add16(t, t, u);
  return mont_reduce<P, true>(t);
}

// ---------------------------------------------------------------------------
// Product by a FIXED multiplier (Shoup): x * w mod p for a plain (non-Montgomery) w < p with the precomputed
// w' = floor(w * 2^256 / p).  q = floor(x * w' / 2^256) is floor(x * w / p) or one less, so x*w - q*p lies in [0, 2p);
// both products of that difference are needed modulo 2^256 only, and of x * w' only the top half:
//   28 + 8 (low 256 bits of x*w)  +  36 + 7 (columns >= 7 of x*w', the guard column's high words only)
//   + 28 + 8 (low 256 bits of q*p)  =  92 wide, 7 high and 16 low multiplies against 128 wide + 8 low of `mul`.
// Dropping the columns below the guard makes q at most one smaller again: the result is in [0, 3p), two conditional
// subtractions.  x in Montgomery form gives x*w in Montgomery form: the transforms multiply by table constants only
// (roots of unity), so their intra-pass products take this path (csrc/ntt.cu).
// ---------------------------------------------------------------------------
// r[0..7] = low 256 bits of x * y
H2B_HD void mul_low8(uint32_t* r, const uint32_t* x, const uint32_t* y) {
  uint32_t A[2][8];
#pragma unroll
  for (int c = 0; c < 8; ++c) A[0][c] = A[1][c] = 0;
  // row i multiplies y[i] into x[0 .. 7-i]: even-j products to accumulator i & 1 at column i, odd-j to the other at i + 1
  chain_low<4, false>(&A[0][0], &x[0], y[0]);
  chain_low<3, true>(&A[1][1], &x[1], y[0]);
  chain_low<3, true>(&A[1][1], &x[0], y[1]);
  chain_low<3, false>(&A[0][2], &x[1], y[1]);
  chain_low<3, false>(&A[0][2], &x[0], y[2]);
  chain_low<2, true>(&A[1][3], &x[1], y[2]);
  chain_low<2, true>(&A[1][3], &x[0], y[3]);
  chain_low<2, false>(&A[0][4], &x[1], y[3]);
  chain_low<2, false>(&A[0][4], &x[0], y[4]);
  chain_low<1, true>(&A[1][5], &x[1], y[4]);
  chain_low<1, true>(&A[1][5], &x[0], y[5]);
  chain_low<1, false>(&A[0][6], &x[1], y[5]);
  chain_low<1, false>(&A[0][6], &x[0], y[6]);
  chain_low<0, true>(&A[1][7], &x[1], y[6]);
  chain_low<0, true>(&A[1][7], &x[0], y[7]);
  add8(r, A[0], A[1]);  // modulo 2^256
}

// q[0..7] = floor(x * y / 2^256) or one less (columns below 6 and the low words of column 6 are dropped)
H2B_HD void mul_high8(uint32_t* q, const uint32_t* x, const uint32_t* y) {
  uint32_t H[2][10];  // columns 7 .. 16
#pragma unroll
  for (int c = 0; c < 10; ++c) H[0][c] = H[1][c] = 0;
  // row i multiplies y[i] into x[6-i] (high word only, column 7) and x[7-i .. 7]; products with i + j even go to
  // accumulator 0, odd to accumulator 1; rows ascend, so every carry-out lands on an empty word or an earlier carry
  chain_high<true, 0>(&H[0][0], &x[6], y[0]);
  chain_high<false, 1>(&H[1][0], &x[7], y[0]);
  chain_high<true, 1>(&H[0][0], &x[5], y[1]);
  chain_high<false, 1>(&H[1][0], &x[6], y[1]);
  chain_high<true, 1>(&H[0][0], &x[4], y[2]);
  chain_high<false, 2>(&H[1][0], &x[5], y[2]);
  chain_high<true, 2>(&H[0][0], &x[3], y[3]);
  chain_high<false, 2>(&H[1][0], &x[4], y[3]);
  chain_high<true, 2>(&H[0][0], &x[2], y[4]);
  chain_high<false, 3>(&H[1][0], &x[3], y[4]);
  chain_high<true, 3>(&H[0][0], &x[1], y[5]);
  chain_high<false, 3>(&H[1][0], &x[2], y[5]);
  chain_high<true, 3>(&H[0][0], &x[0], y[6]);
  chain_high<false, 4>(&H[1][0], &x[1], y[6]);
  chain_high<false, 4>(&H[0][1], &x[1], y[7]);
  chain_high<false, 4>(&H[1][0], &x[0], y[7]);
  add8_seed(q, &H[0][1], &H[1][1], H[0][0], H[1][0]);  // columns 8 .. 15 + the carry out of column 7
}

// x * w mod p, given w (plain, < p) and wp = floor(w * 2^256 / p)
// FULL = false: one conditional subtraction only, the result is below 2p for any x < 4p (lazy residues)
template <class P, bool FULL = true>
H2B_HD Fp<P> mul_shoup(const Fp<P>& x, const Fp<P>& w, const Fp<P>& wp) {
  uint32_t q[8], lo[8], qp[8], pm[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) pm[i] = P::mod(i);
  mul_high8(q, x.v, wp.v);
  mul_low8(lo, x.v, w.v);
  mul_low8(qp, q, pm);
  Fp<P> r;
  sub8(r.v, lo, qp);  // exact: the true difference is below p (2 + x / 2^256) < 3p < 2^256
  reduce_once(r);
  if (FULL) reduce_once(r);
  return r;
}

// ---------------------------------------------------------------------------
// Lazy residues: values kept in [0, 2p) (4p < 2^256 leaves the room).  The transforms run on them inside and between
// their passes and reduce once at the very end: a product then needs no final subtraction at all (Montgomery: inputs
// below 2p and p give (2p*p + R*p)/R < 1.38p) or one instead of two (Shoup), and additions / subtractions cost what
// they cost on canonical residues.
// ---------------------------------------------------------------------------
template <class P>
H2B_HD constexpr uint32_t mod2(int i) {  // limb i of 2p
  return (P::mod(i) << 1) | (i ? P::mod(i - 1) >> 31 : 0u);
}

// a + b for a, b in [0, 2p): result in [0, 2p)
template <class P>
H2B_HD Fp<P> add_lazy(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  add8(r.v, a.v, b.v);  // below 4p < 2^256
  uint32_t m[8], t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = mod2<P>(i);
  const uint32_t borrow = sub8(t, r.v, m);
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = borrow ? r.v[i] : t[i];
  return r;
}

// a - b for a, b in [0, 2p): result in [0, 2p)
template <class P>
H2B_HD Fp<P> sub_lazy(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  const uint32_t mask = sub8(r.v, a.v, b.v);
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = mod2<P>(i) & mask;
  add8(r.v, r.v, m);
  return r;
}

// a + b for a, b in [0, 2p), NOT brought back below 2p (below 4p): for a sum that goes straight into a product
template <class P>
H2B_HD Fp<P> add_wide(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  add8(r.v, a.v, b.v);
  return r;
}

// a - b + 2p for a, b in [0, 2p): in (0, 4p), NOT brought back below 2p -- for a difference that goes straight into a
// Shoup product (which takes any operand below 4p and returns a lazy residue)
template <class P>
H2B_HD Fp<P> sub_wide(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  sub8(r.v, a.v, b.v);  // modulo 2^256; the true value a - b + 2p is positive and below 4p < 2^256
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = mod2<P>(i);
  add8(r.v, r.v, m);
  return r;
}

// [0, 2p) -> [0, p)
template <class P>
H2B_HD Fp<P> canon(Fp<P> a) {
  reduce_once(a);
  return a;
}

// wp = floor(w * 2^256 / p) for a plain w < p, from its Montgomery form wm = w * 2^256 mod p:
// w * 2^256 = wp * p + wm, so wp = -wm * p^-1 modulo 2^256 (table builders; generic loops)
template <class P>
H2B_HD Fp<P> shoup_companion(const Fp<P>& wm) {
  uint32_t pm[8], inv[8], t[8], u[8];
  for (int i = 0; i < 8; ++i) {
    pm[i] = P::mod(i);
    inv[i] = 0;
  }
  inv[0] = 1;
  for (int it = 0; it < 8; ++it) {  // Newton: inv <- inv * (2 - p * inv) doubles the correct low bits (1 -> 256)
    mul_low8(t, pm, inv);
    uint32_t borrow = 0;
    for (int i = 0; i < 8; ++i) {  // u = 2 - t
      const uint64_t d = (uint64_t)(i == 0 ? 2u : 0u) - t[i] - borrow;
      u[i] = (uint32_t)d;
      borrow = (uint32_t)(d >> 32) & 1u;
    }
    mul_low8(t, inv, u);
    for (int i = 0; i < 8; ++i) inv[i] = t[i];
  }
  uint32_t neg[8];
  uint32_t borrow = 0;
  for (int i = 0; i < 8; ++i) {  // neg = 2^256 - wm
    const uint64_t d = (uint64_t)0 - wm.v[i] - borrow;
    neg[i] = (uint32_t)d;
    borrow = (uint32_t)(d >> 32) & 1u;
  }
  Fp<P> r;
  mul_low8(r.v, neg, inv);
  return r;
}

// Montgomery -> canonical residue: a / R mod p, the reduction alone (half the multiplies of a product by 1)
template <class P>
H2B_HD Fp<P> from_mont(const Fp<P>& a) {
  return mont_reduce<P, false>(a.v);
}

// a^e for a small public exponent (host finishers, table builders)
template <class P>
H2B_HD Fp<P> pow_u64(const Fp<P>& a, uint64_t e) {
  Fp<P> r = Fp<P>::one();
  Fp<P> base = a;
  while (e) {
    if (e & 1) r = mul(r, base);
    base = sqr(base);
    e >>= 1;
  }
  return r;
}

// a^(p-2): inversion by Fermat (host finishers, table builders; not a hot op)
template <class P>
H2B_HD Fp<P> inv(const Fp<P>& a) {
  // exponent p-2, processed MSB first
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) e[i] = P::mod(i);
  e[0] -= 2;  // low limb of both moduli is >= 2, no borrow
  Fp<P> r = Fp<P>::one();
  for (int i = 7; i >= 0; --i) {
    for (int bit = 31; bit >= 0; --bit) {
      r = sqr(r);
      if ((e[i] >> bit) & 1) r = mul(r, a);
    }
  }
  return r;
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace h2b
