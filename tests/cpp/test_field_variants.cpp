// Host-side check of the product variants of csrc/field.cuh against the plain Montgomery product `mul` (which
// tests/test_emulator.py and tests/test_gpu_parity.py check against big integers): the same functions the kernels
// compile, through their host emulation of the carry chains.  Random operands + edge values, Fr and Fq.
//   sqr                 == mul(a, a)                        (36 + 64 wide multiplies)
//   mul_sub(a,b,c,d)    == mul(a,b) - mul(c,d)              (two products, one reduction)
//   mul_shoup(x, w, w') == mul(x, wm)                       (fixed multiplier: w plain, w' its companion, wm Montgomery)
//   lazy residues: add_lazy / sub_lazy / sub_wide / add_wide / mul<false> / mul_shoup<false> on operands in [0, 2p)
//   (and below 4p where a product follows), results congruent and inside their stated ranges.
// g++ -std=c++17 -O2 -DH2B_EMU tests/cpp/test_field_variants.cpp && ./a.out
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#ifndef H2B_EMU
#define H2B_EMU 1
#endif
#include "../../halo2-pse_b200/csrc/field.cuh"
using namespace h2b;

template <class F, class P>
static int run(const char* nm, int iters) {
  uint64_t s = 88172645463325252ull;
  auto nx = [&]() {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return (uint32_t)(s >> 11) ^ (uint32_t)(s >> 40);
  };
  F pmod;
  uint32_t m2[8], tt[8];
  for (int j = 0; j < 8; ++j) {
    pmod.v[j] = P::mod(j);
    m2[j] = mod2<P>(j);
  }
  auto below2p = [&](const F& t) { return sub8(tt, t.v, m2) != 0; };
  auto lift = [&](const F& a, int on) {  // a or a + p: both represent a, both in [0, 2p)
    F r = a;
    if (on) add8(r.v, a.v, pmod.v);
    return r;
  };
  int bad = 0;
  for (int it = 0; it < iters; ++it) {
    F x, y, z, w, wm;
    for (int j = 0; j < 8; ++j) {
      x.v[j] = nx();
      y.v[j] = nx();
      z.v[j] = nx();
      w.v[j] = nx();
      wm.v[j] = nx();
    }
    for (F* f : {&x, &y, &z, &w, &wm}) {
      f->v[7] &= 0x3fffffff;
      for (int k = 0; k < 4; ++k) reduce_once(*f);
    }
    const F minus1 = sub(F::zero(), F::one());
    if (it < 64) {  // zeros, ones, p - 1 in every position
      if (it & 1) x = F::zero();
      if (it & 2) y = (it & 32) ? minus1 : F::zero();
      if (it & 4) z = minus1;
      if (it & 8) w = (it & 32) ? F::one() : minus1;
      if (it & 16) x = minus1;
    }
    if (wm.is_zero()) wm = F::one();
    if (it >= 64 && it < 80) wm = (it & 1) ? minus1 : F::one();
    int b = 0;
    b += !(sqr(x) == mul(x, x));
    b += !(mul_sub(x, y, z, w) == sub(mul(x, y), mul(z, w)));
    const F wpl = from_mont(wm), wsh = shoup_companion(wm);
    b += !(mul_shoup(x, wpl, wsh) == mul(x, wm));
    // lazy residues
    const F xl = lift(x, it & 1), yl = lift(y, (it >> 1) & 1);
    F t = add_lazy(xl, yl);
    b += !below2p(t) + !(canon(t) == add(x, y));
    t = sub_lazy(xl, yl);
    b += !below2p(t) + !(canon(t) == sub(x, y));
    t = mul<P, false>(xl, wm);
    b += !below2p(t) + !(canon(t) == mul(x, wm));
    t = mul_shoup<P, false>(xl, wpl, wsh);
    b += !below2p(t) + !(canon(t) == mul(x, wm));
    // operands below 4p into a product: results lazy again
    const F dw = sub_wide(xl, yl), sw = add_wide(xl, yl);
    t = mul_shoup<P, false>(dw, wpl, wsh);
    b += !below2p(t) + !(canon(t) == mul(sub(x, y), wm));
    t = mul_shoup<P, false>(sw, wpl, wsh);
    b += !below2p(t) + !(canon(t) == mul(add(x, y), wm));
    t = mul<P, false>(dw, wm);
    b += !below2p(t) + !(canon(t) == mul(sub(x, y), wm));
    t = mul<P, false>(sw, wm);
    b += !below2p(t) + !(canon(t) == mul(add(x, y), wm));
    b += !(mul(sw, wm) == mul(add(x, y), wm));  // full product of a wide operand: canonical
    if (b) {
      if (bad < 5) printf("%s: mismatch at iteration %d (%d checks)\n", nm, it, b);
      ++bad;
    }
  }
  printf("%s: %d iterations, %d mismatches\n", nm, iters, bad);
  return bad;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200000;
  const int bad = run<Fr, FrParams>("Fr", iters) + run<Fq, FqParams>("Fq", iters);
  return bad ? 1 : 0;
}
