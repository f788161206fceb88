// bn256 G1 (y^2 = x^3 + 3 over Fq) group law for the MSM kernels.
//
// Replaces halo2curves 0.3.1 `bn256::G1Affine` / `bn256::G1` group operations
// as used by /root/reference/halo2_proofs/src/arithmetic.rs:48,59-68,95-99.
// Accumulators are kept in extended Jacobian ("XYZZ": x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2) coordinates: mixed addition 8M+2S, full addition 12M+2S; the two products of y3 share one Montgomery reduction (mul_sub).
// All exceptional cases (identity operands, P + P, P + (-P)) are handled, so
// every routine is a complete group law on the curve.
#pragma once
#include "field.cuh"

namespace h2b {

// Boundary layout of halo2curves G1Affine: {x, y} in Montgomery form, 64 B.
// The identity is encoded as (0, 0) (not a curve point since b = 3 != 0).
struct G1Affine {
  Fq x, y;
  H2B_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct G1Xyzz {
  Fq x, y, zz, zzz;
  H2B_HD bool is_identity() const { return zz.is_zero(); }
  static H2B_HD G1Xyzz identity() {
    G1Xyzz r;
    r.x = Fq::zero();
    r.y = Fq::zero();
    r.zz = Fq::zero();
    r.zzz = Fq::zero();
    return r;
  }
  static H2B_HD G1Xyzz from_affine(const G1Affine& p) {
    G1Xyzz r;
    if (p.is_identity()) return identity();
    r.x = p.x;
    r.y = p.y;
    r.zz = Fq::one();
    r.zzz = Fq::one();
    return r;
  }
};

// 2 * (affine p), p != identity
H2B_HD G1Xyzz xyzz_double_affine(const G1Affine& p) {
  // mdbl-2008-s-1 with a = 0
  G1Xyzz r;
  Fq u = dbl(p.y);
  Fq v = sqr(u);
  Fq w = mul(u, v);
  Fq s = mul(p.x, v);
  Fq xx = sqr(p.x);
  Fq m = add(dbl(xx), xx);
  r.x = sub(sqr(m), dbl(s));
  r.y = mul_sub(m, sub(s, r.x), w, p.y);
  r.zz = v;
  r.zzz = w;
  return r;
}

H2B_HD G1Xyzz xyzz_double(const G1Xyzz& p) {
  // dbl-2008-s-1 with a = 0.  y = 0 never happens on this curve (odd order).
  if (p.is_identity()) return p;
  G1Xyzz r;
  Fq u = dbl(p.y);
  Fq v = sqr(u);
  Fq w = mul(u, v);
  Fq s = mul(p.x, v);
  Fq xx = sqr(p.x);
  Fq m = add(dbl(xx), xx);
  r.x = sub(sqr(m), dbl(s));
  r.y = mul_sub(m, sub(s, r.x), w, p.y);
  r.zz = mul(v, p.zz);
  r.zzz = mul(w, p.zzz);
  return r;
}

// acc += (affine p)
H2B_HD void xyzz_add_affine(G1Xyzz& acc, const G1Affine& p) {
  if (p.is_identity()) return;
  if (acc.is_identity()) {
    acc.x = p.x;
    acc.y = p.y;
    acc.zz = Fq::one();
    acc.zzz = Fq::one();
    return;
  }
  // madd-2008-s
  Fq u2 = mul(p.x, acc.zz);
  Fq s2 = mul(p.y, acc.zzz);
  Fq pp_ = sub(u2, acc.x);
  Fq r = sub(s2, acc.y);
  if (pp_.is_zero()) {
    if (r.is_zero()) {
      acc = xyzz_double_affine(p);
    } else {
      acc = G1Xyzz::identity();
    }
    return;
  }
  Fq pp = sqr(pp_);
  Fq ppp = mul(pp_, pp);
  Fq q = mul(acc.x, pp);
  Fq x3 = sub(sub(sqr(r), ppp), dbl(q));
  Fq y3 = mul_sub(r, sub(q, x3), acc.y, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = mul(acc.zz, pp);
  acc.zzz = mul(acc.zzz, ppp);
}

// acc += b
H2B_HD void xyzz_add(G1Xyzz& acc, const G1Xyzz& b) {
  if (b.is_identity()) return;
  if (acc.is_identity()) {
    acc = b;
    return;
  }
  // add-2008-s
  Fq u1 = mul(acc.x, b.zz);
  Fq u2 = mul(b.x, acc.zz);
  Fq s1 = mul(acc.y, b.zzz);
  Fq s2 = mul(b.y, acc.zzz);
  Fq pp_ = sub(u2, u1);
  Fq r = sub(s2, s1);
  if (pp_.is_zero()) {
    if (r.is_zero()) {
      acc = xyzz_double(acc);
    } else {
      acc = G1Xyzz::identity();
    }
    return;
  }
  Fq pp = sqr(pp_);
  Fq ppp = mul(pp_, pp);
  Fq q = mul(u1, pp);
  Fq x3 = sub(sub(sqr(r), ppp), dbl(q));
  Fq y3 = mul_sub(r, sub(q, x3), s1, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = mul(mul(acc.zz, b.zz), pp);
  acc.zzz = mul(mul(acc.zzz, b.zzz), ppp);
}

H2B_HD G1Affine g1_neg(const G1Affine& p) {
  G1Affine r;
  r.x = p.x;
  r.y = neg(p.y);
  return r;
}

// XYZZ -> affine (one inversion; host finishers only)
H2B_HD G1Affine xyzz_to_affine(const G1Xyzz& p) {
  G1Affine r;
  if (p.is_identity()) {
    r.x = Fq::zero();
    r.y = Fq::zero();
    return r;
  }
  // 1/zzz, then 1/zz = zzz^-1 * (zzz/zz) and zzz/zz = z  =>  use zz^-1 = (zzz^-1)^2 * zz^2
  Fq izzz = inv(p.zzz);
  Fq izz = mul(sqr(izzz), sqr(p.zz));  // zz^2 / zzz^2 = z^4 / z^6 = 1 / z^2
  r.x = mul(p.x, izz);
  r.y = mul(p.y, izzz);
  return r;
}

}  // namespace h2b
