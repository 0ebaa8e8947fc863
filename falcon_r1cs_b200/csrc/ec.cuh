// BLS12-381 G1 / G2 group arithmetic in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2; infinity <=> ZZ = 0), templated on the coordinate field (Fq or Fq2).
// Curve: y^2 = x^3 + b with a = 0, so no curve constant appears in add/double.
// Replaces ark-ec 0.3.0's Jacobian arithmetic (reached from ark-groth16's prover,
// examples/pok_sig.rs:32); results are compared in affine form, where any correct
// group law gives identical bytes.
#pragma once
#include "ff32.cuh"

namespace ec {

template <class F>
struct Affine {  // (0,0) encodes the point at infinity (not on either curve)
  F x, y;
  FF_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  FF_HD static Affine infinity() { return {F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;
  FF_HD static XYZZ infinity() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  FF_HD bool is_inf() const { return zz.is_zero(); }
  FF_HD static XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return infinity();
    return {p.x, p.y, F::one(), F::one()};
  }
  FF_HD XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

  // dbl-2008-s-1 (a = 0): 6M + 4S for a general point
  FF_HD XYZZ dbl() const {
    if (is_inf()) return *this;
    F u = y.dbl();
    F v = u.sqr();
    F w = u * v;
    F s = x * v;
    F x2 = x.sqr();
    F m = x2.dbl() + x2;
    F x3 = m.sqr() - s.dbl();
    F y3 = F::mul_sub2(m, s - x3, w, y);
    return {x3, y3, v * zz, w * zzz};
  }
  // mdbl-2008-s-1: doubling of an affine point
  FF_HD static XYZZ dbl_affine(const Affine<F>& p) {
    if (p.is_inf()) return infinity();
    F u = p.y.dbl();
    F v = u.sqr();
    F w = u * v;
    F s = p.x * v;
    F x2 = p.x.sqr();
    F m = x2.dbl() + x2;
    F x3 = m.sqr() - s.dbl();
    F y3 = F::mul_sub2(m, s - x3, w, p.y);
    return {x3, y3, v, w};
  }
  // madd-2008-s: 8M + 2S (Y3 = R (Q - X3) - Y1 PPP as one fused product pair, ff32.cuh mul_sub2).  neg = add -p.
  FF_HD void add_mixed(const Affine<F>& p, bool negate = false) {
    if (p.is_inf()) return;
    F py = negate ? p.y.neg() : p.y;
    if (is_inf()) {
      x = p.x;
      y = py;
      zz = F::one();
      zzz = F::one();
      return;
    }
    F u2 = p.x * zz;
    F s2 = py * zzz;
    F pp = u2 - x;
    F r = s2 - y;
    if (pp.is_zero()) {
      if (r.is_zero()) {
        Affine<F> q = {p.x, py};
        *this = dbl_affine(q);
      } else {
        *this = infinity();
      }
      return;
    }
    F p2 = pp.sqr();
    F p3 = pp * p2;
    F q = x * p2;
    F x3 = r.sqr() - p3 - q.dbl();
    zz = zz * p2;  // (before y: p2 is dead by then, one value fewer alive across the fused product)
    y = F::mul_sub2(r, q - x3, y, p3);
    x = x3;
    zzz = zzz * p3;
  }
  // add-2008-s: 12M + 2S
  FF_HD void add(const XYZZ& o) {
    if (o.is_inf()) return;
    if (is_inf()) {
      *this = o;
      return;
    }
    F u1 = x * o.zz, u2 = o.x * zz;
    F s1 = y * o.zzz, s2 = o.y * zzz;
    F pp = u2 - u1;
    F r = s2 - s1;
    if (pp.is_zero()) {
      if (r.is_zero())
        *this = dbl();
      else
        *this = infinity();
      return;
    }
    F p2 = pp.sqr();
    F p3 = pp * p2;
    F q = u1 * p2;
    F x3 = r.sqr() - p3 - q.dbl();
    y = F::mul_sub2(r, q - x3, s1, p3);
    x = x3;
    zz = zz * o.zz * p2;
    zzz = zzz * o.zzz * p3;
  }
  FF_HD Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::infinity();
    F zi = zzz.inverse();        // 1/ZZZ
    F zi2 = (zi * zz).sqr();     // (ZZ/ZZZ)^2 = 1/ZZ   (ZZ^3 = ZZZ^2)
    return {x * zi2, y * zi};
  }
  // k: canonical little-endian limbs, nbits significant bits (MSB-first double-and-add)
  FF_HD XYZZ mul(const uint32_t* k, int nbits) const {
    XYZZ r = infinity();
    for (int i = nbits - 1; i >= 0; i--) {
      r = r.dbl();
      if ((k[i >> 5] >> (i & 31)) & 1) r.add(*this);
    }
    return r;
  }
};

// out[k] = 2^(CB(k+1)) p in affine form for k = 0 .. W-2 (the MSM's per-window copies of a
// base; CB = window width in bits).  One chain of doublings; a single batched inversion of the W-1 ZZZ coordinates.
template <class F, int W, int CB = 16>
FF_NOINLINE void window_multiples(const Affine<F>& p, Affine<F>* out) {
  if (p.is_inf()) {
    for (int k = 0; k < W - 1; k++) out[k] = p;
    return;
  }
  F zz[W - 1], zzz[W - 1], pre[W - 1];
  XYZZ<F> q = XYZZ<F>::from_affine(p);
  F acc = F::one();
  for (int k = 0; k < W - 1; k++) {
    for (int d = 0; d < CB; d++) q = q.dbl();
    out[k] = {q.x, q.y};  // X, Y for now
    zz[k] = q.zz;
    zzz[k] = q.zzz;
    pre[k] = acc;
    acc = acc * q.zzz;  // a point of odd prime order never doubles to infinity
  }
  F inv = acc.inverse();
  for (int k = W - 2; k >= 0; k--) {
    F zi = inv * pre[k];         // 1/ZZZ_k
    inv = inv * zzz[k];
    F zi2 = (zi * zz[k]).sqr();  // 1/ZZ_k  (ZZ^3 = ZZZ^2)
    out[k] = {out[k].x * zi2, out[k].y * zi};
  }
}

// ---- affine + affine with the inversion taken out ("batched affine" bucket accumulation, msm_impl.cuh pair_kernel) ----
// pair_classify names the case and, for the two cases that divide, the denominator; pair_finish completes the sum given
// 1 / den: 1 multiplication for lambda, 1 squaring, 1 multiplication for y3 (against 8M + 2S for a mixed addition in XYZZ
// coordinates); the inversion is shared by a whole batch of pairs (Montgomery's trick, 3 more multiplications a pair).
// has2 = false: the pair is a single point.
enum PairKind { PAIR_GEN = 0, PAIR_DBL = 1, PAIR_P1 = 2, PAIR_P2 = 3, PAIR_INF = 4 };
template <class F>
FF_HD int pair_classify(const Affine<F>& p1, const Affine<F>& p2, bool has2, F& den) {
  if (!has2 || p2.is_inf()) return PAIR_P1;
  if (p1.is_inf()) return PAIR_P2;
  den = p2.x - p1.x;
  if (!den.is_zero()) return PAIR_GEN;
  if (p1.y == p2.y && !p1.y.is_zero()) {
    den = p1.y.dbl();
    return PAIR_DBL;
  }
  return PAIR_INF;  // p2 = -p1 (or a point of order two)
}
template <class F>
FF_HD Affine<F> pair_finish(int kind, const Affine<F>& p1, const Affine<F>& p2, const F& dinv) {
  if (kind == PAIR_P1) return p1;
  if (kind == PAIR_P2) return p2;
  if (kind == PAIR_INF) return Affine<F>::infinity();
  F num;
  if (kind == PAIR_GEN) {
    num = p2.y - p1.y;
  } else {
    F x2 = p1.x.sqr();
    num = x2.dbl() + x2;
  }
  F lam = num * dinv;
  F x3 = lam.sqr() - p1.x - p2.x;  // doubling: p2.x == p1.x
  F y3 = lam * (p1.x - x3) - p1.y;
  return {x3, y3};
}

typedef Affine<ff::Fq> G1Affine;
typedef Affine<ff::Fq2> G2Affine;
typedef XYZZ<ff::Fq> G1;
typedef XYZZ<ff::Fq2> G2;

}  // namespace ec
