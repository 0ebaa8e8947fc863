// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of ark-ec 0.3.0 short
// Weierstrass Jacobian arithmetic for BLS12-381 G1 (over Fq) and G2 (over Fq2),
// [EXT] to /root/reference (reached via ark-groth16, examples/pok_sig.rs:30-32).
// Group results are compared after into_affine(), so any correct group law is
// byte-identical to arkworks.
#pragma once
#include <vector>

#include "ff.hpp"

namespace orc {

template <class F>
struct Affine {
  F x, y;
  bool inf;
  static Affine infinity() { return {F::zero(), F::zero(), true}; }
  Affine neg() const { return {x, -y, inf}; }
  bool operator==(const Affine& o) const { return inf == o.inf && (inf || (x == o.x && y == o.y)); }
};

template <class F>
struct Jac {
  F x, y, z;
  static Jac infinity() { return {F::one(), F::one(), F::zero()}; }
  static Jac from_affine(const Affine<F>& a) {
    if (a.inf) return infinity();
    return {a.x, a.y, F::one()};
  }
  bool is_inf() const { return z.is_zero(); }
  Jac neg() const { return {x, -y, z}; }

  Jac dbl() const {  // dbl-2009-l, a = 0
    if (is_inf()) return *this;
    F a = x.square(), b = y.square(), c = b.square();
    F d = ((x + b).square() - a - c).dbl();
    F e = a.dbl() + a;
    F f = e.square();
    F x3 = f - d.dbl();
    F y3 = e * (d - x3) - c.dbl().dbl().dbl();
    F z3 = (y * z).dbl();
    return {x3, y3, z3};
  }
  Jac add(const Jac& o) const {  // add-2007-bl
    if (is_inf()) return o;
    if (o.is_inf()) return *this;
    F z1z1 = z.square(), z2z2 = o.z.square();
    F u1 = x * z2z2, u2 = o.x * z1z1;
    F s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2) {
      if (s1 == s2) return dbl();
      return infinity();
    }
    F h = u2 - u1;
    F i = h.dbl().square();
    F j = h * i;
    F r = (s2 - s1).dbl();
    F v = u1 * i;
    F x3 = r.square() - j - v.dbl();
    F y3 = r * (v - x3) - (s1 * j).dbl();
    F z3 = ((z + o.z).square() - z1z1 - z2z2) * h;
    return {x3, y3, z3};
  }
  Jac add_mixed(const Affine<F>& o) const {  // madd-2007-bl
    if (o.inf) return *this;
    if (is_inf()) return from_affine(o);
    F z1z1 = z.square();
    F u2 = o.x * z1z1;
    F s2 = o.y * z * z1z1;
    if (x == u2) {
      if (y == s2) return dbl();
      return infinity();
    }
    F h = u2 - x;
    F hh = h.square();
    F i = hh.dbl().dbl();
    F j = h * i;
    F r = (s2 - y).dbl();
    F v = x * i;
    F x3 = r.square() - j - v.dbl();
    F y3 = r * (v - x3) - (y * j).dbl();
    F z3 = (z + h).square() - z1z1 - hh;
    return {x3, y3, z3};
  }
  template <int M>
  Jac mul(const Big<M>& k) const {
    Jac r = infinity();
    for (int i = k.num_bits() - 1; i >= 0; i--) {
      r = r.dbl();
      if (k.bit(i)) r = r.add(*this);
    }
    return r;
  }
  Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::infinity();
    F zi = z.inverse();
    F zi2 = zi.square();
    return {x * zi2, y * zi2 * zi, false};
  }
};

// Montgomery-trick batch normalisation.
template <class F>
static std::vector<Affine<F>> batch_to_affine(const std::vector<Jac<F>>& pts) {
  size_t n = pts.size();
  std::vector<Affine<F>> out(n);
  std::vector<F> pre(n);
  F acc = F::one();
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    if (!pts[i].is_inf()) acc = acc * pts[i].z;
  }
  F inv = acc.inverse();
  for (size_t i = n; i-- > 0;) {
    if (pts[i].is_inf()) {
      out[i] = Affine<F>::infinity();
      continue;
    }
    F zi = inv * pre[i];
    inv = inv * pts[i].z;
    F zi2 = zi.square();
    out[i] = {pts[i].x * zi2, pts[i].y * zi2 * zi, false};
  }
  return out;
}

typedef Affine<Fq> G1A;
typedef Jac<Fq> G1J;
typedef Affine<Fq2> G2A;
typedef Jac<Fq2> G2J;

static inline Big<6> big6_from_hex(const char* hex) {
  Big<6> b;
  memset(b.l, 0, sizeof b.l);
  size_t n = strlen(hex);
  for (size_t i = 0; i < n; i++) {
    char ch = hex[n - 1 - i];
    uint64_t d = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
    b.l[i / 16] |= d << (4 * (i % 16));
  }
  return b;
}

// Standard BLS12-381 generators (same constants as ark-bls12-381 0.3.0 g1.rs/g2.rs).
static inline G1A g1_generator() {
  return {Fq::from_big(big6_from_hex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")),
          Fq::from_big(big6_from_hex("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1")),
          false};
}
static inline G2A g2_generator() {
  Fq2 x = {Fq::from_big(big6_from_hex("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")),
           Fq::from_big(big6_from_hex("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"))};
  Fq2 y = {Fq::from_big(big6_from_hex("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801")),
           Fq::from_big(big6_from_hex("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"))};
  return {x, y, false};
}
static inline bool g1_on_curve(const G1A& p) {
  if (p.inf) return true;
  return p.y.square() == p.x.square() * p.x + Fq::from_u64(4);
}
static inline bool g2_on_curve(const G2A& p) {
  if (p.inf) return true;
  Fq2 b = {Fq::from_u64(4), Fq::from_u64(4)};
  return p.y.square() == p.x.square() * p.x + b;
}

}  // namespace orc
