// Host-side Groth16 verifier over BLS12-381: what ark_groth16::verify_proof(&pvk, &proof, &public_inputs)
// (examples/pok_sig.rs:45-47) computes, i.e.
//     e(A, B) == e(alpha, beta) * e(IC_0 + sum_i x_i IC_{i+1}, gamma) * e(C, delta).
// [EXT] ark-ec 0.3 / ark-bls12-381 0.3 implement the optimal ate pairing with sparse line
// functions and a cyclotomic final exponentiation.  The check only needs *a* non-degenerate
// bilinear pairing evaluated the same way on both sides, so this implementation favours
// simplicity over speed (a verification is a few tens of milliseconds on one core):
//   * Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - xi), xi = 1 + u, schoolbook/Karatsuba towers;
//   * G2 points are mapped to E(Fq12) by the untwist (x', y') -> (x'/w^2, y'/w^3) and the Miller
//     loop over |x| runs with plain affine chord-and-tangent lines in Fq12;
//   * the final exponentiation is one square-and-multiply by (p^12 - 1)/r.
// The four Miller loops are multiplied before a single final exponentiation.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/falcon_r1cs_b200.h"
#include "ec.cuh"
#include "hostff.hpp"
#include "pairing_consts.hpp"

namespace {

using namespace hostff;
typedef Fq64 Fq;
typedef Fq2_64 Fq2;

Fq2 fq2(const Fq& a, const Fq& b) { return Fq2{a, b}; }
Fq2 mul_xi(const Fq2& a) { return {a.c0 - a.c1, a.c0 + a.c1}; }  // (1 + u)(c0 + c1 u)

struct Fq6 {
  Fq2 c0, c1, c2;  // c0 + c1 v + c2 v^2
  static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
  bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  friend Fq6 operator+(const Fq6& a, const Fq6& b) { return {a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
  friend Fq6 operator-(const Fq6& a, const Fq6& b) { return {a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
  Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  friend Fq6 operator*(const Fq6& a, const Fq6& b) {
    Fq2 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1, t2 = a.c2 * b.c2;
    Fq2 r0 = t0 + mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - t1 - t2);
    Fq2 r1 = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1 + mul_xi(t2);
    Fq2 r2 = (a.c0 + a.c2) * (b.c0 + b.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  Fq6 mul_v() const { return {mul_xi(c2), c0, c1}; }  // v * (c0 + c1 v + c2 v^2), v^3 = xi
  Fq6 inverse() const {
    Fq2 A = c0.sqr() - mul_xi(c1 * c2);
    Fq2 B = mul_xi(c2.sqr()) - c0 * c1;
    Fq2 C = c1.sqr() - c0 * c2;
    Fq2 F = (c0 * A + mul_xi(c2 * B + c1 * C)).inverse();
    return {A * F, B * F, C * F};
  }
};

struct Fq12 {
  Fq6 c0, c1;  // c0 + c1 w, w^2 = v
  static Fq12 zero() { return {Fq6::zero(), Fq6::zero()}; }
  static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
  friend Fq12 operator+(const Fq12& a, const Fq12& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  friend Fq12 operator-(const Fq12& a, const Fq12& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  friend Fq12 operator*(const Fq12& a, const Fq12& b) {
    Fq6 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
    return {t0 + t1.mul_v(), (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1};
  }
  Fq12 sqr() const { return *this * *this; }
  Fq12 inverse() const {
    Fq6 d = (c0 * c0 - (c1 * c1).mul_v()).inverse();
    return {c0 * d, (c1 * d).neg()};
  }
  static Fq12 from_fq(const Fq& x) { return {{fq2(x, Fq::zero()), Fq2::zero(), Fq2::zero()}, Fq6::zero()}; }
  static Fq12 from_fq2(const Fq2& x) { return {{x, Fq2::zero(), Fq2::zero()}, Fq6::zero()}; }
  static Fq12 w() { return {Fq6::zero(), Fq6::one()}; }
};

Fq ld(const uint64_t* p) {
  Fq r;
  memcpy(r.v, p, 48);
  return r;
}
bool g1_is_inf(const uint64_t* p) {
  for (int i = 0; i < 12; i++)
    if (p[i]) return false;
  return true;
}
bool g2_is_inf(const uint64_t* p) {
  for (int i = 0; i < 24; i++)
    if (p[i]) return false;
  return true;
}

// Miller loop f_{|x|, psi(Q)}(P) with affine lines over Fq12 (vertical lines omitted)
Fq12 miller(const uint64_t* g1, const uint64_t* g2) {
  if (g1_is_inf(g1) || g2_is_inf(g2)) return Fq12::one();
  static const Fq12 wi = Fq12::w().inverse(), wi2 = wi * wi, wi3 = wi2 * wi;
  const Fq12 xp = Fq12::from_fq(ld(g1)), yp = Fq12::from_fq(ld(g1 + 6));
  const Fq12 xq = Fq12::from_fq2(fq2(ld(g2), ld(g2 + 6))) * wi2, yq = Fq12::from_fq2(fq2(ld(g2 + 12), ld(g2 + 18))) * wi3;
  Fq12 xt = xq, yt = yq, f = Fq12::one();
  const Fq12 three = Fq12::from_fq(Fq::one() + Fq::one() + Fq::one());
  int top = 63;
  while (!((BLS_X_ABS >> top) & 1)) top--;
  for (int i = top - 1; i >= 0; i--) {
    // tangent at T
    Fq12 lam = three * xt * xt * (yt + yt).inverse();
    f = f.sqr() * (yp - yt - lam * (xp - xt));
    Fq12 x3 = lam * lam - xt - xt;
    yt = lam * (xt - x3) - yt;
    xt = x3;
    if ((BLS_X_ABS >> i) & 1) {
      Fq12 lam2 = (yq - yt) * (xq - xt).inverse();
      f = f * (yp - yt - lam2 * (xp - xt));
      Fq12 x4 = lam2 * lam2 - xt - xq;
      yt = lam2 * (xt - x4) - yt;
      xt = x4;
    }
  }
  return f;
}

Fq12 final_exp(const Fq12& f) {
  Fq12 r = Fq12::one();
  bool started = false;
  for (int i = FINAL_EXP_LIMBS * 64 - 1; i >= 0; i--) {
    if (started) r = r.sqr();
    if ((FINAL_EXP[i >> 6] >> (i & 63)) & 1) {
      r = started ? r * f : f;
      started = true;
    }
  }
  return r;
}

typedef ec::XYZZ<Fq> G1h;
typedef ec::Affine<Fq> G1ah;

// canonical little-endian u32 limbs of a Montgomery Fr (4 x u64)
void fr_canonical(const uint64_t* mont, uint32_t* out) {
  ff::Fr x;
  for (int i = 0; i < 4; i++) {
    x.v[2 * i] = (uint32_t)mont[i];
    x.v[2 * i + 1] = (uint32_t)(mont[i] >> 32);
  }
  x = x.from_mont();
  for (int i = 0; i < 8; i++) out[i] = x.v[i];
}

// ---- point validation: what arkworks does when it deserialises a Proof / VerifyingKey ([EXT] ark-ec 0.3
// GroupAffine::deserialize: coordinates below p, on the curve, in the prime-order subgroup).  This ABI takes raw
// affine limbs, so the verifier checks them itself.  (0, 0) is the encoding of the point at infinity.
typedef ec::XYZZ<Fq2> G2h;
typedef ec::Affine<Fq2> G2ah;
const uint32_t R_LIMBS[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};

bool limbs_in_range(const uint64_t* p, int n_fq) {
  for (int i = 0; i < n_fq; i++)
    if (Fq::geq_mod(p + 6 * i)) return false;
  return true;
}
// 1: infinity or a point of the order-r subgroup; 0: coordinates out of range, off the curve or outside the subgroup
int g1_valid(const uint64_t* p) {
  if (g1_is_inf(p)) return 1;
  if (!limbs_in_range(p, 2)) return 0;
  const Fq x = ld(p), y = ld(p + 6), one = Fq::one(), four = (one + one).dbl();
  if (y.sqr() != x.sqr() * x + four) return 0;  // y^2 = x^3 + 4
  return G1h::from_affine(G1ah{x, y}).mul(R_LIMBS, 255).is_inf() ? 1 : 0;
}
int g2_valid(const uint64_t* p) {
  if (g2_is_inf(p)) return 1;
  if (!limbs_in_range(p, 4)) return 0;
  const Fq2 x = fq2(ld(p), ld(p + 6)), y = fq2(ld(p + 12), ld(p + 18));
  const Fq one = Fq::one(), four = (one + one).dbl();
  if (y.sqr() != x.sqr() * x + fq2(four, four)) return 0;  // y^2 = x^3 + 4 (1 + u)
  return G2h::from_affine(G2ah{x, y}).mul(R_LIMBS, 255).is_inf() ? 1 : 0;
}

}  // namespace

extern "C" {

// 1 = infinity or a valid point of the prime-order subgroup, 0 = not (range, curve equation or subgroup check failed)
int32_t frcs_g1_validate(const uint64_t* p) { return p ? g1_valid(p) : FRCS_E_INVALID_ARG; }
int32_t frcs_g2_validate(const uint64_t* p) { return p ? g2_valid(p) : FRCS_E_INVALID_ARG; }
// all points of a verifying key (done once per key; frcs_verify_proof re-checks only the proof's own points)
int32_t frcs_vk_validate(const uint64_t* vk_alpha_g1, const uint64_t* vk_g2, const uint64_t* gamma_abc_g1, uint64_t n_inputs) {
  if (!vk_alpha_g1 || !vk_g2 || !gamma_abc_g1) return FRCS_E_INVALID_ARG;
  if (!g1_valid(vk_alpha_g1)) return 0;
  for (int i = 0; i < 3; i++)
    if (!g2_valid(vk_g2 + 24 * i) || g2_is_inf(vk_g2 + 24 * i)) return 0;
  for (uint64_t i = 0; i <= n_inputs; i++)
    if (!g1_valid(gamma_abc_g1 + 12 * i)) return 0;
  return 1;
}

// e(p1, q1) == e(p2, q2)?  (G1 affine 12 u64, G2 affine 24 u64) -- exposed for the bilinearity tests
int32_t frcs_pairing_eq(const uint64_t* p1, const uint64_t* q1, const uint64_t* p2, const uint64_t* q2) {
  if (!p1 || !q1 || !p2 || !q2) return FRCS_E_INVALID_ARG;
  // e(p1,q1) * e(-p2,q2) == 1
  uint64_t neg[12];
  memcpy(neg, p2, 96);
  Fq y = ld(p2 + 6).neg();
  if (!g1_is_inf(p2)) memcpy(neg + 6, y.v, 48);
  Fq12 f = miller(p1, q1) * miller(neg, q2);
  return final_exp(f) == Fq12::one() ? 1 : 0;
}

// e(p, q) == 1?  (degeneracy test)
int32_t frcs_pairing_is_one(const uint64_t* p, const uint64_t* q) {
  if (!p || !q) return FRCS_E_INVALID_ARG;
  return final_exp(miller(p, q)) == Fq12::one() ? 1 : 0;
}

// ark_groth16::verify_proof: 1 = valid, 0 = invalid, negative = error.
// vk_alpha_g1 (12), vk_g2 = beta_g2 | gamma_g2 | delta_g2 (3 x 24), gamma_abc_g1 ((n_inputs + 1) x 12),
// proof = A (12) | B (24) | C (12) affine, public_inputs: n_inputs x 4 Montgomery Fr (without the leading One).
int32_t frcs_verify_proof(const uint64_t* vk_alpha_g1, const uint64_t* vk_g2, const uint64_t* gamma_abc_g1,
                          uint64_t n_inputs, const uint64_t* public_inputs, const uint64_t* proof) {
  if (!vk_alpha_g1 || !vk_g2 || !gamma_abc_g1 || !proof || (n_inputs && !public_inputs)) return FRCS_E_INVALID_ARG;
  // The proof comes from an untrusted party: its three points must be valid group elements before they enter the
  // pairing (arkworks checks this when it deserialises the Proof).  The verifying key is checked once with
  // frcs_vk_validate; here only the coordinate ranges of the points that go into the Miller loops.
  if (!g1_valid(proof) || !g2_valid(proof + 12) || !g1_valid(proof + 36)) return FRCS_E_INVALID_POINT;
  if (!limbs_in_range(vk_alpha_g1, 2) || !limbs_in_range(vk_g2, 12)) return FRCS_E_INVALID_POINT;
  for (uint64_t i = 0; i < n_inputs; i++) {  // public inputs must be canonical field elements
    bool lt = false;
    for (int k = 3; k >= 0 && !lt; k--) {
      const uint64_t rk = (uint64_t)R_LIMBS[2 * k] | ((uint64_t)R_LIMBS[2 * k + 1] << 32);
      if (public_inputs[4 * i + k] > rk) return FRCS_E_INVALID_ARG;
      lt = public_inputs[4 * i + k] < rk;
    }
    if (!lt) return FRCS_E_INVALID_ARG;
  }
  // prepare_inputs: IC_0 + sum x_i IC_{i+1}
  G1h acc = G1h::from_affine(G1ah{ld(gamma_abc_g1), ld(gamma_abc_g1 + 6)});
  for (uint64_t i = 0; i < n_inputs; i++) {
    const uint64_t* b = gamma_abc_g1 + 12 * (i + 1);
    if (g1_is_inf(b)) continue;
    uint32_t k[8];
    fr_canonical(public_inputs + 4 * i, k);
    int nbits = 256;
    while (nbits > 0 && !((k[(nbits - 1) >> 5] >> ((nbits - 1) & 31)) & 1)) nbits--;
    if (nbits == 0) continue;
    acc.add(G1h::from_affine(G1ah{ld(b), ld(b + 6)}).mul(k, nbits));
  }
  G1ah vkx = acc.to_affine();
  uint64_t nx[12], nc[12], na[12];
  auto neg_g1 = [](const uint64_t* p, uint64_t* out) {
    memcpy(out, p, 96);
    if (!g1_is_inf(p)) {
      Fq y = ld(p + 6).neg();
      memcpy(out + 6, y.v, 48);
    }
  };
  uint64_t vkx_raw[12];
  memcpy(vkx_raw, vkx.x.v, 48);
  memcpy(vkx_raw + 6, vkx.y.v, 48);
  neg_g1(vkx_raw, nx);
  neg_g1(proof + 36, nc);
  neg_g1(vk_alpha_g1, na);
  // e(A, B) * e(-vk_x, gamma) * e(-C, delta) * e(-alpha, beta) == 1
  Fq12 f = miller(proof, proof + 12) * miller(nx, vk_g2 + 24) * miller(nc, vk_g2 + 48) * miller(na, vk_g2);
  return final_exp(f) == Fq12::one() ? 1 : 0;
}

}  // extern "C"
