// Host-side tail of create_proof (ark-groth16 0.3.0 prover.rs, reached from
// examples/pok_sig.rs:32): the O(1) group operations that follow the five MSMs,
//     C = s*A + r*B1 + (L - r s delta_1) + H,      A, B2, C -> affine,
// and ark-serialize's compressed form.  Two 255-bit variable-base scalar
// multiplications are latency-bound (a sequential chain of ~255 doublings), which a
// single CPU core with 64-bit limbs finishes in ~0.2 ms — faster than one GPU thread —
// and which overlaps with the next proof's kernels.  Everything proportional to the
// circuit size stays on the GPU.
//
// Plain C++ (compiled by g++): the curve formulas are the same ec.cuh templates the
// kernels use, instantiated over a 6 x u64 Montgomery field.
#include <cstdint>
#include <cstring>

#include "ec.cuh"
#include "finalize.hpp"
#include "hostff.hpp"

namespace {

using namespace hostff;

typedef ec::XYZZ<Fq64> G1h;
typedef ec::Affine<Fq64> G1ah;

Fq64 ld(const uint64_t* p) {
  Fq64 r;
  memcpy(r.v, p, 48);
  return r;
}
G1h ld_g1(const uint64_t* p) { return {ld(p), ld(p + 6), ld(p + 12), ld(p + 18)}; }

// r (Montgomery Fr, 4 x u64) -> canonical little-endian u32 limbs
void fr_canonical(const uint64_t* mont, uint32_t* out) {
  ff::Fr x;
  for (int i = 0; i < 4; i++) {
    x.v[2 * i] = (uint32_t)mont[i];
    x.v[2 * i + 1] = (uint32_t)(mont[i] >> 32);
  }
  x = x.from_mont();
  for (int i = 0; i < 8; i++) out[i] = x.v[i];
}

// 4-bit fixed-window scalar multiplication: 252 doublings + 64 additions
G1h scalar_mul(const G1h& p, const uint32_t* k) {
  G1h tab[16];
  tab[0] = G1h::infinity();
  tab[1] = p;
  for (int i = 2; i < 16; i++) {
    tab[i] = tab[i - 1];
    tab[i].add(p);
  }
  G1h r = G1h::infinity();
  for (int w = 63; w >= 0; w--) {
    if (w != 63)
      for (int d = 0; d < 4; d++) r = r.dbl();
    uint32_t dig = (k[w >> 3] >> ((w & 7) * 4)) & 15;
    if (dig) r.add(tab[dig]);
  }
  return r;
}

bool fq_gt(const Fq64& a, const Fq64& b) {  // on canonical integers
  Fq64 x = a.from_mont(), y = b.from_mont();
  for (int i = 5; i >= 0; i--) {
    if (x.v[i] != y.v[i]) return x.v[i] > y.v[i];
  }
  return false;
}

}  // namespace

void host_finalize_proof(const uint64_t* msm, const uint64_t* r_mont, const uint64_t* s_mont, uint64_t* proof) {
  // msm: A (24 u64, XYZZ) | B1 (24) | L (24) | H (24) | B2 (48, XYZZ over Fq2)
  uint32_t r[8], s[8];
  fr_canonical(r_mont, r);
  fr_canonical(s_mont, s);
  G1h A = ld_g1(msm), B1 = ld_g1(msm + 24), Lp = ld_g1(msm + 48), H = ld_g1(msm + 72);
  G1h C = scalar_mul(A, s);
  bool r_zero = true;
  for (int i = 0; i < 8; i++) r_zero &= r[i] == 0;
  if (!r_zero) C.add(scalar_mul(B1, r));  // g1_b is skipped when r == 0 (prover.rs)
  C.add(Lp);
  C.add(H);
  G1ah a = A.to_affine(), c = C.to_affine();
  memcpy(proof, a.x.v, 48);
  memcpy(proof + 6, a.y.v, 48);
  memcpy(proof + 36, c.x.v, 48);
  memcpy(proof + 42, c.y.v, 48);
  // B2: XYZZ over Fq2 -> affine
  const uint64_t* b = msm + 96;
  Fq2_64 X = {ld(b), ld(b + 6)}, Y = {ld(b + 12), ld(b + 18)}, ZZ = {ld(b + 24), ld(b + 30)}, ZZZ = {ld(b + 36), ld(b + 42)};
  if (ZZ.is_zero()) {
    memset(proof + 12, 0, 192);
  } else {
    Fq2_64 zi = ZZZ.inverse();
    Fq2_64 zi2 = (zi * ZZ).sqr();
    Fq2_64 x = X * zi2, y = Y * zi;
    memcpy(proof + 12, x.c0.v, 48);
    memcpy(proof + 18, x.c1.v, 48);
    memcpy(proof + 24, y.c0.v, 48);
    memcpy(proof + 30, y.c1.v, 48);
  }
}

void host_sum_partials(uint32_t n_shards, const uint64_t* const* partials, uint64_t* out) {
  typedef ec::XYZZ<Fq2_64> G2h;
  for (int k = 0; k < 4; k++) {  // A, B1, L+H, (unused) in G1
    G1h acc = G1h::infinity();
    for (uint32_t sh = 0; sh < n_shards; sh++) acc.add(ld_g1(partials[sh] + 24 * k));
    memcpy(out + 24 * k, acc.x.v, 48);
    memcpy(out + 24 * k + 6, acc.y.v, 48);
    memcpy(out + 24 * k + 12, acc.zz.v, 48);
    memcpy(out + 24 * k + 18, acc.zzz.v, 48);
  }
  G2h acc = G2h::infinity();
  for (uint32_t sh = 0; sh < n_shards; sh++) {
    const uint64_t* b = partials[sh] + 96;
    G2h p = {{ld(b), ld(b + 6)}, {ld(b + 12), ld(b + 18)}, {ld(b + 24), ld(b + 30)}, {ld(b + 36), ld(b + 42)}};
    acc.add(p);
  }
  const Fq64* f[8] = {&acc.x.c0, &acc.x.c1, &acc.y.c0, &acc.y.c1, &acc.zz.c0, &acc.zz.c1, &acc.zzz.c0, &acc.zzz.c1};
  for (int i = 0; i < 8; i++) memcpy(out + 96 + 6 * i, f[i]->v, 48);
}

// ark-serialize 0.3 compressed form: x little-endian canonical, flags
// in the top bits of the last byte: bit7 = y > -y (lexicographic; Fq2: c1 first), bit6 = infinity
static void ser_fq(const uint64_t* mont, uint8_t* out) {
  Fq64 c = ld(mont).from_mont();
  memcpy(out, c.v, 48);
}
void host_compress_proof(const uint64_t* pa, uint8_t* out) {
  memset(out, 0, 192);
  auto g1 = [&](const uint64_t* p, uint8_t* o) {
    Fq64 x = ld(p), y = ld(p + 6);
    if (x.is_zero() && y.is_zero()) {
      o[47] |= 1 << 6;
      return;
    }
    ser_fq(p, o);
    if (fq_gt(y, y.neg())) o[47] |= 1 << 7;
  };
  g1(pa, out);
  g1(pa + 36, out + 144);
  const uint64_t* b = pa + 12;
  Fq64 x0 = ld(b), x1 = ld(b + 6), y0 = ld(b + 12), y1 = ld(b + 18);
  uint8_t* o = out + 48;
  if (x0.is_zero() && x1.is_zero() && y0.is_zero() && y1.is_zero()) {
    o[95] |= 1 << 6;
  } else {
    ser_fq(b, o);
    ser_fq(b + 6, o + 48);
    Fq64 n0 = y0.neg(), n1 = y1.neg();
    bool gt = y1 != n1 ? fq_gt(y1, n1) : fq_gt(y0, n0);
    if (gt) o[95] |= 1 << 7;
  }
}
