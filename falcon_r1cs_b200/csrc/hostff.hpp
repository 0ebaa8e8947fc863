// Host-side BLS12-381 base-field arithmetic (6 x u64 Montgomery limbs, the memory image of ark_ff::Fp384)
// and its quadratic extension, shared by the prover's host tail (finalize.cpp) and the verifier (pairing.cpp).
#pragma once
#include <cstdint>
#include <cstring>

#include "ff_consts.cuh"

namespace hostff {

typedef unsigned __int128 u128;

struct Fq64 {  // BLS12-381 base field, 6 x u64 limbs, Montgomery (R = 2^384): same image as 12 x u32
  uint64_t v[6];
  static constexpr uint64_t MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                      0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
  static constexpr uint64_t INV = 0x89f3fffcfffcfffdULL;  // -p^-1 mod 2^64
  static Fq64 zero() {
    Fq64 r;
    memset(r.v, 0, sizeof r.v);
    return r;
  }
  static Fq64 one() {
    Fq64 r;
    for (int i = 0; i < 6; i++) r.v[i] = (uint64_t)FqParams::R1(2 * i) | ((uint64_t)FqParams::R1(2 * i + 1) << 32);
    return r;
  }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3] | v[4] | v[5]) == 0; }
  bool operator==(const Fq64& o) const { return memcmp(v, o.v, sizeof v) == 0; }
  bool operator!=(const Fq64& o) const { return !(*this == o); }
  static bool geq_mod(const uint64_t* a) {
    for (int i = 5; i >= 0; i--) {
      if (a[i] > MOD[i]) return true;
      if (a[i] < MOD[i]) return false;
    }
    return true;
  }
  static void sub_mod(uint64_t* a) {
    uint64_t br = 0;
    for (int i = 0; i < 6; i++) {
      u128 d = (u128)a[i] - MOD[i] - br;
      a[i] = (uint64_t)d;
      br = (uint64_t)(d >> 64) & 1;
    }
  }
  friend Fq64 operator+(const Fq64& a, const Fq64& b) {
    Fq64 r;
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)a.v[i] + b.v[i];
      r.v[i] = (uint64_t)c;
      c >>= 64;
    }
    if (geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  friend Fq64 operator-(const Fq64& a, const Fq64& b) {
    Fq64 r;
    uint64_t br = 0;
    for (int i = 0; i < 6; i++) {
      u128 d = (u128)a.v[i] - b.v[i] - br;
      r.v[i] = (uint64_t)d;
      br = (uint64_t)(d >> 64) & 1;
    }
    if (br) {
      u128 c = 0;
      for (int i = 0; i < 6; i++) {
        c += (u128)r.v[i] + MOD[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    return r;
  }
  friend Fq64 operator*(const Fq64& a, const Fq64& b) {  // CIOS
    uint64_t t[8] = {0};
    for (int i = 0; i < 6; i++) {
      u128 c = 0;
      for (int j = 0; j < 6; j++) {
        c += (u128)a.v[j] * b.v[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[6];
      t[6] = (uint64_t)c;
      t[7] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * INV;
      c = ((u128)m * MOD[0] + t[0]) >> 64;
      for (int j = 1; j < 6; j++) {
        c += (u128)m * MOD[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[6];
      t[5] = (uint64_t)c;
      t[6] = t[7] + (uint64_t)(c >> 64);
    }
    Fq64 r;
    memcpy(r.v, t, sizeof r.v);
    if (t[6] || geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  Fq64 sqr() const { return *this * *this; }
  Fq64 neg() const { return zero() - *this; }
  static Fq64 mul_sub2(const Fq64& a, const Fq64& b, const Fq64& c, const Fq64& d) { return a * b - c * d; }
  Fq64 dbl() const { return *this + *this; }
  Fq64 inverse() const {  // a^(p-2)
    uint64_t e[6];
    memcpy(e, MOD, sizeof e);
    e[0] -= 2;
    Fq64 r = one(), b = *this;
    for (int i = 0; i < 384; i++) {
      if ((e[i >> 6] >> (i & 63)) & 1) r = r * b;
      b = b.sqr();
    }
    return r;
  }
  Fq64 from_mont() const {
    Fq64 o = zero();
    o.v[0] = 1;
    return *this * o;
  }
};
constexpr uint64_t Fq64::MOD[6];

struct Fq2_64 {
  Fq64 c0, c1;
  static Fq2_64 zero() { return {Fq64::zero(), Fq64::zero()}; }
  static Fq2_64 one() { return {Fq64::one(), Fq64::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2_64& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fq2_64& o) const { return !(*this == o); }
  friend Fq2_64 operator+(const Fq2_64& a, const Fq2_64& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  friend Fq2_64 operator-(const Fq2_64& a, const Fq2_64& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  Fq2_64 neg() const { return {c0.neg(), c1.neg()}; }
  static Fq2_64 mul_sub2(const Fq2_64& a, const Fq2_64& b, const Fq2_64& c, const Fq2_64& d) { return a * b - c * d; }
  Fq2_64 dbl() const { return {c0.dbl(), c1.dbl()}; }
  friend Fq2_64 operator*(const Fq2_64& a, const Fq2_64& b) {
    Fq64 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1, t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
    return {t0 - t1, t2 - t0 - t1};
  }
  Fq2_64 sqr() const {
    Fq64 a = (c0 + c1) * (c0 - c1), b = c0 * c1;
    return {a, b + b};
  }
  Fq2_64 inverse() const {
    Fq64 n = (c0.sqr() + c1.sqr()).inverse();
    return {c0 * n, (c1 * n).neg()};
  }
};


}  // namespace hostff
