// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of ark-ff 0.3.0's prime
// fields (Fp256 / Fp384, 64-bit limbs, Montgomery form), which are [EXT] to
// /root/reference (crates.io dependency `ark-ff ^0.3.0`, falcon-r1cs/Cargo.toml:15).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
//
// Parity note: ark-ff's source is not in /root/reference; representation facts
// restated here: little-endian u64 limbs, value stored as
// x*R mod m with R = 2^(64*N).  Results of field arithmetic are exact, hence
// algorithm-independent.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace orc {

typedef unsigned __int128 u128;

template <int N>
struct Big {
  uint64_t l[N];
  bool operator==(const Big& o) const { return memcmp(l, o.l, sizeof l) == 0; }
  bool is_zero() const {
    for (int i = 0; i < N; i++)
      if (l[i]) return false;
    return true;
  }
  bool bit(int i) const { return i < 64 * N && ((l[i / 64] >> (i % 64)) & 1); }
  int num_bits() const {
    for (int i = N - 1; i >= 0; i--)
      if (l[i]) return 64 * i + 64 - __builtin_clzll(l[i]);
    return 0;
  }
};

template <int N>
static inline int cmp_n(const uint64_t* a, const uint64_t* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}
template <int N>
static inline uint64_t add_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 c = 0;
  for (int i = 0; i < N; i++) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
template <int N>
static inline uint64_t sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < N; i++) {
    u128 d = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}

// Field parameters derived at start-up from the modulus alone.
template <int N>
struct FieldParams {
  uint64_t mod[N];
  uint64_t r1[N];  // R mod m  (Montgomery one)
  uint64_t r2[N];  // R^2 mod m
  uint64_t inv;    // -m^{-1} mod 2^64
  Big<N> mod_minus_2;
  void init(const uint64_t* m) {
    memcpy(mod, m, sizeof mod);
    uint64_t x = 1;  // Newton iteration for m^{-1} mod 2^64
    for (int i = 0; i < 6; i++) x *= 2 - m[0] * x;
    inv = (uint64_t)0 - x;
    // r1 = 2^(64N) mod m by repeated doubling of 1
    uint64_t t[N] = {0};
    t[0] = 1;
    auto dbl = [&](uint64_t* a) {
      uint64_t c = add_n<N>(a, a, a);
      if (c || cmp_n<N>(a, mod) >= 0) sub_n<N>(a, a, mod);
    };
    for (int i = 0; i < 64 * N; i++) dbl(t);
    memcpy(r1, t, sizeof r1);
    for (int i = 0; i < 64 * N; i++) dbl(t);
    memcpy(r2, t, sizeof r2);
    uint64_t two[N] = {0};
    two[0] = 2;
    sub_n<N>(mod_minus_2.l, mod, two);
  }
};

// Tag must provide: static constexpr int N; static const uint64_t MOD[N].
template <class Tag>
struct Fp {
  static constexpr int N = Tag::N;
  uint64_t v[N];  // Montgomery form, same memory image as ark-ff's BigInteger limbs

  static const FieldParams<N>& P() {
    static FieldParams<N> p = [] {
      FieldParams<N> q;
      q.init(Tag::MOD);
      return q;
    }();
    return p;
  }
  static Fp zero() {
    Fp r;
    memset(r.v, 0, sizeof r.v);
    return r;
  }
  static Fp one() {
    Fp r;
    memcpy(r.v, P().r1, sizeof r.v);
    return r;
  }
  static Fp from_raw(const uint64_t* limbs) {  // limbs already Montgomery
    Fp r;
    memcpy(r.v, limbs, sizeof r.v);
    return r;
  }
  static Fp from_big(const Big<N>& b) {  // canonical integer < m
    Fp r;
    memcpy(r.v, b.l, sizeof r.v);
    Fp r2 = from_raw(P().r2);
    return r * r2;
  }
  static Fp from_u64(uint64_t x) {
    Big<N> b;
    memset(b.l, 0, sizeof b.l);
    b.l[0] = x;
    return from_big(b);
  }
  static Fp from_i64(int64_t x) { return x >= 0 ? from_u64((uint64_t)x) : -from_u64((uint64_t)(-x)); }
  Big<N> to_big() const {  // into_repr(): canonical integer
    Fp o;
    memset(o.v, 0, sizeof o.v);
    o.v[0] = 1;
    Fp r = mont_mul(*this, o);
    Big<N> b;
    memcpy(b.l, r.v, sizeof b.l);
    return b;
  }
  bool is_zero() const {
    for (int i = 0; i < N; i++)
      if (v[i]) return false;
    return true;
  }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof v) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  static Fp mont_mul(const Fp& a, const Fp& b) {  // CIOS
    const FieldParams<N>& p = P();
    uint64_t t[N + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < N; i++) {
      u128 c = 0;
      for (int j = 0; j < N; j++) {
        c += (u128)a.v[j] * b.v[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[N];
      t[N] = (uint64_t)c;
      t[N + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * p.inv;
      c = (u128)m * p.mod[0] + t[0];
      c >>= 64;
      for (int j = 1; j < N; j++) {
        c += (u128)m * p.mod[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[N];
      t[N - 1] = (uint64_t)c;
      t[N] = t[N + 1] + (uint64_t)(c >> 64);
    }
    Fp r;
    if (t[N] || cmp_n<N>(t, p.mod) >= 0)
      sub_n<N>(r.v, t, p.mod);
    else
      memcpy(r.v, t, sizeof r.v);
    return r;
  }
  Fp operator*(const Fp& o) const { return mont_mul(*this, o); }
  Fp& operator*=(const Fp& o) { return *this = mont_mul(*this, o); }
  Fp square() const { return mont_mul(*this, *this); }
  Fp operator+(const Fp& o) const {
    Fp r;
    uint64_t c = add_n<N>(r.v, v, o.v);
    if (c || cmp_n<N>(r.v, P().mod) >= 0) sub_n<N>(r.v, r.v, P().mod);
    return r;
  }
  Fp& operator+=(const Fp& o) { return *this = *this + o; }
  Fp operator-(const Fp& o) const {
    Fp r;
    if (sub_n<N>(r.v, v, o.v)) add_n<N>(r.v, r.v, P().mod);
    return r;
  }
  Fp& operator-=(const Fp& o) { return *this = *this - o; }
  Fp operator-() const { return zero() - *this; }
  Fp dbl() const { return *this + *this; }
  template <int M>
  Fp pow(const Big<M>& e) const {
    Fp r = one();
    for (int i = e.num_bits() - 1; i >= 0; i--) {
      r = r.square();
      if (e.bit(i)) r *= *this;
    }
    return r;
  }
  Fp pow_u64(uint64_t e) const {
    Big<1> b;
    b.l[0] = e;
    return pow(b);
  }
  Fp inverse() const { return pow(P().mod_minus_2); }  // 0 -> 0
};

struct FrTag {
  static constexpr int N = 4;
  static constexpr uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                      0x73eda753299d7d48ULL};
};
struct FqTag {
  static constexpr int N = 6;
  static constexpr uint64_t MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                      0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
};
typedef Fp<FrTag> Fr;  // BLS12-381 scalar field (= ark_ed_on_bls12_381::Fq, falcon_ntt.rs:130)
typedef Fp<FqTag> Fq;  // BLS12-381 base field

// Quadratic extension Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
  Fq c0, c1;
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fq2& o) const { return !(*this == o); }
  Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fq2 operator-() const { return {-c0, -c1}; }
  Fq2& operator+=(const Fq2& o) { return *this = *this + o; }
  Fq2& operator-=(const Fq2& o) { return *this = *this - o; }
  Fq2 operator*(const Fq2& o) const {
    Fq a = c0 * o.c0, b = c1 * o.c1;
    Fq c = (c0 + c1) * (o.c0 + o.c1);
    return {a - b, c - a - b};
  }
  Fq2& operator*=(const Fq2& o) { return *this = *this * o; }
  Fq2 square() const {
    Fq a = (c0 + c1) * (c0 - c1);
    Fq b = c0 * c1;
    return {a, b + b};
  }
  Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fq2 mul_fq(const Fq& s) const { return {c0 * s, c1 * s}; }
  Fq2 conj() const { return {c0, -c1}; }
  Fq2 inverse() const {
    Fq n = (c0.square() + c1.square()).inverse();
    return {c0 * n, -(c1 * n)};
  }
  Fq2 mul_by_nonresidue() const {  // times (1+u)
    return {c0 - c1, c0 + c1};
  }
};

}  // namespace orc
