// ORACLE — TEST INFRASTRUCTURE ONLY.  Clear-text Falcon helpers that the
// reference takes from falcon-rust ([EXT]: floating git dependency,
// falcon-r1cs/Cargo.toml:11): MODULUS, N, LOG_N, SIG_L2_BOUND, NTT_TABLE,
// Polynomial mul/sub, NTTPolynomial::from.  Restated from SURVEY.md App. D/E;
// NTT_TABLE is pinned against script/ntt_param.sage:3-132 by
// tests/golden/ntt_table.json.
#pragma once
#include <cstdint>
#include <vector>

namespace orc {

static const uint32_t FALCON_Q = 12289;
static inline uint32_t sig_l2_bound(int logn) {  // range_proofs.rs:104,196
  return logn == 9 ? 34034726u : 70265242u;
}

static inline uint32_t powmod_q(uint32_t b, uint32_t e) {
  uint64_t r = 1, x = b % FALCON_Q;
  while (e) {
    if (e & 1) r = r * x % FALCON_Q;
    x = x * x % FALCON_Q;
    e >>= 1;
  }
  return (uint32_t)r;
}
static inline uint32_t bitrev(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
// NTT_TABLE[i] = 7^bitrev10(i) mod q (script/ntt_param.sage:3-132: vrfy.c GMb / 4091);
// Falcon-512 uses the first 512 entries (gadgets/misc.rs:72).
static inline std::vector<uint32_t> ntt_table(int n) {
  std::vector<uint32_t> t(n);
  for (int i = 0; i < n; i++) t[i] = powmod_q(7, bitrev(i, 10));
  return t;
}
// NTTPolynomial::from(&Polynomial): loop shape of gadgets/poly.rs:115-149 with
// reduction mod q at each step.
static inline std::vector<uint32_t> ntt_clear(const std::vector<uint32_t>& in, int logn) {
  int n = 1 << logn;
  std::vector<uint32_t> tab = ntt_table(n), out(in);
  int t = n;
  for (int l = 0; l < logn; l++) {
    int m = 1 << l, ht = t / 2, j1 = 0;
    for (int i = 0; i < m; i++) {
      uint32_t s = tab[m + i];
      for (int j = j1; j < j1 + ht; j++) {
        uint32_t u = out[j], v = out[j + ht] * s % FALCON_Q;
        out[j] = (u + v) % FALCON_Q;
        out[j + ht] = (u + FALCON_Q - v) % FALCON_Q;
      }
      j1 += t;
    }
    t = ht;
  }
  return out;
}
// Polynomial * Polynomial in Z_q[x]/(x^N+1), schoolbook
static inline std::vector<uint32_t> poly_mul(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
  size_t n = a.size();
  std::vector<int64_t> acc(n, 0);
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < n; j++) {
      int64_t p = (int64_t)a[i] * b[j];
      if (i + j < n)
        acc[i + j] += p;
      else
        acc[i + j - n] -= p;
    }
  std::vector<uint32_t> r(n);
  for (size_t i = 0; i < n; i++) {
    int64_t x = acc[i] % (int64_t)FALCON_Q;
    if (x < 0) x += FALCON_Q;
    r[i] = (uint32_t)x;
  }
  return r;
}
// hm - uh lifted to [0, q)   (falcon_ntt.rs:47-49)
static inline std::vector<uint32_t> poly_sub(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
  std::vector<uint32_t> r(a.size());
  for (size_t i = 0; i < a.size(); i++) r[i] = (a[i] + FALCON_Q - b[i]) % FALCON_Q;
  return r;
}

}  // namespace orc
