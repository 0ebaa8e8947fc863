// Device helpers shared by the witness kernels (witness.cu) and the stand-alone gadget kernels (gadgets.cu):
// 32-byte stores of Fr values and Boolean witnesses, the Montgomery look-up of small integers, and the closed forms
// of the range-proof gadgets' Boolean witnesses.
#pragma once
#include "ff32.cuh"

namespace wdev {

using ff::Fr;

constexpr uint32_t Q = 12289;
__device__ __forceinline__ uint32_t modq(uint32_t x) { return x % Q; }

__device__ __forceinline__ void store_fr(uint64_t* dst, const Fr& x) {
  uint64_t a = (uint64_t)x.v[0] | ((uint64_t)x.v[1] << 32), b = (uint64_t)x.v[2] | ((uint64_t)x.v[3] << 32);
  uint64_t c = (uint64_t)x.v[4] | ((uint64_t)x.v[5] << 32), d = (uint64_t)x.v[6] | ((uint64_t)x.v[7] << 32);
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void store_bit(uint64_t* dst, bool bit) {
  // 0 or R mod r (Montgomery one)
  uint64_t m = bit ? ~0ull : 0ull;
  uint64_t a = ((uint64_t)FrParams::R1(0) | ((uint64_t)FrParams::R1(1) << 32)) & m;
  uint64_t b = ((uint64_t)FrParams::R1(2) | ((uint64_t)FrParams::R1(3) << 32)) & m;
  uint64_t c = ((uint64_t)FrParams::R1(4) | ((uint64_t)FrParams::R1(5) << 32)) & m;
  uint64_t d = ((uint64_t)FrParams::R1(6) | ((uint64_t)FrParams::R1(7) << 32)) & m;
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// bit j of the 27-witness enforce_less_than_q gadget on value x (range_proofs.rs:42-94):
// b0..b13, o1..o11 (o_k = b0|..|b_k), x1 = o11 & b12, x2 = x1 & b13
__device__ __forceinline__ bool ltq_bit(uint32_t x, uint32_t j) {
  if (j < 14) return (x >> j) & 1;
  if (j < 25) return (x & ((2u << (j - 13)) - 1)) != 0;
  bool x1 = ((x & 0xfffu) != 0) && ((x >> 12) & 1);
  if (j == 25) return x1;
  return x1 && ((x >> 13) & 1);
}

// mont(x) for x < 2^28 from two tables: T0[x mod 2^14] + T1[x >> 14], T1[j] = mont(2^14 j)
__device__ __forceinline__ Fr ld_tab(const uint32_t* p) {
  Fr r;
  uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
// x < 2^28
__device__ __forceinline__ Fr mont_small(const uint32_t* __restrict__ tab, uint32_t x) {
  Fr lo = ld_tab(tab + 8 * (x & 0x3fffu));
  if (x < 16384u) return lo;
  return lo + ld_tab(tab + 8 * (16384u + (x >> 14)));
}

// all 27 witnesses of enforce_less_than_q(x) at once (bit j = ltq_bit(x, j)), branch-free
__device__ __forceinline__ uint32_t ltq_mask(uint32_t x) {
  const uint32_t low12 = x & 0xfffu;
  const uint32_t p = __ffs(low12 | 0x1000u) - 1;          // lowest set bit of b_0..b_11 (12 if none)
  const uint32_t o = ((0xfffu << p) & 0xffeu) << 13;      // o_k = b_0 | .. | b_k, k = 1..11 -> bits 14..24
  const uint32_t x1 = (low12 != 0) & (x >> 12) & 1u;      // o_11 & b_12
  return (x & 0x3fffu) | o | (x1 << 25) | ((x1 & (x >> 13)) << 26);
}
// the 16 boolean witnesses of one l2 element: 14 bits, y1 = b11 & b12, y2 = !b13 & !y1
__device__ __forceinline__ uint32_t l2_mask(uint32_t e) {
  const uint32_t y1 = (e >> 11) & (e >> 12) & 1u;
  return (e & 0x3fffu) | (y1 << 14) | (((((e >> 13) | y1) & 1u) ^ 1u) << 15);
}
}  // namespace wdev
