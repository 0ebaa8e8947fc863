// 32-bit-limb Montgomery prime fields for sm_100a (and, for set-up code and unit
// tests, the host).  Memory image is identical to ark-ff 0.3.0's Fp256/Fp384
// (little-endian limbs of x*R mod m, R = 2^256 / 2^384), so
// values cross the C ABI without conversion.
//
// Multiplication is the even/odd column scheme: products a[j]*b_i with even j are
// accumulated into `even`, odd j into `odd` (which sits one limb higher), so each
// row is a single carry chain of mad.lo.cc/madc.hi.cc pairs that ptxas fuses into
// IMAD.WIDE.U32(.X).  One Montgomery reduction step per b_i, interleaved.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#ifdef __CUDA_ARCH__
#define FF_HD __host__ __device__ __forceinline__
#else
#define FF_HD __host__ __device__ inline
#endif
#define FF_NOINLINE __host__ __device__ __noinline__
#else
#define FF_HD inline
#define FF_NOINLINE __attribute__((noinline))
#endif

#include "ff_consts.cuh"

namespace ff {

#if defined(__CUDACC__) && defined(FF_OPAQUE_M0)
// Fr has -m^-1 mod 2^32 = 0xffffffff.  Knowing that, ptxas turns mi = t * M0 into a negation and then emits every
// mi * m[j] row as IMAD.X + IMAD.HI.U32.X pairs (6 issue cycles of the multiplier pipe per limb product) instead of
// IMAD.WIDE.U32.X (4): 37 % of the limb products of an Fr multiplication.  With M0 read from constant memory the
// rows stay on IMAD.WIDE (one extra 32-bit IMAD per step).  Defined by the translation units whose Fr
// multiplications are hot (ntt.cu).
static __constant__ uint32_t ff_m0_c[2] = {FrParams::M0, FqParams::M0};
#define FF_M0(P) (P::N == 8 ? ff_m0_c[0] : P::M0)
#else
#define FF_M0(P) (P::M0)
#endif

// ---- carry-chain building blocks ---------------------------------------------------
// On the device the carry lives in the PTX condition code between consecutive asm
// statements of one chain; on the host it is the explicit `cf` argument.

template <int N>
FF_HD void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < N; j += 2) {
#ifdef __CUDA_ARCH__
    asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(bi));
#else
    uint64_t p = (uint64_t)a[j] * bi;
    acc[j] = (uint32_t)p;
    acc[j + 1] = (uint32_t)(p >> 32);
#endif
  }
}

// acc += sum_{j even} a[j]*bi*2^(32j); carry out left in CC (device) / cf (host)
template <int N>
FF_HD void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi, uint32_t& cf) {
#ifdef __CUDA_ARCH__
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(acc[0]), "+r"(acc[1])
               : "r"(a[0]), "r"(bi));
#pragma unroll
  for (int j = 2; j < N; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(acc[j]), "+r"(acc[j + 1])
                 : "r"(a[j]), "r"(bi));
#else
  uint64_t c = 0;
  for (int j = 0; j < N; j += 2) {
    uint64_t p = (uint64_t)a[j] * bi;
    uint64_t lo = (uint64_t)acc[j] + (uint32_t)p + c;
    acc[j] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[j + 1] + (p >> 32) + (lo >> 32);
    acc[j + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  cf = (uint32_t)c;
#endif
}

// (odd[j],odd[j+1]) = a[j]*bi + (odd[j+2],odd[j+3]) + carry, j even < N-2;
// (odd[N-2],odd[N-1]) = a[N-2]*bi + carry.  Consumes the incoming carry.
template <int N>
FF_HD void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi, uint32_t cf) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int j = 0; j < N - 2; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(odd[j]), "=r"(odd[j + 1])
                 : "r"(a[j]), "r"(bi), "r"(odd[j + 2]), "r"(odd[j + 3]));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;"
               : "=r"(odd[N - 2]), "=r"(odd[N - 1])
               : "r"(a[N - 2]), "r"(bi));
#else
  uint64_t c = cf;
  for (int j = 0; j < N; j += 2) {
    uint64_t p = (uint64_t)a[j] * bi;
    uint64_t lo = (uint64_t)(uint32_t)p + (j < N - 2 ? odd[j + 2] : 0) + c;
    uint64_t hi = (p >> 32) + (j < N - 2 ? odd[j + 3] : 0) + (lo >> 32);
    odd[j] = (uint32_t)lo;
    odd[j + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
#endif
}

FF_HD uint32_t add_cc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a + b;
  cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
#endif
}
FF_HD uint32_t addc_cc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a + b + cf;
  cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
#endif
}
FF_HD uint32_t addc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  return a + b + cf;
#endif
}
FF_HD uint32_t sub_cc(uint32_t a, uint32_t b, uint32_t& bf) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a - b;
  bf = (uint32_t)(t >> 63);
  return (uint32_t)t;
#endif
}
FF_HD uint32_t subc_cc(uint32_t a, uint32_t b, uint32_t& bf) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  uint64_t t = (uint64_t)a - b - bf;
  bf = (uint32_t)(t >> 63);
  return (uint32_t)t;
#endif
}
// returns 0 or 0xffffffff: the final borrow spread over a word
FF_HD uint32_t subc_mask(uint32_t& bf) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(r));
  return r;
#else
  return bf ? 0xffffffffu : 0u;
#endif
}

// One interleaved multiply+reduce step.  E = the array whose limb 0 is the current
// lowest limb, O = the other one (holding limb k+1 at index k before the call,
// i.e. the previous step's E shifted: O[0] is the zeroed limb, O[1] is limb 0).
template <int N, class P>
FF_HD void mad_redc_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, const uint32_t* m, bool first) {
  uint32_t cf = 0;
  if (first) {
    mul_n<N>(O, a + 1, bi);
    mul_n<N>(E, a, bi);
  } else {
    E[0] = add_cc(E[0], O[1], cf);
    madc_n_rshift<N>(O, a + 1, bi, cf);
    cmad_n<N>(E, a, bi, cf);
    O[N - 1] = addc(O[N - 1], 0, cf);
  }
#ifdef __CUDA_ARCH__
  uint32_t mi = E[0] * FF_M0(P);
#else
  uint32_t mi = E[0] * P::M0;
#endif
  cmad_n<N>(O, m + 1, mi, cf);
  cmad_n<N>(E, m, mi, cf);
  O[N - 1] = addc(O[N - 1], 0, cf);
}

template <class P>
struct Fp {
  static constexpr int N = P::N;
  uint32_t v[N];

  FF_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  FF_HD static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::R1(i);
    return r;
  }
  FF_HD static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::R2(i);
    return r;
  }
  FF_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i];
    return o == 0;
  }
  FF_HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  FF_HD bool operator!=(const Fp& b) const { return !(*this == b); }

  // r = r - m if r >= m  (r < 2m assumed)
  FF_HD void reduce_once() {
    uint32_t t[N], bf = 0;
    t[0] = sub_cc(v[0], P::MOD(0), bf);
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = subc_cc(v[i], P::MOD(i), bf);
    uint32_t mask = subc_mask(bf);  // all ones if v < m
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = (v[i] & mask) | (t[i] & ~mask);
  }
  FF_HD friend Fp operator+(const Fp& a, const Fp& b) {
    Fp r;
    uint32_t cf = 0;
    r.v[0] = add_cc(a.v[0], b.v[0], cf);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i], cf);
    r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1], cf);
    r.reduce_once();
    return r;
  }
  FF_HD friend Fp operator-(const Fp& a, const Fp& b) {
    Fp r;
    uint32_t bf = 0;
    r.v[0] = sub_cc(a.v[0], b.v[0], bf);
#pragma unroll
    for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i], bf);
    uint32_t mask = subc_mask(bf);  // all ones if a < b
    uint32_t cf = 0;
    r.v[0] = add_cc(r.v[0], P::MOD(0) & mask, cf);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], P::MOD(i) & mask, cf);
    r.v[N - 1] = addc(r.v[N - 1], P::MOD(N - 1) & mask, cf);
    return r;
  }
  FF_HD Fp neg() const { return zero() - *this; }
  FF_HD Fp dbl() const { return *this + *this; }

  // r = a*b*R^-1 mod m
  FF_HD static void mul_inline(Fp& r, const Fp& a, const Fp& b) {
    uint32_t even[N], odd[N], m[N];
#pragma unroll
    for (int i = 0; i < N; i++) m[i] = P::MOD(i);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      mad_redc_step<N, P>(even, odd, a.v, b.v[i], m, i == 0);
      mad_redc_step<N, P>(odd, even, a.v, b.v[i + 1], m, false);
    }
    // value = even + (odd >> 32)
    uint32_t cf = 0;
    r.v[0] = add_cc(even[0], odd[1], cf);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(even[i], odd[i + 1], cf);
    r.v[N - 1] = addc(even[N - 1], 0, cf);
    r.reduce_once();
  }
  // r = (a*b - c*d) * R^-1 mod m with ONE Montgomery reduction for the two products: every step adds the row a * b[i]
  // and the row (m - c) * d[i] before it clears the lowest limb, N^2 limb products fewer than two multiplications and a
  // subtraction (Fq: 432 instead of 576).  Needs 3 m < 2^(32 N), so that the running value (< 3 m) fits the N limbs:
  // Fq (m ~ 0.10 * 2^384), not Fr.  The result before the final conditional subtraction is
  // (a b + (m - c) d + sum q_i m 2^(32 i)) / R < m (1 + 2 m / R) < 2 m.
  FF_HD static void mul_sub2_inline(Fp& r, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    static_assert(P::MOD(N - 1) < 0x40000000u, "mul_sub2 needs 3 m < 2^(32 N)");
    uint32_t even[N], odd[N], m[N], cn[N];
    {
      uint32_t bf = 0;  // cn = m - c, in (0, m]
      cn[0] = sub_cc(P::MOD(0), c.v[0], bf);
#pragma unroll
      for (int i = 1; i < N; i++) cn[i] = subc_cc(P::MOD(i), c.v[i], bf);
    }
#pragma unroll
    for (int i = 0; i < N; i++) m[i] = P::MOD(i);
#pragma unroll
    for (int i = 0; i < N; i++) {
      uint32_t* E = (i & 1) ? odd : even;
      uint32_t* O = (i & 1) ? even : odd;
      uint32_t cf = 0;
      if (i == 0) {
        mul_n<N>(O, a.v + 1, b.v[0]);
        mul_n<N>(E, a.v, b.v[0]);
      } else {
        E[0] = add_cc(E[0], O[1], cf);
        madc_n_rshift<N>(O, a.v + 1, b.v[i], cf);
        cmad_n<N>(E, a.v, b.v[i], cf);
        O[N - 1] = addc(O[N - 1], 0, cf);
      }
      cmad_n<N>(O, cn + 1, d.v[i], cf);
      cmad_n<N>(E, cn, d.v[i], cf);
      O[N - 1] = addc(O[N - 1], 0, cf);
      const uint32_t mi = E[0] * P::M0;
      cmad_n<N>(O, m + 1, mi, cf);
      cmad_n<N>(E, m, mi, cf);
      O[N - 1] = addc(O[N - 1], 0, cf);
    }
    uint32_t cf = 0;
    r.v[0] = add_cc(even[0], odd[1], cf);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(even[i], odd[i + 1], cf);
    r.v[N - 1] = addc(even[N - 1], 0, cf);
    r.reduce_once();
  }
  // a*b - c*d; the fused form where the modulus leaves room for it and multiplications are inlined (the MSM kernels)
  FF_HD static Fp mul_sub2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
#if defined(FF_INLINE_MUL) && defined(__CUDA_ARCH__) && !defined(FF_NO_MUL_SUB2)
    if constexpr (P::MOD(N - 1) < 0x40000000u) {
      Fp r;
      mul_sub2_inline(r, a, b, c, d);
      return r;
    } else
#endif
      return a * b - c * d;
  }
  // Out-of-line copy: keeps code size and compile time sane where a multiplication is
  // not on a hot path (curve formulas for G2, inversions, host set-up code).
  FF_NOINLINE static Fp mul_call(Fp a, Fp b) {  // by value: operands and result travel in registers
    Fp r;
    mul_inline(r, a, b);
    return r;
  }
  FF_HD friend Fp operator*(const Fp& a, const Fp& b) {
#if defined(FF_INLINE_MUL) && defined(__CUDA_ARCH__)
    Fp r;
    mul_inline(r, a, b);
    return r;
#else
    return mul_call(a, b);
#endif
  }
#if defined(__CUDA_ARCH__)
  // one carry chain of the wide square: t[i + j], t[i + j + 1] += v[i] * v[j] for j = j0, j0 + 2, ... < N; the carry out
  // lands on the limb above the chain (zero or at most one there, see sqr)
  __device__ __forceinline__ static void sqr_chain(uint32_t* t, const uint32_t* v, int i, int j0) {
    if (j0 >= N) return;
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(t[i + j0]), "+r"(t[i + j0 + 1])
                 : "r"(v[i]), "r"(v[j0]));
    int last = j0;
#pragma unroll
    for (int j = j0 + 2; j < N; j += 2) {
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                   : "+r"(t[i + j]), "+r"(t[i + j + 1])
                   : "r"(v[i]), "r"(v[j]));
      last = j;
    }
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(t[i + last + 2]));
  }
  // Dedicated squaring: the wide square from its N (N - 1) / 2 off-diagonal products (doubled by a one-bit shift) plus
  // the N diagonal ones, then a separate Montgomery reduction: N (N + 1) / 2 + N^2 limb products instead of 2 N^2
  // (Fq: 222 instead of 288).  The shifts and the carry bookkeeping run on the integer ALU, not on the multiplier pipe.
  __device__ __forceinline__ Fp sqr_dev() const {
    uint32_t t[2 * N], m[N], cnt[N + 1];
#pragma unroll
    for (int i = 0; i < 2 * N; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) m[i] = P::MOD(i);
#pragma unroll
    for (int i = 0; i <= N; i++) cnt[i] = 0;
    // sum_{i < j} v[i] v[j] 2^(32 (i + j)).  Row i has two chains (j - i odd / even); the one that ends lower goes first:
    // its carry lands on limb i + N, where the earlier rows left at most one, and the other chain's on limb i + N + 1,
    // untouched so far (rows 0 .. i add up to less than 2^(32 (i + N + 1) + 1)).
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
      if (((N - 1 - (i + 1)) & 1) == 0) {  // j = N - 1 belongs to the chain that starts at i + 1
        sqr_chain(t, v, i, i + 2);
        sqr_chain(t, v, i, i + 1);
      } else {
        sqr_chain(t, v, i, i + 1);
        sqr_chain(t, v, i, i + 2);
      }
    }
#pragma unroll
    for (int k = 2 * N - 1; k > 0; k--) t[k] = __funnelshift_l(t[k - 1], t[k], 1);
    t[0] = 0;  // (no product lands on limb 0)
    asm volatile("mad.lo.cc.u32 %0, %2, %2, %0; madc.hi.cc.u32 %1, %2, %2, %1;" : "+r"(t[0]), "+r"(t[1]) : "r"(v[0]));
#pragma unroll
    for (int i = 1; i < N - 1; i++)
      asm volatile("madc.lo.cc.u32 %0, %2, %2, %0; madc.hi.cc.u32 %1, %2, %2, %1;"
                   : "+r"(t[2 * i]), "+r"(t[2 * i + 1])
                   : "r"(v[i]));
    asm volatile("madc.lo.cc.u32 %0, %2, %2, %0; madc.hi.u32 %1, %2, %2, %1;"
                 : "+r"(t[2 * N - 2]), "+r"(t[2 * N - 1])
                 : "r"(v[N - 1]));
    // Montgomery reduction: step i clears limb i with mi * m; the two chains' carries (onto limbs i + N and i + N + 1)
    // are counted aside and added once at the end (t + sum mi m 2^(32 i) < 2 m R fits the 2 N limbs)
#pragma unroll
    for (int i = 0; i < N; i++) {
      const uint32_t mi = t[i] * P::M0;
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(t[i]), "+r"(t[i + 1]) : "r"(m[0]), "r"(mi));
#pragma unroll
      for (int j = 2; j < N; j += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(t[i + j]), "+r"(t[i + j + 1])
                     : "r"(m[j]), "r"(mi));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(cnt[i]));
      if (i + N < 2 * N - 1) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(t[i + 1]), "+r"(t[i + 2])
                     : "r"(m[1]), "r"(mi));
#pragma unroll
        for (int j = 3; j < N; j += 2)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                       : "+r"(t[i + j]), "+r"(t[i + j + 1])
                       : "r"(m[j]), "r"(mi));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(cnt[i + 1]));
      } else {  // last step: the odd chain ends on the top limb, nothing can carry out of it
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(t[i + 1]), "+r"(t[i + 2])
                     : "r"(m[1]), "r"(mi));
#pragma unroll
        for (int j = 3; j < N - 2; j += 2)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                       : "+r"(t[i + j]), "+r"(t[i + j + 1])
                       : "r"(m[j]), "r"(mi));
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(t[i + N - 1]), "+r"(t[i + N])
                     : "r"(m[N - 1]), "r"(mi));
      }
    }
    Fp r;
    uint32_t cf = 0;
    r.v[0] = add_cc(t[N], cnt[0], cf);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(t[N + i], cnt[i], cf);
    r.v[N - 1] = addc(t[2 * N - 1], cnt[N - 1], cf);
    r.reduce_once();
    return r;
  }
#endif
  // (sqr_dev is opt-in: in accum0_kernel it was measured slower than the plain multiplication, 351.6 vs 371.8 proofs/s --
  // the 2 N-limb square, the carry counters and the modulus are 49 live registers on top of a kernel already at its
  // 168-register budget (ptxas: spills appear), and the shifts / carry bookkeeping add ~100 ALU instructions.)
  FF_HD Fp sqr() const {
#if defined(FF_INLINE_MUL) && defined(__CUDA_ARCH__) && defined(FF_USE_SQR)
    return sqr_dev();
#else
    return *this * *this;
#endif
  }

  // x (canonical, < m) -> Montgomery form, and back
  FF_HD Fp to_mont() const { return *this * r2(); }
  FF_HD Fp from_mont() const {
    Fp o = zero();
    o.v[0] = 1;
    return *this * o;
  }
  FF_HD static Fp from_u32(uint32_t x) {
    Fp o = zero();
    o.v[0] = x;
    return o.to_mont();
  }
  // a^(m-2); 0 -> 0
  FF_HD Fp inverse() const {
    Fp r = one(), b = *this;
    uint32_t e[N];
#pragma unroll
    for (int i = 0; i < N; i++) e[i] = P::MOD(i);
    {  // e = m - 2
      uint32_t borrow = 2;
#pragma unroll
      for (int i = 0; i < N; i++) {
        uint32_t t = e[i] - borrow;
        borrow = e[i] < borrow ? 1 : 0;
        e[i] = t;
      }
    }
    for (int i = 0; i < 32 * N; i++) {
      if ((e[i >> 5] >> (i & 31)) & 1) r = r * b;
      b = b.sqr();
    }
    return r;
  }
  // The same power with a fixed 4-bit window: 4 N * 8 squarings + at most 8 N + 14 multiplications (Fq: 384 + 110
  // instead of ~574).  Out of line: called once per batch of shared inversions (msm_impl.cuh, pair_kernel).
  FF_NOINLINE static Fp inverse_w4(Fp a) {
    Fp tab[16];
    tab[0] = one();
    tab[1] = a;
    for (int i = 2; i < 16; i++) tab[i] = tab[i - 1] * a;
    uint32_t e[N];
#pragma unroll
    for (int i = 0; i < N; i++) e[i] = P::MOD(i);
    {  // e = m - 2
      uint32_t borrow = 2;
#pragma unroll
      for (int i = 0; i < N; i++) {
        uint32_t t = e[i] - borrow;
        borrow = e[i] < borrow ? 1 : 0;
        e[i] = t;
      }
    }
    Fp r = tab[(e[N - 1] >> 28) & 15u];
    for (int i = 8 * N - 2; i >= 0; i--) {
      r = r.sqr();
      r = r.sqr();
      r = r.sqr();
      r = r.sqr();
      const uint32_t d = (e[i >> 3] >> (4 * (i & 7))) & 15u;
      if (d) r = r * tab[d];
    }
    return r;
  }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
  Fq c0, c1;
  FF_HD static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  FF_HD static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  FF_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  FF_HD bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  FF_HD bool operator!=(const Fq2& o) const { return !(*this == o); }
  FF_HD friend Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  FF_HD friend Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  FF_HD friend Fq2 operator*(const Fq2& a, const Fq2& b) {
    Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
    Fq t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
    return {t0 - t1, t2 - t0 - t1};
  }
  FF_HD Fq2 sqr() const {
    Fq a = (c0 + c1) * (c0 - c1);
    Fq b = c0 * c1;
    return {a, b + b};
  }
  FF_HD static Fq2 mul_sub2(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) { return a * b - c * d; }
  FF_HD Fq2 neg() const { return {c0.neg(), c1.neg()}; }
  FF_HD Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  FF_HD Fq2 inverse() const {
    Fq n = (c0.sqr() + c1.sqr()).inverse();
    return {c0 * n, (c1 * n).neg()};
  }
  FF_HD static Fq2 inverse_w4(const Fq2& a) {
    Fq n = Fq::inverse_w4(a.c0.sqr() + a.c1.sqr());
    return {a.c0 * n, (a.c1 * n).neg()};
  }
};

}  // namespace ff
