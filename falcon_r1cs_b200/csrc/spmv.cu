// Subsystem (2): R1CS evaluation  a_i = <A_i, z>, b_i = <B_i, z>, c_i = <C_i, z>  and the
// satisfaction check a_i * b_i == c_i, over the circuit's CSR matrices.  Replaces
// `evaluate_constraint` in ark-groth16 0.3.0's R1CStoQAP::witness_map and
// ark-relations' `which_is_unsatisfied` ([EXT]; reached from examples/pok_sig.rs:32
// and the `cs.is_satisfied()` asserts, e.g. circuits/falcon_ntt.rs:159).
//
// Row classes and their kernels, every one producing the same a_i, b_i, c_i as the term-by-term
// field evaluation:
//   * ~151k rows whose matrices are each z[p] - z[n]                      r1cs_pm1_kernel (thread per row and signature)
//   * ~9k other short rows (bit decompositions, small coefficients)      r1cs_fast_short_kernel
//   * 2N+1 long rows with integer coefficients (inlined NTT, norm row)   r1cs_bundle_kernel + r1cs_bundle_finish_kernel
//     for batches of 64 signatures and more, r1cs_signed_long_kernel (warp per 4 rows x 8 signatures) below that
//   * long rows with field-sized coefficients (none in the NTT circuits) r1cs_fast_long_kernel
//   * assignments whose "small" columns are not small (invalid ones)     exact fall-backs: row_dot / warp_row_dot
// plus the plain CSR kernels r1cs_short_kernel / r1cs_long_kernel used by the set-up (launch_matvec3).
#define FF_OPAQUE_M0  // (ff32.cuh: keeps the Fr multiplications on IMAD.WIDE.U32)
#include <algorithm>
#include <array>
#include <map>
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "nvtx.hpp"
#define FF_INLINE_MUL
#include "ff32.cuh"

using ff::Fr;

namespace {

constexpr uint32_t LONG_ROW = 64;

__device__ __forceinline__ Fr load_fr(const uint32_t* p) {
  Fr r;
  uint64_t a, b, c, d;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  r.v[0] = (uint32_t)a; r.v[1] = (uint32_t)(a >> 32);
  r.v[2] = (uint32_t)b; r.v[3] = (uint32_t)(b >> 32);
  r.v[4] = (uint32_t)c; r.v[5] = (uint32_t)(c >> 32);
  r.v[6] = (uint32_t)d; r.v[7] = (uint32_t)(d >> 32);
  return r;
}
__device__ __forceinline__ void store_fr(uint32_t* p, const Fr& x) {
  uint64_t a = (uint64_t)x.v[0] | ((uint64_t)x.v[1] << 32), b = (uint64_t)x.v[2] | ((uint64_t)x.v[3] << 32);
  uint64_t c = (uint64_t)x.v[4] | ((uint64_t)x.v[5] << 32), d = (uint64_t)x.v[6] | ((uint64_t)x.v[7] << 32);
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ bool is_one(const Fr& c) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= c.v[i] ^ FrParams::R1(i);
  return o == 0;
}

__global__ void to_montgomery_kernel(uint32_t* vals, uint64_t count) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fr x = load_fr(vals + 8 * i);
  store_fr(vals + 8 * i, x.to_mont());
}

// dot product of one CSR row with z, serial
__device__ __forceinline__ Fr row_dot(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                      const uint32_t* __restrict__ val, const uint32_t* __restrict__ z, uint32_t row,
                                      const Fr& minus_one) {
  Fr acc = Fr::zero();
  uint32_t k0 = row_ptr[row], k1 = row_ptr[row + 1];
  for (uint32_t k = k0; k < k1; k++) {
    Fr c = load_fr(val + 8 * (uint64_t)k);
    Fr x = load_fr(z + 8 * (uint64_t)col[k]);
    if (is_one(c))
      acc = acc + x;
    else if (c == minus_one)
      acc = acc - x;
    else
      acc = acc + c * x;
  }
  return acc;
}

struct EvalArgs {
  const uint32_t *a_ptr, *a_col, *a_val;
  const uint32_t *b_ptr, *b_col, *b_val;
  const uint32_t *c_ptr, *c_col, *c_val;
  uint32_t n_cons, n_z;
  uint64_t out_stride;  // Fr elements between consecutive signatures in az / bz / cz
};

// one thread per (signature, row) for rows whose A-row is short
__global__ void __launch_bounds__(256)
    r1cs_short_kernel(EvalArgs g, const uint32_t* __restrict__ z_all, uint32_t* az, uint32_t* bz, uint32_t* cz,
                      unsigned long long* first_unsat) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t sid = blockIdx.y;
  if (row >= g.n_cons) return;
  if (g.a_ptr[row + 1] - g.a_ptr[row] > LONG_ROW || g.b_ptr[row + 1] - g.b_ptr[row] > LONG_ROW ||
      g.c_ptr[row + 1] - g.c_ptr[row] > LONG_ROW)
    return;
  const uint32_t* z = z_all + sid * (uint64_t)g.n_z * 8;
  Fr m1 = Fr::one().neg();
  Fr a = row_dot(g.a_ptr, g.a_col, g.a_val, z, row, m1);
  Fr b = row_dot(g.b_ptr, g.b_col, g.b_val, z, row, m1);
  Fr c = row_dot(g.c_ptr, g.c_col, g.c_val, z, row, m1);
  uint64_t o = (sid * g.out_stride + row) * 8;
  if (az) store_fr(az + o, a);
  if (bz) store_fr(bz + o, b);
  if (cz) store_fr(cz + o, c);
  if (first_unsat && a * b != c) atomicMin(first_unsat + sid, (unsigned long long)row);
}

// lane-strided dot product of one CSR row, reduced over the warp
__device__ __forceinline__ Fr warp_row_dot(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                           const uint32_t* __restrict__ val, const uint32_t* __restrict__ z,
                                           uint32_t row, uint32_t lane) {
  Fr acc = Fr::zero();
  uint32_t k0 = row_ptr[row], k1 = row_ptr[row + 1];
  for (uint32_t k = k0 + lane; k < k1; k += 32) {
    Fr c = load_fr(val + 8 * (uint64_t)k);
    Fr x = load_fr(z + 8 * (uint64_t)col[k]);
    acc = acc + c * x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Fr other;
#pragma unroll
    for (int i = 0; i < 8; i++) other.v[i] = __shfl_xor_sync(0xffffffffu, acc.v[i], o);
    acc = acc + other;
  }
  return acc;
}

// one warp per (signature, long row): a row is "long" if it has > LONG_ROW non-zeros in any matrix
__global__ void __launch_bounds__(256)
    r1cs_long_kernel(EvalArgs g, const uint32_t* __restrict__ long_rows, uint32_t n_long,
                     const uint32_t* __restrict__ z_all, uint32_t* az, uint32_t* bz, uint32_t* cz,
                     unsigned long long* first_unsat) {
  uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint64_t sid = blockIdx.y;
  if (wid >= n_long) return;
  uint32_t row = long_rows[wid];
  const uint32_t* z = z_all + sid * (uint64_t)g.n_z * 8;
  Fr a = warp_row_dot(g.a_ptr, g.a_col, g.a_val, z, row, lane);
  Fr b = warp_row_dot(g.b_ptr, g.b_col, g.b_val, z, row, lane);
  Fr c = warp_row_dot(g.c_ptr, g.c_col, g.c_val, z, row, lane);
  if (lane == 0) {
    uint64_t o = (sid * g.out_stride + row) * 8;
    if (az) store_fr(az + o, a);
    if (bz) store_fr(bz + o, b);
    if (cz) store_fr(cz + o, c);
    if (first_unsat && a * b != c) atomicMin(first_unsat + sid, (unsigned long long)row);
  }
}

__global__ void init_unsat_kernel(unsigned long long* p, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = ~0ull;  // == (int64_t)-1 when no row is violated
}

}  // namespace

int32_t launch_to_montgomery(frcs_ctx* ctx, uint32_t* d_vals, uint64_t count, cudaStream_t st) {
  if (count == 0) return FRCS_OK;
  // Set-up path only.  The values were just uploaded with cudaMemcpy from pageable memory, which returns once the
  // data is staged: the DMA itself may still be running on the legacy stream, and `st` is a non-blocking stream
  // that does not wait for it.  (Seen as a few thousand coefficients left unconverted, one run in three.)
  FRCS_CUDA_CHECK(cudaDeviceSynchronize());
  to_montgomery_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_vals, count);
  ctx->launches++;
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

// =============================================================================================
// Fast path.  The coefficients of the circuit fall in two classes: +-1 / small
// integers (bits, powers of two, q) and the full-width entries of the 2N inlined NTT rows, whose
// multiplicands are the 14-bit sig / v inputs.  Either way a term is (32-bit integer) x (255-bit
// residue), so a row is accumulated lazily as a 10-limb integer,  S += small * wide  (8 IMAD.WIDE
// per term instead of a 128-product Montgomery multiplication), and reduced once per row:
// S = S_lo + 2^256 S_hi == S_lo + R S_hi (mod r), i.e. one multiplication by R^2.  All terms are in
// Montgomery form, so the reduced sum is bit-identical to the term-by-term evaluation.
// Multiplicands that are not small (possible only for invalid assignments) take a full multiplication.
// =============================================================================================
namespace {

constexpr uint32_t CODE_NEG = 0x80000000u, CODE_FULL = 0x40000000u, CODE_MASK = 0x3fffffffu, NOT_SMALL = 0xffffffffu;
// multiplicands of full-width coefficients count as small below 2^20: a lane then adds at most 2^12 products
// m * limb < 2^52 into a 64-bit accumulator per limb without any carry handling (rows have < 2^17 terms)
constexpr uint32_t SMALL_LIMIT = 1u << 20;
// the transposed view of the small columns keeps values below 2^28; the signed-digit kernel applies a per-row limit
// (build_signed_long) and the generic long-row kernel SMALL_LIMIT
constexpr uint32_t VIEW_LIMIT = 1u << 28;

struct Lazy {  // 320-bit unsigned accumulator
  uint32_t v[10];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 10; i++) v[i] = 0;
  }
  // v += m * w   (m < 2^32, w < 2^256): two carry chains (even and odd limbs of w), 8 IMAD.WIDE in all
  __device__ __forceinline__ void fma(uint32_t m, const Fr& w) {
    asm volatile(
        "mad.lo.cc.u32 %0, %10, %11, %0; madc.hi.cc.u32 %1, %10, %11, %1;"
        "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;"
        "madc.lo.cc.u32 %4, %10, %15, %4; madc.hi.cc.u32 %5, %10, %15, %5;"
        "madc.lo.cc.u32 %6, %10, %17, %6; madc.hi.cc.u32 %7, %10, %17, %7;"
        "addc.cc.u32 %8, %8, 0; addc.u32 %9, %9, 0;"
        "mad.lo.cc.u32 %1, %10, %12, %1; madc.hi.cc.u32 %2, %10, %12, %2;"
        "madc.lo.cc.u32 %3, %10, %14, %3; madc.hi.cc.u32 %4, %10, %14, %4;"
        "madc.lo.cc.u32 %5, %10, %16, %5; madc.hi.cc.u32 %6, %10, %16, %6;"
        "madc.lo.cc.u32 %7, %10, %18, %7; madc.hi.cc.u32 %8, %10, %18, %8;"
        "addc.u32 %9, %9, 0;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
          "+r"(v[9])
        : "r"(m), "r"(w.v[0]), "r"(w.v[1]), "r"(w.v[2]), "r"(w.v[3]), "r"(w.v[4]), "r"(w.v[5]), "r"(w.v[6]), "r"(w.v[7]));
  }
  __device__ __forceinline__ void add(const Lazy& o) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      c += (uint64_t)v[i] + o.v[i];
      v[i] = (uint32_t)c;
      c >>= 32;
    }
  }
  // the residue of the accumulated integer mod r
  __device__ __forceinline__ Fr reduce() const {
    Fr lo, hi = Fr::zero();
#pragma unroll
    for (int i = 0; i < 8; i++) lo.v[i] = v[i];
    lo.reduce_once();  // lo < 2^256 < 3r: two conditional subtractions
    lo.reduce_once();
    hi.v[0] = v[8];
    hi.v[1] = v[9];
    return lo + hi * Fr::r2();
  }
};

struct FastMat {
  const uint32_t *row_ptr, *col, *code, *fval, *full_end;
};
struct FastArgs {
  FastMat m[3];
  const uint2* hdr;     // per row: x = first merged term, y = nA | nB << 7 | nC << 14 (short rows only)
  const uint2* mterm;   // merged terms of the short rows, A then B then C of each row: (col, code)
  const uint32_t* mfval;  // full-width coefficients of the merged terms
  const uint32_t* small_cols;
  uint32_t n_small, n_cons, n_z;
  uint32_t xs_stride;   // signatures per small column in the transposed small view
  const uint32_t* mont_tab;  // [2][2^14] Fr: mont(j), mont(2^14 j)
  uint64_t out_stride;
  // slow path (exact term-by-term evaluation) for assignments whose "small" columns are not small
  EvalArgs slow;
};

__device__ __forceinline__ Fr neg_fr(const Fr& x) { return x.is_zero() ? x : Fr::zero() - x; }

// Montgomery image of a small signed integer: two table look-ups below 2^28 (mont(m) = T0[m mod 2^14] + T1[m >> 14])
__device__ __forceinline__ Fr small_mont(int64_t sv, const uint32_t* __restrict__ tab) {
  const bool neg = sv < 0;
  const uint64_t m = neg ? (uint64_t)(-sv) : (uint64_t)sv;
  Fr v;
  if (m < (1ull << 28)) {
    v = load_fr(tab + 8 * (m & 0x3fffu));
    if (m >> 14) v = v + load_fr(tab + 8 * (16384u + (uint32_t)(m >> 14)));
  } else {
    v = Fr::zero();
    v.v[0] = (uint32_t)m;
    v.v[1] = (uint32_t)(m >> 32);
    v = v.to_mont();
  }
  return neg ? neg_fr(v) : v;
}

struct MulItem {
  uint32_t a[8], b[8], c[8], row, sid;
};

constexpr int SS = 2;  // signatures per thread in the short-row kernel (1 with 3 blocks per SM: 6 % slower)
// one thread per (short row, pair of signatures); rows in class order
__global__ void __launch_bounds__(256)
    r1cs_fast_short_kernel(FastArgs g, const uint32_t* __restrict__ perm, uint32_t n_short,
                           const uint32_t* __restrict__ z_all, const uint32_t* __restrict__ xs_t, uint32_t n_sig,
                           uint32_t* az, uint32_t* bz, uint32_t* cz, unsigned long long* first_unsat) {
  __shared__ MulItem s_items[64];
  __shared__ uint32_t s_count;
  // rows are visited in class order (same term counts and the same lazy / exact mix per warp)
  const uint32_t t_row = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t sid0 = blockIdx.y * SS;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const bool live = t_row < n_short;
  const uint32_t row = live ? perm[t_row] : 0;
  Fr res[3][SS];
  uint32_t need_mul = 0;  // bit s: signature s of this thread needs the real product check
  if (live) {
    const uint32_t* z[SS];
    bool valid[SS];
#pragma unroll
    for (int s = 0; s < SS; s++) {
      valid[s] = sid0 + s < n_sig;
      z[s] = z_all + (uint64_t)(valid[s] ? sid0 + s : sid0) * g.n_z * 8;
    }
    const uint2 hdr = g.hdr[row];
    uint32_t k = hdr.x;
    const uint32_t cnt3[3] = {hdr.y & 0x7fu, (hdr.y >> 7) & 0x7fu, (hdr.y >> 14) & 0x7fu};
#pragma unroll
    for (int m = 0; m < 3; m++) {
      Fr acc[SS];
      Lazy lazy[SS];
      int64_t isum[SS];  // sum of the small coefficients whose multiplicand is the bit 1 (decomposition rows)
      bool used = false;
#pragma unroll
      for (int s = 0; s < SS; s++) {
        acc[s] = Fr::zero();
        lazy[s].clear();
        isum[s] = 0;
      }
      const uint32_t k1 = k + cnt3[m];
#pragma unroll 2
      for (; k < k1; k++) {
        const uint2 t = g.mterm[k];
        const uint32_t col = t.x, code = t.y;
        if (code & CODE_FULL) {
          Fr c = load_fr(g.mfval + 8 * (uint64_t)(code & CODE_MASK));
#pragma unroll
          for (int s = 0; s < SS; s++) {
            uint32_t x = xs_t[(uint64_t)col * g.xs_stride + (valid[s] ? sid0 + s : sid0)];
            if (x != NOT_SMALL)
              lazy[s].fma(x, c);
            else
              acc[s] = acc[s] + c * load_fr(z[s] + 8 * (uint64_t)g.small_cols[col]);
          }
          used = true;
        } else {
          const uint32_t mag = code & CODE_MASK;
          Fr x[SS];
#pragma unroll
          for (int s = 0; s < SS; s++) x[s] = load_fr(z[s] + 8 * (uint64_t)col);
          if (mag == 1) {
#pragma unroll
            for (int s = 0; s < SS; s++) acc[s] = (code & CODE_NEG) ? acc[s] - x[s] : acc[s] + x[s];
          } else {
            // multiplicands are mostly bits: 0 adds nothing, 1 adds the coefficient to an integer sum
#pragma unroll
            for (int s = 0; s < SS; s++) {
              if (x[s].is_zero()) continue;
              if (is_one(x[s])) {
                isum[s] += (code & CODE_NEG) ? -(int64_t)mag : (int64_t)mag;
              } else {
                lazy[s].fma(mag, (code & CODE_NEG) ? neg_fr(x[s]) : x[s]);
                used = true;
              }
            }
          }
        }
      }
#pragma unroll
      for (int s = 0; s < SS; s++) {
        res[m][s] = used ? acc[s] + lazy[s].reduce() : acc[s];
        if (isum[s] != 0) res[m][s] = res[m][s] + small_mont(isum[s], g.mont_tab);
      }
    }
#pragma unroll
    for (int s = 0; s < SS; s++) {
      if (!valid[s]) continue;
      const uint64_t o = ((uint64_t)(sid0 + s) * g.out_stride + row) * 8;
      if (az) store_fr(az + o, res[0][s]);
      if (bz) store_fr(bz + o, res[1][s]);
      if (cz) store_fr(cz + o, res[2][s]);
      if (first_unsat) {
        const Fr &a = res[0][s], &b = res[1][s], &c = res[2][s];
        bool bad = false;
        if (a.is_zero() || b.is_zero())
          bad = !c.is_zero();
        else if (is_one(a))
          bad = b != c;
        else if (is_one(b))
          bad = a != c;
        else
          need_mul |= 1u << s;
        if (bad) atomicMin(first_unsat + sid0 + s, (unsigned long long)row);
      }
    }
  }
  if (!first_unsat) return;
  // the rows that need a real multiplication are compacted in shared memory and handled by a few threads
  for (;;) {
#pragma unroll
    for (int s = 0; s < SS; s++) {
      if ((need_mul >> s) & 1) {
        uint32_t slot = atomicAdd(&s_count, 1u);
        if (slot < 64) {
#pragma unroll
          for (int i = 0; i < 8; i++) {
            s_items[slot].a[i] = res[0][s].v[i];
            s_items[slot].b[i] = res[1][s].v[i];
            s_items[slot].c[i] = res[2][s].v[i];
          }
          s_items[slot].row = row;
          s_items[slot].sid = sid0 + s;
          need_mul &= ~(1u << s);
        }
      }
    }
    __syncthreads();
    const uint32_t total = s_count, n = min(total, 64u);
    if (threadIdx.x < n) {
      Fr x, y, w;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        x.v[i] = s_items[threadIdx.x].a[i];
        y.v[i] = s_items[threadIdx.x].b[i];
        w.v[i] = s_items[threadIdx.x].c[i];
      }
      if (x * y != w) atomicMin(first_unsat + s_items[threadIdx.x].sid, (unsigned long long)s_items[threadIdx.x].row);
    }
    __syncthreads();
    if (total <= 64) break;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
  }
}

// Rows whose every matrix is  z[p] - z[n]  (either side optional): 94 % of the rows of the Falcon circuits (Boolean
// constraints, and / or gates, selections, products of two variables).  One thread per (row, signature), straight-line:
// the six 32-byte loads are issued together, so a thread has its whole working set in flight and the kernel runs at
// the rate the rows' z entries stream in from HBM.  Descriptor = 8 words: row, (p, n) of A, B, C, pad; PM1_NONE = absent.
constexpr uint32_t PM1_NONE = 0xffffffffu;
#ifndef PM1_BLOCKS
#define PM1_BLOCKS 4
#endif
// z[col], or zero without touching memory for an absent term (predicated load: no branch, so the six loads of a
// thread are issued back to back)
__device__ __forceinline__ Fr pm1_load(const uint32_t* z, uint32_t col) {
  uint64_t a, b, c, d;
  asm volatile(
      "{.reg .pred p; setp.ne.u32 p, %5, 0xffffffff;\n\t"
      "mov.b64 %0, 0; mov.b64 %1, 0; mov.b64 %2, 0; mov.b64 %3, 0;\n\t"
      "@p ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];}"
      : "=&l"(a), "=&l"(b), "=&l"(c), "=&l"(d)
      : "l"(z + 8 * (uint64_t)(col != PM1_NONE ? col : 0u)), "r"(col));
  Fr r;
  r.v[0] = (uint32_t)a; r.v[1] = (uint32_t)(a >> 32);
  r.v[2] = (uint32_t)b; r.v[3] = (uint32_t)(b >> 32);
  r.v[4] = (uint32_t)c; r.v[5] = (uint32_t)(c >> 32);
  r.v[6] = (uint32_t)d; r.v[7] = (uint32_t)(d >> 32);
  return r;
}
__global__ void __launch_bounds__(256, PM1_BLOCKS)
    r1cs_pm1_kernel(const uint4* __restrict__ desc, uint32_t n_rows, const uint32_t* __restrict__ z_all, uint32_t n_z,
                    uint64_t out_stride, uint32_t* az, uint32_t* bz, uint32_t* cz, unsigned long long* first_unsat) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t sid = blockIdx.y;
  if (t >= n_rows) return;
  const uint4 d0 = desc[2 * t], d1 = desc[2 * t + 1];
  const uint32_t row = d0.x;
  const uint32_t* z = z_all + (uint64_t)sid * n_z * 8;
  // all six loads first (absent terms: predicated off, value zero), then the arithmetic
  const uint32_t cols[6] = {d0.y, d0.z, d0.w, d1.x, d1.y, d1.z};
  Fr v[6];
#pragma unroll
  for (int i = 0; i < 6; i++) v[i] = pm1_load(z, cols[i]);
  // rows are in class order (which of the six terms exist), so these branches are uniform over a warp
  Fr a = v[0], b = v[2], c = v[4];
  if (cols[1] != PM1_NONE) a = a - v[1];
  if (cols[3] != PM1_NONE) b = b - v[3];
  if (cols[5] != PM1_NONE) c = c - v[5];
  const uint64_t o = ((uint64_t)sid * out_stride + row) * 8;
  if (az) store_fr(az + o, a);
  if (bz) store_fr(bz + o, b);
  if (cz) store_fr(cz + o, c);
  if (first_unsat) {
    bool bad;
    if (a.is_zero() || b.is_zero())
      bad = !c.is_zero();
    else if (is_one(a))
      bad = b != c;
    else if (is_one(b))
      bad = a != c;
    else
      bad = a * b != c;  // rows are in class order: the warps of the product rows all come here, the Boolean ones never
    if (bad) atomicMin(first_unsat + sid, (unsigned long long)row);
  }
}


// =============================================================================================
// Streaming short rows (all batch sizes): z is cut into windows of SW consecutive columns; every short row lives in
// the window of its highest column (a gadget's rows sit right behind its witnesses, so almost all of a row's columns
// fall in that window); the few columns a window's rows reference outside it are its "remote" slots.  One CTA takes
// one window for a run of signatures: the window's row program (16-bit slot references) is loaded into shared memory
// once, then per signature the window is read from HBM once, in contiguous 32 KB pieces, and all of its rows are
// evaluated.  Replaces the two kernels above, which read z as scattered 32-byte sectors and a second time for the bit
// columns of the decomposition rows.
// =============================================================================================
#ifndef FRCS_SW
#define FRCS_SW 1024
#endif
#ifndef FRCS_STREAM_CTAS
#define FRCS_STREAM_CTAS 4
#endif
constexpr uint32_t SW = FRCS_SW;     // columns per window (32 B each)
constexpr uint32_t SREF_NONE = 0xffffu;  // host side only: the kernel sees absent terms as references to a zero slot
constexpr int STREAM_THREADS = 256;
struct StreamWin {
  uint32_t col_lo, n_cols, n_remote, n_pm1, n_gen, n_terms;
  uint32_t desc_bytes, pad;
  uint64_t desc_off;  // byte offset of the window's program in the blob (16-byte aligned):
                      // [remote columns u32 x n_remote (padded to 4)] [pm1 rows: uint4] [general rows: uint4] [terms: uint2]
};
struct StreamArgs {
  const StreamWin* wins;
  const uint8_t* desc;
  const uint32_t* mont_tab;
  uint32_t n_z, slots, desc_max;  // slots = SW + max n_remote (+ the zero slot)
  uint32_t qmax;                  // most z[p] - z[n] rows of any window (capacity of a slow-row queue), even
  uint64_t out_stride;
};

// a * b == c without a multiplication when a or b is 0 or 1 (Boolean rows)
__device__ __forceinline__ bool row_violated(const Fr& a, const Fr& b, const Fr& c) {
  if (a.is_zero() || b.is_zero()) return !c.is_zero();
  if (is_one(a)) return b != c;
  if (is_one(b)) return a != c;
  return a * b != c;
}

// -1, 0 or 1 as a field element (Montgomery form)
__device__ __forceinline__ Fr fr_of_sign(int v) {
  if (v == 0) return Fr::zero();
  const Fr one = Fr::one();
  return v > 0 ? one : Fr::zero() - one;
}

// class of one 32-byte entry given as two 16-byte halves: 0 (zero), 1 (Montgomery one), 2 (anything else)
__device__ __forceinline__ uint32_t entry_class(const uint4& lo, const uint4& hi) {
  const uint32_t any = lo.x | lo.y | lo.z | lo.w | hi.x | hi.y | hi.z | hi.w;
  const uint32_t dif = (lo.x ^ FrParams::R1(0)) | (lo.y ^ FrParams::R1(1)) | (lo.z ^ FrParams::R1(2)) | (lo.w ^ FrParams::R1(3)) |
                       (hi.x ^ FrParams::R1(4)) | (hi.y ^ FrParams::R1(5)) | (hi.z ^ FrParams::R1(6)) | (hi.w ^ FrParams::R1(7));
  return any == 0 ? 0u : dif == 0 ? 1u : 2u;
}

// Per signature a CTA (A) classifies every entry of its window -- 0, 1 (Montgomery one) or something else -- into one
// byte of shared memory: every thread loads its share of the window (plus at most one column from outside it) straight
// from HBM, all loads in flight at once; 91 % of an assignment are Boolean witnesses, and a row over bits is decided
// from six class bytes; (B) evaluates the window's rows from the row program in shared memory; the few rows that
// involve a non-bit are queued and evaluated together one iteration later (whole warps instead of a few lanes of
// many), reading their 32-byte operands back through L1 / L2.  One __syncthreads per signature: class bytes are double
// buffered, the slow-row queues triple buffered; 4 CTAs per SM hide the barrier and the load latency.
// Where the time goes (ncu source-level sampling, 592 signatures): 47 % of the warp-stall samples sit at the barrier and
// 9 % wait for the window's loads; issue slots are 45 % busy.  Measured without effect on the 1.24 ms: the next
// signature's window prefetched into L2 while the rows are evaluated (387 k checks/s either way); run lengths fitted to
// whole waves; deciding the Boolean constraints (One - x) * x = 0 -- half of the z[p] - z[n] rows -- from one class
// byte in the verdict-only mode (387 k/s); four terms of a term-list row in flight per thread (360 k/s, 44 B of
// spills at the 64-register bound); z[p] - z[n] rows claimed 32 at a time by whichever warp is free instead of a fixed
// split between the term-list warps and the others (381 k/s).  So the waiting at the barrier is not an imbalance of the
// row work: it is the window's load latency, seen by the warps whose own loads came back first.
// Measured alternatives, per 592 signatures (this version: see profiles/): staging the 32 KB window in shared memory
// with cp.async.bulk + mbarrier and evaluating rows from the 32-byte values there: 1.9 ms (shared-memory reads, three
// barriers per window; every remote column as its own 32-byte bulk copy cost ~46 cycles of TMA issue each); the same
// with class bytes, the bulk copy issued one signature ahead and 3 CTAs per SM: 1.48 ms.
__global__ void __launch_bounds__(STREAM_THREADS, FRCS_STREAM_CTAS)
    r1cs_stream_kernel(StreamArgs g, const uint32_t* __restrict__ z_all, uint32_t n_sig, uint32_t sig_per_cta,
                       uint32_t* az, uint32_t* bz, uint32_t* cz, unsigned long long* first_unsat) {
  extern __shared__ __align__(16) uint8_t stream_smem[];
  uint32_t* qcount = reinterpret_cast<uint32_t*>(stream_smem);         // [3] slow-row queue lengths
  const uint32_t cls_bytes = (g.slots + 15) & ~15u;
  uint8_t* cls0 = stream_smem + 16;                                    // [2][slots] class of every entry
  uint16_t* queue0 = reinterpret_cast<uint16_t*>(cls0 + 2 * cls_bytes);  // [3][qmax] rows waiting for the slow path
  uint8_t* dsm = reinterpret_cast<uint8_t*>(queue0) + 3 * (size_t)g.qmax * 2;  // the window's program
  const StreamWin win = g.wins[blockIdx.y];
  const uint32_t tid = threadIdx.x;
  const uint32_t s0 = blockIdx.x * sig_per_cta;
  if (s0 >= n_sig) return;
  const uint32_t S = min(sig_per_cta, n_sig - s0);
  {
    const uint4* src = reinterpret_cast<const uint4*>(g.desc + win.desc_off);
    uint4* dst = reinterpret_cast<uint4*>(dsm);
    for (uint32_t i = tid; i < win.desc_bytes / 16; i += STREAM_THREADS) dst[i] = src[i];
  }
  const uint32_t zero_slot = SW + win.n_remote;  // absent terms reference this slot (value 0, class 0)
  if (tid < 2) cls0[tid * cls_bytes + zero_slot] = 0;
  if (tid < 3) qcount[tid] = 0;
  __syncthreads();
  const uint32_t* remote = reinterpret_cast<const uint32_t*>(dsm);
  const uint32_t rem_words = (win.n_remote + 3) & ~3u;
  const uint4* pm1 = reinterpret_cast<const uint4*>(dsm + rem_words * 4);
  const uint4* gen = pm1 + win.n_pm1;
  const uint2* terms = reinterpret_cast<const uint2*>(gen + win.n_gen);
  const uint32_t my_remote = tid < win.n_remote ? remote[tid] : 0u;
  // warps [0, gen_warps) take the term-list rows first; the z[p] - z[n] rows are spread over the other warps
  const uint32_t gen_threads = min((win.n_gen + 31u) & ~31u, (uint32_t)STREAM_THREADS - 64u);
  const uint32_t pm1_threads = STREAM_THREADS - gen_threads;
  const bool want_out = az || bz || cz;
  constexpr uint32_t PER = SW / STREAM_THREADS;  // window entries per thread
  // the rows of signature `sid` that involve a non-bit: evaluated from the 32-byte entries (through L1 / L2)
  auto slow_rows = [&](uint32_t sid, const uint16_t* queue, uint32_t n) {
    const uint32_t* z = z_all + (uint64_t)sid * g.n_z * 8;
    auto value = [&](uint32_t slot) -> Fr {
      if (slot == zero_slot) return Fr::zero();
      FRCS_ASSERT(slot < zero_slot && (slot >= SW || slot < win.n_cols));
      const uint32_t col = slot < SW ? win.col_lo + slot : remote[slot - SW];
      FRCS_ASSERT(col < g.n_z);
      return load_fr(z + (uint64_t)col * 8);
    };
    for (uint32_t i = tid; i < n; i += STREAM_THREADS) {
      const uint4 d = pm1[queue[i]];
      const uint32_t row = d.x;
      const Fr a = value(d.y & 0xffffu) - value(d.y >> 16), b = value(d.z & 0xffffu) - value(d.z >> 16),
               cc = value(d.w & 0xffffu) - value(d.w >> 16);
      const uint64_t o = ((uint64_t)sid * g.out_stride + row) * 8;
      if (az) store_fr(az + o, a);
      if (bz) store_fr(bz + o, b);
      if (cz) store_fr(cz + o, cc);
      if (first_unsat && row_violated(a, b, cc)) atomicMin(first_unsat + sid, (unsigned long long)row);
    }
  };
#pragma unroll 1
  for (uint32_t k = 0; k < S; k++) {
    const uint32_t sid = s0 + k;
    const uint32_t* z = z_all + (uint64_t)sid * g.n_z * 8;
    uint8_t* cls = cls0 + (k & 1) * cls_bytes;
    // (A) classify
    {
      const uint4* zw = reinterpret_cast<const uint4*>(z + (uint64_t)win.col_lo * 8);
      uint4 lo[PER], hi[PER], rlo = make_uint4(0, 0, 0, 0), rhi = rlo;
#pragma unroll
      for (uint32_t j = 0; j < PER; j++) {
        const uint32_t e = j * STREAM_THREADS + tid;
        if (e < win.n_cols) {
          lo[j] = __ldg(zw + 2 * e);
          hi[j] = __ldg(zw + 2 * e + 1);
        } else {
          lo[j] = hi[j] = make_uint4(0, 0, 0, 0);
        }
      }
      if (tid < win.n_remote) {
        const uint4* p = reinterpret_cast<const uint4*>(z + (uint64_t)my_remote * 8);
        rlo = __ldg(p);
        rhi = __ldg(p + 1);
      }
#pragma unroll
      for (uint32_t j = 0; j < PER; j++) cls[j * STREAM_THREADS + tid] = (uint8_t)entry_class(lo[j], hi[j]);
      if (tid < win.n_remote) cls[SW + tid] = (uint8_t)entry_class(rlo, rhi);
    }
    __syncthreads();
    if (tid == 0) qcount[(k + 1) % 3] = 0;  // last read one iteration ago, before the barrier everybody just passed
    if (k > 0) slow_rows(sid - 1, queue0 + ((k - 1) % 3) * (size_t)g.qmax, qcount[(k - 1) % 3]);
    uint16_t* queue = queue0 + (k % 3) * (size_t)g.qmax;
    uint32_t* qn = qcount + (k % 3);
    const uint64_t obase = (uint64_t)sid * g.out_stride;
    // (B.2) the term-list rows (bit decompositions, add_mod, selections): bit multiplicands go to an integer sum
    if (tid < gen_threads)
      for (uint32_t i = tid; i < win.n_gen; i += gen_threads) {
        const uint4 h = gen[i];
        const uint32_t row = h.x;
        uint32_t t = h.y;
        Fr res0, res1, res2;
#pragma unroll
        for (int m = 0; m < 3; m++) {
          Fr acc = Fr::zero();
          Lazy lazy;
          lazy.clear();
          int64_t isum = 0;
          bool used = false;
          const uint32_t t1 = t + ((h.z >> (8 * m)) & 0xffu);
          for (; t < t1; t++) {
            const uint2 tm = terms[t];
            FRCS_ASSERT(t < win.n_terms && tm.x < zero_slot);
            const uint32_t code = tm.y, mag = code & CODE_MASK, cl = cls[tm.x];
            if (cl == 0) continue;
            if (cl == 1) {
              isum += (code & CODE_NEG) ? -(int64_t)mag : (int64_t)mag;
              continue;
            }
            const uint32_t col = tm.x < SW ? win.col_lo + tm.x : remote[tm.x - SW];
            const Fr x = load_fr(z + (uint64_t)col * 8);
            if (mag == 1) {
              acc = (code & CODE_NEG) ? acc - x : acc + x;
            } else {
              lazy.fma(mag, (code & CODE_NEG) ? neg_fr(x) : x);
              used = true;
            }
          }
          if (used) acc = acc + lazy.reduce();
          if (isum != 0) acc = acc + small_mont(isum, g.mont_tab);
          if (m == 0) res0 = acc;
          if (m == 1) res1 = acc;
          if (m == 2) res2 = acc;
        }
        const uint64_t o = (obase + row) * 8;
        if (az) store_fr(az + o, res0);
        if (bz) store_fr(bz + o, res1);
        if (cz) store_fr(cz + o, res2);
        if (first_unsat && row_violated(res0, res1, res2)) atomicMin(first_unsat + sid, (unsigned long long)row);
      }
    // (B.1) rows whose matrices are each z[p] - z[n]
    if (tid >= gen_threads)
      for (uint32_t i = tid - gen_threads; i < win.n_pm1; i += pm1_threads) {
        const uint4 d = pm1[i];
        const uint32_t row = d.x;
        const uint32_t r[6] = {d.y & 0xffffu, d.y >> 16, d.z & 0xffffu, d.z >> 16, d.w & 0xffffu, d.w >> 16};
        uint32_t c[6];
#pragma unroll
        for (int q = 0; q < 6; q++) {
          FRCS_ASSERT(r[q] <= zero_slot && r[q] < g.slots);
          c[q] = cls[r[q]];
        }
        FRCS_ASSERT(row < g.out_stride || !want_out);
        if ((c[0] | c[1] | c[2] | c[3] | c[4] | c[5]) < 2) {  // all bits: decided over the integers (|values| <= 1)
          const int ia = (int)c[0] - (int)c[1], ib = (int)c[2] - (int)c[3], ic = (int)c[4] - (int)c[5];
          if (want_out) {
            const uint64_t o = (obase + row) * 8;
            if (az) store_fr(az + o, fr_of_sign(ia));
            if (bz) store_fr(bz + o, fr_of_sign(ib));
            if (cz) store_fr(cz + o, fr_of_sign(ic));
          }
          if (first_unsat && ia * ib != ic) atomicMin(first_unsat + sid, (unsigned long long)row);
        } else {
          const uint32_t at = atomicAdd(qn, 1u);
          FRCS_ASSERT(at < g.qmax);
          queue[at] = (uint16_t)i;
        }
      }
  }
  __syncthreads();
  slow_rows(s0 + S - 1, queue0 + ((S - 1) % 3) * (size_t)g.qmax, qcount[(S - 1) % 3]);
}

constexpr int LS = 8;  // signatures per warp in the long-row kernel
constexpr int LONG_THREADS = 128;  // 4 rows per block; <= 170 registers -> 3 blocks (12 warps) per SM

// one term of a row for one signature (serial evaluation by a single lane)
__device__ __forceinline__ void serial_term(const FastArgs& g, const FastMat& M, uint32_t k, const uint32_t* z,
                                            const uint32_t* xs_t, uint32_t sid, Fr& acc, Lazy& lazy) {
  const uint32_t code = M.code[k], col = M.col[k];
  if (code & CODE_FULL) {
    Fr c = load_fr(M.fval + 8 * (uint64_t)(code & CODE_MASK));
    uint32_t x = xs_t[(uint64_t)col * g.xs_stride + sid];
    if (x != NOT_SMALL)
      lazy.fma(x, c);
    else
      acc = acc + c * load_fr(z + 8 * (uint64_t)g.small_cols[col]);
  } else {
    Fr x = load_fr(z + 8 * (uint64_t)col);
    lazy.fma(code & CODE_MASK, (code & CODE_NEG) ? neg_fr(x) : x);
  }
}

// one warp per (long row, tile of 8 signatures): every coefficient is loaded once per tile; lanes
// stride over the terms and accumulate lazily; the 8 accumulators are combined by a reduce-scatter
// over the lanes (sig s ends up in lane 4 s).  Matrices with <= 8 terms in the row (B and C of the
// NTT rows) are evaluated serially by one lane per signature.
__global__ void __launch_bounds__(LONG_THREADS, 3)
    r1cs_fast_long_kernel(FastArgs g, const uint32_t* __restrict__ long_rows, uint32_t n_long,
                          const uint32_t* __restrict__ z_all, const uint32_t* __restrict__ xs_t, uint32_t n_sig,
                          uint32_t* az, uint32_t* bz, uint32_t* cz, unsigned long long* first_unsat) {
  // blockIdx.x = signature tile (fastest-varying), blockIdx.y = group of rows: the blocks that share a group's
  // coefficients run back to back, so those are fetched from HBM once and then hit in L2
  const uint32_t wid = (blockIdx.y * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t sid0 = blockIdx.x * LS;
  if (wid >= n_long) return;
  const uint32_t row = long_rows[wid];
  const uint32_t n_here = min((uint32_t)LS, n_sig - sid0);
  const uint32_t my_s = lane >> 2;                      // signature owned by this lane after the reduce-scatter
  const bool owner = (lane & 3) == 0 && my_s < n_here;  // lane 4 s finishes signature s
  const uint32_t* my_z = z_all + (uint64_t)(sid0 + (my_s < n_here ? my_s : 0)) * g.n_z * 8;
  Fr res0, res1, res2;
  uint32_t slow = 0;
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const FastMat& M = g.m[m];
    const uint32_t k0 = M.row_ptr[row], k1 = M.row_ptr[row + 1];
    Fr r = Fr::zero();
    if (k1 - k0 <= 8) {
      if (owner) {
        Lazy lz;
        lz.clear();
        for (uint32_t k = k0; k < k1; k++) serial_term(g, M, k, my_z, xs_t, sid0 + my_s, r, lz);
        if (k1 > k0) r = r + lz.reduce();
      }
    } else {
      Lazy lazy[LS];
      {
        // (1) terms in the full-coefficient form (sorted first): carry-free 64-bit accumulation per limb
        uint64_t acc[LS][8];  // 64-bit accumulator per (signature, limb): an aligned register pair
#pragma unroll
        for (int s = 0; s < LS; s++)
#pragma unroll
          for (int i = 0; i < 8; i++) acc[s][i] = 0;
        const uint32_t kf = M.full_end[row];
        // software pipeline, two deep: the (code, col) pair of term k + 64 and the coefficient / multiplicands of
        // term k + 32 are requested before term k is consumed, so neither level of the dependent loads
        // (indices -> data) is waited for
        uint32_t k = k0 + lane;
        uint32_t code_n = 0, col_n = 0;  // indices of term k + 32
        Fr c = Fr::zero();
        uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;
        if (k < kf) {
          const uint32_t code = M.code[k], col = M.col[k];
          c = load_fr(M.fval + 8 * (uint64_t)(code & CODE_MASK));
          const uint4* xp = reinterpret_cast<const uint4*>(xs_t + (uint64_t)col * g.xs_stride + sid0);
          x0 = xp[0];
          x1 = xp[1];
        }
        if (k + 32 < kf) {
          code_n = M.code[k + 32];
          col_n = M.col[k + 32];
        }
        while (k < kf) {
          const uint32_t kn = k + 32, knn = k + 64;
          Fr cn = Fr::zero();
          uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
          uint32_t code_nn = 0, col_nn = 0;
          if (kn < kf) {
            cn = load_fr(M.fval + 8 * (uint64_t)(code_n & CODE_MASK));
            const uint4* xp = reinterpret_cast<const uint4*>(xs_t + (uint64_t)col_n * g.xs_stride + sid0);
            n0 = xp[0];
            n1 = xp[1];
          }
          if (knn < kf) {
            code_nn = M.code[knn];
            col_nn = M.col[knn];
          }
          const uint32_t xv[LS] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int s = 0; s < LS; s++) {
            const bool big = xv[s] >= SMALL_LIMIT;  // includes NOT_SMALL
            slow |= (big ? 1u : 0u) << s;  // recomputed exactly below; the lazy value is then unused
            const uint32_t x = big ? 0u : xv[s];
#pragma unroll
            for (int i = 0; i < 8; i++)  // one IMAD.WIDE.U32 with 64-bit accumulate
              asm("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.lo.cc.u32 lo, %1, %2, lo; madc.hi.u32 hi, %1, %2, hi; "
                  "mov.b64 %0, {lo, hi};}"
                  : "+l"(acc[s][i])
                  : "r"(x), "r"(c.v[i]));
          }
          c = cn;
          x0 = n0;
          x1 = n1;
          code_n = code_nn;
          col_n = col_nn;
          k = kn;
        }
#pragma unroll
        for (int s = 0; s < LS; s++) {
          uint64_t carry = 0;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            // carry < 2^32 here and acc < 2^64 - 2^32 (at most 2^12 products < 2^52), so the sum fits
            carry += acc[s][i];
            lazy[s].v[i] = (uint32_t)carry;
            carry >>= 32;
          }
          lazy[s].v[8] = (uint32_t)carry;
          lazy[s].v[9] = (uint32_t)(carry >> 32);
        }
        // (2) the remaining terms: small coefficient x full-width multiplicand
        for (uint32_t k = kf + lane; k < k1; k += 32) {
          const uint32_t code = M.code[k], col = M.col[k];
          const uint32_t mag = code & CODE_MASK;
#pragma unroll
          for (int s = 0; s < LS; s++) {
            Fr x = load_fr(z_all + ((uint64_t)(sid0 + (s < (int)n_here ? s : 0)) * g.n_z + col) * 8);
            lazy[s].fma(mag, (code & CODE_NEG) ? neg_fr(x) : x);
          }
        }
      }
      // reduce-scatter: offsets 16, 8, 4 halve the set of accumulators a lane keeps; 2, 1 finish
#pragma unroll
      for (int half = 4, o = 16; half >= 1; half >>= 1, o >>= 1) {
        const bool hi = lane & o;
#pragma unroll
        for (int j = 0; j < half; j++) {
          Lazy send = hi ? lazy[j] : lazy[j + half], other;
#pragma unroll
          for (int i = 0; i < 10; i++) other.v[i] = __shfl_xor_sync(0xffffffffu, send.v[i], o);
          if (hi) lazy[j] = lazy[j + half];
          lazy[j].add(other);
        }
      }
#pragma unroll
      for (int o = 2; o > 0; o >>= 1) {
        Lazy other;
#pragma unroll
        for (int i = 0; i < 10; i++) other.v[i] = __shfl_xor_sync(0xffffffffu, lazy[0].v[i], o);
        lazy[0].add(other);
      }
      if (owner) r = lazy[0].reduce();
    }
    if (m == 0) res0 = r;
    if (m == 1) res1 = r;
    if (m == 2) res2 = r;
  }
  slow = __reduce_or_sync(0xffffffffu, slow);
  // exact fall-back for signatures whose "small" columns are not small (invalid assignments only)
  for (uint32_t s = 0; s < n_here; s++) {
    if (!((slow >> s) & 1)) continue;
    const uint32_t* z = z_all + (uint64_t)(sid0 + s) * g.n_z * 8;
    Fr a = warp_row_dot(g.slow.a_ptr, g.slow.a_col, g.slow.a_val, z, row, lane);
    Fr b = warp_row_dot(g.slow.b_ptr, g.slow.b_col, g.slow.b_val, z, row, lane);
    Fr c = warp_row_dot(g.slow.c_ptr, g.slow.c_col, g.slow.c_val, z, row, lane);
    if (lane == 4 * s) {
      res0 = a;
      res1 = b;
      res2 = c;
    }
  }
  if (owner) {
    const uint64_t o = ((uint64_t)(sid0 + my_s) * g.out_stride + row) * 8;
    if (az) store_fr(az + o, res0);
    if (bz) store_fr(bz + o, res1);
    if (cz) store_fr(cz + o, res2);
    if (first_unsat) {
      bool bad;
      if (is_one(res1))
        bad = res0 != res2;
      else
        bad = res0 * res1 != res2;
      if (bad) atomicMin(first_unsat + sid0 + my_s, (unsigned long long)row);
    }
  }
}

// =============================================================================================
// Signed-digit long rows.  The wide matrix of an inlined NTT row holds *integers* (products of twiddles of the
// lazy butterflies, circuits/falcon_ntt.rs:31-39, gadgets/poly.rs:115-149): |c| < q^10 < 2^136, stored by arkworks
// as c mod r.  Here each is kept as five balanced base-2^32 digits d_i in [-2^31, 2^31), so a term costs
// 5 signed IMAD.WIDE per signature (not 8), the row sum is an exact integer S = sum_i acc_i 2^(32 i), |S| < 2^200,
// and <A_row, z> = mont(S mod r) is bit-identical to the term-by-term field evaluation.
// Record of a term = 8 words: d0..d4, small-column index, 0, 0; the records of a row are contiguous, so the 32 lanes
// of a warp read 1 KB per step.  A warp takes RW rows in turn for a tile of 8 signatures and finishes them together:
// after the reduce-scatter lane 4 s + j owns (row j, signature s), so the per-row epilogue (conversion to
// Montgomery form, the few remaining terms, B and C, the product check) runs on all 32 lanes.
// =============================================================================================
constexpr int RW = 4;
struct SLong {
  const uint32_t *rows, *ptr, *rec, *wide, *limit;  // limit: multiplicands below it cannot overflow the 64-bit sums
  uint32_t n_rows;
};

__device__ __forceinline__ void load8(const uint32_t* p, uint32_t (&w)[8]) {
  uint64_t a, b, c, d;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32);
  w[2] = (uint32_t)b; w[3] = (uint32_t)(b >> 32);
  w[4] = (uint32_t)c; w[5] = (uint32_t)(c >> 32);
  w[6] = (uint32_t)d; w[7] = (uint32_t)(d >> 32);
}
// acc += x * d (signed 32 x 32 -> 64): one IMAD.WIDE with a 64-bit accumulate on an aligned register pair
__device__ __forceinline__ void smad(int64_t& acc, uint32_t x, uint32_t d) {
  asm("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.lo.cc.u32 lo, %1, %2, lo; madc.hi.s32 hi, %1, %2, hi; "
      "mov.b64 %0, {lo, hi};}"
      : "+l"(acc)
      : "r"(x), "r"(d));
}
struct S224 {  // two's complement, 7 words
  uint32_t v[7];
  __device__ __forceinline__ void add(const S224& o) {
    asm("add.cc.u32 %0, %0, %7; addc.cc.u32 %1, %1, %8; addc.cc.u32 %2, %2, %9; addc.cc.u32 %3, %3, %10;"
        "addc.cc.u32 %4, %4, %11; addc.cc.u32 %5, %5, %12; addc.u32 %6, %6, %13;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6])
        : "r"(o.v[0]), "r"(o.v[1]), "r"(o.v[2]), "r"(o.v[3]), "r"(o.v[4]), "r"(o.v[5]), "r"(o.v[6]));
  }
  // the Montgomery image of the integer mod r  (|value| < 2^223 < r)
  __device__ __forceinline__ Fr to_fr() const {
    const bool neg = v[6] >> 31;
    const uint32_t m = neg ? 0xffffffffu : 0u;
    Fr x;
    uint32_t carry = neg ? 1u : 0u;
#pragma unroll
    for (int i = 0; i < 7; i++) {  // magnitude = neg ? ~v + 1 : v
      const uint64_t t = (uint64_t)(v[i] ^ m) + carry;
      x.v[i] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
    x.v[7] = 0;
    x = x.to_mont();
    return neg ? neg_fr(x) : x;
  }
};

// one term of a short matrix row for one signature; +-1 coefficients by modular add / sub
__device__ __forceinline__ void serial_term2(const FastArgs& g, const FastMat& M, uint32_t k, const uint32_t* z,
                                             const uint32_t* xs_t, uint32_t sid, Fr& acc, Lazy& lazy, bool& used) {
  const uint32_t code = M.code[k], col = M.col[k];
  if (code & CODE_FULL) {
    Fr c = load_fr(M.fval + 8 * (uint64_t)(code & CODE_MASK));
    uint32_t x = xs_t[(uint64_t)col * g.xs_stride + sid];
    if (x != NOT_SMALL) {
      lazy.fma(x, c);
      used = true;
    } else {
      acc = acc + c * load_fr(z + 8 * (uint64_t)g.small_cols[col]);
    }
  } else {
    Fr x = load_fr(z + 8 * (uint64_t)col);
    const uint32_t mag = code & CODE_MASK;
    if (mag == 1) {
      acc = (code & CODE_NEG) ? acc - x : acc + x;
    } else {
      lazy.fma(mag, (code & CODE_NEG) ? neg_fr(x) : x);
      used = true;
    }
  }
}

__global__ void __launch_bounds__(LONG_THREADS, 3)
    r1cs_signed_long_kernel(FastArgs g, SLong L, const uint32_t* __restrict__ z_all, const uint32_t* __restrict__ xs_t,
                            uint32_t n_sig, uint32_t* az, uint32_t* bz, uint32_t* cz, unsigned long long* first_unsat) {
  // blockIdx.x = signature tile (fastest-varying): the blocks sharing a group of rows run back to back and find its
  // records in L2
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t r0 = (blockIdx.y * (LONG_THREADS / 32) + (threadIdx.x >> 5)) * RW;
  if (r0 >= L.n_rows) return;
  const uint32_t sid0 = blockIdx.x * LS;
  const uint32_t n_here = min((uint32_t)LS, n_sig - sid0);
  const uint32_t my_s = lane >> 2, my_j = lane & 3;
  S224 mine;
#pragma unroll
  for (int i = 0; i < 7; i++) mine.v[i] = 0;
  uint32_t slow = 0;  // bit 8 j + s: (row j, signature s) has a multiplicand that is not small
#pragma unroll 1
  for (uint32_t j = 0; j < RW; j++) {
    if (r0 + j >= L.n_rows) break;
    const uint32_t k0 = L.ptr[r0 + j], k1 = L.ptr[r0 + j + 1], x_limit = L.limit[r0 + j];
    int64_t acc[LS][5];
#pragma unroll
    for (int s = 0; s < LS; s++)
#pragma unroll
      for (int i = 0; i < 5; i++) acc[s][i] = 0;
    // Software pipeline without register rotation (a copy of a loaded register waits for the load, which would put
    // the wait back into the step that issued it): three record buffers and three multiplicand buffers, the loop
    // unrolled three steps.  In step t the multiplicands of term t + 1 are requested (their column came with a
    // record requested two steps earlier), term t is consumed, and its record buffer is refilled for term t + 3.
    uint32_t k = k0 + lane;
    uint32_t recA[8], recB[8], recC[8];
    uint4 xA0, xA1, xB0, xB1, xC0, xC1;
    uint32_t slow_j = 0;
    auto issue_rec = [&](uint32_t (&rec)[8], uint32_t kk) {
      if (kk < k1) {
        load8(L.rec + 8 * (uint64_t)kk, rec);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) rec[i] = 0;
      }
    };
    auto issue_x = [&](const uint32_t (&rec)[8], uint4& a, uint4& b, uint32_t kk) {
      a = b = make_uint4(0, 0, 0, 0);
      if (kk < k1) {
        const uint4* xp = reinterpret_cast<const uint4*>(xs_t + (uint64_t)rec[5] * g.xs_stride + sid0);
        a = xp[0];
        b = xp[1];
      }
    };
    auto consume = [&](const uint32_t (&rec)[8], const uint4& a, const uint4& b) {
      const uint32_t xv[LS] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int s = 0; s < LS; s++) {
        const bool big = xv[s] >= x_limit;  // includes NOT_SMALL
        slow_j |= (big ? 1u : 0u) << s;     // recomputed exactly below; the integer sum is then unused
        const uint32_t x = big ? 0u : xv[s];
#pragma unroll
        for (int i = 0; i < 5; i++) smad(acc[s][i], x, rec[i]);
      }
    };
    issue_rec(recA, k);
    issue_rec(recB, k + 32);
    issue_rec(recC, k + 64);
    issue_x(recA, xA0, xA1, k);
#pragma unroll 1
    while (k < k1) {
      issue_x(recB, xB0, xB1, k + 32);
      consume(recA, xA0, xA1);
      issue_rec(recA, k + 96);
      issue_x(recC, xC0, xC1, k + 64);
      consume(recB, xB0, xB1);
      issue_rec(recB, k + 128);
      issue_x(recA, xA0, xA1, k + 96);
      consume(recC, xC0, xC1);
      issue_rec(recC, k + 160);
      k += 96;
    }
    slow |= slow_j << (8 * j);
    // per signature: the integer sum of this lane's terms as a 224-bit two's complement number
    S224 t[LS];
#pragma unroll
    for (int s = 0; s < LS; s++) {
      int64_t carry = 0;
#pragma unroll
      for (int i = 0; i < 5; i++) {  // |acc| < 2^62 by the row's multiplicand limit, |carry| < 2^31
        carry += acc[s][i];
        t[s].v[i] = (uint32_t)carry;
        carry >>= 32;
      }
      t[s].v[5] = (uint32_t)carry;
      t[s].v[6] = (uint32_t)(carry >> 32);
    }
    // reduce-scatter over the lanes (offsets 16, 8, 4: lane quad s keeps signature s), then all-reduce inside the quad
#pragma unroll
    for (int half = 4, o = 16; half >= 1; half >>= 1, o >>= 1) {
      const bool hi = lane & o;
#pragma unroll
      for (int q = 0; q < half; q++) {
        S224 send = hi ? t[q] : t[q + half], other;
#pragma unroll
        for (int i = 0; i < 7; i++) other.v[i] = __shfl_xor_sync(0xffffffffu, send.v[i], o);
        if (hi) t[q] = t[q + half];
        t[q].add(other);
      }
    }
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) {
      S224 other;
#pragma unroll
      for (int i = 0; i < 7; i++) other.v[i] = __shfl_xor_sync(0xffffffffu, t[0].v[i], o);
      t[0].add(other);
    }
    if (my_j == j) mine = t[0];
  }
  slow = __reduce_or_sync(0xffffffffu, slow);
  // epilogue: lane 4 s + j finishes (row j, signature s)
  const bool live = r0 + my_j < L.n_rows && my_s < n_here;
  const uint32_t row = L.rows[live ? r0 + my_j : r0];
  const uint32_t sid = sid0 + (live ? my_s : 0);
  Fr res[3];
  if (live) {
    const uint32_t wide = L.wide[r0 + my_j];
    const uint32_t* z = z_all + (uint64_t)sid * g.n_z * 8;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const FastMat& M = g.m[m];
      const uint32_t k1 = M.row_ptr[row + 1];
      const uint32_t kb = (uint32_t)m == wide ? M.full_end[row] : M.row_ptr[row];
      Fr r = Fr::zero();
      Lazy lz;
      lz.clear();
      bool used = false;
      for (uint32_t k = kb; k < k1; k++) serial_term2(g, M, k, z, xs_t, sid, r, lz, used);
      if (used) r = r + lz.reduce();
      if ((uint32_t)m == wide) r = r + mine.to_fr();
      res[m] = r;
    }
  }
  // exact fall-back for assignments whose "small" columns are not small (invalid assignments only)
  while (slow) {
    const uint32_t b = __ffs(slow) - 1;
    slow &= slow - 1;
    const uint32_t j = b >> 3, s = b & 7;
    if (r0 + j >= L.n_rows || s >= n_here) continue;
    const uint32_t srow = L.rows[r0 + j];
    const uint32_t* z = z_all + (uint64_t)(sid0 + s) * g.n_z * 8;
    Fr a = warp_row_dot(g.slow.a_ptr, g.slow.a_col, g.slow.a_val, z, srow, lane);
    Fr bb = warp_row_dot(g.slow.b_ptr, g.slow.b_col, g.slow.b_val, z, srow, lane);
    Fr c = warp_row_dot(g.slow.c_ptr, g.slow.c_col, g.slow.c_val, z, srow, lane);
    if (lane == 4 * s + j) {
      res[0] = a;
      res[1] = bb;
      res[2] = c;
    }
  }
  if (live) {
    const uint64_t o = ((uint64_t)sid * g.out_stride + row) * 8;
    if (az) store_fr(az + o, res[0]);
    if (bz) store_fr(bz + o, res[1]);
    if (cz) store_fr(cz + o, res[2]);
    if (first_unsat) {
      bool bad;
      if (is_one(res[1]))
        bad = res[0] != res[2];
      else
        bad = res[0] * res[1] != res[2];
      if (bad) atomicMin(first_unsat + sid, (unsigned long long)row);
    }
  }
}

// =============================================================================================
// Bundled long rows, for batches of 64 signatures and more: one *lane per signature*.  The rows of one NTT all run
// over the same columns, so four rows with identical column lists form a bundle: a term is then (column, 4 x 5 digits),
// one coalesced 128-byte load of the multiplicands (xs[col][signature .. signature + 31]) feeds 20 multiply-adds per
// lane, and a lane keeps the whole row sums of its two signatures (no reduction over lanes).  The sums go to a scratch
// buffer; r1cs_bundle_finish_kernel turns them into <A_row, z> and does the rest of the row.
// The records of a bundle are staged through shared memory in chunks of 64 terms (cp.async, double buffered) and read
// back as warp-uniform LDS.128; the column offsets sit in shared memory for the 8-terms-ahead multiplicand prefetch.
// Two digit formats:
//  * DBL: balanced base-2^28 digits held as doubles, accumulated with DFMA.  Every partial sum is an integer below 2^53
//    in magnitude (the bundle's multiplicand limit = 2^53 / (T max|d|), T terms), so the arithmetic is exact; DFMA issues
//    at 1.68e13 /s on B200 against 7.0e12 /s for IMAD.WIDE with a 64-bit accumulate (profiles/r01_pipe_rates_ubench.txt).
//    Used for the NTT rows (14-bit multiplicands).  (Keeping 32-bit digits in shared memory and making them doubles in
//    registers with the 2^52 trick halves the LDS.128 stream but adds 3 integer/FP64 instructions per digit: 9 % slower.)
//  * integer: balanced base-2^32 digits, IMAD.WIDE into signed 64-bit sums: for bundles whose multiplicands are too large
//    for the 53-bit budget (the norm decomposition row: squares up to 2^26).
// A coefficient outside the five digits (the constant term of the last NTT layers, ~2^159) is listed as an extra term
// of its row and evaluated with the field-sized coefficient in the epilogue.
// =============================================================================================
constexpr int BR = 4;     // rows per bundle
constexpr int BCH = 64;   // terms per staged chunk
constexpr int BT = 32;    // threads (signatures) per block: one warp
constexpr int BQ = 8;     // multiplicand prefetch distance (terms)
constexpr int BX = 2;     // extra (field-sized) terms per row
constexpr int BS = 2;     // signatures per lane
constexpr int ND = BR * 5;                 // digits per term record
constexpr int QD = ND / 2, QI = (ND + 3) / 4;  // 16-byte words per record: doubles, integers
static_assert(ND % 2 == 0 && (BCH * QD) % BT == 0 && (BCH * QI) % BT == 0, "record staging");
struct Bundles {
  const uint32_t *rows, *ptr, *cols, *wide, *limit, *extra, *dbl;
  const uint64_t* rec_off;  // first 16-byte word of the bundle's records
  const void* rec;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// t += v * 2^SH  (two's complement, 224 bits)
template <int SH>
__device__ __forceinline__ void add_shifted(S224& t, int64_t v) {
  constexpr int W = SH / 32, BIT = SH % 32;
  const uint64_t lo = (uint64_t)v << BIT;
  const uint32_t hi = BIT ? (uint32_t)(v >> (64 - BIT)) : (uint32_t)(v >> 63);  // sign-extended third word
  const uint32_t ext = (uint32_t)(v >> 63);
  const uint32_t w[3] = {(uint32_t)lo, (uint32_t)(lo >> 32), hi};
  uint64_t carry = 0;
#pragma unroll
  for (int i = W; i < 7; i++) {
    carry += (uint64_t)t.v[i] + (i - W < 3 ? w[i - W] : ext);
    t.v[i] = (uint32_t)carry;
    carry >>= 32;
  }
}

// lane-private: a short matrix row (or the terms of the wide matrix that are not in digit form) for one signature
__device__ __forceinline__ Fr serial_row(const FastArgs& g, const FastMat& M, uint32_t kb, uint32_t k1, const uint32_t* z,
                                         const uint32_t* xs_t, uint32_t sid, const uint32_t* extra) {
  Fr r = Fr::zero();
  Lazy lz;
  lz.clear();
  bool used = false;
  for (uint32_t k = kb; k < k1; k++) serial_term2(g, M, k, z, xs_t, sid, r, lz, used);
  if (extra)
    for (int e = 0; e < BX; e++)
      if (extra[e] != 0xffffffffu) serial_term2(g, M, extra[e], z, xs_t, sid, r, lz, used);
  if (used) r = r + lz.reduce();
  return r;
}

// Second pass of the bundled rows: thread per (bundle slot, signature).  Takes the integer row sum left by
// r1cs_bundle_kernel (224-bit two's complement; word 7 != 0 marks a signature whose multiplicands were not small),
// brings it to the field, adds the terms that are not in digit form, evaluates B and C, writes the outputs and checks
// the product.  Kept out of the first pass because its chains of dependent loads need many resident warps to hide.
__global__ void __launch_bounds__(128)
    r1cs_bundle_finish_kernel(FastArgs g, Bundles B, uint32_t n_slots, const uint32_t* __restrict__ sums,
                              const uint32_t* __restrict__ z_all, const uint32_t* __restrict__ xs_t, uint32_t n_sig,
                              uint32_t* az, uint32_t* bz, uint32_t* cz, unsigned long long* first_unsat) {
  const uint32_t sid = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
  if (sid >= n_sig || p >= n_slots) return;
  const uint32_t row = B.rows[p];
  if (row == 0xffffffffu) return;
  const uint32_t wide = B.wide[p / BR];
  const uint32_t* z = z_all + (uint64_t)sid * g.n_z * 8;
  uint32_t w[8];
  load8(sums + ((uint64_t)p * n_sig + sid) * 8, w);
  Fr res[3];
  if (w[7] == 0) {
    S224 t;
#pragma unroll
    for (int i = 0; i < 7; i++) t.v[i] = w[i];
    const Fr wide_val = t.to_fr();
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const FastMat& M = g.m[m];
      const bool is_wide = (uint32_t)m == wide;
      const uint32_t kb = is_wide ? M.full_end[row] : M.row_ptr[row];
      res[m] = serial_row(g, M, kb, M.row_ptr[row + 1], z, xs_t, sid, is_wide ? B.extra + p * BX : nullptr);
      if (is_wide) res[m] = res[m] + wide_val;
    }
  } else {  // exact term-by-term evaluation (invalid assignments only)
    const Fr m1 = Fr::one().neg();
    res[0] = row_dot(g.slow.a_ptr, g.slow.a_col, g.slow.a_val, z, row, m1);
    res[1] = row_dot(g.slow.b_ptr, g.slow.b_col, g.slow.b_val, z, row, m1);
    res[2] = row_dot(g.slow.c_ptr, g.slow.c_col, g.slow.c_val, z, row, m1);
  }
  const uint64_t o = ((uint64_t)sid * g.out_stride + row) * 8;
  if (az) store_fr(az + o, res[0]);
  if (bz) store_fr(bz + o, res[1]);
  if (cz) store_fr(cz + o, res[2]);
  if (first_unsat) {
    bool bad;
    if (is_one(res[1]))
      bad = res[0] != res[2];
    else
      bad = res[0] * res[1] != res[2];
    if (bad) atomicMin(first_unsat + sid, (unsigned long long)row);
  }
}

// one bundle for BT x BS = 64 signatures (lane l: signatures l and l + 32 of the block), digit format DBL
// (BR x BS = 2 x 4 and a prefetch distance of 4 were measured 15 % slower than 4 x 2 and 8)
template <bool DBL>
__device__ __forceinline__ void bundle_run(const FastArgs& g, const Bundles& B, uint4* bsm,
                                           const uint32_t* __restrict__ xs_t, uint32_t n_sig,
                                           uint32_t* __restrict__ sums) {
  using Acc = typename std::conditional<DBL, double, int64_t>::type;
  constexpr int Q = DBL ? QD : QI;  // 16-byte words per term record
  uint4* recbuf = bsm;              // [2][BCH * QD] records, then the bundle's column offsets (+ BQ of padding)
  uint32_t* cols = reinterpret_cast<uint32_t*>(bsm + 2 * BCH * QD);
  const uint32_t b = blockIdx.y, lane = threadIdx.x;
  const uint32_t t0 = B.ptr[b], T = B.ptr[b + 1] - t0;  // a multiple of BQ (zero records as padding)
  uint32_t sid[BS];
  bool valid[BS];
#pragma unroll
  for (int s = 0; s < BS; s++) {
    const uint32_t raw = blockIdx.x * (BT * BS) + s * BT + lane;
    valid[s] = raw < n_sig;
    sid[s] = valid[s] ? raw : n_sig - 1;  // idle slots repeat the last signature and store nothing
  }
  const uint4* src = reinterpret_cast<const uint4*>(B.rec) + B.rec_off[b];
  auto fetch = [&](uint32_t c) {  // (the last chunk of a bundle may run into the next bundle's records: never consumed)
    uint4* dst = recbuf + (c & 1) * (BCH * QD);
    const uint4* s4 = src + (uint64_t)c * (BCH * Q);
#pragma unroll
    for (int i = 0; i < BCH * Q / BT; i++) cp_async16(dst + lane + i * BT, s4 + lane + i * BT);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(0);
  for (uint32_t i = lane; i < T + BQ; i += BT) cols[i] = i < T ? B.cols[t0 + i] * g.xs_stride : 0u;
  __syncwarp();
  const uint32_t x_limit = B.limit[b];
  Acc acc[BS][BR][5];
#pragma unroll
  for (int s = 0; s < BS; s++)
#pragma unroll
    for (int r = 0; r < BR; r++)
#pragma unroll
      for (int i = 0; i < 5; i++) acc[s][r][i] = 0;
  uint32_t xq[BQ][BS];
#pragma unroll
  for (int u = 0; u < BQ; u++)
#pragma unroll
    for (int s = 0; s < BS; s++) xq[u][s] = xs_t[cols[u] + sid[s]];
  bool slow[BS];
#pragma unroll
  for (int s = 0; s < BS; s++) slow[s] = false;
  const uint32_t nch = (T + BCH - 1) / BCH;
#pragma unroll 1
  for (uint32_t c = 0; c < nch; c++) {
    if (c + 1 < nch) {
      fetch(c + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const uint4* rb = recbuf + (c & 1) * (BCH * QD);
    const uint32_t n_here = min((uint32_t)BCH, T - c * BCH);
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < n_here; i0 += BQ) {
#pragma unroll
      for (int u = 0; u < BQ; u++) {
        uint32_t x[BS];
        const uint32_t cn = cols[c * BCH + i0 + u + BQ];
#pragma unroll
        for (int s = 0; s < BS; s++) {
          const uint32_t xv = xq[u][s];
          xq[u][s] = xs_t[cn + sid[s]];
          const bool big = xv >= x_limit;  // includes NOT_SMALL
          slow[s] |= big;                  // recomputed exactly below; the sums are then unused
          x[s] = big ? 0u : xv;
        }
        const uint4* r4 = rb + (i0 + u) * Q;
        if constexpr (DBL) {
          double xd[BS];
#pragma unroll
          for (int s = 0; s < BS; s++) xd[s] = (double)x[s];
#pragma unroll
          for (int q = 0; q < QD; q++) {
            const uint4 w = r4[q];
            const double d0 = __hiloint2double((int)w.y, (int)w.x), d1 = __hiloint2double((int)w.w, (int)w.z);
#pragma unroll
            for (int s = 0; s < BS; s++) {
              acc[s][(2 * q) / 5][(2 * q) % 5] = fma(xd[s], d0, acc[s][(2 * q) / 5][(2 * q) % 5]);
              acc[s][(2 * q + 1) / 5][(2 * q + 1) % 5] = fma(xd[s], d1, acc[s][(2 * q + 1) / 5][(2 * q + 1) % 5]);
            }
          }
        } else {
          uint32_t d[4 * QI];
#pragma unroll
          for (int q = 0; q < QI; q++) {
            const uint4 w = r4[q];
            d[4 * q] = w.x;
            d[4 * q + 1] = w.y;
            d[4 * q + 2] = w.z;
            d[4 * q + 3] = w.w;
          }
#pragma unroll
          for (int s = 0; s < BS; s++)
#pragma unroll
            for (int r = 0; r < BR; r++)
#pragma unroll
              for (int i = 0; i < 5; i++) smad(acc[s][r][i], x[s], d[5 * r + i]);
        }
      }
    }
    __syncwarp();  // the buffer is refilled two chunks later
  }
  // the row sums as 224-bit two's complement integers, for the second pass
  constexpr int DB = DBL ? 28 : 32;  // digit width; |digit sums| < 2^53 (DBL: exact in a double) or 2^63 - 2^33
#pragma unroll  // (rolled loops would index acc[] dynamically and park the accumulators in local memory)
  for (int s = 0; s < BS; s++) {
    if (!valid[s]) continue;
#pragma unroll
    for (int r = 0; r < BR; r++) {
      S224 t;
#pragma unroll
      for (int i = 0; i < 7; i++) t.v[i] = 0;
      add_shifted<0 * DB>(t, (int64_t)acc[s][r][0]);
      add_shifted<1 * DB>(t, (int64_t)acc[s][r][1]);
      add_shifted<2 * DB>(t, (int64_t)acc[s][r][2]);
      add_shifted<3 * DB>(t, (int64_t)acc[s][r][3]);
      add_shifted<4 * DB>(t, (int64_t)acc[s][r][4]);
      uint32_t* out = sums + ((uint64_t)(BR * b + r) * n_sig + sid[s]) * 8;
      *reinterpret_cast<uint4*>(out) = make_uint4(t.v[0], t.v[1], t.v[2], t.v[3]);
      *reinterpret_cast<uint4*>(out + 4) = make_uint4(t.v[4], t.v[5], t.v[6], slow[s] ? 1u : 0u);
    }
  }
}

__global__ void __launch_bounds__(BT)
    r1cs_bundle_kernel(FastArgs g, Bundles B, const uint32_t* __restrict__ xs_t, uint32_t n_sig,
                       uint32_t* __restrict__ sums) {
  extern __shared__ uint4 bsm[];
  if (B.dbl[blockIdx.y])  // uniform over the block
    bundle_run<true>(g, B, bsm, xs_t, n_sig, sums);
  else
    bundle_run<false>(g, B, bsm, xs_t, n_sig, sums);
}

// =============================================================================================
// The rows of an ntt_circuit block through the butterfly network itself.  Row k of a block is
//   < L_k(x) - q t_k - b_k | z_0 | 0 >,   L_k = output k of the unreduced Cooley-Tukey network of gadgets/poly.rs:115-149
// over the inputs x (and the constant column z_0 for the bound constants 2^(l+1) q^(l+2)).  The circuit matrices hold
// L_k inlined, N rows of N + 1 terms with ~136-bit integer coefficients (what the bundle kernels multiply out: 2.1 M
// multiply-adds per signature); the network computes the same N linear forms with (N / 2) log2 N butterflies of one
// field multiplication each (10 k per signature).  All arithmetic is in Fr, so the result is the same field element
// for any z, small or not.  One CTA per (block, signature); values limb-major in shared memory.
// =============================================================================================
struct NttRowBlocks {
  uint32_t in_col0[4], out_col0[4], row0[4];
};
__device__ __forceinline__ Fr lds_limbs(const uint32_t* s, int n, int p) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = s[i * n + p];
  return r;
}
__device__ __forceinline__ void sts_limbs(uint32_t* s, int n, int p, const Fr& x) {
#pragma unroll
  for (int i = 0; i < 8; i++) s[i * n + p] = x.v[i];
}
template <int LOGN>
__global__ void __launch_bounds__(256)
    r1cs_ntt_rows_kernel(NttRowBlocks blk, const uint32_t* __restrict__ tw_mont, const uint32_t* __restrict__ cst_mont,
                         const uint32_t* __restrict__ z_all, uint32_t n_z, uint64_t out_stride, uint32_t* az, uint32_t* bz,
                         uint32_t* cz, unsigned long long* first_unsat) {
  constexpr int N = 1 << LOGN, NT = 256;
  extern __shared__ uint32_t ntt_sm[];  // [8][N]
  const int tid = threadIdx.x;
  const uint32_t b = blockIdx.x, sid = blockIdx.y;
  const uint32_t* z = z_all + (uint64_t)sid * n_z * 8;
  const Fr z0 = load_fr(z);
  for (int j = tid; j < N; j += NT) sts_limbs(ntt_sm, N, j, load_fr(z + (uint64_t)(blk.in_col0[b] + j) * 8));
  __syncthreads();
  int t = N;
#pragma unroll 1
  for (int l = 0; l < LOGN; l++) {
    const int ht = t >> 1;
    const Fr cl = load_fr(cst_mont + 8 * l) * z0;  // 2^(l+1) q^(l+2) on the constant column
    for (int idx = tid; idx < N / 2; idx += NT) {
      const int i = idx / ht, j = idx - i * ht;
      const int p0 = i * t + j, p1 = p0 + ht;
      const Fr s = load_fr(tw_mont + 8 * ((1 << l) + i));
      const Fr u = lds_limbs(ntt_sm, N, p0), v = lds_limbs(ntt_sm, N, p1) * s;
      sts_limbs(ntt_sm, N, p0, u + v);         // out[j]      = u + v
      sts_limbs(ntt_sm, N, p1, u + (cl - v));  // out[j + ht] = u + (const[l+1] - v)
    }
    t = ht;
    __syncthreads();
  }
  const Fr q = load_fr(cst_mont + 8 * LOGN);  // q
  for (int k = tid; k < N; k += NT) {
    const uint32_t* w = z + (uint64_t)(blk.out_col0[b] + 29 * k) * 8;
    const Fr a = lds_limbs(ntt_sm, N, k) - q * load_fr(w) - load_fr(w + 8);
    const uint32_t row = blk.row0[b] + 30 * k;
    const uint64_t o = ((uint64_t)sid * out_stride + row) * 8;
    if (az) store_fr(az + o, a);
    if (bz) store_fr(bz + o, z0);
    if (cz) store_fr(cz + o, Fr::zero());
    if (first_unsat && !a.is_zero() && !z0.is_zero()) atomicMin(first_unsat + sid, (unsigned long long)row);
  }
}

// canonical values of the small columns of every signature, transposed: xs_t[col][signature]
__global__ void __launch_bounds__(256)
    small_view_kernel(const uint32_t* __restrict__ z_all, const uint32_t* __restrict__ small_cols, uint32_t n_small,
                      uint32_t n_z, uint32_t n_sig, uint32_t xs_stride, uint32_t* __restrict__ xs_t) {
  const uint32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = blockIdx.y;
  if (sid >= xs_stride) return;
  uint32_t out = NOT_SMALL;
  if (sid < n_sig) {
    Fr x = load_fr(z_all + ((uint64_t)sid * n_z + small_cols[i]) * 8).from_mont();
    uint32_t hi = 0;
#pragma unroll
    for (int k = 1; k < 8; k++) hi |= x.v[k];
    if (hi == 0 && x.v[0] < VIEW_LIMIT) out = x.v[0];
  }
  xs_t[(uint64_t)i * xs_stride + sid] = out;
}

int32_t upload_terms(const circuit::HostCSR& h, const std::vector<int64_t>& small_index, DevTerms* d) {
  using circuit::U256;
  const size_t nnz = h.col.size();
  std::vector<uint32_t> col(nnz + 1), code(nnz + 1), fval;
  auto small = [](const U256& x) {
    for (int i = 1; i < 8; i++)
      if (x.v[i]) return false;
    return x.v[0] < CODE_FULL && x.v[0] != 0;
  };
  // These tables serve the warp-per-row kernel.  A term on a "small" column always takes the
  // (small multiplicand) x (full-width coefficient) form, whatever its coefficient, and within a row those
  // terms come first, so that the 32 lanes striding over a dense NTT row take the same branch in all but
  // the last one or two iterations (the order of the terms of a dot product is immaterial).
  const size_t n_rows = h.row_ptr.size() - 1;
  std::vector<uint32_t> full_end(n_rows + 1, 0);
  size_t k = 0;
  for (size_t r = 0; r < n_rows; r++) {
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1) full_end[r] = (uint32_t)k;
      for (uint32_t e = h.row_ptr[r]; e < h.row_ptr[r + 1]; e++) {
        const U256& c = h.val[e];
        U256 n = circuit::fr_neg(c);
        const bool on_small = small_index[h.col[e]] >= 0;
        const bool as_full = on_small || (!small(c) && !small(n));
        if (as_full != (pass == 0)) continue;
        if (as_full) {
          col[k] = (uint32_t)small_index[h.col[e]];
          code[k] = CODE_FULL | (uint32_t)(fval.size() / 8);
          for (int i = 0; i < 8; i++) fval.push_back(c.v[i]);
        } else if (small(c)) {
          col[k] = h.col[e];
          code[k] = c.v[0];
        } else {
          col[k] = h.col[e];
          code[k] = CODE_NEG | n.v[0];
        }
        k++;
      }
    }
  }
  FRCS_CUDA_CHECK(cudaMalloc(&d->full_end, full_end.size() * 4));
  FRCS_CUDA_CHECK(cudaMemcpy(d->full_end, full_end.data(), full_end.size() * 4, cudaMemcpyHostToDevice));
  d->nnz = nnz;
  d->n_full = fval.size() / 8;
  FRCS_CUDA_CHECK(cudaMalloc(&d->row_ptr, h.row_ptr.size() * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->col, (nnz + 1) * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->code, (nnz + 1) * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->fval, (fval.size() + 8) * 4));
  FRCS_CUDA_CHECK(cudaMemcpy(d->row_ptr, h.row_ptr.data(), h.row_ptr.size() * 4, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMemcpy(d->col, col.data(), nnz * 4, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMemcpy(d->code, code.data(), nnz * 4, cudaMemcpyHostToDevice));
  if (!fval.empty()) FRCS_CUDA_CHECK(cudaMemcpy(d->fval, fval.data(), fval.size() * 4, cudaMemcpyHostToDevice));
  return FRCS_OK;
}

// five balanced base-2^32 digits of the integer behind c (c or c - r, whichever is small); false if it has none
static bool signed_digits(const circuit::U256& c, uint32_t d[5]) {
  const circuit::U256 n = circuit::fr_neg(c);
  auto fits = [](const circuit::U256& x) { return x.v[5] == 0 && x.v[6] == 0 && x.v[7] == 0; };
  uint32_t w[8];
  if (fits(c)) {
    for (int i = 0; i < 8; i++) w[i] = c.v[i];
  } else if (fits(n)) {  // two's complement of the magnitude
    uint64_t carry = 1;
    for (int i = 0; i < 8; i++) {
      carry += (uint64_t)(uint32_t)~n.v[i];
      w[i] = (uint32_t)carry;
      carry >>= 32;
    }
  } else {
    return false;
  }
  for (int i = 0; i < 5; i++) {
    const int32_t di = (int32_t)w[0];
    d[i] = (uint32_t)di;
    // w = (w - di) >> 32 (arithmetic) = (w >> 32) + (di < 0)
    uint64_t carry = di < 0 ? 1 : 0;
    const uint32_t ext = (w[7] >> 31) ? 0xffffffffu : 0u;
    for (int k = 0; k < 8; k++) {
      carry += (uint64_t)(k < 7 ? w[k + 1] : ext);
      w[k] = (uint32_t)carry;
      carry >>= 32;
    }
  }
  for (int i = 0; i < 8; i++)
    if (w[i]) return false;
  return true;
}

// five balanced base-2^bits digits (each in [-2^(bits-1), 2^(bits-1))) of the integer behind c (c or c - r); false if
// the integer does not fit
static bool balanced_digits(const circuit::U256& c, int bits, int64_t d[5]) {
  const circuit::U256 n = circuit::fr_neg(c);
  auto fits = [](const circuit::U256& x) { return x.v[5] == 0 && x.v[6] == 0 && x.v[7] == 0; };
  uint32_t w[8];
  if (fits(c)) {
    for (int i = 0; i < 8; i++) w[i] = c.v[i];
  } else if (fits(n)) {
    uint64_t carry = 1;
    for (int i = 0; i < 8; i++) {
      carry += (uint64_t)(uint32_t)~n.v[i];
      w[i] = (uint32_t)carry;
      carry >>= 32;
    }
  } else {
    return false;
  }
  const uint64_t mask = (1ull << bits) - 1, half = 1ull << (bits - 1);
  for (int i = 0; i < 5; i++) {
    const uint64_t low = (((uint64_t)w[1] << 32) | w[0]) & mask;
    const int64_t di = low >= half ? (int64_t)low - (int64_t)(1ull << bits) : (int64_t)low;
    d[i] = di;
    // w = (w - di) >> bits (arithmetic):  subtract the sign-extended digit, then shift
    uint64_t borrow_in = 0;
    const uint64_t sub_lo = (uint64_t)di;             // two's complement of the digit, low 64 bits
    const uint32_t sub_ext = di < 0 ? 0xffffffffu : 0u;  // its sign extension
    uint32_t t[8];
    for (int k = 0; k < 8; k++) {
      const uint64_t sk = k == 0 ? (uint32_t)sub_lo : k == 1 ? (uint32_t)(sub_lo >> 32) : sub_ext;
      const uint64_t lhs = w[k], rhs = sk + borrow_in;
      t[k] = (uint32_t)(lhs - rhs);
      borrow_in = lhs < rhs ? 1 : 0;
    }
    const uint32_t ext = (t[7] >> 31) ? 0xffffffffu : 0u;
    for (int k = 0; k < 8; k++) {
      const uint32_t lo = t[k], hi = k < 7 ? t[k + 1] : ext;
      w[k] = bits == 32 ? hi : (lo >> bits) | (hi << (32 - bits));
    }
  }
  for (int i = 0; i < 8; i++)
    if (w[i]) return false;
  return true;
}

// splits the long rows into signed-digit rows and generic ones and uploads the digit records
// unbounded[i]: small column i was added only because an all-integer row uses it (third rule of build_fast_r1cs); unlike
// the columns that meet field-sized coefficients nothing bounds its values to ~14 bits
static int32_t build_signed_long(frcs_ctx* ctx, const circuit::Matrices& m, const std::vector<int64_t>& small_index,
                                 const std::vector<uint8_t>& unbounded) {
  const circuit::HostCSR* hs[3] = {&m.a, &m.b, &m.c};
  std::vector<uint32_t> sl_rows, sl_ptr{0}, sl_rec, sl_wide, sl_limit, gl_rows;

  for (uint32_t r : ctx->long_rows_host) {
    int wide = -1, n_wide = 0;
    for (int k = 0; k < 3; k++)
      if (hs[k]->row_ptr[r + 1] - hs[k]->row_ptr[r] > 8) {
        wide = k;
        n_wide++;
      }
    bool ok = n_wide == 1;
    uint32_t lim = 0;
    std::vector<uint32_t> rec;
    if (ok) {
      const circuit::HostCSR& h = *hs[wide];
      uint32_t n_full = 0, n_rest = 0;
      uint64_t max_d = 1;
      for (uint32_t e = h.row_ptr[r]; e < h.row_ptr[r + 1] && ok; e++) {
        if (small_index[h.col[e]] < 0) {
          n_rest++;
          continue;
        }
        // a coefficient just beyond the digit range (the constant term of the last NTT layers, ~2^159.4) is entered
        // as two records of half the size on the same column
        circuit::U256 parts[2] = {h.val[e], h.val[e]};
        int n_parts = 1;
        uint32_t d[5];
        if (!signed_digits(h.val[e], d)) {
          const circuit::U256 neg = circuit::fr_neg(h.val[e]);
          const bool is_neg = !(h.val[e].v[5] == 0 && h.val[e].v[6] == 0 && h.val[e].v[7] == 0);
          const circuit::U256& mag = is_neg ? neg : h.val[e];
          circuit::U256 half;
          for (int i = 0; i < 8; i++) half.v[i] = (mag.v[i] >> 1) | (i < 7 ? mag.v[i + 1] << 31 : 0u);
          const circuit::U256 rest = circuit::u256_sub(mag, half);
          parts[0] = is_neg ? circuit::fr_neg(half) : half;
          parts[1] = is_neg ? circuit::fr_neg(rest) : rest;
          n_parts = 2;
        }
        for (int p = 0; p < n_parts && ok; p++) {
          ok = signed_digits(parts[p], d);
          for (int i = 0; i < 5; i++) {
            rec.push_back(d[i]);
            const int64_t v = (int32_t)d[i];
            max_d = std::max<uint64_t>(max_d, (uint64_t)(v < 0 ? -v : v));
          }
          rec.push_back((uint32_t)small_index[h.col[e]]);
          rec.push_back(0);
          rec.push_back(0);
          n_full++;
        }
      }
      // a lane adds ceil(n_full / 32) products |d| x into a signed 64-bit sum: x below `lim` keeps it under 2^62
      const uint64_t per_lane = (n_full + 31) / 32;
      const uint64_t cap = per_lane ? (1ull << 62) / (max_d * per_lane) : VIEW_LIMIT;
      lim = 1;
      while (lim < VIEW_LIMIT && 2ull * lim <= cap) lim *= 2;
      ok = ok && n_rest <= 8 && lim >= (1u << 14);
    }
    if (!ok) {
      if (getenv("FRCS_DEBUG"))
        fprintf(stderr, "long row %u stays generic: %d wide matrices, %zu digit records, limit %u\n", r, n_wide, rec.size() / 8, lim);
      gl_rows.push_back(r);
      continue;
    }
    sl_rows.push_back(r);
    sl_wide.push_back((uint32_t)wide);
    sl_limit.push_back(lim);
    sl_rec.insert(sl_rec.end(), rec.begin(), rec.end());
    sl_ptr.push_back((uint32_t)(sl_rec.size() / 8));
  }
  ctx->n_sl_rows = (uint32_t)sl_rows.size();
  ctx->n_sl_rest = 0;  // long_rows_host lists the rows of the ntt_circuit blocks last
  for (uint32_t r : sl_rows) {
    bool blk = false;
    for (const circuit::NttBlock& b : m.ntt_blocks) blk |= r >= b.row0 && (r - b.row0) % 30 == 0 && (r - b.row0) / 30 < m.L.n;
    if (blk) break;
    ctx->n_sl_rest++;
  }
  ctx->n_gl_rows = (uint32_t)gl_rows.size();
  ctx->n_gl_rest = 0;
  for (uint32_t r : gl_rows) {
    bool blk = false;
    for (const circuit::NttBlock& b : m.ntt_blocks) blk |= r >= b.row0 && (r - b.row0) % 30 == 0 && (r - b.row0) / 30 < m.L.n;
    if (blk) break;
    ctx->n_gl_rest++;
  }
  auto up = [](uint32_t** d, const std::vector<uint32_t>& v, size_t pad) -> cudaError_t {
    cudaError_t e = cudaMalloc(d, (v.size() + pad) * 4);
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpy(*d, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
  };
  FRCS_CUDA_CHECK(up(&ctx->sl_rows, sl_rows, 1));
  FRCS_CUDA_CHECK(up(&ctx->sl_ptr, sl_ptr, 1));
  FRCS_CUDA_CHECK(up(&ctx->sl_rec, sl_rec, 8));
  FRCS_CUDA_CHECK(up(&ctx->sl_wide, sl_wide, 1));
  FRCS_CUDA_CHECK(up(&ctx->sl_limit, sl_limit, 1));
  FRCS_CUDA_CHECK(up(&ctx->gl_rows, gl_rows, 1));
  if (getenv("FRCS_DEBUG"))
    fprintf(stderr, "long rows: %u signed-digit (%zu records), %u generic\n", ctx->n_sl_rows, sl_rec.size() / 8, ctx->n_gl_rows);
  // Bundles of BR rows with identical (wide matrix, column list), terms sorted by column; two sets (digit formats).
  {
    struct Term { uint32_t col, k; int64_t d28[5], d32[5]; bool ok28, ok32; };
    struct RowInfo { std::vector<Term> terms; uint32_t wide; };
    std::vector<RowInfo> info(sl_rows.size());
    for (size_t i = 0; i < sl_rows.size(); i++) {
      const uint32_t r = sl_rows[i];
      const circuit::HostCSR& h = *hs[sl_wide[i]];
      info[i].wide = sl_wide[i];
      uint32_t j = 0;  // position among the row's terms on small columns = its offset in the term tables (upload_terms)
      for (uint32_t e = h.row_ptr[r]; e < h.row_ptr[r + 1]; e++) {
        if (small_index[h.col[e]] < 0) continue;
        Term t;
        t.col = (uint32_t)small_index[h.col[e]];
        t.k = h.row_ptr[r] + j++;
        t.ok28 = balanced_digits(h.val[e], 28, t.d28);
        t.ok32 = balanced_digits(h.val[e], 32, t.d32);
        info[i].terms.push_back(t);
      }
    }
    size_t covered = 0;
    std::vector<uint32_t> bd_rows, bd_ptr{0}, bd_cols, bd_wide, bd_limit, bd_extra, bd_dbl;
    std::vector<uint64_t> bd_off;   // in 16-byte words
    std::vector<uint32_t> rec;      // digit records of every bundle (4-byte integers or 8-byte doubles), 16-byte aligned
    uint32_t max_terms = 0;
    auto in_ntt_block = [&](uint32_t r) {
      for (const circuit::NttBlock& b : m.ntt_blocks)
        if (r >= b.row0 && (r - b.row0) % 30 == 0 && (r - b.row0) / 30 < m.L.n) return true;
      return false;
    };
    uint32_t n_rest_bundles = 0;
    // pass 0: rows outside the ntt_circuit blocks, pass 1: the block rows (skipped at run time when
    // r1cs_ntt_rows_kernel takes them); within a pass first the double-digit bundles, then the integer ones
    for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) n_rest_bundles = (uint32_t)bd_wide.size();
    for (int set = 1; set >= 0; set--) {  // 1: double digits, 0: integer digits
      const bool dbl = set == 1;
      // rows of this set: the double format if its 53-bit budget leaves room for 14-bit multiplicands, else integers
      std::map<std::pair<uint32_t, std::vector<uint32_t>>, std::vector<uint32_t>> groups;
      std::vector<std::vector<Term>> digit_terms(sl_rows.size());
      std::vector<std::array<uint32_t, BX>> extras(sl_rows.size());
      std::vector<uint64_t> row_cap(sl_rows.size(), 0);
      for (size_t i = 0; i < sl_rows.size(); i++) {
        if (in_ntt_block(sl_rows[i]) != (pass == 1)) continue;
        auto cap_of = [&](bool use28, std::vector<Term>& out, std::array<uint32_t, BX>& ex) -> uint64_t {
          out.clear();
          ex.fill(0xffffffffu);
          uint32_t n_ex = 0;
          uint64_t max_d = 1;
          for (const Term& t : info[i].terms) {
            if (use28 ? t.ok28 : t.ok32) {
              out.push_back(t);
              for (int q = 0; q < 5; q++) {
                const int64_t v = use28 ? t.d28[q] : t.d32[q];
                max_d = std::max<uint64_t>(max_d, (uint64_t)(v < 0 ? -v : v));
              }
            } else {
              if (n_ex == BX) return 0;
              ex[n_ex++] = t.k;
            }
          }
          if (out.empty() || out.size() > 8192) return 0;
          const uint64_t budget = use28 ? (1ull << 53) - 1 : (1ull << 63) - (1ull << 33);
          return std::min<uint64_t>(budget / (max_d * out.size()), VIEW_LIMIT);
        };
        std::vector<Term> t28, t32;
        std::array<uint32_t, BX> e28, e32;
        const uint64_t c28 = cap_of(true, t28, e28), c32 = cap_of(false, t32, e32);
        bool to_dbl = c28 >= (1u << 14) && !getenv("FRCS_NO_DBL");
        for (const Term& t : info[i].terms) to_dbl = to_dbl && !unbounded[t.col];
        if (dbl != to_dbl) continue;
        if (!dbl && c32 < (1u << 14)) continue;  // neither format: the row is not covered, bundles stay unused
        digit_terms[i] = dbl ? t28 : t32;
        extras[i] = dbl ? e28 : e32;
        row_cap[i] = dbl ? c28 : c32;
        std::stable_sort(digit_terms[i].begin(), digit_terms[i].end(), [](const Term& x, const Term& y) { return x.col < y.col; });
        std::vector<uint32_t> key;
        for (const Term& t : digit_terms[i]) key.push_back(t.col);
        groups[{info[i].wide, key}].push_back((uint32_t)i);
        covered++;
      }
      for (auto& kv : groups) {
        const std::vector<uint32_t>& members = kv.second;
        const uint32_t T = (uint32_t)kv.first.second.size();
        const uint32_t T_pad = (T + BQ - 1) / BQ * BQ;
        for (size_t m0 = 0; m0 < members.size(); m0 += BR) {
          uint64_t cap = VIEW_LIMIT;
          for (int r = 0; r < BR; r++) {
            const bool have = m0 + r < members.size();
            bd_rows.push_back(have ? sl_rows[members[m0 + r]] : 0xffffffffu);
            for (int e = 0; e < BX; e++) bd_extra.push_back(have ? extras[members[m0 + r]][e] : 0xffffffffu);
            if (have) cap = std::min(cap, row_cap[members[m0 + r]]);
          }
          bd_off.push_back(rec.size() / 4);
          for (uint32_t t = 0; t < T_pad; t++) {
            bd_cols.push_back(t < T ? kv.first.second[t] : 0u);
            for (int r = 0; r < BR; r++)
              for (int q = 0; q < 5; q++) {
                const bool have = t < T && m0 + r < members.size();
                if (dbl) {
                  const double d = have ? (double)digit_terms[members[m0 + r]][t].d28[q] : 0.0;
                  uint64_t bits;
                  memcpy(&bits, &d, 8);
                  rec.push_back((uint32_t)bits);
                  rec.push_back((uint32_t)(bits >> 32));
                } else {
                  rec.push_back(have ? (uint32_t)(int32_t)digit_terms[members[m0 + r]][t].d32[q] : 0u);
                }
              }
            if (!dbl)
              for (int q = ND; q < 4 * QI; q++) rec.push_back(0u);
          }
          bd_ptr.push_back((uint32_t)bd_cols.size());
          bd_wide.push_back(kv.first.first);
          bd_limit.push_back((uint32_t)cap);
          bd_dbl.push_back(dbl ? 1u : 0u);
          max_terms = std::max(max_terms, T_pad);
        }
      }
      if (getenv("FRCS_DEBUG"))
        fprintf(stderr, "bundles (%s digits): %zu so far, %zu column lists, %zu padded terms so far\n", dbl ? "double" : "integer",
                bd_wide.size(), groups.size(), bd_cols.size());
    }
    }
    DevBundles& D = ctx->bd;
    D.n = (uint32_t)bd_wide.size();
    D.n_rest = n_rest_bundles;
    D.max_terms = max_terms;
    if (D.n) {
      FRCS_CUDA_CHECK(up(&D.rows, bd_rows, 4));
      FRCS_CUDA_CHECK(up(&D.ptr, bd_ptr, 1));
      FRCS_CUDA_CHECK(up(&D.cols, bd_cols, 8));
      FRCS_CUDA_CHECK(up(&D.wide, bd_wide, 1));
      FRCS_CUDA_CHECK(up(&D.limit, bd_limit, 1));
      FRCS_CUDA_CHECK(up(&D.extra, bd_extra, 4));
      FRCS_CUDA_CHECK(up(&D.dbl, bd_dbl, 1));
      FRCS_CUDA_CHECK(cudaMalloc(&D.rec_off, bd_off.size() * 8));
      FRCS_CUDA_CHECK(cudaMemcpy(D.rec_off, bd_off.data(), bd_off.size() * 8, cudaMemcpyHostToDevice));
      // records, padded by one chunk (the staging of a bundle's last chunk reads a whole chunk)
      const size_t bytes = rec.size() * 4, pad = (size_t)BCH * QD * 16;
      FRCS_CUDA_CHECK(cudaMalloc(&D.rec, bytes + pad));
      FRCS_CUDA_CHECK(cudaMemset(D.rec, 0, bytes + pad));
      FRCS_CUDA_CHECK(cudaMemcpy(D.rec, rec.data(), bytes, cudaMemcpyHostToDevice));
    }
    ctx->bundles_usable = covered == sl_rows.size() && !sl_rows.empty();
  }
  return FRCS_OK;
}


// Plan of r1cs_stream_kernel: windows of SW columns, every short row in the window of its highest column, columns
// outside the window as "remote" slots.  Not used (stream_usable = false) when a short row carries a field-sized
// coefficient or a window would need too many remote columns (the schoolbook products: both factors are far away).
static int32_t build_stream_plan(frcs_ctx* ctx, const circuit::Matrices& m, const std::vector<uint32_t>& bm,
                                 const std::vector<uint32_t>& hdr, const std::vector<uint32_t>& mt) {
  ctx->stream_usable = false;
  const uint32_t n_z = m.L.n_z, n_win = (n_z + SW - 1) / SW;
  struct Row {
    uint32_t row, k0, cnt[3];
  };
  std::vector<std::vector<Row>> rows_of(n_win);
  for (uint32_t r = 0; r < m.L.n_cons; r++) {
    if ((bm[r >> 5] >> (r & 31)) & 1) continue;
    Row R{r, hdr[2 * r], {hdr[2 * r + 1] & 0x7fu, (hdr[2 * r + 1] >> 7) & 0x7fu, (hdr[2 * r + 1] >> 14) & 0x7fu}};
    uint32_t maxcol = 0;
    for (uint32_t e = 0; e < R.cnt[0] + R.cnt[1] + R.cnt[2]; e++) {
      if (mt[2 * (R.k0 + e) + 1] & CODE_FULL) return FRCS_OK;  // a field-sized coefficient: keep the table-driven kernels
      maxcol = std::max(maxcol, mt[2 * (R.k0 + e)]);
    }
    rows_of[maxcol / SW].push_back(R);
  }
  std::vector<StreamWin> wins(n_win);
  std::vector<uint8_t> blob;
  uint32_t max_remote = 0, max_desc = 0, max_pm1 = 0;
  for (uint32_t w = 0; w < n_win; w++) {
    const uint32_t lo = w * SW, nc = std::min(SW, n_z - lo);
    std::map<uint32_t, uint32_t> remote_slot;
    std::vector<uint32_t> remote;
    auto ref = [&](uint32_t col) -> uint32_t {
      if (col >= lo && col < lo + nc) return col - lo;
      auto it = remote_slot.find(col);
      if (it == remote_slot.end()) {
        it = remote_slot.emplace(col, SW + (uint32_t)remote.size()).first;
        remote.push_back(col);
      }
      return it->second;
    };
    std::vector<std::pair<uint32_t, std::array<uint32_t, 4>>> pm1;       // (class key, descriptor)
    std::vector<std::pair<uint32_t, Row>> gen;                            // (class key, row)
    for (const Row& R : rows_of[w]) {
      std::array<uint32_t, 6> slot;
      slot.fill(SREF_NONE);
      bool is_pm1 = true;
      uint32_t k = R.k0;
      for (int mm = 0; mm < 3 && is_pm1; mm++) {
        for (uint32_t e = 0; e < R.cnt[mm] && is_pm1; e++) {
          const uint32_t code = mt[2 * (k + e) + 1];
          uint32_t& sl = slot[2 * mm + ((code & CODE_NEG) ? 1 : 0)];
          if ((code & CODE_MASK) != 1 || sl != SREF_NONE) is_pm1 = false;
          sl = 0;  // occupied (the reference is resolved below)
        }
        k += R.cnt[mm];
      }
      if (is_pm1) {
        slot.fill(SREF_NONE);
        k = R.k0;
        uint32_t key = 0;
        for (int mm = 0; mm < 3; mm++) {
          for (uint32_t e = 0; e < R.cnt[mm]; e++) {
            const uint32_t code = mt[2 * (k + e) + 1];
            slot[2 * mm + ((code & CODE_NEG) ? 1 : 0)] = ref(mt[2 * (k + e)]);
          }
          k += R.cnt[mm];
        }
        for (int i = 0; i < 6; i++) key |= (slot[i] != SREF_NONE ? 1u : 0u) << i;
        pm1.push_back({key, {R.row, slot[0] | (slot[1] << 16), slot[2] | (slot[3] << 16), slot[4] | (slot[5] << 16)}});
      } else {
        gen.push_back({R.cnt[0] | (R.cnt[1] << 8) | (R.cnt[2] << 16), R});
      }
    }
    std::stable_sort(pm1.begin(), pm1.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    std::stable_sort(gen.begin(), gen.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    std::vector<uint32_t> gen_hdr, terms;
    for (auto& kr : gen) {
      const Row& R = kr.second;
      gen_hdr.insert(gen_hdr.end(), {R.row, (uint32_t)(terms.size() / 2), kr.first, 0u});
      for (uint32_t e = 0; e < R.cnt[0] + R.cnt[1] + R.cnt[2]; e++) {
        terms.push_back(ref(mt[2 * (R.k0 + e)]));
        terms.push_back(mt[2 * (R.k0 + e) + 1]);
      }
    }
    if (remote.size() > (size_t)STREAM_THREADS) return FRCS_OK;  // too much outside the window: not a streaming circuit
    {  // absent terms reference the zero slot behind the remote slots
      const uint32_t zs = SW + (uint32_t)remote.size();
      for (auto& kr : pm1)
        for (int q = 1; q < 4; q++) {
          uint32_t lo16 = kr.second[q] & 0xffffu, hi16 = kr.second[q] >> 16;
          if (lo16 == SREF_NONE) lo16 = zs;
          if (hi16 == SREF_NONE) hi16 = zs;
          kr.second[q] = lo16 | (hi16 << 16);
        }
    }
    StreamWin& W = wins[w];
    W.col_lo = lo;
    W.n_cols = nc;
    W.n_remote = (uint32_t)remote.size();
    W.n_pm1 = (uint32_t)pm1.size();
    W.n_gen = (uint32_t)gen.size();
    W.n_terms = (uint32_t)(terms.size() / 2);
    W.pad = 0;
    W.desc_off = blob.size();
    auto put = [&](const uint32_t* p, size_t n) {
      const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
      blob.insert(blob.end(), b, b + 4 * n);
    };
    remote.resize((remote.size() + 3) & ~(size_t)3, 0u);
    put(remote.data(), remote.size());
    for (auto& kr : pm1) put(kr.second.data(), 4);
    put(gen_hdr.data(), gen_hdr.size());
    if (terms.size() % 4) terms.insert(terms.end(), 4 - terms.size() % 4, 0u);
    put(terms.data(), terms.size());
    W.desc_bytes = (uint32_t)(blob.size() - W.desc_off);
    max_remote = std::max(max_remote, W.n_remote);
    max_desc = std::max(max_desc, W.desc_bytes);
    max_pm1 = std::max(max_pm1, W.n_pm1);
  }
  ctx->stream_slots = (SW + max_remote + 1 + 15) & ~15u;  // + the zero slot
  ctx->stream_desc_max = max_desc;
  ctx->n_stream_win = n_win;
  if (max_pm1 > 0xffffu) return FRCS_OK;  // queue entries are 16-bit row indices
  ctx->stream_qmax = (max_pm1 + 8) & ~7u;
  const size_t smem = 16 + 2 * (size_t)((ctx->stream_slots + 15) & ~15u) + 3 * (size_t)ctx->stream_qmax * 2 + max_desc;
  if (smem > (size_t)(226 / FRCS_STREAM_CTAS) * 1024) return FRCS_OK;  // FRCS_STREAM_CTAS CTAs per SM must fit
  FRCS_CUDA_CHECK(cudaMalloc(&ctx->stream_wins, wins.size() * sizeof(StreamWin)));
  FRCS_CUDA_CHECK(cudaMemcpy(ctx->stream_wins, wins.data(), wins.size() * sizeof(StreamWin), cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMalloc(&ctx->stream_desc, blob.size() + 16));
  FRCS_CUDA_CHECK(cudaMemcpy(ctx->stream_desc, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  ctx->stream_smem = smem;
  ctx->stream_usable = true;
  if (getenv("FRCS_DEBUG"))
    fprintf(stderr, "stream plan: %u windows of %u columns, <= %u remote columns, program <= %u B, %zu B smem per CTA\n", n_win,
            SW, max_remote, max_desc, smem);
  return FRCS_OK;
}

}  // namespace

// Test hook (host only, no GPU): the digit decompositions behind the long-row kernels.
extern "C" int32_t frcs_debug_digits(const uint64_t* coeffs, uint64_t n, int32_t bits, int64_t* digits, int32_t* ok) {
  if (!coeffs || !digits || !ok || (bits != 0 && bits != 28 && bits != 32)) {
    frcs_set_error("frcs_debug_digits: null buffer or bits not in {0, 28, 32}");
    return FRCS_E_INVALID_ARG;
  }
  for (uint64_t i = 0; i < n; i++) {
    circuit::U256 c;
    for (int k = 0; k < 4; k++) {
      c.v[2 * k] = (uint32_t)coeffs[4 * i + k];
      c.v[2 * k + 1] = (uint32_t)(coeffs[4 * i + k] >> 32);
    }
    int64_t d[5] = {0, 0, 0, 0, 0};
    bool good;
    if (bits == 0) {  // the 32-bit records of r1cs_signed_long_kernel
      uint32_t w[5] = {0, 0, 0, 0, 0};
      good = signed_digits(c, w);
      for (int k = 0; k < 5; k++) d[k] = (int32_t)w[k];
    } else {
      good = balanced_digits(c, bits, d);
    }
    ok[i] = good ? 1 : 0;
    for (int k = 0; k < 5; k++) digits[5 * i + k] = good ? d[k] : 0;
  }
  return FRCS_OK;
}

// Classifies the coefficients of A, B, C (canonical on the host) and uploads the term tables.
int32_t build_fast_r1cs(frcs_ctx* ctx, const circuit::Matrices& m) {
  using circuit::U256;
  // columns that meet a full-width coefficient anywhere
  std::vector<int64_t> small_index(m.L.n_z, -1);
  std::vector<uint32_t> small_cols;
  for (const circuit::HostCSR* h : {&m.a, &m.b, &m.c})
    for (size_t k = 0; k < h->col.size(); k++) {
      const U256& c = h->val[k];
      U256 n = circuit::fr_neg(c);
      bool cs = true, ns = true;
      for (int i = 1; i < 8; i++) {
        cs &= c.v[i] == 0;
        ns &= n.v[i] == 0;
      }
      cs &= c.v[0] < CODE_FULL;
      ns &= n.v[0] < CODE_FULL;
      if (!cs && !ns && small_index[h->col[k]] < 0) {
        small_index[h->col[k]] = (int64_t)small_cols.size();
        small_cols.push_back(h->col[k]);
      }
    }
  // ... and columns used by many long rows (the inputs of an NTT whose coefficients all happen to be small):
  // inside the dense rows they should take the same form as their neighbours
  {
    std::vector<uint32_t> uses(m.L.n_z, 0);
    for (const circuit::HostCSR* h : {&m.a, &m.b, &m.c})
      for (uint32_t r : ctx->long_rows_host)
        for (uint32_t e = h->row_ptr[r]; e < h->row_ptr[r + 1]; e++) uses[h->col[e]]++;
    for (uint32_t col = 0; col < m.L.n_z; col++)
      if (uses[col] > 64 && small_index[col] < 0) {
        small_index[col] = (int64_t)small_cols.size();
        small_cols.push_back(col);
      }
  }
  // ... and the columns of a wide long row that has only integer coefficients but more than a handful of terms off the
  // small columns (the norm-bound decomposition row: 2N squares below 2^26 and the bits of the norm): as small columns
  // the row qualifies for the signed-digit kernel.  Values that turn out not to be small only cost the exact fall-back.
  std::vector<uint8_t> unbounded;
  {
    std::vector<std::pair<uint32_t, const circuit::HostCSR*>> cand;  // (row, its wide all-integer matrix)
    std::vector<uint8_t> counted(m.L.n_z, 0);
    size_t extra_cols = 0;
    for (uint32_t r : ctx->long_rows_host)
      for (const circuit::HostCSR* h : {&m.a, &m.b, &m.c}) {
        if (h->row_ptr[r + 1] - h->row_ptr[r] <= 8) continue;
        uint32_t off = 0;
        bool integers = true;
        for (uint32_t e = h->row_ptr[r]; e < h->row_ptr[r + 1] && integers; e++) {
          uint32_t d[5];
          integers = signed_digits(h->val[e], d);
          off += small_index[h->col[e]] < 0;
        }
        if (!integers || off <= 8) continue;
        cand.push_back({r, h});
        for (uint32_t e = h->row_ptr[r]; e < h->row_ptr[r + 1]; e++)
          if (small_index[h->col[e]] < 0 && !counted[h->col[e]]) {
            counted[h->col[e]] = 1;
            extra_cols++;
          }
      }
    // all such rows or none (the schoolbook circuit has N of them over N^2 product columns: the view would not pay)
    if (small_cols.size() + extra_cols <= 16384)
      for (auto& rh : cand)
        for (uint32_t e = rh.second->row_ptr[rh.first]; e < rh.second->row_ptr[rh.first + 1]; e++)
          if (small_index[rh.second->col[e]] < 0) {
            small_index[rh.second->col[e]] = (int64_t)small_cols.size();
            small_cols.push_back(rh.second->col[e]);
            unbounded.resize(small_cols.size(), 0);
            unbounded.back() = 1;
          }
  }
  unbounded.resize(small_cols.size(), 0);
  int32_t rc;
  if ((rc = upload_terms(m.a, small_index, &ctx->TA)) || (rc = upload_terms(m.b, small_index, &ctx->TB)) ||
      (rc = upload_terms(m.c, small_index, &ctx->TC)))
    return rc;
  for (DevTerms* t : {&ctx->TA, &ctx->TB, &ctx->TC})
    if ((rc = launch_to_montgomery(ctx, t->fval, t->n_full, ctx->stream))) return rc;
  if ((rc = build_signed_long(ctx, m, small_index, unbounded))) return rc;
  if ((rc = ensure_mont_table(ctx, ctx->stream))) return rc;
  ctx->n_small = (uint32_t)small_cols.size();
  if (getenv("FRCS_DEBUG")) {
    size_t nf = 0, ns = 0;
    for (uint32_t r : ctx->long_rows_host)
      for (uint32_t e = m.a.row_ptr[r]; e < m.a.row_ptr[r + 1]; e++) (small_index[m.a.col[e]] >= 0 ? nf : ns)++;
    fprintf(stderr, "fast r1cs: %zu small columns, long rows %zu: A terms on small columns %zu, others %zu, TA full %llu\n",
            small_cols.size(), ctx->long_rows_host.size(), nf, ns, (unsigned long long)ctx->TA.n_full);
  }
  FRCS_CUDA_CHECK(cudaMalloc(&ctx->small_cols, (small_cols.size() + 1) * 4));
  FRCS_CUDA_CHECK(cudaMemcpy(ctx->small_cols, small_cols.data(), small_cols.size() * 4, cudaMemcpyHostToDevice));
  // bitmap of the long rows
  std::vector<uint32_t> bm((m.L.n_cons + 31) / 32 + 1, 0);
  for (uint32_t r : ctx->long_rows_host) bm[r >> 5] |= 1u << (r & 31);
  FRCS_CUDA_CHECK(cudaMalloc(&ctx->is_long, bm.size() * 4));
  FRCS_CUDA_CHECK(cudaMemcpy(ctx->is_long, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice));
  // merged term list of the short rows (A, B, C terms of a row contiguous) + one header per row
  {
    std::vector<uint32_t> hdr(2 * (size_t)m.L.n_cons + 2, 0), mt, mf;
    const circuit::HostCSR* hs[3] = {&m.a, &m.b, &m.c};
    for (uint32_t r = 0; r < m.L.n_cons; r++) {
      if ((bm[r >> 5] >> (r & 31)) & 1) continue;
      hdr[2 * r] = (uint32_t)(mt.size() / 2);
      uint32_t packed = 0;
      for (int k = 0; k < 3; k++) {
        const circuit::HostCSR& h = *hs[k];
        uint32_t cnt = h.row_ptr[r + 1] - h.row_ptr[r];
        packed |= cnt << (7 * k);
        for (uint32_t e = h.row_ptr[r]; e < h.row_ptr[r + 1]; e++) {
          const U256& c = h.val[e];
          U256 n = circuit::fr_neg(c);
          auto small = [](const U256& x) {
            for (int i = 1; i < 8; i++)
              if (x.v[i]) return false;
            return x.v[0] < CODE_FULL && x.v[0] != 0;
          };
          if (small(c)) {
            mt.push_back(h.col[e]);
            mt.push_back(c.v[0]);
          } else if (small(n)) {
            mt.push_back(h.col[e]);
            mt.push_back(CODE_NEG | n.v[0]);
          } else {
            mt.push_back((uint32_t)small_index[h.col[e]]);
            mt.push_back(CODE_FULL | (uint32_t)(mf.size() / 8));
            for (int i = 0; i < 8; i++) mf.push_back(c.v[i]);
          }
        }
      }
      hdr[2 * r + 1] = packed;
    }
    // rows made of +-1 terms only, at most one of each sign per matrix, go to r1cs_pm1_kernel
    std::vector<uint8_t> is_pm1(m.L.n_cons, 0);
    {
      std::vector<std::pair<uint32_t, std::array<uint32_t, 8>>> rows;  // (class key, descriptor)
      for (uint32_t r = 0; r < m.L.n_cons; r++) {
        if ((bm[r >> 5] >> (r & 31)) & 1) continue;
        std::array<uint32_t, 8> d;
        d.fill(PM1_NONE);
        d[0] = r;
        uint32_t k = hdr[2 * r], key = 0;
        bool ok = true;
        for (int mm = 0; mm < 3 && ok; mm++) {
          uint32_t cnt = (hdr[2 * r + 1] >> (7 * mm)) & 0x7f;
          for (uint32_t e = 0; e < cnt && ok; e++) {
            const uint32_t col = mt[2 * (k + e)], code = mt[2 * (k + e) + 1];
            if ((code & CODE_FULL) || (code & CODE_MASK) != 1) ok = false;
            uint32_t& slot = d[1 + 2 * mm + ((code & CODE_NEG) ? 1 : 0)];
            if (slot != PM1_NONE) ok = false;  // two terms of the same sign
            slot = col;
          }
          k += cnt;
        }
        if (!ok) continue;
        for (int i = 1; i < 7; i++) key |= (d[i] != PM1_NONE ? 1u : 0u) << i;
        is_pm1[r] = 1;
        rows.push_back({key, d});
      }
      std::stable_sort(rows.begin(), rows.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
      std::vector<uint32_t> flat;
      flat.reserve(rows.size() * 8 + 8);
      for (auto& kr : rows) flat.insert(flat.end(), kr.second.begin(), kr.second.end());
      ctx->n_pm1_rows = (uint32_t)rows.size();
      FRCS_CUDA_CHECK(cudaMalloc(&ctx->r_pm1, (flat.size() + 8) * 4));
      if (!flat.empty()) FRCS_CUDA_CHECK(cudaMemcpy(ctx->r_pm1, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice));
    }
    // class order of the other short rows: key = term counts and which matrices take the lazy path
    std::vector<std::pair<uint64_t, uint32_t>> keyed;
    for (uint32_t r = 0; r < m.L.n_cons; r++) {
      if (((bm[r >> 5] >> (r & 31)) & 1) || is_pm1[r]) continue;
      uint64_t key = hdr[2 * r + 1];
      uint32_t k = hdr[2 * r];
      for (int mm = 0; mm < 3; mm++) {
        uint32_t cnt = (hdr[2 * r + 1] >> (7 * mm)) & 0x7f;
        bool lazy = false;
        for (uint32_t e = 0; e < cnt; e++) {
          uint32_t code = mt[2 * (k + e) + 1];
          lazy |= (code & CODE_FULL) || (code & CODE_MASK) != 1;
        }
        k += cnt;
        key |= (uint64_t)lazy << (32 + mm);
      }
      keyed.push_back({key, r});
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<uint64_t, uint32_t>& x, const std::pair<uint64_t, uint32_t>& y) { return x.first < y.first; });
    std::vector<uint32_t> perm;
    for (auto& kr : keyed) perm.push_back(kr.second);
    ctx->n_short_rows = (uint32_t)perm.size();
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->r_perm, (perm.size() + 1) * 4));
    FRCS_CUDA_CHECK(cudaMemcpy(ctx->r_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice));
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->r_hdr, hdr.size() * 4));
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->r_mterm, (mt.size() + 2) * 4));
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->r_mfval, (mf.size() + 8) * 4));
    FRCS_CUDA_CHECK(cudaMemcpy(ctx->r_hdr, hdr.data(), hdr.size() * 4, cudaMemcpyHostToDevice));
    FRCS_CUDA_CHECK(cudaMemcpy(ctx->r_mterm, mt.data(), mt.size() * 4, cudaMemcpyHostToDevice));
    if (!mf.empty()) {
      FRCS_CUDA_CHECK(cudaMemcpy(ctx->r_mfval, mf.data(), mf.size() * 4, cudaMemcpyHostToDevice));
      if ((rc = launch_to_montgomery(ctx, ctx->r_mfval, mf.size() / 8, ctx->stream))) return rc;
    }
    if ((rc = build_stream_plan(ctx, m, bm, hdr, mt))) return rc;
  }
  // tables of r1cs_ntt_rows_kernel: twiddles, bound constants 2^(l+1) q^(l+2) (l < LOG_N), then q
  ctx->ntt_rows_usable = false;
  if (!m.ntt_blocks.empty() && m.ntt_blocks.size() <= 4) {
    const uint32_t n = m.L.n, logn = m.L.logn;
    std::vector<uint32_t> tab = circuit::ntt_table(n), tw(8 * (size_t)n, 0), cst(8 * (size_t)(logn + 1), 0);
    for (uint32_t i = 0; i < n; i++) tw[8 * i] = tab[i];
    for (uint32_t l = 0; l < logn; l++) {
      circuit::U256 c = circuit::u256_pow2(l + 1);
      for (uint32_t e = 0; e < l + 2; e++) c = circuit::u256_mul_small(c, circuit::Q);
      for (int k = 0; k < 8; k++) cst[8 * l + k] = c.v[k];
    }
    cst[8 * logn] = circuit::Q;
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->ntt_tw_mont, tw.size() * 4));
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->ntt_cst_mont, cst.size() * 4));
    FRCS_CUDA_CHECK(cudaMemcpy(ctx->ntt_tw_mont, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice));
    FRCS_CUDA_CHECK(cudaMemcpy(ctx->ntt_cst_mont, cst.data(), cst.size() * 4, cudaMemcpyHostToDevice));
    if ((rc = launch_to_montgomery(ctx, ctx->ntt_tw_mont, n, ctx->stream))) return rc;
    if ((rc = launch_to_montgomery(ctx, ctx->ntt_cst_mont, logn + 1, ctx->stream))) return rc;
    ctx->ntt_rows_usable = true;
  }
  return FRCS_OK;
}

void free_fast_r1cs(frcs_ctx* ctx) {
  for (DevTerms* t : {&ctx->TA, &ctx->TB, &ctx->TC}) {
    cudaFree(t->row_ptr);
    cudaFree(t->col);
    cudaFree(t->code);
    cudaFree(t->fval);
    cudaFree(t->full_end);
  }
  cudaFree(ctx->small_cols);
  cudaFree(ctx->is_long);
  cudaFree(ctx->xs);
  cudaFree(ctx->r_perm);
  cudaFree(ctx->r_pm1);
  cudaFree(ctx->sl_rows);
  cudaFree(ctx->sl_ptr);
  cudaFree(ctx->sl_rec);
  cudaFree(ctx->sl_wide);
  cudaFree(ctx->sl_limit);
  {
    DevBundles& D = ctx->bd;
    cudaFree(D.rows);
    cudaFree(D.ptr);
    cudaFree(D.cols);
    cudaFree(D.wide);
    cudaFree(D.limit);
    cudaFree(D.extra);
    cudaFree(D.dbl);
    cudaFree(D.rec_off);
    cudaFree(D.rec);
    cudaFree(ctx->bd_sums);
  }
  cudaFree(ctx->gl_rows);
  cudaFree(ctx->r_hdr);
  cudaFree(ctx->r_mterm);
  cudaFree(ctx->r_mfval);
  cudaFree(ctx->stream_wins);
  cudaFree(ctx->stream_desc);
}

int32_t launch_r1cs_eval(frcs_ctx* ctx, uint64_t n, const uint64_t* d_z, uint64_t* d_az, uint64_t* d_bz,
                         uint64_t* d_cz, int64_t* d_first_unsat, cudaStream_t st, uint64_t out_stride) {
  if (n == 0) return FRCS_OK;
  NvtxRange nvtx("frcs:r1cs_eval");
  if (out_stride == 0) out_stride = ctx->L.n_cons;
  auto fm = [](const DevTerms& t) { return FastMat{t.row_ptr, t.col, t.code, t.fval, t.full_end}; };
  // gridDim.y is limited to 65535: chunk the batch; the small-column view lives in a per-context buffer
  const uint64_t CH = 16384;
  const uint32_t xs_stride = (uint32_t)(((n < CH ? n : CH) + 7) & ~7ull);
  FastArgs g{{fm(ctx->TA), fm(ctx->TB), fm(ctx->TC)},
             (const uint2*)ctx->r_hdr,
             (const uint2*)ctx->r_mterm,
             ctx->r_mfval,
             ctx->small_cols,
             ctx->n_small,
             ctx->L.n_cons,
             ctx->L.n_z,
             xs_stride,
             ctx->mont_tab,
             out_stride,
             EvalArgs{ctx->A.row_ptr, ctx->A.col, ctx->A.val, ctx->B.row_ptr, ctx->B.col, ctx->B.val, ctx->C.row_ptr,
                      ctx->C.col, ctx->C.val, ctx->L.n_cons, ctx->L.n_z, out_stride}};
  unsigned long long* fu = (unsigned long long*)d_first_unsat;
  const size_t need = (size_t)xs_stride * (ctx->n_small + 1) * 4;
  if (ctx->xs_bytes < need) {
    FRCS_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(ctx->xs);
    ctx->xs = nullptr;
    ctx->xs_bytes = 0;
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->xs, need));
    ctx->xs_bytes = need;
  }
  int ph = prof_begin(ctx, PROF_R1CS, st);
  if (fu) {
    init_unsat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fu, n);
    ctx->launches++;
  }
  for (uint64_t s0 = 0; s0 < n; s0 += CH) {
    unsigned ny = (unsigned)(n - s0 < CH ? n - s0 : CH);
    const uint32_t* z = (const uint32_t*)d_z + s0 * ctx->L.n_z * 8;
    uint64_t oo = s0 * out_stride * 8;
    uint32_t* az = d_az ? (uint32_t*)d_az + oo : nullptr;
    uint32_t* bz = d_bz ? (uint32_t*)d_bz + oo : nullptr;
    uint32_t* cz = d_cz ? (uint32_t*)d_cz + oo : nullptr;
    if (ctx->n_small)
      small_view_kernel<<<dim3((xs_stride + 255) / 256, ctx->n_small), 256, 0, st>>>(z, ctx->small_cols, ctx->n_small,
                                                                                     ctx->L.n_z, ny, xs_stride, ctx->xs);
    if (ctx->n_small) ctx->launches++;
    static const bool no_stream = getenv("FRCS_NO_STREAM") != nullptr;
    const bool stream = ctx->stream_usable && !no_stream;
    if (stream) {
      // signatures per CTA: enough CTAs for ~8 waves of 2 per SM, at least 8 signatures per program load
      int sms = 0;
      FRCS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
      uint64_t per = ((uint64_t)ny * ctx->n_stream_win + 16ull * sms - 1) / (16ull * sms);
      per = std::max<uint64_t>(per, std::min<uint64_t>(ny, 8));
      per = std::min<uint64_t>(per, 64);
      // (fitting the run length to a whole number of waves -- 592 signatures: 2340 CTAs = 3.95 waves instead of 2496 =
      // 4.2 -- was measured neutral, 383 k checks/s either way: the windows' programs differ too much in length)
      StreamArgs sa{(const StreamWin*)ctx->stream_wins, (const uint8_t*)ctx->stream_desc, ctx->mont_tab, ctx->L.n_z,
                    ctx->stream_slots, ctx->stream_desc_max, ctx->stream_qmax, out_stride};
      FRCS_CUDA_CHECK(cudaFuncSetAttribute(r1cs_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->stream_smem));
      r1cs_stream_kernel<<<dim3((unsigned)((ny + per - 1) / per), ctx->n_stream_win), STREAM_THREADS, ctx->stream_smem, st>>>(
          sa, z, ny, (uint32_t)per, az, bz, cz, fu ? fu + s0 : nullptr);
      ctx->launches++;
    }
    if (ctx->n_pm1_rows && !stream) {
      r1cs_pm1_kernel<<<dim3((ctx->n_pm1_rows + 255) / 256, ny), 256, 0, st>>>(
          (const uint4*)ctx->r_pm1, ctx->n_pm1_rows, z, ctx->L.n_z, out_stride, az, bz, cz, fu ? fu + s0 : nullptr);
      ctx->launches++;
    }
    // (running the two short-row kernels over sub-chunks of 8..64 signatures, so that the second finds the bits in L2,
    // was measured 7-33 % slower than whole-batch launches)
    // (a variant of the kernel below with four terms' loads in flight per thread ran at the same speed: it re-reads
    // 2.2 GB of scattered 32-byte sectors per 592 signatures at 3.4 TB/s, which is what bounds it)
    if (ctx->n_short_rows && !stream) {
      r1cs_fast_short_kernel<<<dim3((ctx->n_short_rows + 255) / 256, (ny + SS - 1) / SS), 256, 0, st>>>(
          g, ctx->r_perm, ctx->n_short_rows, z, ctx->xs, ny, az, bz, cz, fu ? fu + s0 : nullptr);
      ctx->launches++;
    }
    // the rows of the ntt_circuit blocks through the butterfly network (FRCS_NO_NTT_ROWS=1: through the generic
    // long-row kernels below, like every other long row)
    const bool ntt_fast = ctx->ntt_rows_usable && getenv("FRCS_NO_NTT_ROWS") == nullptr;
    if (ntt_fast) {
      NttRowBlocks nb;
      for (size_t i = 0; i < ctx->ntt_blocks.size(); i++) {
        nb.in_col0[i] = ctx->ntt_blocks[i].in_col0;
        nb.out_col0[i] = ctx->ntt_blocks[i].out_col0;
        nb.row0[i] = ctx->ntt_blocks[i].row0;
      }
      const dim3 gn((unsigned)ctx->ntt_blocks.size(), ny);
      const size_t smem_n = (size_t)32 * ctx->L.n;
      if (ctx->L.logn == 10)
        r1cs_ntt_rows_kernel<10><<<gn, 256, smem_n, st>>>(nb, ctx->ntt_tw_mont, ctx->ntt_cst_mont, z, ctx->L.n_z, out_stride, az,
                                                         bz, cz, fu ? fu + s0 : nullptr);
      else
        r1cs_ntt_rows_kernel<9><<<gn, 256, smem_n, st>>>(nb, ctx->ntt_tw_mont, ctx->ntt_cst_mont, z, ctx->L.n_z, out_stride, az,
                                                        bz, cz, fu ? fu + s0 : nullptr);
      ctx->launches++;
    }
    static const bool no_bundles = getenv("FRCS_NO_BUNDLES") != nullptr;
    const uint32_t n_bundles = ntt_fast ? ctx->bd.n_rest : ctx->bd.n;
    const uint32_t n_sl = ntt_fast ? ctx->n_sl_rest : ctx->n_sl_rows;
    // bundles pay when there are many of them (one warp per bundle and 64 signatures); a handful of long rows (the
    // norm decomposition row once the NTT blocks are gone) is spread over more warps by the warp-per-row kernel
    if (ctx->bundles_usable && ny >= 64 && !no_bundles && n_bundles * BR > 16) {
      if (n_bundles) {
        const DevBundles& D = ctx->bd;
        Bundles bd{D.rows, D.ptr, D.cols, D.wide, D.limit, D.extra, D.dbl, D.rec_off, D.rec};
        const size_t smem = 2 * BCH * QD * sizeof(uint4) + (D.max_terms + BQ) * sizeof(uint32_t);
        const size_t need_sums = (size_t)D.n * BR * ny * 32;
        if (ctx->bd_sums_bytes < need_sums) {
          FRCS_CUDA_CHECK(cudaDeviceSynchronize());
          cudaFree(ctx->bd_sums);
          ctx->bd_sums = nullptr;
          ctx->bd_sums_bytes = 0;
          FRCS_CUDA_CHECK(cudaMalloc(&ctx->bd_sums, need_sums));
          ctx->bd_sums_bytes = need_sums;
        }
        r1cs_bundle_kernel<<<dim3((ny + BT * BS - 1) / (BT * BS), n_bundles), BT, smem, st>>>(g, bd, ctx->xs, ny, ctx->bd_sums);
        r1cs_bundle_finish_kernel<<<dim3((ny + 127) / 128, n_bundles * BR), 128, 0, st>>>(
            g, bd, n_bundles * BR, ctx->bd_sums, z, ctx->xs, ny, az, bz, cz, fu ? fu + s0 : nullptr);
        ctx->launches += 2;
      }
    } else if (n_sl) {
      const uint32_t rows_per_block = RW * (LONG_THREADS / 32);
      dim3 g2((ny + LS - 1) / LS, (n_sl + rows_per_block - 1) / rows_per_block);
      SLong sl{ctx->sl_rows, ctx->sl_ptr, ctx->sl_rec, ctx->sl_wide, ctx->sl_limit, n_sl};
      r1cs_signed_long_kernel<<<g2, LONG_THREADS, 0, st>>>(g, sl, z, ctx->xs, ny, az, bz, cz, fu ? fu + s0 : nullptr);
      ctx->launches++;
    }
    const uint32_t n_gl = ntt_fast ? ctx->n_gl_rest : ctx->n_gl_rows;
    if (n_gl) {
      dim3 g2((ny + LS - 1) / LS, (n_gl * 32 + LONG_THREADS - 1) / LONG_THREADS);
      r1cs_fast_long_kernel<<<g2, LONG_THREADS, 0, st>>>(g, ctx->gl_rows, n_gl, z, ctx->xs, ny, az, bz, cz,
                                                fu ? fu + s0 : nullptr);
      ctx->launches++;
    }
  }
  prof_end(ctx, ph, st);
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

// generic y = M x for three CSR matrices sharing the row space (used by the set-up on the
// transposed circuit matrices); long_rows: rows with > 64 non-zeros in any of the three
int32_t launch_matvec3(frcs_ctx* ctx, const DevCSR* m, uint32_t n_rows, const uint32_t* d_long, uint32_t n_long,
                       const uint32_t* d_x, uint32_t* ya, uint32_t* yb, uint32_t* yc, cudaStream_t st) {
  EvalArgs g{m[0].row_ptr, m[0].col, m[0].val, m[1].row_ptr, m[1].col, m[1].val,
             m[2].row_ptr, m[2].col, m[2].val, n_rows,       0,        n_rows};
  r1cs_short_kernel<<<dim3((n_rows + 255) / 256, 1), 256, 0, st>>>(g, d_x, ya, yb, yc, nullptr);
  ctx->launches++;
  if (n_long) {
    r1cs_long_kernel<<<dim3((n_long * 32 + 255) / 256, 1), 256, 0, st>>>(g, d_long, n_long, d_x, ya, yb, yc, nullptr);
    ctx->launches++;
  }
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}
