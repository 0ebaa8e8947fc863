// Subsystem (4): G1 / G2 multi-scalar multiplication, sum_i s_i * P_i.
// Replaces ark-ec 0.3.0's VariableBaseMSM::multi_scalar_mul ([EXT]),
// called five times by ark-groth16's create_proof (examples/pok_sig.rs:32).
//
// B200-first design (not arkworks' windowed Pippenger):
//  * bases are fixed per proving key, and HBM is 180 GB: for every base P and window k we
//    keep 2^(16k) P in affine form (16 windows, 96 B each for G1).  All windows then share
//    ONE set of 2^15 buckets, no per-window doubling pass exists, and the whole MSM is
//    n*16 mixed additions + one bucket reduction.
//  * signed 16-bit digits (|d| <= 2^15), counting sort of (digit, point) pairs by bucket
//    with warp-aggregated atomics;
//  * bucket accumulation by repeated slicing: every bucket's list is cut into slices of
//    <= Lc entries, one thread per slice (perfect load balance even when one bucket holds
//    half the points, as with the 0/1-heavy witness scalars), slice sums are reduced the
//    same way until one point per bucket remains;
//  * bucket reduction sum_b (b+1) B_b by running sums over 8-bucket runs + a tree (32768 buckets), or by a
//    suffix scan + tree inside one block (128 buckets);
//  * two window geometries (msm.hpp): 16-bit digits for dense scalars, 8-bit digits for the 0/1-heavy
//    assignment z, where the bucket reduction would otherwise dominate.
// Field arithmetic: 12 x u32 Montgomery, IMAD.WIDE carry chains (ff32.cuh).
// Included by msm_g1.cu (MSM_FIELD = ff::Fq, inlined multiplications) and msm_g2.cu
// (MSM_FIELD = ff::Fq2, out-of-line multiplications to bound code size).
#pragma once
#include <cstring>
#include "ctx.hpp"
#include "nvtx.hpp"
#include "ec.cuh"
#include "msm.hpp"

using namespace ff;

namespace {

// window geometry (msm.hpp): CB-bit signed digits, NB = 2^(CB-1) buckets (|digit| - 1), 256 / CB windows
template <int CB_>
struct Geo {
  static constexpr int CB = CB_;
  static constexpr uint32_t NB = 1u << (CB_ - 1);
  static constexpr uint32_t WINDOWS = 256u / CB_;
  static_assert(32 % CB_ == 0, "a digit must not straddle a 32-bit limb");
};
typedef Geo<MSM_CB_WIDE> Wide;
typedef Geo<MSM_CB_NARROW> Narrow;

// Batch geometry.  A "sort problem" is one scalar vector (digits, sorted entries, level plan);
// an "accumulation problem" is one (base table, sort problem) pair: acc problem q uses table
// q / n_sort and sort problem q % n_sort, so several MSMs over the same scalars (the a, b_g1 and
// b_g2 queries of Groth16 all take the assignment z) share one sort.
struct BatchStrides {
  uint64_t sort;    // u32 words between consecutive sort problems (all sort arrays)
  uint64_t acc;     // u32 words between consecutive accumulation problems (slice sums, partials)
  uint32_t n_sort;  // sort problems in the batch
};
struct Tables {
  const uint32_t* pts[2];
};
// scalar sources of one sort problem: up to three consecutive segments
struct ScalarSegs {
  const uint32_t* ptr[3];
  uint64_t stride[3];  // u32 words between problems
  uint64_t end[3];     // cumulative element counts
  const uint32_t* skip;  // optional bitmask over the elements: scalars that count as zero (see msm_sort)
};

static uint32_t msm_env_u32(const char* name, uint32_t dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : -1;
  return v >= 0 ? (uint32_t)v : dflt;
}

// Software prefetch of the points a thread will need a few additions from now.  The accumulation kernels gather 96-byte
// points from a table of hundreds of MB at ~3 warps per scheduler: without it every addition waits for HBM.
// mode 1: into L1 and L2, 2: into L2 only, 0: off.  (pf = mode | distance << 8)
__device__ __forceinline__ void prefetch_line(const void* p, uint32_t mode) {
  if (mode == 1)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
  else if (mode == 2)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// words [0, nw) from a: the first and the last word's cache lines
__device__ __forceinline__ void prefetch_words(const uint32_t* a, int nw, uint32_t mode) {
  prefetch_line(a, mode);
  if ((((uintptr_t)a) & 127u) + 4u * nw > 128u) prefetch_line(a + nw - 1, mode);
}

template <class F>
struct Words {
  static constexpr int N = sizeof(F) / 4;
};

template <class F>
__device__ __forceinline__ F ld_field(const uint32_t* p) {
  F r;
  uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < Words<F>::N; i += 4) {
    uint4 v = *reinterpret_cast<const uint4*>(p + i);
    w[i] = v.x;
    w[i + 1] = v.y;
    w[i + 2] = v.z;
    w[i + 3] = v.w;
  }
  return r;
}
template <class F>
__device__ __forceinline__ void st_field(uint32_t* p, const F& x) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
  for (int i = 0; i < Words<F>::N; i += 4) *reinterpret_cast<uint4*>(p + i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
}
template <class F>
__device__ __forceinline__ ec::Affine<F> ld_affine(const uint32_t* p) {
  return {ld_field<F>(p), ld_field<F>(p + Words<F>::N)};
}
template <class F>
__device__ __forceinline__ void st_affine(uint32_t* p, const ec::Affine<F>& a) {
  st_field<F>(p, a.x);
  st_field<F>(p + Words<F>::N, a.y);
}
template <class F>
__device__ __forceinline__ ec::XYZZ<F> ld_xyzz(const uint32_t* p) {
  constexpr int W = Words<F>::N;
  return {ld_field<F>(p), ld_field<F>(p + W), ld_field<F>(p + 2 * W), ld_field<F>(p + 3 * W)};
}
template <class F>
__device__ __forceinline__ void st_xyzz(uint32_t* p, const ec::XYZZ<F>& a) {
  constexpr int W = Words<F>::N;
  st_field<F>(p, a.x);
  st_field<F>(p + W, a.y);
  st_field<F>(p + 2 * W, a.zz);
  st_field<F>(p + 3 * W, a.zzz);
}

// ---- base pre-processing: pts[k][i] = 2^(CB k) P_i, affine --------------------------------
template <class F, class G>
__global__ void __launch_bounds__(128) precompute_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ pts, uint64_t n) {
  constexpr int AW = 2 * Words<F>::N;
  constexpr uint32_t WINDOWS = G::WINDOWS;
  // windows 1 .. WINDOWS-1 in at most two chains of <= 16 (bounds the per-thread scratch of the batched inversion)
  constexpr int FIRST = WINDOWS - 1 > 16 ? 16 : (int)WINDOWS - 1, SECOND = (int)WINDOWS - 1 - FIRST;
  static_assert(SECOND <= FIRST, "two chains cover at most 32 windows");
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  ec::Affine<F> p = ld_affine<F>(in + i * AW);
  st_affine<F>(pts + i * AW, p);
  if (p.is_inf()) {
    for (uint32_t k = 1; k < WINDOWS; k++) st_affine<F>(pts + (k * n + i) * AW, p);
    return;
  }
  ec::Affine<F> win[FIRST];
  ec::window_multiples<F, FIRST + 1, G::CB>(p, win);
  for (uint32_t k = 1; k <= (uint32_t)FIRST; k++) st_affine<F>(pts + (k * n + i) * AW, win[k - 1]);
  if constexpr (SECOND > 0) {
    const ec::Affine<F> mid = win[FIRST - 1];
    ec::window_multiples<F, SECOND + 1, G::CB>(mid, win);
    for (uint32_t k = 1; k <= (uint32_t)SECOND; k++) st_affine<F>(pts + ((FIRST + k) * n + i) * AW, win[k - 1]);
  }
}

// ---- scalars -> signed digits + histogram ----------------------------------------------
__device__ __forceinline__ void warp_agg_inc(uint32_t* counters, uint32_t key, bool active, uint32_t* pos_out) {
  // warp-aggregated atomicAdd(counters[key], 1) for the active lanes; returns each lane's slot
  unsigned am = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  unsigned peers = __match_any_sync(am, key);
  int leader = __ffs(peers) - 1;
  int lane = threadIdx.x & 31;
  uint32_t basepos = 0;
  if (lane == leader) basepos = atomicAdd(counters + key, (uint32_t)__popc(peers));
  basepos = __shfl_sync(peers, basepos, leader);
  if (pos_out) *pos_out = basepos + __popc(peers & ((1u << lane) - 1));
}

// The same without the warp aggregation: for the windows above the lowest one of the wide geometry.  There the digits
// of a warp are almost always distinct (dense scalars: random 16-bit digits; the 0/1-heavy assignment has none), so
// the match / ballot / shuffle sequence of warp_agg_inc finds nothing to merge and only delays the atomic.
__device__ __forceinline__ void plain_inc(uint32_t* counters, uint32_t key, bool active, uint32_t* pos_out) {
  if (!active) return;
  const uint32_t p = atomicAdd(counters + key, 1u);
  if (pos_out) *pos_out = p;
}
#ifndef FRCS_SORT_AGG_ALL
#define FRCS_SORT_AGG_ALL 0
#endif

__global__ void __launch_bounds__(256) zero_hist_kernel(uint32_t* __restrict__ hist, uint32_t nb_buckets, BatchStrides bs) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb_buckets) hist[blockIdx.y * bs.sort + b] = 0;
}

template <class G>
__global__ void __launch_bounds__(256)
    digits_kernel(ScalarSegs sg, uint64_t n_total, int mont, uint32_t* __restrict__ digits, uint32_t* __restrict__ hist,
                  BatchStrides bs) {
  constexpr uint32_t WINDOWS = G::WINDOWS, CB = G::CB, FULL = 1u << CB, HALF = FULL >> 1;
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint32_t p = blockIdx.y;
  if (digits) digits += p * bs.sort;
  hist += p * bs.sort;
  bool valid = i < n_total;
  Fr k = Fr::zero();
  if (valid) {
    const uint32_t* src;
    if (i < sg.end[0])
      src = sg.ptr[0] + p * sg.stride[0] + 8 * i;
    else if (i < sg.end[1])
      src = sg.ptr[1] + p * sg.stride[1] + 8 * (i - sg.end[0]);
    else
      src = sg.ptr[2] + p * sg.stride[2] + 8 * (i - sg.end[1]);
    uint4 lo = *reinterpret_cast<const uint4*>(src), hi = *reinterpret_cast<const uint4*>(src + 4);
    k.v[0] = lo.x; k.v[1] = lo.y; k.v[2] = lo.z; k.v[3] = lo.w;
    k.v[4] = hi.x; k.v[5] = hi.y; k.v[6] = hi.z; k.v[7] = hi.w;
    if (sg.skip && ((sg.skip[i >> 5] >> (i & 31)) & 1u)) k = Fr::zero();
    if (mont) k = k.from_mont();
  }
  uint32_t carry = 0;
#pragma unroll
  for (uint32_t w = 0; w < WINDOWS; w++) {
    uint32_t raw = ((k.v[(w * CB) >> 5] >> ((w * CB) & 31u)) & (FULL - 1u)) + carry;
    uint32_t neg = raw > HALF;
    uint32_t mag = neg ? FULL - raw : raw;
    carry = neg;
    bool nz = valid && mag != 0;
    if (valid && digits) digits[w * n_total + i] = nz ? ((mag << 1) | neg) : 0u;
    if (CB == 16 && w > 0 && !FRCS_SORT_AGG_ALL)
      plain_inc(hist, mag - 1, nz, nullptr);
    else
      warp_agg_inc(hist, mag - 1, nz, nullptr);
  }
}

// ---- block-histogram variant of the wide sort (batches) -----------------------------------------------------------
// With 2^15 buckets per problem and dense 16-bit digits every entry is its own global atomic in digits_kernel
// (69 M per group of 16 Falcon-1024 proofs, 0.68 ms).  Here a block owns a contiguous slice of one problem's scalars
// and keeps the problem's whole histogram in shared memory (128 KB), then adds one count per non-empty (block, bucket)
// to the global histogram: 0.32 ms.  The digits are derived from the scalars again by the scatter pass
// (scatter_scalar_kernel: 32 B read per scalar instead of 64 B of digits written and read).
// Measured and dropped: the same idea for the scatter (rebuild the block's histogram, reserve the block's range of
// every bucket with one global atomic, place the entries with shared-memory atomics) took 2.35 ms against 1.04 ms for
// one global atomic per entry -- 8 warps per SM wait on every shared-memory atomic before the dependent store; the
// per-entry scatter itself runs at ~130 G L2 operations/s (atomic + 4-byte store per entry) whatever its shape.
// Blocks of 256 threads: they must fit next to the resident accumulation blocks of the previous group (a 1024-thread
// block waits for a whole SM's register file, see plan_kernel).
constexpr int BS_THREADS = 256;
__device__ __forceinline__ Fr load_scalar(const ScalarSegs& sg, uint32_t p, uint64_t i) {
  const uint32_t* src;
  if (i < sg.end[0])
    src = sg.ptr[0] + p * sg.stride[0] + 8 * i;
  else if (i < sg.end[1])
    src = sg.ptr[1] + p * sg.stride[1] + 8 * (i - sg.end[0]);
  else
    src = sg.ptr[2] + p * sg.stride[2] + 8 * (i - sg.end[1]);
  const uint4 lo = *reinterpret_cast<const uint4*>(src), hi = *reinterpret_cast<const uint4*>(src + 4);
  Fr k;
  k.v[0] = lo.x; k.v[1] = lo.y; k.v[2] = lo.z; k.v[3] = lo.w;
  k.v[4] = hi.x; k.v[5] = hi.y; k.v[6] = hi.z; k.v[7] = hi.w;
  if (sg.skip && ((sg.skip[i >> 5] >> (i & 31)) & 1u)) k = Fr::zero();
  return k;
}
// signed digits of a scalar: d[w] = (|digit| << 1) | negative, 0 for a zero digit (digits_kernel's encoding)
template <class G>
__device__ __forceinline__ void scalar_digits(Fr k, int mont, uint32_t* d) {
  constexpr uint32_t WINDOWS = G::WINDOWS, CB = G::CB, FULL = 1u << CB, HALF = FULL >> 1;
  if (mont) k = k.from_mont();
  uint32_t carry = 0;
#pragma unroll
  for (uint32_t w = 0; w < WINDOWS; w++) {
    const uint32_t raw = ((k.v[(w * CB) >> 5] >> ((w * CB) & 31u)) & (FULL - 1u)) + carry;
    const uint32_t neg = raw > HALF;
    const uint32_t mag = neg ? FULL - raw : raw;
    carry = neg;
    d[w] = mag ? ((mag << 1) | neg) : 0u;
  }
}
constexpr int BS_UNROLL = 4;  // scalars a thread has in flight (8 warps per SM otherwise wait for one 32-byte load each)
// hist[b] += entries of the block's slice in bucket b
template <class G>
__global__ void __launch_bounds__(BS_THREADS)
    block_hist_kernel(ScalarSegs sg, uint64_t n_total, int mont, uint32_t per_block, uint32_t* __restrict__ hist,
                      BatchStrides bs) {
  extern __shared__ uint32_t s_hist[];  // G::NB counters
  constexpr uint32_t NB = G::NB, WINDOWS = G::WINDOWS;
  const uint32_t p = blockIdx.y, tid = threadIdx.x;
  uint32_t* g_cnt = hist + p * bs.sort;
  const uint64_t i0 = (uint64_t)blockIdx.x * per_block, i1 = min(i0 + per_block, n_total);
  for (uint32_t b = tid; b < NB; b += BS_THREADS) s_hist[b] = 0;
  __syncthreads();
  for (uint64_t i = i0 + tid; i < i1; i += BS_UNROLL * BS_THREADS) {
    Fr k[BS_UNROLL];
#pragma unroll
    for (int j = 0; j < BS_UNROLL; j++)
      if (i + j * BS_THREADS < i1) k[j] = load_scalar(sg, p, i + j * BS_THREADS);
#pragma unroll
    for (int j = 0; j < BS_UNROLL; j++) {
      if (i + j * BS_THREADS >= i1) break;
      uint32_t d[WINDOWS];
      scalar_digits<G>(k[j], mont, d);
#pragma unroll
      for (uint32_t w = 0; w < WINDOWS; w++)
        if (d[w]) atomicAdd(&s_hist[(d[w] >> 1) - 1], 1u);
    }
  }
  __syncthreads();
  for (uint32_t b = tid; b < NB; b += BS_THREADS) {
    const uint32_t c = s_hist[b];
    if (c) atomicAdd(g_cnt + b, c);
  }
}

// Scatter straight from the scalars, one thread per scalar, no digits array.
// AGG_ALL = false (wide geometry): the WINDOWS positions are requested back to back (independent atomics in flight)
// before any of the dependent stores; the lowest window keeps the warp aggregation (the Boolean witnesses in the
// l_query part all land in the digit-1 bucket).
// AGG_ALL = true (narrow geometry, 128 buckets, the assignment z): every window is warp-aggregated; almost all scalars
// are bits or 14-bit values, so the windows above the second are skipped by a whole warp at a time.
template <class G, bool AGG_ALL>
__global__ void __launch_bounds__(256)
    scatter_scalar_kernel(ScalarSegs sg, uint64_t n_total, int mont, uint32_t* __restrict__ cursor,
                          uint32_t* __restrict__ sorted, BatchStrides bs) {
  constexpr uint32_t WINDOWS = G::WINDOWS;
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint32_t p = blockIdx.y;
  cursor += p * bs.sort;
  sorted += p * bs.sort;
  const bool valid = i < n_total;
  uint32_t d[WINDOWS];
  if (valid) {
    scalar_digits<G>(load_scalar(sg, p, i), mont, d);
  } else {
#pragma unroll
    for (uint32_t w = 0; w < WINDOWS; w++) d[w] = 0;
  }
  if constexpr (AGG_ALL) {
#pragma unroll
    for (uint32_t w = 0; w < WINDOWS; w++) {
      if (!__any_sync(0xffffffffu, d[w] != 0)) continue;
      uint32_t pos = 0;
      warp_agg_inc(cursor, (d[w] >> 1) - 1, d[w] != 0, &pos);
      if (d[w]) {
        FRCS_ASSERT(pos < n_total * WINDOWS);
        sorted[pos] = (uint32_t)(w * n_total + i) | ((d[w] & 1u) << 31);
      }
    }
  } else {
    uint32_t pos[WINDOWS];
    pos[0] = 0;
    warp_agg_inc(cursor, (d[0] >> 1) - 1, d[0] != 0, &pos[0]);
#pragma unroll
    for (uint32_t w = 1; w < WINDOWS; w++) pos[w] = d[w] ? atomicAdd(cursor + (d[w] >> 1) - 1, 1u) : 0u;
#pragma unroll
    for (uint32_t w = 0; w < WINDOWS; w++)
      if (d[w]) {
        FRCS_ASSERT(pos[w] < n_total * WINDOWS);
        sorted[pos[w]] = (uint32_t)(w * n_total + i) | ((d[w] & 1u) << 31);
      }
  }
}

// Level plan of one problem.  Level 0 cuts the SORTED LIST (not the buckets) into pieces of lc[0] consecutive entries, one
// thread each, so every lane of a warp does the same number of mixed additions; a piece that runs over a bucket
// boundary yields one sum per bucket it touches.  Bucket b (entries [o, o + c) of the list) therefore has
//   cnt[1][b] = (o + c - 1) / lc[0] - o / lc[0] + 1   sums after level 0,
// and the higher levels cut each bucket's sums into slices: cnt[l+1][b] = ceil(cnt[l][b] / lc[l]) (so does level 0 when
// it is a pair level, MSM_LV_PAIR).
// off[l] = exclusive scan of cnt[l] (off[l][NB] = total), cursor = off[0] (scatter positions).  One block per
// (level, problem); every level is derived from cnt[0] directly, so the levels run in parallel.
// Blocks of at most 256 threads and few registers: this kernel sits between the two big kernels of the sort on a
// high-priority stream while the previous group's accumulation fills the SMs; a 1024-thread block needed a whole
// SM's register file to become free and waited ~5 ms for it.
template <class G>
__global__ void __launch_bounds__(G::NB < 256u ? G::NB : 256u)
    plan_kernel(uint32_t* cnt_all, uint32_t* off_all, uint32_t* cursor_all, MsmLevels lv, BatchStrides bs) {
  __shared__ uint32_t s_warp[32];
  constexpr uint32_t NB = G::NB, THREADS = NB < 256u ? NB : 256u, PER = NB / THREADS;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t l = blockIdx.x, p = blockIdx.y;
  uint32_t* cnt = cnt_all + p * bs.sort;
  uint32_t* o = off_all + p * bs.sort + (uint64_t)l * (NB + 1);
  // entries of a bucket at level l, from its level-0 offset and count
  auto level_count = [&](uint32_t o0, uint32_t c) {
    if (l == 0 || c == 0) return c;
    if (lv.kind[0] == MSM_LV_SEG)
      c = (o0 + c - 1) / lv.lc[0] - o0 / lv.lc[0] + 1;
    else
      c = (c + lv.lc[0] - 1) / lv.lc[0];
    for (uint32_t k = 1; k < l; k++) c = (c + lv.lc[k] - 1) / lv.lc[k];
    return c;
  };
  // block-wide exclusive scan of one value per thread
  auto block_excl = [&](uint32_t sum) {
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
      if ((int)lane >= d) incl += v;
    }
    __syncthreads();  // s_warp may still be read by the previous call
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = lane < THREADS / 32 ? s_warp[lane] : 0u, wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
        if ((int)lane >= d) wi += v;
      }
      s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    return s_warp[wid] + incl - sum;
  };
  uint32_t sum0 = 0;
  for (uint32_t j = 0; j < PER; j++) sum0 += cnt[tid * PER + j];
  const uint32_t base0 = block_excl(sum0);  // level-0 offset of this thread's first bucket
  uint32_t run = base0;
  if (l > 0) {
    uint32_t sum = 0, run0 = base0;
    for (uint32_t j = 0; j < PER; j++) {
      const uint32_t c0 = cnt[tid * PER + j], c = level_count(run0, c0);
      cnt[(uint64_t)l * NB + tid * PER + j] = c;
      sum += c;
      run0 += c0;
    }
    run = block_excl(sum);
  }
  uint32_t* cursor = cursor_all + p * bs.sort;
  uint32_t run0 = base0;
  for (uint32_t j = 0; j < PER; j++) {
    o[tid * PER + j] = run;
    if (l == 0) cursor[tid * PER + j] = run;
    const uint32_t c0 = cnt[tid * PER + j];  // level 0 counts are not modified by any block
    run += level_count(run0, c0);
    run0 += c0;
  }
  if (tid == THREADS - 1) o[NB] = run;
}

__global__ void __launch_bounds__(256)
    scatter_kernel(const uint32_t* __restrict__ digits, uint64_t n_total, uint32_t* __restrict__ cursor,
                   uint32_t* __restrict__ sorted, BatchStrides bs, bool wide) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t w = blockIdx.y;
  const uint32_t p = blockIdx.z;
  digits += p * bs.sort;
  sorted += p * bs.sort;
  cursor += p * bs.sort;
  bool valid = i < n_total;
  uint32_t d = valid ? digits[w * n_total + i] : 0;
  bool nz = d != 0;
  uint32_t pos = 0;
  if (wide && w > 0 && !FRCS_SORT_AGG_ALL)  // (uniform over the block)
    plain_inc(cursor, (d >> 1) - 1, nz, &pos);
  else
    warp_agg_inc(cursor, (d >> 1) - 1, nz, &pos);
  FRCS_ASSERT(!nz || pos < n_total * gridDim.y);
  if (nz) sorted[pos] = (uint32_t)(w * n_total + i) | ((d & 1u) << 31);
}

// thread t -> (bucket b, slice k) through the next level's offsets
template <class G>
__device__ __forceinline__ uint32_t find_bucket(const uint32_t* __restrict__ off_next, uint32_t t) {
  uint32_t lo = 0, hi = G::NB;  // off_next[lo] <= t < off_next[hi]
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (off_next[mid] <= t)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

#ifndef ACCUM0_MIN_BLOCKS
#define ACCUM0_MIN_BLOCKS 3
#endif
// level 0: mixed additions of pre-processed affine points.  Thread t takes the entries [t lc, (t + 1) lc) of the sorted
// list whatever buckets they belong to (equal work per lane: with one thread per <= lc-entry slice of a bucket the
// partial last slices left 11 % of the lanes idle, ncu: 28.3 active threads per warp) and writes one sum per bucket
// touched: piece (t - off[b] / lc) of bucket b, at off_next[b] + that (plan_kernel counts the pieces the same way).
template <class F, class G>
__global__ void __launch_bounds__(128, ACCUM0_MIN_BLOCKS)
    accum0_kernel(Tables tabs, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ off,
                  const uint32_t* __restrict__ off_next, uint32_t lc, uint32_t* __restrict__ out, uint32_t pf,
                  BatchStrides bs) {
  constexpr int AW = 2 * Words<F>::N, XW = 4 * Words<F>::N;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t q = blockIdx.y, p = q % bs.n_sort;
  const uint32_t* __restrict__ pts = tabs.pts[q / bs.n_sort];
  sorted += p * bs.sort;
  off += p * bs.sort;
  off_next += p * bs.sort;
  out += q * bs.acc;
  const uint32_t total = off[G::NB];
  uint32_t e = t * lc;
  if (e >= total) return;
  const uint32_t e_end = min(e + lc, total);
  uint32_t b = find_bucket<G>(off, e);  // the (non-empty) bucket that holds entry e
  uint32_t b_end = off[b + 1];
  uint32_t slot = off_next[b] + (t - off[b] / lc);
  ec::XYZZ<F> acc = ec::XYZZ<F>::infinity();
  for (; e < e_end; e++) {
    if (e >= b_end) {  // the list moves on to the next non-empty bucket, which starts inside this piece: its piece 0
      FRCS_ASSERT(slot < off_next[G::NB]);
      st_xyzz<F>(out + (uint64_t)slot * XW, acc);
      acc = ec::XYZZ<F>::infinity();
      do {
        b++;
        b_end = off[b + 1];
      } while (e >= b_end);
      slot = off_next[b];
    }
    if ((pf & 255u) && e + (pf >> 8) < e_end)
      prefetch_words(pts + (uint64_t)(sorted[e + (pf >> 8)] & 0x7fffffffu) * AW, AW, pf & 255u);
    const uint32_t idx = sorted[e];
    const ec::Affine<F> pt = ld_affine<F>(pts + (uint64_t)(idx & 0x7fffffffu) * AW);
    acc.add_mixed(pt, idx >> 31);
  }
  FRCS_ASSERT(slot < off_next[G::NB] && b < G::NB);
  st_xyzz<F>(out + (uint64_t)slot * XW, acc);
}

// level >= 1: sums of XYZZ slice sums
template <class F, class G>
__global__ void __launch_bounds__(128)
    accumN_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                  const uint32_t* __restrict__ off_next, uint32_t lc, uint32_t* __restrict__ out, BatchStrides bs) {
  constexpr int XW = 4 * Words<F>::N;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t q = blockIdx.y, p = q % bs.n_sort;
  in += q * bs.acc;
  off += p * bs.sort;
  cnt += p * bs.sort;
  off_next += p * bs.sort;
  out += q * bs.acc;
  if (t >= off_next[G::NB]) return;
  uint32_t b = find_bucket<G>(off_next, t);
  uint32_t k = t - off_next[b];
  uint32_t e0 = off[b] + k * lc, e1 = min(off[b] + cnt[b], e0 + lc);
  ec::XYZZ<F> acc = ld_xyzz<F>(in + (uint64_t)e0 * XW);
  for (uint32_t e = e0 + 1; e < e1; e++) acc.add(ld_xyzz<F>(in + (uint64_t)e * XW));
  st_xyzz<F>(out + (uint64_t)t * XW, acc);
}

// level >= 1: mixed additions of the affine sums a pair level wrote
template <class F, class G>
__global__ void __launch_bounds__(128, ACCUM0_MIN_BLOCKS)
    accumA_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                  const uint32_t* __restrict__ off_next, uint32_t lc, uint32_t* __restrict__ out, BatchStrides bs) {
  constexpr int AW = 2 * Words<F>::N, XW = 4 * Words<F>::N;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t q = blockIdx.y, p = q % bs.n_sort;
  in += q * bs.acc;
  off += p * bs.sort;
  cnt += p * bs.sort;
  off_next += p * bs.sort;
  out += q * bs.acc;
  if (t >= off_next[G::NB]) return;
  uint32_t b = find_bucket<G>(off_next, t);
  uint32_t k = t - off_next[b];
  uint32_t e0 = off[b] + k * lc, e1 = min(off[b] + cnt[b], e0 + lc);
  ec::XYZZ<F> acc = ec::XYZZ<F>::infinity();
  for (uint32_t e = e0; e < e1; e++) acc.add_mixed(ld_affine<F>(in + (uint64_t)e * AW), false);
  st_xyzz<F>(out + (uint64_t)t * XW, acc);
}

// Pair level ("batched affine"): sum s of bucket b is entry 2k + entry 2k+1 of the bucket's list (k = s - off_next[b];
// a last odd entry passes through), both affine, the result affine.  An affine addition needs 1 / (x2 - x1): thread t
// owns the m consecutive sums [t m, (t + 1) m) and shares one inversion between them (Montgomery's trick): a forward
// pass over the x coordinates keeps the running product of the denominators and parks each prefix in `scratch`
// ([pair][thread], coalesced), one inversion (inverse_w4), then a backward pass peels the individual inverses off and
// finishes the sums.  Per sum 1 + 2 + 3 multiplications and squarings + (494 / m for the inversion), against 10 for a
// mixed addition in XYZZ coordinates.  Doublings, cancellations and points at infinity are classified the same way in
// both passes (ec::pair_classify) and take no part in the product.
// TABLE: the entries are (sign, index) references into the pre-processed base table, else affine sums in `in`.
template <class F, class G, bool TABLE>
__global__ void __launch_bounds__(128, ACCUM0_MIN_BLOCKS)
    pair_kernel(Tables tabs, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ in,
                const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off_next,
                uint32_t m, uint32_t* __restrict__ out, uint32_t* __restrict__ scratch, uint64_t scratch_stride,
                uint32_t pf, BatchStrides bs) {
  constexpr int W = Words<F>::N, AW = 2 * W;
  const uint32_t T = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t q = blockIdx.y, p = q % bs.n_sort;
  const uint32_t* __restrict__ pts = tabs.pts[q / bs.n_sort];
  sorted += p * bs.sort;
  in += q * bs.acc;
  off += p * bs.sort;
  cnt += p * bs.sort;
  off_next += p * bs.sort;
  out += q * bs.acc;
  scratch += q * scratch_stride + (uint64_t)t * W;
  const uint32_t n_out = off_next[G::NB], n_in = off[G::NB];
  const uint32_t s0 = t * m;
  if (s0 >= n_out) return;
  const uint32_t s1 = min(s0 + m, n_out);
  const uint32_t pf_mode = pf & 255u, pf_dist = 2 * (pf >> 8);  // the passes walk the entry list two entries a sum
  // entry e is needed soon: its x coordinate (forward pass) or the whole point (backward pass)
  auto prefetch_entry = [&](uint32_t e, int nw) {
    const uint32_t* a = TABLE ? pts + (uint64_t)(sorted[e] & 0x7fffffffu) * AW : in + (uint64_t)e * AW;
    prefetch_words(a, nw, pf_mode);
  };
  // entry e of the level's input: where its point lives, and whether it is to be negated
  auto src = [&](uint32_t e, bool& neg) -> const uint32_t* {
    if (TABLE) {
      const uint32_t idx = sorted[e];
      neg = idx >> 31;
      return pts + (uint64_t)(idx & 0x7fffffffu) * AW;
    }
    neg = false;
    return in + (uint64_t)e * AW;
  };
  auto load_pt = [&](uint32_t e) {
    bool neg;
    const uint32_t* a = src(e, neg);
    ec::Affine<F> pt = ld_affine<F>(a);
    if (neg) pt.y = pt.y.neg();
    return pt;
  };
  uint32_t b = find_bucket<G>(off_next, s0);  // the (non-empty) bucket that owns sum s0
  uint32_t ob = off_next[b], oe = off_next[b + 1], ib = off[b], ie = ib + cnt[b];
  F run = F::one();
  for (uint32_t s = s0; s < s1; s++) {
    while (s >= oe) {
      b++;
      ob = oe;
      oe = off_next[b + 1];
      ib = off[b];
      ie = ib + cnt[b];
    }
    const uint32_t e0 = ib + 2 * (s - ob);
    const bool has2 = e0 + 1 < ie;
    FRCS_ASSERT(e0 < ie && b < G::NB);
    if (pf_mode && e0 + pf_dist + 1 < n_in) {
      prefetch_entry(e0 + pf_dist, W);
      prefetch_entry(e0 + pf_dist + 1, W);
    }
    bool n1, n2 = false;
    const uint32_t* a1 = src(e0, n1);
    const F x1 = ld_field<F>(a1);
    F den = F::one();
    bool plain = false;
    if (has2) {
      const uint32_t* a2 = src(e0 + 1, n2);
      const F x2 = ld_field<F>(a2);
      den = x2 - x1;
      plain = !x1.is_zero() && !x2.is_zero() && !den.is_zero();  // two finite points with different x
      if (!plain) {
        ec::Affine<F> p1 = {x1, ld_field<F>(a1 + W)}, p2 = {x2, ld_field<F>(a2 + W)};
        if (n1) p1.y = p1.y.neg();
        if (n2) p2.y = p2.y.neg();
        den = F::one();
        plain = ec::pair_classify<F>(p1, p2, true, den) <= ec::PAIR_DBL;
      }
    }
    st_field<F>(scratch + (uint64_t)(s - s0) * T * W, run);
    if (plain) run = run * den;
  }
  F inv = F::inverse_w4(run);
  for (uint32_t s = s1; s-- > s0;) {
    while (s < ob) {
      b--;
      oe = ob;
      ob = off_next[b];
      ib = off[b];
      ie = ib + cnt[b];
    }
    const uint32_t e0 = ib + 2 * (s - ob);
    const bool has2 = e0 + 1 < ie;
    if (pf_mode && e0 >= pf_dist) {
      prefetch_entry(e0 - pf_dist, AW);
      prefetch_entry(e0 - pf_dist + 1, AW);
      if (s - s0 >= (pf >> 8)) prefetch_words(scratch + (uint64_t)(s - s0 - (pf >> 8)) * T * W, W, pf_mode);
    }
    const ec::Affine<F> p1 = load_pt(e0);
    const ec::Affine<F> p2 = has2 ? load_pt(e0 + 1) : p1;
    F den = F::one(), dinv = F::one();
    const int kind = ec::pair_classify<F>(p1, p2, has2, den);
    if (kind <= ec::PAIR_DBL) {
      dinv = inv * ld_field<F>(scratch + (uint64_t)(s - s0) * T * W);
      inv = inv * den;
    }
    st_affine<F>(out + (uint64_t)s * AW, ec::pair_finish<F>(kind, p1, p2, dinv));
  }
}

// shared-memory tree over the block's accumulators; the sum ends up in thread 0's `acc`
template <class F>
__device__ __forceinline__ void block_tree_sum(ec::XYZZ<F>& acc, uint32_t* sm) {
  constexpr int XW = 4 * Words<F>::N;
  const uint32_t tid = threadIdx.x;
  for (uint32_t d = blockDim.x >> 1; d > 0; d >>= 1) {
    if (tid >= d && tid < 2 * d) {
      uint32_t* mine = sm + (uint64_t)(tid - d) * XW;
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&acc);
      for (int i = 0; i < XW; i++) mine[i] = w[i];
    }
    __syncthreads();
    if (tid < d) {
      ec::XYZZ<F> o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
      const uint32_t* src = sm + (uint64_t)tid * XW;
      for (int i = 0; i < XW; i++) ow[i] = src[i];
      acc.add(o);
    }
    __syncthreads();
  }
}

// Buckets that still hold more than one slice sum after the last level (only buckets with more than prod(lc) entries,
// e.g. the digit-1 bucket of the 0/1-heavy witness scalars): the block sums them, thread-strided then a shared-memory
// tree, and writes the result over the bucket's first entry.
// Grid (NB / FIN, nq): a block scans FIN consecutive buckets and works only on those with > 1 entries.  FIN = 64 for the
// 32768 buckets of the wide geometry (few of them qualify); FIN = 1 for the 128 buckets of the narrow one, where every
// bucket may qualify (the schoolbook circuit's product witnesses through a key shard: ~5 sums left in each of the 128
// buckets, which one block per 64 buckets took 6 ms to finish one after the other).
template <class F, class G>
__global__ void __launch_bounds__(64)
    finish_kernel(uint32_t* __restrict__ entries, const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                  BatchStrides bs) {
  constexpr int XW = 4 * Words<F>::N;
  constexpr uint32_t FIN = G::NB >= 4096u ? 64u : 1u;
  extern __shared__ uint32_t sm[];
  __shared__ uint32_t s_cnt[FIN];
  const uint32_t q = blockIdx.y, p = q % bs.n_sort, b0 = blockIdx.x * FIN;
  if (threadIdx.x < FIN) s_cnt[threadIdx.x] = cnt[p * bs.sort + b0 + threadIdx.x];
  __syncthreads();
  entries += q * bs.acc;
  for (uint32_t i = 0; i < FIN; i++) {
    const uint32_t c = s_cnt[i];
    if (c <= 1) continue;
    const uint32_t e0 = off[p * bs.sort + b0 + i];
    ec::XYZZ<F> acc = ec::XYZZ<F>::infinity();
    for (uint32_t e = threadIdx.x; e < c; e += blockDim.x) acc.add(ld_xyzz<F>(entries + (uint64_t)(e0 + e) * XW));
    block_tree_sum<F>(acc, sm);
    if (threadIdx.x == 0) st_xyzz<F>(entries + (uint64_t)e0 * XW, acc);
    __syncthreads();
  }
}

// ---- bucket reduction:  sum_b (b+1) B_b, 32768 buckets (wide geometry) -------------------------
// Stage 1 (one thread per run of K = 8 buckets, b = 8 s + i):
//     T_s = sum_i (i+1) B_{8s+i}   (running sums),      R_s = sum_i B_{8s+i}
//   so that  sum_b (b+1) B_b = sum_s T_s + 8 sum_s s R_s.
// Stage 2: the weights s < 4096 are split into bits,  sum_s s R_s = sum_j 2^j C_j  with
//   C_j = sum_{s : bit j of s} R_s; with sum_s T_s (split in two halves) that is 14 plain sums of
//   2048 points each ("channels"): 4 blocks x 64 threads x 8 points, then a shared-memory tree.
// Stage 3: result = sum_s T_s + sum_j 2^(j+3) C_j: one block takes the 14 x 4 partial sums,
//   doubles them in parallel and adds them with a tree.
// Every running sum starts at a fixed point Q (the group generator) instead of infinity, so that
// `acc += run` never meets acc == run or a point at infinity and all lanes of a warp stay on the
// generic-addition path: T'_s = T_s + 8 Q, R'_s = R_s + Q, and the total picks up
// (NB + 8 * sum_{s < RED_RUNS} s) Q = RED_CORR * Q, which reduce_combine_kernel takes off again
// (reduce_corr_kernel computes -RED_CORR * Q once per context).  The addition formulas still handle
// every exceptional case, so this is about speed only.
template <class F> struct Gen;
template <> struct Gen<Fq> {
  __device__ static ec::Affine<Fq> get() {
    ec::Affine<Fq> g;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      g.x.v[i] = FqParams::G1X(i);
      g.y.v[i] = FqParams::G1Y(i);
    }
    return g;
  }
};
template <> struct Gen<Fq2> {
  __device__ static ec::Affine<Fq2> get() {
    ec::Affine<Fq2> g;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      g.x.c0.v[i] = FqParams::G2X0(i);
      g.x.c1.v[i] = FqParams::G2X1(i);
      g.y.c0.v[i] = FqParams::G2Y0(i);
      g.y.c1.v[i] = FqParams::G2Y1(i);
    }
    return g;
  }
};
constexpr uint32_t RED_K = 8, RED_RUNS = Wide::NB / RED_K, RED_BITS = 12, RED_CH = RED_BITS + 2, RED_BLK = 4, RED_PER = 8;
static_assert((1u << RED_BITS) == RED_RUNS, "weight bits");
static_assert(RED_BLK * 64 * RED_PER == RED_RUNS / 2, "channel geometry");
static_assert(RED_CH * RED_BLK <= 64, "combine block");
constexpr uint32_t RED_CORR = Wide::NB + RED_K * (RED_RUNS * (RED_RUNS - 1) / 2);

template <class F>
__global__ void reduce_corr_kernel(uint32_t* out) {
  const uint32_t k[1] = {RED_CORR};
  ec::XYZZ<F> q = ec::XYZZ<F>::from_affine(Gen<F>::get());
  st_xyzz<F>(out, q.mul(k, 32).neg());
}

template <class F>
__global__ void __launch_bounds__(64)
    bucket_reduce_kernel(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ off,
                         const uint32_t* __restrict__ cnt, uint32_t* __restrict__ partial, BatchStrides bs) {
  constexpr int XW = 4 * Words<F>::N;
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t q = blockIdx.y, p = q % bs.n_sort;
  entries += q * bs.acc;
  off += p * bs.sort;
  cnt += p * bs.sort;
  partial += q * bs.acc;
  if (s >= RED_RUNS) return;
  ec::XYZZ<F> run = ec::XYZZ<F>::from_affine(Gen<F>::get()), acc = ec::XYZZ<F>::infinity();
  for (int i = (int)RED_K - 1; i >= 0; i--) {
    uint32_t b = s * RED_K + i;
    if (cnt[b]) run.add(ld_xyzz<F>(entries + (uint64_t)off[b] * XW));
    acc.add(run);
  }
  st_xyzz<F>(partial + (uint64_t)s * XW, acc);               // T_s
  st_xyzz<F>(partial + (uint64_t)(RED_RUNS + s) * XW, run);  // R_s
}

// stage 2: grid (RED_BLK, RED_CH, nq).  channels 0,1: T_s for s in the lower / upper half;
// channel j+2: R_s over the s with bit j set.  out: [channel][RED_BLK] partial sums.
template <class F>
__global__ void __launch_bounds__(64) reduce_channels_kernel(uint32_t* __restrict__ partial, BatchStrides bs) {
  constexpr int XW = 4 * Words<F>::N;
  extern __shared__ uint32_t sm[];
  partial += blockIdx.z * bs.acc;
  const uint32_t ch = blockIdx.y;
  ec::XYZZ<F> acc = ec::XYZZ<F>::infinity();
  for (uint32_t k = 0; k < RED_PER; k++) {
    uint32_t u = (k * RED_BLK + blockIdx.x) * 64 + threadIdx.x;  // < RED_RUNS / 2
    uint32_t s;
    const uint32_t* src;
    if (ch < 2) {
      s = ch * (RED_RUNS / 2) + u;
      src = partial;
    } else {
      uint32_t bit = ch - 2;
      s = ((u >> bit) << (bit + 1)) | (1u << bit) | (u & ((1u << bit) - 1));
      src = partial + (uint64_t)RED_RUNS * XW;
    }
    acc.add(ld_xyzz<F>(src + (uint64_t)s * XW));
  }
  block_tree_sum<F>(acc, sm);
  if (threadIdx.x == 0) st_xyzz<F>(partial + (uint64_t)(2 * RED_RUNS + ch * RED_BLK + blockIdx.x) * XW, acc);
}
// stage 3: grid (nq), 64 threads
template <class F>
__global__ void __launch_bounds__(64)
    reduce_combine_kernel(const uint32_t* __restrict__ partial, const uint32_t* __restrict__ corr, uint32_t* out0,
                          uint32_t* out1, uint64_t out_stride, BatchStrides bs) {
  constexpr int XW = 4 * Words<F>::N;
  extern __shared__ uint32_t sm[];
  const uint32_t q = blockIdx.x;
  partial += q * bs.acc;
  uint32_t* out = (q / bs.n_sort ? out1 : out0) + (q % bs.n_sort) * out_stride;
  const uint32_t tid = threadIdx.x;
  ec::XYZZ<F> acc = ec::XYZZ<F>::infinity();
  if (tid < RED_CH * RED_BLK) {
    acc = ld_xyzz<F>(partial + (uint64_t)(2 * RED_RUNS + tid) * XW);
    const uint32_t ch = tid / RED_BLK;
    if (ch >= 2)
      for (uint32_t d = 0; d < ch + 1; d++) acc = acc.dbl();  // 2^(j+3), j = ch - 2
  }
  if (tid == 63) acc = ld_xyzz<F>(corr);  // -RED_CORR * Q
  block_tree_sum<F>(acc, sm);
  if (tid == 0) st_xyzz<F>(out, acc);
}

// ---- bucket reduction, 128 buckets (narrow geometry): one block per problem, thread b owns bucket b.
//   sum_b (b+1) B_b = sum_k S_k  with the suffix sums S_k = sum_{j >= k} B_j:
// a Hillis-Steele suffix scan over shared memory (log2 NB steps), then a tree over the S_k.
template <class F, class G>
__global__ void __launch_bounds__(G::NB)
    narrow_reduce_kernel(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ off,
                         const uint32_t* __restrict__ cnt, uint32_t* out0, uint32_t* out1, uint64_t out_stride,
                         BatchStrides bs) {
  constexpr int XW = 4 * Words<F>::N;
  constexpr uint32_t NB = G::NB;
  extern __shared__ uint32_t sm[];  // NB x XYZZ
  const uint32_t q = blockIdx.x, p = q % bs.n_sort, b = threadIdx.x;
  entries += q * bs.acc;
  off += p * bs.sort;
  cnt += p * bs.sort;
  uint32_t* out = (q / bs.n_sort ? out1 : out0) + (q % bs.n_sort) * out_stride;
  FRCS_ASSERT(b < NB && (!cnt[b] || off[b] < off[NB]));
  ec::XYZZ<F> acc = cnt[b] ? ld_xyzz<F>(entries + (uint64_t)off[b] * XW) : ec::XYZZ<F>::infinity();
  for (uint32_t d = 1; d < NB; d <<= 1) {  // after the step: acc_b = sum of B_j over j in [b, b + 2d)
    uint32_t* mine = sm + (uint64_t)b * XW;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&acc);
    for (int i = 0; i < XW; i++) mine[i] = w[i];
    __syncthreads();
    if (b + d < NB) {
      ec::XYZZ<F> o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
      const uint32_t* src = sm + (uint64_t)(b + d) * XW;
      for (int i = 0; i < XW; i++) ow[i] = src[i];
      acc.add(o);
    }
    __syncthreads();
  }
  block_tree_sum<F>(acc, sm);
  if (b == 0) st_xyzz<F>(out, acc);
}

template <class F>
__global__ void to_affine_kernel(const uint32_t* in, uint32_t* out) {
  ec::XYZZ<F> p = ld_xyzz<F>(in);
  st_affine<F>(out, p.to_affine());
}

}  // namespace

// ---------------------------------------------------------------------------------------
#ifdef MSM_DEFINE_LEVELS
// FRCS_MSM_PAIR: 0 (default) = never use the pair levels, 1 = for large batches, 2 = always (wide geometry; tests).
// Off by default: measured on B200 (profiles/r02_ncu_pair.txt, Falcon-1024, groups of 16 proofs) the pair levels do the
// l+h accumulation in 32.7 ms (15.4 + 8.1 ms for two pair levels, 9.1 ms for the mixed additions that follow) against
// 26 ms for the fixed-length XYZZ pieces: 21 % fewer multiplications, but the forward pass has one multiplication per two
// dependent gathers at ~2.7 warps per scheduler (IPC 0.98 against 1.20), 256-pair blocks leave 8-23 % of the SM time in
// the tail of a level, and the inversion still costs 1.9 multiplications per pair.
static int msm_pair_env() {
  static const int v = [] {
    const char* e = getenv("FRCS_MSM_PAIR");
    return e ? atoi(e) : 0;
  }();
  return v;
}

MsmLevels msm_levels(uint64_t n_total, int cb, uint32_t nb_problems) {
  // level 0: pieces of lc[0] consecutive sorted entries (mixed additions; one sum per bucket touched, see plan_kernel);
  // level l > 0: slices of <= lc[l] sums of a bucket;
  // whatever is left per bucket (more than one sum only for buckets with more than prod(lc) entries) is
  // finished by one block per bucket (finish_kernel).
  MsmLevels lv;
  memset(&lv, 0, sizeof lv);
  const uint32_t nb = msm_buckets(cb);
  uint64_t m = n_total * msm_windows(cb);  // bound on the number of non-zero digits
  if (cb == MSM_CB_NARROW) {
    // 128 buckets: the digit-1 bucket of window 0 alone holds ~54 % of z (the Boolean witnesses that are 1), an
    // average bucket a few hundred entries: three levels (16 x 8 x 8) leave one sum in all but the heaviest buckets
    lv.n_levels = 3;
    lv.kind[0] = MSM_LV_SEG;
    lv.kind[1] = lv.kind[2] = MSM_LV_XYZZ;
    lv.lc[0] = 16;
    lv.lc[1] = 8;
    lv.lc[2] = 8;
  } else {
    // pair levels: worth it when every thread gets a few hundred pairs (the inversion costs ~494 multiplications) and
    // the pairs of the batch still fill the GPU: >= 2^25 entries in the batch; buffers are sized for <= 2^23 per problem
    const int pe = msm_pair_env();
    const bool pair = pe == 2 || (pe != 0 && m > (1u << 20) && m <= (1u << 23) && m * nb_problems >= (1ull << 25));
    if (pair) {
      const uint32_t np = msm_env_u32("FRCS_MSM_PAIR_LEVELS", 2);
      lv.pair_m = msm_env_u32("FRCS_MSM_PAIR_M", 256);
      lv.n_levels = (np > 5 ? 5 : np) + 2;
      for (uint32_t l = 0; l + 2 < lv.n_levels; l++) {
        lv.kind[l] = MSM_LV_PAIR;
        lv.lc[l] = 2;
      }
      lv.kind[lv.n_levels - 2] = MSM_LV_MIXED;
      lv.lc[lv.n_levels - 2] = 32;
      lv.kind[lv.n_levels - 1] = MSM_LV_XYZZ;
      lv.lc[lv.n_levels - 1] = 8;
    } else {
      lv.n_levels = 2;
      lv.kind[0] = MSM_LV_SEG;
      lv.kind[1] = MSM_LV_XYZZ;
#ifdef FRCS_LC0_WIDE
      lv.lc[0] = m > (1u << 20) ? FRCS_LC0_WIDE : 8;
#else
      // measured 16 / 32 / 64: 268 / 365 / 371 proofs/s (Falcon-1024, groups of 16); with the fused Y3 64 / 96 / 128:
      // 384 / 389 / 388 (fewer sums for accumN_kernel against a longer tail); FRCS_LC0 overrides
      lv.lc[0] = m > (1u << 20) ? msm_env_u32("FRCS_LC0", 96) : 8;
      if (lv.lc[0] < 8 || lv.lc[0] > 1024) lv.lc[0] = 96;
#endif
      // FRCS_MSM_ONE_WAVE=1: a launch of one to three waves of 64-entry pieces pays for whole waves (a key shard's l+h
      // MSM: 70 k pieces on 56,832 resident threads = two waves): cut the list into one wave of longer pieces instead.
      // Off by default: with other streams' kernels holding part of the SMs the "one wave" spills into a second one of
      // the longer pieces (single Falcon-1024 proof: 5.8 ms instead of 5.5 ms).
      const uint64_t resident = 148ull * 3 * 128, pieces = m * nb_problems / 64;
      if (m > (1u << 20) && pieces > resident && pieces < 3 * resident && msm_env_u32("FRCS_MSM_ONE_WAVE", 0))
        lv.lc[0] = (uint32_t)((m * nb_problems + resident - 1) / resident);
      lv.lc[1] = 8;
    }
  }
  uint64_t prev = m;
  for (uint32_t l = 0; l < lv.n_levels; l++) {
    lv.t_max[l] = prev / lv.lc[l] + nb + 1;
    prev = lv.t_max[l];
  }
  return lv;
}

#endif

// Work layouts (bytes, 256-aligned); a batch is consecutive copies.
struct SortLayout {
  size_t digits, sorted, cnt, off, cursor, total;
};
struct AccLayout {
  size_t buf0, buf1, partial, scratch, total;
};
// (the same for every level structure msm_levels may choose: room for MSM_MAX_LEVELS levels)
static SortLayout sort_layout(uint64_t n_total, int cb) {
  const size_t NB = msm_buckets(cb), WINDOWS = msm_windows(cb);
  SortLayout w;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += (bytes + 255) & ~(size_t)255;
    return at;
  };
  w.digits = take(n_total * WINDOWS * 4);
  w.sorted = take(n_total * WINDOWS * 4);
  w.cnt = take((size_t)(MSM_MAX_LEVELS + 1) * NB * 4);
  w.off = take((size_t)(MSM_MAX_LEVELS + 1) * (NB + 1) * 4);
  w.cursor = take(NB * 4);
  w.total = o;
  return w;
}
// pair slots of a pair level: whole blocks of 128 threads x pair_m pairs
static uint64_t pair_slots(const MsmLevels& lv, uint32_t l) {
  const uint64_t per_block = 128ull * lv.pair_m;
  return (lv.t_max[l] + per_block - 1) / per_block * per_block;
}
static AccLayout acc_layout(uint64_t n_total, size_t xyzz_bytes, int cb) {
  AccLayout w;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += (bytes + 255) & ~(size_t)255;
    return at;
  };
  // level l writes buf[l & 1]: buf0 holds levels 0, 2, ..., buf1 levels 1, 3, ...; sized for both level structures
  // (small and large batches), so that a context's buffers do not depend on the batch
  size_t bb[2] = {0, 0}, sc = 0;
  for (uint32_t nbp : {1u, 1u << 20}) {
    MsmLevels lv = msm_levels(n_total, cb, nbp);
    for (uint32_t l = 0; l < lv.n_levels; l++) {
      const bool pair = lv.kind[l] == MSM_LV_PAIR;
      const size_t bytes = (lv.t_max[l] + 1) * (pair ? xyzz_bytes / 2 : xyzz_bytes);
      bb[l & 1] = bytes > bb[l & 1] ? bytes : bb[l & 1];
      if (pair) {
        const size_t s = pair_slots(lv, l) * (xyzz_bytes / 4);
        sc = s > sc ? s : sc;
      }
    }
  }
  w.buf0 = take(bb[0]);
  w.buf1 = take(bb[1]);
  w.partial = take((2 * RED_RUNS + RED_CH * RED_BLK + 8) * xyzz_bytes);
  w.scratch = take(sc);
  w.total = o;
  return w;
}

#ifdef MSM_DEFINE_LEVELS
size_t msm_sort_bytes(uint64_t n_total, int cb) { return sort_layout(n_total, cb).total; }

template <class G>
static int32_t msm_sort_g(frcs_ctx* ctx, uint64_t n_total, const ScalarSegs& sg, int mont, uint32_t nb, void* sort_work,
                          cudaStream_t st) {
  constexpr uint32_t NB = G::NB;
  MsmLevels lv = msm_levels(n_total, G::CB, nb);
  SortLayout wl = sort_layout(n_total, G::CB);
  uint8_t* w = (uint8_t*)sort_work;
  uint32_t* digits = (uint32_t*)(w + wl.digits);
  uint32_t* sorted = (uint32_t*)(w + wl.sorted);
  uint32_t* cnt = (uint32_t*)(w + wl.cnt);
  uint32_t* off = (uint32_t*)(w + wl.off);
  uint32_t* cursor = (uint32_t*)(w + wl.cursor);
  BatchStrides bs{wl.total / 4, 0, nb};
  const bool wide = G::CB == MSM_CB_WIDE;  // (diagnostic spans for the wide sort only)
  int pd = wide ? prof_begin(ctx, PROF_SORT_DIGITS, st) : -1;
  zero_hist_kernel<<<dim3((NB + 255) / 256, nb), 256, 0, st>>>(cnt, NB, bs);
  prof_end(ctx, pd, st);
  unsigned gs = (unsigned)((n_total + 255) / 256);
  // batches of dense scalars: block histograms in shared memory (block_hist_kernel); a block's slice must hold a few
  // entries per bucket for the per-(block, bucket) atomics to be fewer than the entries
  static const uint32_t block_sort_min = msm_env_u32("FRCS_BLOCK_SORT_MIN", 8);
  const bool block_sort = wide && block_sort_min > 0 && nb >= block_sort_min && n_total * G::WINDOWS >= 8ull * NB;
  if (block_sort) {
    int sms = 0;
    FRCS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    uint32_t bpp = (uint32_t)sms / nb;  // blocks per problem: one block per SM (128 KB of shared memory each)
    const uint32_t bpp_max = (uint32_t)(n_total * G::WINDOWS / (8ull * NB));
    bpp = bpp < 1 ? 1 : bpp > bpp_max ? bpp_max : bpp;
    const uint32_t per_block = (uint32_t)((n_total + bpp - 1) / bpp);
    const dim3 grid((unsigned)((n_total + per_block - 1) / per_block), nb);
    static bool attr_set = false;
    if (!attr_set) {
      FRCS_CUDA_CHECK(cudaFuncSetAttribute(block_hist_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NB * 4)));
      attr_set = true;
    }
    block_hist_kernel<G><<<grid, BS_THREADS, NB * 4, st>>>(sg, n_total, mont, per_block, cnt, bs);
    pd = wide ? prof_begin(ctx, PROF_SORT_SCATTER, st) : -1;
    plan_kernel<G><<<dim3(lv.n_levels + 1, nb), NB < 256u ? NB : 256u, 0, st>>>(cnt, off, cursor, lv, bs);
    prof_end(ctx, pd, st);
    scatter_scalar_kernel<G, false><<<dim3(gs, nb), 256, 0, st>>>(sg, n_total, mont, cursor, sorted, bs);
  } else {
    // narrow geometry: no digits array either (FRCS_SORT_DIGITS=1: the digit-array scatter, kept for small wide sorts)
    static const uint32_t use_digits = msm_env_u32("FRCS_SORT_DIGITS", 0);
    const bool from_scalars = !wide && !use_digits;
    digits_kernel<G><<<dim3(gs, nb), 256, 0, st>>>(sg, n_total, mont, from_scalars ? nullptr : digits, cnt, bs);
    pd = wide ? prof_begin(ctx, PROF_SORT_SCATTER, st) : -1;
    plan_kernel<G><<<dim3(lv.n_levels + 1, nb), NB < 256u ? NB : 256u, 0, st>>>(cnt, off, cursor, lv, bs);
    prof_end(ctx, pd, st);
    if (from_scalars)
      scatter_scalar_kernel<G, true><<<dim3(gs, nb), 256, 0, st>>>(sg, n_total, mont, cursor, sorted, bs);
    else
      scatter_kernel<<<dim3(gs, G::WINDOWS, nb), 256, 0, st>>>(digits, n_total, cursor, sorted, bs, G::CB == 16);
  }
  ctx->launches++;
  ctx->launches += 3;
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

// Signed-digit decomposition + counting sort + level plan of nb scalar vectors.
int32_t msm_sort(frcs_ctx* ctx, uint64_t n_total, const MsmScalars& sc, int mont, uint32_t nb, void* sort_work,
                 cudaStream_t st, int cb, const uint32_t* skip) {
  if (nb == 0) return FRCS_OK;
  NvtxRange nvtx("frcs:msm_sort");
  ScalarSegs sg;
  sg.skip = skip;
  uint64_t end = 0;
  for (int k = 0; k < 3; k++) {
    sg.ptr[k] = sc.ptr[k];
    sg.stride[k] = sc.stride[k];
    end += sc.ptr[k] ? sc.count[k] : 0;
    sg.end[k] = end;
  }
  if (end != n_total) {
    frcs_set_error("msm_sort: scalar segments do not add up to the number of bases");
    return FRCS_E_INVALID_ARG;
  }
  if ((n_total * msm_windows(cb)) >> 31) {
    frcs_set_error("msm_sort: too many (window, base) pairs for the 31-bit sorted entries");
    return FRCS_E_INVALID_ARG;
  }
  return cb == MSM_CB_NARROW ? msm_sort_g<Narrow>(ctx, n_total, sg, mont, nb, sort_work, st)
                             : msm_sort_g<Wide>(ctx, n_total, sg, mont, nb, sort_work, st);
}
#endif

template <class F>
size_t msm_acc_bytes(uint64_t n_total, int cb) {
  return acc_layout(n_total, 4 * sizeof(F), cb).total;
}
template size_t msm_acc_bytes<MSM_FIELD>(uint64_t, int);

template <class F>
int32_t msm_precompute(frcs_ctx* ctx, const uint32_t* d_bases, uint64_t n, uint32_t* d_pts, cudaStream_t st, int cb) {
  if (n == 0) return FRCS_OK;
  if (cb == MSM_CB_NARROW)
    precompute_kernel<F, Narrow><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_bases, d_pts, n);
  else
    precompute_kernel<F, Wide><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_bases, d_pts, n);
  ctx->launches++;
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}
template int32_t msm_precompute<MSM_FIELD>(frcs_ctx*, const uint32_t*, uint64_t, uint32_t*, cudaStream_t, int);

// Bucket accumulation + reduction for n_tables (1 or 2) pre-processed base tables over the nb
// sorted scalar vectors in sort_work: result[t][p] (XYZZ, device, result_stride words apart)
// = sum_i s_{p,i} * table_t[i].  acc_work: n_tables * nb * msm_acc_bytes<F>(n_total, cb) bytes.
template <class F, class G>
static int32_t msm_accumulate_g(frcs_ctx* ctx, uint32_t n_tables, const uint32_t* const* d_pts, uint64_t n_total,
                                uint32_t nb, const void* sort_work, void* acc_work, uint32_t* const* d_result,
                                uint64_t result_stride, cudaStream_t st, int prof_total, int prof_accum) {
  constexpr size_t XW = 4 * sizeof(F) / 4;  // words per XYZZ
  constexpr uint32_t NB = G::NB;
  MsmLevels lv = msm_levels(n_total, G::CB, nb);
  SortLayout sl = sort_layout(n_total, G::CB);
  AccLayout al = acc_layout(n_total, XW * 4, G::CB);
  const uint8_t* sw = (const uint8_t*)sort_work;
  const uint32_t* sorted = (const uint32_t*)(sw + sl.sorted);
  const uint32_t* cnt = (const uint32_t*)(sw + sl.cnt);
  const uint32_t* off = (const uint32_t*)(sw + sl.off);
  uint8_t* aw = (uint8_t*)acc_work;
  uint32_t* buf[2] = {(uint32_t*)(aw + al.buf0), (uint32_t*)(aw + al.buf1)};
  uint32_t* partial = (uint32_t*)(aw + al.partial);
  BatchStrides bs{sl.total / 4, al.total / 4, nb};
  const uint32_t nq = n_tables * nb;
  Tables tabs{{d_pts[0], n_tables > 1 ? d_pts[1] : d_pts[0]}};

  // (prefetch: measured neutral for the XYZZ pieces, 366-371 proofs/s with L1 / L2 / no prefetch; off by default)
  static const uint32_t pf = msm_env_u32("FRCS_MSM_PF", 0) | msm_env_u32("FRCS_MSM_PF_D", 2) << 8;
  int pt = prof_total >= 0 ? prof_begin(ctx, prof_total, st) : -1;
  // the profiler's "accumulation" span: level 0, or with pair levels everything up to the first XYZZ sums
  uint32_t accum_last = 0;
  for (uint32_t l = 0; l < lv.n_levels; l++)
    if (lv.kind[l] == MSM_LV_PAIR || lv.kind[l] == MSM_LV_MIXED) accum_last = l;
  int pa = -1;
  for (uint32_t l = 0; l < lv.n_levels; l++) {
    const uint32_t* o = off + (size_t)l * (NB + 1);
    const uint32_t* c = cnt + (size_t)l * NB;
    const uint32_t* on = off + (size_t)(l + 1) * (NB + 1);
    const uint32_t* src = l ? buf[(l - 1) & 1] : nullptr;
    unsigned g = (unsigned)((lv.t_max[l] + 127) / 128);
    if (l == 0 && prof_accum >= 0) {
      pa = prof_begin(ctx, prof_accum, st);
      ctx->prof.work_dev[prof_accum] = off + NB;  // off[0][NB] of problem 0 = its number of additions
      ctx->prof.work_mul[prof_accum] = nq;
    }
    switch (lv.kind[l]) {
      case MSM_LV_SEG:
        accum0_kernel<F, G><<<dim3(g, nq), 128, 0, st>>>(tabs, sorted, o, on, lv.lc[0], buf[0], pf, bs);
        break;
      case MSM_LV_PAIR: {
        const unsigned gp = (unsigned)(pair_slots(lv, l) / (128ull * lv.pair_m));
        uint32_t* scratch = (uint32_t*)(aw + al.scratch);
        if (l == 0)
          pair_kernel<F, G, true><<<dim3(gp, nq), 128, 0, st>>>(tabs, sorted, nullptr, o, c, on, lv.pair_m, buf[0], scratch,
                                                                 al.total / 4, pf, bs);
        else
          pair_kernel<F, G, false><<<dim3(gp, nq), 128, 0, st>>>(tabs, nullptr, src, o, c, on, lv.pair_m, buf[l & 1], scratch,
                                                                  al.total / 4, pf, bs);
        break;
      }
      case MSM_LV_MIXED:
        accumA_kernel<F, G><<<dim3(g, nq), 128, 0, st>>>(src, o, c, on, lv.lc[l], buf[l & 1], bs);
        break;
      default:
        accumN_kernel<F, G><<<dim3(g, nq), 128, 0, st>>>(src, o, c, on, lv.lc[l], buf[l & 1], bs);
    }
    if (l == accum_last) prof_end(ctx, pa, st);
    ctx->launches++;
  }
  uint32_t* fin = buf[(lv.n_levels - 1) & 1];
  const uint32_t* fo = off + (size_t)lv.n_levels * (NB + 1);
  const uint32_t* fc = cnt + (size_t)lv.n_levels * NB;
  finish_kernel<F, G><<<dim3(NB >= 4096u ? NB / 64 : NB, nq), 64, 64 * XW * 4, st>>>(fin, fo, fc, bs);
  ctx->launches++;
  if constexpr (NB <= 1024u) {
    const size_t smem = (size_t)NB * XW * 4;
    FRCS_CUDA_CHECK(cudaFuncSetAttribute(narrow_reduce_kernel<F, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    narrow_reduce_kernel<F, G><<<dim3(nq), NB, smem, st>>>(fin, fo, fc, d_result[0], n_tables > 1 ? d_result[1] : nullptr,
                                                          result_stride, bs);
    ctx->launches++;
  } else {
    bucket_reduce_kernel<F><<<dim3((RED_RUNS + 63) / 64, nq), 64, 0, st>>>(fin, fo, fc, partial, bs);
    reduce_channels_kernel<F><<<dim3(RED_BLK, RED_CH, nq), 64, 64 * XW * 4, st>>>(partial, bs);
    constexpr int FI = sizeof(F) == sizeof(Fq) ? 0 : 1;
    if (!ctx->red_corr[FI]) {
      FRCS_CUDA_CHECK(cudaMalloc(&ctx->red_corr[FI], XW * 4));
      reduce_corr_kernel<F><<<1, 1, 0, st>>>(ctx->red_corr[FI]);
      ctx->launches++;
      FRCS_CUDA_CHECK(cudaStreamSynchronize(st));  // once per context: other streams read it without an event
    }
    reduce_combine_kernel<F><<<dim3(nq), 64, 64 * XW * 4, st>>>(partial, ctx->red_corr[FI], d_result[0],
                                                               n_tables > 1 ? d_result[1] : nullptr, result_stride, bs);
    ctx->launches += 3;
  }
  prof_end(ctx, pt, st);
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

template <class F>
int32_t msm_accumulate(frcs_ctx* ctx, uint32_t n_tables, const uint32_t* const* d_pts, uint64_t n_total, uint32_t nb,
                       const void* sort_work, void* acc_work, uint32_t* const* d_result, uint64_t result_stride,
                       cudaStream_t st, int cb, int prof_total, int prof_accum) {
  if (nb == 0 || n_tables == 0) return FRCS_OK;
  NvtxRange nvtx(sizeof(F) == sizeof(Fq) ? "frcs:msm_accumulate_g1" : "frcs:msm_accumulate_g2");
  if (cb == MSM_CB_NARROW)
    return msm_accumulate_g<F, Narrow>(ctx, n_tables, d_pts, n_total, nb, sort_work, acc_work, d_result, result_stride, st,
                                       prof_total, prof_accum);
  return msm_accumulate_g<F, Wide>(ctx, n_tables, d_pts, n_total, nb, sort_work, acc_work, d_result, result_stride, st,
                                   prof_total, prof_accum);
}
template int32_t msm_accumulate<MSM_FIELD>(frcs_ctx*, uint32_t, const uint32_t* const*, uint64_t, uint32_t, const void*,
                                           void*, uint32_t* const*, uint64_t, cudaStream_t, int, int, int);

// RAII device allocations of the stand-alone entry points
struct MsmDevBuf {
  void* p = nullptr;
  ~MsmDevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

template <class F>
static int32_t msm_api(frcs_ctx* ctx, int cb, uint64_t n, const uint64_t* bases, const uint64_t* scalars, uint64_t* out) {
  if (!ctx || !bases || !scalars || !out || (cb != MSM_CB_WIDE && cb != MSM_CB_NARROW)) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  constexpr size_t AB = 2 * sizeof(F);
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    memset(out, 0, AB);
    return FRCS_OK;
  }
  MsmDevBuf d_bases, d_pts, d_sc, d_res, work, swork;
  FRCS_CUDA_CHECK(d_bases.alloc(n * AB));
  FRCS_CUDA_CHECK(d_pts.alloc(n * AB * msm_windows(cb)));
  FRCS_CUDA_CHECK(d_sc.alloc(n * 32));
  FRCS_CUDA_CHECK(d_res.alloc(3 * AB));
  FRCS_CUDA_CHECK(work.alloc(msm_acc_bytes<F>(n, cb)));
  FRCS_CUDA_CHECK(swork.alloc(msm_sort_bytes(n, cb)));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_bases.p, bases, n * AB, cudaMemcpyHostToDevice, st));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_sc.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
  int32_t rc = msm_precompute<F>(ctx, (const uint32_t*)d_bases.p, n, (uint32_t*)d_pts.p, st, cb);
  if (!rc) {
    MsmScalars sc{{(const uint32_t*)d_sc.p, nullptr, nullptr}, {0, 0, 0}, {n, 0, 0}};
    rc = msm_sort(ctx, n, sc, 0, 1, swork.p, st, cb);
  }
  if (!rc) {
    const uint32_t* tabs[1] = {(const uint32_t*)d_pts.p};
    uint32_t* outs[1] = {(uint32_t*)d_res.p};
    rc = msm_accumulate<F>(ctx, 1, tabs, n, 1, swork.p, work.p, outs, 0, st, cb, -1, -1);
  }
  if (!rc) {
    uint32_t* res = (uint32_t*)d_res.p;
    to_affine_kernel<F><<<1, 1, 0, st>>>(res, res + 2 * AB / 4);
    ctx->launches++;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(out, res + 2 * AB / 4, AB, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaGetLastError());
  }
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));  // also on the error paths: the buffers are freed on return
  return rc;
}

extern "C" int32_t MSM_API_NAME(frcs_ctx* ctx, uint64_t n, const uint64_t* bases, const uint64_t* scalars, uint64_t* out) {
  return msm_api<MSM_FIELD>(ctx, MSM_CB_WIDE, n, bases, scalars, out);
}
// test hook: the same MSM through either window geometry (window_bits = 16 or 8)
extern "C" int32_t MSM_API_WB_NAME(frcs_ctx* ctx, int32_t window_bits, uint64_t n, const uint64_t* bases,
                                   const uint64_t* scalars, uint64_t* out) {
  return msm_api<MSM_FIELD>(ctx, window_bits, n, bases, scalars, out);
}

// debug/test hook: the pre-processed table of n bases (16 windows x n affine points)
extern "C" int32_t MSM_DEBUG_NAME(frcs_ctx* ctx, uint64_t n, const uint64_t* bases, uint64_t* out) {
  typedef MSM_FIELD F;
  constexpr size_t AB = 2 * sizeof(F);
  if (!ctx || !bases || !out) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const size_t W = msm_windows(MSM_CB_WIDE);
  MsmDevBuf d_bases, d_pts;
  FRCS_CUDA_CHECK(d_bases.alloc(n * AB));
  FRCS_CUDA_CHECK(d_pts.alloc(n * AB * W));
  // on the context's (non-blocking) stream: a legacy-stream copy from pageable memory is not ordered with it
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_bases.p, bases, n * AB, cudaMemcpyHostToDevice, ctx->stream));
  int32_t rc = msm_precompute<F>(ctx, (const uint32_t*)d_bases.p, n, (uint32_t*)d_pts.p, ctx->stream, MSM_CB_WIDE);
  if (!rc) FRCS_CUDA_CHECK(cudaMemcpyAsync(out, d_pts.p, n * AB * W, cudaMemcpyDeviceToHost, ctx->stream));
  FRCS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return rc;
}
