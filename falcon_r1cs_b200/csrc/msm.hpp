// Internal interface of the MSM subsystem (msm_impl.cuh, instantiated in msm_g1.cu / msm_g2.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

// Window geometry: CB-bit signed digits; all 256 / CB windows share one set of 2^(CB-1) buckets (bucket b holds
// the points whose digit has magnitude b + 1), because every base is kept pre-multiplied by 2^(CB k).
//  * WIDE (16 bits, 16 windows, 32768 buckets): dense 255-bit scalars -- the quotient polynomial h against
//    h_query, merged with l_query; also the generic frcs_msm_g1/g2 entry points.
//  * NARROW (8 bits, 32 windows, 128 buckets): the assignment z against a_query / b_g1_query / b_g2_query.  z is
//    ~37 % zeros, ~54 % ones, ~8 % values below 2^28 and 2N quotients of ~146 bits: ~150 k
//    non-zero digits per proof, for which a 32768-bucket reduction (65 k full additions, five latency-bound
//    launches) cost more than the accumulation itself.  With 128 buckets the reduction is one block per problem.
#define MSM_CB_WIDE 16
#define MSM_CB_NARROW 8
#define MSM_MAX_LEVELS 8

static inline uint32_t msm_windows(int cb) { return 256u / (uint32_t)cb; }
static inline uint32_t msm_buckets(int cb) { return 1u << (cb - 1); }

struct frcs_ctx;

// Bucket accumulation runs in levels; level l turns the lists of a bucket's entries into shorter lists of sums.
enum MsmLevelKind : uint32_t {
  MSM_LV_SEG = 0,    // level 0 only: pieces of lc consecutive entries of the whole sorted list, table points -> XYZZ sums
  MSM_LV_PAIR = 1,   // pairs of a bucket's entries (table points at level 0, affine sums above) -> affine sums, the
                     // inversions shared by pair_m pairs per thread ("batched affine")
  MSM_LV_MIXED = 2,  // slices of lc affine sums of a bucket -> XYZZ sums (mixed additions)
  MSM_LV_XYZZ = 3,   // slices of lc XYZZ sums of a bucket -> XYZZ sums
};
struct MsmLevels {
  uint32_t n_levels;
  uint32_t kind[MSM_MAX_LEVELS];
  uint32_t lc[MSM_MAX_LEVELS];     // slice length per level
  uint64_t t_max[MSM_MAX_LEVELS];  // upper bound on the number of sums a level writes
  uint32_t pair_m;                 // pairs per thread of the MSM_LV_PAIR levels
};
// nb = number of scalar vectors sorted (and accumulated) together: the pair levels need many pairs per thread to pay for
// their inversions and enough threads to fill the GPU, so they are used for large batches only
MsmLevels msm_levels(uint64_t n_total, int cb, uint32_t nb);

// Scalars of a batch of MSM problems: up to three consecutive segments (e.g. the assignment z,
// then the randomiser scalars, then the quotient polynomial h); problem p reads segment k at
// ptr[k] + p * stride[k] (u32 words), count[k] scalars of 8 words each.  Unused: ptr = nullptr.
struct MsmScalars {
  const uint32_t* ptr[3];
  uint64_t stride[3];
  uint64_t count[3];
};

size_t msm_sort_bytes(uint64_t n_total, int cb);
// skip: optional device bitmask over the n_total bases (bit i of word i / 32); the scalars of the marked bases count as
// zero (bases that are the point at infinity: ~41 % of the b_g1 / b_g2 queries of the Falcon circuits)
int32_t msm_sort(frcs_ctx* ctx, uint64_t n_total, const MsmScalars& sc, int mont, uint32_t nb, void* sort_work,
                 cudaStream_t st, int cb, const uint32_t* skip = nullptr);

// F = ff::Fq (G1) or ff::Fq2 (G2)
template <class F> size_t msm_acc_bytes(uint64_t n_total, int cb);
template <class F> int32_t msm_precompute(frcs_ctx* ctx, const uint32_t* d_bases, uint64_t n, uint32_t* d_pts, cudaStream_t st,
                                          int cb);
template <class F> int32_t msm_accumulate(frcs_ctx* ctx, uint32_t n_tables, const uint32_t* const* d_pts, uint64_t n_total,
                                          uint32_t nb, const void* sort_work, void* acc_work, uint32_t* const* d_result,
                                          uint64_t result_stride, cudaStream_t st, int cb, int prof_total = -1,
                                          int prof_accum = -1);
