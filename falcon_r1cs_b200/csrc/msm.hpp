// Internal interface of the MSM subsystem (msm_impl.cuh, instantiated in msm_g1.cu / msm_g2.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#define MSM_NB 32768u    // buckets (|signed 16-bit digit| - 1)
#define MSM_WINDOWS 16u  // 16-bit windows covering the 255-bit scalar
#define MSM_MAX_LEVELS 8

struct frcs_ctx;

struct MsmLevels {
  uint32_t n_levels;
  uint32_t lc[MSM_MAX_LEVELS];     // slice length per level
  uint64_t t_max[MSM_MAX_LEVELS];  // upper bound on the number of slices per level
};
MsmLevels msm_levels(uint64_t n_total);

// Scalars of a batch of MSM problems: up to three consecutive segments (e.g. the assignment z,
// then the randomiser scalars, then the quotient polynomial h); problem p reads segment k at
// ptr[k] + p * stride[k] (u32 words), count[k] scalars of 8 words each.  Unused: ptr = nullptr.
struct MsmScalars {
  const uint32_t* ptr[3];
  uint64_t stride[3];
  uint64_t count[3];
};

size_t msm_sort_bytes(uint64_t n_total);
int32_t msm_sort(frcs_ctx* ctx, uint64_t n_total, const MsmScalars& sc, int mont, uint32_t nb, void* sort_work,
                 cudaStream_t st);

// F = ff::Fq (G1) or ff::Fq2 (G2)
template <class F> size_t msm_acc_bytes(uint64_t n_total);
template <class F> int32_t msm_precompute(frcs_ctx* ctx, const uint32_t* d_bases, uint64_t n, uint32_t* d_pts, cudaStream_t st);
template <class F> int32_t msm_accumulate(frcs_ctx* ctx, uint32_t n_tables, const uint32_t* const* d_pts, uint64_t n_total,
                                          uint32_t nb, const void* sort_work, void* acc_work, uint32_t* const* d_result,
                                          uint64_t result_stride, cudaStream_t st, int prof_total = -1, int prof_accum = -1);
