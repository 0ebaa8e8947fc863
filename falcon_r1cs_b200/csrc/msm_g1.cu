// G1 instantiation of the MSM subsystem (see msm_impl.cuh).
#define FF_INLINE_MUL
#define MSM_FIELD ff::Fq
#ifdef FRCS_G1_ACCUM_BLOCKS
#define ACCUM0_MIN_BLOCKS FRCS_G1_ACCUM_BLOCKS
#else
#define ACCUM0_MIN_BLOCKS 3
#endif
#define MSM_API_NAME frcs_msm_g1
#define MSM_API_WB_NAME frcs_debug_msm_g1
#define MSM_DEFINE_LEVELS
#define MSM_DEBUG_NAME frcs_debug_windows_g1
#include "msm_impl.cuh"
