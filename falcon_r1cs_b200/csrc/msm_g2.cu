// G2 instantiation of the MSM subsystem (see msm_impl.cuh).
#define MSM_FIELD ff::Fq2
#ifdef FRCS_G2_ACCUM_BLOCKS  // tuning: resident blocks per SM the G2 accumulation kernel is compiled for (default 3)
#define ACCUM0_MIN_BLOCKS FRCS_G2_ACCUM_BLOCKS
#endif
#define MSM_API_NAME frcs_msm_g2
#define MSM_API_WB_NAME frcs_debug_msm_g2
#define MSM_DEBUG_NAME frcs_debug_windows_g2
#include "msm_impl.cuh"
