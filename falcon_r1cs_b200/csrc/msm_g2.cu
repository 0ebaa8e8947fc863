// G2 instantiation of the MSM subsystem (see msm_impl.cuh).
#define MSM_FIELD ff::Fq2
#define MSM_API_NAME frcs_msm_g2
#define MSM_DEBUG_NAME frcs_debug_windows_g2
#include "msm_impl.cuh"
