// NVTX ranges around the launch sites of every stage ("NVTX ranges + CUDA-event timers per
// stage"): with Nsight Systems attached the kernels of a stage appear under its range; without a tool the calls
// are no-ops (nvtx3 is header-only and loads the injection library lazily).
#pragma once
#include <nvtx3/nvToolsExt.h>

struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
