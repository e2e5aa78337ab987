// Kernel launch helper shared by every launch site of the library.
//
// A fine-tuning step is ~260 dependent launches replayed as one CUDA graph, many of them 5-30 us long, so the
// kernel-to-kernel hand-over (grid drain, launch latency, prologue of the next kernel) is a double-digit share of the
// step.  Two measures, both applied here so that no launch site can forget them:
//   * programmatic dependent launch: every kernel of the library starts with pdl_grid_sync() (griddepcontrol
//     launch_dependents + wait, see below) and is launched with the programmatic-stream-serialization attribute, so
//     the CTAs of launch i+1 are scheduled while launch i drains, run their prologue (barrier init, TMEM allocation,
//     tensor-map prefetch, parameter loads from the constant bank) and block in griddepcontrol.wait until launch i has
//     completed and its memory is visible.  Stream capture turns the attribute into programmatic graph edges.
//   * one shared-memory carve-out for all kernels (the maximum): the GEMMs need 226 KB, the row-wise kernels none;
//     alternating preferences forces an SM-wide L1/shared reconfiguration between launches, which needs an idle SM.
// DP_PDL=0 / DP_CARVEOUT=0 switch either off (A/B measurements, debugging).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace dp {

// Device side: allow the next launch in the stream to start being scheduled, then wait until every launch this one
// depends on has completed (no-ops when the kernel was launched without the attribute).  MUST run before the first
// global-memory access of the kernel; correctness of launch i+2 relies on launch i+1 not completing before i has.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
inline bool pdl_enabled() {
  static const int v = env_flag("DP_PDL", 1);
  return v != 0;
}
inline bool carveout_enabled() {
  static const int v = env_flag("DP_CARVEOUT", 1);
  return v != 0;
}

template <auto Kern, typename... Args>
cudaError_t launch_k(dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  static bool prepared = false;
  if (!prepared) {
    if (carveout_enabled())
      cudaFuncSetAttribute(Kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    prepared = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, Kern, args...);
}

}  // namespace dp
