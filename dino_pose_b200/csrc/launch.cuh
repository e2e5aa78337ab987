// Kernel launch helper shared by every launch site of the library.
//
// A fine-tuning step is ~260 dependent launches replayed as one CUDA graph, many of them 5-30 us long.  Two launch-level
// measures were built here (one place, so that no launch site can forget them) and MEASURED on the benchmark step
// (gpurun_out ab1, tools/gpu_job_pdl_ab.sh):
//   * programmatic dependent launch: every kernel of the library starts with pdl_grid_sync() (griddepcontrol
//     launch_dependents + wait) and can be launched with the programmatic-stream-serialization attribute, so the CTAs
//     of launch i+1 are scheduled while launch i drains, run their prologue (barrier init, TMEM allocation, tensor-map
//     prefetch) and block in griddepcontrol.wait until launch i has completed.  Stream capture turns the attribute
//     into programmatic graph edges.  Result: correct (GPU suite green) but SLOWER, 4.73 vs 4.59 ms per step: graph
//     kernel nodes already hand over in ~1 us, and the early-resident CTAs of the next launch cost more than that.
//   * one shared-memory carve-out (the maximum) for all kernels, to avoid L1/shared reconfiguration between the
//     226 KB GEMMs and the row-wise kernels: 4.66 vs 4.59 ms, also slower (the row-wise kernels lose their L1).
// Both therefore default to OFF; DP_PDL=1 / DP_CARVEOUT=1 switch them on (the device-side prologue is a no-op when
// the kernel was launched without the attribute).
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
  static const int v = env_flag("DP_PDL", 0);
  return v != 0;
}
inline bool carveout_enabled() {
  static const int v = env_flag("DP_CARVEOUT", 0);
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
