// Forward GEMM: run-time-everything variants (any epilogue feature combination, every tile width).  Slower
// than the specialised variants (larger code, more instructions per element); used for rare shapes only.
#include "gemm_kernel.cuh"
namespace dp {
extern const GemmVariant kGemmVariantsGeneric[] = {
    DP_GEMM_VARIANT(32, EO_RUNTIME, EA_RUNTIME, EM_RUNTIME, OP_ALL),
    DP_GEMM_VARIANT(64, EO_RUNTIME, EA_RUNTIME, EM_RUNTIME, OP_ALL),
    DP_GEMM_VARIANT(128, EO_RUNTIME, EA_RUNTIME, EM_RUNTIME, OP_ALL),
    DP_GEMM_VARIANT(192, EO_RUNTIME, EA_RUNTIME, EM_RUNTIME, OP_ALL),
    DP_GEMM_VARIANT(256, EO_RUNTIME, EA_RUNTIME, EM_RUNTIME, OP_ALL),
};
extern const int kNumGemmVariantsGeneric = sizeof(kGemmVariantsGeneric) / sizeof(kGemmVariantsGeneric[0]);
}  // namespace dp
