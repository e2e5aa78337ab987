// Forward GEMM variants of the pose heads in eval mode (folded BatchNorm scale, run-time activation).
#include "gemm_kernel.cuh"
namespace dp {
extern const GemmVariant kGemmVariantsD[] = {
    DP_GEMM_VARIANT(64, EO_BF16, EA_RUNTIME, EM_IDENTITY, OP_SCALE | OP_CONV),
    DP_GEMM_VARIANT(128, EO_BF16, EA_RUNTIME, EM_IDENTITY, OP_SCALE | OP_CONV),
    DP_GEMM_VARIANT(256, EO_BF16, EA_RUNTIME, EM_IDENTITY, OP_SCALE | OP_CONV),
    DP_GEMM_VARIANT(128, EO_BF16, EA_RUNTIME, EM_SHUFFLE, OP_SCALE),
    DP_GEMM_VARIANT(256, EO_BF16, EA_RUNTIME, EM_SHUFFLE, OP_SCALE),
};
extern const int kNumGemmVariantsD = sizeof(kGemmVariantsD) / sizeof(kGemmVariantsD[0]);
}  // namespace dp
