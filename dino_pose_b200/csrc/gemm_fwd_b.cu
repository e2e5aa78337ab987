// Forward GEMM variants: fp32 residual-stream outputs (attention / MLP output projections with LayerScale +
// residual, patch embedding with the position-embedding add and the token row map).
#include "gemm_kernel.cuh"
namespace dp {
extern const GemmVariant kGemmVariantsB[] = {
    DP_GEMM_VARIANT(128, EO_F32, EA_NONE, EM_IDENTITY, OP_LSRES),
    DP_GEMM_VARIANT(192, EO_F32, EA_NONE, EM_IDENTITY, OP_LSRES),
    DP_GEMM_VARIANT(256, EO_F32, EA_NONE, EM_IDENTITY, OP_LSRES),
    DP_GEMM_VARIANT(128, EO_F32, EA_NONE, EM_PATCH, OP_LSRES),
    DP_GEMM_VARIANT(192, EO_F32, EA_NONE, EM_PATCH, OP_LSRES),
};
extern const int kNumGemmVariantsB = sizeof(kGemmVariantsB) / sizeof(kGemmVariantsB[0]);
}  // namespace dp
