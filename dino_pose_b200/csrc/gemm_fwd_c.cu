// Forward GEMM variants of the pose heads: training convolutions (fp32 pre-BatchNorm output + fused batch
// statistics), their input gradients (bf16, optional bf16 residual), eval convolutions (folded BatchNorm scale).
#include "gemm_kernel.cuh"
namespace dp {
extern const GemmVariant kGemmVariantsC[] = {
    DP_GEMM_VARIANT(64, EO_F32, EA_NONE, EM_IDENTITY, OP_STATS | OP_CONV),
    DP_GEMM_VARIANT(128, EO_F32, EA_NONE, EM_IDENTITY, OP_STATS | OP_CONV),
    DP_GEMM_VARIANT(256, EO_F32, EA_NONE, EM_IDENTITY, OP_STATS | OP_CONV),
    DP_GEMM_VARIANT(128, EO_F32, EA_NONE, EM_SHUFFLE, OP_STATS),
    DP_GEMM_VARIANT(256, EO_F32, EA_NONE, EM_SHUFFLE, OP_STATS),
    DP_GEMM_VARIANT(64, EO_BF16, EA_NONE, EM_IDENTITY, OP_CONV | OP_RES_BF16),
    DP_GEMM_VARIANT(128, EO_BF16, EA_NONE, EM_IDENTITY, OP_CONV | OP_RES_BF16),
    DP_GEMM_VARIANT(256, EO_BF16, EA_NONE, EM_IDENTITY, OP_CONV | OP_RES_BF16),
};
extern const int kNumGemmVariantsC = sizeof(kGemmVariantsC) / sizeof(kGemmVariantsC[0]);
}  // namespace dp
