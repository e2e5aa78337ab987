// A-stationary (A in tensor memory) forward GEMM variants: the K <= 512 linear layers of the backbone.
#include "gemm_astat.cuh"
namespace dp {
extern const GemmVariant kGemmVariantsAstat[] = {
    DP_GEMM_ASTAT_VARIANT(EO_BF16, EA_NONE, OP_TMA_OUT),
    DP_GEMM_ASTAT_VARIANT(EO_BF16, EA_GELU, OP_TMA_OUT),
    DP_GEMM_ASTAT_VARIANT(EO_BF16, EA_NONE, 0),
    DP_GEMM_ASTAT_VARIANT(EO_BF16, EA_GELU, OP_AUX_OUT),
    DP_GEMM_ASTAT_VARIANT(EO_BF16, EA_NONE, OP_AUX_IN),
    DP_GEMM_ASTAT_VARIANT(EO_F32, EA_NONE, OP_LSRES),
    DP_GEMM_ASTAT_CL2_VARIANT(EO_BF16, EA_NONE, OP_TMA_OUT),
    DP_GEMM_ASTAT_CL2_VARIANT(EO_BF16, EA_GELU, OP_TMA_OUT),
    DP_GEMM_ASTAT_CL2_VARIANT(EO_BF16, EA_NONE, 0),
    DP_GEMM_ASTAT_CL2_VARIANT(EO_BF16, EA_GELU, OP_AUX_OUT),
    DP_GEMM_ASTAT_CL2_VARIANT(EO_BF16, EA_NONE, OP_AUX_IN),
    DP_GEMM_ASTAT_CL2_VARIANT(EO_F32, EA_NONE, OP_LSRES),
};
extern const int kNumGemmVariantsAstat = sizeof(kGemmVariantsAstat) / sizeof(kGemmVariantsAstat[0]);
}  // namespace dp
