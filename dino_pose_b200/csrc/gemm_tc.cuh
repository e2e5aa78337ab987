// Parameter blocks shared between the host launchers (c_api.cu) and the tcgen05 kernels (gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace dp {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
enum : int { ROWMAP_IDENTITY = 0, ROWMAP_PATCH_TOKENS = 1, ROWMAP_NCHW = 2, ROWMAP_SHUFFLE2X2 = 3 };
enum : int { OUT_BF16 = 0, OUT_F32 = 1 };

// Epilogue of the K-major GEMM / implicit conv:   (per output element, column c, logical row r)
//   v = acc * scale[c] + bias[c]
//   aux_out[r,c] = bf16(act == GELU ? gelu'(v) : v)   (optional: what the backward multiplies with / the branch output)
//   v = act(v)
//   v = v * aux_in[r,c]                       (optional: fused GELU backward for the fc2->fc1 dgrad)
//   v = residual[rr,c] + v * ls[c]            (optional: LayerScale + residual / pos-emb add)
//   out[map(r),c] = v
struct Epilogue {
  void* out;
  const float* bias;
  const float* scale;
  const float* ls;
  const float* residual;
  void* aux_out;
  const void* aux_in;
  double* stats;    // optional fused BatchNorm statistics: stats[c] += sum_rows v, stats[stats_c + c] += sum_rows v^2
  long long ldo, ldr, ld_aux;
  int out_dtype;
  int act;
  int row_map;
  int res_is_bf16;  // residual tensor dtype (0 fp32, 1 bf16)
  int n_valid;      // number of valid output columns (<= N)
  int stats_c;      // number of channels of `stats`
  long long* trace; // debug: per-tile timestamps of CTA 0 (DP_GEMM_TRACE=1), [tile][4] = mma start, mma issued, epilogue start, epilogue end
  int debug;        // tuning knobs (DP_GEMM_DEBUG env): 1 skip stores, 2 skip residual/aux loads, 4 skip phase 2, 8 skip TMEM loads
  int map_a, map_b; // ROWMAP_PATCH_TOKENS: (patches per image, tokens per image); NCHW: (channels K, 0);
                    // SHUFFLE2X2: (Cout, 0)
  // row-owning kernel with the next LayerNorm fused (gemm_rowln.cu): ln_out = LayerNorm(out row; ln_gamma, ln_beta, ln_eps)
  const float* ln_gamma;
  const float* ln_beta;
  void* ln_out;     // bf16 [rows, ld_ln]
  long long ld_ln;
  float ln_eps;
  // same kernel, LoRA mode: x_out = residual + ls * (y + lora_scaling * dropout(y lora_A lora_B)); ln_out then receives y (fp32)
  const float* lora_A;                  // [N, 8]
  const float* lora_B;                  // [8, N]
  float* lora_u_out;                    // [rows, 8] (saved for the backward) or nullptr
  const unsigned long long* lora_seed;  // device scalar
  float lora_scaling, lora_p_drop;
};

struct alignas(64) GemmParams {
  CUtensorMap tmA;  // plain: 2D {K, M} box {64,128}; conv: 4D {C, W, H, B} box {64, bw, bh, bb}
  CUtensorMap tmB;  // 2D {Ktotal, N} box {64, BN}
  CUtensorMap tmC;  // OP_TMA_OUT variants only: bf16 output, 2D {n_valid, M} box {32, 32}, 64B swizzle
  Epilogue epi;
  int M, N, num_k_blocks;
  int m_tiles, n_tiles;
  int a_mode;               // 0 plain, 1 implicit conv
  int kw, pad_x, pad_y, cin_blocks;
  int bw, bh, bb;           // pixel box of one 128-row tile (powers of two)
  int bw_log2, bh_log2;
  int OW, OH, NB;           // output extents (conv) -- also used by the row maps
  int tiles_x, tiles_y;
};

// Weight-gradient GEMM (both operands MN-major, reduction over pixels, split-K, fp32 atomics).
//   out[offm(m) + offn(n) + tap*so_t] += sum_p A[p, m] * B[p (+tap shift), n]
struct alignas(64) WgradParams {
  CUtensorMap tmA;  // 4D {Cm, W, H, B} box {64, bw, bh, bb} (64 pixels)  or 2D {Cm, P} box {64,64}
  CUtensorMap tmB;  // same for the N-side operand
  float* out;
  long long so_m, so_mo, so_n, so_no, so_t;
  int m_inner, n_inner;     // off(m) = (m % m_inner) * so_m + (m / m_inner) * so_mo
  int Mc, Nc;
  int m_tiles, n_tiles;
  int mode;                 // 0 plain 2D, 1 conv 4D
  int taps, kw, pad_x, pad_y;
  int bw, bh, bb;
  int tiles_x, tiles_y, tiles_b;
  int total_k_blocks, splits;
  float* ws;                // split-K workspace [splits][taps][m_tiles*128][n_tiles*BN] fp32 (nullptr: fp32 atomics into out)
  int ws_ld;                // n_tiles * BN
  int debug;                // DP_WGRAD_DEBUG: 1 skip the atomic scatter (tuning aid)
};

}  // namespace dp
