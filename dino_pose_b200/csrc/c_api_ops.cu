// extern "C" wrappers for the row-wise / head / decode kernels (argument checks + launch).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/dinopose.h"

namespace dp {
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
int sm_count();

cudaError_t launch_layernorm_fwd(const float*, const float*, const float*, __nv_bfloat16*, float*, long long, int, int, int,
                                 float, cudaStream_t);
cudaError_t launch_layernorm_bwd(const void*, int, const float*, const float*, const float*, float*, const float*,
                                 __nv_bfloat16*, long long, int, int, int, float, cudaStream_t);
cudaError_t launch_patch_im2col(const float*, __nv_bfloat16*, int, int, int, int, cudaStream_t);
cudaError_t launch_fill_cls(float*, const float*, int, int, int, cudaStream_t);
cudaError_t launch_lora_fwd(const float*, const float*, const float*, const float*, const float*, float*, float*, long long,
                            int, int, float, float, const unsigned long long*, int, cudaStream_t);
cudaError_t launch_lora_bwd(const float*, const float*, const float*, const float*, const float*, float*, float*, float*,
                            long long, int, int, float, float, const unsigned long long*, int, cudaStream_t);
cudaError_t launch_attention_fwd(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, float, cudaStream_t);
cudaError_t launch_attention_bwd(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, float*, int,
                                 int, int, float, cudaStream_t);
cudaError_t launch_layernorm_bwd_params(const void*, int, const float*, float*, float*, long long, int, float, int, cudaStream_t);
cudaError_t launch_colsum_prod(const float*, const __nv_bfloat16*, float*, long long, int, cudaStream_t);
cudaError_t launch_pred1x1_fwd(const __nv_bfloat16*, const float*, const float*, float*, long long, int, int, cudaStream_t);
cudaError_t launch_pred1x1_bwd(const float*, const __nv_bfloat16*, const float*, __nv_bfloat16*, float*, float*, long long, int,
                               int, int, cudaStream_t);
cudaError_t launch_decode(const float*, int, int, int, double, double, int*, double*, float*, cudaStream_t);
cudaError_t launch_im2col(const void*, void*, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
cudaError_t launch_col2im(const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
cudaError_t launch_dwconv3x3(const void*, const float*, const float*, const void*, void*, int, int, int, int, int, int, cudaStream_t);
cudaError_t launch_dwconv3x3_wgrad(const void*, const void*, float*, int, int, int, int, cudaStream_t);
cudaError_t launch_bn_stats(const void*, int, double*, long long, int, cudaStream_t);
cudaError_t launch_bn_finalize(double*, const float*, const float*, float*, float*, float*, float*, float*, float*, int,
                               double, float, float, cudaStream_t);
cudaError_t launch_bn_fold_eval(const float*, const float*, const float*, const float*, const float*, float*, float*, float*,
                                float*, int, float, cudaStream_t);
long long preprocess_workspace_bytes(int, int, int, int, int);
cudaError_t launch_preprocess(const void*, int, int, int, int, int, const float*, const float*, float*, void*, long long, int*,
                              cudaStream_t);
cudaError_t launch_bn_finalize_apply(const void*, int, double*, const float*, const float*, float*, float*, float*, float*, float*,
                                     float*, const void*, const void*, void*, long long, int, int, int, float, float, cudaStream_t);
cudaError_t launch_bn_apply(const void*, int, const float*, const float*, const void*, const void*, void*, long long, int,
                            int, int, cudaStream_t);
cudaError_t launch_bn_bwd_reduce(const void*, const void*, int, const void*, const float*, const float*, const float*,
                                 const float*, double*, long long, int, int, int, cudaStream_t);
cudaError_t launch_bn_bwd_apply(const void*, const void*, int, const void*, const float*, const float*, const float*,
                                const float*, const float*, double*, void*, void*, float*, float*, long long, int, int,
                                int, int, int, int, cudaStream_t);
cudaError_t launch_avgpool2(const float*, float*, long long, int, int, cudaStream_t);
cudaError_t launch_hm_grad_to_nhwc(const float*, void*, int, int, int, int, int, int, cudaStream_t);
cudaError_t launch_mean_tokens(const void*, float*, int, int, int, cudaStream_t);
cudaError_t launch_mean_tokens_bwd(void*, const float*, int, int, int, cudaStream_t);
cudaError_t launch_sgemm_small(const float*, long long, long long, const float*, long long, long long, float*, long long,
                               int, int, int, const float*, int, const float*, long long, float, const unsigned long long*, int,
                               cudaStream_t);
cudaError_t launch_colsum(const void*, int, float*, long long, int, long long, cudaStream_t);
cudaError_t launch_relu_mask(const float*, const float*, float*, long long, float, cudaStream_t);
}  // namespace dp

using namespace dp;
#define ST static_cast<cudaStream_t>(stream)

static bool ln_dim_ok(int D) { return D == 128 || D == 384 || D == 768 || D == 1024; }

extern "C" int dp_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* y32,
                                long long rows, int D, int T, int drop_cls, float eps, void* stream) {
  if (!x || !gamma || !beta || (!y && !y32)) return set_error(-1, "dp_layernorm_fwd: null pointer");
  if (!ln_dim_ok(D)) return set_error(-2, "dp_layernorm_fwd: unsupported D=%d", D);
  if (drop_cls && (T <= 1 || rows % T)) return set_error(-3, "dp_layernorm_fwd: rows %% T != 0");
  return cuda_error(launch_layernorm_fwd(x, gamma, beta, static_cast<__nv_bfloat16*>(y), y32, rows, D, T, drop_cls, eps, ST),
                    "dp_layernorm_fwd");
}
extern "C" int dp_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* add_in,
                                float* dx, const float* ls, void* dx_scaled, long long rows, int D, int T, int drop_cls,
                                float eps, void* stream) {
  if (!dy || !x || !gamma || !dx) return set_error(-1, "dp_layernorm_bwd: null pointer");
  if (!ln_dim_ok(D)) return set_error(-2, "dp_layernorm_bwd: unsupported D=%d", D);
  if (dx_scaled && !ls) return set_error(-3, "dp_layernorm_bwd: dx_scaled needs ls");
  return cuda_error(launch_layernorm_bwd(dy, dy_is_bf16, x, gamma, add_in, dx, ls, static_cast<__nv_bfloat16*>(dx_scaled),
                                         rows, D, T, drop_cls, eps, ST),
                    "dp_layernorm_bwd");
}
extern "C" int dp_patch_im2col(const float* px, void* out, int B, int H, int W, int Kp, void* stream) {
  if (!px || !out) return set_error(-1, "dp_patch_im2col: null pointer");
  if (H % 14 || W % 14 || Kp < 588 || Kp % 8) return set_error(-2, "dp_patch_im2col: bad geometry H=%d W=%d Kp=%d", H, W, Kp);
  return cuda_error(launch_patch_im2col(px, static_cast<__nv_bfloat16*>(out), B, H, W, Kp, ST), "dp_patch_im2col");
}
extern "C" int dp_fill_cls(float* x, const float* cls_row, int B, int T, int D, void* stream) {
  if (!x || !cls_row) return set_error(-1, "dp_fill_cls: null pointer");
  return cuda_error(launch_fill_cls(x, cls_row, B, T, D, ST), "dp_fill_cls");
}
extern "C" int dp_lora_fwd(const float* y, const float* A, const float* B, const float* lambda1, const float* x_in,
                           float* x_out, float* u_save, long long rows, int D, int R, float scaling, float p_drop,
                           const unsigned long long* seed, void* stream) {
  if (!y || !A || !B || !lambda1 || !x_in || !x_out) return set_error(-1, "dp_lora_fwd: null pointer");
  if (R != 4 && R != 8 && R != 16) return set_error(-2, "dp_lora_fwd: rank %d not in {4,8,16}", R);
  return cuda_error(launch_lora_fwd(y, A, B, lambda1, x_in, x_out, u_save, rows, D, R, scaling, p_drop, seed, sm_count(), ST),
                    "dp_lora_fwd");
}
extern "C" int dp_lora_bwd(const float* g, const float* y, const float* u_saved, const float* B, const float* lambda1,
                           float* dA, float* dB, float* gu_ws, long long rows, int D, int R, float scaling, float p_drop,
                           const unsigned long long* seed, void* stream) {
  if (!g || !y || !u_saved || !B || !lambda1 || !dA || !dB || !gu_ws) return set_error(-1, "dp_lora_bwd: null pointer");
  if (R != 4 && R != 8 && R != 16) return set_error(-2, "dp_lora_bwd: rank %d not in {4,8,16}", R);
  return cuda_error(launch_lora_bwd(g, y, u_saved, B, lambda1, dA, dB, gu_ws, rows, D, R, scaling, p_drop, seed, sm_count(), ST),
                    "dp_lora_bwd");
}
extern "C" int dp_attention_fwd(const void* qkv, void* ctx, int B, int T, int heads, float scale, void* stream) {
  if (!qkv || !ctx) return set_error(-1, "dp_attention_fwd: null pointer");
  if (B <= 0 || T <= 0 || heads <= 0) return set_error(-2, "dp_attention_fwd: bad shape");
  return cuda_error(launch_attention_fwd(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(ctx), B, T,
                                         heads, scale, ST),
                    "dp_attention_fwd");
}
extern "C" int dp_decode(const float* hm, int maps, int H, int W, double tw, double th, int* idx, double* xy, float* conf,
                         void* stream) {
  if (!hm || !idx || !xy) return set_error(-1, "dp_decode: null pointer");
  if (maps <= 0) return 0;
  if (H <= 0 || W <= 0) return set_error(-2, "dp_decode: bad map size");
  return cuda_error(launch_decode(hm, maps, H, W, tw, th, idx, xy, conf, ST), "dp_decode");
}
extern "C" int dp_im2col(const void* in, void* col, int NB, int IH, int IW, int C, int OH, int OW, int KH, int KW,
                         int stride, int pad, void* stream) {
  if (!in || !col || C % 8) return set_error(-1, "dp_im2col: bad args");
  return cuda_error(launch_im2col(in, col, NB, IH, IW, C, OH, OW, KH, KW, stride, pad, ST), "dp_im2col");
}
extern "C" int dp_col2im(const void* col, const float* bias, void* big, int big_f32, int NB, int SH, int SW, int C, int BH,
                         int BW, int KH, int KW, int stride, int pad, void* stream) {
  if (!col || !big || C % 8) return set_error(-1, "dp_col2im: bad args");
  return cuda_error(launch_col2im(col, bias, big, big_f32, NB, SH, SW, C, BH, BW, KH, KW, stride, pad, ST), "dp_col2im");
}
extern "C" int dp_dwconv3x3(const void* in, const float* w, const float* bias, const void* add, void* out, int out_f32,
                            int NB, int H, int W, int C, int flip, void* stream) {
  if (!in || !w || !out || C % 8) return set_error(-1, "dp_dwconv3x3: bad args");
  return cuda_error(launch_dwconv3x3(in, w, bias, add, out, out_f32, NB, H, W, C, flip, ST), "dp_dwconv3x3");
}
extern "C" int dp_dwconv3x3_wgrad(const void* in, const void* dout, float* dw, int NB, int H, int W, int C, void* stream) {
  if (!in || !dout || !dw) return set_error(-1, "dp_dwconv3x3_wgrad: bad args");
  return cuda_error(launch_dwconv3x3_wgrad(in, dout, dw, NB, H, W, C, ST), "dp_dwconv3x3_wgrad");
}
extern "C" int dp_bn_stats(const void* raw, int raw_f32, double* sums, long long P, int C, void* stream) {
  if (!raw || !sums || C % 8) return set_error(-1, "dp_bn_stats: bad args");
  return cuda_error(launch_bn_stats(raw, raw_f32, sums, P, C, ST), "dp_bn_stats");
}
extern "C" int dp_bn_finalize(double* sums, const float* gamma, const float* beta, float* rm, float* rv, float* scale,
                              float* shift, float* mean, float* invstd, int C, double count, float eps, float momentum,
                              void* stream) {
  if (!sums || !gamma || !beta || !scale || !shift || !mean || !invstd) return set_error(-1, "dp_bn_finalize: bad args");
  return cuda_error(launch_bn_finalize(sums, gamma, beta, rm, rv, scale, shift, mean, invstd, C, count, eps, momentum, ST),
                    "dp_bn_finalize");
}
extern "C" int dp_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                               const float* conv_bias, float* scale, float* shift, float* mean_out, float* invstd_out,
                               int C, float eps, void* stream) {
  if (!gamma || !beta || !rm || !rv || !scale || !shift) return set_error(-1, "dp_bn_fold_eval: bad args");
  return cuda_error(launch_bn_fold_eval(gamma, beta, rm, rv, conv_bias, scale, shift, mean_out, invstd_out, C, eps, ST),
                    "dp_bn_fold_eval");
}
extern "C" int dp_bn_apply(const void* raw, int raw_f32, const float* scale, const float* shift, const void* add1,
                           const void* add2, void* out, long long P, int C, int relu, int mode, void* stream) {
  if (!raw || !scale || !shift || !out || C % 8) return set_error(-1, "dp_bn_apply: bad args");
  if (mode == 1 && !add1) return set_error(-2, "dp_bn_apply: mode 1 needs add1");
  return cuda_error(launch_bn_apply(raw, raw_f32, scale, shift, add1, add2, out, P, C, relu, mode, ST), "dp_bn_apply");
}
extern "C" long long dp_preprocess_workspace_bytes(int B, int H, int W, int short_edge, int crop) {
  return preprocess_workspace_bytes(B, H, W, short_edge, crop);
}
extern "C" int dp_preprocess_u8(const void* images, int B, int H, int W, int short_edge, int crop, const float* mean255,
                                const float* std255, float* out, void* workspace, long long workspace_bytes, void* stream) {
  if (!images || !mean255 || !std255 || !out || !workspace) return set_error(-1, "dp_preprocess_u8: null pointer");
  int status = 0;
  cudaError_t e = launch_preprocess(images, B, H, W, short_edge, crop, mean255, std255, out, workspace, workspace_bytes, &status, ST);
  if (status == -1) return set_error(-2, "dp_preprocess_u8: bad geometry B=%d H=%d W=%d short_edge=%d crop=%d (crop <= short_edge, crop <= 1024)", B, H, W, short_edge, crop);
  if (status == -2) return set_error(-3, "dp_preprocess_u8: down-scaling factor too large (more than 160 filter taps)");
  if (status == -3) return set_error(-4, "dp_preprocess_u8: workspace too small (%lld bytes given)", workspace_bytes);
  return cuda_error(e, "dp_preprocess_u8");
}
extern "C" int dp_bn_finalize_apply(const void* raw, int raw_f32, double* sums, const float* gamma, const float* beta, float* rm,
                                    float* rv, float* scale, float* shift, float* mean, float* invstd, const void* add1,
                                    const void* add2, void* out, long long P, int C, int relu, int mode, float eps,
                                    float momentum, void* stream) {
  if (!raw || !sums || !gamma || !beta || !scale || !shift || !mean || !invstd || !out || C % 8 || C > 512)
    return set_error(-1, "dp_bn_finalize_apply: bad args");
  if (mode == 1 && !add1) return set_error(-2, "dp_bn_finalize_apply: mode 1 needs add1");
  return cuda_error(launch_bn_finalize_apply(raw, raw_f32, sums, gamma, beta, rm, rv, scale, shift, mean, invstd, add1, add2, out,
                                             P, C, relu, mode, eps, momentum, ST),
                    "dp_bn_finalize_apply");
}
extern "C" int dp_bn_bwd_reduce(const void* dout, const void* raw, int raw_f32, const void* add1, const float* scale,
                                const float* shift, const float* mean, const float* invstd, double* sums, long long P,
                                int C, int relu, int mode, void* stream) {
  if (!dout || !raw || !scale || !shift || !mean || !invstd || !sums || C % 2) return set_error(-1, "dp_bn_bwd_reduce: bad args");
  return cuda_error(launch_bn_bwd_reduce(dout, raw, raw_f32, add1, scale, shift, mean, invstd, sums, P, C, relu, mode, ST),
                    "dp_bn_bwd_reduce");
}
extern "C" int dp_bn_bwd_apply(const void* dout, const void* raw, int raw_f32, const void* add1, const float* gamma,
                               const float* scale,
                               const float* shift, const float* mean, const float* invstd, double* sums, void* draw,
                               void* dres, float* dgamma, float* dbeta, long long P, int C, int relu, int mode,
                               int eval_mode, int shuffle_oh, int shuffle_ow, void* stream) {
  if (!dout || !raw || !scale || !shift || !draw || !sums || C % 8 || (!eval_mode && !gamma) || (eval_mode != 1 && (!mean || !invstd)))
    return set_error(-1, "dp_bn_bwd_apply: bad args");
  return cuda_error(launch_bn_bwd_apply(dout, raw, raw_f32, add1, gamma, scale, shift, mean, invstd, sums, draw, dres, dgamma, dbeta,
                                        P, C, relu, mode, eval_mode, shuffle_oh, shuffle_ow, ST),
                    "dp_bn_bwd_apply");
}
extern "C" int dp_avgpool2(const float* in, float* out, long long planes, int OH, int OW, void* stream) {
  if (!in || !out) return set_error(-1, "dp_avgpool2: bad args");
  return cuda_error(launch_avgpool2(in, out, planes, OH, OW, ST), "dp_avgpool2");
}
extern "C" int dp_pred1x1_fwd(const void* a, const float* w, const float* bias, float* out, long long P, int HW, int C, int K,
                              void* stream) {
  if (!a || !w || !bias || !out) return set_error(-1, "dp_pred1x1_fwd: null pointer");
  if (C != 64 || K < 1 || K > 32 || HW <= 0 || P <= 0 || P % HW) return set_error(-2, "dp_pred1x1_fwd: needs C = 64, K <= 32, P %% HW == 0 (C=%d K=%d)", C, K);
  return cuda_error(launch_pred1x1_fwd(static_cast<const __nv_bfloat16*>(a), w, bias, out, P, HW, K, ST), "dp_pred1x1_fwd");
}
extern "C" int dp_pred1x1_bwd(const float* g, const void* a, const float* w, void* d, float* dW, float* db, long long P, int HW,
                              int C, int K, void* stream) {
  if (!g || !a || !w || !d || !dW || !db) return set_error(-1, "dp_pred1x1_bwd: null pointer");
  if (C != 64 || K < 1 || K > 32 || HW <= 0 || P <= 0 || P % HW) return set_error(-2, "dp_pred1x1_bwd: needs C = 64, K <= 32, P %% HW == 0 (C=%d K=%d)", C, K);
  return cuda_error(launch_pred1x1_bwd(g, static_cast<const __nv_bfloat16*>(a), w, static_cast<__nv_bfloat16*>(d), dW, db, P, HW,
                                       K, sm_count(), ST), "dp_pred1x1_bwd");
}
extern "C" int dp_hm_grad_to_nhwc(const float* g, void* out, int NB, int K, int Kp, int OH, int OW, int up, void* stream) {
  if (!g || !out || (up != 1 && up != 2)) return set_error(-1, "dp_hm_grad_to_nhwc: bad args");
  return cuda_error(launch_hm_grad_to_nhwc(g, out, NB, K, Kp, OH, OW, up, ST), "dp_hm_grad_to_nhwc");
}
extern "C" int dp_mean_tokens(const void* feat, float* out, int B, int N, int D, void* stream) {
  if (!feat || !out) return set_error(-1, "dp_mean_tokens: bad args");
  return cuda_error(launch_mean_tokens(feat, out, B, N, D, ST), "dp_mean_tokens");
}
extern "C" int dp_mean_tokens_bwd(void* dfeat, const float* dmean, int B, int N, int D, void* stream) {
  if (!dfeat || !dmean) return set_error(-1, "dp_mean_tokens_bwd: bad args");
  return cuda_error(launch_mean_tokens_bwd(dfeat, dmean, B, N, D, ST), "dp_mean_tokens_bwd");
}
extern "C" int dp_sgemm_small(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k,
                              long long sb_n, float* C, long long ldc, int M, int N, int K, const float* bias, int relu,
                              const float* mask_ref, long long ld_ref, float p_drop, const unsigned long long* seed,
                              int accumulate, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return set_error(-1, "dp_sgemm_small: bad args");
  return cuda_error(launch_sgemm_small(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, M, N, K, bias, relu, mask_ref, ld_ref, p_drop,
                                       seed, accumulate, ST),
                    "dp_sgemm_small");
}
extern "C" int dp_attention_bwd(const void* qkv, const void* ctx, const void* dctx, void* dqkv, float* stats, int B, int T,
                                int heads, float scale, void* stream) {
  if (!qkv || !ctx || !dctx || !dqkv || !stats) return set_error(-1, "dp_attention_bwd: null pointer");
  if (B <= 0 || T <= 0 || heads <= 0) return set_error(-2, "dp_attention_bwd: bad shape");
  return cuda_error(launch_attention_bwd(static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(ctx),
                                         static_cast<const __nv_bfloat16*>(dctx), static_cast<__nv_bfloat16*>(dqkv), stats, B,
                                         T, heads, scale, ST),
                    "dp_attention_bwd");
}
extern "C" int dp_layernorm_bwd_params(const void* dy, int dy_is_bf16, const float* x, float* dgamma, float* dbeta,
                                       long long rows, int D, float eps, void* stream) {
  if (!dy || !x || !dgamma || !dbeta) return set_error(-1, "dp_layernorm_bwd_params: null pointer");
  if (!ln_dim_ok(D)) return set_error(-2, "dp_layernorm_bwd_params: unsupported D=%d", D);
  return cuda_error(launch_layernorm_bwd_params(dy, dy_is_bf16, x, dgamma, dbeta, rows, D, eps, sm_count(), ST),
                    "dp_layernorm_bwd_params");
}
extern "C" int dp_colsum_prod(const float* g, const void* a, float* out, long long P, int C, void* stream) {
  if (!g || !a || !out || P <= 0 || C <= 0) return set_error(-1, "dp_colsum_prod: bad args");
  return cuda_error(launch_colsum_prod(g, static_cast<const __nv_bfloat16*>(a), out, P, C, ST), "dp_colsum_prod");
}
extern "C" int dp_colsum(const void* x, int is_bf16, float* out, long long P, int C, long long ld, void* stream) {
  if (!x || !out) return set_error(-1, "dp_colsum: bad args");
  return cuda_error(launch_colsum(x, is_bf16, out, P, C, ld, ST), "dp_colsum");
}
extern "C" int dp_relu_mask(const float* d, const float* ref, float* out, long long n, float keep_scale, void* stream) {
  if (!d || !ref || !out) return set_error(-1, "dp_relu_mask: bad args");
  return cuda_error(launch_relu_mask(d, ref, out, n, keep_scale, ST), "dp_relu_mask");
}

namespace dp {
cudaError_t launch_pose_loss(const float*, const float*, const float*, int, const float*, const float*, double*, float*, float*,
                             float*, float*, float*, int, int, int, float, float, int, cudaStream_t);
cudaError_t launch_adamw(float*, const float*, float*, float*, long long, float, float, float, float, float, float, long long*,
                         const float*, int, int, cudaStream_t);
}  // namespace dp
extern "C" int dp_pose_loss(const float* heatmaps, const float* target_heatmaps, const float* keypoints, int kp_stride,
                            const float* z, const float* target_z, double* sums, float* state, float* out, float* scales,
                            float* d_heatmaps, float* d_z, int B, int K, int HW, float momentum, float rate, void* stream) {
  if (!heatmaps || !target_heatmaps || !keypoints || !z || !target_z || !sums || !state || !out || !scales || !d_heatmaps || !d_z)
    return set_error(-1, "dp_pose_loss: null pointer");
  if (B <= 0 || K <= 0 || HW <= 0 || (HW % 4) || kp_stride < 3) return set_error(-2, "dp_pose_loss: bad shape");
  return cuda_error(launch_pose_loss(heatmaps, target_heatmaps, keypoints, kp_stride, z, target_z, sums, state, out, scales,
                                     d_heatmaps, d_z, B, K, HW, momentum, rate, sm_count(), ST), "dp_pose_loss");
}
extern "C" int dp_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, float grad_scale, long long* step_dev, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !step_dev) return set_error(-1, "dp_adamw: null pointer");
  if (n <= 0 || (n % 4)) return set_error(-2, "dp_adamw: n must be a positive multiple of 4");
  return cuda_error(launch_adamw(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, grad_scale,
                                 step_dev, nullptr, 1, sm_count(), ST), "dp_adamw");
}
extern "C" int dp_adamw_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                            const float* hyper_dev, float beta1, float beta2, float eps, float grad_scale, long long* step_dev,
                            int bump_step, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !step_dev || !hyper_dev) return set_error(-1, "dp_adamw_dev: null pointer");
  if (n <= 0 || (n % 4)) return set_error(-2, "dp_adamw_dev: n must be a positive multiple of 4");
  return cuda_error(launch_adamw(params, grads, exp_avg, exp_avg_sq, n, 0.f, beta1, beta2, eps, 0.f, grad_scale, step_dev,
                                 hyper_dev, bump_step, sm_count(), ST), "dp_adamw_dev");
}

namespace dp {
cudaError_t launch_pack_weights(const long long*, int, long long, int, cudaStream_t);
cudaError_t launch_add_i64(const long long*, int, long long, cudaStream_t);
}  // namespace dp
extern "C" int dp_pack_weights_bf16(const long long* jobs_dev, int njobs, long long max_total, void* stream) {
  if (!jobs_dev || njobs <= 0 || njobs > 65535 || max_total <= 0) return set_error(-1, "dp_pack_weights_bf16: bad args");
  return cuda_error(launch_pack_weights(jobs_dev, njobs, max_total, sm_count(), ST), "dp_pack_weights_bf16");
}
extern "C" int dp_add_i64(const long long* ptrs_dev, int n, long long inc, void* stream) {
  if (!ptrs_dev || n <= 0) return set_error(-1, "dp_add_i64: bad args");
  return cuda_error(launch_add_i64(ptrs_dev, n, inc, ST), "dp_add_i64");
}
