/* libdinopose_sm100a.so -- C ABI of the B200-native dino_pose hot path.
 *
 * The reference (seungjoohan/dino_pose) is pure Python/PyTorch: it has no FFI of its own, its
 * "operator API" is the nn.Module surface of model/dinov2_pose.py, model/pose_heads.py,
 * model/lora.py and the numpy decode in src/model_utils.py.  This header is
 * the level BELOW that surface: every ATen / numpy call the reference makes on the hot path is
 * replaced by one of these entry points; dino_pose_b200/model/*.py (same class names, constructor
 * arguments, forward() signatures and state_dict keys as the reference) binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host; no torch types cross the ABI
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on that stream,
 *     allocate nothing and never synchronise (CUDA-graph capturable)
 *   - return value: 0 = ok, negative = invalid argument (dp_last_error() has the text),
 *     positive = cudaError_t from the launch
 *   - activations are bf16 NHWC / [rows, channels] row-major; the residual stream, LayerNorm
 *     statistics, accumulators, gradients of parameters and the heat-map / z outputs are fp32
 */
#ifndef DINOPOSE_H_
#define DINOPOSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_ABI_VERSION 7

const char* dp_last_error(void);
int dp_abi_version(void);
/* sizeof() of the argument structs, so the ctypes mirror can be checked at load time */
int dp_sizeof_gemm_args(void);
int dp_sizeof_wgrad_args(void);

/* ---- activation / row-map / dtype codes (dp_gemm_args) ---- */
enum { DP_ACT_NONE = 0, DP_ACT_RELU = 1, DP_ACT_GELU = 2 };
enum { DP_ROWMAP_IDENTITY = 0, DP_ROWMAP_PATCH_TOKENS = 1, DP_ROWMAP_NCHW = 2, DP_ROWMAP_SHUFFLE2X2 = 3 };
enum { DP_OUT_BF16 = 0, DP_OUT_F32 = 1 };

/* C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue), bf16 x bf16 -> fp32 on tcgen05.
 * Replaces: nn.Linear q/k/v/dense/fc1/fc2 (HF modeling_dinov2.py:211-213,250,325-327), the patch
 * conv (HF:139,148), and -- with a_mode = 1 (implicit convolution over an NHWC activation) -- the
 * stride-1 nn.Conv2d / ConvTranspose2d(k4,s1) of model/pose_heads.py:306-340.
 * Epilogue (per element): v = acc*scale[c] + bias[c]; aux_out = bf16(act == GELU ? gelu'(v) : v); v = act(v);
 *   v *= aux_in; v *= ls[c]; v += residual[row,c]; out[row_map(row), c] = v
 * (aux_out of the fc1 forward is therefore exactly the multiplier aux_in of the fc2 input-gradient GEMM)            */
typedef struct {
  const void* A;          /* bf16 */
  const void* W;          /* bf16 [N, K] row-major, row pitch ldw */
  long long lda, ldw;     /* elements */
  int M, N, K;
  int block_n;            /* 32 / 64 / 128 / 192 / 256, 0 = auto */
  int a_mode;             /* 0: A is [M,K]; 1: A is NHWC [NB,IH,IW,C], K = KH*KW*C, M = NB*OH*OW */
  int C, IW, IH, NB;
  long long a_stride_w, a_stride_h, a_stride_b; /* elements */
  int KH, KW, pad_y, pad_x, OH, OW;
  void* out;
  long long ldo;
  int out_dtype;
  int act;
  const float* bias;
  const float* scale;
  const float* ls;
  const void* residual;   /* fp32 (or bf16 when res_is_bf16) [rows, ldr] */
  long long ldr;
  int res_is_bf16;
  void* aux_out;
  const void* aux_in;
  long long ld_aux;
  int row_map, n_valid, map_a, map_b;
  double* stats;          /* optional fused train-mode BatchNorm statistics of v (before act): fp64 [2*stats_c],
                             stats[c] += sum over rows, stats[stats_c + c] += sum of squares (same contract as
                             dp_bn_stats: zero on entry, dp_bn_finalize re-zeroes).  With DP_ROWMAP_SHUFFLE2X2
                             the channel of column j is j % map_a. */
  int stats_c;
  int cta_pair;           /* 0 auto, 1 force the CTA-pair (cta_group::2, 256-row tile) kernel, 2 force single-CTA,
                             3 / 4 the A-stationary kernels (K <= 512; 4 = clusters of two with multicast weight loads) */
  /* Optional fused LayerNorm of the OUTPUT rows (the nn.LayerNorm that follows the attention / MLP projection in the next
   * sub-block, HF:371,379): ln_out = bf16 LayerNorm(out[row, :]; ln_gamma, ln_beta, ln_eps), row pitch ld_ln.  Needs
   * out_dtype fp32, identity row map, a_mode 0 and N in {128, 256, 384} (one CTA owns complete rows). */
  const float* ln_gamma;
  const float* ln_beta;
  void* ln_out;
  long long ld_ln;
  float ln_eps;
  /* Optional fused LoRA adapter on the projection output (reference model/lora.py:26-28,53-59 + LayerScale + residual,
   * HF:373-376), same row-owning kernel:  y = A W^T + bias;  out = residual + ls * (y + lora_scaling * dropout(y lora_A lora_B)).
   * lora_A fp32 [N, 8], lora_B fp32 [8, N] (rank 8); the dropout mask is the counter-based hash of dp_lora_bwd (key =
   * *lora_seed, element row * N + column, probability lora_p_drop).  lora_y_out fp32 [rows, ld_lora_y] and lora_u_out fp32
   * [rows, 8] receive y and u = y lora_A, the activations dp_lora_bwd reads (either may be NULL).  Same shape limits as the
   * fused LayerNorm; the two cannot be combined. */
  const float* lora_A;
  const float* lora_B;
  float* lora_y_out;
  long long ld_lora_y;
  float* lora_u_out;
  const unsigned long long* lora_seed;
  float lora_scaling, lora_p_drop;
  int lora_rank;
} dp_gemm_args;
int dp_gemm_bf16(const dp_gemm_args* a, void* stream);

/* Weight gradient: out[off(m)+off(n)+tap*so_t] (+)= sum_p A[p,m] * B[p (+tap), n]  (fp32 atomics onto a
 * caller-zeroed `out`, or -- with a workspace -- plain stores of split-K partials + a reduce kernel).  Replaces autograd's conv / linear weight-gradient kernels for
 * model/pose_heads.py layers (train.py:169 loss.backward()).
 * mode 0: A [P,Mc] (lda), B [P,Nc] (ldb).  mode 1: A NHWC [NB,OH,OW,Mc], B NHWC [NB,IH,IW,Nc],
 * B read at (y+ky-pad_y, x+kx-pad_x) for tap (ky,kx).                                           */
typedef struct {
  const void* A;
  const void* B;
  int mode, P;
  long long lda, ldb;
  int Mc, Nc;
  int NB, OH, OW, IH, IW;
  long long a_sw, a_sh, a_sb, b_sw, b_sh, b_sb;
  int KH, KW, pad_y, pad_x;
  float* out;
  long long so_m, so_mo, so_n, so_no, so_t;
  int m_inner, n_inner;
  int block_n, splits;
  void* workspace;            /* optional split-K workspace (fp32, 16-byte aligned): when it holds at least
                                 taps * ceil(Mc/128)*128 * ceil(Nc/block_n)*block_n * 4 bytes per split, the partial
                                 tiles are stored there without atomics and a second kernel reduces them into `out`
                                 (which then needs no zeroing and the result is deterministic) */
  long long workspace_bytes;
} dp_wgrad_args;
int dp_wgrad_bf16(const dp_wgrad_args* a, void* stream);
/* Tuning aid, not part of the reference surface: with DP_GEMM_TRACE=2 in the environment CTA 0 of every dp_gemm_bf16 launch
 * records a cycle / nanosecond timeline in a device buffer; this copies its first n (<= 4096) int64 entries to the host. */
int dp_debug_read_trace(long long* host, int n);
/* Persistent GEMM / weight-gradient grids launched after this call leave k SMs free (0 = use all); returns the previous
 * value.  Host-side state read at launch time (under stream capture: at capture time).  Used by the data-parallel trainer
 * while NCCL all-reduces run beside the backward (no reference counterpart: the reference is single-device). */
int dp_set_reserved_sms(int k);

/* ---------------------------------------------------------------- backbone row-wise kernels */

/* nn.LayerNorm(D, eps) forward (HF modeling_dinov2.py:371,379,477).  x fp32 [rows,D] -> y bf16 and/or
 * y32 fp32 (either may be NULL).  drop_cls != 0: rows are [B,T] tokens; the CLS row of each image is
 * dropped and the output is [B*(T-1), D] (reference model/dinov2_pose.py:147,153).  D in {128,384,768,1024}. */
int dp_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                     long long rows, int D, int T, int drop_cls, float eps, void* stream);
/* LayerNorm input-gradient (autograd of the above; parameters frozen).  dx = LNbwd(dy) (+ add_in);
 * optional dx_scaled_bf16 = bf16(dx * ls) feeds the next dgrad GEMM. */
int dp_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* add_in,
                     float* dx, const float* ls, void* dx_scaled_bf16, long long rows, int D, int T, int drop_cls,
                     float eps, void* stream);
/* im2col of the 14x14/stride-14 patch conv (HF:139-149): pixel_values fp32 NCHW -> bf16 [B*gh*gw, Kp]. */
int dp_patch_im2col(const float* pixel_values, void* out_bf16, int B, int H, int W, int Kp, void* stream);
/* x[b*T + 0, :] = cls_row (cls_token + position_embeddings[0], HF:108-112). */
int dp_fill_cls(float* x, const float* cls_row, int B, int T, int D, void* stream);
/* LoRALayer on the attention output + LayerScale + residual (reference model/lora.py:26-28,57-59; HF:373-376):
 * x_out = x_in + (y + dropout(y A B) * scaling) * lambda1.  u_save [rows, R] optional. */
int dp_lora_fwd(const float* y, const float* A, const float* B, const float* lambda1, const float* x_in,
                float* x_out, float* u_save, long long rows, int D, int R, float scaling, float p_drop,
                const unsigned long long* seed_dev, void* stream);
/* gradients of lora_A [D,R] and lora_B [R,D] (accumulated with atomics; caller zeroes them). */
int dp_lora_bwd(const float* g, const float* y, const float* u_saved, const float* B, const float* lambda1,
                float* dA, float* dB, float* gu_workspace /* [rows,R] */, long long rows, int D, int R, float scaling, float p_drop,
                const unsigned long long* seed_dev, void* stream);
/* Multi-head attention forward (HF:203-234), head dim 64.  qkv bf16 [B*T, 3*heads*64] -> ctx bf16 [B*T, heads*64]. */
int dp_attention_fwd(const void* qkv_bf16, void* ctx_bf16, int B, int T, int heads, float scale, void* stream);

/* ---- backward of an UN-FROZEN encoder layer: Dinov2PoseModel(unfreeze_last_n_layers = n), reference
 * model/dinov2_pose.py:25-39 (autograd of HF:203-234, :371-384) */
/* Attention backward.  qkv, dqkv bf16 [B*T, 3*heads*64] (q | k | v and dq | dk | dv); ctx = forward output, dctx =
 * its gradient, bf16 [B*T, heads*64]; stats = fp32 scratch [2 * B*heads*T] (row log-sum-exp and rowsum(dO o O),
 * recomputed here: the forward stores nothing).  Deterministic (no atomics). */
int dp_attention_bwd(const void* qkv_bf16, const void* ctx_bf16, const void* dctx_bf16, void* dqkv_bf16, float* stats,
                     int B, int T, int heads, float scale, void* stream);
/* LayerNorm parameter gradients: dgamma[c] += sum_rows dy*xhat, dbeta[c] += sum_rows dy (atomics, caller zeroes). */
int dp_layernorm_bwd_params(const void* dy, int dy_is_bf16, const float* x, float* dgamma, float* dbeta, long long rows,
                            int D, float eps, void* stream);
/* LayerScale gradient (HF:272-278): out[c] += sum_rows g[row,c] * a[row,c]; g fp32, a bf16, both [P, C] dense. */
int dp_colsum_prod(const float* g, const void* a_bf16, float* out, long long P, int C, void* stream);

/* ---------------------------------------------------------------- decode */
/* Heat-map -> key-points (reference src/model_utils.py:10-51).  heatmaps fp32 [maps, H, W];
 * idx int32 [maps,2] = (row, col) of the first maximum; xy float64 [maps,2] = refined (x, y) scaled to
 * (target_w, target_h); conf fp32 [maps] = peak value (may be NULL).  Bit-exact vs the numpy reference. */
int dp_decode(const float* heatmaps, int maps, int H, int W, double target_w, double target_h, int* idx,
              double* xy, float* conf, void* stream);

/* ---------------------------------------------------------------- pose-head kernels (NHWC bf16) */
int dp_im2col(const void* in, void* col, int NB, int IH, int IW, int C, int OH, int OW, int KH, int KW, int stride,
              int pad, void* stream);
int dp_col2im(const void* col, const float* bias, void* big, int big_is_f32, int NB, int SH, int SW, int C, int BH, int BW,
              int KH, int KW, int stride, int pad, void* stream);
int dp_dwconv3x3(const void* in, const float* w, const float* bias, const void* add, void* out, int out_is_f32, int NB,
                 int H, int W, int C, int flip, void* stream);
int dp_dwconv3x3_wgrad(const void* in, const void* dout, float* dw, int NB, int H, int W, int C, void* stream);
/* train-mode BatchNorm2d (eps, momentum as torch): sums fp64 [2*C] must be zero on entry of dp_bn_stats;
 * dp_bn_finalize re-zeroes it.  `sums` must be the [DP_BN_BWD_REPLICAS + 1][2*C] buffer described at dp_bn_bwd_reduce
 * (all zero on first use): dp_bn_stats spreads its atomics over the replicas and folds them into [0, 2*C) itself.  `raw` (pre-BN conv output, [P,C]) is bf16 or fp32 (raw_is_f32). */
int dp_bn_stats(const void* raw, int raw_is_f32, double* sums, long long P, int C, void* stream);
int dp_bn_finalize(double* sums, const float* gamma, const float* beta, float* running_mean, float* running_var,
                   float* scale, float* shift, float* mean, float* invstd, int C, double count, float eps,
                   float momentum, void* stream);
/* mean_out / invstd_out (nullable): the running mean and 1/sqrt(running_var + eps), in the layout the backward entry
 * points read saved batch statistics in -- used by a training step whose heads are in eval mode (frozen statistics). */
int dp_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                    const float* conv_bias, float* scale, float* shift, float* mean_out, float* invstd_out, int C,
                    float eps, void* stream);
int dp_bn_apply(const void* raw, int raw_is_f32, const float* scale, const float* shift, const void* add1,
                const void* add2, void* out, long long P, int C, int relu, int mode, void* stream);
/* dp_bn_finalize + dp_bn_apply in ONE launch (pose_heads.py train-mode nn.BatchNorm2d + ReLU / adds): every thread
 * block derives scale / shift from `sums` itself, block 0 writes scale / shift / mean / invstd (kept for the backward)
 * and updates the running statistics, the last block to finish re-zeroes sums[0 .. 2*C).  count = P.  C <= 512.
 * `sums` must be the [DP_BN_BWD_REPLICAS + 1][2*C] buffer described below (its last 4 bytes hold a ticket counter,
 * zero on first use). */
int dp_bn_finalize_apply(const void* raw, int raw_is_f32, double* sums, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                         float* invstd, const void* add1, const void* add2, void* out, long long P, int C, int relu,
                         int mode, float eps, float momentum, void* stream);
/* BatchNorm backward.  `sums` here is fp64 [DP_BN_BWD_REPLICAS + 1][2*C]: the first DP_BN_BWD_REPLICAS blocks are
 * accumulators, zero on entry of dp_bn_bwd_reduce (the thread blocks spread their atomics over the replicas);
 * dp_bn_bwd_apply first adds the replicas up, forms the per-channel coefficients (scratch = the last block), writes
 * dgamma / dbeta and RE-ZEROES the accumulators, then applies draw = A*dy + B*raw + K.
 * eval_mode 0: train-mode BatchNorm (batch statistics in mean / invstd).  1: identity pass, draw = dy * scale, dgamma = 0.
 * 2: BatchNorm with frozen statistics (`module.eval()` inside a training step, torch nn.BatchNorm2d eval semantics):
 *    mean / invstd hold the running statistics (dp_bn_fold_eval), draw = dy * scale, dgamma = sum dy * xhat. */
#define DP_BN_BWD_REPLICAS 8
int dp_bn_bwd_reduce(const void* dout, const void* raw, int raw_is_f32, const void* add1, const float* scale, const float* shift,
                     const float* mean, const float* invstd, double* sums, long long P, int C, int relu, int mode,
                     void* stream);
int dp_bn_bwd_apply(const void* dout, const void* raw, int raw_is_f32, const void* add1, const float* gamma, const float* scale,
                    const float* shift, const float* mean, const float* invstd, double* sums, void* draw,
                    void* dres, float* dgamma, float* dbeta, long long P, int C, int relu, int mode, int eval_mode,
                    int shuffle_oh, int shuffle_ow, void* stream);
int dp_avgpool2(const float* in, float* out, long long planes, int OH, int OW, void* stream);
/* prediction.3 = Conv2d(64, K, 1) + bias (reference model/pose_heads.py:335-340) on CUDA cores, K <= 32, fp32 weights read
 * from the parameter: a bf16 [P, 64] (NHWC rows, P = NB*HW) -> out fp32 NCHW [NB, K, HW]. */
int dp_pred1x1_fwd(const void* a_bf16, const float* w /* [K,64] */, const float* bias, float* out_nchw, long long P, int HW,
                   int C, int K, void* stream);
/* its backward in one launch: g fp32 NCHW [NB, K, HW] -> d bf16 [P, 64] (input gradient); dW [K,64] and db [K] are
 * accumulated with atomics (caller zeroes them). */
int dp_pred1x1_bwd(const float* g_nchw, const void* a_bf16, const float* w, void* d_bf16, float* dW, float* db, long long P,
                   int HW, int C, int K, void* stream);
int dp_hm_grad_to_nhwc(const float* g, void* out_bf16, int NB, int K, int Kp, int OH, int OW, int up, void* stream);
int dp_mean_tokens(const void* feat_bf16, float* out, int B, int N, int D, void* stream);
int dp_mean_tokens_bwd(void* dfeat_bf16, const float* dmean, int B, int N, int D, void* stream);
/* fp32 GEMM for the z-head MLP (M = batch): C[m,n] (+)= sum_k A[m*sa_m + k*sa_k] * B[k*sb_k + n*sb_n]. */
int dp_sgemm_small(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k, long long sb_n,
                   float* C, long long ldc, int M, int N, int K, const float* bias, int relu, const float* mask_ref,
                   long long ld_ref, float p_drop, const unsigned long long* seed_dev, int accumulate, void* stream);
int dp_colsum(const void* x, int is_bf16, float* out, long long P, int C, long long ld, void* stream);
/* out = ref > 0 ? d * keep_scale : 0  (gradient through ReLU + inverted dropout given the saved output) */
int dp_relu_mask(const float* d, const float* ref, float* out, long long n, float keep_scale, void* stream);

/* ---------------------------------------------------------------- training step around the model */
/* Reference losses + their backward seed (train.py:89-120, DynamicLossWeighting :17-69), state on the device.
 * heatmaps / target_heatmaps fp32 [B,K,HW]; keypoints fp32 [B*K, kp_stride] with the visibility at column 2
 * (mask = visibility > 1, train.py:94,114); z / target_z fp32 [B,K].
 * sums fp64[2] (zero on entry, re-zeroed); state fp32[4] = {kp_avg, z_avg, started, weight} (init {0,0,0,0.1});
 * out fp32[3] = {balanced loss, keypoint loss, z loss}; scales fp32[2] workspace;
 * d_heatmaps [B,K,HW], d_z [B,K] = d(balanced loss)/d(heatmaps), /d(z).  HW % 4 == 0. */
int dp_pose_loss(const float* heatmaps, const float* target_heatmaps, const float* keypoints, int kp_stride,
                 const float* z, const float* target_z, double* sums, float* state, float* out, float* scales,
                 float* d_heatmaps, float* d_z, int B, int K, int HW, float momentum, float rate, void* stream);
/* torch.optim.AdamW step (train.py:280-284,170) over flat fp32 buffers of n elements (n % 4 == 0, 16-byte
 * aligned); gradients are multiplied by grad_scale first (1/world_size after the all-reduce); step_dev is a
 * device int64 holding the number of steps taken so far and is incremented. */
int dp_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
             float beta2, float eps, float weight_decay, float grad_scale, long long* step_dev, void* stream);
/* Same step with the learning rate and the weight decay read from DEVICE memory, hyper_dev = fp32 {lr, weight_decay}:
 * the reference's ReduceLROnPlateau (train.py:286-293,341) changes the rate between steps, and a step captured in a
 * CUDA graph would otherwise keep the value it was recorded with.
 * bump_step = 0 leaves the step counter alone: the optimizer step may be issued as several launches over disjoint slices of
 * the flat buffers (the slice whose gradients are final first is updated while the backward of the rest still runs); the
 * last launch of a step passes 1. */
int dp_adamw_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, const float* hyper_dev,
                 float beta1, float beta2, float eps, float grad_scale, long long* step_dev, int bump_step, void* stream);

/* Re-pack trainable conv weights (fp32 [d0,d1,kh,kw] nn.Parameters, reference model/pose_heads.py) into the bf16 GEMM
 * layouts, all layers in one launch.  jobs_dev: device table, 16 int64 per job =
 * { src ptr, dst ptr, n0,n1,n2,n3 (extents in dst order), s0..s3 (signed src element strides: a negative stride +
 *   src offset mirrors the filter taps), t0..t3 (dst element strides), src element offset, element count }. */
int dp_pack_weights_bf16(const long long* jobs_dev, int njobs, long long max_total, void* stream);
/* *ptrs[i] += inc for n device int64 scalars (BatchNorm num_batches_tracked, torch bookkeeping). */
int dp_add_i64(const long long* ptrs_dev, int n, long long inc, void* stream);

/* ---- image pre-processing: replaces `self.image_processor(image, return_tensors="pt")`
 * (reference model/dinov2_pose.py:15,182; demo.py:80,171; benchmark_model.py:35,45; data_loader/data_loader.py:52) =
 * HF BitImageProcessor with the DINOv2 preprocessor config (transformers/image_processing_backends.py
 * TorchvisionBackend.resize / center_crop / rescale_and_normalize over ATen's uint8 anti-aliased bicubic kernel).
 * images: DEVICE uint8 [B, H, W, 3] RGB, interleaved, contiguous.  out: DEVICE fp32 [B, 3, crop, crop].
 * Shortest edge -> short_edge (long edge int(short_edge * long / short)), bicubic + antialias in the CPU kernel's
 * uint8 / int16-weight arithmetic (horizontal pass, uint8 rounding, vertical pass), center crop, then
 * (float(u8) - mean255[c]) / std255[c] in fp32 with IEEE division.  mean255 / std255: HOST arrays of 3 floats
 * (= float32(mean) * float32(1 / rescale_factor), as the reference fuses them).  Bit-identical to the reference.
 * workspace: DEVICE, >= dp_preprocess_workspace_bytes(...) (returns -1 for an unsupported geometry). */
long long dp_preprocess_workspace_bytes(int B, int H, int W, int short_edge, int crop);
int dp_preprocess_u8(const void* images, int B, int H, int W, int short_edge, int crop, const float* mean255,
                     const float* std255, float* out, void* workspace, long long workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DINOPOSE_H_ */
