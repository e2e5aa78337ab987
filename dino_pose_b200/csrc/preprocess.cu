// Image pre-processing on the GPU: uint8 HWC RGB -> resize (shortest edge, bicubic, anti-aliased) ->
// center crop -> (x - 255*mean) / (255*std) -> fp32 NCHW pixel_values.
//
// Replaces the reference's `self.image_processor(image, return_tensors="pt")` (model/dinov2_pose.py:15,182; called at
// demo.py:80,171, benchmark_model.py:35,45, data_loader/data_loader.py:52), i.e. HF BitImageProcessor with the DINOv2
// preprocessor config, whose arithmetic is transformers/image_processing_backends.py (TorchvisionBackend.resize /
// center_crop / rescale_and_normalize) on top of ATen's uint8 anti-aliased bicubic kernel
// (aten/src/ATen/native/cpu/UpSampleKernel.cpp).  The result is BIT-IDENTICAL to that path (oracle/preprocess_oracle.py
// states the algorithm; tests/golden/preprocess.npz holds outputs of the real processor):
//   * per axis, float64 filter weights exactly in ATen's operation order (no FMA contraction: __dmul_rn / __dadd_rn),
//     normalised, quantised to int16 with the axis-wide precision p (largest shift keeping round(max_w * 2^(p+1)) < 2^15);
//   * horizontal pass, result rounded and clamped to uint8, then the vertical pass (same integer arithmetic);
//   * float32 subtract / IEEE divide by the fused mean / std.
// Only what the crop needs is computed: the horizontal pass covers the input rows the crop's vertical taps touch and the
// crop's columns.  Three launches (weights of both axes, horizontal, vertical + normalise), no host synchronisation, no
// allocation: the caller passes the workspace (dp_preprocess_workspace_bytes).  All kernels are byte / HBM bound:
// algorithmic traffic = the touched input rows once + 4 * 3 * crop^2 output bytes.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "launch.cuh"

namespace dp {

constexpr int kPpMaxTaps = 160;     // ksize limit: down-scaling factors up to ~39
constexpr int kPpThreads = 256;

struct PpAxis {
  int in_size, out_size;   // full axis extents (the weights' precision depends on ALL output coordinates)
  int first, count;        // output coordinates that are kept (the crop)
  int ksize;
  double scale, support, invscale;
};

struct PpPlan {
  PpAxis ax[2];            // 0 = horizontal (x), 1 = vertical (y)
  int row0, rows;          // input rows the horizontal pass has to produce: [row0, row0 + rows)
  int H, W, crop, B;
};

__device__ __forceinline__ double pp_cubic(double x) {
  // HelperInterpCubic::aa_filter<double, true>: a = -0.5
  x = fabs(x);
  if (x < 1.0) {   // ((a + 2) * x - (a + 3)) * x * x + 1
    double t = __dmul_rn(1.5, x);
    t = __dsub_rn(t, 2.5);
    t = __dmul_rn(t, x);
    t = __dmul_rn(t, x);
    return __dadd_rn(t, 1.0);
  }
  if (x < 2.0) {   // ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    double t = __dmul_rn(-0.5, x);
    t = __dsub_rn(t, -2.5);
    t = __dmul_rn(t, x);
    t = __dadd_rn(t, -4.0);
    t = __dmul_rn(t, x);
    return __dsub_rn(t, -2.0);
  }
  return 0.0;
}

// _compute_indices_min_size_weights_aa for output coordinate i; returns sum of the raw weights
__device__ __forceinline__ double pp_span(const PpAxis& a, int i, int& xmin, int& xsize, double& center) {
  center = __dmul_rn(a.scale, double(i) + 0.5);
  long long lo = (long long)(__dadd_rn(__dsub_rn(center, a.support), 0.5));
  if (lo < 0) lo = 0;
  long long hi = (long long)(__dadd_rn(__dadd_rn(center, a.support), 0.5));
  if (hi > a.in_size) hi = a.in_size;
  long long n = hi - lo;
  if (n < 0) n = 0;
  if (n > a.ksize) n = a.ksize;
  xmin = int(lo);
  xsize = int(n);
  double total = 0.0;
  for (int j = 0; j < xsize; ++j)
    total = __dadd_rn(total, pp_cubic(__dmul_rn(__dadd_rn(__dsub_rn(double(j + xmin), center), 0.5), a.invscale)));
  return total;
}
__device__ __forceinline__ double pp_weight(const PpAxis& a, int j, int xmin, double center, double total) {
  const double w = pp_cubic(__dmul_rn(__dadd_rn(__dsub_rn(double(j + xmin), center), 0.5), a.invscale));
  return total != 0.0 ? __ddiv_rn(w, total) : w;
}

// one block per axis: (1) largest normalised weight over the WHOLE axis -> precision, (2) int16 weights of the kept range
__global__ void __launch_bounds__(kPpThreads) pp_weights_kernel(const PpPlan plan, short* __restrict__ w_all,
                                                                int* __restrict__ xmin_all, int* __restrict__ xsize_all,
                                                                int* __restrict__ prec_all) {
  pdl_grid_sync();
  const PpAxis a = plan.ax[blockIdx.x];
  short* w = w_all + (long long)blockIdx.x * plan.crop * kPpMaxTaps;
  int* xmin_o = xmin_all + blockIdx.x * plan.crop;
  int* xsize_o = xsize_all + blockIdx.x * plan.crop;
  __shared__ double s_max[kPpThreads];
  __shared__ int s_prec;
  double mx = 0.0;   // the weight table is zero padded, so the maximum is never below 0
  for (int i = threadIdx.x; i < a.out_size; i += kPpThreads) {
    int xmin, xsize;
    double center;
    const double total = pp_span(a, i, xmin, xsize, center);
    for (int j = 0; j < xsize; ++j) mx = fmax(mx, pp_weight(a, j, xmin, center, total));
  }
  s_max[threadIdx.x] = mx;
  __syncthreads();
  for (int s = kPpThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double max_w = s_max[0];
    int prec = 0;
    for (; prec < 22; ++prec) {
      const int next_value = int(__dadd_rn(0.5, __dmul_rn(max_w, double(1 << (prec + 1)))));
      if (next_value >= (1 << 15)) break;
    }
    s_prec = prec;
    prec_all[blockIdx.x] = prec;
  }
  __syncthreads();
  const double mult = double(1 << s_prec);
  for (int k = threadIdx.x; k < a.count; k += kPpThreads) {
    int xmin, xsize;
    double center;
    const double total = pp_span(a, a.first + k, xmin, xsize, center);
    xmin_o[k] = xmin;
    xsize_o[k] = xsize;
    for (int j = 0; j < xsize; ++j) {
      const double v = __dmul_rn(pp_weight(a, j, xmin, center, total), mult);
      w[(long long)k * kPpMaxTaps + j] = short(v < 0 ? int(__dadd_rn(-0.5, v)) : int(__dadd_rn(0.5, v)));
    }
  }
}

__device__ __forceinline__ int pp_round_clamp(int acc, int prec) {
  const int v = acc >> prec;   // arithmetic shift, as the int32 accumulation of the CPU kernel
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: tmp[b, r, x, c] for the needed input rows r and the crop's columns x   (grid: rows x B, block: crop)
__global__ void pp_horizontal_kernel(const PpPlan plan, const uint8_t* __restrict__ img, const short* __restrict__ w_all,
                                     const int* __restrict__ xmin_all, const int* __restrict__ xsize_all,
                                     const int* __restrict__ prec_all, uint8_t* __restrict__ tmp) {
  pdl_grid_sync();
  const int x = threadIdx.x, r = blockIdx.x, b = blockIdx.y;
  if (x >= plan.crop) return;
  const int prec = prec_all[0];
  const int xmin = xmin_all[x], n = xsize_all[x];
  const short* w = w_all + (long long)x * kPpMaxTaps;
  const uint8_t* src = img + ((long long)b * plan.H + plan.row0 + r) * plan.W * 3 + (long long)xmin * 3;
  int a0 = 1 << (prec - 1), a1 = a0, a2 = a0;
  for (int j = 0; j < n; ++j) {
    const int wj = w[j];
    a0 += int(src[3 * j]) * wj;
    a1 += int(src[3 * j + 1]) * wj;
    a2 += int(src[3 * j + 2]) * wj;
  }
  uint8_t* dst = tmp + (((long long)b * plan.rows + r) * plan.crop + x) * 3;
  dst[0] = uint8_t(pp_round_clamp(a0, prec));
  dst[1] = uint8_t(pp_round_clamp(a1, prec));
  dst[2] = uint8_t(pp_round_clamp(a2, prec));
}

// vertical pass + normalisation: out[b, c, y, x] fp32   (grid: crop rows x B, block: crop)
__global__ void pp_vertical_kernel(const PpPlan plan, const uint8_t* __restrict__ tmp, const short* __restrict__ w_all,
                                   const int* __restrict__ xmin_all, const int* __restrict__ xsize_all,
                                   const int* __restrict__ prec_all, float m0, float m1, float m2, float s0, float s1, float s2,
                                   float* __restrict__ out) {
  pdl_grid_sync();
  const int x = threadIdx.x, y = blockIdx.x, b = blockIdx.y;
  if (x >= plan.crop) return;
  const int prec = prec_all[1];
  const int ymin = xmin_all[plan.crop + y] - plan.row0, n = xsize_all[plan.crop + y];
  const short* w = w_all + ((long long)plan.crop + y) * kPpMaxTaps;
  const uint8_t* src = tmp + (((long long)b * plan.rows + ymin) * plan.crop + x) * 3;
  const long long pitch = (long long)plan.crop * 3;
  int a0 = 1 << (prec - 1), a1 = a0, a2 = a0;
  for (int j = 0; j < n; ++j) {
    const int wj = w[j];
    a0 += int(src[j * pitch]) * wj;
    a1 += int(src[j * pitch + 1]) * wj;
    a2 += int(src[j * pitch + 2]) * wj;
  }
  const long long plane = (long long)plan.crop * plan.crop;
  float* o = out + (long long)b * 3 * plane + (long long)y * plan.crop + x;
  o[0] = __fdiv_rn(__fsub_rn(float(pp_round_clamp(a0, prec)), m0), s0);
  o[plane] = __fdiv_rn(__fsub_rn(float(pp_round_clamp(a1, prec)), m1), s1);
  o[2 * plane] = __fdiv_rn(__fsub_rn(float(pp_round_clamp(a2, prec)), m2), s2);
}

// ---- host: geometry exactly as the Python side of the reference computes it
static void pp_axis(PpAxis& a, int in_size, int out_size, int first, int count) {
  a.in_size = in_size; a.out_size = out_size; a.first = first; a.count = count;
  a.scale = double(in_size) / double(out_size);                       // area_pixel_compute_scale<double>, no scale factor
  a.support = a.scale >= 1.0 ? 2.0 * a.scale : 2.0;                   // (interp_size * 0.5) * scale
  a.ksize = int(ceil(a.support)) * 2 + 1;
  a.invscale = a.scale >= 1.0 ? 1.0 / a.scale : 1.0;
}
static void pp_span_host(const PpAxis& a, int i, int& xmin, int& xsize) {
  const double center = a.scale * (double(i) + 0.5);
  long long lo = (long long)(center - a.support + 0.5);
  if (lo < 0) lo = 0;
  long long hi = (long long)(center + a.support + 0.5);
  if (hi > a.in_size) hi = a.in_size;
  long long n = hi - lo;
  if (n < 0) n = 0;
  if (n > a.ksize) n = a.ksize;
  xmin = int(lo); xsize = int(n);
}

// returns 0 and fills the plan, or a negative code
int pp_make_plan(PpPlan& p, int B, int H, int W, int short_edge, int crop) {
  if (B <= 0 || H <= 0 || W <= 0 || short_edge <= 0 || crop <= 0 || crop > short_edge || crop > 1024) return -1;
  // get_resize_output_image_size(size=short_edge, default_to_square=False): int(short_edge * long / short)
  const int shortv = W <= H ? W : H, longv = W <= H ? H : W;
  const int new_long = int(double((long long)short_edge * longv) / double(shortv));
  const int new_h = W <= H ? new_long : short_edge, new_w = W <= H ? short_edge : new_long;
  const int top = int(double(new_h - crop) / 2.0), left = int(double(new_w - crop) / 2.0);   // center_crop
  pp_axis(p.ax[0], W, new_w, left, crop);
  pp_axis(p.ax[1], H, new_h, top, crop);
  if (p.ax[0].ksize > kPpMaxTaps || p.ax[1].ksize > kPpMaxTaps) return -2;
  int lo, n, hi_lo, hi_n;
  pp_span_host(p.ax[1], top, lo, n);
  pp_span_host(p.ax[1], top + crop - 1, hi_lo, hi_n);
  p.row0 = lo;
  p.rows = hi_lo + hi_n - lo;
  p.H = H; p.W = W; p.crop = crop; p.B = B;
  return 0;
}

static long long pp_align(long long v) { return (v + 255) & ~255LL; }
struct PpWorkspace { long long w, xmin, xsize, prec, tmp, total; };
static PpWorkspace pp_layout(const PpPlan& p) {
  PpWorkspace ws;
  long long o = 0;
  ws.w = o; o += pp_align(2LL * p.crop * kPpMaxTaps * sizeof(short));
  ws.xmin = o; o += pp_align(2LL * p.crop * sizeof(int));
  ws.xsize = o; o += pp_align(2LL * p.crop * sizeof(int));
  ws.prec = o; o += 256;
  ws.tmp = o; o += pp_align((long long)p.B * p.rows * p.crop * 3);
  ws.total = o;
  return ws;
}

long long preprocess_workspace_bytes(int B, int H, int W, int short_edge, int crop) {
  PpPlan p;
  if (pp_make_plan(p, B, H, W, short_edge, crop) != 0) return -1;
  return pp_layout(p).total;
}

cudaError_t launch_preprocess(const void* images, int B, int H, int W, int short_edge, int crop, const float* mean255,
                              const float* std255, float* out, void* workspace, long long workspace_bytes, int* status,
                              cudaStream_t s) {
  PpPlan p;
  *status = pp_make_plan(p, B, H, W, short_edge, crop);
  if (*status != 0) return cudaSuccess;
  const PpWorkspace ws = pp_layout(p);
  if (workspace_bytes < ws.total) { *status = -3; return cudaSuccess; }
  uint8_t* base = static_cast<uint8_t*>(workspace);
  short* w = reinterpret_cast<short*>(base + ws.w);
  int* xmin = reinterpret_cast<int*>(base + ws.xmin);
  int* xsize = reinterpret_cast<int*>(base + ws.xsize);
  int* prec = reinterpret_cast<int*>(base + ws.prec);
  uint8_t* tmp = base + ws.tmp;
  const int threads = (crop + 31) / 32 * 32;
  launch_k<pp_weights_kernel>(2, kPpThreads, 0, s, p, w, xmin, xsize, prec);
  launch_k<pp_horizontal_kernel>(dim3(p.rows, B), threads, 0, s, p, static_cast<const uint8_t*>(images), w, xmin, xsize, prec, tmp);
  launch_k<pp_vertical_kernel>(dim3(crop, B), threads, 0, s, p, tmp, w, xmin, xsize, prec, mean255[0], mean255[1], mean255[2],
                               std255[0], std255[1], std255[2], out);
  return cudaGetLastError();
}

}  // namespace dp
