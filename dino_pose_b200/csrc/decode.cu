// Heat-map -> key-point decode, bit-exact against the reference numpy implementation
// (reference src/model_utils.py:10-36; evaluation order documented in oracle/decode_oracle.py).
//
// One warp per [H,W] map: vectorised scan for the first maximum in row-major order (NaN counts as the
// maximum, first NaN wins -- np.argmax), shuffle reduction with an index tie-break, then lane 0
// evaluates the 5x5 clipped-window centroid in exactly numpy's order: float32 sequential column / row
// sums, float32 pairwise-8 window total, float64 weighted sums.  HBM-bound: reads 4*H*W bytes per map.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "launch.cuh"

namespace dp {

struct Best {
  float v;
  int i;
  int nan;
};

__device__ __forceinline__ bool better(const Best& a, const Best& b) {
  if (a.nan != b.nan) return a.nan > b.nan;
  if (a.nan) return a.i < b.i;
  if (a.v != b.v) return a.v > b.v;
  return a.i < b.i;
}

__device__ __forceinline__ void consider(Best& best, float v, int i) {
  Best c;
  c.v = v;
  c.i = i;
  c.nan = (v != v) ? 1 : 0;
  if (better(c, best)) best = c;
}

__global__ void __launch_bounds__(128) decode_kernel(const float* __restrict__ hm, int maps, int H, int W, double tw,
                                                     double th, int* __restrict__ idx_out, double* __restrict__ xy_out,
                                                     float* __restrict__ conf_out) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int map = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (map >= maps) return;
  const int n = H * W;
  const float* m = hm + (long long)map * n;
  Best best;
  best.v = -INFINITY;
  best.i = 0x7fffffff;
  best.nan = 0;
  if (n == 2304 && ((reinterpret_cast<uintptr_t>(m) & 15) == 0)) {
    // 48 x 48 maps (every configuration of the pose models): all 18 loads of a lane are in flight before the first compare
    const float4* m4 = reinterpret_cast<const float4*>(m);
    float4 t[18];
#pragma unroll
    for (int u = 0; u < 18; ++u) t[u] = __ldg(m4 + lane + 32 * u);
#pragma unroll
    for (int u = 0; u < 18; ++u) {
      const int j = lane + 32 * u;
      consider(best, t[u].x, 4 * j);
      consider(best, t[u].y, 4 * j + 1);
      consider(best, t[u].z, 4 * j + 2);
      consider(best, t[u].w, 4 * j + 3);
    }
  } else if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(m) & 15) == 0)) {
    const float4* m4 = reinterpret_cast<const float4*>(m);
    for (int j = lane; j < n / 4; j += 32) {
      const float4 t = __ldg(m4 + j);
      consider(best, t.x, 4 * j);
      consider(best, t.y, 4 * j + 1);
      consider(best, t.z, 4 * j + 2);
      consider(best, t.w, 4 * j + 3);
    }
  } else {
    for (int j = lane; j < n; j += 32) consider(best, __ldg(m + j), j);
  }
  // a lane that saw only -inf values keeps i = INT_MAX; index 0 is the answer if everything is -inf
  if (lane == 0 && best.i == 0x7fffffff) best.i = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best other;
    other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
    other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
    other.nan = __shfl_xor_sync(0xffffffffu, best.nan, o);
    if (better(other, best)) best = other;
  }
  const int cy = best.i / W, cx = best.i % W;
  const int x0 = max(0, cx - 2), x1 = min(W, cx + 3);
  const int y0 = max(0, cy - 2), y1 = min(H, cy + 3);
  const int ww = x1 - x0, wh = y1 - y0;
  // the (up to) 25 window values are fetched by 25 lanes at once and handed to lane 0, which does the order-sensitive sums
  float mine = 0.f;
  if (lane < ww * wh) mine = m[(y0 + lane / ww) * W + (x0 + lane % ww)];
  float win[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) win[k] = __shfl_sync(0xffffffffu, mine, k);
  if (lane != 0) return;
  // float32 pairwise-8 total over the row-major window (numpy pairwise sum, 9..25 values)
  const int cnt = ww * wh;
  float total;
  if (cnt < 8) {
    total = -0.0f;
    for (int i = 0; i < cnt; ++i) total = __fadd_rn(total, win[i]);
  } else {
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = win[j];
    int i = 8;
    for (; i < cnt - (cnt % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], win[i + j]);
    total = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                      __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < cnt; ++i) total = __fadd_rn(total, win[i]);
  }
  double sx = -0.0, sy = -0.0;
  for (int x = 0; x < ww; ++x) {
    float c = win[x];
    for (int y = 1; y < wh; ++y) c = __fadd_rn(c, win[y * ww + x]);
    sx = __dadd_rn(sx, __dmul_rn(0.5 + double(x0 + x), double(c)));
  }
  for (int y = 0; y < wh; ++y) {
    float rsum = win[y * ww];
    for (int x = 1; x < ww; ++x) rsum = __fadd_rn(rsum, win[y * ww + x]);
    sy = __dadd_rn(sy, __dmul_rn(0.5 + double(y0 + y), double(rsum)));
  }
  const double lx = __dmul_rn(__ddiv_rn(__ddiv_rn(sx, double(total)), double(W)), tw);
  const double ly = __dmul_rn(__ddiv_rn(__ddiv_rn(sy, double(total)), double(H)), th);
  idx_out[2 * map] = cy;
  idx_out[2 * map + 1] = cx;
  xy_out[2 * map] = lx;
  xy_out[2 * map + 1] = ly;
  if (conf_out != nullptr) conf_out[map] = m[best.i];
}

cudaError_t launch_decode(const float* hm, int maps, int H, int W, double tw, double th, int* idx, double* xy,
                          float* conf, cudaStream_t s) {
  const int wpb = 4;   // 128 threads; 1536 maps (batch 64) = 384 blocks = 2.6 per SM
  launch_k<decode_kernel>((maps + wpb - 1) / wpb, wpb * 32, 0, s, hm, maps, H, W, tw, th, idx, xy, conf);
  return cudaGetLastError();
}

}  // namespace dp
