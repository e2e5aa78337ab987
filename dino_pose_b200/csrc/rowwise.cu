// HBM-bound row-wise kernels: LayerNorm forward / backward (warp per token, 128-bit loads, shuffle
// reductions), patch im2col for the patch-embed GEMM, CLS row fill, LoRA adapter forward / backward.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.cuh"

#include <type_traits>

#include "ptx.cuh"

namespace dp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward.  x fp32 [rows, D] -> y bf16 (and / or fp32).  V = D / 128 float4 per lane.
// drop_cls: rows are tokens [B, T]; token 0 of every image is skipped and token t goes to row
// b*(T-1) + t-1 of the output (reference model/dinov2_pose.py:147 drops CLS before the heads).
template <int V>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                            float* __restrict__ y32, long long rows, int T, int drop_cls,
                                                            float eps) {
  pdl_grid_sync();
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long orow = row;
  if (drop_cls) {
    const long long b = row / T;
    const int t = int(row % T);
    if (t == 0) return;
    orow = b * (T - 1) + (t - 1);
  }
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = __ldg(xr + lane + 32 * i);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (y != nullptr) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y + orow * D)[lane + 32 * i] = pk;
    }
    if (y32 != nullptr) reinterpret_cast<float4*>(y32 + orow * D)[lane + 32 * i] = o;
  }
}

// LayerNorm backward (input gradient only -- the backbone's LayerNorm parameters are frozen,
// reference model/dinov2_pose.py:193-194).  dy bf16 or fp32 [rows_out, D], x fp32 [rows, D]:
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))  (+ add_in)
// drop_cls mirrors the forward: dy rows exclude the CLS token, whose dx is add_in only (or 0).
template <int V, typename TDY>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const TDY* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ add_in, float* __restrict__ dx,
                                                            const float* __restrict__ ls,
                                                            __nv_bfloat16* __restrict__ dx_scaled, long long rows, int T,
                                                            int drop_cls, float eps) {
  pdl_grid_sync();
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long grow = row;
  bool has_dy = true;
  if (drop_cls) {
    const long long b = row / T;
    const int t = int(row % T);
    has_dy = (t != 0);
    grow = b * (T - 1) + (t - 1);
  }
  float4* dxr = reinterpret_cast<float4*>(dx + row * D);
  if (!has_dy) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (add_in != nullptr) o = __ldg(reinterpret_cast<const float4*>(add_in + row * D) + lane + 32 * i);
      dxr[lane + 32 * i] = o;
      if (dx_scaled != nullptr) {
        const float4 l = __ldg(reinterpret_cast<const float4*>(ls) + lane + 32 * i);
        uint2 pk;
        pk.x = pack_bf16x2(o.x * l.x, o.y * l.y);
        pk.y = pack_bf16x2(o.z * l.z, o.w * l.w);
        reinterpret_cast<uint2*>(dx_scaled + row * D)[lane + 32 * i] = pk;
      }
    }
    return;
  }
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[V], g[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = __ldg(xr + lane + 32 * i);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 d;
    if constexpr (sizeof(TDY) == 2) {
      const uint2 pk = __ldg(reinterpret_cast<const uint2*>(dy + grow * D) + lane + 32 * i);
      const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&pk.x);
      const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
      d = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
    } else {
      d = __ldg(reinterpret_cast<const float4*>(dy + grow * D) + lane + 32 * i);
    }
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
    g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
    sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
  }
  const float mg = warp_sum(sg) * (1.0f / D);
  const float mgx = warp_sum(sgx) * (1.0f / D);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 o;
    o.x = rstd * (g[i].x - mg - v[i].x * mgx);
    o.y = rstd * (g[i].y - mg - v[i].y * mgx);
    o.z = rstd * (g[i].z - mg - v[i].z * mgx);
    o.w = rstd * (g[i].w - mg - v[i].w * mgx);
    if (add_in != nullptr) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(add_in + row * D) + lane + 32 * i);
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    dxr[lane + 32 * i] = o;
    if (dx_scaled != nullptr) {
      const float4 l = __ldg(reinterpret_cast<const float4*>(ls) + lane + 32 * i);
      uint2 pk;
      pk.x = pack_bf16x2(o.x * l.x, o.y * l.y);
      pk.y = pack_bf16x2(o.z * l.z, o.w * l.w);
      reinterpret_cast<uint2*>(dx_scaled + row * D)[lane + 32 * i] = pk;
    }
  }
}

template <typename F> static cudaError_t dispatch_v(int D, F&& f) {
  switch (D) {
    case 128: f(std::integral_constant<int, 1>{}); break;
    case 384: f(std::integral_constant<int, 3>{}); break;
    case 768: f(std::integral_constant<int, 6>{}); break;
    case 1024: f(std::integral_constant<int, 8>{}); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_layernorm_fwd(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y, float* y32,
                                 long long rows, int D, int T, int drop_cls, float eps, cudaStream_t s) {
  const int wpb = 8;
  const unsigned grid = unsigned((rows + wpb - 1) / wpb);
  return dispatch_v(D, [&](auto v) {
    launch_k<layernorm_fwd_kernel<decltype(v)::value>>(grid, wpb * 32, 0, s, x, gamma, beta, y, y32, rows, T, drop_cls, eps);
  });
}

cudaError_t launch_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma,
                                 const float* add_in, float* dx, const float* ls, __nv_bfloat16* dx_scaled,
                                 long long rows, int D, int T, int drop_cls, float eps, cudaStream_t s) {
  const int wpb = 8;
  const unsigned grid = unsigned((rows + wpb - 1) / wpb);
  return dispatch_v(D, [&](auto v) {
    constexpr int V = decltype(v)::value;
    if (dy_is_bf16)
      launch_k<layernorm_bwd_kernel<V, __nv_bfloat16>>(grid, wpb * 32, 0, s, reinterpret_cast<const __nv_bfloat16*>(dy), x,
                                                                      gamma, add_in, dx, ls, dx_scaled, rows, T, drop_cls, eps);
    else
      launch_k<layernorm_bwd_kernel<V, float>>(grid, wpb * 32, 0, s, reinterpret_cast<const float*>(dy), x, gamma, add_in, dx,
                                                             ls, dx_scaled, rows, T, drop_cls, eps);
  });
}

// ------------------------------------------------------------------------------------------------
// Patch im2col: pixel_values fp32 NCHW [B,3,H,W] -> bf16 [B*gh*gw, Kp] with column index
// c*196 + ky*14 + kx (the native flattening of the conv weight [D,3,14,14], HF:139) and zero padding
// up to Kp (a multiple of 64, so the GEMM k-loop needs no tail).
__global__ void __launch_bounds__(256) patch_im2col_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ out,
                                                           int B, int H, int W, int gh, int gw, int Kp) {
  pdl_grid_sync();
  // one block per (b, patch row); threads sweep (c, ky, x-pair) with x fastest: 8-byte coalesced reads, and an
  // even x never straddles a 14-wide patch, so each pair is one 4-byte bf16x2 store
  const int b = blockIdx.x / gh, py = blockIdx.x % gh;
  const int Wv = gw * 14, Wh = Wv >> 1;
  const int total = 3 * 14 * Wh;
#pragma unroll 4
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int xh = i % Wh;
    const int t = i / Wh;
    const int ky = t % 14;
    const int c = t / 14;
    const int x = xh * 2;
    const float2 v = __ldg(reinterpret_cast<const float2*>(px + (((long long)b * 3 + c) * H + (py * 14 + ky)) * W + x));
    const long long row = ((long long)b * gh + py) * gw + x / 14;
    *reinterpret_cast<uint32_t*>(out + row * Kp + c * 196 + ky * 14 + (x % 14)) = pack_bf16x2(v.x, v.y);
  }
  // zero the K padding of the gw rows owned by this block
  const int padw = Kp - 588;
  for (int i = threadIdx.x; i < gw * padw; i += blockDim.x) {
    const long long row = ((long long)b * gh + py) * gw + i / padw;
    out[row * Kp + 588 + i % padw] = __float2bfloat16_rn(0.f);
  }
}

cudaError_t launch_patch_im2col(const float* px, __nv_bfloat16* out, int B, int H, int W, int Kp, cudaStream_t s) {
  const int gh = H / 14, gw = W / 14;
  launch_k<patch_im2col_kernel>(B * gh, 256, 0, s, px, out, B, H, W, gh, gw, Kp);
  return cudaGetLastError();
}

// x[b*T + 0, :] = cls_row[:]   (cls token + its position embedding, HF:108-112)
__global__ void fill_cls_kernel(float* __restrict__ x, const float* __restrict__ cls_row, int B, int T, int D) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i % D;
  x[(long long)b * T * D + d] = __ldg(cls_row + d);
}
cudaError_t launch_fill_cls(float* x, const float* cls_row, int B, int T, int D, cudaStream_t s) {
  launch_k<fill_cls_kernel>((B * D + 255) / 256, 256, 0, s, x, cls_row, B, T, D);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Counter-based dropout mask: keep iff hash(seed, index) >= p * 2^32.  Same function in fwd and bwd.
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return uint32_t((z ^ (z >> 31)) >> 16);
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t thresh) {
  return mix32(seed * 0xD1342543DE82EF95ull + idx) >= thresh;
}

// LoRA adapter on the attention-block output (reference model/lora.py:26-28,57-59) fused with
// LayerScale + residual (HF:373-376):
//   u = y A            [r]
//   v = dropout(u B) * s
//   x_out = x_in + (y + v) * lambda1
// y fp32 [rows, D] (output of the out-proj GEMM), A [D, r], Bm [r, D] fp32 (trainable, read directly).
// One warp per row; A and B staged in shared memory.  Saves u (fp32 [rows, R]) for the backward.
template <int R>
__global__ void __launch_bounds__(256) lora_fwd_kernel(const float* __restrict__ y, const float* __restrict__ A,
                                                       const float* __restrict__ Bm, const float* __restrict__ lambda1,
                                                       const float* __restrict__ x_in, float* __restrict__ x_out,
                                                       float* __restrict__ u_save, long long rows, int D, float scaling,
                                                       float p_drop, const unsigned long long* __restrict__ seed_ptr) {
  pdl_grid_sync();
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  extern __shared__ float sm[];
  float* sA = sm;           // [D][R]
  float* sB = sm + D * R;   // [R][D]
  for (int i = threadIdx.x; i < D * R; i += blockDim.x) {
    sA[i] = A[i];
    sB[i] = Bm[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    float u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = 0.f;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) {
      const float yv = __ldg(y + row * D + d);
#pragma unroll
      for (int r = 0; r < R; ++r) u[r] += yv * sA[d * R + r];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = warp_sum(u[r]);
    if (u_save != nullptr && lane < R) {
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) mine = (lane == r) ? u[r] : mine;
      u_save[row * R + lane] = mine;
    }
#pragma unroll 4
    for (int d = lane; d < D; d += 32) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) v += u[r] * sB[r * D + d];
      if (p_drop > 0.f) v = dropout_keep(seed, uint64_t(row) * D + d, thresh) ? v * keep_scale : 0.f;
      const float yv = y[row * D + d];
      x_out[row * D + d] = x_in[row * D + d] + (yv + v * scaling) * lambda1[d];
    }
  }
}

// LoRA backward.  g = dL/dx_out [rows, D] fp32 (gradient of the residual stream after the attention
// branch).  With gv = g * lambda1 * s * mask/(1-p) (gradient wrt u B):
//   gu[row, r] = sum_d gv[row, d] * B[r, d]                     (kernel 1, warp per row)
//   dB[r, d]  += sum_rows u[row, r] * gv[row, d]                 (kernel 2, thread per column d)
//   dA[d, r]  += sum_rows y[row, d] * gu[row, r]
// (no gradient flows further: y is produced by frozen parameters from a frozen input.)
template <int R>
__global__ void __launch_bounds__(256) lora_bwd_gu_kernel(const float* __restrict__ g, const float* __restrict__ Bm,
                                                          const float* __restrict__ lambda1, float* __restrict__ gu_out,
                                                          long long rows, int D, float scaling, float p_drop,
                                                          const unsigned long long* __restrict__ seed_ptr) {
  pdl_grid_sync();
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  extern __shared__ float sm[];
  float* sB = sm;  // [R][D]
  for (int i = threadIdx.x; i < D * R; i += blockDim.x) sB[i] = Bm[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    float gu[R];
#pragma unroll
    for (int r = 0; r < R; ++r) gu[r] = 0.f;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) {
      float gv = __ldg(g + row * D + d) * __ldg(lambda1 + d) * scaling;
      if (p_drop > 0.f) gv = dropout_keep(seed, uint64_t(row) * D + d, thresh) ? gv * keep_scale : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) gu[r] += gv * sB[r * D + d];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) gu[r] = warp_sum(gu[r]);
    if (lane < R) {
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) mine = (lane == r) ? gu[r] : mine;
      gu_out[row * R + lane] = mine;
    }
  }
}

template <int R>
__global__ void __launch_bounds__(256) lora_bwd_acc_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                           const float* __restrict__ u_saved, const float* __restrict__ gu,
                                                           const float* __restrict__ lambda1, float* __restrict__ dA,
                                                           float* __restrict__ dB, long long rows, int D, float scaling,
                                                           float p_drop, const unsigned long long* __restrict__ seed_ptr,
                                                           int rows_per_block) {
  pdl_grid_sync();
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const float ld = lambda1[d] * scaling;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, rows);
  float aB[R], aA[R];
#pragma unroll
  for (int r = 0; r < R; ++r) aB[r] = aA[r] = 0.f;
  // 4 rows per iteration: the g / y loads of all four rows are in flight together (the loop is latency bound otherwise)
  for (long long row = r0; row < r1; row += 4) {
    float gv[4], yv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long rr = row + k < r1 ? row + k : r1 - 1;
      gv[k] = __ldg(g + rr * D + d);
      yv[k] = __ldg(y + rr * D + d);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (row + k >= r1) break;
      float gk = gv[k] * ld;
      if (p_drop > 0.f) gk = dropout_keep(seed, uint64_t(row + k) * D + d, thresh) ? gk * keep_scale : 0.f;
      const float4* up = reinterpret_cast<const float4*>(u_saved + (row + k) * R);
      const float4* gp = reinterpret_cast<const float4*>(gu + (row + k) * R);
#pragma unroll
      for (int r4 = 0; r4 < R / 4; ++r4) {
        const float4 uu = __ldg(up + r4), gg = __ldg(gp + r4);
        aB[4 * r4 + 0] += uu.x * gk; aB[4 * r4 + 1] += uu.y * gk; aB[4 * r4 + 2] += uu.z * gk; aB[4 * r4 + 3] += uu.w * gk;
        aA[4 * r4 + 0] += yv[k] * gg.x; aA[4 * r4 + 1] += yv[k] * gg.y; aA[4 * r4 + 2] += yv[k] * gg.z; aA[4 * r4 + 3] += yv[k] * gg.w;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    atomicAdd(dB + (long long)r * D + d, aB[r]);
    atomicAdd(dA + (long long)d * R + r, aA[r]);
  }
}

cudaError_t launch_lora_fwd(const float* y, const float* A, const float* Bm, const float* lambda1, const float* x_in,
                            float* x_out, float* u_save, long long rows, int D, int R, float scaling, float p_drop,
                            const unsigned long long* seed, int sms, cudaStream_t s) {
  const size_t smem = size_t(2) * D * R * sizeof(float);
  const int grid = sms * 6;
  if (R == 8) {
    cudaFuncSetAttribute(lora_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    launch_k<lora_fwd_kernel<8>>(grid, 256, smem, s, y, A, Bm, lambda1, x_in, x_out, u_save, rows, D, scaling, p_drop, seed);
  } else if (R == 4) {
    cudaFuncSetAttribute(lora_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    launch_k<lora_fwd_kernel<4>>(grid, 256, smem, s, y, A, Bm, lambda1, x_in, x_out, u_save, rows, D, scaling, p_drop, seed);
  } else if (R == 16) {
    cudaFuncSetAttribute(lora_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    launch_k<lora_fwd_kernel<16>>(grid, 256, smem, s, y, A, Bm, lambda1, x_in, x_out, u_save, rows, D, scaling, p_drop, seed);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <int R>
static cudaError_t lora_bwd_t(const float* g, const float* y, const float* u_saved, const float* Bm, const float* lambda1,
                              float* dA, float* dB, float* gu_ws, long long rows, int D, float scaling, float p_drop,
                              const unsigned long long* seed, int sms, cudaStream_t s) {
  const size_t smem = size_t(D) * R * sizeof(float);
  cudaFuncSetAttribute(lora_bwd_gu_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  launch_k<lora_bwd_gu_kernel<R>>(sms * 6, 256, smem, s, g, Bm, lambda1, gu_ws, rows, D, scaling, p_drop, seed);
  const int bx = 128;
  const int gy = (D + bx - 1) / bx;
  int gx = (sms * 4) / gy;   // more blocks = more same-address atomics on the 2*D*R outputs (measured slower)
  if (gx < 1) gx = 1;
  int rpb = int((rows + gx - 1) / gx);
  if (rpb < 16) rpb = 16;
  gx = int((rows + rpb - 1) / rpb);
  launch_k<lora_bwd_acc_kernel<R>>(dim3(gx, gy), bx, 0, s, g, y, u_saved, gu_ws, lambda1, dA, dB, rows, D, scaling, p_drop, seed, rpb);
  return cudaGetLastError();
}

cudaError_t launch_lora_bwd(const float* g, const float* y, const float* u_saved, const float* Bm, const float* lambda1,
                            float* dA, float* dB, float* gu_ws, long long rows, int D, int R, float scaling, float p_drop,
                            const unsigned long long* seed, int sms, cudaStream_t s) {
  if (R == 8) return lora_bwd_t<8>(g, y, u_saved, Bm, lambda1, dA, dB, gu_ws, rows, D, scaling, p_drop, seed, sms, s);
  if (R == 4) return lora_bwd_t<4>(g, y, u_saved, Bm, lambda1, dA, dB, gu_ws, rows, D, scaling, p_drop, seed, sms, s);
  if (R == 16) return lora_bwd_t<16>(g, y, u_saved, Bm, lambda1, dA, dB, gu_ws, rows, D, scaling, p_drop, seed, sms, s);
  return cudaErrorInvalidValue;
}

}  // namespace dp
