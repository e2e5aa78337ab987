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

// LayerNorm parameter gradients of the un-frozen encoder layers (reference model/dinov2_pose.py:25-39 un-freezes
// norm1 / norm2 of the last n blocks):  dgamma[c] += sum_rows dy[row,c] * xhat[row,c],  dbeta[c] += sum_rows dy[row,c].
// Persistent blocks of 8 warps; a warp walks rows (statistics recomputed from x like the input-gradient kernel), a lane
// keeps the partial sums of its 4*V columns in registers; one shared-memory reduction and 2*D atomics per block.
template <int V, typename TDY>
__global__ void __launch_bounds__(256) layernorm_bwd_params_kernel(const TDY* __restrict__ dy, const float* __restrict__ x,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   long long rows, float eps) {
  pdl_grid_sync();
  constexpr int D = V * 128;
  __shared__ float red[4][2 * D];   // two-step reduction over the 8 warps (32 KB at D = 1024)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 ag[V], ab[V];
#pragma unroll
  for (int i = 0; i < V; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
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
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 d;
      if constexpr (sizeof(TDY) == 2) {
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(dy + row * D) + lane + 32 * i);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&pk.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
        d = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
      } else {
        d = __ldg(reinterpret_cast<const float4*>(dy + row * D) + lane + 32 * i);
      }
      ag[i].x = fmaf(d.x, v[i].x * rstd, ag[i].x); ag[i].y = fmaf(d.y, v[i].y * rstd, ag[i].y);
      ag[i].z = fmaf(d.z, v[i].z * rstd, ag[i].z); ag[i].w = fmaf(d.w, v[i].w * rstd, ag[i].w);
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
    }
  }
  if (warp >= 4) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      *reinterpret_cast<float4*>(&red[warp - 4][4 * (lane + 32 * i)]) = ag[i];
      *reinterpret_cast<float4*>(&red[warp - 4][D + 4 * (lane + 32 * i)]) = ab[i];
    }
  }
  __syncthreads();
  if (warp < 4) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 g2 = *reinterpret_cast<const float4*>(&red[warp][4 * (lane + 32 * i)]);
      const float4 b2 = *reinterpret_cast<const float4*>(&red[warp][D + 4 * (lane + 32 * i)]);
      ag[i].x += g2.x; ag[i].y += g2.y; ag[i].z += g2.z; ag[i].w += g2.w;
      ab[i].x += b2.x; ab[i].y += b2.y; ab[i].z += b2.z; ab[i].w += b2.w;
      *reinterpret_cast<float4*>(&red[warp][4 * (lane + 32 * i)]) = ag[i];
      *reinterpret_cast<float4*>(&red[warp][D + 4 * (lane + 32 * i)]) = ab[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    const float t = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
    atomicAdd((c < D ? dgamma + c : dbeta + (c - D)), t);
  }
}

// out[c] += sum_rows g[row,c] * a[row,c]  (LayerScale gradient: g fp32 = gradient of the residual stream, a bf16 = the
// branch output the scale multiplied; HF:272-278)
__global__ void __launch_bounds__(128) colsum_prod_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ a,
                                                          float* __restrict__ out, long long P, int C, int rows_per_block) {
  pdl_grid_sync();
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long p0 = (long long)blockIdx.x * rows_per_block;
  const long long p1 = min(p0 + rows_per_block, P);
  float s0 = 0.f, s1 = 0.f;
  long long p = p0;
  for (; p + 1 < p1; p += 2) {
    s0 = fmaf(__ldg(g + p * C + c), __bfloat162float(a[p * C + c]), s0);
    s1 = fmaf(__ldg(g + (p + 1) * C + c), __bfloat162float(a[(p + 1) * C + c]), s1);
  }
  if (p < p1) s0 = fmaf(__ldg(g + p * C + c), __bfloat162float(a[p * C + c]), s0);
  atomicAdd(out + c, s0 + s1);
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

cudaError_t launch_layernorm_bwd_params(const void* dy, int dy_is_bf16, const float* x, float* dgamma, float* dbeta,
                                        long long rows, int D, float eps, int sms, cudaStream_t s) {
  long long want = (rows + 7) / 8;
  const unsigned grid = unsigned(want < 2LL * sms ? (want < 1 ? 1 : want) : 2LL * sms);
  return dispatch_v(D, [&](auto v) {
    constexpr int V = decltype(v)::value;
    if (dy_is_bf16)
      launch_k<layernorm_bwd_params_kernel<V, __nv_bfloat16>>(grid, 256, 0, s, reinterpret_cast<const __nv_bfloat16*>(dy), x,
                                                             dgamma, dbeta, rows, eps);
    else
      launch_k<layernorm_bwd_params_kernel<V, float>>(grid, 256, 0, s, reinterpret_cast<const float*>(dy), x, dgamma, dbeta,
                                                     rows, eps);
  });
}

cudaError_t launch_colsum_prod(const float* g, const __nv_bfloat16* a, float* out, long long P, int C, cudaStream_t s) {
  int rpb = int((P + 295) / 296);
  if (rpb < 32) rpb = 32;
  dim3 grid(unsigned((P + rpb - 1) / rpb), unsigned((C + 127) / 128));
  launch_k<colsum_prod_kernel>(grid, 128, 0, s, g, a, out, P, C, rpb);
  return cudaGetLastError();
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

// Same result through shared memory: the 3 x 14 image rows of one patch row are read with coalesced 8-byte loads and kept
// as bf16 [c][ky][x]; the gw output rows (Kp bf16 = 1280 bytes each) then leave as full 16-byte pieces, 8 consecutive k
// gathered from the staged band (the direct kernel above writes 4-byte pieces 28 bytes apart: 32 us for 60 MB).
__global__ void __launch_bounds__(256) patch_im2col_staged_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ out,
                                                                  int B, int H, int W, int gh, int gw, int Kp) {
  pdl_grid_sync();
  extern __shared__ __nv_bfloat16 s_band[];      // [42][Wv]
  const int b = blockIdx.x / gh, py = blockIdx.x % gh;
  const int Wv = gw * 14, Wh = Wv >> 1;
  const int total = 42 * Wh;
#pragma unroll 4
  for (int i = threadIdx.x; i < total; i += 256) {
    const int r = i / Wh, xh = i - r * Wh;       // r = c * 14 + ky
    const int c = r / 14, ky = r - c * 14;
    const float2 v = __ldg(reinterpret_cast<const float2*>(px + (((long long)b * 3 + c) * H + (py * 14 + ky)) * W + 2 * xh));
    *reinterpret_cast<uint32_t*>(s_band + r * Wv + 2 * xh) = pack_bf16x2(v.x, v.y);
  }
  __syncthreads();
  const int K8 = Kp >> 3;                        // 16-byte pieces per output row
  const unsigned short* sb = reinterpret_cast<const unsigned short*>(s_band);
  for (int i = threadIdx.x; i < gw * K8; i += 256) {
    const int p = i / K8, k0 = (i - p * K8) * 8;
    uint32_t w[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      uint32_t lo = 0u, hi = 0u;
      const int ka = k0 + 2 * h, kb = ka + 1;
      if (ka < 588) { const int r = ka / 14; lo = sb[r * Wv + p * 14 + (ka - r * 14)]; }
      if (kb < 588) { const int r = kb / 14; hi = sb[r * Wv + p * 14 + (kb - r * 14)]; }
      w[h] = lo | (hi << 16);
    }
    const long long row = ((long long)b * gh + py) * gw + p;
    *reinterpret_cast<uint4*>(out + row * Kp + k0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

cudaError_t launch_patch_im2col(const float* px, __nv_bfloat16* out, int B, int H, int W, int Kp, cudaStream_t s) {
  const int gh = H / 14, gw = W / 14;
  const size_t smem = size_t(42) * gw * 14 * sizeof(__nv_bfloat16);
  static int staged = -1;     // DP_PATCH_STAGED=0: the direct kernel (A/B)
  if (staged < 0) { const char* v = getenv("DP_PATCH_STAGED"); staged = v ? atoi(v) : 1; }
  if (staged && smem <= 48 * 1024 && Kp % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    launch_k<patch_im2col_staged_kernel>(B * gh, 256, smem, s, px, out, B, H, W, gh, gw, Kp);
    return cudaGetLastError();
  }
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
// (counter-based dropout mask: mix32 / dropout_keep live in ptx.cuh, shared with the LoRA-fused GEMM epilogue)

// LoRA adapter on the attention-block output (reference model/lora.py:26-28,57-59) fused with
// LayerScale + residual (HF:373-376):
//   u = y A            [r]
//   v = dropout(u B) * s
//   x_out = x_in + (y + v) * lambda1
// y fp32 [rows, D] (output of the out-proj GEMM), A [D, r], Bm [r, D] fp32 (trainable, read directly).
// One warp per row; A and B staged in shared memory.  Saves u (fp32 [rows, R]) for the backward.
template <int R, int NV>
__global__ void __launch_bounds__(256) lora_fwd_kernel(const float* __restrict__ y, const float* __restrict__ A,
                                                       const float* __restrict__ Bm, const float* __restrict__ lambda1,
                                                       const float* __restrict__ x_in, float* __restrict__ x_out,
                                                       float* __restrict__ u_save, long long rows, int D, float scaling,
                                                       float p_drop, const unsigned long long* __restrict__ seed_ptr) {
  pdl_grid_sync();
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  extern __shared__ float sm[];
  float* sAT = sm;          // [R][D]  (A transposed: a lane's float4 of columns is contiguous, conflict free)
  float* sB = sm + D * R;   // [R][D]
  for (int i = threadIdx.x; i < D * R; i += blockDim.x) {
    sAT[(i % R) * D + i / R] = A[i];
    sB[i] = Bm[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  // a lane owns the float4 column groups 4*lane + 128*i (D % 128 == 0, D <= 1024): every global access is a 16-byte
  // load / store, and all of a row's loads (y and x_in) are issued before the first use -- with scalar loads and four
  // of them in flight per lane the kernel ran at a third of the HBM rate
  // NV = D / 128 is a template parameter: with run-time trip counts the compiler kept 8 float4 pairs per lane alive
  // (139 registers, ONE resident block per SM, and 6 waves of blocks each re-staging A and B: 47 us for 76 MB)
  constexpr int kMaxV = NV;
  constexpr int nv = NV;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    float4 yv[kMaxV], xv[kMaxV];
#pragma unroll
    for (int i = 0; i < kMaxV; ++i)
      if (i < nv) {
        yv[i] = __ldg(reinterpret_cast<const float4*>(y + row * D) + lane + 32 * i);
        xv[i] = __ldg(reinterpret_cast<const float4*>(x_in + row * D) + lane + 32 * i);
      }
    float u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i)
      if (i < nv) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(sAT + r * D + 4 * lane + 128 * i);
          u[r] = fmaf(yv[i].x, a.x, fmaf(yv[i].y, a.y, fmaf(yv[i].z, a.z, fmaf(yv[i].w, a.w, u[r]))));
        }
      }
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = warp_sum(u[r]);
    if (u_save != nullptr && lane < R) {
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) mine = (lane == r) ? u[r] : mine;
      u_save[row * R + lane] = mine;
    }
#pragma unroll
    for (int i = 0; i < kMaxV; ++i)
      if (i < nv) {
        const int d = 4 * lane + 128 * i;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 b = *reinterpret_cast<const float4*>(sB + r * D + d);
          v[0] = fmaf(u[r], b.x, v[0]); v[1] = fmaf(u[r], b.y, v[1]); v[2] = fmaf(u[r], b.z, v[2]); v[3] = fmaf(u[r], b.w, v[3]);
        }
        if (p_drop > 0.f) {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = dropout_keep(seed, uint64_t(row) * D + d + k, thresh) ? v[k] * keep_scale : 0.f;
        }
        const float4 l = __ldg(reinterpret_cast<const float4*>(lambda1 + d));
        float4 o;
        o.x = xv[i].x + (yv[i].x + v[0] * scaling) * l.x;
        o.y = xv[i].y + (yv[i].y + v[1] * scaling) * l.y;
        o.z = xv[i].z + (yv[i].z + v[2] * scaling) * l.z;
        o.w = xv[i].w + (yv[i].w + v[3] * scaling) * l.w;
        *(reinterpret_cast<float4*>(x_out + row * D) + lane + 32 * i) = o;
      }
  }
}

// LoRA backward.  g = dL/dx_out [rows, D] fp32 (gradient of the residual stream after the attention
// branch).  With gv = g * lambda1 * s * mask/(1-p) (gradient wrt u B):
//   gu[row, r] = sum_d gv[row, d] * B[r, d]                     (kernel 1, warp per row)
//   dB[r, d]  += sum_rows u[row, r] * gv[row, d]                 (kernel 2, thread per column d)
//   dA[d, r]  += sum_rows y[row, d] * gu[row, r]
// (no gradient flows further: y is produced by frozen parameters from a frozen input.)
template <int R, int NV>
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
  constexpr int kMaxV = NV;     // float4 column groups per lane (D = 128 * NV), all loads of a row in flight
  constexpr int nv = NV;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    float4 gv4[kMaxV];
#pragma unroll
    for (int i = 0; i < kMaxV; ++i)
      if (i < nv) gv4[i] = __ldg(reinterpret_cast<const float4*>(g + row * D) + lane + 32 * i);
    float gu[R];
#pragma unroll
    for (int r = 0; r < R; ++r) gu[r] = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i)
      if (i < nv) {
        const int d = 4 * lane + 128 * i;
        const float4 l = __ldg(reinterpret_cast<const float4*>(lambda1 + d));
        float gv[4] = {gv4[i].x * l.x * scaling, gv4[i].y * l.y * scaling, gv4[i].z * l.z * scaling, gv4[i].w * l.w * scaling};
        if (p_drop > 0.f) {
#pragma unroll
          for (int k = 0; k < 4; ++k) gv[k] = dropout_keep(seed, uint64_t(row) * D + d + k, thresh) ? gv[k] * keep_scale : 0.f;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 b = *reinterpret_cast<const float4*>(sB + r * D + d);
          gu[r] = fmaf(gv[0], b.x, fmaf(gv[1], b.y, fmaf(gv[2], b.z, fmaf(gv[3], b.w, gu[r]))));
        }
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

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// A thread owns FOUR columns (one float4 of g and y per row) and walks over rows eight at a time: 16 loads of 16 bytes in
// flight per thread.  blockDim.x = (D / 4) * row_lanes: the row lanes split the block's rows (the D / 4 = 96 threads of
// ViT-S alone are 6 warps per SM -- issue bound) and are combined through shared memory before the atomics.
template <int R>
__global__ void __launch_bounds__(384) lora_bwd_acc_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                           const float* __restrict__ u_saved, const float* __restrict__ gu,
                                                           const float* __restrict__ lambda1, float* __restrict__ dA,
                                                           float* __restrict__ dB, long long rows, int D, float scaling,
                                                           float p_drop, const unsigned long long* __restrict__ seed_ptr,
                                                           int rows_per_block, int row_lanes) {
  pdl_grid_sync();
  extern __shared__ float s_red[];   // [row_lanes - 1][D / 4][8 * R]
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  const int cols = D >> 2;
  const int ct = threadIdx.x % cols, rl = threadIdx.x / cols;
  const int d = 4 * ct;
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const float4 l4 = __ldg(reinterpret_cast<const float4*>(lambda1 + d));
  const float ld[4] = {l4.x * scaling, l4.y * scaling, l4.z * scaling, l4.w * scaling};
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, rows);
  float aB[R][4], aA[4][R];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) aB[r][k] = aA[k][r] = 0.f;
  constexpr int kRows = 8;
  for (long long base = r0 + rl; base < r1; base += (long long)kRows * row_lanes) {
    float4 gv[kRows], yv[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      long long rr = base + (long long)k * row_lanes;
      if (rr >= r1) rr = base;
      gv[k] = __ldg(reinterpret_cast<const float4*>(g + rr * D + d));
      yv[k] = __ldg(reinterpret_cast<const float4*>(y + rr * D + d));
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const long long row = base + (long long)k * row_lanes;
      if (row >= r1) break;
      float gk[4] = {gv[k].x * ld[0], gv[k].y * ld[1], gv[k].z * ld[2], gv[k].w * ld[3]};
      if (p_drop > 0.f) {
#pragma unroll
        for (int c = 0; c < 4; ++c) gk[c] = dropout_keep(seed, uint64_t(row) * D + d + c, thresh) ? gk[c] * keep_scale : 0.f;
      }
      const float yk[4] = {yv[k].x, yv[k].y, yv[k].z, yv[k].w};
      const float4* up = reinterpret_cast<const float4*>(u_saved + row * R);   // same address in every thread of a lane: broadcast
      const float4* gp = reinterpret_cast<const float4*>(gu + row * R);
#pragma unroll
      for (int r4 = 0; r4 < R / 4; ++r4) {
        const float4 uu = __ldg(up + r4), gg = __ldg(gp + r4);
        const float uv[4] = {uu.x, uu.y, uu.z, uu.w}, gq[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            aB[4 * r4 + q][c] = fmaf(uv[q], gk[c], aB[4 * r4 + q][c]);
            aA[c][4 * r4 + q] = fmaf(yk[c], gq[q], aA[c][4 * r4 + q]);
          }
      }
    }
  }
  if (rl > 0) {
    float* dst = s_red + ((long long)(rl - 1) * cols + ct) * (8 * R);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        dst[r * 4 + c] = aB[r][c];
        dst[4 * R + c * R + r] = aA[c][r];
      }
  }
  __syncthreads();
  if (rl == 0) {
    for (int l = 1; l < row_lanes; ++l) {
      const float* src = s_red + ((long long)(l - 1) * cols + ct) * (8 * R);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          aB[r][c] += src[r * 4 + c];
          aA[c][r] += src[4 * R + c * R + r];
        }
    }
    // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the atomic operations -- with ~300 blocks adding
    // into the same 2*D*R addresses the L2 atomic rate, not the 50 MB of loads, bounded this kernel
#pragma unroll
    for (int r = 0; r < R; ++r) red_add_v4(dB + (long long)r * D + d, aB[r][0], aB[r][1], aB[r][2], aB[r][3]);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r4 = 0; r4 < R / 4; ++r4)
        red_add_v4(dA + (long long)(d + c) * R + 4 * r4, aA[c][4 * r4], aA[c][4 * r4 + 1], aA[c][4 * r4 + 2], aA[c][4 * r4 + 3]);
  }
}

template <int R, int NV>
static cudaError_t lora_fwd_t(const float* y, const float* A, const float* Bm, const float* lambda1, const float* x_in,
                              float* x_out, float* u_save, long long rows, int D, float scaling, float p_drop,
                              const unsigned long long* seed, int sms, cudaStream_t s) {
  const size_t smem = size_t(2) * D * R * sizeof(float);
  cudaFuncSetAttribute(lora_fwd_kernel<R, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  // persistent blocks (3 per SM fit by registers and shared memory): A and B are staged once per block
  launch_k<lora_fwd_kernel<R, NV>>(sms * 3, 256, smem, s, y, A, Bm, lambda1, x_in, x_out, u_save, rows, D, scaling, p_drop, seed);
  return cudaGetLastError();
}
#define DP_LORA_DISPATCH(FN, ...)                                                     \
  switch (R * 16 + D / 128) {                                                         \
    case 4 * 16 + 1: return FN<4, 1>(__VA_ARGS__);                                    \
    case 4 * 16 + 3: return FN<4, 3>(__VA_ARGS__);                                    \
    case 4 * 16 + 6: return FN<4, 6>(__VA_ARGS__);                                    \
    case 4 * 16 + 8: return FN<4, 8>(__VA_ARGS__);                                    \
    case 8 * 16 + 1: return FN<8, 1>(__VA_ARGS__);                                    \
    case 8 * 16 + 3: return FN<8, 3>(__VA_ARGS__);                                    \
    case 8 * 16 + 6: return FN<8, 6>(__VA_ARGS__);                                    \
    case 8 * 16 + 8: return FN<8, 8>(__VA_ARGS__);                                    \
    case 16 * 16 + 1: return FN<16, 1>(__VA_ARGS__);                                  \
    case 16 * 16 + 3: return FN<16, 3>(__VA_ARGS__);                                  \
    case 16 * 16 + 6: return FN<16, 6>(__VA_ARGS__);                                  \
    case 16 * 16 + 8: return FN<16, 8>(__VA_ARGS__);                                  \
    default: return cudaErrorInvalidValue;                                            \
  }
cudaError_t launch_lora_fwd(const float* y, const float* A, const float* Bm, const float* lambda1, const float* x_in,
                            float* x_out, float* u_save, long long rows, int D, int R, float scaling, float p_drop,
                            const unsigned long long* seed, int sms, cudaStream_t s) {
  // float4 column groups per lane: D = 128 * {1, 3, 6, 8} (test / ViT-S / ViT-B / ViT-L), rank 4 / 8 / 16
  if (D % 128) return cudaErrorInvalidValue;
  DP_LORA_DISPATCH(lora_fwd_t, y, A, Bm, lambda1, x_in, x_out, u_save, rows, D, scaling, p_drop, seed, sms, s)
}

template <int R, int NV>
static cudaError_t lora_bwd_t(const float* g, const float* y, const float* u_saved, const float* Bm, const float* lambda1,
                              float* dA, float* dB, float* gu_ws, long long rows, int D, float scaling, float p_drop,
                              const unsigned long long* seed, int sms, cudaStream_t s) {
  const size_t smem = size_t(D) * R * sizeof(float);
  cudaFuncSetAttribute(lora_bwd_gu_kernel<R, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  launch_k<lora_bwd_gu_kernel<R, NV>>(sms * 3, 256, smem, s, g, Bm, lambda1, gu_ws, rows, D, scaling, p_drop, seed);
  if (D % 128 || D > 1024) return cudaErrorInvalidValue;
  const int cols = D / 4;          // one thread per float4 column group ...
  int lanes = 384 / cols;          // ... times up to 4 row lanes (ViT-S: 4, ViT-B: 2, ViT-L: 1)
  if (lanes < 1) lanes = 1;
  if (lanes > 4) lanes = 4;
  if (R == 16) lanes = 1;          // 250 registers per thread
  int gx = sms * 2;                // more blocks = more same-address atomics on the 2*D*R outputs (measured slower)
  int rpb = int((rows + gx - 1) / gx);
  if (rpb < 16) rpb = 16;
  gx = int((rows + rpb - 1) / rpb);
  const size_t red = size_t(lanes - 1) * cols * 8 * R * sizeof(float);
  cudaFuncSetAttribute(lora_bwd_acc_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  launch_k<lora_bwd_acc_kernel<R>>(gx, cols * lanes, red, s, g, y, u_saved, gu_ws, lambda1, dA, dB, rows, D, scaling, p_drop, seed,
                                   rpb, lanes);
  return cudaGetLastError();
}

cudaError_t launch_lora_bwd_mma(const float* g, const float* y, const float* u, const float* Bm, const float* lambda1,
                                float* dA, float* dB, long long rows, int D, int R, float scaling, float p_drop,
                                const unsigned long long* seed, int sms, cudaStream_t s);

cudaError_t launch_lora_bwd(const float* g, const float* y, const float* u_saved, const float* Bm, const float* lambda1,
                            float* dA, float* dB, float* gu_ws, long long rows, int D, int R, float scaling, float p_drop,
                            const unsigned long long* seed, int sms, cudaStream_t s) {
  if (D % 128) return cudaErrorInvalidValue;
  // rank 8, D <= 384: one pass over g and y with warp-level MMAs (lora_bwd_mma.cu); DP_LORA_BWD_MMA=0 keeps the two fp32
  // kernels below (A/B, and the shapes the MMA kernel does not cover)
  static int use_mma = -1;
  if (use_mma < 0) { const char* v = getenv("DP_LORA_BWD_MMA"); use_mma = v ? atoi(v) : 1; }
  if (use_mma) {
    const cudaError_t e = launch_lora_bwd_mma(g, y, u_saved, Bm, lambda1, dA, dB, rows, D, R, scaling, p_drop, seed, sms, s);
    if (e != cudaErrorNotSupported) return e;
  }
  DP_LORA_DISPATCH(lora_bwd_t, g, y, u_saved, Bm, lambda1, dA, dB, gu_ws, rows, D, scaling, p_drop, seed, sms, s)
}

}  // namespace dp
