// Memory-bound kernels of the pose heads (reference model/pose_heads.py:211-400), NHWC bf16 activations:
// im2col / col2im for the strided and transposed convolutions, depthwise 3x3, train-mode BatchNorm
// (statistics, finalize + running-stat update, apply with fused ReLU / residual adds) and its backward,
// bilinear 2x reduction, token mean-pool, fp32 SIMT GEMM for the tiny z-head MLP, layout conversions.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.cuh"

#include "ptx.cuh"

namespace dp {

namespace {

__device__ __forceinline__ void unpack8(const uint4& t, float (&f)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
    f[2 * k] = __low2float(h);
    f[2 * k + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 t;
  t.x = pack_bf16x2(f[0], f[1]);
  t.y = pack_bf16x2(f[2], f[3]);
  t.z = pack_bf16x2(f[4], f[5]);
  t.w = pack_bf16x2(f[6], f[7]);
  return t;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]);
__device__ __forceinline__ void store8(float* p, const float (&f)[8]);
inline unsigned grid_for(long long n, int block, int max_blocks = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return unsigned(g);
}

// ------------------------------------------------------------------------------------------------
// im2col: in NHWC [NB,IH,IW,C] -> col [NB*OH*OW, KH*KW*C], column = (ky*KW + kx)*C + c, zero padding.
__global__ void __launch_bounds__(256) im2col_kernel(const uint4* __restrict__ in, uint4* __restrict__ col, int NB, int IH,
                                                     int IW, int C8, int OH, int OW, int KH, int KW, int stride, int pad) {
  pdl_grid_sync();
  const long long total = (long long)NB * OH * OW * KH * KW * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C8);
    long long r = i / C8;
    const int kx = int(r % KW); r /= KW;
    const int ky = int(r % KH); r /= KH;
    const int ox = int(r % OW); r /= OW;
    const int oy = int(r % OH);
    const int b = int(r / OH);
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < IH && ix >= 0 && ix < IW) v = __ldg(in + (((long long)b * IH + iy) * IW + ix) * C8 + c);
    col[i] = v;
  }
}

// col2im (gather form): big[b,y,x,c] = bias[c] + sum_{ky,kx} col[(b,sy,sx), (ky*KW+kx)*C + c]
// over taps with y = sy*stride - pad + ky, x = sx*stride - pad + kx, (sy,sx) inside the small SHxSW grid.
// Forward of ConvTranspose2d (small = input grid) and input-gradient of a strided Conv2d (small = output grid).
template <typename OutT>
__global__ void __launch_bounds__(256) col2im_kernel(const uint4* __restrict__ col, const float* __restrict__ bias,
                                                     OutT* __restrict__ big, int NB, int SH, int SW, int C8, int BH, int BW,
                                                     int KH, int KW, int stride, int pad) {
  pdl_grid_sync();
  const long long total = (long long)NB * BH * BW * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C8);
    long long r = i / C8;
    const int x = int(r % BW); r /= BW;
    const int y = int(r % BH);
    const int b = int(r / BH);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? __ldg(bias + c * 8 + j) : 0.f;
    // only the taps with (y + pad - ky) % stride == 0 contribute: start at ky0 = (y + pad) % stride and step by the
    // stride (1-2 taps per axis for k4 s3) instead of testing all KH x KW taps with a division each
    const int ky0 = (y + pad) % stride, kx0 = (x + pad) % stride;
    for (int ky = ky0; ky < KH; ky += stride) {
      const int ty = y + pad - ky;
      if (ty < 0) break;
      const int sy = ty / stride;
      if (sy >= SH) continue;
      for (int kx = kx0; kx < KW; kx += stride) {
        const int tx = x + pad - kx;
        if (tx < 0) break;
        const int sx = tx / stride;
        if (sx >= SW) continue;
        const uint4 t = __ldg(col + ((((long long)b * SH + sy) * SW + sx) * (KH * KW) + (ky * KW + kx)) * C8 + c);
        float f[8];
        unpack8(t, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    store8(big + i * 8, acc);
  }
}

// depthwise 3x3, pad 1 (HourglassModule.depthwise_conv[0], pose_heads.py:219).  w fp32 [C,1,3,3].
// flip = 1 gives the input-gradient (correlation with the flipped kernel).
// A thread owns ONE group of 8 channels for the whole kernel: its 72 filter taps and 8 biases live in registers,
// and it walks over pixels (4 pixel lanes per 256-thread block when C = 512), 9 x 16-byte loads + 72 FMAs per pixel.
template <typename OutT>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const uint4* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const uint4* __restrict__ add,
                                                        OutT* __restrict__ out, int NB, int H, int W, int C8, int flip) {
  pdl_grid_sync();
  const int c = threadIdx.x % C8;
  const int ppb = blockDim.x / C8;          // pixels per block iteration
  const int pl = threadIdx.x / C8;
  // the filter [C, 9] goes through shared memory: read coalesced once per block (the per-thread gather of 72 taps at a
  // 36-byte stride touched 32 sectors per load instruction and cost more than the pixel loop; with ~1200 short-lived
  // blocks it was paid 8 times per SM -- the kernel ran at 60 us for 50 MB of traffic)
  extern __shared__ float s_w[];
  for (int i = threadIdx.x; i < C8 * 72; i += blockDim.x) s_w[i] = __ldg(w + i);
  __syncthreads();
  if (pl >= ppb) return;
  float wt[9][8], bs[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int tap = flip ? 8 - t : t;
#pragma unroll
    for (int j = 0; j < 8; ++j) wt[t][j] = s_w[(c * 8 + j) * 9 + tap];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = bias ? __ldg(bias + c * 8 + j) : 0.f;
  const long long P = (long long)NB * H * W;
  for (long long p = (long long)blockIdx.x * ppb + pl; p < P; p += (long long)gridDim.x * ppb) {
    const int x = int(p % W);
    const int y = int((p / W) % H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bs[j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y + ky - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x + kx - 1;
        if (ix < 0 || ix >= W) continue;
        float f[8];
        unpack8(__ldg(in + (p + (long long)(ky - 1) * W + (kx - 1)) * C8 + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wt[ky * 3 + kx][j], acc[j]);
      }
    }
    if (add != nullptr) {
      float f[8];
      unpack8(__ldg(add + p * C8 + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    store8(out + (p * C8 + c) * 8, acc);
  }
}

// Tiled versions (C % 64 == 0): a block owns (image, slab of 64 channels, band of rows), stages the band + one halo row
// above / below in shared memory with 16-byte loads and computes from there, so every input element is read from global
// memory ONCE.  The untiled kernels above re-read each element 9 times from L2 (151 MB instead of 17 MB for the
// hourglass layer) and ran at 50-70 us, 5-7x their HBM time (tools/op_bench.py).
// thread = (channel group g = tid % 8, pixel lane = tid / 8); smem tile [rows + 2][W][8] uint4, conflict free.
constexpr int kDwSlab = 8;   // channel groups (of 8 channels) per block
template <typename OutT>
__global__ void __launch_bounds__(256) dwconv3x3_tiled_kernel(const uint4* __restrict__ in, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const uint4* __restrict__ add,
                                                              OutT* __restrict__ out, int H, int W, int C8, int flip, int TH) {
  pdl_grid_sync();
  extern __shared__ uint4 s_tile[];                                   // [(TH + 2) * W * 8]
  float* s_wt = reinterpret_cast<float*>(s_tile + (TH + 2) * W * kDwSlab);   // [64 * 9]
  const int y0 = blockIdx.x * TH, slab = blockIdx.y, b = blockIdx.z;
  const int rows = min(TH, H - y0);
  const int g = threadIdx.x % kDwSlab, lane = threadIdx.x / kDwSlab;
  const int cg = slab * kDwSlab + g;                                  // channel group of this thread
  for (int i = threadIdx.x; i < kDwSlab * 72; i += 256) s_wt[i] = __ldg(w + slab * kDwSlab * 72 + i);
  const uint4* src = in + (long long)b * H * W * C8;
  for (int i = threadIdx.x; i < (rows + 2) * W * kDwSlab; i += 256) {
    const int gg = i % kDwSlab, px = i / kDwSlab;
    const int ty = px / W, x = px - ty * W;
    const int y = y0 + ty - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H) v = __ldg(src + ((long long)y * W + x) * C8 + slab * kDwSlab + gg);
    s_tile[i] = v;
  }
  __syncthreads();
  float wt[9][8], bs[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int tap = flip ? 8 - t : t;
#pragma unroll
    for (int j = 0; j < 8; ++j) wt[t][j] = s_wt[(g * 8 + j) * 9 + tap];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = bias ? __ldg(bias + cg * 8 + j) : 0.f;
  for (int px = lane; px < rows * W; px += 256 / kDwSlab) {
    const int ty = px / W, x = px - ty * W;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bs[j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x + kx - 1;
        if (ix < 0 || ix >= W) continue;       // rows outside the image are zero in the tile
        float f[8];
        unpack8(s_tile[((ty + ky) * W + ix) * kDwSlab + g], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wt[ky * 3 + kx][j], acc[j]);
      }
    const long long p = ((long long)b * H + y0 + ty) * W + x;
    if (add != nullptr) {
      float f[8];
      unpack8(__ldg(add + p * C8 + cg), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    store8(out + (p * C8 + cg) * 8, acc);
  }
}

// weight gradient, same tiling; 72 accumulators per thread, reduced over the 32 pixel lanes (shuffles inside a warp, shared
// memory across the 8 warps), then one fp32 atomic per (channel, tap) and block.
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_tiled_kernel(const uint4* __restrict__ in, const uint4* __restrict__ dout,
                                                                    float* __restrict__ dw, int H, int W, int C8, int TH) {
  pdl_grid_sync();
  extern __shared__ uint4 s_tile[];                                   // in: [(TH + 2) * W * 8], then dout: [TH * W * 8]
  uint4* s_d = s_tile + (TH + 2) * W * kDwSlab;
  const int y0 = blockIdx.x * TH, slab = blockIdx.y, b = blockIdx.z;
  const int rows = min(TH, H - y0);
  const int g = threadIdx.x % kDwSlab, lane = threadIdx.x / kDwSlab;
  const uint4* src = in + (long long)b * H * W * C8;
  const uint4* dsrc = dout + (long long)b * H * W * C8;
  for (int i = threadIdx.x; i < (rows + 2) * W * kDwSlab; i += 256) {
    const int gg = i % kDwSlab, px = i / kDwSlab;
    const int ty = px / W, x = px - ty * W;
    const int y = y0 + ty - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H) v = __ldg(src + ((long long)y * W + x) * C8 + slab * kDwSlab + gg);
    s_tile[i] = v;
  }
  for (int i = threadIdx.x; i < rows * W * kDwSlab; i += 256) {
    const int gg = i % kDwSlab, px = i / kDwSlab;
    s_d[i] = __ldg(dsrc + ((long long)y0 * W + px) * C8 + slab * kDwSlab + gg);
  }
  __syncthreads();
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  for (int px = lane; px < rows * W; px += 256 / kDwSlab) {
    const int ty = px / W, x = px - ty * W;
    float gd[8];
    unpack8(s_d[px * kDwSlab + g], gd);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x + kx - 1;
        if (ix < 0 || ix >= W) continue;
        float f[8];
        unpack8(s_tile[((ty + ky) * W + ix) * kDwSlab + g], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[ky * 3 + kx][j] = fmaf(gd[j], f[j], acc[ky * 3 + kx][j]);
      }
  }
  // lanes of one group inside a warp differ in bits 3 and 4 of the thread index
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], 8);
      acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], 16);
    }
  __syncthreads();                                                    // the tiles are dead: reuse the memory
  float* red = reinterpret_cast<float*>(s_tile);                      // [8 warps][8 groups][72]
  const int wrp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) < kDwSlab) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(wrp * kDwSlab + g) * 72 + j * 9 + t] = acc[t][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kDwSlab * 72; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k * kDwSlab * 72 + i];
    atomicAdd(dw + (long long)slab * kDwSlab * 72 + i, t);   // i = (g * 8 + j) * 9 + tap = the [C, 9] weight layout
  }
}

// depthwise weight gradient: dW[c,tap] += sum_p dRaw[p,c] * in[p+tap,c].  Same thread mapping as the forward:
// 8 channels x 9 taps = 72 fp32 accumulators per thread over its pixels, then one shared-memory reduction over the
// block's pixel lanes and fp32 atomics (grid = a few hundred blocks, so <= a few hundred atomics per address).
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_kernel(const uint4* __restrict__ in, const uint4* __restrict__ dout,
                                                              float* __restrict__ dw, int NB, int H, int W, int C8) {
  pdl_grid_sync();
  const int c = threadIdx.x % C8;
  const int ppb = blockDim.x / C8;
  const int pl = threadIdx.x / C8;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const long long P = (long long)NB * H * W;
  if (pl < ppb) {
    for (long long p = (long long)blockIdx.x * ppb + pl; p < P; p += (long long)gridDim.x * ppb) {
      const int x = int(p % W);
      const int y = int((p / W) % H);
      float g[8];
      unpack8(__ldg(dout + p * C8 + c), g);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = x + kx - 1;
          if (ix < 0 || ix >= W) continue;
          float f[8];
          unpack8(__ldg(in + (p + (long long)(ky - 1) * W + (kx - 1)) * C8 + c), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[ky * 3 + kx][j] = fmaf(g[j], f[j], acc[ky * 3 + kx][j]);
        }
      }
    }
  }
  // reduce over the pixel lanes of the block: lane 0 of each channel group accumulates through shared memory
  extern __shared__ float red[];   // [ppb - 1][C8][72]
  if (pl > 0 && pl < ppb) {
    float* r = red + ((long long)(pl - 1) * C8 + c) * 72;
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) r[t * 8 + j] = acc[t][j];
  }
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < ppb; ++l) {
      const float* r = red + ((long long)(l - 1) * C8 + c) * 72;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] += r[t * 8 + j];
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dw + (c * 8 + j) * 9 + t, acc[t][j]);
  }
}

// ------------------------------------------------------------------------------------------------
// Train-mode BatchNorm over channels-last [P, C] tensors.  The pre-BN conv output ("raw") may be stored
// in fp32 (training: bf16 rounding of raw is amplified by |mean|/std through the normalisation) or bf16.
// Thread mapping for all four kernels: a thread owns ONE group of 8 consecutive channels for the whole
// kernel (per-channel parameters live in registers) and strides over rows; 16/32-byte vector accesses.
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
}
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) { *reinterpret_cast<uint4*>(p) = pack8(f); }
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

constexpr int kBnThreads = 256;
constexpr int kBnBwdReplicas = 8;   // == DP_BN_BWD_REPLICAS (include/dinopose.h)

// block-level reduction of per-thread partial sums that share a channel group, then fp64 atomics
__device__ __forceinline__ void bn_block_reduce_atomic(float (&s)[8], float (&q)[8], int c8, int C8, int C,
                                                       double* __restrict__ sums) {
  __shared__ float red[kBnThreads][17];
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[tid][j] = s[j];
    red[tid][8 + j] = q[j];
  }
  __syncthreads();
  if (tid < C8) {
    float ts[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) ts[j] = 0.f;
    for (int t = tid; t < kBnThreads; t += C8)
#pragma unroll
      for (int j = 0; j < 16; ++j) ts[j] += red[t][j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(sums + c8 * 8 + j, double(ts[j]));
      atomicAdd(sums + C + c8 * 8 + j, double(ts[8 + j]));
    }
  }
}

// The ~600 blocks spread their fp64 atomics over the kBnBwdReplicas copies of the accumulator that the backward uses
// anyway (one copy = 600 dependent read-modify-writes per address in L2, ~20 us of a 40 us launch); the last block to
// finish (ticket counter shared with the fused finalize kernels) folds the copies into the first and re-zeroes the
// others, so dp_bn_finalize / dp_bn_finalize_apply see the [2*C] layout they always did.
template <typename RawT>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const RawT* __restrict__ raw, double* __restrict__ sums,
                                                              unsigned int* __restrict__ counter, long long P, int C8) {
  pdl_grid_sync();
  const int C = C8 * 8;
  const int c8 = threadIdx.x % C8;
  const int rpb = kBnThreads / C8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  // two rows per iteration, both loads issued before the math
  const long long stride = (long long)gridDim.x * rpb;
  for (long long row = (long long)blockIdx.x * rpb + threadIdx.x / C8; row < P; row += 2 * stride) {
    float f0[8], f1[8];
    const bool two = row + stride < P;
    load8(raw + row * C + c8 * 8, f0);
    load8(raw + (two ? row + stride : row) * C + c8 * 8, f1);
    const float w1 = two ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b = f1[j] * w1;
      s[j] += f0[j] + b;
      q[j] = fmaf(f0[j], f0[j], fmaf(b, f1[j], q[j]));
    }
  }
  bn_block_reduce_atomic(s, q, c8, C8, C, sums + (long long)(blockIdx.x % kBnBwdReplicas) * 2 * C);
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence();
    // ALL loads of up to four channels per thread first, then the stores: with the stores in between, the compiler must
    // keep the 8 x 4 L2 round trips of this one block in order (ncu: the launch took 20 us whatever the tensor size)
    for (int c0 = threadIdx.x; c0 < 2 * C; c0 += 4 * kBnThreads) {
      double v[4][kBnBwdReplicas];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + i * kBnThreads;
#pragma unroll
        for (int r = 0; r < kBnBwdReplicas; ++r) v[i][r] = c < 2 * C ? __ldcg(sums + (long long)r * 2 * C + c) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + i * kBnThreads;
        if (c >= 2 * C) break;
        double a = v[i][0];
#pragma unroll
        for (int r = 1; r < kBnBwdReplicas; ++r) {
          a += v[i][r];
          sums[(long long)r * 2 * C + c] = 0.0;
        }
        sums[c] = a;
      }
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// finalize: batch mean / biased var -> scale, shift (y = raw*scale + shift), saved mean / invstd,
// running statistics update with momentum (unbiased variance), then zero `sums` for the next step.
__global__ void bn_finalize_kernel(double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int C, double count, float eps, float momentum) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0) var = 0;
  const float invstd = float(1.0 / sqrt(var + double(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - float(mean) * sc;
  mean_out[c] = float(mean);
  invstd_out[c] = invstd;
  if (running_mean != nullptr) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * float(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * float(unbiased);
  }
  sums[c] = 0.0;
  sums[C + c] = 0.0;
}

// eval-mode fold: scale = gamma / sqrt(running_var + eps); shift = beta + (conv_bias - running_mean) * scale
__global__ void bn_fold_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv,
                                    const float* __restrict__ conv_bias, float* __restrict__ scale,
                                    float* __restrict__ shift, float* __restrict__ mean_out,
                                    float* __restrict__ invstd_out, int C, float eps) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float is = rsqrtf(rv[c] + eps);
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - rm[c]) * sc;
  // training step with the heads in eval mode (frozen statistics): the backward reads these like saved batch statistics
  if (mean_out != nullptr) mean_out[c] = rm[c];
  if (invstd_out != nullptr) invstd_out[c] = is;
}

// apply: y = raw*scale + shift;  mode 0: out = relu?(y) + add1 + add2      (HourglassModule 3-way sum, :285)
//                                 mode 1: out = relu(y + add1)              (bottleneck residual, :277-278)
template <typename RawT>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const RawT* __restrict__ raw, const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const __nv_bfloat16* __restrict__ add1,
                                                              const __nv_bfloat16* __restrict__ add2,
                                                              __nv_bfloat16* __restrict__ out, long long P, int C8, int relu,
                                                              int mode) {
  pdl_grid_sync();
  const int C = C8 * 8;
  const int c8 = threadIdx.x % C8;
  const int rpb = kBnThreads / C8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + c8 * 8 + j);
    sh[j] = __ldg(shift + c8 * 8 + j);
  }
  for (long long row = (long long)blockIdx.x * rpb + threadIdx.x / C8; row < P; row += (long long)gridDim.x * rpb) {
    const long long o = row * C + c8 * 8;
    float f[8], a[8];
    load8(raw + o, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
    if (mode == 1) {
      load8(add1 + o, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j] + a[j], 0.f);
    } else {
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (add1 != nullptr) {
        load8(add1 + o, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += a[j];
      }
      if (add2 != nullptr) {
        load8(add2 + o, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += a[j];
      }
    }
    store8(out + o, f);
  }
}

// Fused bn_finalize + bn_apply (train mode): every block derives scale / shift of all C channels from the fp64 sums
// itself (2*C doubles = 8 KB at C = 512, an L2 hit) instead of waiting for a C-thread kernel in between: 14 launches of
// ~4 us less per step.  Block 0 also writes scale / shift / mean / invstd for the backward and updates the running
// statistics; the LAST block to finish (ticket counter) zeroes the sums for the next step -- every block has read
// them before it takes its ticket.
constexpr int kBnMaxC = 512;
template <typename RawT>
__global__ void __launch_bounds__(kBnThreads) bn_finalize_apply_kernel(
    const RawT* __restrict__ raw, double* __restrict__ sums, unsigned int* __restrict__ counter,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
    float* __restrict__ running_var, float* __restrict__ scale_out, float* __restrict__ shift_out,
    float* __restrict__ mean_out, float* __restrict__ invstd_out, const __nv_bfloat16* __restrict__ add1,
    const __nv_bfloat16* __restrict__ add2, __nv_bfloat16* __restrict__ out, long long P, int C8, int relu, int mode,
    double count, float eps, float momentum) {
  pdl_grid_sync();
  __shared__ float s_sc[kBnMaxC], s_sh[kBnMaxC];
  __shared__ bool s_last;
  const int C = C8 * 8;
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const double mean = __ldcg(sums + c) / count;
    double var = __ldcg(sums + C + c) / count - mean * mean;
    if (var < 0) var = 0;
    const float invstd = float(1.0 / sqrt(var + double(eps)));
    const float sc = __ldg(gamma + c) * invstd;
    const float sh = __ldg(beta + c) - float(mean) * sc;
    s_sc[c] = sc;
    s_sh[c] = sh;
    if (blockIdx.x == 0) {
      scale_out[c] = sc;
      shift_out[c] = sh;
      mean_out[c] = float(mean);
      invstd_out[c] = invstd;
      if (running_mean != nullptr) {
        const double unbiased = count > 1 ? var * count / (count - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * float(mean);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * float(unbiased);
      }
    }
  }
  __syncthreads();
  const int c8 = threadIdx.x % C8;
  const int rpb = kBnThreads / C8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_sc[c8 * 8 + j];
    sh[j] = s_sh[c8 * 8 + j];
  }
  for (long long row = (long long)blockIdx.x * rpb + threadIdx.x / C8; row < P; row += (long long)gridDim.x * rpb) {
    const long long o = row * C + c8 * 8;
    float f[8], a[8];
    load8(raw + o, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
    if (mode == 1) {
      load8(add1 + o, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j] + a[j], 0.f);
    } else {
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (add1 != nullptr) {
        load8(add1 + o, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += a[j];
      }
      if (add2 != nullptr) {
        load8(add2 + o, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += a[j];
      }
    }
    store8(out + o, f);
  }
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    for (int c = threadIdx.x; c < 2 * C; c += kBnThreads) sums[c] = 0.0;
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// ---- 4-channel helpers for the BatchNorm backward kernels (a thread owns 4 channels: ~50 registers, 5 blocks / SM)
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&f)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&f)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
  const __nv_bfloat162 h1 = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
  f[0] = __low2float(h0); f[1] = __high2float(h0); f[2] = __low2float(h1); f[3] = __high2float(h1);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&f)[4]) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&h0);
  t.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(p) = t;
}

// BatchNorm backward, pass 1: with y = raw*scale+shift,
//   dy = dout * [y > 0]            (mode 0, relu)      | dy = dout * [y + add1 > 0]   (mode 1)
//   sums[c] += dy ; sums[C + c] += dy * raw
// (sum dy*xhat = invstd * (sum dy*raw - mean * sum dy) is formed in fp64 by pass 2: the loop needs no mean / invstd)
template <typename RawT>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ dout, const RawT* __restrict__ raw, const __nv_bfloat16* __restrict__ add1,
    const float* __restrict__ scale, const float* __restrict__ shift, double* __restrict__ sums, long long P, int C4,
    int relu, int mode) {
  pdl_grid_sync();
  const int C = C4 * 4;
  const int c4 = threadIdx.x % C4;
  const int rpb = kBnThreads / C4;
  float sc[4], sh[4], s[4], q[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = __ldg(scale + c4 * 4 + j);
    sh[j] = __ldg(shift + c4 * 4 + j);
    s[j] = q[j] = 0.f;
  }
  const bool masked = relu || mode == 1;
  // two rows per iteration: all loads of both rows are issued before the math
  const long long stride = (long long)gridDim.x * rpb;
  for (long long row = (long long)blockIdx.x * rpb + threadIdx.x / C4; row < P; row += 2 * stride) {
    const long long o0 = row * C + c4 * 4;
    const bool two = row + stride < P;
    const long long o1 = two ? (row + stride) * C + c4 * 4 : o0;
    float r0[4], g0[4], a0[4], r1[4], g1[4], a1[4];
    load4(raw + o0, r0);
    load4(dout + o0, g0);
    load4(raw + o1, r1);
    load4(dout + o1, g1);
    if (mode == 1) {
      load4(add1 + o0, a0);
      load4(add1 + o1, a1);
    }
    const float w1 = two ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y0 = fmaf(r0[j], sc[j], sh[j]), y1 = fmaf(r1[j], sc[j], sh[j]);
      if (mode == 1) { y0 += a0[j]; y1 += a1[j]; }
      const float gj0 = (masked && !(y0 > 0.f)) ? 0.f : g0[j];
      const float gj1 = (masked && !(y1 > 0.f)) ? 0.f : g1[j] * w1;
      s[j] += gj0 + gj1;
      q[j] = fmaf(gj0, r0[j], fmaf(gj1, r1[j], q[j]));
    }
  }
  // block reduction over the threads that share a channel group, then fp64 atomics
  __shared__ float red[kBnThreads][9];
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[tid][j] = s[j];
    red[tid][4 + j] = q[j];
  }
  __syncthreads();
  if (tid < C4) {
    float ts[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ts[j] = 0.f;
    for (int t = tid; t < kBnThreads; t += C4)
#pragma unroll
      for (int j = 0; j < 8; ++j) ts[j] += red[t][j];
    // kBnBwdReplicas copies of the accumulators: ~700 blocks adding to the same 2*C addresses serialise in L2
    double* dst = sums + (long long)(blockIdx.x % kBnBwdReplicas) * 2 * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(dst + c4 * 4 + j, double(ts[j]));
      atomicAdd(dst + C + c4 * 4 + j, double(ts[4 + j]));
    }
  }
}

// between the passes (C threads): add the replicas up, form the three per-channel coefficients of pass 2 in fp64,
// emit dgamma / dbeta, and re-zero the accumulators for the next step.  coef = [A | B | K] fp32, 3*C.
__global__ void bn_bwd_coeffs_kernel(double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ scale,
                                     const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ coef,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int C, double invP, int eval_mode) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double S1 = 0.0, Q = 0.0;
#pragma unroll
  for (int rpl = 0; rpl < kBnBwdReplicas; ++rpl) {
    S1 += sums[(long long)rpl * 2 * C + c];
    Q += sums[(long long)rpl * 2 * C + C + c];
    sums[(long long)rpl * 2 * C + c] = 0.0;
    sums[(long long)rpl * 2 * C + C + c] = 0.0;
  }
  if (eval_mode) {
    // 1: identity pass (element-wise adds routed through this kernel); 2: BatchNorm with frozen (running) statistics --
    // the input gradient is dy * scale, gamma's gradient is sum dy * xhat with xhat from the running statistics
    coef[c] = scale[c];
    coef[C + c] = 0.f;
    coef[2 * C + c] = 0.f;
    if (dgamma != nullptr) {
      dgamma[c] = eval_mode == 2 ? float(double(invstd[c]) * (Q - double(mean[c]) * S1)) : 0.f;
      dbeta[c] = float(S1);
    }
    return;
  }
  const double mu = double(mean[c]), is = double(invstd[c]);
  const double S2 = is * (Q - mu * S1);
  const double A = double(gamma[c]) * is, k2 = is * S2 * invP;
  coef[c] = float(A);
  coef[C + c] = float(-A * k2);
  coef[2 * C + c] = float(A * (k2 * mu - S1 * invP));
  if (dgamma != nullptr) { dgamma[c] = float(S2); dbeta[c] = float(S1); }
}

// pass 2: draw = gamma*invstd * (dy - S1/P - xhat*S2/P)  (train)   |   draw = dy * scale (eval_mode)
// with S1 = sum dy, S2 = sum dy*xhat = invstd * (sums[C+c] - mean*S1).  Per channel this is  A*dy + B*raw + K
// (A = gamma*invstd, B = -A*invstd*S2/P, K = A*(invstd*S2/P*mean - S1/P)), so the row loop needs 5 coefficients.
// Also emits dgamma = S2, dbeta = S1 (block 0) and, for mode 1, the masked gradient of the residual branch.
// shuffle_oh > 0: write draw in the un-shuffled [P/4, 4*Cout] "col" layout of a k2 s2 transposed conv
// (row = input pixel, column = tap*Cout + co), the operand layout of its weight / input gradients.
// FUSED: the coefficient step of bn_bwd_coeffs_kernel runs inside this kernel -- every block adds the replicas up for
// all C channels (16 doubles per channel from L2) and keeps A / B / K in shared memory; block 0 emits dgamma / dbeta;
// the last block to finish (ticket counter) re-zeroes the accumulators.  Saves one ~4 us launch per BatchNorm layer.
struct BnBwdFused {
  double* sums;
  unsigned int* counter;
  const float* gamma;
  const float* mean;
  const float* invstd;
  float* dgamma;
  float* dbeta;
  double invP;
  int eval_mode;
};
template <typename RawT, bool FUSED>
__global__ void __launch_bounds__(kBnThreads, 4) bn_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ dout, const RawT* __restrict__ raw, const __nv_bfloat16* __restrict__ add1,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ coef,
    __nv_bfloat16* __restrict__ draw, __nv_bfloat16* __restrict__ dres, long long P, int C4, int relu, int mode,
    int shuffle_oh, int shuffle_ow, const BnBwdFused fz) {
  pdl_grid_sync();
  const int C = C4 * 4;
  const int c4 = threadIdx.x % C4;
  const int rpb = kBnThreads / C4;
  __shared__ float s_coef[FUSED ? 3 * kBnMaxC : 1];
  __shared__ bool s_last;
  if constexpr (FUSED) {
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
      double S1 = 0.0, Q = 0.0;
#pragma unroll
      for (int rpl = 0; rpl < kBnBwdReplicas; ++rpl) {
        S1 += __ldcg(fz.sums + (long long)rpl * 2 * C + c);
        Q += __ldcg(fz.sums + (long long)rpl * 2 * C + C + c);
      }
      float A, B, K, dg;
      if (fz.eval_mode) {
        A = __ldg(scale + c); B = 0.f; K = 0.f;
        dg = fz.eval_mode == 2 ? float(double(__ldg(fz.invstd + c)) * (Q - double(__ldg(fz.mean + c)) * S1)) : 0.f;
      } else {
        const double mu = double(__ldg(fz.mean + c)), is = double(__ldg(fz.invstd + c));
        const double S2 = is * (Q - mu * S1);
        const double Ad = double(__ldg(fz.gamma + c)) * is, k2 = is * S2 * fz.invP;
        A = float(Ad); B = float(-Ad * k2); K = float(Ad * (k2 * mu - S1 * fz.invP)); dg = float(S2);
      }
      s_coef[c] = A; s_coef[kBnMaxC + c] = B; s_coef[2 * kBnMaxC + c] = K;
      if (blockIdx.x == 0 && fz.dgamma != nullptr) { fz.dgamma[c] = dg; fz.dbeta[c] = float(S1); }
    }
    __syncthreads();
  }
  float sc[4], sh[4], ka[4], kb[4], kc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c4 * 4 + j;
    sc[j] = __ldg(scale + c);
    sh[j] = __ldg(shift + c);
    if constexpr (FUSED) {
      ka[j] = s_coef[c]; kb[j] = s_coef[kBnMaxC + c]; kc[j] = s_coef[2 * kBnMaxC + c];
    } else {
      ka[j] = __ldg(coef + c);
      kb[j] = __ldg(coef + C + c);
      kc[j] = __ldg(coef + 2 * C + c);
    }
  }
  const bool masked = relu || mode == 1;
  const long long stride = (long long)gridDim.x * rpb;
  for (long long row = (long long)blockIdx.x * rpb + threadIdx.x / C4; row < P; row += 2 * stride) {
    const bool two = row + stride < P;
    const long long rows[2] = {row, two ? row + stride : row};
    float r[2][4], g[2][4], a[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long o = rows[u] * C + c4 * 4;
      load4(raw + o, r[u]);
      load4(dout + o, g[u]);
      if (mode == 1) load4(add1 + o, a[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const long long o = rows[u] * C + c4 * 4;
      float ov[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y = fmaf(r[u][j], sc[j], sh[j]);
        if (mode == 1) y += a[u][j];
        if (masked && !(y > 0.f)) g[u][j] = 0.f;
        ov[j] = fmaf(ka[j], g[u][j], fmaf(kb[j], r[u][j], kc[j]));
      }
      long long oo = o;
      if (shuffle_oh > 0) {
        // row indexes the shuffled NHWC output [NB, 2*ih, 2*iw, C]; destination is [NB*ih*iw, 4*C]
        long long p = rows[u];
        const int x = int(p % shuffle_ow); p /= shuffle_ow;
        const int y = int(p % shuffle_oh);
        const long long b = p / shuffle_oh;
        const int ih = shuffle_oh / 2, iw = shuffle_ow / 2;
        const int tap = (y & 1) * 2 + (x & 1);
        oo = ((((b * ih + (y >> 1)) * iw + (x >> 1)) * 4 + tap) * C4 + c4) * 4;
      }
      store4(draw + oo, ov);
      if (dres != nullptr) store4(dres + o, g[u]);
    }
  }
  if constexpr (FUSED) {
    if (threadIdx.x == 0) s_last = atomicAdd(fz.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
      for (int c = threadIdx.x; c < kBnBwdReplicas * 2 * C; c += kBnThreads) fz.sums[c] = 0.0;
      if (threadIdx.x == 0) *fz.counter = 0u;
    }
  }
}

__global__ void zero_f64_kernel(double* p, int n) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// 2x2 mean (bilinear align_corners=False at exact scale 1/2, pose_heads.py:353-359) fp32 NCHW; and its adjoint.
__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int OH, int OW) {
  pdl_grid_sync();
  const long long total = planes * OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % OW);
    const int y = int((i / OW) % OH);
    const long long pl = i / ((long long)OW * OH);
    const float* s = in + (pl * 2 * OH + 2 * y) * (2 * OW) + 2 * x;
    // same association as the separable bilinear kernel: horizontal lerp then vertical lerp, weights 0.5
    const float top = 0.5f * s[0] + 0.5f * s[1];
    const float bot = 0.5f * s[2 * OW] + 0.5f * s[2 * OW + 1];
    out[i] = 0.5f * top + 0.5f * bot;
  }
}

// heat-map gradient fp32 NCHW [NB,K,GH,GW] -> bf16 NHWC [NB*OH*OW, Kp] (zero padded channels);
// up = 2 additionally applies the adjoint of avgpool2 (each fine pixel gets 0.25 * coarse gradient).
// A block transposes 128 consecutive pixels of one image through shared memory: NCHW reads are coalesced along the
// pixels of each channel plane (4 loads in flight per lane), NHWC writes are 16-byte pieces of the padded channel rows.
constexpr int kHmPix = 128;
__global__ void __launch_bounds__(256) hm_grad_to_nhwc_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ out,
                                                              int NB, int K, int Kp, int OH, int OW, int up) {
  pdl_grid_sync();
  __shared__ float tile[64][kHmPix + 1];
  const int GH = OH / up, GW = OW / up;
  const float w = up == 2 ? 0.25f : 1.0f;
  const int hw = OH * OW;
  const int chunks = (hw + kHmPix - 1) / kHmPix;
  const int b = blockIdx.x / chunks, p0 = (blockIdx.x % chunks) * kHmPix;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int k = wrp; k < Kp; k += 8) {
    float v[kHmPix / 32];
#pragma unroll
    for (int i = 0; i < kHmPix / 32; ++i) {
      const int pix = p0 + lane + 32 * i;
      v[i] = 0.f;
      if (k < K && pix < hw) {
        const int y = pix / OW, x = pix - y * OW;
        v[i] = w * __ldg(g + (((long long)b * K + k) * GH + y / up) * GW + x / up);
      }
    }
#pragma unroll
    for (int i = 0; i < kHmPix / 32; ++i) tile[k][lane + 32 * i] = v[i];
  }
  __syncthreads();
  const int kp8 = Kp / 8;   // Kp % 8 == 0 (checked by the launcher)
  for (int i = threadIdx.x; i < kHmPix * kp8; i += 256) {
    const int pl = i / kp8, k8 = i - pl * kp8;
    if (p0 + pl < hw) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = tile[k8 * 8 + j][pl];
      store8(out + ((long long)b * hw + p0 + pl) * Kp + k8 * 8, f);
    }
  }
}

// mean over the N patch tokens: feat bf16 [B, N, D] -> fp32 [B, D]   (pose_heads.py:397).  One block per image:
// thread = (8-channel group, token lane), 16-byte loads, token lanes combined through shared memory.
__global__ void __launch_bounds__(512) mean_tokens_kernel(const __nv_bfloat16* __restrict__ feat, float* __restrict__ out, int N, int D) {
  pdl_grid_sync();
  extern __shared__ float s_part[];   // [lanes][D]
  const int b = blockIdx.x;
  const int C8 = D / 8;
  const int lanes = blockDim.x / C8;
  const int c = threadIdx.x % C8, l = threadIdx.x / C8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (l < lanes) {
    const uint4* src = reinterpret_cast<const uint4*>(feat + (long long)b * N * D) + c;
    for (int n = l; n < N; n += lanes) {
      float f[8];
      unpack8(__ldg(src + (long long)n * C8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[l * D + c * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < lanes; ++k) t += s_part[k * D + d];
    out[(long long)b * D + d] = t / float(N);
  }
}
// adjoint: dfeat[b, n, :] += dmean[b, :] / N   (bf16 in place), 8 channels per thread
__global__ void __launch_bounds__(256) mean_tokens_bwd_kernel(__nv_bfloat16* __restrict__ dfeat, const float* __restrict__ dmean,
                                                              int N, int D, long long total) {
  pdl_grid_sync();
  const int C8 = D / 8;
  const long long total8 = total / 8;
  const float inv = 1.f / float(N);
  uint4* p = reinterpret_cast<uint4*>(dfeat);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C8);
    const long long b = i / ((long long)N * C8);
    float f[8];
    unpack8(p[i], f);
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(dmean + b * D + c * 8));
    const float4 m1 = __ldg(reinterpret_cast<const float4*>(dmean + b * D + c * 8) + 1);
    // same rounding as before: dmean / N (a division), added in fp32, one bf16 rounding
    f[0] += m0.x / float(N); f[1] += m0.y / float(N); f[2] += m0.z / float(N); f[3] += m0.w / float(N);
    f[4] += m1.x / float(N); f[5] += m1.y / float(N); f[6] += m1.z / float(N); f[7] += m1.w / float(N);
    (void)inv;
    p[i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------
// Small fp32 GEMM for the z-head MLP (M = batch): C[m,n] = sum_k A(m,k) * B(k,n), generic strides,
// 64x64x16 shared-memory tiles, 256 threads x (4x4).  Epilogue: + bias[n]; relu; dropout (forward);
// or multiply by the saved forward activation's mask (backward): C *= (ref[m,n] > 0) [* dropout mask].
__device__ __forceinline__ uint32_t mix32h(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return uint32_t((z ^ (z >> 31)) >> 16);
}

__global__ void __launch_bounds__(128) sgemm_small_kernel(const float* __restrict__ A, long long sa_m, long long sa_k,
                                                          const float* __restrict__ Bm, long long sb_k, long long sb_n,
                                                          float* __restrict__ Cc, long long ldc, int M, int N, int K,
                                                          const float* __restrict__ bias, int relu,
                                                          const float* __restrict__ mask_ref, long long ld_ref,
                                                          float p_drop, const unsigned long long* __restrict__ seed_ptr,
                                                          int accumulate, int k_per_split) {
  pdl_grid_sync();
  // 32x32 output tile, 32-deep k-steps, 128 threads x (2 rows x 4 cols); loads are coalesced along whichever
  // operand dimension has unit stride.  gridDim.z > 1: split-K -- each z-slice adds its partial sums to C with
  // fp32 atomics (C pre-zeroed or accumulated onto) and the epilogue runs in sgemm_small_finish_kernel.
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  const bool split = gridDim.z > 1;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (sa_k == 1), b_kfast = (sb_k == 1);
  for (int k0 = k_begin; k0 < k_end; k0 += 32) {
#pragma unroll
    for (int i = threadIdx.x; i < 32 * 32; i += 128) {
      const int lo = i & 31, hi = i >> 5;
      {
        const int kk = a_kfast ? lo : hi, mm = a_kfast ? hi : lo;
        const int m = m0 + mm, k = k0 + kk;
        As[kk][mm] = (m < M && k < k_end) ? __ldg(A + m * sa_m + k * sa_k) : 0.f;
      }
      {
        const int kk = b_kfast ? lo : hi, nn = b_kfast ? hi : lo;
        const int n = n0 + nn, k = k0 + kk;
        Bs[kk][nn] = (n < N && k < k_end) ? __ldg(Bm + k * sb_k + n * sb_n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
      float b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[0][j] += a0 * b[j];
        acc[1][j] += a1 * b[j];
      }
    }
    __syncthreads();
  }
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty * 2 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (split) {
        atomicAdd(Cc + m * ldc + n, v);
        continue;
      }
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      if (mask_ref) v = mask_ref[m * ld_ref + n] > 0.f ? v : 0.f;
      if (p_drop > 0.f) v = (mix32h(seed * 0xD1342543DE82EF95ull + uint64_t(m) * N + n) >= thresh) ? v * keep : 0.f;
      if (accumulate) v += Cc[m * ldc + n];
      Cc[m * ldc + n] = v;
    }
  }
}

// epilogue of the split-K path, in place on C
__global__ void sgemm_small_finish_kernel(float* __restrict__ Cc, long long ldc, int M, int N, const float* __restrict__ bias,
                                          int relu, const float* __restrict__ mask_ref, long long ld_ref, float p_drop,
                                          const unsigned long long* __restrict__ seed_ptr) {
  pdl_grid_sync();
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = int(i / N), n = int(i % N);
  float v = Cc[m * ldc + n];
  if (bias) v += bias[n];
  if (relu) v = fmaxf(v, 0.f);
  if (mask_ref) v = mask_ref[m * ld_ref + n] > 0.f ? v : 0.f;
  if (p_drop > 0.f) v = (mix32h(seed * 0xD1342543DE82EF95ull + uint64_t(m) * N + n) >= thresh) ? v * keep : 0.f;
  Cc[m * ldc + n] = v;
}

// out = (ref > 0) ? d * keep_scale : 0      (gradient through ReLU [+ inverted dropout] given the saved output)
__global__ void relu_mask_kernel(const float* __restrict__ d, const float* __restrict__ ref, float* __restrict__ out,
                                 long long n, float keep_scale) {
  pdl_grid_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ref[i] > 0.f ? d[i] * keep_scale : 0.f;
}

// column sums: out[c] (+)= sum_p x[p, c]   (bias gradients).  x fp32 or bf16.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long P, int C, long long ld,
                              int rows_per_block) {
  pdl_grid_sync();
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long p0 = (long long)blockIdx.x * rows_per_block;
  const long long p1 = min(p0 + rows_per_block, P);
  float s = 0.f;
  for (long long p = p0; p < p1; ++p) {
    if constexpr (sizeof(T) == 2) s += __bfloat162float(x[p * ld + c]);
    else s += x[p * ld + c];
  }
  atomicAdd(out + c, s);
}

}  // namespace

// ----------------------------------------------------------------------------------------------- launchers
cudaError_t launch_im2col(const void* in, void* col, int NB, int IH, int IW, int C, int OH, int OW, int KH, int KW,
                          int stride, int pad, cudaStream_t s) {
  const long long total = (long long)NB * OH * OW * KH * KW * (C / 8);
  launch_k<im2col_kernel>(grid_for(total, 256), 256, 0, s, reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(col), NB,
                                                     IH, IW, C / 8, OH, OW, KH, KW, stride, pad);
  return cudaGetLastError();
}
cudaError_t launch_col2im(const void* col, const float* bias, void* big, int big_f32, int NB, int SH, int SW, int C, int BH,
                          int BW, int KH, int KW, int stride, int pad, cudaStream_t s) {
  const long long total = (long long)NB * BH * BW * (C / 8);
  if (big_f32)
    launch_k<col2im_kernel<float>>(grid_for(total, 256), 256, 0, s, reinterpret_cast<const uint4*>(col), bias,
                                                              reinterpret_cast<float*>(big), NB, SH, SW, C / 8, BH, BW, KH,
                                                              KW, stride, pad);
  else
    launch_k<col2im_kernel<__nv_bfloat16>>(grid_for(total, 256), 256, 0, s, reinterpret_cast<const uint4*>(col), bias,
                                                                      reinterpret_cast<__nv_bfloat16*>(big), NB, SH, SW,
                                                                      C / 8, BH, BW, KH, KW, stride, pad);
  return cudaGetLastError();
}
cudaError_t launch_dwconv3x3(const void* in, const float* w, const float* bias, const void* add, void* out, int out_f32,
                             int NB, int H, int W, int C, int flip, cudaStream_t s) {
  const int C8 = C / 8;
  if (C % 8 || C8 > 256) return cudaErrorInvalidValue;
  if (C % 64 == 0 && W * 128 * 3 <= 44 * 1024) {   // tiled kernel: rows + 2 halo rows of W pixels x 128 bytes fit in 44 KB
    int TH = (44 * 1024) / (W * 128) - 2;
    if (TH > H) TH = H;
    // (smaller bands = more blocks were measured SLOWER: 29 us at 16 rows, 35 at 8, 43 at 4 for the hourglass layer)
    const size_t smem = size_t(TH + 2) * W * kDwSlab * sizeof(uint4) + kDwSlab * 72 * sizeof(float);
    const dim3 grid((H + TH - 1) / TH, C8 / kDwSlab, NB);
    if (out_f32)
      launch_k<dwconv3x3_tiled_kernel<float>>(grid, 256, smem, s, reinterpret_cast<const uint4*>(in), w, bias,
                                              reinterpret_cast<const uint4*>(add), reinterpret_cast<float*>(out), H, W, C8, flip, TH);
    else
      launch_k<dwconv3x3_tiled_kernel<__nv_bfloat16>>(grid, 256, smem, s, reinterpret_cast<const uint4*>(in), w, bias,
                                                      reinterpret_cast<const uint4*>(add), reinterpret_cast<__nv_bfloat16*>(out),
                                                      H, W, C8, flip, TH);
    return cudaGetLastError();
  }
  const int ppb = 256 / C8;
  const long long P = (long long)NB * H * W;
  long long g = (P + ppb - 1) / ppb;
  if (g > 148 * 2) g = 148 * 2;   // two resident blocks per SM (~110 registers per thread): one wave, the taps are loaded once
  const size_t smem = size_t(C) * 9 * sizeof(float);   // <= 72 KB... C <= 2048 -> checked below
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  if (out_f32)
    launch_k<dwconv3x3_kernel<float>>(unsigned(g), 256, smem, s, reinterpret_cast<const uint4*>(in), w, bias,
                                                        reinterpret_cast<const uint4*>(add), reinterpret_cast<float*>(out),
                                                        NB, H, W, C8, flip);
  else
    launch_k<dwconv3x3_kernel<__nv_bfloat16>>(unsigned(g), 256, smem, s, reinterpret_cast<const uint4*>(in), w, bias,
                                                                reinterpret_cast<const uint4*>(add),
                                                                reinterpret_cast<__nv_bfloat16*>(out), NB, H, W, C8, flip);
  return cudaGetLastError();
}
cudaError_t launch_dwconv3x3_wgrad(const void* in, const void* dout, float* dw, int NB, int H, int W, int C,
                                   cudaStream_t s) {
  const int C8 = C / 8;
  if (C % 8 || C8 > 256) return cudaErrorInvalidValue;
  if (C % 64 == 0 && W * 128 * 4 <= 46 * 1024) {   // tiled kernel: (rows + 2) input rows + rows gradient rows in <= 46 KB
    int TH = ((46 * 1024) / (W * 128) - 2) / 2;
    if (TH > H) TH = H;
    size_t smem = size_t(2 * TH + 2) * W * kDwSlab * sizeof(uint4);
    if (smem < size_t(8) * kDwSlab * 72 * sizeof(float)) smem = size_t(8) * kDwSlab * 72 * sizeof(float);   // reduction scratch
    const dim3 grid((H + TH - 1) / TH, C8 / kDwSlab, NB);
    launch_k<dwconv3x3_wgrad_tiled_kernel>(grid, 256, smem, s, reinterpret_cast<const uint4*>(in), reinterpret_cast<const uint4*>(dout),
                                           dw, H, W, C8, TH);
    return cudaGetLastError();
  }
  const int ppb = 256 / C8;
  const long long P = (long long)NB * H * W;
  long long g = (P + ppb - 1) / ppb;
  if (g > 148 * 2) g = 148 * 2;
  const size_t smem = size_t(ppb > 1 ? ppb - 1 : 1) * C8 * 72 * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dwconv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  launch_k<dwconv3x3_wgrad_kernel>(unsigned(g), 256, smem, s, reinterpret_cast<const uint4*>(in), reinterpret_cast<const uint4*>(dout),
                                                        dw, NB, H, W, C8);
  return cudaGetLastError();
}
// backward kernels: 4 channels per thread, 2 rows per iteration, grid-stride over the rows.  The grid is ONE wave of
// resident blocks of the kernel actually launched (occupancy query, cached per instantiation): the round-1 cap of 5 blocks
// per SM was 1.25 waves for the reduce kernel (55 registers: 4 resident blocks) and 1.67 for the apply kernel (80
// registers: 3), i.e. a second, mostly empty wave per launch (ncu: 37-43 % / 30 % warps active).
template <typename K>
static int bn_resident_blocks(K kernel) {
  static int blocks = 0;
  if (blocks == 0) {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, kBnThreads, 0) != cudaSuccess || b < 1) b = 3;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    blocks = b * (sms > 0 ? sms : 148);
  }
  return blocks;
}
static unsigned bn_grid4(long long P, int C4, int resident) {
  const int rpb = kBnThreads / C4;
  long long g = (P + 2 * rpb - 1) / (2 * rpb);
  if (g > resident) g = resident;
  return unsigned(g < 1 ? 1 : g);
}
static unsigned bn_grid(long long P, int C8, int resident) {
  const int rpb = kBnThreads / C8;
  long long g = (P + rpb - 1) / rpb;
  if (g > resident) g = resident;      // one wave of resident blocks (see bn_resident_blocks)
  return unsigned(g < 1 ? 1 : g);
}
static bool bn_c_ok(int C) { return C % 8 == 0 && kBnThreads % (C / 8) == 0 && C / 8 <= kBnThreads; }

cudaError_t launch_bn_stats(const void* raw, int raw_f32, double* sums, long long P, int C, cudaStream_t s) {
  if (!bn_c_ok(C)) return cudaErrorInvalidValue;
  // ticket counter: the last float slot of the coefficient block behind the replicas (same slot as the fused kernels)
  unsigned int* counter = reinterpret_cast<unsigned int*>(sums + (long long)kBnBwdReplicas * 2 * C) + 3 * C;
  if (raw_f32)
    launch_k<bn_stats_kernel<float>>(bn_grid(P, C / 8, bn_resident_blocks(bn_stats_kernel<float>)), kBnThreads, 0, s, reinterpret_cast<const float*>(raw), sums, counter, P, C / 8);
  else
    launch_k<bn_stats_kernel<__nv_bfloat16>>(bn_grid(P, C / 8, bn_resident_blocks(bn_stats_kernel<__nv_bfloat16>)), kBnThreads, 0, s, reinterpret_cast<const __nv_bfloat16*>(raw), sums, counter, P, C / 8);
  return cudaGetLastError();
}
cudaError_t launch_bn_finalize(double* sums, const float* gamma, const float* beta, float* rm, float* rv, float* scale,
                               float* shift, float* mean, float* invstd, int C, double count, float eps, float momentum,
                               cudaStream_t s) {
  launch_k<bn_finalize_kernel>((C + 127) / 128, 128, 0, s, sums, gamma, beta, rm, rv, scale, shift, mean, invstd, C, count, eps,
                                                     momentum);
  return cudaGetLastError();
}
cudaError_t launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                                const float* conv_bias, float* scale, float* shift, float* mean_out, float* invstd_out, int C,
                                float eps, cudaStream_t s) {
  launch_k<bn_fold_eval_kernel>((C + 127) / 128, 128, 0, s, gamma, beta, rm, rv, conv_bias, scale, shift, mean_out, invstd_out, C,
                                eps);
  return cudaGetLastError();
}
cudaError_t launch_bn_apply(const void* raw, int raw_f32, const float* scale, const float* shift, const void* add1,
                            const void* add2, void* out, long long P, int C, int relu, int mode, cudaStream_t s) {
  if (!bn_c_ok(C)) return cudaErrorInvalidValue;
  auto a1 = reinterpret_cast<const __nv_bfloat16*>(add1);
  auto a2 = reinterpret_cast<const __nv_bfloat16*>(add2);
  auto o = reinterpret_cast<__nv_bfloat16*>(out);
  if (raw_f32)
    launch_k<bn_apply_kernel<float>>(bn_grid(P, C / 8, bn_resident_blocks(bn_apply_kernel<float>)), kBnThreads, 0, s, reinterpret_cast<const float*>(raw), scale, shift, a1, a2, o, P, C / 8, relu, mode);
  else
    launch_k<bn_apply_kernel<__nv_bfloat16>>(bn_grid(P, C / 8, bn_resident_blocks(bn_apply_kernel<__nv_bfloat16>)), kBnThreads, 0, s, reinterpret_cast<const __nv_bfloat16*>(raw), scale, shift, a1, a2, o, P, C / 8, relu, mode);
  return cudaGetLastError();
}
cudaError_t launch_bn_finalize_apply(const void* raw, int raw_f32, double* sums, const float* gamma, const float* beta, float* rm,
                                     float* rv, float* scale, float* shift, float* mean, float* invstd, const void* add1,
                                     const void* add2, void* out, long long P, int C, int relu, int mode, float eps,
                                     float momentum, cudaStream_t s) {
  if (!bn_c_ok(C) || C > kBnMaxC) return cudaErrorInvalidValue;
  auto a1 = reinterpret_cast<const __nv_bfloat16*>(add1);
  auto a2 = reinterpret_cast<const __nv_bfloat16*>(add2);
  auto o = reinterpret_cast<__nv_bfloat16*>(out);
  // ticket counter: last 4 bytes of the coefficient-scratch block of `sums` (see launch_bn_bwd_apply)
  unsigned int* counter = reinterpret_cast<unsigned int*>(sums + (long long)kBnBwdReplicas * 2 * C) + 3 * C;
  if (raw_f32)
    launch_k<bn_finalize_apply_kernel<float>>(bn_grid(P, C / 8, bn_resident_blocks(bn_finalize_apply_kernel<float>)), kBnThreads, 0, s, reinterpret_cast<const float*>(raw), sums, counter,
                                              gamma, beta, rm, rv, scale, shift, mean, invstd, a1, a2, o, P, C / 8, relu, mode,
                                              double(P), eps, momentum);
  else
    launch_k<bn_finalize_apply_kernel<__nv_bfloat16>>(bn_grid(P, C / 8, bn_resident_blocks(bn_finalize_apply_kernel<__nv_bfloat16>)), kBnThreads, 0, s, reinterpret_cast<const __nv_bfloat16*>(raw),
                                                      sums, counter, gamma, beta, rm, rv, scale, shift, mean, invstd, a1, a2, o, P,
                                                      C / 8, relu, mode, double(P), eps, momentum);
  return cudaGetLastError();
}
cudaError_t launch_bn_bwd_reduce(const void* dout, const void* raw, int raw_f32, const void* add1, const float* scale,
                                 const float* shift, const float* mean, const float* invstd, double* sums, long long P,
                                 int C, int relu, int mode, cudaStream_t s) {
  if (!bn_c_ok(C) || C / 4 > kBnThreads) return cudaErrorInvalidValue;
  auto d = reinterpret_cast<const __nv_bfloat16*>(dout);
  auto a1 = reinterpret_cast<const __nv_bfloat16*>(add1);
  if (raw_f32)
    launch_k<bn_bwd_reduce_kernel<float>>(bn_grid4(P, C / 4, bn_resident_blocks(bn_bwd_reduce_kernel<float>)), kBnThreads, 0, s, d, reinterpret_cast<const float*>(raw), a1, scale, shift, sums, P, C / 4, relu, mode);
  else
    launch_k<bn_bwd_reduce_kernel<__nv_bfloat16>>(bn_grid4(P, C / 4, bn_resident_blocks(bn_bwd_reduce_kernel<__nv_bfloat16>)), kBnThreads, 0, s, d, reinterpret_cast<const __nv_bfloat16*>(raw), a1, scale, shift, sums, P, C / 4, relu, mode);
  return cudaGetLastError();
}
cudaError_t launch_bn_bwd_apply(const void* dout, const void* raw, int raw_f32, const void* add1, const float* gamma,
                                const float* scale, const float* shift, const float* mean, const float* invstd,
                                double* sums, void* draw, void* dres, float* dgamma, float* dbeta, long long P, int C,
                                int relu, int mode, int eval_mode, int shuffle_oh, int shuffle_ow, cudaStream_t s) {
  if (!bn_c_ok(C) || C / 4 > kBnThreads) return cudaErrorInvalidValue;
  auto d = reinterpret_cast<const __nv_bfloat16*>(dout);
  auto a1 = reinterpret_cast<const __nv_bfloat16*>(add1);
  auto dr = reinterpret_cast<__nv_bfloat16*>(draw);
  auto ds = reinterpret_cast<__nv_bfloat16*>(dres);
  // coefficient scratch: the (DP_BN_BWD_REPLICAS + 1)-th 2*C block of `sums` (16*C bytes >= 3*C floats); its last
  // float slot [3*C] is the ticket counter of the fused kernels
  float* coef = reinterpret_cast<float*>(sums + (long long)kBnBwdReplicas * 2 * C);
  static const int fuse = env_flag("DP_BN_FUSE", 1);
  if (fuse && C <= kBnMaxC) {
    BnBwdFused fz;
    fz.sums = sums; fz.counter = reinterpret_cast<unsigned int*>(coef) + 3 * C;
    fz.gamma = gamma; fz.mean = mean; fz.invstd = invstd; fz.dgamma = dgamma; fz.dbeta = dbeta;
    fz.invP = 1.0 / double(P); fz.eval_mode = eval_mode;
    if (raw_f32)
      launch_k<bn_bwd_apply_kernel<float, true>>(bn_grid4(P, C / 4, bn_resident_blocks(bn_bwd_apply_kernel<float, true>)), kBnThreads, 0, s, d, reinterpret_cast<const float*>(raw), a1, scale, shift, coef, dr, ds, P, C / 4, relu, mode, shuffle_oh, shuffle_ow, fz);
    else
      launch_k<bn_bwd_apply_kernel<__nv_bfloat16, true>>(bn_grid4(P, C / 4, bn_resident_blocks(bn_bwd_apply_kernel<__nv_bfloat16, true>)), kBnThreads, 0, s, d, reinterpret_cast<const __nv_bfloat16*>(raw), a1, scale, shift, coef, dr, ds, P, C / 4, relu, mode, shuffle_oh, shuffle_ow, fz);
    return cudaGetLastError();
  }
  const BnBwdFused none = {};
  launch_k<bn_bwd_coeffs_kernel>((C + 127) / 128, 128, 0, s, sums, gamma, scale, mean, invstd, coef, dgamma, dbeta, C, 1.0 / double(P),
                                                      eval_mode);
  if (raw_f32)
    launch_k<bn_bwd_apply_kernel<float, false>>(bn_grid4(P, C / 4, bn_resident_blocks(bn_bwd_apply_kernel<float, false>)), kBnThreads, 0, s, d, reinterpret_cast<const float*>(raw), a1, scale, shift, coef, dr, ds, P, C / 4, relu, mode, shuffle_oh, shuffle_ow, none);
  else
    launch_k<bn_bwd_apply_kernel<__nv_bfloat16, false>>(bn_grid4(P, C / 4, bn_resident_blocks(bn_bwd_apply_kernel<__nv_bfloat16, false>)), kBnThreads, 0, s, d, reinterpret_cast<const __nv_bfloat16*>(raw), a1, scale, shift, coef, dr, ds, P, C / 4, relu, mode, shuffle_oh, shuffle_ow, none);
  return cudaGetLastError();
}
cudaError_t launch_zero_f64(double* p, int n, cudaStream_t s) {
  launch_k<zero_f64_kernel>((n + 255) / 256, 256, 0, s, p, n);
  return cudaGetLastError();
}
cudaError_t launch_avgpool2(const float* in, float* out, long long planes, int OH, int OW, cudaStream_t s) {
  launch_k<avgpool2_kernel>(grid_for(planes * OH * OW, 256), 256, 0, s, in, out, planes, OH, OW);
  return cudaGetLastError();
}
cudaError_t launch_hm_grad_to_nhwc(const float* g, void* out, int NB, int K, int Kp, int OH, int OW, int up,
                                   cudaStream_t s) {
  if (Kp > 64 || K > Kp || Kp % 8) return cudaErrorInvalidValue;
  const int chunks = (OH * OW + kHmPix - 1) / kHmPix;
  launch_k<hm_grad_to_nhwc_kernel>(unsigned(NB * chunks), 256, 0, s, g, reinterpret_cast<__nv_bfloat16*>(out), NB, K, Kp, OH, OW, up);
  return cudaGetLastError();
}
cudaError_t launch_mean_tokens(const void* feat, float* out, int B, int N, int D, cudaStream_t s) {
  if (D % 8 || D / 8 > 512) return cudaErrorInvalidValue;
  const int lanes = 512 / (D / 8);
  launch_k<mean_tokens_kernel>(B, 512, size_t(lanes) * D * sizeof(float), s, reinterpret_cast<const __nv_bfloat16*>(feat), out, N, D);
  return cudaGetLastError();
}
cudaError_t launch_mean_tokens_bwd(void* dfeat, const float* dmean, int B, int N, int D, cudaStream_t s) {
  if (D % 8) return cudaErrorInvalidValue;
  const long long total = (long long)B * N * D;
  launch_k<mean_tokens_bwd_kernel>(grid_for(total / 8, 256), 256, 0, s, reinterpret_cast<__nv_bfloat16*>(dfeat), dmean, N, D, total);
  return cudaGetLastError();
}
cudaError_t launch_sgemm_small(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_k,
                               long long sb_n, float* C, long long ldc, int M, int N, int K, const float* bias, int relu,
                               const float* mask_ref, long long ld_ref, float p_drop, const unsigned long long* seed,
                               int accumulate, cudaStream_t s) {
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  // M is the batch (or a layer width) and the tile count is small: split K until ~2 CTAs per SM are busy
  const int base = int(grid.x * grid.y), ksteps = (K + 31) / 32;
  int splits = (2 * 148 + base - 1) / base;
  if (splits > ksteps / 2) splits = ksteps / 2;
  if (splits < 1) splits = 1;
  const int k_per_split = ((ksteps + splits - 1) / splits) * 32;
  splits = (K + k_per_split - 1) / k_per_split;
  if (splits > 1) {
    grid.z = unsigned(splits);
    if (!accumulate) {
      cudaError_t e = cudaMemset2DAsync(C, size_t(ldc) * sizeof(float), 0, size_t(N) * sizeof(float), size_t(M), s);
      if (e != cudaSuccess) return e;
    }
    launch_k<sgemm_small_kernel>(grid, 128, 0, s, A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, M, N, K, nullptr, 0, nullptr, 0, 0.f, nullptr,
                                            1, k_per_split);
    if (bias != nullptr || relu || mask_ref != nullptr || p_drop > 0.f) {
      const long long total = (long long)M * N;
      launch_k<sgemm_small_finish_kernel>(unsigned((total + 255) / 256), 256, 0, s, C, ldc, M, N, bias, relu, mask_ref, ld_ref, p_drop,
                                                                             seed);
    }
    return cudaGetLastError();
  }
  launch_k<sgemm_small_kernel>(grid, 128, 0, s, A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, M, N, K, bias, relu, mask_ref, ld_ref,
                                          p_drop, seed, accumulate, K);
  return cudaGetLastError();
}
cudaError_t launch_relu_mask(const float* d, const float* ref, float* out, long long n, float keep_scale, cudaStream_t s) {
  launch_k<relu_mask_kernel>(unsigned((n + 255) / 256), 256, 0, s, d, ref, out, n, keep_scale);
  return cudaGetLastError();
}
cudaError_t launch_colsum(const void* x, int is_bf16, float* out, long long P, int C, long long ld, cudaStream_t s) {
  // narrow matrices (the 24-channel heat-map gradient: one warp per block) need many more blocks than two per SM to
  // cover the load latency: 296 one-warp blocks walked 498 rows each and took 35 us for 9.4 MB (profiles/r1h_step_metrics)
  const int target = C >= 128 ? 296 : 296 * 16;
  int rpb = int((P + target - 1) / target);
  if (rpb < 32) rpb = 32;
  const int block = C >= 128 ? 128 : 32;
  dim3 grid(unsigned((P + rpb - 1) / rpb), unsigned((C + block - 1) / block));
  if (is_bf16)
    launch_k<colsum_kernel<__nv_bfloat16>>(grid, block, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), out, P, C, ld, rpb);
  else
    launch_k<colsum_kernel<float>>(grid, block, 0, s, reinterpret_cast<const float*>(x), out, P, C, ld, rpb);
  return cudaGetLastError();
}

}  // namespace dp
