// The prediction layer of the heat-map head: Conv2d(64, K, kernel 1) with bias (reference model/pose_heads.py:335-340,
// `prediction.3`), K = number of key-points (24), on the 48 x 48 (or 96 x 96) map, output fp32 NCHW -- and its backward.
//
// 2 * P * 64 * K FLOPs on P = 147 456 pixels (batch 64) is 0.45 GFLOP against 33 MB of traffic: HBM bound by a factor
// of 30, and a poor fit for the 128 x N tcgen05 tile pipeline (N = 24: one k-block per tile, 1152 tiles whose fixed
// costs dominate -- 28 us forward, and the backward needed a layout-conversion kernel, a column sum, a weight-gradient
// GEMM and an input-gradient GEMM: 25 + 35 + 18 + 18 us, profiles/r1h_step_metrics.md).  Here both directions are one
// CUDA-core kernel each, fp32 weights straight from the parameter (no bf16 re-packing), fp32 accumulation:
//   forward   out[b, k, pix] = bias[k] + sum_c a[p, c] * w[k, c]
//   backward  d[p, c]   = sum_k g[b, k, pix] * w[k, c]                   (input gradient, bf16)
//             dW[k, c] += sum_p g[b, k, pix] * a[p, c],  db[k] += sum_p g[b, k, pix]
// a: bf16 [P, 64] (NHWC rows), g / out: fp32 NCHW [NB, K, HW], P = NB * HW.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.cuh"
#include "ptx.cuh"

namespace dp {
namespace {

constexpr int kC = 64;            // input channels of prediction.3
constexpr int kTilePix = 256;     // pixels per tile
constexpr int kARow = kC + 8;     // bf16 elements per shared-memory row of the activation tile (144 B: conflict-free 16 B reads)

__device__ __forceinline__ void unpack8f(const uint4& t, float (&f)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}

// coalesced load of a [256 pixels][64 channels] bf16 tile into padded shared memory; rows beyond P are zero
template <int THREADS>
__device__ __forceinline__ void load_a_tile(__nv_bfloat16* sA, const __nv_bfloat16* __restrict__ a, long long p0, long long P) {
#pragma unroll
  for (int i = 0; i < kTilePix * 8 / THREADS; ++i) {
    const int idx = threadIdx.x + THREADS * i;
    const int row = idx >> 3, ch = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p0 + row < P) v = __ldg(reinterpret_cast<const uint4*>(a + (p0 + row) * kC) + ch);
    *reinterpret_cast<uint4*>(sA + row * kARow + ch * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------ forward
// 128 threads, two pixels per thread (each broadcast weight load feeds two FMAs), one tile per block.
template <int KP>
__global__ void __launch_bounds__(128) pred1x1_fwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out, long long P,
                                                          int HW, int K) {
  pdl_grid_sync();
  __shared__ __align__(16) float sW[kC][KP];                 // transposed, zero padded: sW[c][k]
  __shared__ __align__(16) __nv_bfloat16 sA[kTilePix * kARow];
  for (int i = threadIdx.x; i < kC * KP; i += 128) {
    const int c = i / KP, k = i - c * KP;
    sW[c][k] = k < K ? __ldg(w + k * kC + c) : 0.f;
  }
  const long long p0 = (long long)blockIdx.x * kTilePix;
  load_a_tile<128>(sA, a, p0, P);
  __syncthreads();
  float acc0[KP], acc1[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) acc0[k] = acc1[k] = (k < K) ? __ldg(bias + k) : 0.f;
  const int t = threadIdx.x;
#pragma unroll 2
  for (int c8 = 0; c8 < 8; ++c8) {
    float x0[8], x1[8];
    unpack8f(*reinterpret_cast<const uint4*>(sA + t * kARow + c8 * 8), x0);
    unpack8f(*reinterpret_cast<const uint4*>(sA + (t + 128) * kARow + c8 * 8), x1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(&sW[c8 * 8 + j][0]);
#pragma unroll
      for (int k4 = 0; k4 < KP / 4; ++k4) {
        const float4 wv = wr[k4];
        acc0[4 * k4 + 0] = fmaf(x0[j], wv.x, acc0[4 * k4 + 0]); acc1[4 * k4 + 0] = fmaf(x1[j], wv.x, acc1[4 * k4 + 0]);
        acc0[4 * k4 + 1] = fmaf(x0[j], wv.y, acc0[4 * k4 + 1]); acc1[4 * k4 + 1] = fmaf(x1[j], wv.y, acc1[4 * k4 + 1]);
        acc0[4 * k4 + 2] = fmaf(x0[j], wv.z, acc0[4 * k4 + 2]); acc1[4 * k4 + 2] = fmaf(x1[j], wv.z, acc1[4 * k4 + 2]);
        acc0[4 * k4 + 3] = fmaf(x0[j], wv.w, acc0[4 * k4 + 3]); acc1[4 * k4 + 3] = fmaf(x1[j], wv.w, acc1[4 * k4 + 3]);
      }
    }
  }
  // NCHW stores: for a fixed k the 32 lanes of a warp write 32 consecutive pixels
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long p = p0 + t + 128 * u;
    if (p < P) {
      const long long b = p / HW;
      const int pix = int(p - b * HW);
      float* o = out + (b * K) * HW + pix;
#pragma unroll
      for (int k = 0; k < KP; ++k)
        if (k < K) o[(long long)k * HW] = u == 0 ? acc0[k] : acc1[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// Persistent blocks of 256 threads.  Per 256-pixel tile: phase 1, thread = pixel: the K gradients of the pixel (coalesced
// NCHW reads) -> input gradient of its 64 channels; phase 2, thread = (channel pair, key-point half, pixel quarter): weight
// gradient partial sums over the tile's pixels from shared memory, kept in registers across tiles.
constexpr int kGRow = kTilePix + 4;   // fp32 elements per row of the gradient tile (k-major), 16-byte aligned rows
template <int KP> constexpr int pred_bwd_smem() { return KP * kC * 4 + KP * kGRow * 4 + 2 * kTilePix * kARow * 2; }

template <int KP>
__global__ void __launch_bounds__(256) pred1x1_bwd_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ a,
                                                          const float* __restrict__ w, __nv_bfloat16* __restrict__ d,
                                                          float* __restrict__ dW, float* __restrict__ db, long long P, int HW,
                                                          int K, int tiles) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t psm[];
  float* sW = reinterpret_cast<float*>(psm);                          // [KP][64]
  float* sG = sW + KP * kC;                                           // [KP][kGRow]
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(sG + KP * kGRow);   // [256][kARow]
  __nv_bfloat16* sD = sA + kTilePix * kARow;                               // [256][kARow] input-gradient tile (coalesced write-out)
  for (int i = threadIdx.x; i < KP * kC; i += 256) sW[i] = (i / kC) < K ? __ldg(w + i) : 0.f;
  const int t = threadIdx.x;
  // phase 2: thread = (channel pair cp, key-point half kh, pixel quarter pg): 2 x KP/2 partial sums in registers.  Per four
  // pixels a thread issues 4 + KP/2 shared-memory loads for 4 * KP FMAs (a (channel, KP/4 key-points) mapping needed
  // 4 + KP/4 loads for KP FMAs and was bound by the shared-memory pipe)
  constexpr int KH = KP / 2;
  const int cp = t & 31, kh = (t >> 5) & 1, pg = t >> 6;
  float wacc[KH][2], bacc = 0.f;
#pragma unroll
  for (int j = 0; j < KH; ++j) wacc[j][0] = wacc[j][1] = 0.f;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long p0 = (long long)tile * kTilePix;
    __syncthreads();                        // previous tile's phase 2 has finished with sG / sA (and sW is staged)
    load_a_tile<256>(sA, a, p0, P);
    const long long p = p0 + t;
    float gk[KP];
    {
      const bool ok = p < P;
      const long long b = ok ? p / HW : 0;
      const int pix = ok ? int(p - b * HW) : 0;
      const float* gp = g + (b * K) * HW + pix;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        gk[k] = (ok && k < K) ? __ldg(gp + (long long)k * HW) : 0.f;
        sG[k * kGRow + t] = gk[k];
      }
    }
    // phase 1: d[p, c] = sum_k g[k] * w[k, c]
    if (p < P) {
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float dacc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) dacc[c] = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const float4* wr = reinterpret_cast<const float4*>(sW + k * kC + half * 32);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 wv = wr[c4];
            dacc[4 * c4 + 0] = fmaf(gk[k], wv.x, dacc[4 * c4 + 0]);
            dacc[4 * c4 + 1] = fmaf(gk[k], wv.y, dacc[4 * c4 + 1]);
            dacc[4 * c4 + 2] = fmaf(gk[k], wv.z, dacc[4 * c4 + 2]);
            dacc[4 * c4 + 3] = fmaf(gk[k], wv.w, dacc[4 * c4 + 3]);
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(sD + t * kARow + half * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_uint4(pack_bf16x2(dacc[8 * q], dacc[8 * q + 1]), pack_bf16x2(dacc[8 * q + 2], dacc[8 * q + 3]),
                              pack_bf16x2(dacc[8 * q + 4], dacc[8 * q + 5]), pack_bf16x2(dacc[8 * q + 6], dacc[8 * q + 7]));
      }
    }
    __syncthreads();                        // sG / sA / sD of this tile complete
#pragma unroll
    for (int i = 0; i < kTilePix * 8 / 256; ++i) {       // 16-byte pieces, consecutive threads -> consecutive addresses
      const int idx = t + 256 * i;
      const int row = idx >> 3, ch = idx & 7;
      if (p0 + row < P)
        *(reinterpret_cast<uint4*>(d + (p0 + row) * kC) + ch) = *reinterpret_cast<const uint4*>(sD + row * kARow + ch * 8);
    }
    // phase 2: dW[k, c] += sum_p g[p, k] * a[p, c] over this thread's quarter of the tile, four pixels per step
#pragma unroll 2
    for (int pp = pg * (kTilePix / 4); pp < (pg + 1) * (kTilePix / 4); pp += 4) {
      float a0[4], a1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(sA + (pp + i) * kARow + 2 * cp);
        a0[i] = __uint_as_float(v << 16);
        a1[i] = __uint_as_float(v & 0xffff0000u);
      }
#pragma unroll
      for (int j = 0; j < KH; ++j) {
        const float4 gv = *reinterpret_cast<const float4*>(sG + (kh * KH + j) * kGRow + pp);
        wacc[j][0] = fmaf(gv.x, a0[0], fmaf(gv.y, a0[1], fmaf(gv.z, a0[2], fmaf(gv.w, a0[3], wacc[j][0]))));
        wacc[j][1] = fmaf(gv.x, a1[0], fmaf(gv.y, a1[1], fmaf(gv.z, a1[2], fmaf(gv.w, a1[3], wacc[j][1]))));
      }
    }
    if (t < KP) {                           // bias gradient: thread k sums its row of the gradient tile
      float sacc = 0.f;
      const float4* gr = reinterpret_cast<const float4*>(sG + t * kGRow);
#pragma unroll 4
      for (int i = 0; i < kTilePix / 4; ++i) {
        const float4 v = gr[i];
        sacc += (v.x + v.y) + (v.z + v.w);
      }
      bacc += sacc;
    }
  }
  // combine the four pixel quarters in shared memory (the gradient tile is dead), then ONE atomic per output and block
  __syncthreads();
  float* red = sG;                          // [4][KP * 64] <= KP * kGRow floats
#pragma unroll
  for (int j = 0; j < KH; ++j) {
    const int k = kh * KH + j;
    red[pg * (KP * kC) + k * kC + 2 * cp] = wacc[j][0];
    red[pg * (KP * kC) + k * kC + 2 * cp + 1] = wacc[j][1];
  }
  __syncthreads();
  for (int i = t; i < K * kC; i += 256)
    atomicAdd(dW + i, (red[i] + red[KP * kC + i]) + (red[2 * KP * kC + i] + red[3 * KP * kC + i]));
  if (t < K) atomicAdd(db + t, bacc);
}

}  // namespace

cudaError_t launch_pred1x1_fwd(const __nv_bfloat16* a, const float* w, const float* bias, float* out, long long P, int HW, int K,
                               cudaStream_t s) {
  const unsigned grid = unsigned((P + kTilePix - 1) / kTilePix);
  if (K <= 24)
    launch_k<pred1x1_fwd_kernel<24>>(grid, 128, 0, s, a, w, bias, out, P, HW, K);
  else
    launch_k<pred1x1_fwd_kernel<32>>(grid, 128, 0, s, a, w, bias, out, P, HW, K);
  return cudaGetLastError();
}

cudaError_t launch_pred1x1_bwd(const float* g, const __nv_bfloat16* a, const float* w, __nv_bfloat16* d, float* dW, float* db,
                               long long P, int HW, int K, int sms, cudaStream_t s) {
  const int tiles = int((P + kTilePix - 1) / kTilePix);
  const unsigned grid = unsigned(tiles < 2 * sms ? tiles : 2 * sms);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pred1x1_bwd_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, pred_bwd_smem<24>());
    cudaFuncSetAttribute(pred1x1_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, pred_bwd_smem<32>());
    attr = true;
  }
  if (K <= 24)
    launch_k<pred1x1_bwd_kernel<24>>(grid, 256, pred_bwd_smem<24>(), s, g, a, w, d, dW, db, P, HW, K, tiles);
  else
    launch_k<pred1x1_bwd_kernel<32>>(grid, 256, pred_bwd_smem<32>(), s, g, a, w, d, dW, db, P, HW, K, tiles);
  return cudaGetLastError();
}

}  // namespace dp
