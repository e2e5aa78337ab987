// LoRA backward (reference model/lora.py:26-28 under autograd) in ONE pass over g and y.
//
// With gv = g * lambda1 * (alpha / r) * mask / (1 - p)   (gradient w.r.t. u B, u = y A saved by the forward):
//     gu[row, r]  = sum_d  gv[row, d] * B[r, d]
//     dB[r, d]   += sum_rows u[row, r] * gv[row, d]
//     dA[d, r]   += sum_rows y[row, d] * gu[row, r]
// The two-kernel fp32 version (rowwise.cu: lora_bwd_gu / lora_bwd_acc) reads g twice and y once at ~1.1 TB/s and sits
// at the very end of the backward (66 us of the 3.77 ms step, nothing left to overlap it).  Here a 4-warp block stages
// 32 rows of gv and y as bf16 in shared memory and runs the three rank-8 products as warp-level MMAs
// (mma.sync m16n8k16: rank 8 is the N -- or, zero padded, the M -- of one instruction; there is no tcgen05 shape for
// it), so the kernel is one HBM pass: 2 * rows * D * 4 bytes.  Accumulators stay in registers over the block's tiles and
// leave through shared memory as 16-byte vector reductions.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "launch.cuh"
#include "ptx.cuh"

namespace dp {
namespace {

constexpr int kLbRows = 32;       // rows per tile = two k-steps of 16
constexpr int kLbThreads = 128;   // 4 warps: warp w owns columns [w * D / 4, (w + 1) * D / 4)
constexpr int kLbRank = 8;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int D>
__global__ void __launch_bounds__(kLbThreads) lora_bwd_mma_kernel(
    const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ u, const float* __restrict__ Bm,
    const float* __restrict__ lambda1, float* __restrict__ dA, float* __restrict__ dB, long long rows, float scaling,
    float p_drop, const unsigned long long* __restrict__ seed_ptr, int rows_per_block) {
  constexpr int R = kLbRank;
  constexpr int LD = D + 8;          // padded row stride (elements): consecutive rows start 16 bytes apart modulo 128
  constexpr int C4 = D / 4;          // float4 column groups per row
  constexpr int CW = D / 4;          // columns owned by a warp
  constexpr int NT = CW / 8;         // n-tiles of the dB product per warp
  constexpr int MT = CW / 16;        // m-tiles of the dA product per warp
  constexpr int KQ = CW / 16;        // k-steps of the gu product per warp (its quarter of D)
  static_assert(D % 64 == 0 && NT % 2 == 0, "D must be a multiple of 64");
  extern __shared__ __align__(16) uint8_t lb_smem[];
  __nv_bfloat16* sGV = reinterpret_cast<__nv_bfloat16*>(lb_smem);   // [32][LD]
  __nv_bfloat16* sY = sGV + kLbRows * LD;                            // [32][LD]
  __nv_bfloat16* sU = sY + kLbRows * LD;                             // [32][8]
  __nv_bfloat16* sGU = sU + kLbRows * R;                             // [32][8]
  float* sPart = reinterpret_cast<float*>(sGU + kLbRows * R);        // [4 warps][32][8] partial gu

  pdl_grid_sync();
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  const uint32_t thresh = p_drop > 0.f ? uint32_t(fminf(p_drop, 0.999999f) * 4294967296.0f) : 0u;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;

  // B fragments of the gu product for this warp's quarter of D: b0 = {B[n][k], B[n][k + 1]}, n = lane / 4, k = 2 * (lane % 4)
  uint32_t bfr[KQ][2];
#pragma unroll
  for (int s = 0; s < KQ; ++s) {
    const float* bp = Bm + (lane >> 2) * D + w * CW + 16 * s + 2 * (lane & 3);
    bfr[s][0] = pack_bf16x2(__ldg(bp), __ldg(bp + 1));
    bfr[s][1] = pack_bf16x2(__ldg(bp + 8), __ldg(bp + 9));
  }
  float accB[NT][4], accA[MT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) accB[j][k] = 0.f;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) accA[i][k] = 0.f;

  const long long rbeg = (long long)blockIdx.x * rows_per_block;
  const long long rend = min(rbeg + rows_per_block, rows);
  const uint32_t gv_s = smem_u32(sGV), y_s = smem_u32(sY), u_s = smem_u32(sU), gu_s = smem_u32(sGU);
  for (long long r0 = rbeg; r0 < rend; r0 += kLbRows) {
    // ---- stage 32 rows: gv = g * lambda1 * s * mask and y, both rounded to bf16; rows past the end are zero
    constexpr int F = kLbRows * C4;      // float4 per tile and array
    constexpr int kBatch = 6;
#pragma unroll 1
    for (int base = 0; base < F; base += kLbThreads * kBatch) {
      float4 gq[kBatch], yq[kBatch];
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        const int f = base + t + kLbThreads * i;
        const int row = f / C4, c4 = f - row * C4;
        const long long rr = r0 + row;
        if (f < F && rr < rend) {
          gq[i] = __ldg(reinterpret_cast<const float4*>(g + rr * D) + c4);
          yq[i] = __ldg(reinterpret_cast<const float4*>(y + rr * D) + c4);
        } else {
          gq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          yq[i] = gq[i];
        }
      }
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        const int f = base + t + kLbThreads * i;
        if (f >= F) break;
        const int row = f / C4, c4 = f - row * C4;
        const float4 l = __ldg(reinterpret_cast<const float4*>(lambda1) + c4);
        float gv[4] = {gq[i].x * l.x * scaling, gq[i].y * l.y * scaling, gq[i].z * l.z * scaling, gq[i].w * l.w * scaling};
        if (p_drop > 0.f) {
          const uint64_t e0 = uint64_t(r0 + row) * D + 4 * c4;
#pragma unroll
          for (int k = 0; k < 4; ++k) gv[k] = dropout_keep(seed, e0 + k, thresh) ? gv[k] * keep_scale : 0.f;
        }
        *reinterpret_cast<uint2*>(sGV + row * LD + 4 * c4) = make_uint2(pack_bf16x2(gv[0], gv[1]), pack_bf16x2(gv[2], gv[3]));
        *reinterpret_cast<uint2*>(sY + row * LD + 4 * c4) =
            make_uint2(pack_bf16x2(yq[i].x, yq[i].y), pack_bf16x2(yq[i].z, yq[i].w));
      }
    }
    if (t < kLbRows * R / 4) {   // u: 32 rows x 8 ranks, one float4 per thread
      const int row = t >> 1, hf = t & 1;
      const long long rr = r0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr < rend) v = __ldg(reinterpret_cast<const float4*>(u + rr * R) + hf);
      *reinterpret_cast<uint2*>(sU + row * R + 4 * hf) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    __syncthreads();

    // ---- gu[32, 8] = gv[32, D] B^T: every warp reduces its quarter of D for both 16-row groups
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const int arow = 16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int s = 0; s < KQ; ++s) {
        uint32_t a[4];
        ldsm_x4(gv_s + uint32_t(arow * LD + w * CW + 16 * s + (lane >> 4) * 8) * 2u, a);
        mma_16816(c, a[0], a[1], a[2], a[3], bfr[s][0], bfr[s][1]);
      }
      float* pp = sPart + (w * kLbRows + 16 * mt + (lane >> 2)) * R + 2 * (lane & 3);
      pp[0] = c[0]; pp[1] = c[1];
      pp[8 * R] = c[2]; pp[8 * R + 1] = c[3];
    }
    __syncthreads();
#pragma unroll
    for (int o = t; o < kLbRows * R; o += kLbThreads) {
      const float v = sPart[o] + sPart[kLbRows * R + o] + sPart[2 * kLbRows * R + o] + sPart[3 * kLbRows * R + o];
      sGU[o] = __float2bfloat16(v);
    }
    __syncthreads();

    // ---- dB[8, D] += u^T gv  (A = u^T: ranks padded to 16 with zero rows) and dA[D, 8] += y^T gu
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int krow = 16 * s + (lane & 7) + ((lane >> 3) & 1) * 8;      // address rows of the x2 loads (lanes 0..15 count)
      uint32_t ua0, ua2, gb0, gb1;
      ldsm_x2_t(u_s + uint32_t(krow * R) * 2u, ua0, ua2);                // a0: (r, k 0..7), a2: (r, k 8..15); a1 = a3 = 0
      ldsm_x2_t(gu_s + uint32_t(krow * R) * 2u, gb0, gb1);               // b0: (k 0..7, r), b1: (k 8..15, r)
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        uint32_t b[4];   // {k 0..7, k 8..15} of n-tile j, then of n-tile j + 1
        ldsm_x4_t(gv_s + uint32_t(krow * LD + w * CW + 8 * (j + (lane >> 4))) * 2u, b);
        mma_16816(accB[j], ua0, 0u, ua2, 0u, b[0], b[1]);
        mma_16816(accB[j + 1], ua0, 0u, ua2, 0u, b[2], b[3]);
      }
      const int yrow = 16 * s + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        uint32_t a[4];   // a0: (d 0..7, k 0..7), a1: (d 8..15, k 0..7), a2: (d 0..7, k 8..15), a3: (d 8..15, k 8..15)
        ldsm_x4_t(y_s + uint32_t(yrow * LD + w * CW + 16 * i + ((lane >> 3) & 1) * 8) * 2u, a);
        mma_16816(accA[i], a[0], a[1], a[2], a[3], gb0, gb1);
      }
    }
    __syncthreads();   // the next tile overwrites the staged rows
  }

  // ---- hand the block's partial sums over: [dB (8 x D) | dA (D x 8)] through shared memory, 16-byte reductions
  float* sOut = reinterpret_cast<float*>(lb_smem);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    float* p = sOut + (lane >> 2) * D + w * CW + 8 * j + 2 * (lane & 3);   // rows 8..15 of the padded product are zero
    p[0] = accB[j][0]; p[1] = accB[j][1];
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    float* p = sOut + R * D + (w * CW + 16 * i + (lane >> 2)) * R + 2 * (lane & 3);
    p[0] = accA[i][0]; p[1] = accA[i][1];
    p[8 * R] = accA[i][2]; p[8 * R + 1] = accA[i][3];
  }
  __syncthreads();
  for (int o = t; o < 2 * R * D / 4; o += kLbThreads) {
    const float4 v = reinterpret_cast<const float4*>(sOut)[o];
    red_add_v4(o < R * D / 4 ? dB + 4 * o : dA + 4 * (o - R * D / 4), v);
  }
}

template <int D>
cudaError_t launch_t(const float* g, const float* y, const float* u, const float* Bm, const float* lambda1, float* dA,
                     float* dB, long long rows, float scaling, float p_drop, const unsigned long long* seed, int sms,
                     cudaStream_t s) {
  constexpr int LD = D + 8;
  size_t smem = size_t(2) * kLbRows * LD * 2 + 2 * kLbRows * kLbRank * 2 + 4 * kLbRows * kLbRank * 4;
  const size_t stage = size_t(2) * kLbRank * D * 4;
  if (smem < stage) smem = stage;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lora_bwd_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    attr = true;
  }
  // two blocks per SM; every block adds 2 * 8 * D floats into dA / dB at the end, so more blocks only add atomics
  long long grid = 2LL * sms;
  long long rpb = (rows + grid - 1) / grid;
  rpb = (rpb + kLbRows - 1) / kLbRows * kLbRows;
  grid = (rows + rpb - 1) / rpb;
  launch_k<lora_bwd_mma_kernel<D>>(unsigned(grid), kLbThreads, smem, s, g, y, u, Bm, lambda1, dA, dB, rows, scaling, p_drop, seed,
                                   int(rpb));
  return cudaGetLastError();
}

}  // namespace

// cudaErrorNotSupported: shape not covered (the caller falls back to the two-kernel fp32 version)
cudaError_t launch_lora_bwd_mma(const float* g, const float* y, const float* u, const float* Bm, const float* lambda1,
                                float* dA, float* dB, long long rows, int D, int R, float scaling, float p_drop,
                                const unsigned long long* seed, int sms, cudaStream_t s) {
  if (R != kLbRank || rows <= 0) return cudaErrorNotSupported;
  switch (D) {
    case 128: return launch_t<128>(g, y, u, Bm, lambda1, dA, dB, rows, scaling, p_drop, seed, sms, s);
    case 256: return launch_t<256>(g, y, u, Bm, lambda1, dA, dB, rows, scaling, p_drop, seed, sms, s);
    case 384: return launch_t<384>(g, y, u, Bm, lambda1, dA, dB, rows, scaling, p_drop, seed, sms, s);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace dp
