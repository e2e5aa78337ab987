// Fused flash-style multi-head attention forward over the CLS+patch tokens (HF modeling_dinov2.py:203-234:
// softmax(q k^T / sqrt(dh)) v, non-causal, no mask, dropout 0), head dim 64.
//
// Input  qkv bf16 [B*T, 3*D]  (q | k | v column blocks written by the fused QKV GEMM)
// Output ctx bf16 [B*T, D]    (heads merged, HF:231-232)
//
// Round-1 implementation: 64 query rows per CTA (4 warps x 16 rows), K/V streamed in 64-key tiles through a
// cp.async double buffer with XOR-swizzled shared memory, mma.sync.m16n8k16 bf16 with fp32 online softmax
// in registers (scores never touch HBM).  The tcgen05/TMEM version replaces this kernel behind the same
// entry point.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.cuh"
#include <stdlib.h>

#include "ptx.cuh"

namespace dp {
namespace {

constexpr int kBQ = 64, kBK = 64, kDH = 64;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `chunk` of row `row` in a [rows][64 bf16] tile with XOR swizzle
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// load a [64 rows][64 dh] bf16 tile (rows t0.. of image b, column block col0) into swizzled smem
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* __restrict__ qkv, long long row_base,
                                          int t0, int T, int ld, int col0, int tid) {
#pragma unroll
  for (int i = 0; i < (kBK * 8) / 128; ++i) {
    const int idx = tid + i * 128;
    const int row = idx >> 3, chunk = idx & 7;
    const int t = t0 + row;
    const bool ok = t < T;
    const __nv_bfloat16* src = qkv + (row_base + (ok ? t : 0)) * ld + col0 + chunk * 8;
    cp_async16(smem_base + sw_off(row, chunk), src, ok ? 16 : 0);
  }
}

__global__ void __launch_bounds__(128) attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ ctx, int T, int D, float scale_log2) {
  pdl_grid_sync();
  __shared__ __align__(128) uint8_t smem[kBQ * 128 + 2 * kBK * 128 + 2 * kBK * 128];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + kBQ * 128;
  const uint32_t sV = sK + 2 * kBK * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBQ, h = blockIdx.y, b = blockIdx.z;
  const int ld = 3 * D;
  const long long row_base = (long long)b * T;
  const int nkt = (T + kBK - 1) / kBK;

  load_tile(sQ, qkv, row_base, q0, T, ld, h * kDH, tid);
  load_tile(sK, qkv, row_base, 0, T, ld, D + h * kDH, tid);
  load_tile(sV, qkv, row_base, 0, T, ld, 2 * D + h * kDH, tid);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) {
      load_tile(sK + (buf ^ 1) * kBK * 128, qkv, row_base, (kt + 1) * kBK, T, ld, D + h * kDH, tid);
      load_tile(sV + (buf ^ 1) * kBK * 128, qkv, row_base, (kt + 1) * kBK, T, ld, 2 * D + h * kDH, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int m = lane >> 3;
        const int row = warp * 16 + (m & 1) * 8 + (lane & 7);
        const int chunk = ks * 2 + (m >> 1);
        ldsm_x4(sQ + sw_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t kb = sK + buf * kBK * 128, vb = sV + buf * kBK * 128;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int nb = 0; nb < 8; nb += 2) {
        const int m = lane >> 3;
        const int row = (nb + (m >> 1)) * 8 + (lane & 7);
        const int chunk = ks * 2 + (m & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + sw_off(row, chunk), b0, b1, b2, b3);
        mma_bf16(s[nb], qf[ks], b0, b1);
        mma_bf16(s[nb + 1], qf[ks], b2, b3);
      }
    }
    // scale, mask keys beyond T, online softmax
    const int key0 = kt * kBK + 2 * (lane & 3);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + nb * 8 + (j & 1);
        const float v = (key < T) ? s[nb][j] * scale_log2 : -INFINITY;
        s[nb][j] = v;
        mx[j >> 1] = fmaxf(mx[j >> 1], v);
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(mrow[r], mx[r]);
      alpha[r] = exp2f(mrow[r] - mnew);  // first tile: exp2(-inf) = 0
      mrow[r] = mnew;
    }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = exp2f(s[nb][j] - mrow[j >> 1]);
        s[nb][j] = p;
        ls[j >> 1] += p;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * alpha[r] + ls[r];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o[nb][0] *= alpha[0];
      o[nb][1] *= alpha[0];
      o[nb][2] *= alpha[1];
      o[nb][3] *= alpha[1];
    }
    // O += P V
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      a[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      a[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      a[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int nb = 0; nb < 8; nb += 2) {
        const int m = lane >> 3;
        const int row = j * 16 + (m & 1) * 8 + (lane & 7);
        const int chunk = nb + (m >> 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vb + sw_off(row, chunk), b0, b1, b2, b3);
        mma_bf16(o[nb], a, b0, b1);
        mma_bf16(o[nb + 1], a, b2, b3);
      }
    }
    __syncthreads();  // all warps done with this K/V buffer before it is refilled
  }
  // finalize: row sums across the quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv0 = 1.0f / lrow[0], inv1 = 1.0f / lrow[1];
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int col = h * kDH + nb * 8 + 2 * (lane & 3);
    if (r0 < T)
      *reinterpret_cast<uint32_t*>(ctx + (row_base + r0) * D + col) = pack_bf16x2(o[nb][0] * inv0, o[nb][1] * inv0);
    if (r1 < T)
      *reinterpret_cast<uint32_t*>(ctx + (row_base + r1) * D + col) = pack_bf16x2(o[nb][2] * inv1, o[nb][3] * inv1);
  }
}

}  // namespace

cudaError_t launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, int B, int T, int heads, float scale,
                                cudaStream_t s);

cudaError_t launch_attention_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, int B, int T, int heads, float scale,
                                 cudaStream_t s) {
  // short sequences (224x224 images: 257 tokens) run on the tcgen05 kernel; longer ones (448x448: 1025) on the
  // mma.sync flash kernel below.  DP_ATTN_LEGACY=1 forces the latter (A/B comparison).
  static int legacy = -1;
  if (legacy < 0) { const char* v = getenv("DP_ATTN_LEGACY"); legacy = v ? atoi(v) : 0; }
  if (!legacy) {
    cudaError_t e = launch_attention_tc(qkv, ctx, B, T, heads, scale, s);
    if (e != cudaErrorNotSupported) return e;
  }
  const int D = heads * kDH;
  dim3 grid((T + kBQ - 1) / kBQ, heads, B);
  launch_k<attention_fwd_kernel>(grid, 128, 0, s, qkv, ctx, T, D, scale * 1.4426950408889634f);
  return cudaGetLastError();
}

}  // namespace dp
