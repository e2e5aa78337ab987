// Backward of the multi-head self-attention (HF modeling_dinov2.py:203-234) for the un-frozen encoder layers of
// Dinov2PoseModel(unfreeze_last_n_layers = n) (reference model/dinov2_pose.py:25-39).
//
//   S = q k^T * scale,  P = softmax(S),  O = P v
//   delta_i = sum_d dO[i,d] * O[i,d]
//   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - delta),   dQ = scale * dS K,   dK = scale * dS^T Q
//
// Inputs  qkv  bf16 [B*T, 3*D] (q | k | v column blocks), ctx = O bf16 [B*T, D], dctx = dO bf16 [B*T, D]
// Output  dqkv bf16 [B*T, 3*D] (dq | dk | dv in the layout of qkv, i.e. the operand of the QKV weight / input gradients)
// Scratch stats fp32 [2][B*heads*T]: log2-domain log-sum-exp of every score row, and delta
//
// Two kernels, both with the tiling of the round-1 mma.sync forward kernel (64 rows per CTA, 4 warps x 16 rows,
// 64-row tiles of the other side streamed through a cp.async double buffer, mma.sync m16n8k16 bf16, fp32 accumulation,
// scores never leave the SM):
//   1. dq kernel, one CTA per 64 queries: pass 1 recomputes the row statistics (the forward kernels do not store
//      them), pass 2 walks the key tiles again and accumulates dQ; writes the statistics for kernel 2.
//   2. dkv kernel, one CTA per 64 keys: walks the query tiles with the TRANSPOSED score tile S^T = K Q^T, so that P^T and
//      dS^T come out of the tensor cores directly in A-operand layout for dV += P^T dO and dK += dS^T Q.
// No atomics: every output element has exactly one writer (deterministic).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.cuh"
#include "ptx.cuh"

namespace dp {
namespace {

constexpr int kBQ = 64, kBK = 64, kDH = 64;
constexpr int kTile = 64 * 128;   // bytes of one [64][64] bf16 tile

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
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ float dot_bf16x2(uint32_t a, uint32_t b) {
  return __uint_as_float(a << 16) * __uint_as_float(b << 16) +
         __uint_as_float(a & 0xffff0000u) * __uint_as_float(b & 0xffff0000u);
}

// [64 rows][64 cols] bf16 tile: rows t0.. of image (row_base), columns col0.. of a matrix with leading dimension ld;
// rows beyond T are zero-filled
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* __restrict__ src, long long row_base,
                                          int t0, int T, int ld, int col0, int tid) {
#pragma unroll
  for (int i = 0; i < (64 * 8) / 128; ++i) {
    const int idx = tid + i * 128;
    const int row = idx >> 3, chunk = idx & 7;
    const int t = t0 + row;
    const bool ok = t < T;
    const __nv_bfloat16* p = src + (row_base + (ok ? t : 0)) * ld + col0 + chunk * 8;
    cp_async16(smem_base + sw_off(row, chunk), p, ok ? 16 : 0);
  }
}

// A-operand fragments of this warp's 16 rows of a tile
__device__ __forceinline__ void load_a_frags(uint32_t tile, int warp, int lane, uint32_t (&f)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int m = lane >> 3;
    const int row = warp * 16 + (m & 1) * 8 + (lane & 7);
    const int chunk = ks * 2 + (m >> 1);
    ldsm_x4(tile + sw_off(row, chunk), f[ks][0], f[ks][1], f[ks][2], f[ks][3]);
  }
}

// c[16 x 64] += A[16 x 64] * Bt^T, Bt = [64 n][64 k] tile (row-major in n): S = Q K^T form
__device__ __forceinline__ void mma_a_bt(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t bt, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int nb = 0; nb < 8; nb += 2) {
      const int m = lane >> 3;
      const int row = (nb + (m >> 1)) * 8 + (lane & 7);
      const int chunk = ks * 2 + (m & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(bt + sw_off(row, chunk), b0, b1, b2, b3);
      mma_bf16(c[nb], a[ks], b0, b1);
      mma_bf16(c[nb + 1], a[ks], b2, b3);
    }
  }
}

// c[16 x 64] += P[16 x 64] * Bm, P given as an fp32 accumulator-layout tile (rounded to bf16), Bm = [64 k][64 n] tile
__device__ __forceinline__ void mma_p_b(float (&c)[8][4], const float (&p)[8][4], uint32_t bm, int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * j][0], p[2 * j][1]);
    a[1] = pack_bf16x2(p[2 * j][2], p[2 * j][3]);
    a[2] = pack_bf16x2(p[2 * j + 1][0], p[2 * j + 1][1]);
    a[3] = pack_bf16x2(p[2 * j + 1][2], p[2 * j + 1][3]);
#pragma unroll
    for (int nb = 0; nb < 8; nb += 2) {
      const int m = lane >> 3;
      const int row = j * 16 + (m & 1) * 8 + (lane & 7);
      const int chunk = nb + (m >> 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(bm + sw_off(row, chunk), b0, b1, b2, b3);
      mma_bf16(c[nb], a, b0, b1);
      mma_bf16(c[nb + 1], a, b2, b3);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}

// ------------------------------------------------------------------------------------------------ kernel 1: dQ
__global__ void __launch_bounds__(128) attention_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                               const __nv_bfloat16* __restrict__ ctx,
                                                               const __nv_bfloat16* __restrict__ dctx,
                                                               __nv_bfloat16* __restrict__ dqkv, float* __restrict__ lse_out,
                                                               float* __restrict__ delta_out, int T, int D, float scale,
                                                               float scale_log2) {
  pdl_grid_sync();
  __shared__ __align__(128) uint8_t smem[6 * kTile];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sdO = sQ + kTile;
  const uint32_t sK = sdO + kTile;      // 2 buffers
  const uint32_t sV = sK + 2 * kTile;   // 2 buffers
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBQ, h = blockIdx.y, b = blockIdx.z, heads = gridDim.y;
  const int ld = 3 * D;
  const long long row_base = (long long)b * T;
  const int nkt = (T + kBK - 1) / kBK;

  load_tile(sQ, qkv, row_base, q0, T, ld, h * kDH, tid);
  load_tile(sdO, dctx, row_base, q0, T, D, h * kDH, tid);
  load_tile(sK, qkv, row_base, 0, T, ld, D + h * kDH, tid);
  cp_async_commit();

  uint32_t qf[4][4], dof[4][4];
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

  // ---- pass 1: row maximum and sum of exponentials (online, log2 domain)
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) {
      load_tile(sK + (buf ^ 1) * kTile, qkv, row_base, (kt + 1) * kBK, T, ld, D + h * kDH, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
      load_a_frags(sQ, warp, lane, qf);
      load_a_frags(sdO, warp, lane, dof);
    }
    float s[8][4];
    zero_acc(s);
    mma_a_bt(s, qf, sK + buf * kTile, lane);
    const int key0 = kt * kBK + 2 * (lane & 3);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + nb * 8 + (j & 1);
        const float v = (key < T) ? s[nb][j] * scale_log2 : -INFINITY;
        s[nb][j] = v;
        mx[j >> 1] = fmaxf(mx[j >> 1], v);
      }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(mrow[r], mx[r]);
      lrow[r] *= exp2f(mrow[r] - mnew);
      mrow[r] = mnew;
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) ls[j >> 1] += exp2f(s[nb][j] - mrow[j >> 1]);
    lrow[0] += ls[0];
    lrow[1] += ls[1];
    __syncthreads();
  }
  float lse[2], delta[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    lse[r] = mrow[r] + log2f(lrow[r]);
  }
  // ---- delta = rowsum(dO o O): O read from global memory in the fragment layout of dof
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  {
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c0 = h * kDH + ks * 16 + 2 * (lane & 3);
      uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
      if (r0 < T) {
        o0 = __ldg(reinterpret_cast<const uint32_t*>(ctx + (row_base + r0) * D + c0));
        o2 = __ldg(reinterpret_cast<const uint32_t*>(ctx + (row_base + r0) * D + c0 + 8));
      }
      if (r1 < T) {
        o1 = __ldg(reinterpret_cast<const uint32_t*>(ctx + (row_base + r1) * D + c0));
        o3 = __ldg(reinterpret_cast<const uint32_t*>(ctx + (row_base + r1) * D + c0 + 8));
      }
      d0 += dot_bf16x2(dof[ks][0], o0) + dot_bf16x2(dof[ks][2], o2);
      d1 += dot_bf16x2(dof[ks][1], o1) + dot_bf16x2(dof[ks][3], o3);
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    delta[0] = d0;
    delta[1] = d1;
  }
  if ((lane & 3) == 0) {
    const long long sb = ((long long)b * heads + h) * T;
    if (r0 < T) { lse_out[sb + r0] = lse[0]; delta_out[sb + r0] = delta[0]; }
    if (r1 < T) { lse_out[sb + r1] = lse[1]; delta_out[sb + r1] = delta[1]; }
  }

  // ---- pass 2: dQ = scale * sum_tiles (P o (dO V^T - delta)) K
  load_tile(sK, qkv, row_base, 0, T, ld, D + h * kDH, tid);
  load_tile(sV, qkv, row_base, 0, T, ld, 2 * D + h * kDH, tid);
  cp_async_commit();
  float dq[8][4];
  zero_acc(dq);
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) {
      load_tile(sK + (buf ^ 1) * kTile, qkv, row_base, (kt + 1) * kBK, T, ld, D + h * kDH, tid);
      load_tile(sV + (buf ^ 1) * kTile, qkv, row_base, (kt + 1) * kBK, T, ld, 2 * D + h * kDH, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t kb = sK + buf * kTile, vb = sV + buf * kTile;
    float s[8][4], dp_[8][4];
    zero_acc(s);
    zero_acc(dp_);
    mma_a_bt(s, qf, kb, lane);
    mma_a_bt(dp_, dof, vb, lane);
    const int key0 = kt * kBK + 2 * (lane & 3);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + nb * 8 + (j & 1);
        const float p = (key < T) ? exp2f(s[nb][j] * scale_log2 - lse[j >> 1]) : 0.f;
        s[nb][j] = p * (dp_[nb][j] - delta[j >> 1]);
      }
    mma_p_b(dq, s, kb, lane);
    __syncthreads();
  }
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int col = h * kDH + nb * 8 + 2 * (lane & 3);
    if (r0 < T)
      *reinterpret_cast<uint32_t*>(dqkv + (row_base + r0) * ld + col) = pack_bf16x2(dq[nb][0] * scale, dq[nb][1] * scale);
    if (r1 < T)
      *reinterpret_cast<uint32_t*>(dqkv + (row_base + r1) * ld + col) = pack_bf16x2(dq[nb][2] * scale, dq[nb][3] * scale);
  }
}

// ------------------------------------------------------------------------------------------------ kernel 2: dK, dV
// dynamic shared memory: K tile, V tile, 2 x Q tile, 2 x dO tile, 2 x {lse[64], delta[64]}
constexpr int kDkvSmem = 6 * kTile + 2 * 2 * 64 * 4;

__global__ void __launch_bounds__(128) attention_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                const __nv_bfloat16* __restrict__ dctx,
                                                                __nv_bfloat16* __restrict__ dqkv,
                                                                const float* __restrict__ lse_in,
                                                                const float* __restrict__ delta_in, int T, int D, float scale,
                                                                float scale_log2) {
  pdl_grid_sync();
  extern __shared__ __align__(128) uint8_t dsm[];
  const uint32_t sK = smem_u32(dsm);
  const uint32_t sV = sK + kTile;
  const uint32_t sQ = sV + kTile;        // 2 buffers
  const uint32_t sdO = sQ + 2 * kTile;   // 2 buffers
  float* sStat = reinterpret_cast<float*>(dsm + 6 * kTile);   // [2 buffers][lse 64 | delta 64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * kBK, h = blockIdx.y, b = blockIdx.z, heads = gridDim.y;
  const int ld = 3 * D;
  const long long row_base = (long long)b * T;
  const long long sb = ((long long)b * heads + h) * T;
  const int nqt = (T + kBQ - 1) / kBQ;

  auto load_stats = [&](int qt, int buf) {
    const int i = tid & 63;
    const int t = qt * kBQ + i;
    const float* src = (tid < 64) ? lse_in : delta_in;
    sStat[buf * 128 + (tid < 64 ? 0 : 64) + i] = (t < T) ? __ldg(src + sb + t) : 0.f;
  };

  load_tile(sK, qkv, row_base, k0, T, ld, D + h * kDH, tid);
  load_tile(sV, qkv, row_base, k0, T, ld, 2 * D + h * kDH, tid);
  load_tile(sQ, qkv, row_base, 0, T, ld, h * kDH, tid);
  load_tile(sdO, dctx, row_base, 0, T, D, h * kDH, tid);
  cp_async_commit();
  load_stats(0, 0);

  uint32_t kf[4][4], vf[4][4];
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);

  for (int qt = 0; qt < nqt; ++qt) {
    const int buf = qt & 1;
    if (qt + 1 < nqt) {
      load_tile(sQ + (buf ^ 1) * kTile, qkv, row_base, (qt + 1) * kBQ, T, ld, h * kDH, tid);
      load_tile(sdO + (buf ^ 1) * kTile, dctx, row_base, (qt + 1) * kBQ, T, D, h * kDH, tid);
      cp_async_commit();
      load_stats(qt + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qt == 0) {
      load_a_frags(sK, warp, lane, kf);
      load_a_frags(sV, warp, lane, vf);
    }
    const uint32_t qb = sQ + buf * kTile, dob = sdO + buf * kTile;
    const float* st = sStat + buf * 128;
    float s[8][4], dp_[8][4];
    zero_acc(s);
    zero_acc(dp_);
    mma_a_bt(s, kf, qb, lane);       // S^T  [keys x queries]
    mma_a_bt(dp_, vf, dob, lane);    // dP^T [keys x queries]
    const int qc0 = 2 * (lane & 3);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int c = nb * 8 + qc0;
      const float2 l2 = *reinterpret_cast<const float2*>(st + c);
      const float2 d2 = *reinterpret_cast<const float2*>(st + 64 + c);
      const bool ok0 = qt * kBQ + c < T, ok1 = qt * kBQ + c + 1 < T;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (j & 1) ? ok1 : ok0;
        const float l = (j & 1) ? l2.y : l2.x;
        const float dl = (j & 1) ? d2.y : d2.x;
        const float p = ok ? exp2f(s[nb][j] * scale_log2 - l) : 0.f;
        s[nb][j] = p;
        dp_[nb][j] = p * (dp_[nb][j] - dl);
      }
    }
    mma_p_b(dv, s, dob, lane);     // dV += P^T dO
    mma_p_b(dk, dp_, qb, lane);    // dK += dS^T Q
    __syncthreads();
  }
  const int r0 = k0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int col = h * kDH + nb * 8 + 2 * (lane & 3);
    if (r0 < T) {
      *reinterpret_cast<uint32_t*>(dqkv + (row_base + r0) * ld + D + col) = pack_bf16x2(dk[nb][0] * scale, dk[nb][1] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + (row_base + r0) * ld + 2 * D + col) = pack_bf16x2(dv[nb][0], dv[nb][1]);
    }
    if (r1 < T) {
      *reinterpret_cast<uint32_t*>(dqkv + (row_base + r1) * ld + D + col) = pack_bf16x2(dk[nb][2] * scale, dk[nb][3] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + (row_base + r1) * ld + 2 * D + col) = pack_bf16x2(dv[nb][2], dv[nb][3]);
    }
  }
}

}  // namespace

cudaError_t launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                                 __nv_bfloat16* dqkv, float* stats, int B, int T, int heads, float scale, cudaStream_t s) {
  const int D = heads * kDH;
  const float scale_log2 = scale * 1.4426950408889634f;
  float* lse = stats;
  float* delta = stats + (long long)B * heads * T;
  dim3 grid((T + kBQ - 1) / kBQ, heads, B);
  cudaError_t e = launch_k<attention_bwd_dq_kernel>(grid, 128, 0, s, qkv, ctx, dctx, dqkv, lse, delta, T, D, scale, scale_log2);
  if (e != cudaSuccess) return e;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attention_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem);
    attr = true;
  }
  e = launch_k<attention_bwd_dkv_kernel>(grid, 128, kDkvSmem, s, qkv, dctx, dqkv, lse, delta, T, D, scale, scale_log2);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace dp
