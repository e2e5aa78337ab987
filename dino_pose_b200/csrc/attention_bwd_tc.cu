// Backward of the multi-head self-attention (HF modeling_dinov2.py:203-234) on tcgen05 / TMEM, for the un-frozen encoder
// layers of Dinov2PoseModel(unfreeze_last_n_layers = n) (reference model/dinov2_pose.py:25-39).
//
//   S = q k^T * scale,  P = softmax(S),  O = P v
//   delta_i = sum_d dO[i,d] * O[i,d]
//   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - delta),   dQ = scale * dS K,   dK = scale * dS^T Q
//
// Inputs  qkv  bf16 [B*T, 3*D] (q | k | v column blocks), ctx = O bf16 [B*T, D], dctx = dO bf16 [B*T, D]
// Output  dqkv bf16 [B*T, 3*D] (dq | dk | dv in the layout of qkv, i.e. the operand of the QKV weight / input gradients)
// Scratch stats fp32 [2][B*heads*T]: log2-domain log-sum-exp of every score row, and delta
//
// Three launches, no atomics (every output element has exactly one writer: deterministic):
//   1. stats   one CTA per (image, head, 128 queries): S tiles on the tensor cores, online max / sum by the softmax warps
//              (the forward kernels store no statistics), delta from the O and dO rows.
//   2. dQ      one CTA per (image, head, 128 queries), two per SM; 64-key tiles of K / V through a TMA ring:
//                S = Q K_j^T and dP = dO V_j^T (TMEM columns [0,64) / [64,128)), thread = query row forms dS_j in bf16 as a
//                K-major operand tile, dQ += dS_j K_j (B = K_j MN-major) accumulates in TMEM columns [128,192).
//   3. dK, dV  one CTA per (image, head, 128 keys), K_j / V_j resident; 128-query tiles of Q / dO through a TMA ring:
//                S and dP as above (N = 128), thread = query row writes P and dS as bf16 tiles [query][key]; the SAME bytes
//                read as MN-major A operands are P^T and dS^T:  dV += P^T dO_i,  dK += dS^T Q_i  (TMEM [256,320) / [320,384)).
// Ragged ends (T = 257 = 2 * 128 + 1): the last key tile is rounded up to 16 keys only (N = 16 MMAs), the last query tile
// contributes K = 16 query rows to the dV / dK MMAs, rows / keys beyond T are exact zeros in P and dS.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "launch.cuh"
#include "ptx.cuh"

namespace dp {
namespace {

constexpr int kDh = 64;
constexpr int kRowBytes = kDh * 2;                 // 128 B: one row of a [rows][64] bf16 tile = one swizzle row
constexpr int kTile128 = 128 * kRowBytes;          // 16 KB
constexpr int kTile64 = 64 * kRowBytes;            // 8 KB
constexpr int kBThreads = 192;                     // warp 0 MMA issue, warp 1 TMEM alloc + TMA, warps 2..5 softmax (thread = row)
constexpr int kDkvThreads = 64 + 8 * 32;           // dK / dV kernel: eight softmax warps (two per TMEM lane quarter)

struct BwdParams {
  CUtensorMap tm128;   // qkv: 2-D {3*D, B*T} bf16, box {64, 128}, 128B swizzle
  CUtensorMap tm64;    // qkv: same tensor, box {64, 64}
  CUtensorMap tmdo;    // dctx: 2-D {D, B*T} bf16, box {64, 128}
  const __nv_bfloat16* ctx;
  const __nv_bfloat16* dctx;
  __nv_bfloat16* dqkv;
  float* lse;          // [B*heads*T]
  float* delta;        // [B*heads*T]
  int T, D, heads, tiles128;
  float scale, scale_log2;
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void unpack8(uint4 u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// ------------------------------------------------------------------------------------------------ 1. statistics
// smem: Q 16 KB, K ring 2 x 16 KB.  TMEM: S columns [0, 128).
constexpr int kStatsSmem = 3 * kTile128 + 256;

__global__ void __launch_bounds__(kBThreads, 2) attention_bwd_stats_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTile128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sK + 2 * kTile128);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // 2
  uint64_t* k_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;     // 4 warp arrivals
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 7);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.tiles128, bh = blockIdx.x / p.tiles128;
  const int h = bh % p.heads, b = bh / p.heads;
  const int row0 = b * p.T;
  const int kv_tiles = p.tiles128;
  const int n_last = ((p.T - (kv_tiles - 1) * 128 + 15) / 16) * 16;
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&p.tm128);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_holder, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();
  if (warp == 1) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kTile128);
      tma_load_2d(sQ, &p.tm128, q_full, h * kDh, row0 + qt * 128);
      for (int j = 0; j < kv_tiles; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&k_empty[st], ((j >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&k_full[st], kTile128);
        tma_load_2d(sK + st * kTile128, &p.tm128, &k_full[st], p.D + h * kDh, row0 + j * 128);
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {
      const uint32_t q_addr = smem_u32(sQ);
      mbar_wait(q_full, 0);
      for (int j = 0; j < kv_tiles; ++j) {
        const int st = j & 1;
        const int nj = (j == kv_tiles - 1) ? n_last : 128;
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);     // the softmax warps have read S_{j-1}
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * kTile128);
        const uint32_t idesc = make_idesc_bf16(128, nj, 0, 0);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base, make_sdesc_sw128(q_addr + k * 32, 0, 1024), make_sdesc_sw128(k_addr + k * 32, 0, 1024), idesc,
                    k != 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl = p.scale_log2;
    const bool warp_active = qt * 128 + q * 32 < p.T;
    float ms = -INFINITY, l = 0.f;
    for (int j = 0; j < kv_tiles; ++j) {
      const bool last = j == kv_tiles - 1;
      const int nj = last ? n_last : 128;
      const int valid = last ? p.T - j * 128 : 128;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (warp_active) {
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < nj; c += 16) {
          uint32_t v[16];
          tmem_ld_32x16(trow + uint32_t(c), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (!last || c + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
        }
        const float ms_new = fmaxf(ms, m * sl);
        const float alpha = (ms == -INFINITY) ? 0.f : ex2_approx(ms - ms_new);
        ms = ms_new;
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < nj; c += 16) {
          uint32_t v[16];
          tmem_ld_32x16(trow + uint32_t(c), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (!last || c + i < valid) sum += ex2_approx(fmaf(__uint_as_float(v[i]), sl, -ms));
        }
        l = fmaf(l, alpha, sum);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
    }
    const int t = qt * 128 + r;
    if (warp_active && t < p.T) {
      // delta = sum_d dO * O of this row (two 128-byte rows from global memory)
      const uint4* orow = reinterpret_cast<const uint4*>(p.ctx + (long long)(row0 + t) * p.D + h * kDh);
      const uint4* grow = reinterpret_cast<const uint4*>(p.dctx + (long long)(row0 + t) * p.D + h * kDh);
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float a[8], g[8];
        unpack8(__ldg(orow + c), a);
        unpack8(__ldg(grow + c), g);
#pragma unroll
        for (int i = 0; i < 8; ++i) d = fmaf(a[i], g[i], d);
      }
      const long long si = (long long)bh * p.T + t;
      p.lse[si] = ms + log2f(l);
      p.delta[si] = d;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------ 2. dQ
// smem: Q 16 KB, dO 16 KB, K ring 2 x 8 KB, V ring 2 x 8 KB, dS 8 KB ([128 q][64 keys], the bytes of one K-major tile row are
// 128 B) -> 16 KB.  TMEM (256 columns, two CTAs per SM): S [0,64), dP [64,128), dQ [128,192).
constexpr int kDqSmem = 2 * kTile128 + 4 * kTile64 + kTile128 + 256;

__global__ void __launch_bounds__(kBThreads, 2) attention_bwd_dq_tc_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + kTile128;
  uint8_t* sK = sdO + kTile128;        // 2 stages of 64 keys
  uint8_t* sV = sK + 2 * kTile64;
  uint8_t* sdS = sV + 2 * kTile64;     // [128 q rows][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + kTile128);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // 2
  uint64_t* kv_empty = bars + 3;   // 2
  uint64_t* s_full = bars + 5;
  uint64_t* ds_ready = bars + 6;   // 4 warp arrivals
  uint64_t* dq_full = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.tiles128, bh = blockIdx.x / p.tiles128;
  const int h = bh % p.heads, b = bh / p.heads;
  const int row0 = b * p.T;
  const int kv_tiles = (p.T + 63) / 64;
  const int n_last = ((p.T - (kv_tiles - 1) * 64 + 15) / 16) * 16;
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&p.tm128);
    tma_prefetch_desc(&p.tm64);
    tma_prefetch_desc(&p.tmdo);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(ds_ready, 4);
    mbar_init(dq_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_holder, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();
  if (warp == 1) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kTile128);
      tma_load_2d(sQ, &p.tm128, q_full, h * kDh, row0 + qt * 128);
      tma_load_2d(sdO, &p.tmdo, q_full, h * kDh, row0 + qt * 128);
      for (int j = 0; j < kv_tiles; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&kv_empty[st], ((j >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTile64);
        tma_load_2d(sK + st * kTile64, &p.tm64, &kv_full[st], p.D + h * kDh, row0 + j * 64);
        tma_load_2d(sV + st * kTile64, &p.tm64, &kv_full[st], 2 * p.D + h * kDh, row0 + j * 64);
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), ds_addr = smem_u32(sdS);
      const uint32_t idesc_dq = make_idesc_bf16(128, kDh, 0, 1);
      mbar_wait(q_full, 0);
      for (int j = 0; j < kv_tiles; ++j) {
        const int st = j & 1;
        const int nj = (j == kv_tiles - 1) ? n_last : 64;
        const uint32_t k_addr = smem_u32(sK + st * kTile64), v_addr = smem_u32(sV + st * kTile64);
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        // S_j and dP_j: their TMEM columns are free (the softmax warps read tile j-1 before ds_ready(j-1), waited below)
        const uint32_t idesc_s = make_idesc_bf16(128, nj, 0, 0);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base, make_sdesc_sw128(q_addr + k * 32, 0, 1024), make_sdesc_sw128(k_addr + k * 32, 0, 1024), idesc_s,
                    k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base + 64, make_sdesc_sw128(do_addr + k * 32, 0, 1024), make_sdesc_sw128(v_addr + k * 32, 0, 1024),
                    idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(ds_ready, j & 1);
        tc_fence_after();
        // dQ += dS_j K_j: A = dS_j K-major (keys are the reduction dimension), B = K_j MN-major
        for (int k = 0; k < nj / 16; ++k)
          umma_bf16(tmem_base + 128, make_sdesc_sw128(ds_addr + k * 32, 0, 1024), make_sdesc_sw128(k_addr + k * 2048, kTile64, 1024),
                    idesc_dq, (j | k) != 0 ? 1u : 0u);
        umma_commit(&kv_empty[st]);
      }
      umma_commit(dq_full);
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl = p.scale_log2;
    const int t = qt * 128 + r;
    const bool warp_active = qt * 128 + q * 32 < p.T;
    const bool row_valid = t < p.T;
    float lse = 0.f, delta = 0.f;
    if (row_valid) {
      lse = p.lse[(long long)bh * p.T + t];
      delta = p.delta[(long long)bh * p.T + t];
    }
    for (int j = 0; j < kv_tiles; ++j) {
      const bool last = j == kv_tiles - 1;
      const int nj = last ? n_last : 64;
      const int valid = last ? p.T - j * 64 : 64;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (warp_active) {
        uint8_t* tile = sdS + r * 128;
#pragma unroll 1
        for (int c = 0; c < nj; c += 16) {
          uint32_t s[16], d[16];
          tmem_ld_32x16(trow + uint32_t(c), s);
          tmem_ld_32x16(trow + 64 + uint32_t(c), d);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float e0 = ex2_approx(fmaf(__uint_as_float(s[i]), sl, -lse)) * (__uint_as_float(d[i]) - delta);
            float e1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), sl, -lse)) * (__uint_as_float(d[i + 1]) - delta);
            if (!row_valid || (last && c + i >= valid)) e0 = 0.f;
            if (!row_valid || (last && c + i + 1 >= valid)) e1 = 0.f;
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
          const int g = c >> 3;
          *reinterpret_cast<uint4*>(tile + ((g ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(tile + (((g + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
    }
    mbar_wait(dq_full, 0);
    tc_fence_after();
    if (warp_active) {    // warp-uniform: tcgen05.ld is a warp-collective instruction; only the stores are per-row
      __nv_bfloat16* dst = p.dqkv + (long long)(row0 + t) * (3 * p.D) + h * kDh;
#pragma unroll
      for (int c = 0; c < kDh; c += 16) {
        uint32_t v[16];
        tmem_ld_32x16(trow + 128 + uint32_t(c), v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * p.scale, __uint_as_float(v[2 * i + 1]) * p.scale);
        if (row_valid) {
          reinterpret_cast<uint4*>(dst + c)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          reinterpret_cast<uint4*>(dst + c)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ 3. dK, dV
// smem: K_j 16 KB, V_j 16 KB (resident), Q ring 2 x 16 KB, dO ring 2 x 16 KB, P 32 KB, dS 32 KB (each two sub-tiles of 64 keys,
// [128 q rows][128 B]).  TMEM (512 columns, one CTA per SM): S [0,128), dP [128,256), dV [256,320), dK [320,384).
constexpr int kDkvSmem = 2 * kTile128 + 4 * kTile128 + 4 * kTile128 + 256;

__global__ void __launch_bounds__(kDkvThreads, 1) attention_bwd_dkv_tc_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTile128;
  uint8_t* sQ = sV + kTile128;         // 2 stages
  uint8_t* sdO = sQ + 2 * kTile128;    // 2 stages
  uint8_t* sP = sdO + 2 * kTile128;    // 2 sub-tiles
  uint8_t* sdS = sP + 2 * kTile128;    // 2 sub-tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * kTile128);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;     // 2
  uint64_t* q_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;
  uint64_t* pds_ready = bars + 6;  // 4 warp arrivals
  uint64_t* dkv_full = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x % p.tiles128, bh = blockIdx.x / p.tiles128;
  const int h = bh % p.heads, b = bh / p.heads;
  const int row0 = b * p.T;
  const int q_tiles = p.tiles128;
  const int keys_valid = min(128, p.T - kt * 128);
  const int nk = ((keys_valid + 15) / 16) * 16;                       // S / dP width of this key tile
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&p.tm128);
    tma_prefetch_desc(&p.tmdo);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(pds_ready, 8);
    mbar_init(dkv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();
  if (warp == 1) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * kTile128);
      tma_load_2d(sK, &p.tm128, kv_full, p.D + h * kDh, row0 + kt * 128);
      tma_load_2d(sV, &p.tm128, kv_full, 2 * p.D + h * kDh, row0 + kt * 128);
      for (int i = 0; i < q_tiles; ++i) {
        const int st = i & 1;
        if (i >= 2) mbar_wait(&q_empty[st], ((i >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&q_full[st], 2 * kTile128);
        tma_load_2d(sQ + st * kTile128, &p.tm128, &q_full[st], h * kDh, row0 + i * 128);
        tma_load_2d(sdO + st * kTile128, &p.tmdo, &q_full[st], h * kDh, row0 + i * 128);
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      const uint32_t idesc_s = make_idesc_bf16(128, nk, 0, 0);
      const uint32_t idesc_t = make_idesc_bf16(128, kDh, 1, 1);     // A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
      mbar_wait(kv_full, 0);
      for (int i = 0; i < q_tiles; ++i) {
        const int st = i & 1;
        const uint32_t q_addr = smem_u32(sQ + st * kTile128), do_addr = smem_u32(sdO + st * kTile128);
        const int rows_valid = min(128, p.T - i * 128);
        const int ksteps = (rows_valid + 15) / 16;                  // query rows (the reduction dimension) in steps of 16
        mbar_wait(&q_full[st], (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base, make_sdesc_sw128(q_addr + k * 32, 0, 1024), make_sdesc_sw128(k_addr + k * 32, 0, 1024), idesc_s,
                    k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base + 128, make_sdesc_sw128(do_addr + k * 32, 0, 1024), make_sdesc_sw128(v_addr + k * 32, 0, 1024),
                    idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(pds_ready, i & 1);
        tc_fence_after();
        for (int k = 0; k < ksteps; ++k)
          umma_bf16(tmem_base + 256, make_sdesc_sw128(p_addr + k * 2048, kTile128, 1024),
                    make_sdesc_sw128(do_addr + k * 2048, kTile128, 1024), idesc_t, (i | k) != 0 ? 1u : 0u);
        for (int k = 0; k < ksteps; ++k)
          umma_bf16(tmem_base + 320, make_sdesc_sw128(ds_addr + k * 2048, kTile128, 1024),
                    make_sdesc_sw128(q_addr + k * 2048, kTile128, 1024), idesc_t, (i | k) != 0 ? 1u : 0u);
        umma_commit(&q_empty[st]);
      }
      umma_commit(dkv_full);
    }
  } else {
    // EIGHT softmax warps: two per TMEM lane quarter (warp % 4), each taking one 64-key half of the tile -- with one CTA per
    // SM four warps left the SM at 9 % warp occupancy (ncu, profiles/r2_full_attnbwd.md)
    const int q = warp & 3, r = q * 32 + lane;
    const int half = (warp - 2) >> 2;
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl = p.scale_log2;
    for (int i = 0; i < q_tiles; ++i) {
      const int t = i * 128 + r;
      const bool row_valid = t < p.T;
      const int rows_valid = min(128, p.T - i * 128);
      const bool warp_needed = q * 32 < ((rows_valid + 15) / 16) * 16;   // rows of this warp are read by the dV / dK MMAs
      float lse = 0.f, delta = 0.f;
      if (row_valid) {
        lse = p.lse[(long long)bh * p.T + t];
        delta = p.delta[(long long)bh * p.T + t];
      }
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      if (warp_needed) {
#pragma unroll 1
        for (int c = half * 64; c < min(nk, half * 64 + 64); c += 16) {
          uint32_t s[16], d[16];
          tmem_ld_32x16(trow + uint32_t(c), s);
          tmem_ld_32x16(trow + 128 + uint32_t(c), d);
          tmem_ld_wait();
          uint32_t pp[8], pd[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(s[e]), sl, -lse));
            float p1 = ex2_approx(fmaf(__uint_as_float(s[e + 1]), sl, -lse));
            if (!row_valid || c + e >= keys_valid) p0 = 0.f;
            if (!row_valid || c + e + 1 >= keys_valid) p1 = 0.f;
            pp[e >> 1] = pack_bf16x2(p0, p1);
            pd[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(d[e]) - delta), p1 * (__uint_as_float(d[e + 1]) - delta));
          }
          const int sub = c >> 6, g = (c & 63) >> 3;
          uint8_t* prow = sP + sub * kTile128 + r * 128;
          uint8_t* drow = sdS + sub * kTile128 + r * 128;
          *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
          *reinterpret_cast<uint4*>(prow + (((g + 1) ^ (r & 7)) << 4)) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
          *reinterpret_cast<uint4*>(drow + ((g ^ (r & 7)) << 4)) = make_uint4(pd[0], pd[1], pd[2], pd[3]);
          *reinterpret_cast<uint4*>(drow + (((g + 1) ^ (r & 7)) << 4)) = make_uint4(pd[4], pd[5], pd[6], pd[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
    }
    mbar_wait(dkv_full, 0);
    tc_fence_after();
    const int key = kt * 128 + r;
    if (q * 32 < keys_valid) {   // warp-uniform (tcgen05.ld is warp-collective); the stores are predicated per key row
      // half 0 writes dV (TMEM [256,320)), half 1 writes dK (TMEM [320,384), times the softmax scale)
      __nv_bfloat16* dst = p.dqkv + (long long)(row0 + key) * (3 * p.D) + (half == 0 ? 2 : 1) * p.D + h * kDh;
      const float mul = half == 0 ? 1.f : p.scale;
#pragma unroll
      for (int c = 0; c < kDh; c += 16) {
        uint32_t v[16];
        tmem_ld_32x16(trow + (half == 0 ? 256u : 320u) + uint32_t(c), v);
        tmem_ld_wait();
        uint32_t pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) pv[e] = pack_bf16x2(__uint_as_float(v[2 * e]) * mul, __uint_as_float(v[2 * e + 1]) * mul);
        if (key < p.T) {
          reinterpret_cast<uint4*>(dst + c)[0] = make_uint4(pv[0], pv[1], pv[2], pv[3]);
          reinterpret_cast<uint4*>(dst + c)[1] = make_uint4(pv[4], pv[5], pv[6], pv[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn bwd_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
bool encode2d(EncodeTiledFn fn, CUtensorMap* m, const void* base, int cols, long long rows, int box_rows) {
  const cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(cols) * 2};
  const cuuint32_t box[2] = {kDh, cuuint32_t(box_rows)}, es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

cudaError_t launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                                 __nv_bfloat16* dqkv, float* stats, int B, int T, int heads, float scale, cudaStream_t s) {
  EncodeTiledFn fn = bwd_encode_fn();
  if (!fn || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(dctx) & 15) ||
      (reinterpret_cast<uintptr_t>(ctx) & 15) || (reinterpret_cast<uintptr_t>(dqkv) & 15))
    return cudaErrorInvalidValue;
  BwdParams p;
  const int D = heads * kDh;
  const long long rows = (long long)B * T;
  if (!encode2d(fn, &p.tm128, qkv, 3 * D, rows, 128) || !encode2d(fn, &p.tm64, qkv, 3 * D, rows, 64) ||
      !encode2d(fn, &p.tmdo, dctx, D, rows, 128))
    return cudaErrorInvalidValue;
  p.ctx = ctx; p.dctx = dctx; p.dqkv = dqkv;
  p.lse = stats;
  p.delta = stats + (long long)B * heads * T;
  p.T = T; p.D = D; p.heads = heads;
  p.tiles128 = (T + 127) / 128;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatsSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const unsigned grid = unsigned(B) * heads * p.tiles128;
  static int stages = -1;    // DP_ATTN_BWD_STAGES: bit mask of the launches to run (debug aid; default all three)
  if (stages < 0) { const char* v = getenv("DP_ATTN_BWD_STAGES"); stages = v ? atoi(v) : 7; }
  if (stages & 1) launch_k<attention_bwd_stats_kernel>(grid, kBThreads, kStatsSmem, s, p);
  if (stages & 2) launch_k<attention_bwd_dq_tc_kernel>(grid, kBThreads, kDqSmem, s, p);
  if (stages & 4) launch_k<attention_bwd_dkv_tc_kernel>(grid, kDkvThreads, kDkvSmem, s, p);
  return cudaGetLastError();
}

}  // namespace dp
