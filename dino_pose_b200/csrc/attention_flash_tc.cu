// Flash-style multi-head attention forward on tcgen05 / TMEM for ANY sequence length (HF modeling_dinov2.py:203-234,
// softmax(q k^T / sqrt(dh)) v, non-causal, no mask, dropout 0, head dim 64).  Used for the sequences the resident-K/V
// kernels of attention_tc.cu do not cover (T > 272: the 1025 tokens of a 448x448 image); replaces the mma.sync kernel.
//
// One CTA per (image, head, 128-query tile), two CTAs per SM (113 KB of shared memory, 256 TMEM columns each): while one
// CTA's softmax warps run, the other's MMAs and loads proceed.  Per 128-key tile j of K / V (2-stage TMA ring):
//     S_j[128 q, n_j] = Q K_j^T          tcgen05.mma, both operands K-major, accumulator in TMEM columns [0, 128)
//     online softmax                     4 warps, thread = query row: m' = max(m, rowmax S_j), P_j = exp2(c S_j - c m'),
//                                        l = l * a + rowsum P_j with a = exp2(c m - c m'); P_j -> bf16 K-major swizzled smem
//     O_j[128 q, 64] = P_j V_j           tcgen05.mma (A = P_j K-major, B = V_j MN-major), FRESH accumulator in TMEM [128, 192)
//     O = O * a + O_j                    in registers (64 fp32 per thread): no TMEM read-modify-write, no rescale hazard
// ctx = O / l.  The last key tile is n_j = (T mod 128 rounded up to 16) wide, so T = 1025 = 8 * 128 + 1 costs one N = 16
// S-MMA and one K = 16 PV-MMA for its last key instead of a full tile.  Rows of a K / V / Q box beyond T belong to the next
// image (or are zero-filled past the end of the tensor): their scores are masked, their outputs are not stored.
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

constexpr int kTile = 128;
constexpr int kDh = 64;
constexpr int kTileBytes = kTile * kDh * 2;                 // 16 KB
constexpr int kFThreads = 192;                              // warp 0 MMA issue, warp 1 TMEM alloc + TMA, warps 2..5 softmax
constexpr int kFSmemBytes = (1 + 2 + 2 + 2) * kTileBytes + 256;   // Q, K ring, V ring, P (two 64-key sub-tiles), barriers
constexpr uint32_t kFTmemCols = 256;
constexpr uint32_t kFColO = 128;

struct FlashParams {
  CUtensorMap tm;      // 2-D {3*D, B*T} bf16, box {64, 128}, 128B swizzle
  __nv_bfloat16* ctx;  // [B*T, D]
  int T, D, heads, q_tiles, kv_tiles, n_last;
  float scale_log2;
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Online-softmax step of one thread (= one query row) on the S tile in TMEM: returns the rescale factor of the running
// output, updates the running maximum `ms` (in units of c = scale * log2 e) and row sum `l`, writes P_j (bf16) to the
// K-major swizzled operand tiles.  MASKED = the last key tile (keys >= `valid` do not exist); the other tiles carry no
// per-element predicate.
template <bool MASKED>
__device__ __forceinline__ float softmax_tile(uint32_t trow, uint8_t* sP, int r, int nj, int valid, float sl, float& ms,
                                              float& l) {
  float m = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < nj; c += 32) {
    if (nj - c >= 32) {
      uint32_t v[32];
      tmem_ld_32x32(trow + uint32_t(c), v);
      tmem_ld_wait();
      if constexpr (MASKED) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) m = fmax3(m, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      }
    } else {
      uint32_t v[16];
      tmem_ld_32x16(trow + uint32_t(c), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (!MASKED || c + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
    }
  }
  const float ms_new = fmaxf(ms, m * sl);
  const float alpha = (ms == -INFINITY) ? 0.f : ex2_approx(ms - ms_new);
  ms = ms_new;
  float sum = 0.f;
#pragma unroll 1
  for (int c = 0; c < nj; c += 32) {
    uint8_t* tile = sP + (c >> 6) * kTileBytes + r * 128;
    const int g0 = (c & 63) >> 3;      // first 16-byte chunk (8 keys) of this 32-key group inside the 64-key sub-tile
    if (nj - c >= 32) {
      uint32_t v[32];
      tmem_ld_32x32(trow + uint32_t(c), v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = ex2_approx(fmaf(__uint_as_float(v[g * 8 + i]), sl, -ms));
          e[i] = (!MASKED || c + g * 8 + i < valid) ? x : 0.f;
        }
        sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
        *reinterpret_cast<uint4*>(tile + (((g0 + g) ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
      }
    } else {
      uint32_t v[16];
      tmem_ld_32x16(trow + uint32_t(c), v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          e[i] = (!MASKED || c + g * 8 + i < valid) ? ex2_approx(fmaf(__uint_as_float(v[g * 8 + i]), sl, -ms)) : 0.f;
        sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
        *reinterpret_cast<uint4*>(tile + (((g0 + g) ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
      }
    }
  }
  l = fmaf(l, alpha, sum);
  return alpha;
}

__global__ void __launch_bounds__(kFThreads, 2) attention_flash_tc_kernel(const __grid_constant__ FlashParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;            // 2 stages
  uint8_t* sV = sK + 2 * kTileBytes;        // 2 stages
  uint8_t* sP = sV + 2 * kTileBytes;        // [128 q x 128 keys] = 2 sub-tiles of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // 2
  uint64_t* kv_empty = bars + 3;      // 2
  uint64_t* s_full = bars + 5;        // 1
  uint64_t* p_ready = bars + 6;       // 1 (4 warp arrivals)
  uint64_t* o_full = bars + 7;        // 1
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads, b = bh / p.heads;
  const int row0 = b * p.T;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&p.tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, kFTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();

  if (warp == 1) {
    if (elect_one()) {
      // ---- TMA producer: Q tile, then the K / V tiles through the 2-stage ring
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_2d(sQ, &p.tm, q_full, h * kDh, row0 + qt * kTile);
      for (int j = 0; j < p.kv_tiles; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&kv_empty[st], ((j >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTileBytes);
        tma_load_2d(sK + st * kTileBytes, &p.tm, &kv_full[st], p.D + h * kDh, row0 + j * kTile);
        tma_load_2d(sV + st * kTileBytes, &p.tm, &kv_full[st], 2 * p.D + h * kDh, row0 + j * kTile);
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      const uint32_t idesc_pv = make_idesc_bf16(kTile, kDh, 0, 1);
      mbar_wait(q_full, 0);
      for (int j = 0; j < p.kv_tiles; ++j) {
        const int st = j & 1;
        const int nj = (j == p.kv_tiles - 1) ? p.n_last : kTile;
        const uint32_t k_addr = smem_u32(sK + st * kTileBytes), v_addr = smem_u32(sV + st * kTileBytes);
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        // S_j: the S columns are free -- the softmax warps finished reading S_{j-1} before they signalled p_ready(j-1),
        // which this thread waited for before issuing PV_{j-1}
        const uint32_t idesc_s = make_idesc_bf16(kTile, nj, 0, 0);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base, make_sdesc_sw128(q_addr + k * 32, 0, 1024), make_sdesc_sw128(k_addr + k * 32, 0, 1024), idesc_s,
                    k != 0 ? 1u : 0u);
        umma_commit(s_full);
        // PV_j once P_j is in shared memory (the softmax warps consumed O_{j-1} before starting on S_j)
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const int ksteps = nj / 16;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adesc = make_sdesc_sw128(p_addr + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024);
          const uint64_t bdesc = make_sdesc_sw128(v_addr + k * 2048, kTileBytes, 1024);
          umma_bf16(tmem_base + kFColO, adesc, bdesc, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);   // K_j / V_j are free once these MMAs have completed
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int r = q * 32 + lane;            // query row inside the tile
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl = p.scale_log2;
    const bool warp_active = qt * kTile + q * 32 < p.T;   // warp-uniform: a warp whose 32 rows are all past T only hand-shakes
    float o[kDh];
#pragma unroll
    for (int i = 0; i < kDh; ++i) o[i] = 0.f;
    float ms = -INFINITY, l = 0.f;          // running maximum (already multiplied by c = scale * log2 e) and row sum
    for (int j = 0; j < p.kv_tiles; ++j) {
      const bool last = j == p.kv_tiles - 1;
      const int nj = last ? p.n_last : kTile;
      const int valid = last ? p.T - j * kTile : kTile;   // keys of this tile that exist
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float alpha = 1.f;
      if (warp_active)
        alpha = last ? softmax_tile<true>(trow, sP, r, nj, valid, sl, ms, l) : softmax_tile<false>(trow, sP, r, nj, valid, sl, ms, l);
      fence_proxy_async_smem();   // make the st.shared of P visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // rescale the running output while PV_j runs, then add the tile's contribution
      if (warp_active && alpha != 1.f) {
#pragma unroll
        for (int i = 0; i < kDh; ++i) o[i] *= alpha;
      }
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      if (warp_active) {
#pragma unroll
        for (int c = 0; c < kDh; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(trow + kFColO + uint32_t(c), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
        }
      }
      tc_fence_before();   // S_j and O_j are consumed: the next tile's MMAs may overwrite them
    }
    const int t = qt * kTile + r;
    if (warp_active && t < p.T) {
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = p.ctx + (long long)(row0 + t) * p.D + h * kDh;   // 128 contiguous bytes per query row
#pragma unroll
      for (int c = 0; c < kDh / 8; ++c)
        reinterpret_cast<uint4*>(dst)[c] = make_uint4(pack_bf16x2(o[c * 8] * inv, o[c * 8 + 1] * inv),
                                                      pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv),
                                                      pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv),
                                                      pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kFTmemCols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn flash_encode_fn() {
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

}  // namespace

cudaError_t launch_attention_flash_tc(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, int B, int T, int heads, float scale,
                                      cudaStream_t s) {
  EncodeTiledFn fn = flash_encode_fn();
  if (!fn || (reinterpret_cast<uintptr_t>(qkv) & 15) || T < 1) return cudaErrorInvalidValue;
  FlashParams p;
  const int D = heads * kDh;
  const cuuint64_t dims[2] = {cuuint64_t(3 * D), cuuint64_t(B) * cuuint64_t(T)};
  const cuuint64_t strides[1] = {cuuint64_t(3 * D) * 2};
  const cuuint32_t box[2] = {kDh, kTile}, es[2] = {1, 1};
  if (fn(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(qkv), dims, strides, box, es,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  p.ctx = ctx;
  p.T = T; p.D = D; p.heads = heads;
  p.q_tiles = (T + kTile - 1) / kTile;
  p.kv_tiles = (T + kTile - 1) / kTile;
  const int rem = T - (p.kv_tiles - 1) * kTile;          // 1 .. 128 keys in the last tile
  p.n_last = ((rem + 15) / 16) * 16;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  launch_k<attention_flash_tc_kernel>(unsigned(B) * heads * p.q_tiles, kFThreads, kFSmemBytes, s, p);
  return cudaGetLastError();
}

cudaError_t launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, int B, int T, int heads, float scale,
                                cudaStream_t s);

// Entry point behind dp_attention_fwd: the resident-K/V tcgen05 kernels for short sequences (attention_tc.cu: T <= 272,
// i.e. the 257 tokens of a 224x224 image), the flash-style tcgen05 kernel above for every other length.
// DP_ATTN_FLASH=1 forces the flash kernel for all lengths (A/B: 37.9 vs 23.9 us at T = 257, batch 64).
cudaError_t launch_attention_fwd(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, int B, int T, int heads, float scale,
                                 cudaStream_t s) {
  static int force_flash = -1;
  if (force_flash < 0) { const char* v = getenv("DP_ATTN_FLASH"); force_flash = v ? atoi(v) : 0; }
  if (!force_flash) {
    cudaError_t e = launch_attention_tc(qkv, ctx, B, T, heads, scale, s);
    if (e != cudaErrorNotSupported) return e;
  }
  return launch_attention_flash_tc(qkv, ctx, B, T, heads, scale, s);
}

}  // namespace dp
