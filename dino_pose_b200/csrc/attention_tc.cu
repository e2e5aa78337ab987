// Multi-head attention forward on tcgen05 / TMEM for short sequences (T <= 272: the 257 tokens of a 224x224 image).
// HF modeling_dinov2.py:203-234: softmax(q k^T / sqrt(dh)) v, non-causal, no mask, dropout 0, head dim 64.
//
// One CTA per (image, head); the whole K and V of that head stay in shared memory (TMA, 128B swizzle):
//     S[128 q, Tk] = Q_tile K^T      tcgen05.mma, both operands K-major, N = 256 (+16), accumulator in TMEM
//     P = exp2(S*c - max*c)          4 warps, thread = query row, two passes over the TMEM row (max, then exp + row sum),
//                                    P written as bf16 into a K-major 128B-swizzled shared-memory operand
//     O[128 q, 64] = P V             tcgen05.mma, A = P (K-major), B = V (MN-major: keys are the reduction dim)
//     ctx = O / rowsum               TMEM -> registers -> 128 contiguous bytes per query row
// Query tiles of 128 rows (3 per head at T = 257; the last holds the single remaining row).  One thread issues TMA and
// MMA; the S-MMA of tile i+1 is issued right behind the PV-MMA of tile i so it runs under the epilogue of tile i.
// Rows of the K / V boxes beyond T belong to the next image (or are zero-filled past the end of the tensor): their
// scores are masked to -inf, so their P is exactly 0.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "launch.cuh"
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

namespace dp {
namespace {

constexpr int kTile = 128;          // query rows per tile, key rows per TMA box
constexpr int kDh = 64;
constexpr int kTileBytes = kTile * kDh * 2;   // 16 KB
constexpr int kMaxKeyTiles = 3;     // 384 key rows staged, 272 used at most
constexpr int kMaxPTiles = 5;       // 5 x 64 keys
constexpr int kThreads = 192;       // warp 0: TMA + MMA issue, warp 1: TMEM alloc, warps 2..5: softmax / epilogue
constexpr int kSmemBytes = (2 * kMaxKeyTiles + 2 + kMaxPTiles) * kTileBytes + 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO = 320;     // O accumulator columns [320, 384)

struct AttnParams {
  CUtensorMap tm;      // 2-D {3*D, B*T} bf16, box {64, 128}, 128B swizzle
  __nv_bfloat16* ctx;  // [B*T, D]
  int T, D, heads, tk_pad, q_tiles;
  float scale_log2;
};

__global__ void __launch_bounds__(kThreads, 1) attention_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + kMaxKeyTiles * kTileBytes;
  uint8_t* sQ = sV + kMaxKeyTiles * kTileBytes;       // 2 buffers
  uint8_t* sP = sQ + 2 * kTileBytes;                   // 5 tiles [128 q x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kMaxPTiles * kTileBytes);
  uint64_t* kv_full = bars;          // 1
  uint64_t* q_full = bars + 1;       // 2
  uint64_t* s_full = bars + 3;       // 1
  uint64_t* p_ready = bars + 4;      // 1 (4 warp arrivals)
  uint64_t* o_full = bars + 5;       // 1
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.heads, b = blockIdx.x / p.heads;
  const int row0 = b * p.T;
  const int key_tiles = (p.tk_pad + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&p.tm);
    mbar_init(kv_full, 1);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();   // programmatic dependent launch: no global-memory access above this line

  if (warp == 0) {
    if (elect_one()) {
      // ---- K, V of this (image, head), and the first Q tile
      mbar_arrive_expect_tx(kv_full, 2 * key_tiles * kTileBytes);
      for (int t = 0; t < key_tiles; ++t) {
        tma_load_2d(sK + t * kTileBytes, &p.tm, kv_full, p.D + h * kDh, row0 + t * kTile);
        tma_load_2d(sV + t * kTileBytes, &p.tm, kv_full, 2 * p.D + h * kDh, row0 + t * kTile);
      }
      mbar_arrive_expect_tx(&q_full[0], kTileBytes);
      tma_load_2d(sQ, &p.tm, &q_full[0], h * kDh, row0);
      const int n1 = p.tk_pad < 256 ? p.tk_pad : 256, n2 = p.tk_pad - n1;   // S = [N = n1] (+ [N = n2 = 16])
      const uint32_t idesc_s1 = make_idesc_bf16(kTile, n1, 0, 0);
      const uint32_t idesc_s2 = make_idesc_bf16(kTile, n2 > 0 ? n2 : 16, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(kTile, kDh, 0, 1);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      mbar_wait(kv_full, 0);
      auto issue_s = [&](int i) {
        mbar_wait(&q_full[i & 1], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + (i & 1) * kTileBytes);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) {
          const uint64_t adesc = make_sdesc_sw128(q_addr + k * 32, 0, 1024);
          umma_bf16(tmem_base, adesc, make_sdesc_sw128(k_addr + k * 32, 0, 1024), idesc_s1, k != 0 ? 1u : 0u);
          if (n2 > 0)
            umma_bf16(tmem_base + 256, adesc, make_sdesc_sw128(k_addr + 2 * kTileBytes + k * 32, 0, 1024), idesc_s2,
                      k != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      };
      issue_s(0);
      for (int i = 0; i < p.q_tiles; ++i) {
        if (i + 1 < p.q_tiles) {   // prefetch the next Q tile (its buffer was last read by S-MMA i-1, long complete)
          mbar_arrive_expect_tx(&q_full[(i + 1) & 1], kTileBytes);
          tma_load_2d(sQ + ((i + 1) & 1) * kTileBytes, &p.tm, &q_full[(i + 1) & 1], h * kDh, row0 + (i + 1) * kTile);
        }
        mbar_wait(p_ready, i & 1);
        tc_fence_after();
        const int ksteps = p.tk_pad / 16;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adesc = make_sdesc_sw128(p_addr + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024);
          const uint64_t bdesc = make_sdesc_sw128(v_addr + k * 2048, kTileBytes, 1024);
          umma_bf16(tmem_base + kColO, adesc, bdesc, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        if (i + 1 < p.q_tiles) issue_s(i + 1);   // S was fully consumed before p_ready(i)
      }
    }
  } else if (warp >= 2) {
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int r = q * 32 + lane;            // query row inside the tile
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl = p.scale_log2;
    const int nchunks = p.tk_pad / 16;
    for (int i = 0; i < p.q_tiles; ++i) {
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      const bool warp_active = i * kTile + q * 32 < p.T;   // warp-uniform: a warp whose 32 rows are all past T skips
      float sum = 1.f;
      if (warp_active) {
        const int nfull = p.tk_pad >> 6, ntail = (p.tk_pad & 63) >> 4;   // 64-key blocks + 16-key chunks
        // pass 1: row maximum over the valid keys (64 columns per tcgen05.wait::ld)
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < nfull; ++c) {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(trow + uint32_t(c * 64), v0);
          tmem_ld_32x32(trow + uint32_t(c * 64 + 32), v1);
          tmem_ld_wait();
          const bool all_valid = c * 64 + 64 <= p.T;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (all_valid || c * 64 + j < p.T) m = fmaxf(m, __uint_as_float(v0[j]));
            if (all_valid || c * 64 + 32 + j < p.T) m = fmaxf(m, __uint_as_float(v1[j]));
          }
        }
#pragma unroll 1
        for (int c = 0; c < ntail; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(trow + uint32_t(nfull * 64 + c * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nfull * 64 + c * 16 + j < p.T) m = fmaxf(m, __uint_as_float(v[j]));
        }
        const float ms = m * sl;
        // pass 2: P = exp2(s*c - m*c) as bf16 into the K-major swizzled operand tiles, row sum in fp32
        sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < nfull; ++c) {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(trow + uint32_t(c * 64), v0);
          tmem_ld_32x32(trow + uint32_t(c * 64 + 32), v1);
          tmem_ld_wait();
          const bool all_valid = c * 64 + 64 <= p.T;
          uint8_t* tile = sP + c * kTileBytes + r * 128;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int g = 0; g < 8; ++g) {        // 8 x 16-byte chunks (8 keys each) of this row of the P tile
            const uint32_t* v = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = ex2_approx(fmaf(__uint_as_float(v[j]), sl, -ms));
              e[j] = (all_valid || c * 64 + g * 8 + j < p.T) ? x : 0.f;
            }
            s0 += (e[0] + e[1]) + (e[2] + e[3]);
            s1 += (e[4] + e[5]) + (e[6] + e[7]);
            *reinterpret_cast<uint4*>(tile + ((g ^ (r & 7)) << 4)) =
                make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
          }
          sum += s0 + s1;
        }
#pragma unroll 1
        for (int c = 0; c < ntail; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(trow + uint32_t(nfull * 64 + c * 16), v);
          tmem_ld_wait();
          uint8_t* tile = sP + nfull * kTileBytes + r * 128;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              e[j] = (nfull * 64 + c * 16 + g * 8 + j < p.T) ? ex2_approx(fmaf(__uint_as_float(v[g * 8 + j]), sl, -ms)) : 0.f;
            sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
            *reinterpret_cast<uint4*>(tile + (((c * 2 + g) ^ (r & 7)) << 4)) =
                make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
          }
        }
      }
      fence_proxy_async_smem();   // make the st.shared of P visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // epilogue of this tile: O / sum -> bf16 -> ctx
      mbar_wait(o_full, i & 1);
      tc_fence_after();
      if (!warp_active) continue;
      const float inv = 1.0f / sum;
      const int t = i * kTile + r;
      __nv_bfloat16* dst = p.ctx + (long long)(row0 + t) * p.D + h * kDh;
#pragma unroll
      for (int c = 0; c < kDh / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(trow + kColO + uint32_t(c * 16), v);
        tmem_ld_wait();
        if (t < p.T) {
          uint4 a, bq;
          a.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
          a.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
          a.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
          a.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
          bq.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
          bq.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
          bq.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
          bq.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
          reinterpret_cast<uint4*>(dst + c * 16)[0] = a;
          reinterpret_cast<uint4*>(dst + c * 16)[1] = bq;
        }
      }
      tc_fence_before();   // O and S of this tile are consumed before the next tile's MMAs may overwrite them
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// T == 257 (CLS + 16x16 patches of a 224x224 image): the shape of every benchmark configuration at 224^2.
// One CTA per (image, head, 128-query tile), TWO CTAs per SM (96 KB shared memory, 256 TMEM columns each), so the
// TMA / MMA / softmax / epilogue phases of two tiles interleave on one SM.  257 = 2*128 + 1 on both axes:
//   * the 257th KEY is a rank-1 correction on the CUDA cores: s_x = q.k_256 joins the row max / sum, and
//     p_x * v_256 is added to O in the epilogue -- the tensor-core part is exactly N = 256 keys;
//   * the 257th QUERY is computed entirely on the CUDA cores by the otherwise idle warp 1 of the second tile's CTA.
// Shared memory is re-used inside the tile: P (4 tiles of 64 keys) overwrites K (dead after the S-MMA, tiles 0-1) and
// Q (tile 2); O overwrites TMEM columns [0, 64) of S.
constexpr int kSmem257 = 6 * kTileBytes + 1024 + 2048 + 2048;   // K0 K1 V0 V1 Q P3 + extra rows / barriers + max / sum exchange + extra-query scores

struct Attn257Params {
  CUtensorMap tm;
  const __nv_bfloat16* qkv;
  __nv_bfloat16* ctx;
  long long* trace;   // debug (DP_ATTN_TRACE=1): timestamps of CTA 1 (second query tile of image 0, head 0)
  int D, heads;
  float scale_log2;
};

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

__device__ __forceinline__ void unpack8_bf16(const uint4& t, float (&f)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
    f[2 * k] = __low2float(hh);
    f[2 * k + 1] = __high2float(hh);
  }
}

// HALVES = 2: EIGHT softmax warps, two threads per query row (keys [0,128) and [128,256)); the row maximum is exchanged
// through shared memory between the two passes, the row sum before the epilogue (each half normalises and stores 32 of
// the 64 output dims).  The per-thread serial chain -- the exposed phase in the timeline of the HALVES = 1 version --
// halves, and the SM holds 16 instead of 8 warps that issue MUFU / FMA work.
template <int HALVES>
__global__ void __launch_bounds__(64 + 128 * HALVES, 2) attention_tc257_kernel(const __grid_constant__ Attn257Params p) {
  constexpr int T = 257;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;                          // 2 tiles; later P tiles 0, 1
  uint8_t* sV = sK + 2 * kTileBytes;           // 2 tiles
  uint8_t* sQ = sV + 2 * kTileBytes;           // 1 tile; later P tile 2
  uint8_t* sP3 = sQ + kTileBytes;              // P tile 3
  float* xk = reinterpret_cast<float*>(sP3 + kTileBytes);   // k_256, v_256, q_256 as fp32 [64] each
  float* xv = xk + 64;
  float* xq = xv + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xq + 64);
  uint64_t* ld_full = bars;
  uint64_t* s_full = bars + 1;
  uint64_t* k_free = bars + 2;
  uint64_t* p_ready = bars + 3;
  uint64_t* o_full = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);
  float* ex_max = reinterpret_cast<float*>(sP3 + kTileBytes + 1024);   // [2 halves][128 rows]
  float* ex_sum = ex_max + 256;                                          // [2 halves][128 rows]
  float* xs = ex_sum + 256;                                              // scores of the extra query row [256 + pad]
  __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(xs + 272);        // its unnormalised probabilities [256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x & 1;
  const int h = (blockIdx.x >> 1) % p.heads, b = (blockIdx.x >> 1) / p.heads;
  const int row0 = b * T;
  const int ld = 3 * p.D;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&p.tm);
    mbar_init(ld_full, 1);
    mbar_init(s_full, 1);
    mbar_init(k_free, 1);
    mbar_init(p_ready, 4 * HALVES);
    mbar_init(o_full, 1);
    fence_barrier_init();
    pdl_grid_sync();   // programmatic dependent launch: every thread passes this before its first global-memory access
    // loads are issued before the CTA-wide sync: their latency overlaps the TMEM allocation and the extra-row loads
    mbar_arrive_expect_tx(ld_full, 5 * kTileBytes);
    tma_load_2d(sQ, &p.tm, ld_full, h * kDh, row0 + qt * kTile);
    tma_load_2d(sK, &p.tm, ld_full, p.D + h * kDh, row0);
    tma_load_2d(sK + kTileBytes, &p.tm, ld_full, p.D + h * kDh, row0 + kTile);
    tma_load_2d(sV, &p.tm, ld_full, 2 * p.D + h * kDh, row0);
    tma_load_2d(sV + kTileBytes, &p.tm, ld_full, 2 * p.D + h * kDh, row0 + kTile);
  } else if (warp != 1) {
    pdl_grid_sync();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
    pdl_grid_sync();
    // the 257th token's k / v / q rows of this head (plain loads: 3 x 128 bytes)
    const __nv_bfloat16* xrow = p.qkv + (long long)(row0 + 256) * ld + h * kDh;
    for (int i = lane; i < 64; i += 32) {
      xq[i] = __bfloat162float(xrow[i]);
      xk[i] = __bfloat162float(xrow[p.D + i]);
      xv[i] = __bfloat162float(xrow[2 * p.D + i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  long long* tr = (p.trace != nullptr && blockIdx.x == 1) ? p.trace : nullptr;
  if (tr && threadIdx.x == 64) tr[0] = clock64();

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kTile, 256, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, kDh, 0, 1);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), p3_addr = smem_u32(sP3);
      mbar_wait(ld_full, 0);
      tc_fence_after();
      if (tr) tr[1] = clock64();
#pragma unroll
      for (int k = 0; k < kDh / 16; ++k)
        umma_bf16(tmem_base, make_sdesc_sw128(q_addr + k * 32, 0, 1024), make_sdesc_sw128(k_addr + k * 32, 0, 1024), idesc_s,
                  k != 0 ? 1u : 0u);
      umma_commit(s_full);
      mbar_wait(p_ready, 0);
      tc_fence_after();
      if (tr) tr[6] = clock64();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int pt = k >> 2;   // P tile: 0, 1 live where K was, 2 where Q was, 3 in its own buffer
        const uint32_t pa = (pt < 2 ? k_addr + pt * kTileBytes : (pt == 2 ? q_addr : p3_addr)) + (k & 3) * 32;
        umma_bf16(tmem_base, make_sdesc_sw128(pa, 0, 1024), make_sdesc_sw128(v_addr + k * 2048, kTileBytes, 1024), idesc_pv,
                  k != 0 ? 1u : 0u);
      }
      umma_commit(o_full);
    }
  } else if (warp == 1) {
    if (qt == 0) {
      if (lane == 0) mbar_arrive(k_free);
    } else {
      // ---- query row 256 with warp-level MMAs (mma.sync m16n8k16: a 16-row A tile whose row 0 is q_256, the other 15
      // rows zero).  The first version did this row with FMAs on the CUDA cores (2 x 257 x 64 MACs by one warp): under
      // the issue pressure of the softmax warps it finished at 19 k cycles against 13.8 k for the rest of the tile
      // (DP_ATTN_TRACE), i.e. it set the lifetime of every second CTA.
      mbar_wait(ld_full, 0);
      const bool g0 = (lane >> 2) == 0;             // lanes 0..3 own row 0 of the A / accumulator fragments
      const int c2 = 2 * (lane & 3);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const float sl = p.scale_log2;
      uint32_t qa0[4], qa2[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        qa0[ks] = g0 ? pack_bf16x2(xq[ks * 16 + c2], xq[ks * 16 + c2 + 1]) : 0u;
        qa2[ks] = g0 ? pack_bf16x2(xq[ks * 16 + 8 + c2], xq[ks * 16 + 8 + c2 + 1]) : 0u;
      }
      // pass 1: scores of the 256 tile keys, 64 at a time, row 0 parked in shared memory
#pragma unroll 1
      for (int kc = 0; kc < 4; ++kc) {
        float sacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
        const uint32_t tile = k_addr + (kc >> 1) * kTileBytes;
        const int rbase = (kc & 1) * 64;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t af[4] = {qa0[ks], 0u, qa2[ks], 0u};
#pragma unroll
          for (int nb = 0; nb < 8; nb += 2) {
            const int m = lane >> 3;
            const int row = rbase + (nb + (m >> 1)) * 8 + (lane & 7);
            const int chunk = ks * 2 + (m & 1);
            uint32_t b0, b1, b2, b3;
            ldsm_x4(tile + row * 128 + ((chunk ^ (lane & 7)) << 4), b0, b1, b2, b3);
            mma_bf16(sacc[nb], af, b0, b1);
            mma_bf16(sacc[nb + 1], af, b2, b3);
          }
        }
        if (g0) {
#pragma unroll
          for (int nb = 0; nb < 8; ++nb)
            *reinterpret_cast<float2*>(xs + kc * 64 + nb * 8 + c2) = make_float2(sacc[nb][0], sacc[nb][1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(k_free);   // K has been read: the softmax warps may overwrite it with P
      if (tr && lane == 0) tr[12] = clock64();
      // softmax over the 257 scores: lane owns keys lane + 32*i (i < 8), lane 0 also key 256
      float sc[9];
#pragma unroll
      for (int i = 0; i < 8; ++i) sc[i] = xs[lane + 32 * i];
      {
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 64; ++e) acc = fmaf(xk[e], xq[e], acc);
        sc[8] = (lane == 0) ? acc : -INFINITY;
      }
      float m = sc[0];
#pragma unroll
      for (int i = 1; i < 9; ++i) m = fmaxf(m, sc[i]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        sc[i] = ex2_approx((sc[i] - m) * sl);   // exp2(-inf) = 0 for the padding lanes of i = 8
        sum += sc[i];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float px = __shfl_sync(0xffffffffu, sc[8], 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) xp[lane + 32 * i] = __float2bfloat16_rn(sc[i]);
      __syncwarp();
      // pass 2: O = P V, 16 keys per step
      float oacc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
#pragma unroll 4
      for (int kk = 0; kk < 16; ++kk) {
        uint32_t af[4] = {0u, 0u, 0u, 0u};
        if (g0) {
          af[0] = *reinterpret_cast<const uint32_t*>(xp + kk * 16 + c2);
          af[2] = *reinterpret_cast<const uint32_t*>(xp + kk * 16 + 8 + c2);
        }
        const uint32_t tile = v_addr + (kk >> 3) * kTileBytes;
        const int rbase = (kk & 7) * 16;
#pragma unroll
        for (int nb = 0; nb < 8; nb += 2) {
          const int mm = lane >> 3;
          const int row = rbase + (mm & 1) * 8 + (lane & 7);
          const int chunk = nb + (mm >> 1);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(tile + row * 128 + ((chunk ^ (lane & 7)) << 4), b0, b1, b2, b3);
          mma_bf16(oacc[nb], af, b0, b1);
          mma_bf16(oacc[nb + 1], af, b2, b3);
        }
      }
      if (g0) {
        const float inv = 1.0f / sum;
        __nv_bfloat16* dst = p.ctx + (long long)(row0 + 256) * p.D + h * kDh;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const int d = nb * 8 + c2;
          *reinterpret_cast<uint32_t*>(dst + d) =
              pack_bf16x2(fmaf(px, xv[d], oacc[nb][0]) * inv, fmaf(px, xv[d + 1], oacc[nb][1]) * inv);
        }
      }
    }
  } else if constexpr (HALVES == 2) {
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // 0: keys [0,128) -> P tiles 0, 1 (over K);  1: keys [128,256) -> P tiles 2, 3
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl = p.scale_log2;
    mbar_wait(s_full, 0);
    tc_fence_after();
    if (tr && threadIdx.x == 64) tr[2] = clock64();
    float sx = 0.f;
    {
      const uint8_t* qrow = sQ + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float f[8];
        unpack8_bf16(*reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) sx = fmaf(f[e], xk[c * 8 + e], sx);
      }
    }
    float m = sx;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = 2 * half + cc;
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(trow + uint32_t(c * 64), v0);
      tmem_ld_32x32(trow + uint32_t(c * 64 + 32), v1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
    }
    ex_max[half * 128 + r] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 softmax warps: maxima exchanged, every read of the Q tile done
    m = fmaxf(m, ex_max[(half ^ 1) * 128 + r]);
    const float ms = m * sl;
    const float px = ex2_approx(fmaf(sx, sl, -ms));
    float sum = half == 0 ? px : 0.f;
    if (tr && threadIdx.x == 64) tr[3] = clock64();
    if (half == 0) {
      if (tr && threadIdx.x == 64) tr[4] = clock64();
      mbar_wait(k_free, 0);     // P tiles 0, 1 overwrite K, which the extra-query warp may still be reading
      if (tr && threadIdx.x == 64) tr[5] = clock64();
    }
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = 2 * half + cc;
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(trow + uint32_t(c * 64), v0);
      tmem_ld_32x32(trow + uint32_t(c * 64 + 32), v1);
      tmem_ld_wait();
      uint8_t* tile = (c < 2 ? sK + c * kTileBytes : (c == 2 ? sQ : sP3)) + r * 128;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t* v = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = ex2_approx(fmaf(__uint_as_float(v[j]), sl, -ms));
        s0 += (e[0] + e[1]) + (e[2] + e[3]);
        s1 += (e[4] + e[5]) + (e[6] + e[7]);
        *reinterpret_cast<uint4*>(tile + ((g ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
      }
      sum += s0 + s1;
    }
    ex_sum[half * 128 + r] = sum;     // read by the other half after o_full (ordered through p_ready -> MMA -> o_full)
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    if (tr && threadIdx.x == 64) tr[7] = clock64();
    mbar_wait(o_full, 0);
    tc_fence_after();
    if (tr && threadIdx.x == 64) tr[8] = clock64();
    sum += ex_sum[(half ^ 1) * 128 + r];
    const float inv = 1.0f / sum;
    __nv_bfloat16* dst = p.ctx + (long long)(row0 + qt * kTile + r) * p.D + h * kDh + half * 32;
    uint32_t v0[32];
    tmem_ld_32x32(trow + uint32_t(half * 32), v0);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t* v = v0 + g * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(px, xv[half * 32 + g * 8 + j], __uint_as_float(v[j])) * inv;
      reinterpret_cast<uint4*>(dst)[g] =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl = p.scale_log2;
    mbar_wait(s_full, 0);
    tc_fence_after();
    if (tr && threadIdx.x == 64) tr[2] = clock64();
    // score against the 257th key: s_x = q_r . k_256 (Q row r from the swizzled tile, k_256 broadcast)
    float sx = 0.f;
    {
      const uint8_t* qrow = sQ + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float f[8];
        unpack8_bf16(*reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) sx = fmaf(f[e], xk[c * 8 + e], sx);
      }
    }
    float m = sx;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(trow + uint32_t(c * 64), v0);
      tmem_ld_32x32(trow + uint32_t(c * 64 + 32), v1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
    }
    const float ms = m * sl;
    const float px = ex2_approx(fmaf(sx, sl, -ms));
    float sum = px;
    if (tr && threadIdx.x == 64) tr[3] = clock64();
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int c = (cc + 2) & 3;   // P tiles 2, 3 first: tiles 0, 1 overwrite K, which the extra-query warp may still read
      if (cc == 2) {
        if (tr && threadIdx.x == 64) tr[4] = clock64();
        mbar_wait(k_free, 0);
        if (tr && threadIdx.x == 64) tr[5] = clock64();
      }
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(trow + uint32_t(c * 64), v0);
      tmem_ld_32x32(trow + uint32_t(c * 64 + 32), v1);
      tmem_ld_wait();
      uint8_t* tile = (c < 2 ? sK + c * kTileBytes : (c == 2 ? sQ : sP3)) + r * 128;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t* v = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = ex2_approx(fmaf(__uint_as_float(v[j]), sl, -ms));
        s0 += (e[0] + e[1]) + (e[2] + e[3]);
        s1 += (e[4] + e[5]) + (e[6] + e[7]);
        *reinterpret_cast<uint4*>(tile + ((g ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
      }
      sum += s0 + s1;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    if (tr && threadIdx.x == 64) tr[7] = clock64();
    mbar_wait(o_full, 0);
    tc_fence_after();
    if (tr && threadIdx.x == 64) tr[8] = clock64();
    const float inv = 1.0f / sum;
    uint32_t v0[32], v1[32];
    tmem_ld_32x32(trow, v0);
    tmem_ld_32x32(trow + 32, v1);
    tmem_ld_wait();
    // thread = row gives 16-byte stores 768 bytes apart (32 half-used sectors per instruction).  Stage the warp's
    // 32 x 64 bf16 block in the dead P-tile-3 buffer (PV has completed), swizzled like the operand tiles, and write it
    // out as four full 128-byte row segments per instruction.
    uint8_t* stage = sP3 + q * (32 * 128);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint32_t* v = g < 4 ? v0 + g * 8 : v1 + (g - 4) * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(px, xv[g * 8 + j], __uint_as_float(v[j])) * inv;
      *reinterpret_cast<uint4*>(stage + lane * 128 + ((g ^ (lane & 7)) << 4)) =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
    __syncwarp();
    {
      const int rr = lane >> 3, ch = lane & 7;
      __nv_bfloat16* dst = p.ctx + (long long)(row0 + qt * kTile + q * 32) * p.D + h * kDh + ch * 8;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + rr;
        *reinterpret_cast<uint4*>(dst + (long long)row * p.D) =
            *reinterpret_cast<const uint4*>(stage + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
  }
  if (tr && threadIdx.x == 64) tr[9] = clock64();
  if (tr && threadIdx.x == 32) tr[10] = clock64();
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 64) tr[11] = clock64();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn attn_encode_fn() {
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

// returns cudaErrorNotSupported when the shape is outside this kernel's range (caller falls back to the mma.sync kernel)
cudaError_t launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, int B, int T, int heads, float scale,
                                cudaStream_t s) {
  const int tk_pad = ((T + 15) / 16) * 16;
  if (tk_pad > 272 || T < 1) return cudaErrorNotSupported;
  EncodeTiledFn fn = attn_encode_fn();
  if (!fn || (reinterpret_cast<uintptr_t>(qkv) & 15)) return cudaErrorNotSupported;
  static int v1_only = -1;
  if (v1_only < 0) { const char* v = getenv("DP_ATTN_V1"); v1_only = v ? atoi(v) : 0; }
  if (T == 257 && !v1_only) {
    Attn257Params q;
    const int Dm = heads * kDh;
    const cuuint64_t dims[2] = {cuuint64_t(3 * Dm), cuuint64_t(B) * cuuint64_t(T)};
    const cuuint64_t strides[1] = {cuuint64_t(3 * Dm) * 2};
    const cuuint32_t box[2] = {kDh, kTile}, es[2] = {1, 1};
    if (fn(&q.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(qkv), dims, strides, box, es,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
    q.qkv = qkv; q.ctx = ctx; q.D = Dm; q.heads = heads;
    q.scale_log2 = scale * 1.4426950408889634f;
    q.trace = nullptr;
    static int trace_on = -1;
    if (trace_on < 0) { const char* v = getenv("DP_ATTN_TRACE"); trace_on = v ? atoi(v) : 0; }
    static long long* dbuf = nullptr;
    if (trace_on) {
      if (!dbuf) cudaMalloc(&dbuf, 16 * sizeof(long long));
      cudaMemsetAsync(dbuf, 0, 16 * sizeof(long long), s);
      q.trace = dbuf;
    }
    static bool attr2 = false;
    if (!attr2) {
      cudaError_t e = cudaFuncSetAttribute(attention_tc257_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem257);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_tc257_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem257);
      if (e != cudaSuccess) return e;
      attr2 = true;
    }
    // DP_ATTN_HALVES=2: eight softmax warps, two threads per row.  Measured equal to the four-warp version at the kernel
    // level (26.3 vs 26.6 us, gpurun_out/ab5): the exponential pass is bound by the MUFU rate of the SM, not by the
    // number of warps issuing it -> the simpler four-warp kernel stays the default
    static int halves = -1;
    if (halves < 0) { const char* v = getenv("DP_ATTN_HALVES"); halves = v ? atoi(v) : 1; }
    if (halves == 2)
      launch_k<attention_tc257_kernel<2>>(B * heads * 2, 64 + 128 * 2, kSmem257, s, q);
    else
      launch_k<attention_tc257_kernel<1>>(B * heads * 2, kThreads, kSmem257, s, q);
    if (trace_on) {   // debug only: synchronises
      long long h[16];
      cudaStreamSynchronize(s);
      cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[attn trace, cycles after CTA sync] ld_full %lld  s_full %lld  pass1 %lld  k_free wait %lld..%lld  "
              "p_ready(mma) %lld  p_done %lld  o_full %lld  epilogue %lld  warp1 scores %lld  warp1 end %lld  cta end %lld\n",
              h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0],
              h[9] - h[0], h[12] - h[0], h[10] - h[0], h[11] - h[0]);
    }
    return cudaGetLastError();
  }
  AttnParams p;
  const int D = heads * kDh;
  const cuuint64_t dims[2] = {cuuint64_t(3 * D), cuuint64_t(B) * cuuint64_t(T)};
  const cuuint64_t strides[1] = {cuuint64_t(3 * D) * 2};
  const cuuint32_t box[2] = {kDh, kTile}, es[2] = {1, 1};
  if (fn(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(qkv), dims, strides, box, es,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  p.ctx = ctx;
  p.T = T; p.D = D; p.heads = heads; p.tk_pad = tk_pad;
  p.q_tiles = (T + kTile - 1) / kTile;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  launch_k<attention_tc_kernel>(B * heads, kThreads, kSmemBytes, s, p);
  return cudaGetLastError();
}

}  // namespace dp
