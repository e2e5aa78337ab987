// Forward tcgen05 GEMM / implicit-conv kernel template (sm_100a) and its variant registry.
//
//   C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
// A is either a plain K-major matrix (2-D TMA) or an NHWC activation read as an IMPLICIT convolution: one
// 4-D TMA box {64 ch, bw, bh, bb} per (filter tap, 64-channel block), shifted by the tap offset, out-of-bounds
// pixels zero-filled by TMA (= the conv padding).  Persistent CTAs, one per SM, 320 threads:
//     warp 0      TMA producer (one elected lane)
//     warp 1      tcgen05.mma issuer (one elected lane), cta_group::1, 128 x BN x 16 per instruction
//     warps 2..9  epilogue: tcgen05.ld -> registers -> 32x32 smem transpose -> fused math -> coalesced stores
// smem ring of kStages {A 128x64, W BNx64} bf16 tiles (128B swizzle), double-buffered fp32 accumulators in TMEM
// (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// The epilogue is specialised at COMPILE time (template parameters OUT / ACT / MAP / OPT): the first version of
// this kernel tested every optional feature at run time inside the fully unrolled row loop, which made the
// kernel 12.7k SASS instructions (I-cache misses = 30 % of all stall samples, ncu profiles/r1b_*) and ~850
// instructions per 32x32 chunk.  A variant is compiled only for the feature sets the engine uses; the launcher
// picks the smallest compiled superset, and a run-time-everything variant exists for each tile width.
#pragma once
#include "gemm_tc.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace dp {

// ---- compile-time epilogue selectors
enum : int { EO_BF16 = 0, EO_F32 = 1, EO_RUNTIME = 2 };                 // output dtype
enum : int { EA_NONE = 0, EA_RELU = 1, EA_GELU = 2, EA_RUNTIME = 3 };   // activation
enum : int { EM_IDENTITY = 0, EM_PATCH = 1, EM_NCHW = 2, EM_SHUFFLE = 3, EM_RUNTIME = 4 };  // row map
enum : int {
  OP_SCALE = 1,      // per-column scale (folded eval BatchNorm)
  OP_LSRES = 2,      // LayerScale and / or fp32 residual
  OP_RES_BF16 = 4,   // bf16 residual
  OP_AUX_OUT = 8,    // bf16 side output: gelu'(v) when the activation is GELU, else v
  OP_AUX_IN = 16,    // multiply by aux_in (the saved gelu' of the forward)
  OP_STATS = 32,     // fused BatchNorm statistics
  OP_CONV = 64,      // implicit-conv row decoding (a_mode = 1)
  OP_ALL = 127,
  OP_TMA_OUT = 128   // bf16 identity-map output written with TMA tile stores (epilogue_tile_tma); not part of OP_ALL
};

struct GemmVariant {
  int bn, out, act, map, opt;
  int pair;   // 1: CTA-pair kernel (gemm_kernel2.cuh): the weight tensor map's box is bn/2 rows, grid = 2 x clusters
  cudaError_t (*launch)(const GemmParams&, int grid, cudaStream_t);
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 bf16 = 128 B = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kEpiWarps = 16;                     // four warps per TMEM lane quarter
constexpr int kEpiGroups = kEpiWarps / 4;         // column groups: warp (q, g) handles chunks c with c % kEpiGroups == g
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr int kStagingBytes = kEpiWarps * 32 * 32 * 4;   // one 32x32 fp32 transpose buffer per epilogue warp
constexpr int kSmemLimit = 232448;                // 227 KB per CTA

// [stages x {A 128x64, W BNx64}] [staging] [column statistics 2*BN fp32] [barriers]
template <int BN> struct GCfg {
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kFixedBytes = kStagingBytes + 2 * BN * 4 + 256;
  static constexpr int kFit = (kSmemLimit - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kStageBytes % 1024 == 0, "stage tiles must stay 1024 B aligned (128B swizzle atoms)");
};

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  template <int N> __device__ __forceinline__ void advance() {
    if (++stage == N) { stage = 0; phase ^= 1; }
  }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void unpack_bf16x4(uint2 t, float (&f)[4]) {
  const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
  const __nv_bfloat162 h1 = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
  f[0] = __low2float(h0); f[1] = __high2float(h0); f[2] = __low2float(h1); f[3] = __high2float(h1);
}

constexpr uint32_t kInvalidRow = 0xffffffffu;
constexpr int kRowBatch = 4;   // rows (x4 row groups) whose loads are in flight together in the epilogue's phase 2

// ------------------------------------------------------------------------------------------------
// Epilogue of one 128 x BN accumulator tile, executed by 16 warps: warp (q, g) owns TMEM lanes
// [32q, 32q+32) and the 32-column chunks c with c % 4 == g (argument `half`).  Per chunk:
//   phase 1  tcgen05.ld 32x32b.x32 (thread = row) -> 32x32 fp32 transpose buffer in shared memory
//            (16-byte XOR swizzle, conflict free both ways)
//   phase 2  lane = (row group rr = lane / 8, column group cg = lane % 8): each warp instruction covers
//            4 rows x 32 columns, so every global access (residual / aux loads, output store) is a set of
//            fully used 128-byte (fp32) or 64-byte (bf16) row segments, and the per-column parameters
//            (scale, bias, LayerScale) are 3 float4 registers per lane loaded once per chunk.
// Row bookkeeping: lane r of the warp computes, once per tile, the element offsets of ITS row in the output /
// residual / aux tensors (32-bit, the launcher checks the ranges); phase 2 fetches them with one shuffle each.
// Optional fused BatchNorm statistics: per-column sum / sum of squares of the pre-activation value,
// reduced over the warp's rows with shuffles, accumulated per CTA in shared memory (fp32) and flushed
// with fp64 atomics when the CTA moves to another column block.
template <int BN, int OUT, int ACT, int MAP, int OPT>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int q, int half, int lane,
                                              int m_blk, int n_blk, float* __restrict__ stg,
                                              float* __restrict__ colstats, uint64_t* tfull, uint32_t tfull_phase,
                                              long long* __restrict__ tr = nullptr, int next_m_blk = -1, int next_n_blk = -1) {
  const Epilogue& e = p.epi;
  const int map = (MAP == EM_RUNTIME) ? e.row_map : MAP;
  if constexpr ((OPT & OP_AUX_IN) != 0 && (OPT & OP_CONV) == 0) {
    // The gelu' operand (fc1 pre-activation saved by the forward, 50 MB at batch 64) is cold in HBM, and this kernel is
    // epilogue bound: the accumulator is already waiting when a warp gets here, so the ~2 us DRAM latency of the aux
    // loads below was fully exposed once per tile (55 us for the fc2 input gradient against 26 us for fc1 forward with
    // the same traffic).  Pull the NEXT tile's rows into L2 now, one 64-byte segment per lane and chunk.
    if (next_m_blk >= 0 && e.aux_in != nullptr) {
      const long long nrow = (long long)next_m_blk * kBlockM + q * 32 + lane;
      if (nrow < p.M) {
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(e.aux_in) + nrow * e.ld_aux + next_n_blk * BN;
        for (int c = half; c < BN / 32; c += kEpiGroups)
          if (next_n_blk * BN + c * 32 < e.n_valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + c * 32));
      }
    }
  }
  const bool conv = (OPT & OP_CONV) && p.a_mode == 1;
  const int r = q * 32 + lane;
  int logical;
  bool valid;
  if (conv) {
    const int xt = m_blk % p.tiles_x;
    const int t2 = m_blk / p.tiles_x;
    const int yt = t2 % p.tiles_y;
    const int bt = t2 / p.tiles_y;
    // bw, bh, bb are powers of two (choose_box): shifts instead of divisions
    const int c = r & (p.bw - 1);
    const int rr2 = (r >> p.bw_log2) & (p.bh - 1);
    const int bi = r >> (p.bw_log2 + p.bh_log2);
    const int x = xt * p.bw + c, y = yt * p.bh + rr2, b = bt * p.bb + bi;
    valid = (x < p.OW) && (y < p.OH) && (b < p.NB);
    logical = (b * p.OH + y) * p.OW + x;
  } else {
    logical = m_blk * kBlockM + r;
    valid = logical < p.M;
  }
  int out_row = logical, res_row = logical;
  int img = 0, pix = 0;
  if (map == EM_PATCH) {
    const int bimg = logical / e.map_a;
    const int n = logical - bimg * e.map_a;
    out_row = bimg * e.map_b + 1 + n;
    res_row = 1 + n;
  } else if (map == EM_NCHW || map == EM_SHUFFLE) {
    const int hw = p.OH * p.OW;
    img = logical / hw;
    pix = logical - img * hw;
    if (map == EM_SHUFFLE) {
      const int py = pix / p.OW, px = pix - py * p.OW;
      out_row = (img * (2 * p.OH) + 2 * py) * (2 * p.OW) + 2 * px;
    }
  }

  if (MAP == EM_NCHW || (MAP == EM_RUNTIME && map == EM_NCHW)) {
    // fp32 NCHW output (heat-maps): for a fixed channel the 32 rows of a warp are 32 consecutive pixels,
    // so the thread = row layout is already coalesced.
    float* o = reinterpret_cast<float*>(e.out);
    const long long hw = (long long)p.OH * p.OW;
    const int act = (ACT == EA_RUNTIME) ? e.act : ACT;
    mbar_wait(tfull, tfull_phase);
    tc_fence_after();
#pragma unroll 1
    for (int c = half; c < BN / 32; c += kEpiGroups) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c * 32), v);
      tmem_ld_wait();
      const int col0 = n_blk * BN + c * 32;
      if (valid && col0 < e.n_valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col < e.n_valid) {
            float f = __uint_as_float(v[j]);
            if (e.scale != nullptr) f *= __ldg(e.scale + col);
            if (e.bias != nullptr) f += __ldg(e.bias + col);
            if (act == ACT_RELU) f = fmaxf(f, 0.f);
            o[((long long)img * e.map_a + col) * hw + pix] = f;
          }
        }
      }
      __syncwarp();
    }
    return;
  }

  // element offsets of this lane's row (as row r of the tile) in the tensors touched by phase 2
  const uint32_t off_out = valid ? uint32_t(out_row) * uint32_t(e.ldo) : kInvalidRow;
  uint32_t off_res = 0, off_aux = 0;
  if constexpr ((OPT & (OP_LSRES | OP_RES_BF16)) != 0) off_res = uint32_t(res_row) * uint32_t(e.ldr);
  if constexpr ((OPT & (OP_AUX_OUT | OP_AUX_IN)) != 0) off_aux = uint32_t(logical) * uint32_t(e.ld_aux);

  const bool out_f32 = (OUT == EO_RUNTIME) ? (e.out_dtype == OUT_F32) : (OUT == EO_F32);
  const int act = (ACT == EA_RUNTIME) ? e.act : ACT;
  const bool has_scale = (OPT & OP_SCALE) && e.scale != nullptr;
  const bool has_ls = (OPT & OP_LSRES) && e.ls != nullptr;
  const bool has_res32 = (OPT & OP_LSRES) && e.residual != nullptr && !e.res_is_bf16 && !(e.debug & 2);
  const bool has_res16 = (OPT & OP_RES_BF16) && e.residual != nullptr && e.res_is_bf16 && !(e.debug & 2);
  const bool has_aux_out = (OPT & OP_AUX_OUT) && e.aux_out != nullptr;
  const bool has_aux_in = (OPT & OP_AUX_IN) && e.aux_in != nullptr && !(e.debug & 2);
  const bool has_stats = (OPT & OP_STATS) && e.stats != nullptr;
  (void)has_ls;

  constexpr uint32_t kFull = 0xffffffffu;
  // (Tried and removed: applying bias / activation BEFORE the transpose with thread = row, staging bf16 and copying
  //  16-byte pieces in phase 2 halves the staging traffic but needs 16 broadcast parameter loads per chunk and thread
  //  and serialises the GELU per row: qkv 20.4 -> 23.0 us, fc1 32.2 -> 45.0 us.)
  const int rr = lane >> 3, cg = lane & 7;
  // The accumulator-ready wait sits INSIDE the chunk loop, after every global load of the first chunk (per-column
  // parameters, residual / aux rows) has been issued: those loads do not depend on the accumulator, and an L2 round
  // trip under load (1-1.5 k cycles) is as long as the whole main loop of a K = 384 tile.
  bool waited = false;
#pragma unroll 1
  for (int c = half; c < BN / 32; c += kEpiGroups) {
    if (e.debug & 8) continue;
    const int col0 = n_blk * BN + c * 32;
    if (col0 >= e.n_valid) continue;  // warp-uniform
    const int ccol = col0 + cg * 4;
    const bool cvalid = ccol < e.n_valid;  // n_valid % 4 == 0 (checked by the launcher)
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bi = make_float4(0.f, 0.f, 0.f, 0.f), lsv = sc;
    if (cvalid) {
      if (has_scale) sc = ldg4(e.scale + ccol);
      if (e.bias != nullptr) bi = ldg4(e.bias + ccol);
      if (has_ls) lsv = ldg4(e.ls + ccol);
    }
    // row offsets (one shuffle each) and every global LOAD of the chunk, all 8 row groups back to back.  The residual
    // may alias the output (in-place residual stream), which stops the compiler from hoisting loads above the stores
    // of earlier rows by itself -- each element is read before it is written by the same thread, so this is safe.
    float4 res32[(OPT & OP_LSRES) ? 8 : 1];
    uint2 res16[(OPT & OP_RES_BF16) ? 8 : 1];
    uint2 auxin[(OPT & OP_AUX_IN) ? 8 : 1];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + rr;
      const bool ok = __shfl_sync(kFull, off_out, row) != kInvalidRow && cvalid;
      if constexpr ((OPT & (OP_LSRES | OP_RES_BF16)) != 0) {
        const uint32_t r_off = __shfl_sync(kFull, off_res, row);
        if constexpr ((OPT & OP_LSRES) != 0) {
          res32[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_res32 && ok) res32[it] = ldg4(e.residual + r_off + ccol);
        }
        if constexpr ((OPT & OP_RES_BF16) != 0) {
          res16[it] = make_uint2(0u, 0u);
          if (has_res16 && ok)
            res16[it] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(e.residual) + r_off + ccol));
        }
      }
      if constexpr ((OPT & (OP_AUX_OUT | OP_AUX_IN)) != 0) {
        const uint32_t ao = __shfl_sync(kFull, off_aux, row);
        if constexpr ((OPT & OP_AUX_IN) != 0) {
          auxin[it] = make_uint2(0u, 0u);
          if (has_aux_in && ok)
            auxin[it] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(e.aux_in) + ao + ccol));
        }
      }
    }
    if (!waited) {
      mbar_wait(tfull, tfull_phase);
      tc_fence_after();
      waited = true;
    }
    uint32_t v[32];
    if (tr) tr[0] = clock64();
    tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c * 32), v);
    tmem_ld_wait();
    if (tr) tr[1] = clock64();
    {
      float4* srow = reinterpret_cast<float4*>(stg + lane * 32);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        srow[jj ^ (lane & 7)] = make_float4(__uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]),
                                            __uint_as_float(v[4 * jj + 2]), __uint_as_float(v[4 * jj + 3]));
    }
    __syncwarp();
    if (tr) tr[2] = clock64();
    if (e.debug & 4) continue;
    uint32_t ocol = uint32_t(ccol), tap_off = 0;
    if (map == EM_SHUFFLE) {
      const int tap = col0 / e.map_a;
      ocol = uint32_t(ccol - tap * e.map_a);
      tap_off = uint32_t((tap >> 1) * (2 * p.OW) + (tap & 1)) * uint32_t(e.ldo);
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    // branch-free math for all 8 row groups (the compiler interleaves the rows), predicated stores only
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + rr;
      const float4 x = *reinterpret_cast<const float4*>(stg + row * 32 + ((cg ^ (row & 7)) << 2));
      const uint32_t o_off = __shfl_sync(kFull, off_out, row);   // re-fetched instead of kept live across the TMEM load
      const bool ok = o_off != kInvalidRow && cvalid;
      float f[4];
      if constexpr ((OPT & OP_SCALE) != 0) {
        f[0] = fmaf(x.x, sc.x, bi.x); f[1] = fmaf(x.y, sc.y, bi.y); f[2] = fmaf(x.z, sc.z, bi.z); f[3] = fmaf(x.w, sc.w, bi.w);
      } else {
        f[0] = x.x + bi.x; f[1] = x.y + bi.y; f[2] = x.z + bi.z; f[3] = x.w + bi.w;
      }
      if constexpr ((OPT & OP_STATS) != 0) {
        const float m = ok ? 1.f : 0.f;   // rows / columns outside the problem do not count
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s1[k] = fmaf(f[k], m, s1[k]);
          s2[k] = fmaf(f[k] * m, f[k], s2[k]);
        }
      }
      // aux_out: what the backward of this layer multiplies with.  With a GELU it is gelu'(v) (same tanh as the forward
      // value, six more FMAs), so the backward epilogue is a plain multiply; without an activation it is v itself.
      const bool gelu_here = (ACT == EA_GELU) || (ACT == EA_RUNTIME && act == ACT_GELU);
      float dg[4] = {f[0], f[1], f[2], f[3]};
      if (gelu_here) {
        if constexpr ((OPT & OP_AUX_OUT) != 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) f[k] = gelu_fast_both(f[k], dg[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) f[k] = gelu_fast(f[k]);
        }
      } else if ((ACT == EA_RELU) || (ACT == EA_RUNTIME && act == ACT_RELU)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = fmaxf(f[k], 0.f);
      }
      if constexpr ((OPT & OP_AUX_OUT) != 0) {
        const uint32_t a_off = __shfl_sync(kFull, off_aux, row);   // all lanes take part
        if (has_aux_out && ok) {
          uint2 t;
          t.x = pack_bf16x2(dg[0], dg[1]);
          t.y = pack_bf16x2(dg[2], dg[3]);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.aux_out) + a_off + ccol) = t;
        }
      }
      if constexpr ((OPT & OP_AUX_IN) != 0) {
        float t[4];
        unpack_bf16x4(auxin[it], t);
        if (has_aux_in) {
#pragma unroll
          for (int k = 0; k < 4; ++k) f[k] *= t[k];
        }
      }
      if constexpr ((OPT & OP_LSRES) != 0) {
        f[0] = fmaf(f[0], lsv.x, res32[it].x); f[1] = fmaf(f[1], lsv.y, res32[it].y);
        f[2] = fmaf(f[2], lsv.z, res32[it].z); f[3] = fmaf(f[3], lsv.w, res32[it].w);
      }
      if constexpr ((OPT & OP_RES_BF16) != 0) {
        float t[4];
        unpack_bf16x4(res16[it], t);
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] += t[k];
      }
      const uint32_t off = o_off + tap_off + ocol;
      if (ok) {
        if (out_f32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off) = make_float4(f[0], f[1], f[2], f[3]);
        } else {
          uint2 t;
          t.x = pack_bf16x2(f[0], f[1]);
          t.y = pack_bf16x2(f[2], f[3]);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + off) = t;
        }
      }
      if (tr && (it & 3) == 3) tr[3 + it / 4] = clock64();
    }
    if constexpr ((OPT & OP_STATS) != 0) {
      if (has_stats) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s1[k] += __shfl_xor_sync(kFull, s1[k], 8);
          s1[k] += __shfl_xor_sync(kFull, s1[k], 16);
          s2[k] += __shfl_xor_sync(kFull, s2[k], 8);
          s2[k] += __shfl_xor_sync(kFull, s2[k], 16);
        }
        if (rr == 0 && cvalid) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            atomicAdd(colstats + c * 32 + cg * 4 + k, s1[k]);
            atomicAdd(colstats + BN + c * 32 + cg * 4 + k, s2[k]);
          }
        }
      }
    }
    __syncwarp();  // staging buffer is rewritten by the next chunk
  }
  if (!waited) {   // every chunk of this warp skipped: still consume the phase before releasing the accumulator
    mbar_wait(tfull, tfull_phase);
    tc_fence_after();
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogue for plain bf16 outputs (QKV, fc1): bias / GELU applied with thread = row straight out of TMEM, the
// 32 x 32 bf16 chunk staged in shared memory in the 64B-swizzled layout of a TMA box and written with ONE
// cp.async.bulk.tensor store per chunk (rows / columns outside the tensor are clipped by the tensor map).
// Against epilogue_tile this removes the fp32 transpose round trip through shared memory, the per-row offset
// shuffles and the 8 predicated global stores per lane and chunk: ~70 instead of ~340 instructions per chunk
// and warp before the activation, and the store latency is off the warp's critical path (two staging buffers per
// warp, bulk_wait_read before a buffer is reused).
template <int BN>
struct TmaEpiBias {
  float v[(BN / 32 + kEpiGroups - 1) / kEpiGroups];   // lane l holds bias[col0 + l] of every chunk this warp handles
};
// Issued BEFORE the wait for the accumulator: the bias segment of a tile is new to the SM (L1 is invalidated per launch,
// n_blk changes every tile), i.e. an L2 round trip of 1-1.5 k cycles under load that would otherwise sit on the
// epilogue's critical path (in-kernel timeline, DP_GEMM_TRACE=1: 1.8-2.3 k cycles per chunk before).
template <int BN>
__device__ __forceinline__ void epilogue_tma_prefetch(const GemmParams& p, int half, int lane, int n_blk, TmaEpiBias<BN>& pre) {
  const Epilogue& e = p.epi;
  int i = 0;
#pragma unroll
  for (int c = half; c < BN / 32; c += kEpiGroups, ++i) {
    const int col = n_blk * BN + c * 32 + lane;
    pre.v[i] = (e.bias != nullptr && col < e.n_valid) ? __ldg(e.bias + col) : 0.f;
  }
}
template <int BN, int ACT>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, uint32_t tmem_acc, int q, int half, int lane,
                                                  int m_blk, int n_blk, uint8_t* __restrict__ stg, uint32_t& nstore,
                                                  const TmaEpiBias<BN>& pre, long long* __restrict__ tr = nullptr) {
  const Epilogue& e = p.epi;
  constexpr uint32_t kFull = 0xffffffffu;
  int i = 0;
#pragma unroll
  for (int c = half; c < BN / 32; c += kEpiGroups, ++i) {
    const int col0 = n_blk * BN + c * 32;
    if (col0 >= e.n_valid) continue;  // warp-uniform
    uint32_t v[32];
    if (tr && i == 0) tr[0] = clock64();
    tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c * 32), v);
    tmem_ld_wait();
    if (tr && i == 0) tr[1] = clock64();
    const float bl = pre.v[i];
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float f0 = __uint_as_float(v[2 * j]) + __shfl_sync(kFull, bl, 2 * j);
      float f1 = __uint_as_float(v[2 * j + 1]) + __shfl_sync(kFull, bl, 2 * j + 1);
      if constexpr (ACT == EA_GELU) {
        f0 = gelu_fast(f0); f1 = gelu_fast(f1);
      } else if constexpr (ACT == EA_RELU) {
        f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f);
      }
      pk[j] = pack_bf16x2(f0, f1);
    }
    uint8_t* buf = stg + (nstore & 1u) * 2048u;
    if (tr && i == 0) tr[2] = clock64();
    if (nstore >= 2) {   // the store issued two chunks ago has to be done reading this buffer
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    if (tr && i == 0) tr[3] = clock64();
    // row = lane, 64 bytes per row; 16-byte piece j lands at j ^ ((row >> 1) & 3) (CU_TENSOR_MAP_SWIZZLE_64B)
    uint8_t* rowp = buf + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&p.tmC, buf, col0, m_blk * kBlockM + q * 32);
      bulk_commit();
    }
    if (tr && i == 0) tr[4] = clock64();
    ++nstore;
  }
}

template <int BN, int OUT, int ACT, int MAP, int OPT>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_fwd_kernel(const __grid_constant__ GemmParams p) {
  using C = GCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_gemm[];
  uint8_t* smem = smem_gemm;
  float* staging = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  float* colstats = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + kStagingBytes + 2 * BN * 4);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzle atoms need a 1024 B aligned base
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if constexpr ((OPT & OP_TMA_OUT) != 0) tma_prefetch_desc(&p.tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, C::kTmemCols);
    tmem_relinquish();
  }
  if constexpr ((OPT & OP_STATS) != 0)
    for (int i = threadIdx.x; i < 2 * BN; i += kGemmThreads) colstats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // everything above touches only shared memory / TMEM / the parameter bank: it overlaps the tail of the previous
  // launch (programmatic dependent launch); global memory is read or written only below this line
  pdl_grid_sync();
  if (p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) {   // SM clock probe: cycles and ns of this CTA's lifetime
    p.epi.trace[4000] = clock64();
    p.epi.trace[4001] = global_timer_ns();
  }

  const int num_tiles = p.m_tiles * p.n_tiles;
  // tile order: n fastest (A tile shared by neighbouring CTAs through L2) unless column statistics are
  // accumulated per CTA, in which case m is fastest so that a CTA rarely changes its column block
  const bool m_fast = (OPT & OP_STATS) && p.epi.stats != nullptr;
#define DP_TILE_COORDS(tile, m_blk, n_blk)                            \
  const int m_blk = m_fast ? (tile) % p.m_tiles : (tile) / p.n_tiles; \
  const int n_blk = m_fast ? (tile) / p.m_tiles : (tile) % p.n_tiles;

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      const bool conv = (OPT & OP_CONV) && p.a_mode == 1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        DP_TILE_COORDS(tile, m_blk, n_blk)
        int x0 = 0, y0 = 0, b0 = 0;
        if (conv) {
          x0 = (m_blk % p.tiles_x) * p.bw;
          y0 = ((m_blk / p.tiles_x) % p.tiles_y) * p.bh;
          b0 = (m_blk / (p.tiles_x * p.tiles_y)) * p.bb;
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          uint8_t* sa = smem + ps.stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          // tuning knobs: 16 = A tile loaded only for the first k-blocks of the launch (A-stationary emulation),
          // 32 = same for the weight tile; results are garbage by design
          const bool skip_a = (p.epi.debug & 16) && (tile != (int)blockIdx.x);
          const bool skip_b = (p.epi.debug & 32) && (tile != (int)blockIdx.x);
          mbar_arrive_expect_tx(&full_bar[ps.stage], (skip_a ? 0 : kABytes) + (skip_b ? 0 : C::kBBytes));
          if (skip_a) {
          } else if (conv) {
            const int ky = tap / p.kw, kx = tap - ky * p.kw;
            tma_load_4d(sa, &p.tmA, &full_bar[ps.stage], cb * kBlockK, x0 + kx - p.pad_x, y0 + ky - p.pad_y, b0);
            if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          } else {
            tma_load_2d(sa, &p.tmA, &full_bar[ps.stage], kb * kBlockK, m_blk * kBlockM);
          }
          if (!skip_b) tma_load_2d(sb, &p.tmB, &full_bar[ps.stage], kb * kBlockK, n_blk * BN);
          ps.template advance<C::kStages>();
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      PipeState ps;
      int acc = 0;
      uint32_t acc_phase = 0;
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        if (p.epi.trace != nullptr && blockIdx.x == 0) p.epi.trace[(tile / gridDim.x) * 4 + 0] = clock64();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + ps.stage * C::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = make_sdesc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bdesc = make_sdesc_sw128(b_addr + k * 32, 0, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[ps.stage]);
          ps.template advance<C::kStages>();
        }
        umma_commit(&tfull_bar[acc]);
        if (p.epi.trace != nullptr && blockIdx.x == 0) p.epi.trace[(tile / gridDim.x) * 4 + 1] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // column group of this warp
    float* stg = staging + (warp - 2) * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t nstore = 0;   // OP_TMA_OUT: tile stores issued by this warp so far (selects the staging buffer)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      DP_TILE_COORDS(tile, m_blk, n_blk)
      if (p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) p.epi.trace[(tile / gridDim.x) * 4 + 2] = clock64();
      long long* tr = (p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) ? p.epi.trace + 2048 + (tile / gridDim.x) * 8 : nullptr;
      if constexpr ((OPT & OP_TMA_OUT) != 0) {
        TmaEpiBias<BN> pre;
        epilogue_tma_prefetch<BN>(p, half, lane, n_blk, pre);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        epilogue_tile_tma<BN, ACT>(p, tmem_base + uint32_t(acc * BN), q, half, lane, m_blk, n_blk,
                                   reinterpret_cast<uint8_t*>(stg), nstore, pre, tr);
      } else {
        int nm = -1, nn = -1;
        if constexpr ((OPT & OP_AUX_IN) != 0) {
          const int next = tile + gridDim.x;
          if (next < num_tiles) {
            DP_TILE_COORDS(next, nm2, nn2)
            nm = nm2;
            nn = nn2;
          }
        }
        epilogue_tile<BN, OUT, ACT, MAP, OPT>(p, tmem_base + uint32_t(acc * BN), q, half, lane, m_blk, n_blk, stg, colstats,
                                              &tfull_bar[acc], acc_phase, tr, nm, nn);
      }
      tc_fence_before();
      __syncwarp();
      if (tr) tr[5] = clock64();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) p.epi.trace[(tile / gridDim.x) * 4 + 3] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      if constexpr ((OPT & OP_STATS) != 0) {
        if (p.epi.stats != nullptr) {
          const int next = tile + gridDim.x;
          const int next_n = next < num_tiles ? next / p.m_tiles : -1;
          if (next_n != n_blk) {  // CTA-uniform: flush this column block's statistics
            const int et = threadIdx.x - 64;  // 0 .. 32 * kEpiWarps - 1
            named_bar_sync(1, 32 * kEpiWarps);
            for (int i = et; i < 2 * BN; i += 32 * kEpiWarps) {
              const int which = i / BN, cl = i - which * BN;
              const int col = n_blk * BN + cl;
              if (col < p.epi.n_valid) {
                const int ch = (p.epi.row_map == ROWMAP_SHUFFLE2X2) ? col % p.epi.map_a : col;
                atomicAdd(p.epi.stats + which * p.epi.stats_c + ch, double(colstats[i]));
              }
              colstats[i] = 0.f;
            }
            named_bar_sync(1, 32 * kEpiWarps);
          }
        }
      }
    }
    if constexpr ((OPT & OP_TMA_OUT) != 0) {
      if (lane == 0) bulk_wait_read<0>();   // shared memory must stay allocated until the last tile store has read it
    }
  }
#undef DP_TILE_COORDS
  tc_fence_before();
  __syncthreads();
  if (p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) {
    p.epi.trace[4002] = clock64();
    p.epi.trace[4003] = global_timer_ns();
  }
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BN, int OUT, int ACT, int MAP, int OPT>
cudaError_t launch_gemm_variant(const GemmParams& p, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_fwd_kernel<BN, OUT, ACT, MAP, OPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  launch_k<gemm_fwd_kernel<BN, OUT, ACT, MAP, OPT>>(grid, kGemmThreads, GCfg<BN>::kSmemBytes, s, p);
  return cudaGetLastError();
}

#define DP_GEMM_VARIANT(BN, OUT, ACT, MAP, OPT) \
  GemmVariant { BN, OUT, ACT, MAP, OPT, 0, &launch_gemm_variant<BN, OUT, ACT, MAP, OPT> }

}  // namespace dp
