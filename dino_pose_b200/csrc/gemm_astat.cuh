// A-stationary tcgen05 GEMM for the short-K linear layers of the backbone (K <= 512: QKV, fc1 and the MLP input
// gradients of ViT-S, K = D = 384):   C[M,N] = epilogue(A[M,K] * W[N,K]^T),  128 x 128 tiles.
//
// Why: the in-kernel timeline of gemm_fwd_kernel (profiles/r2a_gemm_timeline.md) shows a tile period of ~3.7 k cycles
// against an MMA floor of 1.5 k at K = 384, and the period tracks the SHARED-MEMORY byte count of a tile, not its L2 or
// TMEM traffic: a 128 x 128 x 16 SS-mode MMA reads 8 KB of operands per 64 cycles (= the 128 B/clk the SM has), and the
// TMA writes of the same 192 KB per tile plus the epilogue staging ride on the same port.  This kernel removes the two
// avoidable streams:
//   * every CTA works on a CONTIGUOUS range of the m-major tile list, so its consecutive tiles share the 128 x K block
//     of A: A is fetched once per row block (at most twice per CTA) instead of once per tile;
//   * that block lives in TENSOR MEMORY (tcgen05.cp, shared -> TMEM, 32 columns per 64-wide k-block) and the MMAs
//     run in the TS form (A from TMEM, B from shared memory): operand reads from shared memory halve.
// Per 128 x 128 tile at K = 384: 96 KB weight reads + 96 KB weight TMA writes + ~24 KB for A + the epilogue staging,
// against 192 + 192 + staging before.
//
// Pipeline: warp 0 TMA producer, warp 1 tcgen05 issuer, 16 epilogue warps (shared with gemm_kernel.cuh).  ONE ring
// of 16 KB slots carries, in consumption order, the A k-blocks of a row block (the issuer turns each into four
// tcgen05.cp) and the weight k-blocks of its tiles (four MMAs each); tcgen05.cp and tcgen05.mma execute in issue order,
// which is what makes overwriting the A columns safe while the previous row block's last MMAs are still in flight.
// TMEM: columns [0,128) and [128,256) double-buffered accumulators, [256, 256 + 32 * k-blocks) the A block.
#pragma once
#include "gemm_kernel.cuh"

namespace dp {

constexpr int kAsBN = 128;
constexpr int kAsSlotBytes = kABytes;                                             // 16 KB: 128 rows x 64 bf16
constexpr int kAsSlots = (kSmemLimit - kStagingBytes - 256) / kAsSlotBytes;       // 10
constexpr int kAsSmemBytes = kAsSlots * kAsSlotBytes + kStagingBytes + 256;
constexpr int kAsTmemA = 256;                                                     // first TMEM column of the A block
constexpr int kAsMaxKBlocks = (512 - kAsTmemA) / 32;                              // 8 k-blocks: K <= 512
static_assert(kAsSlots >= 8, "ring too shallow");

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared -> TMEM: 128 lanes x 256 bits (one K = 16 slice of a K-major bf16 operand), source given by a matrix descriptor
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t dst_tmem, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(dst_tmem), "l"(sdesc) : "memory");
}

// TMA tile load delivered to the same shared-memory offset (and signalled on the mbarrier at the same offset) in every CTA
// of the cluster named by `mask`
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// arrive (once all previously issued tcgen05 operations of this thread completed) on the mbarrier at this offset in every
// CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// CL = 1: independent CTAs.  CL = 2: clusters of two CTAs that work on two row blocks and the SAME column-tile sequence in
// lock step; each CTA fetches half of every weight k-block and multicasts it to both, so the weight stream crosses the
// L2 -> SM fabric once per 256 rows.  (The MMAs stay cta_group::1: each CTA multiplies its own row block.)
template <int CL, int OUT, int ACT, int OPT>
__device__ __forceinline__ void gemm_astat_body(const GemmParams& p) {
  constexpr int BN = kAsBN;
  extern __shared__ __align__(1024) uint8_t smem_gemm[];
  uint8_t* smem = smem_gemm;
  float* staging = reinterpret_cast<float*>(smem + kAsSlots * kAsSlotBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kAsSlots * kAsSlotBytes + kStagingBytes);
  uint64_t* empty_bar = full_bar + kAsSlots;
  uint64_t* tfull_bar = empty_bar + kAsSlots;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if constexpr ((OPT & OP_TMA_OUT) != 0) tma_prefetch_desc(&p.tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kAsSlots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CL);   // a slot is free once the MMA threads of ALL CTAs that receive it have released it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // the peer's barriers are initialised before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();
  if (p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) {
    p.epi.trace[4000] = clock64();
    p.epi.trace[4001] = global_timer_ns();
  }

  // contiguous range of the m-major tile list (n fastest); with CL = 2 the list is over PAIRS of row blocks and CTA r of
  // the cluster takes row block 2 * pair + r (an odd last row block leaves a phantom half: zero-filled loads, clipped stores)
  const int rank = CL > 1 ? int(cluster_ctarank()) : 0;
  const int unit = CL > 1 ? int(blockIdx.x) / CL : int(blockIdx.x);
  const int units = int(gridDim.x) / CL;
  const long long num_tiles = (long long)((p.m_tiles + CL - 1) / CL) * p.n_tiles;
  const int t_begin = int(num_tiles * unit / units);
  const int t_end = int(num_tiles * (unit + 1) / units);
  const int nkb = p.num_k_blocks;
  constexpr uint16_t kMask = uint16_t((1u << CL) - 1u);
#define DP_AS_COORDS(tile, m_blk, n_blk)                     \
  const int m_blk = ((tile) / p.n_tiles) * CL + rank;        \
  const int n_blk = (tile) - ((tile) / p.n_tiles) * p.n_tiles;

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      int cur_m = -1;
      for (int tile = t_begin; tile < t_end; ++tile) {
        DP_AS_COORDS(tile, m_blk, n_blk)
        if (m_blk != cur_m) {
          cur_m = m_blk;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[ps.stage], kABytes);
            tma_load_2d(smem + ps.stage * kAsSlotBytes, &p.tmA, &full_bar[ps.stage], kb * kBlockK, m_blk * kBlockM);
            ps.template advance<kAsSlots>();
          }
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[ps.stage], BN * kBlockK * 2);
          if constexpr (CL > 1) {
            // this CTA's share of the weight k-block (BN / CL rows), delivered to every CTA of the cluster
            constexpr int kRows = BN / CL;
            tma_load_2d_mc(smem + ps.stage * kAsSlotBytes + rank * (kRows * kBlockK * 2), &p.tmB, &full_bar[ps.stage],
                           kb * kBlockK, n_blk * BN + rank * kRows, kMask);
          } else {
            tma_load_2d(smem + ps.stage * kAsSlotBytes, &p.tmB, &full_bar[ps.stage], kb * kBlockK, n_blk * BN);
          }
          ps.template advance<kAsSlots>();
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      PipeState ps;
      int acc = 0, cur_m = -1;
      uint32_t acc_phase = 0;
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, 0);
      const uint32_t a_tmem = tmem_base + uint32_t(kAsTmemA);
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int m_blk = (tile / p.n_tiles) * CL + rank;
        if (m_blk != cur_m) {
          cur_m = m_blk;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[ps.stage], ps.phase);
            tc_fence_after();
            const uint32_t s_addr = smem_u32(smem + ps.stage * kAsSlotBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              tmem_cp_128x256b(a_tmem + uint32_t(kb * 32 + k * 8), make_sdesc_sw128(s_addr + k * 32, 0, 1024));
            if constexpr (CL > 1) umma_commit_mc(&empty_bar[ps.stage], kMask);
            else umma_commit(&empty_bar[ps.stage]);
            ps.template advance<kAsSlots>();
          }
        }
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        if (p.epi.trace != nullptr && blockIdx.x == 0) p.epi.trace[(tile - t_begin) * 4 + 0] = clock64();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smem + ps.stage * kAsSlotBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_ts(d_tmem, a_tmem + uint32_t(kb * 32 + k * 8), make_sdesc_sw128(b_addr + k * 32, 0, 1024), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          if constexpr (CL > 1) umma_commit_mc(&empty_bar[ps.stage], kMask);
          else umma_commit(&empty_bar[ps.stage]);
          ps.template advance<kAsSlots>();
        }
        umma_commit(&tfull_bar[acc]);
        if (p.epi.trace != nullptr && blockIdx.x == 0) p.epi.trace[(tile - t_begin) * 4 + 1] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = staging + (warp - 2) * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t nstore = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      DP_AS_COORDS(tile, m_blk, n_blk)
      const bool tracer = p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
      long long* tr = tracer ? p.epi.trace + 2048 + (tile - t_begin) * 8 : nullptr;
      if (tracer) p.epi.trace[(tile - t_begin) * 4 + 2] = clock64();
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
          if (tile + 1 < t_end) {
            DP_AS_COORDS(tile + 1, nm2, nn2)
            nm = nm2;
            nn = nn2;
          }
        }
        epilogue_tile<BN, OUT, ACT, EM_IDENTITY, OPT>(p, tmem_base + uint32_t(acc * BN), q, half, lane, m_blk, n_blk, stg,
                                                      nullptr, &tfull_bar[acc], acc_phase, tr, nm, nn);
      }
      tc_fence_before();
      __syncwarp();
      if (tr) tr[5] = clock64();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (tracer) p.epi.trace[(tile - t_begin) * 4 + 3] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr ((OPT & OP_TMA_OUT) != 0) {
      if (lane == 0) bulk_wait_read<0>();
    }
  }
#undef DP_AS_COORDS
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // nobody leaves while the peer may still arrive on this CTA's barriers
  if (p.epi.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) {
    p.epi.trace[4002] = clock64();
    p.epi.trace[4003] = global_timer_ns();
  }
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int OUT, int ACT, int OPT>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_astat_kernel(const __grid_constant__ GemmParams p) {
  gemm_astat_body<1, OUT, ACT, OPT>(p);
}
template <int OUT, int ACT, int OPT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
    gemm_astat_cl2_kernel(const __grid_constant__ GemmParams p) {
  gemm_astat_body<2, OUT, ACT, OPT>(p);
}

template <int OUT, int ACT, int OPT>
cudaError_t launch_gemm_astat(const GemmParams& p, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_astat_kernel<OUT, ACT, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kAsSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  launch_k<gemm_astat_kernel<OUT, ACT, OPT>>(grid, kGemmThreads, kAsSmemBytes, s, p);
  return cudaGetLastError();
}

template <int OUT, int ACT, int OPT>
cudaError_t launch_gemm_astat_cl2(const GemmParams& p, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_astat_cl2_kernel<OUT, ACT, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kAsSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  launch_k<gemm_astat_cl2_kernel<OUT, ACT, OPT>>(grid, kGemmThreads, kAsSmemBytes, s, p);
  return cudaGetLastError();
}

// pair field 2 = A-stationary TS-mode kernel, 3 = the same in clusters of two with multicast weight loads
#define DP_GEMM_ASTAT_VARIANT(OUT, ACT, OPT) \
  GemmVariant { kAsBN, OUT, ACT, EM_IDENTITY, OPT, 2, &launch_gemm_astat<OUT, ACT, OPT> }
#define DP_GEMM_ASTAT_CL2_VARIANT(OUT, ACT, OPT) \
  GemmVariant { kAsBN, OUT, ACT, EM_IDENTITY, OPT, 3, &launch_gemm_astat_cl2<OUT, ACT, OPT> }

}  // namespace dp
