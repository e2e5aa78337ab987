// CTA-pair version of the forward GEMM (tcgen05 cta_group::2): two CTAs of a cluster, on the two SMs of a TPC,
// compute one 256 x BN output tile.  Each CTA stages ITS 128 rows of A and HALF of the BN weight rows; the MMA
// (issued by the leader CTA only) reads A from both shared memories and each half of B once, so per CTA and
// k-block the operand traffic is 16 KB + BN/2 * 128 B for 128 x BN x 64 MACs:
//      L2 -> SM bytes per FLOP and shared-memory operand reads per MMA cycle are ~half of the single-CTA kernel
//      (the two walls measured in DESIGN.md section 3.1).
// Pipeline per CTA is the single-CTA kernel's (gemm_kernel.cuh): TMA producer warp, MMA issuer warp (leader only),
// 8 epilogue warps, smem ring, double-buffered TMEM accumulators; the epilogue code is shared.
// Barriers:  full[s]   leader only; expect_tx = both CTAs' bytes; both producers' TMA loads signal it (.cta_group::2)
//            empty[s]  in each CTA; tcgen05.commit multicast from the leader's MMA thread
//            tfull[a]  in each CTA; tcgen05.commit multicast
//            tempty[a] leader only; 2 x 8 epilogue warps arrive (the peer's through shared::cluster)
#pragma once
#include "gemm_kernel.cuh"

namespace dp {

template <int BN> struct PCfg {
  static constexpr int kBBytes = (BN / 2) * kBlockK * 2;    // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kFixedBytes = kStagingBytes + 2 * BN * 4 + 256;
  static constexpr int kFit = (kSmemLimit - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
  static_assert(BN % 32 == 0 && BN <= 256, "pair tile width");
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kStageBytes % 1024 == 0, "stage tiles must stay 1024 B aligned (128B swizzle atoms)");
};

template <int BN, int OUT, int ACT, int MAP, int OPT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
    gemm_fwd_pair_kernel(const __grid_constant__ GemmParams p) {
  using C = PCfg<BN>;
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
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
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
      mbar_init(&tempty_bar[i], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_holder, C::kTmemCols);
    tmem_relinquish_pair();
  }
  if constexpr ((OPT & OP_STATS) != 0)
    for (int i = threadIdx.x; i < 2 * BN; i += kGemmThreads) colstats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers initialised, peer TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();   // see gemm_kernel.cuh: no global-memory access above this line

  const int pm_tiles = (p.m_tiles + 1) >> 1;           // 256-row pair tiles
  const int num_tiles = pm_tiles * p.n_tiles;
  const bool m_fast = (OPT & OP_STATS) && p.epi.stats != nullptr;
#define DP_PAIR_COORDS(tile, m_blk, n_blk)                                                        \
  const int m_blk = 2 * (m_fast ? (tile) % pm_tiles : (tile) / p.n_tiles) + int(cta_rank);        \
  const int n_blk = m_fast ? (tile) / pm_tiles : (tile) % p.n_tiles;

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      const bool conv = (OPT & OP_CONV) && p.a_mode == 1;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        DP_PAIR_COORDS(tile, m_blk, n_blk)
        int x0 = 0, y0 = 0, b0 = 0;
        if (conv) {
          x0 = (m_blk % p.tiles_x) * p.bw;
          y0 = ((m_blk / p.tiles_x) % p.tiles_y) * p.bh;
          b0 = (m_blk / (p.tiles_x * p.tiles_y)) * p.bb;   // >= NB for the phantom half of an odd last pair: zero fill
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          uint8_t* sa = smem + ps.stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[ps.stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[ps.stage], 2 * C::kStageBytes);
          if (conv) {
            const int ky = tap / p.kw, kx = tap - ky * p.kw;
            tma_load_4d_pair(sa, &p.tmA, lead_full, cb * kBlockK, x0 + kx - p.pad_x, y0 + ky - p.pad_y, b0);
            if (++cb == p.cin_blocks) { cb = 0; ++tap; }
          } else {
            tma_load_2d_pair(sa, &p.tmA, lead_full, kb * kBlockK, m_blk * kBlockM);
          }
          tma_load_2d_pair(sb, &p.tmB, lead_full, kb * kBlockK, n_blk * BN + int(cta_rank) * (BN / 2));
          ps.template advance<C::kStages>();
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      PipeState ps;
      int acc = 0;
      uint32_t acc_phase = 0;
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BN, 0, 0);
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
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
            umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[ps.stage]);
          ps.template advance<C::kStages>();
        }
        umma_commit_pair(&tfull_bar[acc]);
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
    uint32_t nstore = 0;   // OP_TMA_OUT: tile stores issued by this warp so far
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      DP_PAIR_COORDS(tile, m_blk, n_blk)
      if constexpr ((OPT & OP_TMA_OUT) != 0) {
        // plain bf16 outputs (QKV, fc1): thread = row out of TMEM, staged bf16 tiles, TMA tile stores (rows of the phantom
        // half of an odd last pair lie beyond M and are clipped by the tensor map)
        TmaEpiBias<BN> pre;
        epilogue_tma_prefetch<BN>(p, half, lane, n_blk, pre);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        epilogue_tile_tma<BN, ACT>(p, tmem_base + uint32_t(acc * BN), q, half, lane, m_blk, n_blk,
                                   reinterpret_cast<uint8_t*>(stg), nstore, pre);
      } else {
        epilogue_tile<BN, OUT, ACT, MAP, OPT>(p, tmem_base + uint32_t(acc * BN), q, half, lane, m_blk, n_blk, stg, colstats,
                                              &tfull_bar[acc], acc_phase);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      if constexpr ((OPT & OP_STATS) != 0) {
        if (p.epi.stats != nullptr) {
          const int next = tile + num_clusters;
          const int next_n = next < num_tiles ? next / pm_tiles : -1;
          if (next_n != n_blk) {
            const int et = threadIdx.x - 64;
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
#undef DP_PAIR_COORDS
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves (or frees tensor memory) while the peer may still touch this CTA
  if (warp == 2) tmem_dealloc_pair(tmem_base, C::kTmemCols);
}

template <int BN, int OUT, int ACT, int MAP, int OPT>
cudaError_t launch_gemm_pair_variant(const GemmParams& p, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_fwd_pair_kernel<BN, OUT, ACT, MAP, OPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  launch_k<gemm_fwd_pair_kernel<BN, OUT, ACT, MAP, OPT>>(grid, kGemmThreads, PCfg<BN>::kSmemBytes, s, p);
  return cudaGetLastError();
}

#define DP_GEMM_PAIR_VARIANT(BN, OUT, ACT, MAP, OPT) \
  GemmVariant { BN, OUT, ACT, MAP, OPT, 1, &launch_gemm_pair_variant<BN, OUT, ACT, MAP, OPT> }

}  // namespace dp
