// tcgen05 / TMEM / TMA GEMM family for sm_100a.
//
//  (the forward kernel lives in gemm_kernel.cuh; its compiled variants in gemm_fwd_*.cu; this file holds the
//   variant dispatcher and the weight-gradient kernel)
//  gemm_fwd_kernel<BN,...> : C[M,N] = A[M,K] * W[N,K]^T with a fused epilogue.  A is either a plain
//                            K-major matrix (2-D TMA) or an NHWC activation read as an IMPLICIT
//                            convolution: one 4-D TMA box {64 ch, bw, bh, bb} per (filter tap, 64-channel
//                            block), shifted by the tap offset, out-of-bounds pixels zero-filled by TMA
//                            (= the conv padding).  Persistent CTAs, one per SM:
//                              warp 0      TMA producer (one elected lane)
//                              warp 1      tcgen05.mma issuer (one elected lane)
//                              warps 2..9  epilogue: tcgen05.ld TMEM -> registers -> smem transpose -> fused
//                                          math -> coalesced global stores (+ fused BatchNorm statistics)
//                            smem ring of kStages {A 128x64, W BNx64} bf16 tiles (128B swizzle), double-
//                            buffered fp32 accumulators in TMEM (2 x BN columns) so the epilogue of tile i
//                            overlaps the MMAs of tile i+1.
//  gemm_wgrad_kernel<BN>   : dW[m,n] += sum_p A[p,m] * B[p(+tap),n]  -- both operands MN-major (pixels are
//                            the reduction dim), implicit-conv tap shift on B, split-K over pixel tiles,
//                            fp32 red.global.add scatter into the PyTorch weight-gradient layout.
#include "gemm_kernel.cuh"

namespace dp {

namespace {

constexpr int kThreads = 192;                     // weight-gradient kernel: TMA warp, MMA warp, 4 epilogue warps

// weight-gradient kernel configuration (manual 1024 B alignment slack)
template <int BN> struct Cfg {
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_wgrad[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_wgrad) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();   // prologue above overlaps the previous launch; no global-memory access before this line

  // work item = (split, tap, m_blk, n_blk)
  const int num_items = p.splits * p.taps * p.m_tiles * p.n_tiles;
  const int kb_per_split = (p.total_k_blocks + p.splits - 1) / p.splits;
  constexpr int kAtomBytes = 64 * 128;  // one 64-channel x 64-pixel MN-major atom block

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int w = item;
        const int n_blk = w % p.n_tiles; w /= p.n_tiles;
        const int m_blk = w % p.m_tiles; w /= p.m_tiles;
        const int tap = w % p.taps;
        const int split = w / p.taps;
        const int ky = tap / p.kw, kx = tap % p.kw;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.total_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          uint8_t* sa = smem + ps.stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[ps.stage], C::kStageBytes);
          if (p.mode == 1) {
            const int x0 = (kb % p.tiles_x) * p.bw;
            const int y0 = ((kb / p.tiles_x) % p.tiles_y) * p.bh;
            const int b0 = (kb / (p.tiles_x * p.tiles_y)) * p.bb;
#pragma unroll
            for (int a = 0; a < kBlockM / 64; ++a)
              tma_load_4d(sa + a * kAtomBytes, &p.tmA, &full_bar[ps.stage], m_blk * kBlockM + a * 64, x0, y0, b0);
#pragma unroll
            for (int a = 0; a < BN / 64; ++a)
              tma_load_4d(sb + a * kAtomBytes, &p.tmB, &full_bar[ps.stage], n_blk * BN + a * 64, x0 + kx - p.pad_x,
                          y0 + ky - p.pad_y, b0);
          } else {
#pragma unroll
            for (int a = 0; a < kBlockM / 64; ++a)
              tma_load_2d(sa + a * kAtomBytes, &p.tmA, &full_bar[ps.stage], m_blk * kBlockM + a * 64, kb * 64);
#pragma unroll
            for (int a = 0; a < BN / 64; ++a)
              tma_load_2d(sb + a * kAtomBytes, &p.tmB, &full_bar[ps.stage], n_blk * BN + a * 64, kb * 64);
          }
          ps.template advance<C::kStages>();
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      PipeState ps;
      int acc = 0;
      uint32_t acc_phase = 0;
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 1, 1);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item / (p.taps * p.m_tiles * p.n_tiles);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.total_k_blocks);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + ps.stage * C::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < 64 / 16; ++k) {
            // 16 k-rows (pixels) x 128 B each per MMA; MN atoms kAtomBytes apart (LBO); 8-row groups 1024 B (SBO)
            const uint64_t adesc = make_sdesc_sw128(a_addr + k * 2048, kAtomBytes, 1024);
            const uint64_t bdesc = make_sdesc_sw128(b_addr + k * 2048, kAtomBytes, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[ps.stage]);
          ps.template advance<C::kStages>();
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int w = item;
      const int n_blk = w % p.n_tiles; w /= p.n_tiles;
      const int m_blk = w % p.m_tiles; w /= p.m_tiles;
      const int tap = w % p.taps;
      const int split = w / p.taps;
      const bool has_k = split * kb_per_split < p.total_k_blocks;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * kBlockM + q * 32 + lane;
      if (p.ws != nullptr) {
        // atomic-free path: this item's 128 x BN partial tile goes to its own slice of the workspace
        // (rows / columns beyond Mc / Nc are stored too: they are zero and never read)
        float* wrow = p.ws + ((long long)(split * p.taps + tap) * (p.m_tiles * kBlockM) + m) * p.ws_ld + n_blk * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + uint32_t(acc * BN) + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
          tmem_ld_wait();
          if (!(p.debug & 1)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(wrow + c0)[j] =
                  has_k ? make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                      __uint_as_float(v[4 * j + 3]))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      const bool mvalid = (m < p.Mc) && has_k;
      float* orow = p.out + (long long)(m % p.m_inner) * p.so_m + (long long)(m / p.m_inner) * p.so_mo +
                    (long long)tap * p.so_t;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + uint32_t(acc * BN) + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
        tmem_ld_wait();
        if (!mvalid || (p.debug & 1)) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n_blk * BN + c0 + j;
          if (n < p.Nc)
            atomicAdd(orow + (long long)(n % p.n_inner) * p.so_n + (long long)(n / p.n_inner) * p.so_no,
                      __uint_as_float(v[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

// out[off(m) + off(n) + tap * so_t] = sum over splits of the workspace partials (overwrites: no pre-zeroing needed).
// One thread per (tap, m, n) -- reads coalesced along n -- with the split loop unrolled over four independent
// accumulators.  (First version: one thread per (m, n) walking taps x splits serially; with 16 k threads for a
// 128 x 128 x 16-tap gradient it ran at 0.5 TB/s, ~15 us per launch, 14 launches per step -- profiles/r1h_step_metrics.md.)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, int splits,
                                                           int taps, int Mpad, int ld, int Mc, int Nc, long long so_m,
                                                           long long so_mo, long long so_n, long long so_no, long long so_t,
                                                           int m_inner, int n_inner) {
  pdl_grid_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_tap = (long long)Mc * Nc;
  if (idx >= per_tap * taps) return;
  const int t = int(idx / per_tap);
  const long long rem = idx - (long long)t * per_tap;
  const int m = int(rem / Nc), n = int(rem - (long long)m * Nc);
  const long long base = (long long)(m % m_inner) * so_m + (long long)(m / m_inner) * so_mo +
                         (long long)(n % n_inner) * so_n + (long long)(n / n_inner) * so_no;
  const long long slice = (long long)Mpad * ld;
  const long long step = (long long)taps * slice;
  const float* p = ws + (long long)t * slice + (long long)m * ld + n;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    a0 += __ldg(p + (long long)s * step);
    a1 += __ldg(p + (long long)(s + 1) * step);
    a2 += __ldg(p + (long long)(s + 2) * step);
    a3 += __ldg(p + (long long)(s + 3) * step);
  }
  for (; s < splits; ++s) a0 += __ldg(p + (long long)s * step);
  out[base + (long long)t * so_t] = (a0 + a1) + (a2 + a3);
}

template <int BN> cudaError_t launch_wgrad_t(const WgradParams& p, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  launch_k<gemm_wgrad_kernel<BN>>(grid, kThreads, Cfg<BN>::kSmemBytes, s, p);
  return cudaGetLastError();
}

}  // namespace

extern const GemmVariant kGemmVariantsA[], kGemmVariantsB[], kGemmVariantsC[], kGemmVariantsD[], kGemmVariantsGeneric[],
    kGemmVariantsPair[], kGemmVariantsAstat[];
extern const int kNumGemmVariantsA, kNumGemmVariantsB, kNumGemmVariantsC, kNumGemmVariantsD, kNumGemmVariantsGeneric,
    kNumGemmVariantsPair, kNumGemmVariantsAstat;

// Pick the cheapest compiled variant (single-CTA or CTA-pair, as requested) whose compile-time feature set covers
// what this launch needs; nullptr if none is compiled for this tile width.
const GemmVariant* select_gemm_variant(const Epilogue& e, int a_mode, int block_n, int pair, bool tma_out_ok) {
  int need = 0;
  if (e.scale) need |= OP_SCALE;
  if (e.ls || (e.residual && !e.res_is_bf16)) need |= OP_LSRES;
  if (e.residual && e.res_is_bf16) need |= OP_RES_BF16;
  if (e.aux_out) need |= OP_AUX_OUT;
  if (e.aux_in) need |= OP_AUX_IN;
  if (e.stats) need |= OP_STATS;
  if (a_mode == 1) need |= OP_CONV;
  const GemmVariant* tables[7] = {kGemmVariantsA, kGemmVariantsB, kGemmVariantsC, kGemmVariantsD, kGemmVariantsGeneric,
                                  kGemmVariantsPair, kGemmVariantsAstat};
  const int counts[7] = {kNumGemmVariantsA, kNumGemmVariantsB, kNumGemmVariantsC, kNumGemmVariantsD, kNumGemmVariantsGeneric,
                         kNumGemmVariantsPair, kNumGemmVariantsAstat};
  const GemmVariant* best = nullptr;
  int best_cost = 1 << 30;
  for (int t = 0; t < 7; ++t)
    for (int i = 0; i < counts[t]; ++i) {
      const GemmVariant& v = tables[t][i];
      if (v.bn != block_n || v.pair != pair) continue;
      if (v.out != EO_RUNTIME && v.out != e.out_dtype) continue;
      if (v.act != EA_RUNTIME && v.act != e.act) continue;
      if (v.map != EM_RUNTIME && v.map != e.row_map) continue;
      if ((v.opt & need) != need) continue;
      if ((v.opt & OP_TMA_OUT) && !tma_out_ok) continue;
      const int cost = __builtin_popcount(v.opt & OP_ALL) + 8 * ((v.out == EO_RUNTIME) + (v.act == EA_RUNTIME) + (v.map == EM_RUNTIME)) -
                       ((v.opt & OP_TMA_OUT) ? 64 : 0);   // a qualifying output prefers the TMA-store epilogue
      if (cost < best_cost) { best_cost = cost; best = &v; }
    }
  return best;
}

cudaError_t launch_wgrad(const WgradParams& p, int block_n, int grid, cudaStream_t s) {
  cudaError_t e;
  switch (block_n) {
    case 64: e = launch_wgrad_t<64>(p, grid, s); break;
    case 128: e = launch_wgrad_t<128>(p, grid, s); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess || p.ws == nullptr) return e;
  const long long total = (long long)p.Mc * p.Nc * p.taps;
  launch_k<wgrad_reduce_kernel>(unsigned((total + 255) / 256), 256, 0, s, p.ws, p.out, p.splits, p.taps, p.m_tiles * kBlockM, p.ws_ld,
                                                                    p.Mc, p.Nc, p.so_m, p.so_mo, p.so_n, p.so_no, p.so_t,
                                                                    p.m_inner, p.n_inner);
  return cudaGetLastError();
}

}  // namespace dp
