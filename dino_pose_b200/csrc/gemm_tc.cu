// tcgen05 / TMEM / TMA GEMM family for sm_100a.
//
//  gemm_kmajor_kernel<BN>  : C[M,N] = A[M,K] * W[N,K]^T with a fused epilogue.  A is either a plain
//                            K-major matrix (2-D TMA) or an NHWC activation read as an IMPLICIT
//                            convolution: one 4-D TMA box {64 ch, bw, bh, bb} per (filter tap, 64-channel
//                            block), shifted by the tap offset, out-of-bounds pixels zero-filled by TMA
//                            (= the conv padding).  Persistent CTAs, one per SM:
//                              warp 0      TMA producer (one elected lane)
//                              warp 1      tcgen05.mma issuer (one elected lane)
//                              warps 2..5  epilogue: tcgen05.ld TMEM -> registers -> fused math -> global
//                            smem ring of kStages {A 128x64, W BNx64} bf16 tiles (128B swizzle), double-
//                            buffered fp32 accumulators in TMEM (2 x BN columns) so the epilogue of tile i
//                            overlaps the MMAs of tile i+1.
//  gemm_wgrad_kernel<BN>   : dW[m,n] += sum_p A[p,m] * B[p(+tap),n]  -- both operands MN-major (pixels are
//                            the reduction dim), implicit-conv tap shift on B, split-K over pixel tiles,
//                            fp32 red.global.add scatter into the PyTorch weight-gradient layout.
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace dp {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 bf16 = 128 B = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;

template <int BN> struct Cfg {
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  template <int N> __device__ __forceinline__ void advance() {
    if (++stage == N) { stage = 0; phase ^= 1; }
  }
};

__device__ __forceinline__ void load32f(const float* __restrict__ p, int col0, bool full, int n_valid, float (&d)[32]) {
  if (full) {
    const float4* p4 = reinterpret_cast<const float4*>(p + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = __ldg(p4 + i);
      d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = __ldg(p + min(col0 + i, n_valid - 1));
  }
}

__device__ __forceinline__ void load32bf(const __nv_bfloat16* __restrict__ p, bool full, int nrem, float (&d)[32]) {
  if (full) {
    const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 t = __ldg(p4 + i);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        d[8 * i + 2 * k] = __low2float(h);
        d[8 * i + 2 * k + 1] = __high2float(h);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = (i < nrem) ? __bfloat162float(p[i]) : 0.f;
  }
}

__device__ __forceinline__ void store32bf(__nv_bfloat16* __restrict__ p, bool full, int nrem, const float (&f)[32]) {
  if (full) {
    uint4* p4 = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 t;
      t.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
      t.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
      t.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
      t.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
      p4[i] = t;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nrem) p[i] = __float2bfloat16_rn(f[i]);
  }
}

// ------------------------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int q, int lane, int m_blk,
                                              int n_blk) {
  const Epilogue& e = p.epi;
  const int r = q * 32 + lane;
  long long logical;
  bool valid;
  if (p.a_mode == 1) {
    const int xt = m_blk % p.tiles_x;
    const int yt = (m_blk / p.tiles_x) % p.tiles_y;
    const int bt = m_blk / (p.tiles_x * p.tiles_y);
    const int c = r % p.bw;
    const int rr = (r / p.bw) % p.bh;
    const int bi = r / (p.bw * p.bh);
    const int x = xt * p.bw + c, y = yt * p.bh + rr, b = bt * p.bb + bi;
    valid = (x < p.OW) && (y < p.OH) && (b < p.NB);
    logical = ((long long)b * p.OH + y) * p.OW + x;
  } else {
    logical = (long long)m_blk * kBlockM + r;
    valid = logical < p.M;
  }
  long long out_row = logical, res_row = logical;
  long long img = 0, pix = 0;
  if (e.row_map == ROWMAP_PATCH_TOKENS) {
    const long long bimg = logical / e.map_a;
    const long long n = logical % e.map_a;
    out_row = bimg * e.map_b + 1 + n;
    res_row = 1 + n;
  } else if (e.row_map == ROWMAP_NCHW || e.row_map == ROWMAP_SHUFFLE2X2) {
    const long long hw = (long long)p.OH * p.OW;
    img = logical / hw;
    pix = logical % hw;
  }

#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
    tmem_ld_wait();
    const int col0 = n_blk * BN + c0;
    if (!valid || col0 >= e.n_valid) continue;
    const bool full = (col0 + 32 <= e.n_valid);
    const int nrem = e.n_valid - col0;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    float t[32];
    if (e.scale != nullptr) {
      load32f(e.scale, col0, full, e.n_valid, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= t[j];
    }
    if (e.bias != nullptr) {
      load32f(e.bias, col0, full, e.n_valid, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += t[j];
    }
    if (e.aux_out != nullptr)
      store32bf(reinterpret_cast<__nv_bfloat16*>(e.aux_out) + logical * e.ld_aux + col0, full, nrem, f);
    if (e.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    } else if (e.act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
    }
    if (e.aux_in != nullptr) {
      load32bf(reinterpret_cast<const __nv_bfloat16*>(e.aux_in) + logical * e.ld_aux + col0, full, nrem, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= gelu_erf_grad(t[j]);
    }
    if (e.ls != nullptr) {
      load32f(e.ls, col0, full, e.n_valid, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= t[j];
    }
    if (e.residual != nullptr) {
      if (e.res_is_bf16)
        load32bf(reinterpret_cast<const __nv_bfloat16*>(e.residual) + res_row * e.ldr + col0, full, nrem, t);
      else
        load32f(e.residual + res_row * e.ldr, col0, full, e.n_valid, t);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += t[j];
    }
    // ---- store
    if (e.row_map == ROWMAP_NCHW) {
      float* o = reinterpret_cast<float*>(e.out);
      const long long hw = (long long)p.OH * p.OW;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nrem) o[(img * e.map_a + (col0 + j)) * hw + pix] = f[j];
      continue;
    }
    long long orow = out_row;
    int ocol = col0;
    if (e.row_map == ROWMAP_SHUFFLE2X2) {
      const int tap = col0 / e.map_a;
      ocol = col0 % e.map_a;
      const int y = int(pix / p.OW), x = int(pix % p.OW);
      orow = (img * (2 * p.OH) + (2 * y + (tap >> 1))) * (2 * p.OW) + (2 * x + (tap & 1));
    }
    if (e.out_dtype == OUT_BF16) {
      store32bf(reinterpret_cast<__nv_bfloat16*>(e.out) + orow * e.ldo + ocol, full, nrem, f);
    } else {
      float* o = reinterpret_cast<float*>(e.out) + orow * e.ldo + ocol;
      if (full) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nrem) o[j] = f[j];
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_kmajor_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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

  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
        int x0 = 0, y0 = 0, b0 = 0;
        if (p.a_mode == 1) {
          x0 = (m_blk % p.tiles_x) * p.bw;
          y0 = ((m_blk / p.tiles_x) % p.tiles_y) * p.bh;
          b0 = (m_blk / (p.tiles_x * p.tiles_y)) * p.bb;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          uint8_t* sa = smem + ps.stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[ps.stage], C::kStageBytes);
          if (p.a_mode == 1) {
            const int tap = kb / p.cin_blocks;
            const int c0 = (kb % p.cin_blocks) * kBlockK;
            const int ky = tap / p.kw, kx = tap % p.kw;
            tma_load_4d(sa, &p.tmA, &full_bar[ps.stage], c0, x0 + kx - p.pad_x, y0 + ky - p.pad_y, b0);
          } else {
            tma_load_2d(sa, &p.tmA, &full_bar[ps.stage], kb * kBlockK, m_blk * kBlockM);
          }
          tma_load_2d(sb, &p.tmB, &full_bar[ps.stage], kb * kBlockK, n_blk * BN);
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
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN>(p, tmem_base + uint32_t(acc * BN), q, lane, m_blk, n_blk);
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

// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
      const bool mvalid = (m < p.Mc) && has_k;
      float* orow = p.out + (long long)(m % p.m_inner) * p.so_m + (long long)(m / p.m_inner) * p.so_mo +
                    (long long)tap * p.so_t;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + uint32_t(acc * BN) + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
        tmem_ld_wait();
        if (!mvalid) continue;
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

template <int BN> cudaError_t launch_gemm_t(const GemmParams& p, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kmajor_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gemm_kmajor_kernel<BN><<<grid, kThreads, Cfg<BN>::kSmemBytes, s>>>(p);
  return cudaGetLastError();
}

template <int BN> cudaError_t launch_wgrad_t(const WgradParams& p, int grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gemm_wgrad_kernel<BN><<<grid, kThreads, Cfg<BN>::kSmemBytes, s>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gemm(const GemmParams& p, int block_n, int grid, cudaStream_t s) {
  switch (block_n) {
    case 32: return launch_gemm_t<32>(p, grid, s);
    case 64: return launch_gemm_t<64>(p, grid, s);
    case 128: return launch_gemm_t<128>(p, grid, s);
    case 256: return launch_gemm_t<256>(p, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_wgrad(const WgradParams& p, int block_n, int grid, cudaStream_t s) {
  switch (block_n) {
    case 64: return launch_wgrad_t<64>(p, grid, s);
    case 128: return launch_wgrad_t<128>(p, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace dp
