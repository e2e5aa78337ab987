// Row-owning tcgen05 GEMM with the NEXT LayerNorm fused into its epilogue: the attention / MLP output projections of the
// frozen encoder layers (HF modeling_dinov2.py:250 + 373-379, 327 + 382-384 followed by :371 / :379 of the next sub-block).
//
//   v[M, N]  = residual + ls * (A[M,K] * W[N,K]^T + bias)        fp32 residual stream, written to `out` (may alias residual)
//   ln_out   = LayerNorm(v; gamma, beta, eps)                    bf16, the A operand of the following QKV / fc1 GEMM
//
// One CTA owns 128 COMPLETE rows (N = 128 * NACC <= 512 fp32 accumulator columns in tensor memory: N = 384 for ViT-S), so
// the row statistics never leave the SM and the separate LayerNorm launch (25 per forward: 38 MB of traffic and a launch
// gap each) disappears.  Main loop: TMA ring of {A 128 x 64, W N x 64} stages, per 16-wide k-step one N = 256 and one
// N = 128 MMA (NACC = 3).  Epilogue (16 warps; warp (q, g) owns TMEM lanes [32q, 32q+32) and the 32-column chunks c with
// c % 4 == g):
//   pass 1  TMEM -> registers -> 32x32 shared-memory transpose -> coalesced row segments: v written to `out`, per-row
//           partial sums of v and v^2 kept in registers, reduced over the 8 lanes that share a row, then over the four
//           column-group warps through shared memory;
//   pass 2  every lane re-reads the v it wrote itself (L1 / L2 hot), normalises and writes bf16.
// The variance is E[v^2] - E[v]^2 in fp32 over N <= 512 values (relative error ~1e-6 * mean^2 / var; the residual stream
// of a ViT is close to zero-mean across channels).  129 tiles at batch 64 = one wave on 148 SMs: the kernel trades 13 % of
// the SMs for not writing, re-reading and re-launching.
//
// MODE 1 of the same kernel fuses the LoRA adapter of the last block instead (reference model/lora.py:26-28,53-59 on the
// output of `attention.output.dense`, then LayerScale + residual, HF:373-376):
//   y = A W^T + bias;  u = y lora_A  [rank 8];  x_out = x_in + lambda1 * (y + s * dropout(u lora_B))
// The rank-8 side product needs complete rows of y -- which is exactly what the row-owning tile has in tensor memory: phase
// A (thread = row, straight out of TMEM) accumulates u, phase B (transposed, coalesced) applies u lora_B, the dropout mask
// (same counter-based hash as dp_lora_bwd), LayerScale and the residual.  y and u are saved for the backward.
#include "gemm_kernel.cuh"

namespace dp {

constexpr int kRlStages = 3;
constexpr int kLoraR = 8;

template <int NACC> struct RlCfg {
  static constexpr int kN = 128 * NACC;
  static constexpr int kStageBytes = kABytes + kN * kBlockK * 2;            // 16 KB + NACC x 16 KB
  static constexpr int kRingBytes = kRlStages * kStageBytes;
  // LayerNorm: [row][column group][sum, sumsq] (4 KB); LoRA: partial u [row][column group][8] + total u [row][8] (20 KB)
  static constexpr int kStatBytes = 128 * 4 * kLoraR * 4 + 128 * kLoraR * 4;
  static constexpr int kSmemBytes = kRingBytes + kStatBytes + 256;
  static_assert(kRingBytes >= kStagingBytes, "the epilogue staging aliases the operand ring");
  static_assert(kSmemBytes <= kSmemLimit, "shared memory");
  static_assert(kN <= 512, "accumulator columns");
};

template <int NACC, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_rowln_kernel(const __grid_constant__ GemmParams p) {
  using C = RlCfg<NACC>;
  extern __shared__ __align__(1024) uint8_t smem_gemm[];
  uint8_t* smem = smem_gemm;
  float* staging = reinterpret_cast<float*>(smem);      // aliases the ring: used only after the last MMA has completed
  float* rowstat = reinterpret_cast<float*>(smem + C::kRingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kRingBytes + C::kStatBytes);
  uint64_t* empty_bar = full_bar + kRlStages;
  uint64_t* tfull_bar = empty_bar + kRlStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Epilogue& e = p.epi;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kRlStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_grid_sync();
  const int m_blk = blockIdx.x;
  const int nkb = p.num_k_blocks;

  if (warp == 0) {
    if (elect_one()) {
      PipeState ps;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
        uint8_t* sa = smem + ps.stage * C::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[ps.stage], C::kStageBytes);
        tma_load_2d(sa, &p.tmA, &full_bar[ps.stage], kb * kBlockK, m_blk * kBlockM);
#pragma unroll
        for (int a = 0; a < NACC; ++a)
          tma_load_2d(sa + kABytes + a * kABytes, &p.tmB, &full_bar[ps.stage], kb * kBlockK, a * 128);
        ps.template advance<kRlStages>();
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      PipeState ps;
      constexpr int kWide = NACC >= 2 ? 256 : 128;                 // first MMA of a k-step
      constexpr int kRest = C::kN - kWide;                          // second (0, 128 or 256 columns)
      constexpr uint32_t idesc_w = make_idesc_bf16(kBlockM, kWide, 0, 0);
      constexpr uint32_t idesc_r = make_idesc_bf16(kBlockM, kRest > 0 ? kRest : 128, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[ps.stage], ps.phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + ps.stage * C::kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t adesc = make_sdesc_sw128(a_addr + k * 32, 0, 1024);
          umma_bf16(tmem_base, adesc, make_sdesc_sw128(b_addr + k * 32, 0, 1024), idesc_w, (kb | k) != 0 ? 1u : 0u);
          if constexpr (kRest > 0)
            umma_bf16(tmem_base + kWide, adesc, make_sdesc_sw128(b_addr + kWide * 128 + k * 32, 0, 1024), idesc_r,
                      (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[ps.stage]);
        ps.template advance<kRlStages>();
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int q = warp & 3;             // TMEM lane quarter
    const int g = (warp - 2) >> 2;      // column group: chunks g, g + 4, g + 8, ...
    float* stg = staging + (warp - 2) * (32 * 32);
    const int rr = lane >> 3, cg = lane & 7;
    constexpr uint32_t kFull = 0xffffffffu;
    const int r = q * 32 + lane;
    const long long row = (long long)m_blk * kBlockM + r;
    const bool valid = row < p.M;
    const uint32_t off_out = valid ? uint32_t(row) * uint32_t(e.ldo) : kInvalidRow;
    const uint32_t off_res = uint32_t(valid ? row : 0) * uint32_t(e.ldr);
    const uint32_t off_ln = uint32_t(valid ? row : 0) * uint32_t(e.ld_ln);
    if constexpr (MODE == 1) {
      // ================= LoRA epilogue
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      float* upart = rowstat;                               // [128][4][8]
      float* utot = rowstat + 128 * 4 * kLoraR;             // [128][8]
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
      {
        // ---- phase A (thread = row): u = y lora_A over this warp's column chunks
        float u[kLoraR];
#pragma unroll
        for (int k = 0; k < kLoraR; ++k) u[k] = 0.f;
#pragma unroll 1
        for (int c = g; c < C::kN / 32; c += kEpiGroups) {
          uint32_t v[32];
          tmem_ld_32x32(trow + uint32_t(c * 32), v);
          tmem_ld_wait();
          const float bl = e.bias != nullptr ? __ldg(e.bias + c * 32 + lane) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float y = __uint_as_float(v[j]) + __shfl_sync(kFull, bl, j);
            const float4 a0 = ldg4(e.lora_A + (c * 32 + j) * kLoraR), a1 = ldg4(e.lora_A + (c * 32 + j) * kLoraR + 4);
            u[0] = fmaf(y, a0.x, u[0]); u[1] = fmaf(y, a0.y, u[1]); u[2] = fmaf(y, a0.z, u[2]); u[3] = fmaf(y, a0.w, u[3]);
            u[4] = fmaf(y, a1.x, u[4]); u[5] = fmaf(y, a1.y, u[5]); u[6] = fmaf(y, a1.z, u[6]); u[7] = fmaf(y, a1.w, u[7]);
          }
        }
        float4* dst = reinterpret_cast<float4*>(upart + (r * 4 + g) * kLoraR);
        dst[0] = make_float4(u[0], u[1], u[2], u[3]);
        dst[1] = make_float4(u[4], u[5], u[6], u[7]);
      }
      named_bar_sync(1, 32 * kEpiWarps);
      {
        const int et = threadIdx.x - 64;
        if (et < 128) {                                     // one thread per row: total u, saved for the backward
          float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const float4 a = *reinterpret_cast<const float4*>(upart + (et * 4 + gg) * kLoraR);
            const float4 b = *reinterpret_cast<const float4*>(upart + (et * 4 + gg) * kLoraR + 4);
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
            s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
          }
          *reinterpret_cast<float4*>(utot + et * kLoraR) = s0;
          *reinterpret_cast<float4*>(utot + et * kLoraR + 4) = s1;
          const long long grow = (long long)m_blk * kBlockM + et;
          if (e.lora_u_out != nullptr && grow < p.M) {
            *reinterpret_cast<float4*>(e.lora_u_out + grow * kLoraR) = s0;
            *reinterpret_cast<float4*>(e.lora_u_out + grow * kLoraR + 4) = s1;
          }
        }
      }
      named_bar_sync(1, 32 * kEpiWarps);
      // ---- phase B (transposed, coalesced): x_out = x_in + lambda1 * (y + s * dropout(u lora_B)); y saved
      const unsigned long long seed = e.lora_seed != nullptr ? *e.lora_seed : 0ull;
      const uint32_t thresh = e.lora_p_drop > 0.f ? uint32_t(fminf(e.lora_p_drop, 0.999999f) * 4294967296.0f) : 0u;
      const float keep_scale = e.lora_p_drop > 0.f ? 1.0f / (1.0f - e.lora_p_drop) : 1.0f;
      const uint32_t off_y = uint32_t(valid ? row : 0) * uint32_t(e.ld_ln);
#pragma unroll 1
      for (int c = g; c < C::kN / 32; c += kEpiGroups) {
        const int ccol = c * 32 + cg * 4;
        const float4 bi = e.bias != nullptr ? ldg4(e.bias + ccol) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 lsv = e.ls != nullptr ? ldg4(e.ls + ccol) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 bm[kLoraR];
#pragma unroll
        for (int k = 0; k < kLoraR; ++k) bm[k] = ldg4(e.lora_B + (long long)k * p.N + ccol);
        float4 res[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rw = it * 4 + rr;
          const bool ok = __shfl_sync(kFull, off_out, rw) != kInvalidRow;
          const uint32_t r_off = __shfl_sync(kFull, off_res, rw);
          res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && e.residual != nullptr) res[it] = ldg4(e.residual + r_off + ccol);
        }
        uint32_t v[32];
        tmem_ld_32x32(trow + uint32_t(c * 32), v);
        tmem_ld_wait();
        {
          float4* srow = reinterpret_cast<float4*>(stg + lane * 32);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            srow[jj ^ (lane & 7)] = make_float4(__uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]),
                                                __uint_as_float(v[4 * jj + 2]), __uint_as_float(v[4 * jj + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rw = it * 4 + rr;
          const float4 x = *reinterpret_cast<const float4*>(stg + rw * 32 + ((cg ^ (rw & 7)) << 2));
          const uint32_t o_off = __shfl_sync(kFull, off_out, rw);
          const uint32_t y_off = __shfl_sync(kFull, off_y, rw);
          const float4 u0 = *reinterpret_cast<const float4*>(utot + (q * 32 + rw) * kLoraR);
          const float4 u1 = *reinterpret_cast<const float4*>(utot + (q * 32 + rw) * kLoraR + 4);
          const float uu[kLoraR] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
          float y[4] = {x.x + bi.x, x.y + bi.y, x.z + bi.z, x.w + bi.w};
          float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < kLoraR; ++k) {
            d[0] = fmaf(uu[k], bm[k].x, d[0]); d[1] = fmaf(uu[k], bm[k].y, d[1]);
            d[2] = fmaf(uu[k], bm[k].z, d[2]); d[3] = fmaf(uu[k], bm[k].w, d[3]);
          }
          if (o_off != kInvalidRow) {
            const long long grow = (long long)m_blk * kBlockM + q * 32 + rw;
            if (e.lora_p_drop > 0.f) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                d[k] = dropout_keep(seed, uint64_t(grow) * uint64_t(p.N) + uint64_t(ccol + k), thresh) ? d[k] * keep_scale : 0.f;
            }
            float4 o;
            o.x = fmaf(fmaf(d[0], e.lora_scaling, y[0]), lsv.x, res[it].x);
            o.y = fmaf(fmaf(d[1], e.lora_scaling, y[1]), lsv.y, res[it].y);
            o.z = fmaf(fmaf(d[2], e.lora_scaling, y[2]), lsv.z, res[it].z);
            o.w = fmaf(fmaf(d[3], e.lora_scaling, y[3]), lsv.w, res[it].w);
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + o_off + ccol) = o;
            if (e.ln_out != nullptr)     // y = the projection output itself (fp32), what dp_lora_bwd reads
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.ln_out) + y_off + ccol) = make_float4(y[0], y[1], y[2], y[3]);
          }
        }
        __syncwarp();
      }
    } else {
    float rs1[8], rs2[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) rs1[it] = rs2[it] = 0.f;
    bool waited = false;
    // ---- pass 1
#pragma unroll 1
    for (int c = g; c < C::kN / 32; c += kEpiGroups) {
      const int ccol = c * 32 + cg * 4;
      const float4 bi = e.bias != nullptr ? ldg4(e.bias + ccol) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 lsv = e.ls != nullptr ? ldg4(e.ls + ccol) : make_float4(1.f, 1.f, 1.f, 1.f);
      float4 res[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rw = it * 4 + rr;
        const bool ok = __shfl_sync(kFull, off_out, rw) != kInvalidRow;
        const uint32_t r_off = __shfl_sync(kFull, off_res, rw);
        res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && e.residual != nullptr) res[it] = ldg4(e.residual + r_off + ccol);
      }
      if (!waited) {
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
        waited = true;
      }
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * 32), v);
      tmem_ld_wait();
      {
        float4* srow = reinterpret_cast<float4*>(stg + lane * 32);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          srow[jj ^ (lane & 7)] = make_float4(__uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]),
                                              __uint_as_float(v[4 * jj + 2]), __uint_as_float(v[4 * jj + 3]));
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rw = it * 4 + rr;
        const float4 x = *reinterpret_cast<const float4*>(stg + rw * 32 + ((cg ^ (rw & 7)) << 2));
        const uint32_t o_off = __shfl_sync(kFull, off_out, rw);
        float4 f;
        f.x = fmaf(x.x + bi.x, lsv.x, res[it].x);
        f.y = fmaf(x.y + bi.y, lsv.y, res[it].y);
        f.z = fmaf(x.z + bi.z, lsv.z, res[it].z);
        f.w = fmaf(x.w + bi.w, lsv.w, res[it].w);
        if (o_off != kInvalidRow) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + o_off + ccol) = f;
          rs1[it] += (f.x + f.y) + (f.z + f.w);
          rs2[it] += fmaf(f.x, f.x, f.y * f.y) + fmaf(f.z, f.z, f.w * f.w);
        }
      }
      __syncwarp();   // the staging buffer is rewritten by the next chunk
    }
    // ---- row statistics: over the 8 lanes that share a row, then over the 4 column-group warps
#pragma unroll
    for (int it = 0; it < 8; ++it) {
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        rs1[it] += __shfl_xor_sync(kFull, rs1[it], o);
        rs2[it] += __shfl_xor_sync(kFull, rs2[it], o);
      }
      if (cg == 0) {
        const int rw = q * 32 + it * 4 + rr;
        rowstat[(rw * 4 + g) * 2] = rs1[it];
        rowstat[(rw * 4 + g) * 2 + 1] = rs2[it];
      }
    }
    named_bar_sync(1, 32 * kEpiWarps);
    float mean[8], rstd[8];
    const float inv_n = 1.0f / float(C::kN);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rw = q * 32 + it * 4 + rr;
      const float4 a = *reinterpret_cast<const float4*>(rowstat + rw * 8);
      const float4 b = *reinterpret_cast<const float4*>(rowstat + rw * 8 + 4);
      const float s1 = (a.x + a.z) + (b.x + b.z), s2 = (a.y + a.w) + (b.y + b.w);
      mean[it] = s1 * inv_n;
      rstd[it] = rsqrtf(fmaxf(s2 * inv_n - mean[it] * mean[it], 0.f) + e.ln_eps);
    }
    // ---- pass 2: every lane re-reads the values it wrote in pass 1 (same thread, program order), normalises, writes bf16
#pragma unroll 1
    for (int c = g; c < C::kN / 32; c += kEpiGroups) {
      const int ccol = c * 32 + cg * 4;
      const float4 ga = ldg4(e.ln_gamma + ccol), be = ldg4(e.ln_beta + ccol);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rw = it * 4 + rr;
        const uint32_t o_off = __shfl_sync(kFull, off_out, rw);
        const uint32_t l_off = __shfl_sync(kFull, off_ln, rw);
        if (o_off != kInvalidRow) {
          const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.out) + o_off + ccol);
          const float m = mean[it], s = rstd[it];
          uint2 t;
          t.x = pack_bf16x2(fmaf((f.x - m) * s, ga.x, be.x), fmaf((f.y - m) * s, ga.y, be.y));
          t.y = pack_bf16x2(fmaf((f.z - m) * s, ga.z, be.z), fmaf((f.w - m) * s, ga.w, be.w));
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.ln_out) + l_off + ccol) = t;
        }
      }
    }
    }   // MODE
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int NACC, int MODE>
cudaError_t launch_gemm_rowln_t(const GemmParams& p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_rowln_kernel<NACC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         RlCfg<NACC>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  launch_k<gemm_rowln_kernel<NACC, MODE>>(p.m_tiles, kGemmThreads, RlCfg<NACC>::kSmemBytes, s, p);
  return cudaGetLastError();
}

// N = 128 * nacc accumulator columns; nacc in {1, 2, 3} (shared memory bounds the operand ring: nacc = 4 needs 80 KB stages)
// mode 0: fused LayerNorm epilogue; mode 1: fused rank-8 LoRA adapter + LayerScale + residual
cudaError_t launch_gemm_rowln(const GemmParams& p, int nacc, int mode, cudaStream_t s) {
  switch (nacc * 2 + (mode ? 1 : 0)) {
    case 2: return launch_gemm_rowln_t<1, 0>(p, s);
    case 3: return launch_gemm_rowln_t<1, 1>(p, s);
    case 4: return launch_gemm_rowln_t<2, 0>(p, s);
    case 5: return launch_gemm_rowln_t<2, 1>(p, s);
    case 6: return launch_gemm_rowln_t<3, 0>(p, s);
    case 7: return launch_gemm_rowln_t<3, 1>(p, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace dp
