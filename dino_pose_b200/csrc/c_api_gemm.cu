// Host side of the tcgen05 GEMM / implicit-conv / weight-gradient entry points: argument checking,
// TMA tensor-map encoding (driver entry point fetched through the runtime, no libcuda link
// dependency), tile-shape selection and launch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dinopose.h"
#include "gemm_tc.cuh"

namespace dp {
struct GemmVariant {
  int bn, out, act, map, opt;
  int pair;
  cudaError_t (*launch)(const GemmParams&, int grid, cudaStream_t);
};
const GemmVariant* select_gemm_variant(const Epilogue& e, int a_mode, int block_n, int pair, bool tma_out_ok);
cudaError_t launch_wgrad(const WgradParams& p, int block_n, int grid, cudaStream_t s);
cudaError_t launch_gemm_rowln(const GemmParams& p, int nacc, int mode, cudaStream_t s);

static thread_local char g_err[512] = "";
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_error(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return int(e);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    // DP_RESERVE_SMS=k: the persistent (one CTA per SM) GEMM / weight-gradient grids leave k SMs free, so that the CTAs of
    // a concurrent NCCL all-reduce can become resident beside them (data-parallel training; A/B in profiles/r2_scaling.md)
    const char* v = getenv("DP_RESERVE_SMS");
    const int k = v ? atoi(v) : 0;
    if (k > 0 && k < n - 16) n -= k;
  }
  return n;
}

// SMs left free by the persistent GEMM / weight-gradient grids launched from now on (dp_set_reserved_sms): the data-parallel
// trainer sets it while gradient all-reduces are in flight.  A persistent grid with one CTA per SM and ~200 KB of shared
// memory cannot share an SM with an NCCL CTA: the CTA that loses its SM runs its tiles as a second wave and the whole
// GEMM takes twice as long; a grid that leaves those SMs alone only loses their share of the throughput.
static int g_reserved_sms = 0;
extern "C" int dp_set_reserved_sms(int k) {
  const int prev = g_reserved_sms;
  g_reserved_sms = (k > 0 && k < sm_count() - 16) ? k : 0;
  return prev;
}
static int sm_avail() { return sm_count() - g_reserved_sms; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor map, 128B swizzle, zero OOB fill.  dims/box innermost first; strides (bytes) for dims 1..rank-1.
static int make_tmap(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(-10, "cuTensorMapEncodeTiled entry point unavailable");
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return set_error(-11, "TMA base pointer %p not 16-byte aligned", ptr);
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return set_error(-12, "TMA box dim %d = %u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    st[i] = strides[i];
    if (strides[i] & 15) return set_error(-13, "TMA stride %d = %llu bytes not a multiple of 16", i,
                                          (unsigned long long)strides[i]);
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, st, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-14, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return 0;
}

static int pow2_ge(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
// pixel box (bw, bh, bb) with bw*bh*bb == pixels for an OW x OH output map
static void choose_box(int OW, int OH, int pixels, int* bw, int* bh, int* bb) {
  int w = OW >= 16 ? 16 : pow2_ge(OW);
  if (w > pixels) w = pixels;
  int h = pow2_ge(OH);
  if (h > pixels / w) h = pixels / w;
  *bw = w;
  *bh = h;
  *bb = pixels / (w * h);
}

}  // namespace dp

using namespace dp;

extern "C" const char* dp_last_error(void) { return g_err; }
extern "C" int dp_abi_version(void) { return DP_ABI_VERSION; }
extern "C" int dp_sizeof_gemm_args(void) { return int(sizeof(dp_gemm_args)); }
extern "C" int dp_sizeof_wgrad_args(void) { return int(sizeof(dp_wgrad_args)); }

static long long* g_trace_buf = nullptr;
// tuning aid (DP_GEMM_TRACE=2): copy the trace buffer of the most recent traced launch to the host (synchronises)
extern "C" int dp_debug_read_trace(long long* host, int n) {
  if (!g_trace_buf || !host || n <= 0 || n > 4096) return set_error(-1, "dp_debug_read_trace: no trace buffer");
  cudaDeviceSynchronize();
  return cuda_error(cudaMemcpy(host, g_trace_buf, size_t(n) * sizeof(long long), cudaMemcpyDeviceToHost), "dp_debug_read_trace");
}
extern "C" int dp_gemm_bf16(const dp_gemm_args* a, void* stream) {
  if (!a || !a->A || !a->W || !a->out) return set_error(-1, "dp_gemm_bf16: null pointer");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(-2, "dp_gemm_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  // ---- epilogue
  Epilogue& e = p.epi;
  e.out = a->out; e.bias = a->bias; e.scale = a->scale; e.ls = a->ls; e.residual = static_cast<const float*>(a->residual); e.res_is_bf16 = a->res_is_bf16;
  e.aux_out = a->aux_out; e.aux_in = a->aux_in;
  e.ldo = a->ldo; e.ldr = a->ldr; e.ld_aux = a->ld_aux;
  e.out_dtype = a->out_dtype; e.act = a->act; e.row_map = a->row_map;
  e.n_valid = a->n_valid > 0 ? a->n_valid : a->N;
  e.map_a = a->map_a; e.map_b = a->map_b;
  if (e.n_valid > a->N) return set_error(-6, "dp_gemm_bf16: n_valid > N");
  if (e.row_map != DP_ROWMAP_NCHW && (e.n_valid % 4)) return set_error(-6, "dp_gemm_bf16: n_valid %% 4 != 0");
  static int dbg = -1, allow_pair = -1;
  if (dbg < 0) { const char* v = getenv("DP_GEMM_DEBUG"); dbg = v ? atoi(v) : 0; }
  // CTA-pair kernels are compiled and tested but NOT chosen automatically: at the backbone shapes they measured
  // 5-20 % slower than the single-CTA kernel (tools/gemm_tune.py); DP_GEMM_PAIR=1 or cta_pair=1 opts in
  if (allow_pair < 0) { const char* v = getenv("DP_GEMM_PAIR"); allow_pair = v ? atoi(v) : 0; }
  e.debug = dbg;
  e.stats = a->stats;
  e.stats_c = a->stats_c > 0 ? a->stats_c : e.n_valid;
  if (e.stats && e.row_map == DP_ROWMAP_NCHW) return set_error(-6, "dp_gemm_bf16: stats with NCHW row map");
  if (e.row_map == DP_ROWMAP_IDENTITY || e.row_map == DP_ROWMAP_PATCH_TOKENS) {
    const long long align = (e.out_dtype == DP_OUT_BF16) ? 8 : 4;
    if (e.ldo % align) return set_error(-7, "dp_gemm_bf16: ldo %lld must be a multiple of %lld", e.ldo, align);
  }
  if (e.row_map == DP_ROWMAP_SHUFFLE2X2 && (e.map_a % 32 || e.map_a <= 0))
    return set_error(-8, "dp_gemm_bf16: shuffle map needs Cout %% 32 == 0");
  if ((e.row_map == DP_ROWMAP_NCHW || e.row_map == DP_ROWMAP_SHUFFLE2X2) && (a->OH <= 0 || a->OW <= 0))
    return set_error(-8, "dp_gemm_bf16: row map needs OH/OW");
  if (e.row_map != DP_ROWMAP_NCHW) {
    // the epilogue addresses rows with 32-bit element offsets
    const long long out_rows = (e.row_map == DP_ROWMAP_SHUFFLE2X2) ? 4LL * a->M
                               : (e.row_map == DP_ROWMAP_PATCH_TOKENS) ? (long long)(a->M / e.map_a + 1) * e.map_b : a->M;
    if (out_rows * e.ldo + a->N >= 0xffffffffLL || (e.residual && (long long)a->M * e.ldr + a->N >= 0xffffffffLL) ||
        ((e.aux_out || e.aux_in) && (long long)a->M * e.ld_aux + a->N >= 0xffffffffLL))
      return set_error(-9, "dp_gemm_bf16: tensor too large for 32-bit epilogue offsets");
  }
  // ---- fused LayerNorm of the output rows: row-owning kernel (gemm_rowln.cu)
  if (a->ln_out || a->lora_A) {
    const bool lora = a->lora_A != nullptr;
    if (lora && a->ln_out) return set_error(-3, "dp_gemm_bf16: fused LoRA and fused LayerNorm cannot be combined");
    if (lora && (!a->lora_B || a->lora_rank != 8)) return set_error(-3, "dp_gemm_bf16: fused LoRA needs lora_B and rank 8");
    if (!lora && (!a->ln_gamma || !a->ln_beta)) return set_error(-1, "dp_gemm_bf16: ln_out without ln_gamma / ln_beta");
    if (a->a_mode != 0 || e.out_dtype != DP_OUT_F32 || e.row_map != DP_ROWMAP_IDENTITY || e.scale || e.aux_out || e.aux_in ||
        e.stats || a->act != DP_ACT_NONE || e.res_is_bf16 || (a->N != 128 && a->N != 256 && a->N != 384) || (a->K % 64))
      return set_error(-3, "dp_gemm_bf16: fused LayerNorm needs a plain fp32-output projection with N in {128,256,384}, K %% 64 == 0");
    const long long ld2 = lora ? a->ld_lora_y : a->ld_ln;
    if ((ld2 % 4) || (e.ldo % 4) || (e.residual && (e.ldr % 4)))
      return set_error(-7, "dp_gemm_bf16: fused LayerNorm / LoRA needs row pitches that are multiples of 4");
    if ((long long)a->M * ld2 + a->N >= 0xffffffffLL) return set_error(-9, "dp_gemm_bf16: tensor too large for 32-bit offsets");
    if (lora) {
      e.lora_A = a->lora_A; e.lora_B = a->lora_B; e.lora_u_out = a->lora_u_out; e.lora_seed = a->lora_seed;
      e.lora_scaling = a->lora_scaling; e.lora_p_drop = a->lora_p_drop;
      e.ln_out = a->lora_y_out; e.ld_ln = a->ld_lora_y;     // the kernel's second output slot carries y (fp32) in LoRA mode
    } else {
      e.ln_gamma = a->ln_gamma; e.ln_beta = a->ln_beta; e.ln_out = a->ln_out; e.ld_ln = a->ld_ln; e.ln_eps = a->ln_eps;
    }
    p.N = a->N; p.n_tiles = 1; p.a_mode = 0; p.M = a->M;
    p.m_tiles = (a->M + 127) / 128;
    p.num_k_blocks = a->K / 64;
    int rc2;
    {
      const uint64_t dims[2] = {uint64_t(a->K), uint64_t(a->M)};
      const uint64_t st[1] = {uint64_t(a->lda) * 2};
      const uint32_t box[2] = {64, 128};
      if ((rc2 = make_tmap(&p.tmA, a->A, 2, dims, st, box))) return rc2;
    }
    {
      const uint64_t dims[2] = {uint64_t(a->K), uint64_t(a->N)};
      const uint64_t st[1] = {uint64_t(a->ldw) * 2};
      const uint32_t box[2] = {64, 128};
      if ((rc2 = make_tmap(&p.tmB, a->W, 2, dims, st, box))) return rc2;
    }
    return cuda_error(launch_gemm_rowln(p, a->N / 128, lora ? 1 : 0, static_cast<cudaStream_t>(stream)),
                      "dp_gemm_bf16 (row-owning kernel) launch");
  }
  // ---- tile shape: CTA-pair 256 x {192, 256, 128} tiles when a variant is compiled for this epilogue and N divides,
  // else single-CTA 128 x {128, 64, 32}.  block_n fixes the width, cta_pair (1 pair / 2 single) the kind.
  // plain bf16 outputs (QKV, fc1) are written with TMA tile stores when the output qualifies (DP_GEMM_TMA_OUT=0: off)
  static int allow_tma_out = -1;
  if (allow_tma_out < 0) { const char* v = getenv("DP_GEMM_TMA_OUT"); allow_tma_out = v ? atoi(v) : 1; }
  const bool tma_out_ok = allow_tma_out && a->a_mode == 0 && e.row_map == DP_ROWMAP_IDENTITY && e.out_dtype == DP_OUT_BF16 &&
                          !e.scale && !e.ls && !e.residual && !e.aux_out && !e.aux_in && !e.stats && a->act != DP_ACT_RELU &&
                          (e.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 && dbg == 0;
  const GemmVariant* var = nullptr;
  const bool may_pair = (allow_pair || a->cta_pair == 1) && a->cta_pair != 2 && a->cta_pair < 3 && a->M > 128;
  const bool may_single = a->cta_pair != 1 && a->cta_pair < 3;
  // A-stationary TS-mode kernel (gemm_astat.cuh): plain K-major A with K <= 512, 128-wide tiles, several tiles per row block.
  // cta_pair = 3 requests it, DP_GEMM_ASTAT=0 keeps it out of the automatic choice.
  static int allow_astat = -1;
  if (allow_astat < 0) { const char* v = getenv("DP_GEMM_ASTAT"); allow_astat = v ? atoi(v) : 0; }
  const bool astat_shape = a->a_mode == 0 && a->K <= 512 && (a->K % 64) == 0 && a->N >= 256 && e.row_map == DP_ROWMAP_IDENTITY &&
                           (a->block_n == 0 || a->block_n == 128) && !e.stats && !e.scale;
  // cta_pair: 3 = A-stationary, 4 = A-stationary in clusters of two (multicast weight loads); DP_GEMM_ASTAT = 0 off, 1 single
  // CTAs, 2 clusters for the automatic choice
  if (astat_shape && (a->cta_pair == 3 || (a->cta_pair == 0 && allow_astat == 1 && a->M >= 2048)))
    var = select_gemm_variant(e, a->a_mode, 128, 2, tma_out_ok);
  if (astat_shape && (a->N % 128) == 0 && (a->cta_pair == 4 || (a->cta_pair == 0 && allow_astat == 2 && a->M >= 2048)))
    var = select_gemm_variant(e, a->a_mode, 128, 3, tma_out_ok);
  if (a->cta_pair >= 3 && !var) return set_error(-3, "dp_gemm_bf16: no A-stationary variant for this shape / epilogue");
  if (var) {
  } else if (a->block_n != 0) {
    if (may_pair) var = select_gemm_variant(e, a->a_mode, a->block_n, 1, tma_out_ok);
    if (!var && may_single) var = select_gemm_variant(e, a->a_mode, a->block_n, 0, tma_out_ok);
  } else {
    if (may_pair) {
      const int cand[3] = {192, 256, 128};
      for (int i = 0; i < 3 && !var; ++i)
        if (a->N % cand[i] == 0 && a->N >= 2 * cand[i] - 128) var = select_gemm_variant(e, a->a_mode, cand[i], 1, tma_out_ok);
    }
    if (!var && may_single) var = select_gemm_variant(e, a->a_mode, a->N >= 128 ? 128 : (a->N > 32 ? 64 : 32), 0, tma_out_ok);
  }
  if (!var) return set_error(-3, "dp_gemm_bf16: no kernel variant for block_n %d (cta_pair %d)", a->block_n, a->cta_pair);
  const int bn = var->bn;
  p.N = a->N;
  p.n_tiles = (a->N + bn - 1) / bn;
  p.a_mode = a->a_mode;
  int rc;
  if (a->a_mode == 0) {
    p.M = a->M;
    p.m_tiles = (a->M + 127) / 128;
    p.num_k_blocks = (a->K + 63) / 64;
    const uint64_t dims[2] = {uint64_t(a->K), uint64_t(a->M)};
    const uint64_t st[1] = {uint64_t(a->lda) * 2};
    const uint32_t box[2] = {64, 128};
    if ((rc = make_tmap(&p.tmA, a->A, 2, dims, st, box))) return rc;
    p.OH = a->OH > 0 ? a->OH : 1;
    p.OW = a->OW > 0 ? a->OW : 1;
    p.NB = a->NB;
  } else {
    if (a->C % 64) return set_error(-4, "dp_gemm_bf16: implicit conv needs C %% 64 == 0 (C=%d)", a->C);
    if (a->K != a->KH * a->KW * a->C) return set_error(-5, "dp_gemm_bf16: K != KH*KW*C");
    if (a->M != a->NB * a->OH * a->OW) return set_error(-5, "dp_gemm_bf16: M != NB*OH*OW");
    choose_box(a->OW, a->OH, 128, &p.bw, &p.bh, &p.bb);
    p.bw_log2 = __builtin_ctz(p.bw);
    p.bh_log2 = __builtin_ctz(p.bh);
    p.OW = a->OW; p.OH = a->OH; p.NB = a->NB;
    p.tiles_x = (a->OW + p.bw - 1) / p.bw;
    p.tiles_y = (a->OH + p.bh - 1) / p.bh;
    const int tiles_b = (a->NB + p.bb - 1) / p.bb;
    p.M = a->M;
    p.m_tiles = p.tiles_x * p.tiles_y * tiles_b;
    p.kw = a->KW; p.pad_x = a->pad_x; p.pad_y = a->pad_y;
    p.cin_blocks = a->C / 64;
    p.num_k_blocks = a->KH * a->KW * p.cin_blocks;
    const uint64_t dims[4] = {uint64_t(a->C), uint64_t(a->IW), uint64_t(a->IH), uint64_t(a->NB)};
    const uint64_t st[3] = {uint64_t(a->a_stride_w) * 2, uint64_t(a->a_stride_h) * 2, uint64_t(a->a_stride_b) * 2};
    const uint32_t box[4] = {64, uint32_t(p.bw), uint32_t(p.bh), uint32_t(p.bb)};
    if ((rc = make_tmap(&p.tmA, a->A, 4, dims, st, box))) return rc;
  }
  {
    const uint64_t dims[2] = {uint64_t(a->K), uint64_t(a->N)};
    const uint64_t st[1] = {uint64_t(a->ldw) * 2};
    const uint32_t box[2] = {64, uint32_t((var->pair == 1 || var->pair == 3) ? bn / 2 : bn)};   // a pair CTA stages half of the weight rows
    if ((rc = make_tmap(&p.tmB, a->W, 2, dims, st, box))) return rc;
  }
  if (var->opt & 128) {   // OP_TMA_OUT: 32 x 32 bf16 boxes, 64-byte rows, 64B swizzle; columns >= n_valid and rows >= M are clipped
    const uint64_t dims[2] = {uint64_t(e.n_valid), uint64_t(a->M)};
    const uint64_t st[1] = {uint64_t(e.ldo) * 2};
    const uint32_t box[2] = {32, 32};
    if ((rc = make_tmap(&p.tmC, e.out, 2, dims, st, box, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  }
  int grid;
  if (var->pair == 2) {
    const int tiles = p.m_tiles * p.n_tiles;
    grid = tiles < sm_avail() ? tiles : sm_avail();
  } else if (var->pair == 3) {
    const int tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int clusters = tiles < sm_avail() / 2 ? tiles : sm_avail() / 2;
    grid = 2 * clusters;
  } else if (var->pair) {
    const int tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int clusters = tiles < sm_avail() / 2 ? tiles : sm_avail() / 2;
    grid = 2 * clusters;
  } else {
    const int tiles = p.m_tiles * p.n_tiles;
    grid = tiles < sm_avail() ? tiles : sm_avail();
  }
  static int trace_on = -1;
  if (trace_on < 0) { const char* v = getenv("DP_GEMM_TRACE"); trace_on = v ? atoi(v) : 0; }
  if (trace_on == 2 && var->pair != 1) {
    // asynchronous probe: CTA 0 of every launch overwrites the buffer; dp_debug_read_trace() fetches the last one
    if (!g_trace_buf) cudaMalloc(&g_trace_buf, 4096 * sizeof(long long));
    p.epi.trace = g_trace_buf;
    return cuda_error(var->launch(p, grid, static_cast<cudaStream_t>(stream)), "dp_gemm_bf16 launch");
  }
  if (trace_on && var->pair != 1) {
    // debug only (synchronises): per-tile timeline of CTA 0, cycles relative to the first MMA start
    static long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 4096 * sizeof(long long));
    cudaMemsetAsync(dbuf, 0, 4096 * sizeof(long long), static_cast<cudaStream_t>(stream));
    p.epi.trace = dbuf;
    cudaError_t le = var->launch(p, grid, static_cast<cudaStream_t>(stream));
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    static long long host[4096];
    cudaMemcpy(host, dbuf, sizeof(host), cudaMemcpyDeviceToHost);
    const int per_cta = (p.m_tiles * p.n_tiles + grid - 1) / grid;
    fprintf(stderr, "[trace] M=%d N=%d K=%d bn=%d tiles/CTA=%d: tile: mma_start mma_issued epi_start epi_end (cycles)\n", a->M,
            a->N, a->K, bn, per_cta);
    for (int t = 0; t < per_cta && t < 12; ++t)
      fprintf(stderr, "[trace]  %2d: %8lld %8lld %8lld %8lld\n", t, host[4 * t] - host[0], host[4 * t + 1] - host[0],
              host[4 * t + 2] - host[0], host[4 * t + 3] - host[0]);
    for (int t = 0; t < per_cta && t < 12; ++t) {
      const long long* r = host + 2048 + 8 * t;
      fprintf(stderr, "[trace]  %2d epilogue warp 2 (first chunk), relative to epi_start: ldtm_issue %lld ldtm_done %lld staged %lld "
              "batch0 %lld batch1 %lld fenced %lld arrived %lld\n", t, r[0] - host[4 * t + 2], r[1] - host[4 * t + 2],
              r[2] - host[4 * t + 2], r[3] - host[4 * t + 2], r[4] - host[4 * t + 2], r[5] - host[4 * t + 2],
              host[4 * t + 3] - host[4 * t + 2]);
    }
    return cuda_error(le, "dp_gemm_bf16 launch");
  }
  return cuda_error(var->launch(p, grid, static_cast<cudaStream_t>(stream)), "dp_gemm_bf16 launch");
}

extern "C" int dp_wgrad_bf16(const dp_wgrad_args* a, void* stream) {
  if (!a || !a->A || !a->B || !a->out) return set_error(-1, "dp_wgrad_bf16: null pointer");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int bn = a->block_n;
  if (bn == 0) bn = a->Nc >= 128 ? 128 : 64;
  if (bn != 64 && bn != 128) return set_error(-3, "dp_wgrad_bf16: block_n %d", bn);
  p.out = a->out;
  p.so_m = a->so_m; p.so_mo = a->so_mo; p.so_n = a->so_n; p.so_no = a->so_no; p.so_t = a->so_t;
  p.m_inner = a->m_inner > 0 ? a->m_inner : 0x7fffffff;
  p.n_inner = a->n_inner > 0 ? a->n_inner : 0x7fffffff;
  p.Mc = a->Mc; p.Nc = a->Nc;
  p.m_tiles = (a->Mc + 127) / 128;
  p.n_tiles = (a->Nc + bn - 1) / bn;
  p.mode = a->mode;
  int rc;
  if (a->mode == 0) {
    p.taps = 1; p.kw = 1;
    p.total_k_blocks = (a->P + 63) / 64;
    const uint64_t da[2] = {uint64_t(a->Mc), uint64_t(a->P)};
    const uint64_t sa[1] = {uint64_t(a->lda) * 2};
    const uint64_t db[2] = {uint64_t(a->Nc), uint64_t(a->P)};
    const uint64_t sb[1] = {uint64_t(a->ldb) * 2};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap(&p.tmA, a->A, 2, da, sa, box))) return rc;
    if ((rc = make_tmap(&p.tmB, a->B, 2, db, sb, box))) return rc;
  } else {
    p.taps = a->KH * a->KW; p.kw = a->KW; p.pad_x = a->pad_x; p.pad_y = a->pad_y;
    choose_box(a->OW, a->OH, 64, &p.bw, &p.bh, &p.bb);
    p.tiles_x = (a->OW + p.bw - 1) / p.bw;
    p.tiles_y = (a->OH + p.bh - 1) / p.bh;
    p.tiles_b = (a->NB + p.bb - 1) / p.bb;
    p.total_k_blocks = p.tiles_x * p.tiles_y * p.tiles_b;
    const uint64_t da[4] = {uint64_t(a->Mc), uint64_t(a->OW), uint64_t(a->OH), uint64_t(a->NB)};
    const uint64_t sa[3] = {uint64_t(a->a_sw) * 2, uint64_t(a->a_sh) * 2, uint64_t(a->a_sb) * 2};
    const uint64_t db[4] = {uint64_t(a->Nc), uint64_t(a->IW), uint64_t(a->IH), uint64_t(a->NB)};
    const uint64_t sb[3] = {uint64_t(a->b_sw) * 2, uint64_t(a->b_sh) * 2, uint64_t(a->b_sb) * 2};
    const uint32_t box[4] = {64, uint32_t(p.bw), uint32_t(p.bh), uint32_t(p.bb)};
    if ((rc = make_tmap(&p.tmA, a->A, 4, da, sa, box))) return rc;
    if ((rc = make_tmap(&p.tmB, a->B, 4, db, sb, box))) return rc;
  }
  const int base_items = p.taps * p.m_tiles * p.n_tiles;
  const long long slice_bytes = (long long)p.taps * p.m_tiles * 128 * p.n_tiles * bn * 4;   // workspace per split
  const bool use_ws = a->workspace != nullptr && a->workspace_bytes >= slice_bytes;
  int splits = a->splits;
  if (splits <= 0) {
    if (use_ws) {
      // atomic-free path: pick the split count with the best SM utilisation (rounds of 148 items per unit of K),
      // a small linear term for the workspace traffic, at least 8 k-blocks per split, workspace permitting
      double best = 1e30;
      splits = 1;
      double per_split = double(slice_bytes) / 4.0e8;   // workspace write + read relative to the operand streaming
      if (per_split < 0.0005) per_split = 0.0005;
      if (per_split > 0.01) per_split = 0.01;
      for (int sN = 1; sN <= 148; ++sN) {
        if (sN > 1 && (p.total_k_blocks / sN < 8 || (long long)sN * slice_bytes > a->workspace_bytes)) break;
        const int rounds = (base_items * sN + sm_count() - 1) / sm_count();
        const double cost = double(rounds) / sN + per_split * sN;
        if (cost < best - 1e-9) { best = cost; splits = sN; }
      }
    } else {
      splits = (2 * sm_count() + base_items - 1) / base_items;
      const int max_by_k = (p.total_k_blocks + 3) / 4;  // at least 4 k-blocks per split
      if (splits > max_by_k) splits = max_by_k;
      if (splits < 1) splits = 1;
    }
  }
  if (use_ws && (long long)splits * slice_bytes > a->workspace_bytes) splits = int(a->workspace_bytes / slice_bytes);
  if (splits > p.total_k_blocks) splits = p.total_k_blocks;
  // make sure no split is empty
  while (splits > 1 && ((p.total_k_blocks + splits - 1) / splits) * (splits - 1) >= p.total_k_blocks) --splits;
  p.splits = splits;
  if (use_ws) {
    if (reinterpret_cast<uintptr_t>(a->workspace) & 15) return set_error(-11, "dp_wgrad_bf16: workspace not 16-byte aligned");
    p.ws = static_cast<float*>(a->workspace);
    p.ws_ld = p.n_tiles * bn;
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* v = getenv("DP_WGRAD_DEBUG"); dbg = v ? atoi(v) : 0; }
    p.debug = dbg;
  }
  const int items = base_items * splits;
  const int grid = items < sm_avail() ? items : sm_avail();
  return cuda_error(launch_wgrad(p, bn, grid, static_cast<cudaStream_t>(stream)), "dp_wgrad_bf16 launch");
}
