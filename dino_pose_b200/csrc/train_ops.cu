// Training-step kernels around the model: the reference losses with their backward seed, and AdamW.
//
//  pose_loss_*   reference train.py:89-120 + DynamicLossWeighting (:17-69), fused: one reduction pass over the
//                heat-maps, a one-thread finalize that keeps the loss-weight EMA state ON THE DEVICE (the
//                reference does 7 .item() host syncs per step, train.py:155-156,173-178), and one element-wise
//                pass that writes d(loss)/d(heatmaps), d(loss)/d(z) -- the seeds of the backward program.
//  adamw         torch.optim.AdamW(lr, weight_decay) (train.py:280-284) over ONE flat fp32 parameter buffer with
//                the flat gradient buffer the backward program fills; bias correction from a device step counter;
//                gradients pre-scaled by 1/world_size after the data-parallel all-reduce.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.cuh"

namespace dp {
namespace {

// sums[0] += sum exp(-d^2) d^2 [conf > 1],  d = pred - target (heat-maps);  sums[1] += sum |pz*m - tz*m|
__global__ void __launch_bounds__(256) pose_loss_reduce_kernel(const float4* __restrict__ hm, const float4* __restrict__ thm,
                                                               const float* __restrict__ kps, int kp_stride,
                                                               const float* __restrict__ z, const float* __restrict__ tz,
                                                               double* __restrict__ sums, long long n4, int map4, int BK) {
  pdl_grid_sync();
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int bk = int(i / map4);
    if (__ldg(kps + (long long)bk * kp_stride + 2) > 1.f) {
      const float4 a = __ldg(hm + i), b = __ldg(thm + i);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      const float q0 = d0 * d0, q1 = d1 * d1, q2 = d2 * d2, q3 = d3 * d3;
      acc += __expf(-q0) * q0 + __expf(-q1) * q1 + __expf(-q2) * q2 + __expf(-q3) * q3;
    }
  }
  float zacc = 0.f;
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < BK; i += blockDim.x) {
      const float m = __ldg(kps + (long long)i * kp_stride + 2) > 1.f ? 1.f : 0.f;
      zacc += fabsf(z[i] * m - tz[i] * m);
    }
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    zacc += __shfl_xor_sync(0xffffffffu, zacc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = acc;
    red[1][threadIdx.x >> 5] = zacc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    atomicAdd(sums, double(a));
    if (blockIdx.x == 0) atomicAdd(sums + 1, double(b));
  }
}

// state: [0] kp_avg  [1] z_avg  [2] started (0/1)  [3] weight.   out: [0] balanced loss  [1] kp  [2] z
// scales: [0] d(loss)/d(hm) factor = 2 / (numel_hm * (kp_avg + 1e-8)),  [1] 1 / (numel_z * (z_avg + 1e-8))
__global__ void pose_loss_finalize_kernel(double* __restrict__ sums, float* __restrict__ state, float* __restrict__ out,
                                          float* __restrict__ scales, double numel_hm, double numel_z, float momentum,
                                          float rate) {
  pdl_grid_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float kp = float(sums[0] / numel_hm), zl = float(sums[1] / numel_z);
  sums[0] = 0.0;
  sums[1] = 0.0;
  float kp_avg, z_avg;
  if (state[2] == 0.f) {
    kp_avg = kp;
    z_avg = zl;
  } else {
    kp_avg = momentum * state[0] + (1.f - momentum) * kp;
    z_avg = momentum * state[1] + (1.f - momentum) * zl;
  }
  const float target = (kp + 1e-8f) / (zl + 1e-8f);
  float w = (1.f - rate) * state[3] + rate * target;
  w = fminf(fmaxf(w, 1e-3f), 10.0f);
  state[0] = kp_avg; state[1] = z_avg; state[2] = 1.f; state[3] = w;
  out[0] = kp / (kp_avg + 1e-8f) + zl / (z_avg + 1e-8f);
  out[1] = kp;
  out[2] = zl;
  scales[0] = float(2.0 / (numel_hm * double(kp_avg + 1e-8f)));
  scales[1] = float(1.0 / (numel_z * double(z_avg + 1e-8f)));
}

__global__ void __launch_bounds__(256) pose_loss_grad_kernel(const float4* __restrict__ hm, const float4* __restrict__ thm,
                                                             const float* __restrict__ kps, int kp_stride,
                                                             const float* __restrict__ z, const float* __restrict__ tz,
                                                             const float* __restrict__ scales, float4* __restrict__ dhm,
                                                             float* __restrict__ dz, long long n4, int map4, int BK) {
  pdl_grid_sync();
  const float s0 = scales[0], s1 = scales[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int bk = int(i / map4);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (__ldg(kps + (long long)bk * kp_stride + 2) > 1.f) {
      const float4 a = __ldg(hm + i), b = __ldg(thm + i);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      g.x = s0 * __expf(-d0 * d0) * d0;
      g.y = s0 * __expf(-d1 * d1) * d1;
      g.z = s0 * __expf(-d2 * d2) * d2;
      g.w = s0 * __expf(-d3 * d3) * d3;
    }
    dhm[i] = g;
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < BK; i += blockDim.x) {
      const float m = __ldg(kps + (long long)i * kp_stride + 2) > 1.f ? 1.f : 0.f;
      const float d = z[i] * m - tz[i] * m;
      dz[i] = s1 * m * ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
    }
  }
}

// torch.optim.AdamW semantics: p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).   step_dev holds t-1 on entry; thread 0 bumps it.
__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                    float4* __restrict__ v, long long n4, float lr, float b1, float b2,
                                                    float eps, float wd, float grad_scale, const long long* __restrict__ step_dev,
                                                    const float* __restrict__ hyper) {
  pdl_grid_sync();
  // hyper = device {lr, weight_decay}: a scheduler (train.py:286-293, ReduceLROnPlateau) changes the rate between
  // replays of the captured step without re-recording it
  if (hyper != nullptr) { lr = hyper[0]; wd = hyper[1]; }
  const float t = float(*step_dev + 1);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
    float* pf = &pp.x; float* mf = &mm.x; float* vf = &vv.x; const float* gf = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gf[k] * grad_scale;
      mf[k] = b1 * mf[k] + (1.f - b1) * gk;
      vf[k] = b2 * vf[k] + (1.f - b2) * gk * gk;
      const float denom = sqrtf(vf[k]) * inv_sqrt_bc2 + eps;
      pf[k] = pf[k] * decay - step_size * (mf[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}
__global__ void bump_step_kernel(long long* step_dev) {
  pdl_grid_sync();
  *step_dev += 1;
}

}  // namespace

cudaError_t launch_pose_loss(const float* hm, const float* thm, const float* kps, int kp_stride, const float* z,
                             const float* tz, double* sums, float* state, float* out, float* scales, float* dhm, float* dz,
                             int B, int K, int HW, float momentum, float rate, int sms, cudaStream_t s) {
  const long long n4 = (long long)B * K * HW / 4;
  const int map4 = HW / 4, BK = B * K;
  int grid = int((n4 + 255) / 256);
  if (grid > sms * 8) grid = sms * 8;
  if (grid < 1) grid = 1;
  launch_k<pose_loss_reduce_kernel>(grid, 256, 0, s, reinterpret_cast<const float4*>(hm), reinterpret_cast<const float4*>(thm), kps,
                                               kp_stride, z, tz, sums, n4, map4, BK);
  launch_k<pose_loss_finalize_kernel>(1, 32, 0, s, sums, state, out, scales, double(B) * K * HW, double(BK), momentum, rate);
  launch_k<pose_loss_grad_kernel>(grid, 256, 0, s, reinterpret_cast<const float4*>(hm), reinterpret_cast<const float4*>(thm), kps,
                                             kp_stride, z, tz, scales, reinterpret_cast<float4*>(dhm), dz, n4, map4, BK);
  return cudaGetLastError();
}

cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                         float wd, float grad_scale, long long* step_dev, const float* hyper, int bump_step, int sms,
                         cudaStream_t s) {
  const long long n4 = n / 4;
  int grid = int((n4 + 255) / 256);
  if (grid > sms * 8) grid = sms * 8;
  if (grid < 1) grid = 1;
  launch_k<adamw_kernel>(grid, 256, 0, s, reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
                                    reinterpret_cast<float4*>(v), n4, lr, b1, b2, eps, wd, grad_scale, step_dev, hyper);
  if (bump_step) launch_k<bump_step_kernel>(1, 1, 0, s, step_dev);
  return cudaGetLastError();
}

}  // namespace dp

// ------------------------------------------------------------------------------------------------
// Weight re-packing after an optimizer step: every trainable conv / transposed-conv weight (fp32 [d0, d1, kh, kw])
// is copied into the bf16 GEMM layouts the forward and input-gradient kernels read (permuted, taps optionally
// mirrored).  One launch for all layers: blockIdx.y selects a job from a device-resident table
//   job[16] = { src ptr, dst ptr, n0..n3 (dst-order extents), s0..s3 (signed src element strides),
//               t0..t3 (dst element strides), src offset, total }
// and BatchNorm's num_batches_tracked counters (14 separate int64 buffers) are bumped by one launch too.
namespace dp {
namespace {
constexpr int kPackChunk = 2048;   // consecutive destination elements per block
__global__ void __launch_bounds__(256) pack_weights_kernel(const long long* __restrict__ jobs) {
  pdl_grid_sync();
  const long long* j = jobs + (long long)blockIdx.y * 16;
  const unsigned total = unsigned(j[15]);
  const unsigned first = blockIdx.x * unsigned(kPackChunk);
  if (first >= total) return;   // the grid is sized for the largest job
  const float* src = reinterpret_cast<const float*>(j[0]) + j[14];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j[1]);
  // every tensor has < 2^31 elements (checked by the launcher): 32-bit index arithmetic, signed source strides
  const unsigned n1 = unsigned(j[3]), n2 = unsigned(j[4]), n3 = unsigned(j[5]);
  const int s0 = int(j[6]), s1 = int(j[7]), s2 = int(j[8]), s3 = int(j[9]);
  const unsigned t0 = unsigned(j[10]), t1 = unsigned(j[11]), t2 = unsigned(j[12]), t3 = unsigned(j[13]);
  const unsigned last = min(total, first + unsigned(kPackChunk));
  // a block covers a contiguous destination range: its strided source reads stay inside a few KB that L1 keeps
  for (unsigned i = first + threadIdx.x; i < last; i += 256) {
    const unsigned c3 = i % n3;
    unsigned r = i / n3;
    const unsigned c2 = r % n2;
    r /= n2;
    const unsigned c1 = r % n1;
    const unsigned c0 = r / n1;
    const long long so = (long long)int(c0) * s0 + (long long)int(c1) * s1 + (long long)int(c2) * s2 + (long long)int(c3) * s3;
    dst[c0 * t0 + c1 * t1 + c2 * t2 + c3 * t3] = __float2bfloat16_rn(__ldg(src + so));
  }
}
__global__ void add_i64_kernel(const long long* __restrict__ ptrs, int n, long long inc) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) *reinterpret_cast<long long*>(ptrs[i]) += inc;
}
}  // namespace

cudaError_t launch_pack_weights(const long long* jobs_dev, int njobs, long long max_total, int sms, cudaStream_t s) {
  if (max_total >= (1LL << 31)) return cudaErrorInvalidValue;
  long long gx = (max_total + kPackChunk - 1) / kPackChunk;   // blocks beyond a job's size return immediately
  if (gx < 1) gx = 1;
  (void)sms;
  launch_k<pack_weights_kernel>(dim3(unsigned(gx), unsigned(njobs)), 256, 0, s, jobs_dev);
  return cudaGetLastError();
}
cudaError_t launch_add_i64(const long long* ptrs_dev, int n, long long inc, cudaStream_t s) {
  launch_k<add_i64_kernel>((n + 127) / 128, 128, 0, s, ptrs_dev, n, inc);
  return cudaGetLastError();
}
}  // namespace dp
