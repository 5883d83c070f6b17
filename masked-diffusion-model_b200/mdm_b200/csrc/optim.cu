// Fused optimiser tail: global-norm clip + Adam/AdamW + EMA + bf16 mirror,
// one pass over the flat parameter buffer.  Replaces `clip_grad_norm_(params, 1.0)`,
// `optimizer.step()` and `ema_model.step(params)` (trainer_masked.py:144-153,
// main_train_masked.py:116-141) -- >= 4 passes over 113.67 M parameters in the reference.
#include <math.h>

#include "common.cuh"

namespace mdm {

__global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  MDM_PDL_ENTER();
  __shared__ float red[32];
  float s = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) s += g[i] * g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  MDM_PDL_ENTER();
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) *out = (float)r;
  }
}

// step statistics -> host-mapped pinned memory, straight from the device: the trainer reads the loss of a step as soon
// as the FORWARD part of the captured step has run (the backward and the optimiser are still in flight), so the host
// prepares and enqueues the next step while the GPU finishes this one instead of idling through a full
// device synchronisation per batch (the reference's `loss.item()`, trainer_masked.py:160).
__global__ void publish_stats_kernel(const float* __restrict__ src, int n, int* __restrict__ counter, volatile float* dst_host) {
  MDM_PDL_ENTER();
  if (threadIdx.x < n) dst_host[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(counter, 1) + 1;
    dst_host[n] = __int_as_float(t);        // the tick the host polls for: written after the values are visible
    __threadfence_system();
  }
}

// scalars exactly as torch.optim derives them (python doubles rounded to fp32 once): 1 - beta from the DOUBLE beta
// (1.0f - 0.999f is 4.7e-5 off 0.001f), step_size = lr / bias_correction1, sqrt(bias_correction2)
struct AdamArgs {
  float lr, beta1, beta2, omb1, omb2, eps, weight_decay, step_size, sqrt_c2, max_norm, ema_decay, grad_scale;
  int decoupled;   // 1 = AdamW, 0 = Adam (L2 added to the gradient)
};

__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, float* __restrict__ ema, __nv_bfloat16* __restrict__ p16,
                                long long n, const float* __restrict__ gnorm_sq, const float* __restrict__ hyper, AdamArgs a) {
  MDM_PDL_ENTER();
  if (hyper) {  // per-step scalars read from device memory so that a captured CUDA graph can be replayed
    a.lr = hyper[0]; a.step_size = hyper[1]; a.sqrt_c2 = hyper[2]; a.ema_decay = hyper[3];
  }
  float coef = a.grad_scale;
  if (gnorm_sq && a.max_norm > 0.f) {
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = sqrtf(*gnorm_sq) * a.grad_scale;
    coef *= fminf(1.0f, a.max_norm / (total + 1e-6f));
  }
  const float step = a.step_size;
  const float sqrt_c2 = a.sqrt_c2;          // torch: denom = sqrt(v) / sqrt(bias_correction2) + eps
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    float4 P = *reinterpret_cast<const float4*>(p + i);
    const float4 Gr = *reinterpret_cast<const float4*>(g + i);
    float4 M = *reinterpret_cast<const float4*>(m + i);
    float4 V = *reinterpret_cast<const float4*>(v + i);
    float pp[4] = {P.x, P.y, P.z, P.w}, gg[4] = {Gr.x, Gr.y, Gr.z, Gr.w}, mm[4] = {M.x, M.y, M.z, M.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gg[k] * coef;
      if (a.decoupled) pp[k] *= (1.0f - a.lr * a.weight_decay);
      else gk += a.weight_decay * pp[k];
      mm[k] = a.beta1 * mm[k] + a.omb1 * gk;
      vv[k] = a.beta2 * vv[k] + a.omb2 * gk * gk;
      const float denom = sqrtf(vv[k]) / sqrt_c2 + a.eps;
      pp[k] -= step * (mm[k] / denom);
    }
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (ema) {
      float4 E = *reinterpret_cast<const float4*>(ema + i);
      const float om = 1.0f - a.ema_decay;
      E.x -= om * (E.x - pp[0]); E.y -= om * (E.y - pp[1]); E.z -= om * (E.z - pp[2]); E.w -= om * (E.w - pp[3]);
      *reinterpret_cast<float4*>(ema + i) = E;
    }
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p16 + i) = u;
    }
  }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ ema,
                           __nv_bfloat16* __restrict__ p16, long long n, const float* __restrict__ gnorm_sq,
                           const float* __restrict__ hyper, AdamArgs a) {
  MDM_PDL_ENTER();
  if (hyper) { a.lr = hyper[0]; a.ema_decay = hyper[3]; }
  float coef = a.grad_scale;
  if (gnorm_sq && a.max_norm > 0.f) coef *= fminf(1.0f, a.max_norm / (sqrtf(*gnorm_sq) * a.grad_scale + 1e-6f));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pp = p[i] - a.lr * (g[i] * coef);
    p[i] = pp;
    if (ema) ema[i] -= (1.0f - a.ema_decay) * (ema[i] - pp);
    if (p16) p16[i] = __float2bfloat16(pp);
  }
}

}  // namespace mdm

using namespace mdm;

extern "C" {

int mdm_grad_sumsq(const float* g, int64_t n, float* ws, float* out, void* stream) {
  MDM_CHECK_ARG(g && ws && out && n > 0, "grad_sumsq: bad arguments");
  const int blocks = 1024;
  launch_pdl(sumsq_partial_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), g, n, ws);
  MDM_LAUNCH_CHECK();
  launch_pdl(sumsq_final_kernel, dim3(1), dim3(1024), 0, as_stream(stream), ws, blocks, out);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_publish_stats(const float* src, int n, int* counter, float* dst_host_mapped, void* stream) {
  MDM_CHECK_ARG(src && counter && dst_host_mapped && n >= 1 && n <= 31, "publish_stats: bad arguments");
  launch_pdl(publish_stats_kernel, dim3(1), dim3(32), 0, as_stream(stream), src, n, counter, (volatile float*)dst_host_mapped);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

static int adam_launch(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, int64_t n,
                       const float* gnorm_sq, const float* hyper, float lr, double beta1, double beta2, float eps,
                       float weight_decay, double bias_c1, double bias_c2, float max_norm, float ema_decay,
                       float grad_scale, int mode, void* stream) {
  MDM_CHECK_ARG(p && g && n > 0 && n % 4 == 0, "adam_ema_step: n must be a positive multiple of 4");
  MDM_CHECK_ARG(mode >= 0 && mode <= 2, "adam_ema_step: mode 0 = Adam, 1 = AdamW, 2 = SGD");
  AdamArgs a{lr, (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), eps, weight_decay,
             (float)((double)lr / bias_c1), (float)sqrt(bias_c2), max_norm, ema_decay, grad_scale, mode == 1};
  const int blocks = kNumSMs * 8;
  if (mode == 2) {
    launch_pdl(sgd_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, ema, (__nv_bfloat16*)p_bf16, n, gnorm_sq, hyper, a);
  } else {
    MDM_CHECK_ARG(m && v, "adam_ema_step: moment buffers are NULL");
    launch_pdl(adam_ema_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, m, v, ema, (__nv_bfloat16*)p_bf16, n, gnorm_sq, hyper, a);
  }
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, int64_t n,
                      const float* gnorm_sq, float lr, double beta1, double beta2, float eps, float weight_decay,
                      double bias_c1, double bias_c2, float max_norm, float ema_decay, float grad_scale, int mode,
                      void* stream) {
  return adam_launch(p, g, m, v, ema, p_bf16, n, gnorm_sq, nullptr, lr, beta1, beta2, eps, weight_decay, bias_c1,
                     bias_c2, max_norm, ema_decay, grad_scale, mode, stream);
}

int mdm_adam_ema_step_dev(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, int64_t n,
                          const float* gnorm_sq, const float* hyper, double beta1, double beta2, float eps,
                          float weight_decay, float max_norm, float grad_scale, int mode, void* stream) {
  MDM_CHECK_ARG(hyper, "adam_ema_step_dev: hyper is NULL");
  return adam_launch(p, g, m, v, ema, p_bf16, n, gnorm_sq, hyper, 0.f, beta1, beta2, eps, weight_decay, 1.0, 1.0,
                     max_norm, 0.f, grad_scale, mode, stream);
}

}  // extern "C"
