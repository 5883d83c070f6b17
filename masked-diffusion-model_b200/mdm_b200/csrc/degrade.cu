// K1 (fill + composite) and K5 (restoration-loop update) -- HBM-bound elementwise kernels.
//
// Reference: scheduler.py:298-321 (degrade_training), :450-475 (degrade_independent_base_
// sampling), :572-598 (degrade_with_mask); sampler.py:143-152,167-216 (loop body).
//
// Layout: images NCHW, one (batch, channel) plane = HW contiguous elements; masks are one byte
// per pixel ([batch, mask_ch, HW], mask_ch = 1 or C), written by rng.cu.  Each plane is cut into
// chunks of CHUNK elements; one CTA owns one chunk, every thread moves 16-byte vectors.
//
// The fill value is a per-sample (or per-channel) masked mean, i.e. a reduction over the
// sample that must finish before the composite can start: pass 1 writes per-chunk partial sums
// (no atomics -> deterministic), pass 2 re-reduces the <= C*nchunk partials of its sample in a
// fixed order (identical in every CTA of that sample) and does the composite.  The second read
// of the image hits L2 for the shapes on this path (<= 50 MB per tensor, 126 MB L2).
//
// Arithmetic is the reference's, op for op and unfused (__fmul_rn/__fadd_rn), so that given the
// same fill the composite is bit-identical, including the 0*NaN = NaN behaviour when a sample
// has no degraded pixel.
#include "common.cuh"

namespace mdm {

constexpr int DG_THREADS = 256;
constexpr int DG_VEC_PER_THREAD = 4;                              // K1: 4 x float4 per thread and tensor in flight
constexpr int K5_VEC_PER_THREAD = 2;                              // K5 reads five tensors: 2 x float4 each keeps 4 CTAs / SM
constexpr int DG_CHUNK = 8192;                                    // elements of a plane per CTA

// sum of N values over the CTA with ONE shared-memory round (every thread gets the results)
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float* red /* >= N * 32 floats */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[k * 32 + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(lane < nw ? red[k * 32 + lane] : 0.f);
}

struct Shift {
  const float* p;
  int64_t sb, sc, sp;
  __device__ __forceinline__ float at(int b, int c, int i) const {
    return p ? p[b * sb + c * sc + (int64_t)i * sp] : 0.0f;
  }
  // four consecutive pixels (vector path: sp is 0 (broadcast) or 1 with 16-byte aligned rows, checked on the host)
  __device__ __forceinline__ float4 at4(int b, int c, int i) const {
    if (!p) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (sp == 0) { const float v = p[b * sb + c * sc]; return make_float4(v, v, v, v); }
    return *reinterpret_cast<const float4*>(p + b * sb + c * sc + i);
  }
};
static bool shift_vec_ok(const float* p, int64_t sb, int64_t sc, int64_t sp) {
  return p == nullptr || sp == 0 || (sp == 1 && (sb & 3) == 0 && (sc & 3) == 0 && ((uintptr_t)p & 15) == 0);
}
// uint8 images (the GPU data-feeding path): the reference's `ToTensor()` + `Normalize(0.5, 0.5)` (utils/mydataset.py:81)
// op for op -- u / 255, then (x - 0.5) / 0.5 -- so a uint8 batch gives the bits the CPU transforms give
__device__ __forceinline__ float u8_normalise(uint8_t u) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), 0.5f), 0.5f);
}
__device__ __forceinline__ float ld_as_float(const uint8_t* p, int64_t i) { return u8_normalise(p[i]); }

// four consecutive pixels of a plane as floats
template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static constexpr int kAlign = 16;
  static __device__ __forceinline__ float4 at(const float* x, int i) { return *reinterpret_cast<const float4*>(x + i); }
};
template <> struct Ld4<__nv_bfloat16> {
  static constexpr int kAlign = 8;
  static __device__ __forceinline__ float4 at(const __nv_bfloat16* x, int i) {
    const uint2 u = *reinterpret_cast<const uint2*>(x + i);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};
template <> struct Ld4<uint8_t> {
  static constexpr int kAlign = 4;
  static __device__ __forceinline__ float4 at(const uint8_t* x, int i) {
    const uchar4 k = *reinterpret_cast<const uchar4*>(x + i);
    return make_float4(u8_normalise(k.x), u8_normalise(k.y), u8_normalise(k.z), u8_normalise(k.w));
  }
};

__device__ __forceinline__ float4 mask4(const uint8_t* m, int64_t i) {
  const uchar4 k = *reinterpret_cast<const uchar4*>(m + i);
  return make_float4((float)k.x, (float)k.y, (float)k.z, (float)k.w);
}

// partials layout per sample: [C][nchunk][3] = {sum img*(1-m), sum img*m, sum (1-m)}
__device__ __forceinline__ float fill_value(const float* part, int c, int C, int nchunk, int fill_mode,
                                            float fill_const, int mean_area) {
  if (fill_mode == MDM_FILL_CONST) return fill_const;
  if (fill_mode == MDM_FILL_DEGRADED_AREA) {
    float s = 0.f, n = 0.f;
    const int c0 = (mean_area == MDM_AREA_IMAGE) ? 0 : c;
    const int c1 = (mean_area == MDM_AREA_IMAGE) ? C : c + 1;
    for (int cc = c0; cc < c1; ++cc)
      for (int k = 0; k < nchunk; ++k) {
        s += part[(cc * nchunk + k) * 3 + 0];
        n += part[(cc * nchunk + k) * 3 + 2];
      }
    return __fdiv_rn(s, n);  // 0/0 -> NaN when nothing is degraded, as in the reference
  }
  float s = 0.f, n = 0.f;  // MDM_FILL_NON_DEGRADED: always per channel (dims (2,3))
  for (int k = 0; k < nchunk; ++k) {
    s += part[(c * nchunk + k) * 3 + 1];
    n += part[(c * nchunk + k) * 3 + 2];
  }
  float v = __fmul_rn(__fdiv_rn(s, n), -1.0f);
  return isnan(v) ? 0.0f : v;
}

__device__ __forceinline__ float composite(float m, float fill, float x) {
  // ((1-masks) * mean_pixel) + masks * img
  return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, m), fill), __fmul_rn(m, x));
}

// ---- K1 pass 1 -----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(DG_THREADS) degrade_stats_kernel(const T* __restrict__ img,
                                                                   const uint8_t* __restrict__ mask,
                                                                   int mask_ch, float* __restrict__ ws,
                                                                   int C, int hw, int nchunk) {
  MDM_PDL_ENTER();
  __shared__ float red[3 * 32];
  const int chunk = blockIdx.x, plane = blockIdx.y;  // plane = b*C + c
  const int b = plane / C, c = plane % C;
  const T* x = img + (int64_t)plane * hw;
  const uint8_t* m = mask + ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw;
  float s0 = 0.f, s1 = 0.f, n0 = 0.f;
  const int base = chunk * DG_CHUNK;
  const int end = min(base + DG_CHUNK, hw);
  if ((hw & 3) == 0 && ((uintptr_t)img & (Ld4<T>::kAlign - 1)) == 0 && ((uintptr_t)mask & 3) == 0) {
    // all of this thread's 16-byte loads are issued before the first use (4 image + 4 mask vectors in flight)
    for (int g0 = base; g0 < end; g0 += DG_THREADS * 4 * DG_VEC_PER_THREAD) {
      float4 v[DG_VEC_PER_THREAD], k[DG_VEC_PER_THREAD];
#pragma unroll
      for (int u = 0; u < DG_VEC_PER_THREAD; ++u) {
        const int i = g0 + (u * DG_THREADS + threadIdx.x) * 4;
        const bool ok = i < end;
        v[u] = ok ? Ld4<T>::at(x, i) : make_float4(0.f, 0.f, 0.f, 0.f);
        k[u] = ok ? mask4(m, i) : make_float4(1.f, 1.f, 1.f, 1.f);      // mask 1, value 0: contributes nothing
      }
#pragma unroll
      for (int u = 0; u < DG_VEC_PER_THREAD; ++u) {
        const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, kk[4] = {k[u].x, k[u].y, k[u].z, k[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s0 += __fmul_rn(vv[e], __fsub_rn(1.0f, kk[e]));
          s1 += __fmul_rn(vv[e], kk[e]);
          n0 += __fsub_rn(1.0f, kk[e]);
        }
      }
    }
  } else
  for (int i = base + threadIdx.x; i < end; i += DG_THREADS) {
    const float v = ld_as_float(x, i);
    const float mk = (float)m[i];
    s0 += __fmul_rn(v, __fsub_rn(1.0f, mk));
    s1 += __fmul_rn(v, mk);
    n0 += __fsub_rn(1.0f, mk);
  }
  float r3[3] = {s0, s1, n0};
  block_sum_n(r3, red);
  if (threadIdx.x == 0) {
    float* o = ws + ((int64_t)plane * nchunk + chunk) * 3;
    o[0] = r3[0]; o[1] = r3[1]; o[2] = r3[2];
  }
}

// ---- K1 pass 2 -----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(DG_THREADS) degrade_apply_kernel(
    const T* __restrict__ img, const uint8_t* __restrict__ mask, int mask_ch, int fill_mode,
    float fill_const, int mean_area, const float* __restrict__ ws, float* __restrict__ x_t,
    float* __restrict__ mask_f32, float* __restrict__ degrade_mask, float* __restrict__ fill_out,
    float* __restrict__ x0_out, int C, int hw, int nchunk) {
  MDM_PDL_ENTER();
  const int chunk = blockIdx.x, plane = blockIdx.y;
  const int b = plane / C, c = plane % C;
  const float fill = fill_value(ws + (int64_t)b * C * nchunk * 3, c, C, nchunk, fill_mode, fill_const, mean_area);
  float* x0o = x0_out ? x0_out + (int64_t)plane * hw : nullptr;     // the image as fp32 (uint8 feed: the loss reads it)
  if (fill_out && chunk == 0 && threadIdx.x == 0) fill_out[plane] = fill;
  const T* x = img + (int64_t)plane * hw;
  const uint8_t* m = mask + ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw;
  float* o = x_t + (int64_t)plane * hw;
  float* dm = degrade_mask ? degrade_mask + (int64_t)plane * hw : nullptr;
  float* mf = (mask_f32 && (mask_ch != 1 || c == 0)) ? mask_f32 + ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw : nullptr;
  const int base = chunk * DG_CHUNK;
  const int end = min(base + DG_CHUNK, hw);
  if ((hw & 3) == 0 && ((uintptr_t)img & (Ld4<T>::kAlign - 1)) == 0 && ((uintptr_t)mask & 3) == 0) {
   for (int g0 = base; g0 < end; g0 += DG_THREADS * 4 * DG_VEC_PER_THREAD) {
    float4 vv[DG_VEC_PER_THREAD], kk[DG_VEC_PER_THREAD];
#pragma unroll
    for (int u = 0; u < DG_VEC_PER_THREAD; ++u) {
      const int i = g0 + (u * DG_THREADS + threadIdx.x) * 4;
      if (i < end) {
        vv[u] = Ld4<T>::at(x, i);
        kk[u] = mask4(m, i);
      }
    }
#pragma unroll
    for (int u = 0; u < DG_VEC_PER_THREAD; ++u) {
      const int i = g0 + (u * DG_THREADS + threadIdx.x) * 4;
      if (i >= end) break;
      const float4 v = vv[u];
      const float m0 = kk[u].x, m1 = kk[u].y, m2 = kk[u].z, m3 = kk[u].w;
      float4 r;
      r.x = composite(m0, fill, v.x); r.y = composite(m1, fill, v.y);
      r.z = composite(m2, fill, v.z); r.w = composite(m3, fill, v.w);
      *reinterpret_cast<float4*>(o + i) = r;
      if (x0o) *reinterpret_cast<float4*>(x0o + i) = v;
      if (dm) {
        float4 d;
        d.x = composite(m0, fill, 1.0f); d.y = composite(m1, fill, 1.0f);
        d.z = composite(m2, fill, 1.0f); d.w = composite(m3, fill, 1.0f);
        *reinterpret_cast<float4*>(dm + i) = d;
      }
      if (mf) *reinterpret_cast<float4*>(mf + i) = make_float4(m0, m1, m2, m3);
    }
   }
  } else {
    for (int i = base + threadIdx.x; i < end; i += DG_THREADS) {
      const float v = ld_as_float(x, i);
      const float mk = (float)m[i];
      o[i] = composite(mk, fill, v);
      if (x0o) x0o[i] = v;
      if (dm) dm[i] = composite(mk, fill, 1.0f);
      if (mf) mf[i] = mk;
    }
  }
}

// ---- K5 pass 1: x0_hat and its masked sums under both masks -----------------------------------
__device__ __forceinline__ float x0_hat(float xt, float net, float sh) {
  // shifted = x_t + shift; shifted_0 = shifted + net; sample_0 = shifted_0 - shift  (sampler.py:143-152)
  return __fsub_rn(__fadd_rn(__fadd_rn(xt, sh), net), sh);
}

__global__ void __launch_bounds__(DG_THREADS) sampler_stats_kernel(
    const float* __restrict__ x_t, const float* __restrict__ net, Shift shift,
    const uint8_t* __restrict__ mask_t, const uint8_t* __restrict__ mask_n, int mask_ch,
    float* __restrict__ ws_t, float* __restrict__ ws_n, int C, int hw, int nchunk, int vec) {
  MDM_PDL_ENTER();
  __shared__ float red[6 * 32];
  const int chunk = blockIdx.x, plane = blockIdx.y;
  const int b = plane / C, c = plane % C;
  const int64_t po = (int64_t)plane * hw;
  const int64_t mo = ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw;
  float a0 = 0.f, a1 = 0.f, an = 0.f, b0 = 0.f, b1 = 0.f, bn = 0.f;
  const int base = chunk * DG_CHUNK;
  const int end = min(base + DG_CHUNK, hw);
  if (vec) {
   for (int g0 = base; g0 < end; g0 += DG_THREADS * 4 * K5_VEC_PER_THREAD) {
    float4 xv[K5_VEC_PER_THREAD], nv[K5_VEC_PER_THREAD], sv[K5_VEC_PER_THREAD], mt4[K5_VEC_PER_THREAD], mn4[K5_VEC_PER_THREAD];
#pragma unroll
    for (int u = 0; u < K5_VEC_PER_THREAD; ++u) {
      const int i = g0 + (u * DG_THREADS + threadIdx.x) * 4;
      if (i < end) {
        xv[u] = *reinterpret_cast<const float4*>(x_t + po + i);
        nv[u] = *reinterpret_cast<const float4*>(net + po + i);
        sv[u] = shift.at4(b, c, i);
        mt4[u] = mask4(mask_t + mo, i);
        mn4[u] = mask4(mask_n + mo, i);
      }
    }
#pragma unroll
    for (int u = 0; u < K5_VEC_PER_THREAD; ++u) {
      const int i = g0 + (u * DG_THREADS + threadIdx.x) * 4;
      if (i >= end) break;
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ns[4] = {nv[u].x, nv[u].y, nv[u].z, nv[u].w};
      const float ss[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
      const float ts[4] = {mt4[u].x, mt4[u].y, mt4[u].z, mt4[u].w}, ms[4] = {mn4[u].x, mn4[u].y, mn4[u].z, mn4[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = x0_hat(xs[e], ns[e], ss[e]);
        a0 += __fmul_rn(v, __fsub_rn(1.0f, ts[e])); a1 += __fmul_rn(v, ts[e]); an += __fsub_rn(1.0f, ts[e]);
        b0 += __fmul_rn(v, __fsub_rn(1.0f, ms[e])); b1 += __fmul_rn(v, ms[e]); bn += __fsub_rn(1.0f, ms[e]);
      }
    }
   }
  } else
  for (int i = base + threadIdx.x; i < end; i += DG_THREADS) {
    const float v = x0_hat(x_t[po + i], net[po + i], shift.at(b, c, i));
    const float mt = (float)mask_t[mo + i], mn = (float)mask_n[mo + i];
    a0 += __fmul_rn(v, __fsub_rn(1.0f, mt)); a1 += __fmul_rn(v, mt); an += __fsub_rn(1.0f, mt);
    b0 += __fmul_rn(v, __fsub_rn(1.0f, mn)); b1 += __fmul_rn(v, mn); bn += __fsub_rn(1.0f, mn);
  }
  float r6[6] = {a0, a1, an, b0, b1, bn};
  block_sum_n(r6, red);
  if (threadIdx.x == 0) {
    float* o = ws_t + ((int64_t)plane * nchunk + chunk) * 3;
    o[0] = r6[0]; o[1] = r6[1]; o[2] = r6[2];
    o = ws_n + ((int64_t)plane * nchunk + chunk) * 3;
    o[0] = r6[3]; o[1] = r6[4]; o[2] = r6[5];
  }
}

// ---- K5 pass 2 -------------------------------------------------------------------------------
__global__ void __launch_bounds__(DG_THREADS) sampler_update_kernel(
    const float* __restrict__ x_t, const float* __restrict__ net, Shift shift,
    const uint8_t* __restrict__ mask_t, const uint8_t* __restrict__ mask_n, int mask_ch, int fill_mode,
    float fill_const, int mean_area, int momentum, int update, Shift shift_next,
    const float* __restrict__ ws_t, const float* __restrict__ ws_n, float* __restrict__ x_next,
    float* __restrict__ x_in_next, float* __restrict__ s0_out, int C, int hw, int nchunk, int vec) {
  MDM_PDL_ENTER();
  const int chunk = blockIdx.x, plane = blockIdx.y;
  const int b = plane / C, c = plane % C;
  const int64_t po = (int64_t)plane * hw;
  const int64_t mo = ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw;
  const int64_t wo = (int64_t)b * C * nchunk * 3;
  const float f_t = update ? fill_value(ws_t + wo, c, C, nchunk, fill_mode, fill_const, mean_area) : 0.f;
  const float f_n = update ? fill_value(ws_n + wo, c, C, nchunk, fill_mode, fill_const, mean_area) : 0.f;
  const int base = chunk * DG_CHUNK;
  const int end = min(base + DG_CHUNK, hw);
  if (vec) {
   for (int g0 = base; g0 < end; g0 += DG_THREADS * 4 * K5_VEC_PER_THREAD) {
    float4 xv[K5_VEC_PER_THREAD], nv[K5_VEC_PER_THREAD], sv[K5_VEC_PER_THREAD], mt4[K5_VEC_PER_THREAD], mn4[K5_VEC_PER_THREAD],
        sn4[K5_VEC_PER_THREAD];
#pragma unroll
    for (int u = 0; u < K5_VEC_PER_THREAD; ++u) {
      const int i = g0 + (u * DG_THREADS + threadIdx.x) * 4;
      if (i < end) {
        xv[u] = *reinterpret_cast<const float4*>(x_t + po + i);
        nv[u] = *reinterpret_cast<const float4*>(net + po + i);
        sv[u] = shift.at4(b, c, i);
        if (update) { mt4[u] = mask4(mask_t + mo, i); mn4[u] = mask4(mask_n + mo, i); }
        if (x_in_next) sn4[u] = shift_next.at4(b, c, i);
      }
    }
#pragma unroll
    for (int u = 0; u < K5_VEC_PER_THREAD; ++u) {
      const int i = g0 + (u * DG_THREADS + threadIdx.x) * 4;
      if (i >= end) break;
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ns[4] = {nv[u].x, nv[u].y, nv[u].z, nv[u].w};
      const float ss[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
      const float ts[4] = {mt4[u].x, mt4[u].y, mt4[u].z, mt4[u].w}, ms[4] = {mn4[u].x, mn4[u].y, mn4[u].z, mn4[u].w};
      const float sn[4] = {sn4[u].x, sn4[u].y, sn4[u].z, sn4[u].w};
      float v0[4], xn[4], xi[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v0[e] = x0_hat(xs[e], ns[e], ss[e]);
        xn[e] = xs[e];
        if (update) {
          const float d_t = composite(ts[e], f_t, v0[e]);
          const float d_n = composite(ms[e], f_n, v0[e]);
          xn[e] = momentum ? __fadd_rn(xs[e], __fsub_rn(d_n, d_t)) : d_n;
        }
        xi[e] = x_in_next ? __fadd_rn(xn[e], sn[e]) : 0.f;
      }
      if (s0_out) *reinterpret_cast<float4*>(s0_out + po + i) = make_float4(v0[0], v0[1], v0[2], v0[3]);
      if (x_next) *reinterpret_cast<float4*>(x_next + po + i) = make_float4(xn[0], xn[1], xn[2], xn[3]);
      if (x_in_next) *reinterpret_cast<float4*>(x_in_next + po + i) = make_float4(xi[0], xi[1], xi[2], xi[3]);
    }
   }
    return;
  }
  for (int i = base + threadIdx.x; i < end; i += DG_THREADS) {
    const float xt = x_t[po + i];
    const float v = x0_hat(xt, net[po + i], shift.at(b, c, i));
    if (s0_out) s0_out[po + i] = v;
    float xn = xt;
    if (update) {
      const float d_t = composite((float)mask_t[mo + i], f_t, v);
      const float d_n = composite((float)mask_n[mo + i], f_n, v);
      xn = momentum ? __fadd_rn(xt, __fsub_rn(d_n, d_t)) : d_n;  // sampler.py:206-216
    }
    if (x_next) x_next[po + i] = xn;
    if (x_in_next) x_in_next[po + i] = __fadd_rn(xn, shift_next.at(b, c, i));
  }
}

__global__ void __launch_bounds__(DG_THREADS) add_shift_kernel(const float* __restrict__ x, Shift shift,
                                                               float* __restrict__ out, int C, int hw) {
  MDM_PDL_ENTER();
  const int plane = blockIdx.y, b = plane / C, c = plane % C;
  const int64_t po = (int64_t)plane * hw;
  const int base = blockIdx.x * DG_CHUNK;
  const int end = min(base + DG_CHUNK, hw);
  for (int i = base + threadIdx.x; i < end; i += DG_THREADS) out[po + i] = __fadd_rn(x[po + i], shift.at(b, c, i));
}

// =================================================================================================================
// Single-pass forms of K1 and K5 (round 2): the per-sample masked mean is a reduction over the sample that must finish
// before the composite can start -- the two-pass kernels above re-read the image for it (K1 49 %, K5 35 % of the HBM
// peak at 256x3x128x128, where the tensors exceed L2).  Here a thread-block CLUSTER owns one sample: CTA (c, j) of the
// cluster holds part j of channel plane c IN REGISTERS (VPT 16-byte vectors per thread and tensor), reduces its sums,
// publishes them in its shared memory, and after one cluster barrier every CTA gathers the C x parts partials through
// distributed shared memory in a fixed order (same fill value everywhere, deterministic), then composites from
// registers.  Every tensor crosses HBM exactly once.  Samples that do not fit a cluster of <= 8 CTAs (3x256x256) keep
// the two-pass kernels.  MEASURED SLOWER than the two-pass kernels (see fused_parts below): kept as an opt-in
// (MDM_DEGRADE_FUSED=1) with its parity test, not the default.
// =================================================================================================================
constexpr int K1_VPT = 12, K5_VPT = 8, FUSED_MAX_CLUSTER = 8, FUSED_MAXC = 4;

__device__ __forceinline__ uint32_t dg_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void dg_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float dg_ld_remote(const float* local_ptr, uint32_t rank) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local_ptr)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}
// fill value of channel c from per-channel totals tot[c][3] = {sum img*(1-m), sum img*m, sum (1-m)}
__device__ __forceinline__ float fill_from_totals(const float (*tot)[3], int c, int C, int fill_mode, float fill_const, int mean_area) {
  if (fill_mode == MDM_FILL_CONST) return fill_const;
  if (fill_mode == MDM_FILL_DEGRADED_AREA) {
    float s = 0.f, n = 0.f;
    if (mean_area == MDM_AREA_IMAGE) { for (int cc = 0; cc < C; ++cc) { s += tot[cc][0]; n += tot[cc][2]; } }
    else { s = tot[c][0]; n = tot[c][2]; }
    return __fdiv_rn(s, n);
  }
  const float v = __fmul_rn(__fdiv_rn(tot[c][1], tot[c][2]), -1.0f);
  return isnan(v) ? 0.0f : v;
}
// gathers the NS sums of every CTA of the cluster: all[q * NS + k] (q = channel * parts + part), then per-channel totals
template <int NS>
__device__ __forceinline__ void gather_cluster_sums(const float* my_part /*shared*/, float* all /*shared*/, int CS) {
  dg_cluster_sync();                                            // every CTA's partial sums are published
  if ((int)threadIdx.x < CS * NS) all[threadIdx.x] = dg_ld_remote(my_part + threadIdx.x % NS, threadIdx.x / NS);
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(DG_THREADS) degrade_fused_kernel(
    const T* __restrict__ img, const uint8_t* __restrict__ mask, int mask_ch, int fill_mode, float fill_const, int mean_area,
    float* __restrict__ x_t, float* __restrict__ mask_f32, float* __restrict__ degrade_mask, float* __restrict__ fill_out,
    float* __restrict__ x0_out, int C, int hw, int parts) {
  MDM_PDL_ENTER();
  __shared__ float red[3 * 32];
  __shared__ float part[3];
  __shared__ float all[FUSED_MAX_CLUSTER * 3];
  const int CS = C * parts;
  const int q = (int)dg_cluster_ctarank(), c = q / parts, j = q - c * parts;
  const int b = blockIdx.x / CS;
  const int plane = b * C + c;
  const T* x = img + (int64_t)plane * hw;
  const uint8_t* m = mask + ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw;
  const int cap = DG_THREADS * 4 * K1_VPT;
  const int base = j * cap, end = min(base + cap, hw);
  float4 v[K1_VPT], k[K1_VPT];
#pragma unroll
  for (int u = 0; u < K1_VPT; ++u) {
    const int i = base + (u * DG_THREADS + threadIdx.x) * 4;
    const bool ok = i < end;
    v[u] = ok ? Ld4<T>::at(x, i) : make_float4(0.f, 0.f, 0.f, 0.f);
    k[u] = ok ? mask4(m, i) : make_float4(1.f, 1.f, 1.f, 1.f);      // mask 1, value 0: contributes nothing
  }
  float r3[3] = {0.f, 0.f, 0.f};
  if (fill_mode != MDM_FILL_CONST) {
#pragma unroll
    for (int u = 0; u < K1_VPT; ++u) {
      const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, kk[4] = {k[u].x, k[u].y, k[u].z, k[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        r3[0] += __fmul_rn(vv[e], __fsub_rn(1.0f, kk[e]));
        r3[1] += __fmul_rn(vv[e], kk[e]);
        r3[2] += __fsub_rn(1.0f, kk[e]);
      }
    }
    block_sum_n(r3, red);
  }
  if (threadIdx.x < 3) part[threadIdx.x] = r3[threadIdx.x];
  gather_cluster_sums<3>(part, all, CS);
  float tot[FUSED_MAXC][3];
#pragma unroll
  for (int cc = 0; cc < FUSED_MAXC; ++cc) {
    tot[cc][0] = tot[cc][1] = tot[cc][2] = 0.f;
    if (cc < C)
      for (int jj = 0; jj < parts; ++jj) {
        tot[cc][0] += all[(cc * parts + jj) * 3];
        tot[cc][1] += all[(cc * parts + jj) * 3 + 1];
        tot[cc][2] += all[(cc * parts + jj) * 3 + 2];
      }
  }
  const float fill = fill_from_totals(tot, c, C, fill_mode, fill_const, mean_area);
  if (fill_out && j == 0 && threadIdx.x == 0) fill_out[plane] = fill;
  float* o = x_t + (int64_t)plane * hw;
  float* x0o = x0_out ? x0_out + (int64_t)plane * hw : nullptr;
  float* dm = degrade_mask ? degrade_mask + (int64_t)plane * hw : nullptr;
  float* mf = (mask_f32 && (mask_ch != 1 || c == 0)) ? mask_f32 + ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw : nullptr;
#pragma unroll
  for (int u = 0; u < K1_VPT; ++u) {
    const int i = base + (u * DG_THREADS + threadIdx.x) * 4;
    if (i < end) {
      const float4 vv = v[u], kk = k[u];
      *reinterpret_cast<float4*>(o + i) = make_float4(composite(kk.x, fill, vv.x), composite(kk.y, fill, vv.y),
                                                      composite(kk.z, fill, vv.z), composite(kk.w, fill, vv.w));
      if (x0o) *reinterpret_cast<float4*>(x0o + i) = vv;
      if (dm) *reinterpret_cast<float4*>(dm + i) = make_float4(composite(kk.x, fill, 1.0f), composite(kk.y, fill, 1.0f),
                                                                composite(kk.z, fill, 1.0f), composite(kk.w, fill, 1.0f));
      if (mf) *reinterpret_cast<float4*>(mf + i) = kk;
    }
  }
  dg_cluster_sync();      // nobody leaves while a peer may still read its partial sums
}

__global__ void __launch_bounds__(DG_THREADS) sampler_fused_kernel(
    const float* __restrict__ x_t, const float* __restrict__ net, Shift shift, const uint8_t* __restrict__ mask_t,
    const uint8_t* __restrict__ mask_n, int mask_ch, int fill_mode, float fill_const, int mean_area, int momentum, int update,
    Shift shift_next, float* __restrict__ x_next, float* __restrict__ x_in_next, float* __restrict__ s0_out, int C, int hw, int parts) {
  MDM_PDL_ENTER();
  __shared__ float red[6 * 32];
  __shared__ float part[6];
  __shared__ float all[FUSED_MAX_CLUSTER * 6];
  const int CS = C * parts;
  const int q = (int)dg_cluster_ctarank(), c = q / parts, j = q - c * parts;
  const int b = blockIdx.x / CS;
  const int64_t po = (int64_t)(b * C + c) * hw;
  const int64_t mo = ((int64_t)b * mask_ch + (mask_ch == 1 ? 0 : c)) * hw;
  const int cap = DG_THREADS * 4 * K5_VPT;
  const int base = j * cap, end = min(base + cap, hw);
  float4 xv[K5_VPT], v0[K5_VPT], mt4[K5_VPT], mn4[K5_VPT];
#pragma unroll
  for (int u = 0; u < K5_VPT; ++u) {
    const int i = base + (u * DG_THREADS + threadIdx.x) * 4;
    if (i < end) {
      xv[u] = *reinterpret_cast<const float4*>(x_t + po + i);
      const float4 nv = *reinterpret_cast<const float4*>(net + po + i);
      const float4 sv = shift.at4(b, c, i);
      v0[u] = make_float4(x0_hat(xv[u].x, nv.x, sv.x), x0_hat(xv[u].y, nv.y, sv.y), x0_hat(xv[u].z, nv.z, sv.z), x0_hat(xv[u].w, nv.w, sv.w));
      if (update) { mt4[u] = mask4(mask_t + mo, i); mn4[u] = mask4(mask_n + mo, i); }
      else { mt4[u] = mn4[u] = make_float4(1.f, 1.f, 1.f, 1.f); }
    } else {
      xv[u] = v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      mt4[u] = mn4[u] = make_float4(1.f, 1.f, 1.f, 1.f);
    }
  }
  float f_t = 0.f, f_n = 0.f;
  if (update) {
    float r6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (fill_mode != MDM_FILL_CONST) {
#pragma unroll
      for (int u = 0; u < K5_VPT; ++u) {
        const float vs[4] = {v0[u].x, v0[u].y, v0[u].z, v0[u].w};
        const float ts[4] = {mt4[u].x, mt4[u].y, mt4[u].z, mt4[u].w}, ms[4] = {mn4[u].x, mn4[u].y, mn4[u].z, mn4[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          r6[0] += __fmul_rn(vs[e], __fsub_rn(1.0f, ts[e])); r6[1] += __fmul_rn(vs[e], ts[e]); r6[2] += __fsub_rn(1.0f, ts[e]);
          r6[3] += __fmul_rn(vs[e], __fsub_rn(1.0f, ms[e])); r6[4] += __fmul_rn(vs[e], ms[e]); r6[5] += __fsub_rn(1.0f, ms[e]);
        }
      }
      block_sum_n(r6, red);
    }
    if (threadIdx.x < 6) part[threadIdx.x] = r6[threadIdx.x];
    gather_cluster_sums<6>(part, all, CS);
    float tt[FUSED_MAXC][3], tn[FUSED_MAXC][3];
#pragma unroll
    for (int cc = 0; cc < FUSED_MAXC; ++cc) {
      tt[cc][0] = tt[cc][1] = tt[cc][2] = tn[cc][0] = tn[cc][1] = tn[cc][2] = 0.f;
      if (cc < C)
        for (int jj = 0; jj < parts; ++jj) {
          const float* a = all + (cc * parts + jj) * 6;
          tt[cc][0] += a[0]; tt[cc][1] += a[1]; tt[cc][2] += a[2];
          tn[cc][0] += a[3]; tn[cc][1] += a[4]; tn[cc][2] += a[5];
        }
    }
    f_t = fill_from_totals(tt, c, C, fill_mode, fill_const, mean_area);
    f_n = fill_from_totals(tn, c, C, fill_mode, fill_const, mean_area);
  }
#pragma unroll
  for (int u = 0; u < K5_VPT; ++u) {
    const int i = base + (u * DG_THREADS + threadIdx.x) * 4;
    if (i < end) {
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, vs[4] = {v0[u].x, v0[u].y, v0[u].z, v0[u].w};
      const float ts[4] = {mt4[u].x, mt4[u].y, mt4[u].z, mt4[u].w}, ms[4] = {mn4[u].x, mn4[u].y, mn4[u].z, mn4[u].w};
      float sn[4] = {0.f, 0.f, 0.f, 0.f};
      if (x_in_next) { const float4 s4 = shift_next.at4(b, c, i); sn[0] = s4.x; sn[1] = s4.y; sn[2] = s4.z; sn[3] = s4.w; }
      float xn[4], xi[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xn[e] = xs[e];
        if (update) {
          const float d_t = composite(ts[e], f_t, vs[e]);
          const float d_n = composite(ms[e], f_n, vs[e]);
          xn[e] = momentum ? __fadd_rn(xs[e], __fsub_rn(d_n, d_t)) : d_n;
        }
        xi[e] = x_in_next ? __fadd_rn(xn[e], sn[e]) : 0.f;
      }
      if (s0_out) *reinterpret_cast<float4*>(s0_out + po + i) = v0[u];
      if (x_next) *reinterpret_cast<float4*>(x_next + po + i) = make_float4(xn[0], xn[1], xn[2], xn[3]);
      if (x_in_next) *reinterpret_cast<float4*>(x_in_next + po + i) = make_float4(xi[0], xi[1], xi[2], xi[3]);
    }
  }
  if (update) dg_cluster_sync();
}

template <typename... KArgs, typename... Args>
static inline void launch_fused_cluster(void (*kernel)(KArgs...), int grid, int cluster, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(DG_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// parts per channel plane for the fused form, 0 when the sample does not fit a cluster (or the fused form is disabled)
static int fused_parts(int channels, int hw, int vpt) {
  // measured on B200 at 256x3x128x128 (bench.py micro entries, L2 flushed): K1 99 us fused vs 38 us two-pass, K5 205 us vs
  // 113 us -- a sample held in the registers of a 6-CTA cluster leaves ONE 256-thread CTA per SM whose load, barrier and
  // store phases do not overlap (10 waves x ~5 us), while the two-pass kernels keep thousands of independent loads in
  // flight and re-read part of the image from L2.  Opt-in (MDM_DEGRADE_FUSED=1, read per call: the tests switch it).
  const char* env = getenv("MDM_DEGRADE_FUSED");
  const int enabled = env ? atoi(env) : 0;
  if (!enabled || (hw & 3) != 0 || channels > FUSED_MAXC) return 0;
  const int cap = DG_THREADS * 4 * vpt;
  const int parts = (hw + cap - 1) / cap;
  return channels * parts <= FUSED_MAX_CLUSTER ? parts : 0;
}

static inline int nchunks(int hw) { return (hw + DG_CHUNK - 1) / DG_CHUNK; }

}  // namespace mdm

using namespace mdm;

extern "C" {

int64_t mdm_degrade_ws_floats(int batch, int channels, int hw) {
  if (batch <= 0 || channels <= 0 || hw <= 0) return 0;
  return (int64_t)batch * channels * nchunks(hw) * 3;
}

static int degrade_launch(const void* img, int img_dtype, const uint8_t* mask, int mask_ch, int fill_mode,
                          float fill_const, int mean_area, float* x_t, float* mask_f32, float* degrade_mask,
                          float* fill_out, float* x0_out, float* ws, int batch, int channels, int hw, void* stream) {
  MDM_CHECK_ARG(img && mask && x_t, "img/mask/x_t is NULL");
  MDM_CHECK_ARG(batch > 0 && channels > 0 && hw > 0, "empty image batch");
  MDM_CHECK_ARG(mask_ch == 1 || mask_ch == channels, "mask_ch must be 1 or C");
  MDM_CHECK_ARG(img_dtype == MDM_F32 || img_dtype == MDM_BF16 || img_dtype == MDM_U8, "img dtype");
  MDM_CHECK_ARG(fill_mode >= 0 && fill_mode <= 2, "fill_mode");
  MDM_CHECK_ARG(fill_mode == MDM_FILL_CONST || ws, "workspace is NULL");
  MDM_CHECK_ARG((int64_t)batch * channels <= 65535, "batch*channels exceeds grid.y");
  cudaStream_t st = as_stream(stream);
  const int esz = img_dtype == MDM_F32 ? 4 : (img_dtype == MDM_BF16 ? 2 : 1);
  if (const int parts = fused_parts(channels, hw, K1_VPT); parts > 0 && ((uintptr_t)img & (4 * esz - 1)) == 0 && ((uintptr_t)mask & 3) == 0 &&
      ((uintptr_t)x_t & 15) == 0 && ((uintptr_t)mask_f32 & 15) == 0 && ((uintptr_t)degrade_mask & 15) == 0 && ((uintptr_t)x0_out & 15) == 0) {
    const int CS = channels * parts;
    if (img_dtype == MDM_F32)
      launch_fused_cluster(degrade_fused_kernel<float>, batch * CS, CS, st, (const float*)img, mask, mask_ch, fill_mode, fill_const, mean_area, x_t, mask_f32, degrade_mask, fill_out, x0_out, channels, hw, parts);
    else if (img_dtype == MDM_BF16)
      launch_fused_cluster(degrade_fused_kernel<__nv_bfloat16>, batch * CS, CS, st, (const __nv_bfloat16*)img, mask, mask_ch, fill_mode, fill_const, mean_area, x_t, mask_f32, degrade_mask, fill_out, x0_out, channels, hw, parts);
    else
      launch_fused_cluster(degrade_fused_kernel<uint8_t>, batch * CS, CS, st, (const uint8_t*)img, mask, mask_ch, fill_mode, fill_const, mean_area, x_t, mask_f32, degrade_mask, fill_out, x0_out, channels, hw, parts);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  const int nc = nchunks(hw);
  dim3 grid(nc, batch * channels);
  if (fill_mode != MDM_FILL_CONST) {
    if (img_dtype == MDM_F32)
      launch_pdl(degrade_stats_kernel<float>, dim3(grid), dim3(DG_THREADS), 0, st, (const float*)img, mask, mask_ch, ws, channels, hw, nc);
    else if (img_dtype == MDM_BF16)
      launch_pdl(degrade_stats_kernel<__nv_bfloat16>, dim3(grid), dim3(DG_THREADS), 0, st, (const __nv_bfloat16*)img, mask, mask_ch, ws, channels, hw, nc);
    else
      launch_pdl(degrade_stats_kernel<uint8_t>, dim3(grid), dim3(DG_THREADS), 0, st, (const uint8_t*)img, mask, mask_ch, ws, channels, hw, nc);
    MDM_LAUNCH_CHECK();
  }
  if (img_dtype == MDM_F32)
    launch_pdl(degrade_apply_kernel<float>, dim3(grid), dim3(DG_THREADS), 0, st, (const float*)img, mask, mask_ch, fill_mode, fill_const, mean_area, ws, x_t, mask_f32, degrade_mask, fill_out, x0_out, channels, hw, nc);
  else if (img_dtype == MDM_BF16)
    launch_pdl(degrade_apply_kernel<__nv_bfloat16>, dim3(grid), dim3(DG_THREADS), 0, st, (const __nv_bfloat16*)img, mask, mask_ch, fill_mode, fill_const, mean_area, ws, x_t, mask_f32, degrade_mask, fill_out, x0_out, channels, hw, nc);
  else
    launch_pdl(degrade_apply_kernel<uint8_t>, dim3(grid), dim3(DG_THREADS), 0, st, (const uint8_t*)img, mask, mask_ch, fill_mode, fill_const, mean_area, ws, x_t, mask_f32, degrade_mask, fill_out, x0_out, channels, hw, nc);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_degrade(const void* img, int img_dtype, const uint8_t* mask, int mask_ch, int fill_mode,
                float fill_const, int mean_area, float* x_t, float* mask_f32, float* degrade_mask,
                float* fill_out, float* ws, int batch, int channels, int hw, void* stream) {
  MDM_CHECK_ARG(img_dtype == MDM_F32 || img_dtype == MDM_BF16, "img dtype (uint8 images go through mdm_degrade_u8)");
  return degrade_launch(img, img_dtype, mask, mask_ch, fill_mode, fill_const, mean_area, x_t, mask_f32, degrade_mask, fill_out,
                        nullptr, ws, batch, channels, hw, stream);
}

int mdm_degrade_u8(const uint8_t* img_u8, const uint8_t* mask, int mask_ch, int fill_mode, float fill_const, int mean_area,
                   float* x_t, float* x0_out, float* mask_f32, float* degrade_mask, float* fill_out, float* ws,
                   int batch, int channels, int hw, void* stream) {
  return degrade_launch(img_u8, MDM_U8, mask, mask_ch, fill_mode, fill_const, mean_area, x_t, mask_f32, degrade_mask, fill_out,
                        x0_out, ws, batch, channels, hw, stream);
}

int mdm_sampler_step(const float* x_t, const float* net, const float* shift, int64_t sb, int64_t sc,
                     int64_t sp, const uint8_t* mask_t, const uint8_t* mask_next, int mask_ch,
                     int fill_mode, float fill_const, int mean_area, int momentum, int update,
                     const float* shift_next, int64_t nb, int64_t nc_, int64_t np_, float* x_next,
                     float* x_in_next, float* s0_out, float* ws, int batch, int channels, int hw,
                     void* stream) {
  MDM_CHECK_ARG(x_t && net, "x_t/net is NULL");
  MDM_CHECK_ARG(batch > 0 && channels > 0 && hw > 0, "empty image batch");
  MDM_CHECK_ARG(!update || (mask_t && mask_next), "masks are NULL");
  MDM_CHECK_ARG(mask_ch == 1 || mask_ch == channels, "mask_ch must be 1 or C");
  MDM_CHECK_ARG(fill_mode >= 0 && fill_mode <= 2, "fill_mode");
  MDM_CHECK_ARG((int64_t)batch * channels <= 65535, "batch*channels exceeds grid.y");
  const int nc = nchunks(hw);
  dim3 grid(nc, batch * channels);
  cudaStream_t st = as_stream(stream);
  Shift s{shift, sb, sc, sp}, sn{shift_next, nb, nc_, np_};
  float* ws_t = ws;
  float* ws_n = ws ? ws + mdm_degrade_ws_floats(batch, channels, hw) : nullptr;
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  const int vec = (hw & 3) == 0 && al16(x_t) && al16(net) && al16(x_next) && al16(x_in_next) && al16(s0_out) &&
                  (((uintptr_t)mask_t | (uintptr_t)mask_next) & 3) == 0 && shift_vec_ok(shift, sb, sc, sp) &&
                  shift_vec_ok(shift_next, nb, nc_, np_);
  if (const int parts = fused_parts(channels, hw, K5_VPT); parts > 0 && vec) {
    const int CS = channels * parts;
    launch_fused_cluster(sampler_fused_kernel, batch * CS, CS, st, x_t, net, s, mask_t, mask_next, mask_ch, fill_mode, fill_const, mean_area,
                         momentum, update, sn, x_next, x_in_next, s0_out, channels, hw, parts);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  if (update && fill_mode != MDM_FILL_CONST) {
    MDM_CHECK_ARG(ws, "workspace is NULL");
    launch_pdl(sampler_stats_kernel, dim3(grid), dim3(DG_THREADS), 0, st, x_t, net, s, mask_t, mask_next, mask_ch, ws_t, ws_n, channels, hw, nc, vec);
    MDM_LAUNCH_CHECK();
  }
  launch_pdl(sampler_update_kernel, dim3(grid), dim3(DG_THREADS), 0, st, x_t, net, s, mask_t, mask_next, mask_ch, fill_mode, fill_const, mean_area, momentum, update, sn, ws_t, ws_n, x_next, x_in_next, s0_out, channels, hw, nc, vec);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_add_shift(const float* x, const float* shift, int64_t sb, int64_t sc, int64_t sp, float* out,
                  int batch, int channels, int hw, void* stream) {
  MDM_CHECK_ARG(x && out, "x/out is NULL");
  MDM_CHECK_ARG(batch > 0 && channels > 0 && hw > 0, "empty image batch");
  dim3 grid(nchunks(hw), batch * channels);
  launch_pdl(add_shift_kernel, dim3(grid), dim3(DG_THREADS), 0, as_stream(stream), x, Shift{shift, sb, sc, sp}, out, channels, hw);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

}  // extern "C"
