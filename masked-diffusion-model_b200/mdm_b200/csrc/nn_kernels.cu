// Denoiser kernels that are not GEMMs: GroupNorm+SiLU (K2) forward/backward, the 3-channel
// first/last convolutions, nearest upsample, low-resolution attention core (K4), timestep
// embedding, bias / time-embedding gradient reductions, fused residual + MSE loss.
//
// All activations NHWC bf16 with channel stride `ld`; statistics and reductions in fp32/fp64.
// These kernels are HBM-bound: 8-byte (4 x bf16) vector accesses, a fixed channel-vector per
// thread (no per-element division), warp-shuffle + shared-memory reductions, per-chunk partial
// sums instead of atomics where the result feeds a normalisation.
#include <math.h>

#include "common.cuh"

namespace mdm {

typedef __nv_bfloat16 bf16;

struct bf16x4 {
  __nv_bfloat162 a, b;
};
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const bf16x4 v = *reinterpret_cast<const bf16x4*>(p);
  const float2 x = __bfloat1622float2(v.a), y = __bfloat1622float2(v.b);
  return make_float4(x.x, x.y, y.x, y.y);
}
__device__ __forceinline__ void st4(bf16* p, float4 f) {
  bf16x4 v;
  v.a = __floats2bfloat162_rn(f.x, f.y);
  v.b = __floats2bfloat162_rn(f.z, f.w);
  *reinterpret_cast<bf16x4*>(p) = v;
}
__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + __expf(-z)); }
__device__ __forceinline__ float siluf_(float z) { return z * sigmoidf_(z); }
__device__ __forceinline__ float dsiluf_(float z) {
  const float s = sigmoidf_(z);
  return s * (1.0f + z * (1.0f - s));
}

// =============================================================================================
// K2: GroupNorm (+SiLU) forward / backward.  HBM-bound streaming kernels, built the same way:
//   * a thread owns a FIXED 8-channel slice (16-byte vectors; its per-channel constants live in registers)
//     and walks the pixels of its CTA's chunk with a pointer increment -- no index arithmetic, no bounds test
//     in the main loop, UNROLL independent loads in flight per thread per tensor;
//   * grid = (pixel chunks, N) with ~32K elements per CTA: a few hundred CTAs, long enough to amortise the
//     per-thread set-up and the block reduction (measured: 8K-element chunks spent > 80 % of their
//     instructions outside the streaming loop);
//   * reductions: per-thread partials -> shared memory per CHANNEL (one smem atomic per channel per thread)
//     -> per group / per channel results; global atomics only once per channel per CTA.
//   fwd : stats pass (per-chunk group sums -> ws, deterministic) + apply pass (the re-read hits the 126 MB L2)
//   bwd : stats pass (dgamma/dbeta; the two group sums follow from them: sum(d*gamma) = sum_c gamma_c dbeta_c,
//         sum(d*gamma*xhat) = sum_c gamma_c dgamma_c) + apply pass, which can also emit the per-sample
//         column sums of dx (the time-embedding / conv-bias gradients) so no separate reduction runs.
// =============================================================================================
// sigmoid with ONE special-function op (tanh.approx, |rel err| <= 2^-11 -- far below the bf16 output rounding);
// exp + reciprocal costs two, and these kernels sit close to the MUFU rate on B200 (16 / clk / SM)
__device__ __forceinline__ float sigmoid_fast(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float silu_fast(float z) { return z * sigmoid_fast(z); }
constexpr int GN_MAX_GROUPS = 32;
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_C = 2048;

struct GnGeom {
  int HW, C, G, cpg, L, R, chunk_pix, nchunk;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  // bf16 -> fp32 is a 16-bit shift: even elements = low halves, odd = high halves
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// the pixels [p0, p1) of sample n handled by this thread: first pixel p0 + r, then every R-th
struct GnWalk {
  long long off;     // element offset of the first pixel's channel slice, for a tensor with pixel stride 1 (x ld)
  int count;         // pixels this thread handles
  int lane, r;
};
__device__ __forceinline__ GnWalk gn_walk(const GnGeom& g, int n, int chunk) {
  GnWalk w;
  w.lane = threadIdx.x % g.L;
  w.r = threadIdx.x / g.L;
  const int p0 = chunk * g.chunk_pix, p1 = min(p0 + g.chunk_pix, g.HW);
  w.count = (w.r < g.R && p0 + w.r < p1) ? (p1 - p0 - w.r + g.R - 1) / g.R : 0;
  w.off = (long long)n * g.HW + p0 + w.r;
  return w;
}

// add this thread's 8-channel partials into the per-channel shared accumulators dst[C].  Threads of a warp that own
// the SAME channel slice (L = C/8 < 32 lanes per pixel, i.e. 32/L pixels per warp) are combined with shuffles first:
// shared-memory float atomics are CAS loops, and same-address lanes of one warp serialise and retry (ncu: the
// short-scoreboard stall of these kernels).  Must be reached by ALL threads of the CTA; threads without data pass zeros.
__device__ __forceinline__ void gn_reduce8(float (&v)[8], float* dst, int lane, int L, bool has_data) {
  if (L < 32 && (32 % L) == 0) {
    for (int o = L; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    }
    if ((int)(threadIdx.x & 31) < L) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&dst[lane * 8 + j], v[j]);
    }
  } else if (has_data) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&dst[lane * 8 + j], v[j]);
  }
}

template <int UNROLL>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const bf16* __restrict__ x, long long ld,
                                                              float* __restrict__ ws, GnGeom g) {
  MDM_PDL_ENTER();
  extern __shared__ float gn_sm[];   // s[C], q[C]
  const int n = blockIdx.y, chunk = blockIdx.x;
  for (int c = threadIdx.x; c < 2 * g.C; c += blockDim.x) gn_sm[c] = 0.f;
  __syncthreads();
  const GnWalk w = gn_walk(g, n, chunk);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (w.count > 0) {
    const bf16* p = x + w.off * ld + w.lane * 8;
    const long long step = (long long)g.R * ld;
    int it = 0;
    for (; it + UNROLL <= w.count; it += UNROLL) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = ldg16(p + u * step);
      p += UNROLL * step;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
      }
    }
    for (; it < w.count; ++it, p += step) {
      float f[8];
      unpack8(ldg16(p), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
    }
  }
  gn_reduce8(s, gn_sm, w.lane, g.L, w.count > 0);
  gn_reduce8(q, gn_sm + g.C, w.lane, g.L, w.count > 0);
  __syncthreads();
  if (threadIdx.x < g.G) {
    float a = 0.f, b = 0.f;
    for (int c = threadIdx.x * g.cpg; c < (threadIdx.x + 1) * g.cpg; ++c) { a += gn_sm[c]; b += gn_sm[g.C + c]; }
    float* o = ws + ((long long)(n * g.nchunk + chunk)) * GN_MAX_GROUPS * 2 + threadIdx.x * 2;
    o[0] = a;
    o[1] = b;
  }
}

template <int UNROLL>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const bf16* __restrict__ x, long long ld, bf16* __restrict__ y,
                                                              long long ldy, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ ws,
                                                              float* __restrict__ stats, float eps, int silu, GnGeom g,
                                                              const float* __restrict__ qa, int qa_quads,
                                                              const float* __restrict__ qb) {
  MDM_PDL_ENTER();
  __shared__ float mr[GN_MAX_GROUPS * 2];
  const int n = blockIdx.y, chunk = blockIdx.x;
  if (threadIdx.x < g.G) {
    double s = 0.0, ss = 0.0;
    if (qa != nullptr) {
      // statistics fused into the producing convolutions' epilogues: (sum, sumsq) per 4-channel quad
      const int qpg = g.cpg >> 2, qb_quads = (g.C >> 2) - qa_quads;
      for (int k = 0; k < qpg; ++k) {
        const int quad = threadIdx.x * qpg + k;
        const float* w = quad < qa_quads ? qa + ((long long)n * qa_quads + quad) * 2
                                         : qb + ((long long)n * qb_quads + (quad - qa_quads)) * 2;
        s += w[0];
        ss += w[1];
      }
    } else
    for (int k = 0; k < g.nchunk; ++k) {
      const float* w = ws + ((long long)(n * g.nchunk + k)) * GN_MAX_GROUPS * 2 + threadIdx.x * 2;
      s += w[0];
      ss += w[1];
    }
    const double cnt = (double)g.HW * g.cpg;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    mr[threadIdx.x * 2] = (float)mean;
    mr[threadIdx.x * 2 + 1] = rstd;
    if (chunk == 0 && stats) {
      stats[((long long)n * g.G + threadIdx.x) * 2] = (float)mean;
      stats[((long long)n * g.G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  const GnWalk w = gn_walk(g, n, chunk);
  if (w.count == 0) return;
  const int c0 = w.lane * 8;
  float sc[8], sh[8];   // y = x * sc + sh
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int grp = (c0 + j) / g.cpg;
    const float mean = mr[grp * 2], rstd = mr[grp * 2 + 1];
    const float ga = gamma[c0 + j], be = beta[c0 + j];
    sc[j] = rstd * ga;
    sh[j] = be - mean * rstd * ga;
  }
  const bf16* p = x + w.off * ld + c0;
  bf16* o = y + w.off * ldy + c0;
  const long long step = (long long)g.R * ld, ostep = (long long)g.R * ldy;
  auto body = [&](const uint4& v, bf16* dst) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float z = fmaf(f[j], sc[j], sh[j]);
      if (silu) z = silu_fast(z);
      f[j] = z;
    }
    *reinterpret_cast<uint4*>(dst) = pack8(f);
  };
  int it = 0;
  for (; it + UNROLL <= w.count; it += UNROLL) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ldg16(p + u * step);
    p += UNROLL * step;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) body(v[u], o + u * ostep);
    o += UNROLL * ostep;
  }
  for (; it < w.count; ++it, p += step, o += ostep) body(ldg16(p), o);
}

// (scale, shift) per sample and channel from the quad sums of the producers' epilogues: what the convolution's
// transform warps apply on the way into its halo tiles (igemm.cu: kNorm).  grid = N, C threads-strided.
__global__ void __launch_bounds__(256) gn_coef_q_kernel(const float* __restrict__ qa, int qa_quads, const float* __restrict__ qb,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ coef, float* __restrict__ stats, int HW, int C, int G, float eps) {
  MDM_PDL_ENTER();
  __shared__ float mr[GN_MAX_GROUPS * 2];
  const int n = blockIdx.x, cpg = C / G;
  if ((int)threadIdx.x < G) {
    double s = 0.0, ss = 0.0;
    const int qpg = cpg >> 2, qb_quads = (C >> 2) - qa_quads;
    for (int k = 0; k < qpg; ++k) {
      const int quad = threadIdx.x * qpg + k;
      const float* w = quad < qa_quads ? qa + ((long long)n * qa_quads + quad) * 2 : qb + ((long long)n * qb_quads + (quad - qa_quads)) * 2;
      s += w[0];
      ss += w[1];
    }
    const double cnt = (double)HW * cpg;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    mr[threadIdx.x * 2] = (float)mean;
    mr[threadIdx.x * 2 + 1] = rstd;
    if (stats) {
      stats[((long long)n * G + threadIdx.x) * 2] = (float)mean;
      stats[((long long)n * G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = mr[(c / cpg) * 2], rstd = mr[(c / cpg) * 2 + 1];
    const float ga = gamma[c], be = beta[c];
    coef[((long long)n * C + c) * 2] = rstd * ga;
    coef[((long long)n * C + c) * 2 + 1] = be - mean * rstd * ga;
  }
}

// d(silu(z))/dz with one exp and one fast division
__device__ __forceinline__ float dsilu_fast(float z) {
  const float s = sigmoid_fast(z);
  return s * fmaf(z, 1.0f - s, 1.0f);
}

// backward pass 1: per-channel dgamma / dbeta partials; per (n, chunk, group) sums of d*gamma and d*gamma*xhat
template <int UNROLL>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_bwd_stats_kernel(const bf16* __restrict__ x, long long ld,
                                                                  const bf16* __restrict__ dy, long long lddy,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ stats, float* __restrict__ ws,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                  int silu, GnGeom g) {
  MDM_PDL_ENTER();
  extern __shared__ float gn_sm[];   // sdg[C], sdb[C]
  float* sdg = gn_sm;
  float* sdb = gn_sm + g.C;
  const int n = blockIdx.y, chunk = blockIdx.x;
  for (int c = threadIdx.x; c < 2 * g.C; c += blockDim.x) gn_sm[c] = 0.f;
  __syncthreads();
  const GnWalk w = gn_walk(g, n, chunk);
  float dg[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = db[j] = 0.f;
  if (w.count > 0) {
    const int c0 = w.lane * 8;
    // z = gamma * xhat + beta = x * za + zb.  The loop accumulates sum(d * x) and sum(d); sum(d * xhat) follows as
    // rstd * sum(d * x) - mean * rstd * sum(d): two per-channel constants live in the loop instead of four
    // (96 instead of 128 registers -> three CTAs per SM)
    float za[8], zb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = (c0 + j) / g.cpg;
      const float mean = stats[((long long)n * g.G + grp) * 2], rstd = stats[((long long)n * g.G + grp) * 2 + 1];
      const float ga = gamma[c0 + j];
      za[j] = rstd * ga;
      zb[j] = fmaf(-mean * rstd, ga, beta[c0 + j]);
    }
    const bf16* px = x + w.off * ld + c0;
    const bf16* pd = dy + w.off * lddy + c0;
    const long long sx = (long long)g.R * ld, sd = (long long)g.R * lddy;
    auto body = [&](const uint4& vx, const uint4& vd) {
      float fx[8], fd[8];
      unpack8(vx, fx);
      unpack8(vd, fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = fd[j];
        if (silu) d *= dsilu_fast(fmaf(fx[j], za[j], zb[j]));
        dg[j] = fmaf(d, fx[j], dg[j]);      // sum d * x (converted to sum d * xhat below)
        db[j] += d;
      }
    };
    int it = 0;
    for (; it + UNROLL <= w.count; it += UNROLL) {
      uint4 vx[UNROLL], vd[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) { vx[u] = ldg16(px + u * sx); vd[u] = ldg16(pd + u * sd); }
      px += UNROLL * sx;
      pd += UNROLL * sd;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) body(vx[u], vd[u]);
    }
    for (; it < w.count; ++it, px += sx, pd += sd) body(ldg16(px), ldg16(pd));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = (c0 + j) / g.cpg;
      const float mean = stats[((long long)n * g.G + grp) * 2], rstd = stats[((long long)n * g.G + grp) * 2 + 1];
      dg[j] = rstd * fmaf(-mean, db[j], dg[j]);
    }
  }
  gn_reduce8(dg, sdg, w.lane, g.L, w.count > 0);
  gn_reduce8(db, sdb, w.lane, g.L, w.count > 0);
  __syncthreads();
  if (dgamma) {
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
      atomicAdd(dgamma + c, sdg[c]);
      atomicAdd(dbeta + c, sdb[c]);
    }
  }
  if (threadIdx.x < g.G) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x * g.cpg; c < (threadIdx.x + 1) * g.cpg; ++c) {
      const float gm = gamma[c];
      s1 = fmaf(gm, sdb[c], s1);
      s2 = fmaf(gm, sdg[c], s2);
    }
    float* o = ws + ((long long)(n * g.nchunk + chunk)) * GN_MAX_GROUPS * 2 + threadIdx.x * 2;
    o[0] = s1;
    o[1] = s2;
  }
}

// backward pass 2: dx = rstd * (d*gamma - mean(d*gamma) - xhat * mean(d*gamma*xhat)) (+ add + add2);
// optionally colsum[n][c] += sum_pixels dx and dbias[c] += the same (conv bias / time-embedding gradients)
template <int UNROLL, bool HAS_ADD, bool HAS_ADD2>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_bwd_apply_kernel(
    const bf16* __restrict__ x, long long ld, const bf16* __restrict__ dy, long long lddy, const bf16* add /* may alias dx */,
    long long ldadd, const bf16* __restrict__ add2, long long ldadd2, bf16* dx, long long lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    const float* __restrict__ ws, int silu, float* __restrict__ colsum, long long ld_colsum, float* __restrict__ dbias,
    GnGeom g) {
  MDM_PDL_ENTER();
  extern __shared__ float gn_sm[];   // scs[C] when colsum
  __shared__ float m12[GN_MAX_GROUPS * 2];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const bool want_cs = colsum != nullptr || dbias != nullptr;
  if (want_cs)
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) gn_sm[c] = 0.f;
  if (threadIdx.x < g.G) {
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < g.nchunk; ++k) {
      const float* w = ws + ((long long)(n * g.nchunk + k)) * GN_MAX_GROUPS * 2 + threadIdx.x * 2;
      s1 += w[0];
      s2 += w[1];
    }
    const double cnt = (double)g.HW * g.cpg;
    m12[threadIdx.x * 2] = (float)(s1 / cnt);
    m12[threadIdx.x * 2 + 1] = (float)(s2 / cnt);
  }
  __syncthreads();
  const GnWalk w = gn_walk(g, n, chunk);
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  if (w.count > 0) {
    const int c0 = w.lane * 8;
    // dx = d * A + x * B + Cc   with d already multiplied by silu'(z):  A = rstd*gamma, B = -rstd^2*m2, Cc = -rstd*(m1 - mean*rstd*m2)
    float zb[8], A[8], Bc[8], Cc[8];       // z = x * A + zb (A = rstd * gamma doubles as the z slope)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = (c0 + j) / g.cpg;
      const float mean = stats[((long long)n * g.G + grp) * 2], rstd = stats[((long long)n * g.G + grp) * 2 + 1];
      const float m1 = m12[grp * 2], m2 = m12[grp * 2 + 1];
      const float mbj = -mean * rstd;
      A[j] = rstd * gamma[c0 + j];
      zb[j] = fmaf(mbj, gamma[c0 + j], beta[c0 + j]);
      Bc[j] = -rstd * rstd * m2;
      Cc[j] = -rstd * (m1 + mbj * m2);
    }
    const bf16* px = x + w.off * ld + c0;
    const bf16* pd = dy + w.off * lddy + c0;
    const bf16* pa = HAS_ADD ? add + w.off * ldadd + c0 : nullptr;
    const bf16* pb = HAS_ADD2 ? add2 + w.off * ldadd2 + c0 : nullptr;
    bf16* po = dx + w.off * lddx + c0;
    const long long sx = (long long)g.R * ld, sd = (long long)g.R * lddy, sa = (long long)g.R * ldadd,
                    sb = (long long)g.R * ldadd2, so = (long long)g.R * lddx;
    auto body = [&](const uint4& vx, const uint4& vd, const uint4& va, const uint4& vb, bf16* dst) {
      float fx[8], fd[8], o[8];
      unpack8(vx, fx);
      unpack8(vd, fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = fd[j];
        if (silu) d *= dsilu_fast(fmaf(fx[j], A[j], zb[j]));
        o[j] = fmaf(d, A[j], fmaf(fx[j], Bc[j], Cc[j]));
      }
      if (want_cs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += o[j];
      }
      if (HAS_ADD) {
        float fa[8];
        unpack8(va, fa);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += fa[j];
      }
      if (HAS_ADD2) {
        float fb[8];
        unpack8(vb, fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += fb[j];
      }
      *reinterpret_cast<uint4*>(dst) = pack8(o);
    };
    int it = 0;
    for (; it + UNROLL <= w.count; it += UNROLL) {
      uint4 vx[UNROLL], vd[UNROLL], va[UNROLL], vb[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        vx[u] = ldg16(px + u * sx);
        vd[u] = ldg16(pd + u * sd);
        if (HAS_ADD) va[u] = *reinterpret_cast<const uint4*>(pa + u * sa);   // may alias dx: plain load
        if (HAS_ADD2) vb[u] = ldg16(pb + u * sb);
      }
      px += UNROLL * sx;
      pd += UNROLL * sd;
      if (HAS_ADD) pa += UNROLL * sa;
      if (HAS_ADD2) pb += UNROLL * sb;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) body(vx[u], vd[u], va[u], vb[u], po + u * so);
      po += UNROLL * so;
    }
    for (; it < w.count; ++it) {
      uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
      if (HAS_ADD) { va = *reinterpret_cast<const uint4*>(pa); pa += sa; }
      if (HAS_ADD2) { vb = ldg16(pb); pb += sb; }
      body(ldg16(px), ldg16(pd), va, vb, po);
      px += sx; pd += sd; po += so;
    }
  }
  if (want_cs) {
    gn_reduce8(cs, gn_sm, w.lane, g.L, w.count > 0);
    __syncthreads();
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
      const float v = gn_sm[c];
      if (colsum) atomicAdd(colsum + (long long)n * ld_colsum + c, v);
      if (dbias) atomicAdd(dbias + c, v);
    }
  }
}

// ---- small maps: one CTA per sample keeps the whole sample in registers -> ONE kernel, one read ------------
// (low-resolution levels of the U-Net: <= 32K elements per sample forward, <= 16K backward)
template <int MAXV>
__global__ void __launch_bounds__(GN_THREADS) gn_small_fwd_kernel(const bf16* __restrict__ x, long long ld, bf16* __restrict__ y,
                                                                  long long ldy, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float* __restrict__ stats,
                                                                  float eps, int silu, GnGeom g) {
  MDM_PDL_ENTER();
  extern __shared__ float gn_sm[];   // s[C], q[C]
  __shared__ float mr[GN_MAX_GROUPS * 2];
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < 2 * g.C; c += blockDim.x) gn_sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x % g.L, r = threadIdx.x / g.L;
  const int count = (r < g.R && r < g.HW) ? (g.HW - r + g.R - 1) / g.R : 0;
  const int c0 = lane * 8;
  const bf16* p = x + ((long long)n * g.HW + r) * ld + c0;
  const long long step = (long long)g.R * ld;
  uint4 v[MAXV];
  {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    if (count > 0) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) v[i] = i < count ? ldg16(p + i * step) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        float f[8];
        unpack8(v[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
      }
    }
    gn_reduce8(s, gn_sm, lane, g.L, count > 0);
    gn_reduce8(q, gn_sm + g.C, lane, g.L, count > 0);
  }
  __syncthreads();
  if (threadIdx.x < g.G) {
    double a = 0.0, b = 0.0;
    for (int c = threadIdx.x * g.cpg; c < (threadIdx.x + 1) * g.cpg; ++c) { a += gn_sm[c]; b += gn_sm[g.C + c]; }
    const double cnt = (double)g.HW * g.cpg;
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    mr[threadIdx.x * 2] = (float)mean;
    mr[threadIdx.x * 2 + 1] = rstd;
    if (stats) {
      stats[((long long)n * g.G + threadIdx.x) * 2] = (float)mean;
      stats[((long long)n * g.G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  if (count == 0) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int grp = (c0 + j) / g.cpg;
    const float mean = mr[grp * 2], rstd = mr[grp * 2 + 1];
    const float ga = gamma[c0 + j], be = beta[c0 + j];
    sc[j] = rstd * ga;
    sh[j] = be - mean * rstd * ga;
  }
  bf16* o = y + ((long long)n * g.HW + r) * ldy + c0;
  const long long ostep = (long long)g.R * ldy;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < count) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = fmaf(f[j], sc[j], sh[j]);
        if (silu) z = silu_fast(z);
        f[j] = z;
      }
      *reinterpret_cast<uint4*>(o + i * ostep) = pack8(f);
    }
  }
}

template <int MAXV>
__global__ void __launch_bounds__(GN_THREADS, 1) gn_small_bwd_kernel(
    const bf16* __restrict__ x, long long ld, const bf16* __restrict__ dy, long long lddy, const bf16* add /* may alias dx */,
    long long ldadd, const bf16* __restrict__ add2, long long ldadd2, bf16* dx, long long lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int silu, float* __restrict__ colsum, long long ld_colsum,
    float* __restrict__ dbias, GnGeom g) {
  MDM_PDL_ENTER();
  extern __shared__ float gn_sm[];   // sdg[C], sdb[C]; reused as scs[C]
  __shared__ float m12[GN_MAX_GROUPS * 2];
  float* sdg = gn_sm;
  float* sdb = gn_sm + g.C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < 2 * g.C; c += blockDim.x) gn_sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x % g.L, r = threadIdx.x / g.L;
  const int count = (r < g.R && r < g.HW) ? (g.HW - r + g.R - 1) / g.R : 0;
  const int c0 = lane * 8;
  const long long first = (long long)n * g.HW + r;
  uint4 vx[MAXV], vd[MAXV];
  float rs[8], mb[8], ga[8], be[8];
  float dg[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = db[j] = 0.f;
  if (count > 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = (c0 + j) / g.cpg;
      const float mean = stats[((long long)n * g.G + grp) * 2], rstd = stats[((long long)n * g.G + grp) * 2 + 1];
      rs[j] = rstd;
      mb[j] = -mean * rstd;
      ga[j] = gamma[c0 + j];
      be[j] = beta[c0 + j];
    }
    const bf16* px = x + first * ld + c0;
    const bf16* pd = dy + first * lddy + c0;
    const long long sx = (long long)g.R * ld, sd = (long long)g.R * lddy;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      vx[i] = i < count ? ldg16(px + i * sx) : make_uint4(0, 0, 0, 0);
      vd[i] = i < count ? ldg16(pd + i * sd) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      float fx[8], fd[8];
      unpack8(vx[i], fx);
      unpack8(vd[i], fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(fx[j], rs[j], mb[j]);
        float d = fd[j];
        if (silu) d *= dsilu_fast(fmaf(xh, ga[j], be[j]));
        dg[j] = fmaf(d, xh, dg[j]);
        db[j] += d;
      }
    }
  }
  gn_reduce8(dg, sdg, lane, g.L, count > 0);
  gn_reduce8(db, sdb, lane, g.L, count > 0);
  __syncthreads();
  if (dgamma) {
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
      atomicAdd(dgamma + c, sdg[c]);
      atomicAdd(dbeta + c, sdb[c]);
    }
  }
  if (threadIdx.x < g.G) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x * g.cpg; c < (threadIdx.x + 1) * g.cpg; ++c) {
      const float gm = gamma[c];
      s1 = fmaf(gm, sdb[c], s1);
      s2 = fmaf(gm, sdg[c], s2);
    }
    const float cnt = (float)g.HW * g.cpg;
    m12[threadIdx.x * 2] = s1 / cnt;
    m12[threadIdx.x * 2 + 1] = s2 / cnt;
  }
  __syncthreads();
  const bool want_cs = colsum != nullptr || dbias != nullptr;
  if (want_cs)
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) gn_sm[c] = 0.f;   // sdg no longer needed
  if (want_cs) __syncthreads();
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  if (count > 0) {
    float A[8], Bc[8], Cc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = (c0 + j) / g.cpg;
      const float m1 = m12[grp * 2], m2 = m12[grp * 2 + 1];
      A[j] = rs[j] * ga[j];
      Bc[j] = -rs[j] * rs[j] * m2;
      Cc[j] = -rs[j] * (m1 + mb[j] * m2);
    }
    const bf16* pa = add ? add + first * ldadd + c0 : nullptr;
    const bf16* pb = add2 ? add2 + first * ldadd2 + c0 : nullptr;
    bf16* po = dx + first * lddx + c0;
    const long long sa = (long long)g.R * ldadd, sb = (long long)g.R * ldadd2, so = (long long)g.R * lddx;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (i < count) {
        float fx[8], fd[8], o[8];
        unpack8(vx[i], fx);
        unpack8(vd[i], fd);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d = fd[j];
          if (silu) d *= dsilu_fast(fmaf(fmaf(fx[j], rs[j], mb[j]), ga[j], be[j]));
          o[j] = fmaf(d, A[j], fmaf(fx[j], Bc[j], Cc[j]));
        }
        if (want_cs) {
#pragma unroll
          for (int j = 0; j < 8; ++j) cs[j] += o[j];
        }
        if (pa) {
          float fa[8];
          unpack8(*reinterpret_cast<const uint4*>(pa + i * sa), fa);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += fa[j];
        }
        if (pb) {
          float fb[8];
          unpack8(ldg16(pb + i * sb), fb);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += fb[j];
        }
        *reinterpret_cast<uint4*>(po + i * so) = pack8(o);
      }
    }
  }
  if (want_cs) {
    gn_reduce8(cs, gn_sm, lane, g.L, count > 0);
    __syncthreads();
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
      const float v = gn_sm[c];
      if (colsum) atomicAdd(colsum + (long long)n * ld_colsum + c, v);
      if (dbias) atomicAdd(dbias + c, v);
    }
  }
}

// ---- medium maps: ONE kernel, one HBM read -- a thread-block CLUSTER per sample --------------------------------
// The sample (up to 16 x 32K elements forward, 16 x 16K backward) is split by pixel range over the K CTAs of a
// cluster.  Every thread copies its vectors with cp.async into thread-private shared-memory slots (slot i of
// thread t at [i][t]: conflict-free, no register cost -> 3 CTAs per SM overlap one CTA's load phase with another's
// store phase), the per-group partial sums are PUSHED into every peer's shared memory (DSMEM stores) before one
// cluster barrier, and the normalisation runs from the staged copy: x (and dy) cross HBM/L2 exactly once.
constexpr int GN_MAX_CLUSTER = 16;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store two floats into the same shared-memory offset of CTA `rank` of this cluster
__device__ __forceinline__ void dsmem_store2(float* local_ptr, uint32_t rank, float a, float b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}

template <int MAXV>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_cluster_fwd_kernel(const bf16* __restrict__ x, long long ld, bf16* __restrict__ y,
                                                                       long long ldy, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, float* __restrict__ stats,
                                                                       float eps, int silu, GnGeom g, int K) {
  MDM_PDL_ENTER();
  extern __shared__ __align__(16) uint8_t gn_dyn[];
  uint4* slots = reinterpret_cast<uint4*>(gn_dyn);                               // [MAXV][GN_THREADS]
  float* ch = reinterpret_cast<float*>(gn_dyn + (size_t)MAXV * GN_THREADS * 16);  // s[C], q[C]
  __shared__ __align__(8) float recv[GN_MAX_CLUSTER * GN_MAX_GROUPS * 2];         // [rank][group]{sum, sumsq}
  __shared__ float mr[GN_MAX_GROUPS * 2];
  const uint32_t rank = cluster_ctarank();
  const int n = blockIdx.x / K;
  const int HWp = g.HW / K, p0 = (int)rank * HWp;
  const int lane = threadIdx.x % g.L, r = threadIdx.x / g.L;
  const int count = (r < g.R && r < HWp) ? (HWp - r + g.R - 1) / g.R : 0;
  const int c0 = lane * 8;
  {
    const bf16* p = x + ((long long)n * g.HW + p0 + r) * ld + c0;
    const long long step = (long long)g.R * ld;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < count) cp_async16(&slots[i * GN_THREADS + threadIdx.x], p + i * step);
  }
  for (int c = threadIdx.x; c < 2 * g.C; c += blockDim.x) ch[c] = 0.f;
  cp_async_wait_all();
  __syncthreads();
  {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
#pragma unroll 4
    for (int i = 0; i < count; ++i) {
      float f[8];
      unpack8(slots[i * GN_THREADS + threadIdx.x], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
    }
    gn_reduce8(s, ch, lane, g.L, count > 0);
    gn_reduce8(q, ch + g.C, lane, g.L, count > 0);
  }
  __syncthreads();
  if (threadIdx.x < g.G) {
    float a = 0.f, b = 0.f;
    for (int c = threadIdx.x * g.cpg; c < (threadIdx.x + 1) * g.cpg; ++c) { a += ch[c]; b += ch[g.C + c]; }
    float* slot = &recv[(rank * GN_MAX_GROUPS + threadIdx.x) * 2];
    for (int peer = 0; peer < K; ++peer) dsmem_store2(slot, (uint32_t)peer, a, b);
  }
  cluster_sync_all();
  if (threadIdx.x < g.G) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < K; ++k) { a += recv[(k * GN_MAX_GROUPS + threadIdx.x) * 2]; b += recv[(k * GN_MAX_GROUPS + threadIdx.x) * 2 + 1]; }
    const double cnt = (double)g.HW * g.cpg;
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    mr[threadIdx.x * 2] = (float)mean;
    mr[threadIdx.x * 2 + 1] = rstd;
    if (rank == 0 && stats) {
      stats[((long long)n * g.G + threadIdx.x) * 2] = (float)mean;
      stats[((long long)n * g.G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  if (count == 0) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int grp = (c0 + j) / g.cpg;
    const float mean = mr[grp * 2], rstd = mr[grp * 2 + 1];
    const float ga = gamma[c0 + j], be = beta[c0 + j];
    sc[j] = rstd * ga;
    sh[j] = be - mean * rstd * ga;
  }
  bf16* o = y + ((long long)n * g.HW + p0 + r) * ldy + c0;
  const long long ostep = (long long)g.R * ldy;
#pragma unroll 4
  for (int i = 0; i < count; ++i) {
    float f[8];
    unpack8(slots[i * GN_THREADS + threadIdx.x], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float z = fmaf(f[j], sc[j], sh[j]);
      if (silu) z = silu_fast(z);
      f[j] = z;
    }
    *reinterpret_cast<uint4*>(o + i * ostep) = pack8(f);
  }
}

template <int MAXV>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_cluster_bwd_kernel(
    const bf16* __restrict__ x, long long ld, const bf16* __restrict__ dy, long long lddy, const bf16* add /* may alias dx */,
    long long ldadd, const bf16* __restrict__ add2, long long ldadd2, bf16* dx, long long lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int silu, float* __restrict__ colsum, long long ld_colsum,
    float* __restrict__ dbias, GnGeom g, int K) {
  MDM_PDL_ENTER();
  extern __shared__ __align__(16) uint8_t gn_dyn[];
  uint4* sx = reinterpret_cast<uint4*>(gn_dyn);                                    // [MAXV][GN_THREADS]
  uint4* sdy = sx + MAXV * GN_THREADS;                                             // [MAXV][GN_THREADS]
  float* sdg = reinterpret_cast<float*>(gn_dyn + (size_t)2 * MAXV * GN_THREADS * 16);   // [C]; reused for the column sums
  float* sdb = sdg + g.C;
  __shared__ __align__(8) float recv[GN_MAX_CLUSTER * GN_MAX_GROUPS * 2];
  __shared__ float m12[GN_MAX_GROUPS * 2];
  const uint32_t rank = cluster_ctarank();
  const int n = blockIdx.x / K;
  const int HWp = g.HW / K, p0 = (int)rank * HWp;
  const int lane = threadIdx.x % g.L, r = threadIdx.x / g.L;
  const int count = (r < g.R && r < HWp) ? (HWp - r + g.R - 1) / g.R : 0;
  const int c0 = lane * 8;
  const long long first = (long long)n * g.HW + p0 + r;
  {
    const bf16* px = x + first * ld + c0;
    const bf16* pd = dy + first * lddy + c0;
    const long long stx = (long long)g.R * ld, std_ = (long long)g.R * lddy;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < count) {
        cp_async16(&sx[i * GN_THREADS + threadIdx.x], px + i * stx);
        cp_async16(&sdy[i * GN_THREADS + threadIdx.x], pd + i * std_);
      }
  }
  for (int c = threadIdx.x; c < 2 * g.C; c += blockDim.x) sdg[c] = 0.f;
  float rs[8], mb[8], ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int grp = (c0 + j) / g.cpg;
    const float mean = stats[((long long)n * g.G + grp) * 2], rstd = stats[((long long)n * g.G + grp) * 2 + 1];
    rs[j] = rstd;
    mb[j] = -mean * rstd;
    ga[j] = gamma[c0 + j];
    be[j] = beta[c0 + j];
  }
  cp_async_wait_all();
  __syncthreads();
  {
    float dg[8], db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[j] = db[j] = 0.f;
#pragma unroll 2
    for (int i = 0; i < count; ++i) {
      float fx[8], fd[8];
      unpack8(sx[i * GN_THREADS + threadIdx.x], fx);
      unpack8(sdy[i * GN_THREADS + threadIdx.x], fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(fx[j], rs[j], mb[j]);
        float d = fd[j];
        if (silu) d *= dsilu_fast(fmaf(xh, ga[j], be[j]));
        dg[j] = fmaf(d, xh, dg[j]);
        db[j] += d;
      }
    }
    gn_reduce8(dg, sdg, lane, g.L, count > 0);
    gn_reduce8(db, sdb, lane, g.L, count > 0);
  }
  __syncthreads();
  if (dgamma) {
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
      atomicAdd(dgamma + c, sdg[c]);
      atomicAdd(dbeta + c, sdb[c]);
    }
  }
  if (threadIdx.x < g.G) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x * g.cpg; c < (threadIdx.x + 1) * g.cpg; ++c) {
      const float gm = gamma[c];
      s1 = fmaf(gm, sdb[c], s1);
      s2 = fmaf(gm, sdg[c], s2);
    }
    float* slot = &recv[(rank * GN_MAX_GROUPS + threadIdx.x) * 2];
    for (int peer = 0; peer < K; ++peer) dsmem_store2(slot, (uint32_t)peer, s1, s2);
  }
  cluster_sync_all();   // (also a CTA barrier: sdg / sdb are free again)
  if (threadIdx.x < g.G) {
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < K; ++k) { s1 += recv[(k * GN_MAX_GROUPS + threadIdx.x) * 2]; s2 += recv[(k * GN_MAX_GROUPS + threadIdx.x) * 2 + 1]; }
    const double cnt = (double)g.HW * g.cpg;
    m12[threadIdx.x * 2] = (float)(s1 / cnt);
    m12[threadIdx.x * 2 + 1] = (float)(s2 / cnt);
  }
  const bool want_cs = colsum != nullptr || dbias != nullptr;
  if (want_cs)
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) sdg[c] = 0.f;
  __syncthreads();
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  if (count > 0) {
    float A[8], Bc[8], Cc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = (c0 + j) / g.cpg;
      const float m1 = m12[grp * 2], m2 = m12[grp * 2 + 1];
      A[j] = rs[j] * ga[j];
      Bc[j] = -rs[j] * rs[j] * m2;
      Cc[j] = -rs[j] * (m1 + mb[j] * m2);
    }
    const bf16* pa = add ? add + first * ldadd + c0 : nullptr;
    const bf16* pb = add2 ? add2 + first * ldadd2 + c0 : nullptr;
    bf16* po = dx + first * lddx + c0;
    const long long sa = (long long)g.R * ldadd, sb = (long long)g.R * ldadd2, so = (long long)g.R * lddx;
#pragma unroll 2
    for (int i = 0; i < count; ++i) {
      float fx[8], fd[8], o[8];
      uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
      if (pa) va = *reinterpret_cast<const uint4*>(pa + i * sa);
      if (pb) vb = ldg16(pb + i * sb);
      unpack8(sx[i * GN_THREADS + threadIdx.x], fx);
      unpack8(sdy[i * GN_THREADS + threadIdx.x], fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = fd[j];
        if (silu) d *= dsilu_fast(fmaf(fmaf(fx[j], rs[j], mb[j]), ga[j], be[j]));
        o[j] = fmaf(d, A[j], fmaf(fx[j], Bc[j], Cc[j]));
      }
      if (want_cs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += o[j];
      }
      if (pa) {
        float fa[8];
        unpack8(va, fa);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += fa[j];
      }
      if (pb) {
        float fb[8];
        unpack8(vb, fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += fb[j];
      }
      *reinterpret_cast<uint4*>(po + i * so) = pack8(o);
    }
  }
  if (want_cs) {
    gn_reduce8(cs, sdg, lane, g.L, count > 0);
    __syncthreads();
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
      const float v = sdg[c];
      if (colsum) atomicAdd(colsum + (long long)n * ld_colsum + c, v);
      if (dbias) atomicAdd(dbias + c, v);
    }
  }
}

template <typename... KArgs, typename... Args>
static inline void launch_cluster(void (*kernel)(KArgs...), int grid, int cluster, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GN_THREADS);
  cfg.dynamicSmemBytes = smem;
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

constexpr int GN_CL_FWD_V = 16, GN_CL_BWD_V = 8;   // 16-byte vectors per thread held in the staged copy
static size_t gn_cluster_smem(int C, bool bwd) {
  return (size_t)(bwd ? 2 * GN_CL_BWD_V : GN_CL_FWD_V) * GN_THREADS * 16 + (size_t)2 * C * sizeof(float);
}
// largest usable cluster size (16 needs the non-portable opt-in and enough co-schedulable SMs per GPC)
static int gn_cluster_limit() {
  static int limit = -1;
  if (limit >= 0) return limit;
  limit = 0;
  const int want = 16;   // (the MDM_GN_CLUSTER cap is applied per call in gn_cluster_size)
  const size_t sm_max = gn_cluster_smem(GN_MAX_C, false);
  if (cudaFuncSetAttribute(gn_cluster_fwd_kernel<GN_CL_FWD_V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_max) != cudaSuccess ||
      cudaFuncSetAttribute(gn_cluster_bwd_kernel<GN_CL_BWD_V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_max) != cudaSuccess) {
    cudaGetLastError();
    return limit;
  }
  limit = 8;
  if (want >= 16 &&
      cudaFuncSetAttribute(gn_cluster_fwd_kernel<GN_CL_FWD_V>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
      cudaFuncSetAttribute(gn_cluster_bwd_kernel<GN_CL_BWD_V>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(GN_THREADS);
    cfg.dynamicSmemBytes = gn_cluster_smem(512, true);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, gn_cluster_bwd_kernel<GN_CL_BWD_V>, &cfg) == cudaSuccess && nclusters >= 4) limit = 16;
  }
  cudaGetLastError();
  if (want < limit) limit = want >= 8 ? 8 : (want >= 4 ? 4 : 2);
  return limit;
}
// cluster size for a sample of HW x C elements with `maxv` vectors per thread; 0: does not fit
static int gn_cluster_size(const GnGeom& g, int maxv) {
  int limit = gn_cluster_limit();
  if (const char* v = getenv("MDM_GN_CLUSTER")) {        // read per call: tests switch the kernel family
    const int want = atoi(v);
    if (want < limit) limit = want;
  }
  for (int K = 2; K <= limit; K *= 2) {
    if (g.HW % K) return 0;
    const int HWp = g.HW / K;
    if ((HWp + g.R - 1) / g.R <= maxv) return K;
  }
  return 0;
}

// CTAs of the two-pass kernel family that are co-resident on the whole GPU (occupancy x SMs), queried once
static int gn_slots(bool bwd) {
  static int slots[2] = {0, 0};
  if (slots[bwd] == 0) {
    int a = 0, b = 0;
    cudaError_t e1, e2;
    if (!bwd) {
      e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, gn_stats_kernel<8>, GN_THREADS, 4096);
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, gn_apply_kernel<4>, GN_THREADS, 0);
    } else {
      e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, gn_bwd_stats_kernel<4>, GN_THREADS, 4096);
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, gn_bwd_apply_kernel<2, true, true>, GN_THREADS, 2048);
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess || a < 1 || b < 1) { cudaGetLastError(); a = b = bwd ? 3 : 5; }
    slots[bwd] = (a < b ? a : b) * kNumSMs;
  }
  return slots[bwd];
}

// geometry of the two-pass kernels: a CTA owns `chunk_pix` pixels of one sample.  The chunk is chosen so that the
// grid (nbatch x nchunk CTAs) is a whole number of WAVES of the co-resident CTAs with as little per-CTA overhead as
// possible: cost(waves) = waves x (chunk elements + ~24K elements' worth of set-up / reduction / tail).  A grid one
// CTA over a wave costs a whole extra wave (measured: 512 CTAs on 444 slots ran as long as 888).
static int gn_geom(GnGeom& g, int HW, int C, int G, int* nchunk, int nbatch = 0, int slots = 0) {
  if (C % 8 != 0 || C % G != 0 || G > GN_MAX_GROUPS || C > GN_MAX_C) {
    set_error("GroupNorm: C=%d must be a multiple of 8 and of G=%d (G <= 32, C <= %d)", C, G, GN_MAX_C);
    return MDM_E_ARG;
  }
  g.HW = HW; g.C = C; g.G = G; g.cpg = C / G;
  g.L = C / 8;
  if (g.L > GN_THREADS) { set_error("GroupNorm: C=%d too wide", C); return MDM_E_ARG; }
  g.R = GN_THREADS / g.L;
  int chunk;
  if (nbatch > 0 && slots > 0) {
    long long best = -1;
    chunk = HW;
    const int max_chunks = HW / g.R > 0 ? HW / g.R : 1;
    for (int w = 1; w <= 8; ++w) {
      int nch = (int)(((long long)slots * w) / nbatch);
      if (nch < 1) nch = 1;
      if (nch > max_chunks) nch = max_chunks;
      int ck = (HW + nch - 1) / nch;
      ck = ((ck + g.R - 1) / g.R) * g.R;
      if (ck > HW) ck = HW;
      if ((long long)ck * C < 8192 && ck < HW) continue;            // too little work per CTA
      const int nch2 = (HW + ck - 1) / ck;
      const long long waves = ((long long)nbatch * nch2 + slots - 1) / slots;
      const long long cost = waves * ((long long)ck * C + 24576);
      if (best < 0 || cost < best) { best = cost; chunk = ck; }
    }
  } else {
    // no batch information: ~32K elements (64 KB of bf16) per CTA, a multiple of R pixels
    chunk = (32768 + C - 1) / C;
    chunk = ((chunk + g.R - 1) / g.R) * g.R;
  }
  if (chunk > HW) chunk = HW;
  if (chunk < 1) chunk = 1;
  g.chunk_pix = chunk;
  g.nchunk = (HW + chunk - 1) / chunk;
  *nchunk = g.nchunk;
  return MDM_OK;
}

// =============================================================================================
// First conv (C_img -> cout, 3x3): planar fp32 image in, NHWC bf16 out.  thread <-> output
// channel, weights in registers, input rows staged in shared memory (broadcast reads).  The same
// kernel is the dgrad of the LAST conv (flip = 1: dy image in, dx NHWC out).
// =============================================================================================
constexpr int PL_MAXC = 4;       // image channels supported (1..4)
constexpr int PL_TW = 32;        // pixels per tile (along W)

template <int FLIP_T>
__global__ void planar_to_nhwc_conv_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                           const float* __restrict__ bias, bf16* __restrict__ y, long long ldy,
                                           int C, int H, int W, int cout, int w_is_out_layout) {
  MDM_PDL_ENTER();
  // tile: one image row h, PL_TW pixels
  __shared__ float tile[PL_MAXC][3][PL_TW + 2];
  const int tiles_w = (W + PL_TW - 1) / PL_TW;
  const int tw = blockIdx.x % tiles_w, h = blockIdx.x / tiles_w, n = blockIdx.y;
  const int w0 = tw * PL_TW;
  for (int i = threadIdx.x; i < C * 3 * (PL_TW + 2); i += blockDim.x) {
    const int c = i / (3 * (PL_TW + 2)), rem = i % (3 * (PL_TW + 2)), rr = rem / (PL_TW + 2), xx = rem % (PL_TW + 2);
    const int hh = h + rr - 1, ww = w0 + xx - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = img[(((long long)n * C + c) * H + hh) * W + ww];
    tile[c][rr][xx] = v;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    float wr[PL_MAXC][9];
#pragma unroll
    for (int c = 0; c < PL_MAXC; ++c)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int tt = FLIP_T ? 8 - t : t;
        // conv_in weight [cout][C][3][3]; conv_out weight [C][cin=cout here][3][3]
        wr[c][t] = c < C ? (w_is_out_layout ? w[((long long)c * cout + co) * 9 + tt] : w[((long long)co * C + c) * 9 + tt]) : 0.f;
      }
    const float b = bias ? bias[co] : 0.f;
    for (int px = 0; px < PL_TW && w0 + px < W; ++px) {
      float acc = b;
#pragma unroll
      for (int c = 0; c < PL_MAXC; ++c)
#pragma unroll
        for (int t = 0; t < 9; ++t)
          if (c < C) acc += tile[c][t / 3][px + (t % 3)] * wr[c][t];
      y[(((long long)n * H + h) * W + w0 + px) * ldy + co] = __float2bfloat16(acc);
    }
  }
}

// wgrad twin: dw[..] += sum_p nhwc[p][co] * img[c][p + tap]; dbias[co] += sum_p nhwc[p][co]
template <int FLIP_T>
__global__ void planar_wgrad_kernel(const float* __restrict__ img, const bf16* __restrict__ t, long long ldt,
                                    float* __restrict__ dw, float* __restrict__ dbias, int C, int H, int W, int cout,
                                    int w_is_out_layout, int rows_per_cta) {
  MDM_PDL_ENTER();
  __shared__ float tile[PL_MAXC][3][PL_TW + 2];
  const int tiles_w = (W + PL_TW - 1) / PL_TW;
  const int n = blockIdx.y;
  const int co = threadIdx.x;
  float acc[PL_MAXC][9];
  float accb = 0.f;
#pragma unroll
  for (int c = 0; c < PL_MAXC; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;
  for (int job = blockIdx.x * rows_per_cta; job < min((blockIdx.x + 1) * rows_per_cta, H * tiles_w); ++job) {
    const int tw = job % tiles_w, h = job / tiles_w, w0 = tw * PL_TW;
    __syncthreads();
    for (int i = threadIdx.x; i < C * 3 * (PL_TW + 2); i += blockDim.x) {
      const int c = i / (3 * (PL_TW + 2)), rem = i % (3 * (PL_TW + 2)), rr = rem / (PL_TW + 2), xx = rem % (PL_TW + 2);
      const int hh = h + rr - 1, ww = w0 + xx - 1;
      float v = 0.f;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = img[(((long long)n * C + c) * H + hh) * W + ww];
      tile[c][rr][xx] = v;
    }
    __syncthreads();
    if (co < cout) {
      for (int px = 0; px < PL_TW && w0 + px < W; ++px) {
        const float d = __bfloat162float(t[(((long long)n * H + h) * W + w0 + px) * ldt + co]);
        accb += d;
#pragma unroll
        for (int c = 0; c < PL_MAXC; ++c)
#pragma unroll
          for (int k = 0; k < 9; ++k)
            if (c < C) acc[c][k] += d * tile[c][k / 3][px + (k % 3)];
      }
    }
  }
  if (co < cout) {
#pragma unroll
    for (int c = 0; c < PL_MAXC; ++c)
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (c >= C) continue;
        const int kk = FLIP_T ? 8 - k : k;
        float* dst = w_is_out_layout ? dw + ((long long)c * cout + co) * 9 + kk : dw + ((long long)co * C + c) * 9 + kk;
        atomicAdd(dst, acc[c][k]);
      }
    if (dbias) atomicAdd(dbias + co, accb);
  }
}

// last conv forward: NHWC bf16 [.,cin] -> planar fp32 [N][C][H][W]; thread <-> pixel
__global__ void nhwc_to_planar_conv_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ w,
                                           const float* __restrict__ bias, float* __restrict__ y, int C, int H, int W, int cin) {
  MDM_PDL_ENTER();
  extern __shared__ float wsm[];  // [9][cin][4]
  for (int i = threadIdx.x; i < 9 * cin * 4; i += blockDim.x) {
    const int c = i & 3, ci = (i >> 2) % cin, t = (i >> 2) / cin;
    wsm[i] = c < C ? w[((long long)c * cin + ci) * 9 + t] : 0.f;
  }
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= (long long)H * W) return;
  const int h = (int)(p / W), ww = (int)(p % W);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int t = 0; t < 9; ++t) {
    const int hh = h + t / 3 - 1, wx = ww + t % 3 - 1;
    if (hh < 0 || hh >= H || wx < 0 || wx >= W) continue;
    const bf16* xp = x + (((long long)n * H + hh) * W + wx) * ldx;
    const float4* wp = reinterpret_cast<const float4*>(wsm + (long long)t * cin * 4);
    for (int ci = 0; ci < cin; ci += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(xp + ci);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        const float4 w0 = wp[ci + e * 2], w1 = wp[ci + e * 2 + 1];
        a0 += f.x * w0.x + f.y * w1.x; a1 += f.x * w0.y + f.y * w1.y;
        a2 += f.x * w0.z + f.y * w1.z; a3 += f.x * w0.w + f.y * w1.w;
      }
    }
  }
  const float accs[4] = {a0, a1, a2, a3};
  for (int c = 0; c < C; ++c) y[(((long long)n * C + c) * H + h) * W + ww] = accs[c] + (bias ? bias[c] : 0.f);
}

__global__ void planar_sum_kernel(const float* __restrict__ img, float* __restrict__ out, int C, long long hw) {
  MDM_PDL_ENTER();
  // out[c] += sum over n, hw of img[n][c][:]
  __shared__ float red[32];
  const int c = blockIdx.y % C, n = blockIdx.y / C;
  const float* p = img + ((long long)n * C + c) * hw;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) s += p[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out + c, s);
}

// =============================================================================================
// First / last convolution as tensor-core GEMMs.  The 3x3 neighbourhood of the (1..4)-channel image is
// gathered into a [pixel][64] bf16 matrix (column tap*C + c, zero padded): the first conv is then a 1x1
// GEMM [pix][64] x W'[cout][64], its wgrad the matching wgrad GEMM; the last conv is a 1x1 GEMM producing
// the 9*C per-tap partial outputs z[pix][32] (fp32) followed by the 9-tap scatter-sum below, its backward
// the same gather (flipped) feeding a dgrad-as-fprop GEMM and a wgrad GEMM.  (csrc/igemm.cu does the GEMMs.)
// =============================================================================================
// thread <-> (pixel, 8-column chunk); only the ceil(9C/8) chunks that hold data are written: the caller zero-fills
// the [pixel][64] matrix ONCE (its columns >= 8*ceil(9C/8) never change).  C is a template parameter: every
// division below has a constant divisor.
template <int C>
__global__ void im2col3x3_kernel(const float* __restrict__ img, bf16* __restrict__ out, int H, int W, int flip,
                                 long long total) {
  MDM_PDL_ENTER();
  constexpr int NCH = (9 * C + 7) / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (pixel, chunk)
  if (i >= total) return;
  long long p = i / NCH;
  const int chunk = (int)(i - p * NCH);
  const long long pix = p;
  const int w = (int)(p % W); p /= W;
  const int h = (int)(p % H); const long long n = p / H;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int j = chunk * 8 + e;
    float v = 0.f;
    if (j < 9 * C) {
      const int tap = j / C, c = j - tap * C;
      const int r = tap / 3, q = tap - r * 3;
      const int dh = flip ? 1 - r : r - 1, dw = flip ? 1 - q : q - 1;
      const int hh = h + dh, ww = w + dw;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(img + ((n * C + c) * H + hh) * W + ww);
    }
    f[e] = v;
  }
  *reinterpret_cast<uint4*>(out + pix * 64 + chunk * 8) = pack8(f);
}

// out[n][c][h][w] = bias[c] + sum_tap z[(n, h+dh, w+dw)][tap*C + c]   (z: fp32 [N*H*W][32])
// CTA <-> a 16 x 16 pixel tile of one image: the 18 x 18 z rows it needs are staged in shared memory with fully
// coalesced float4 loads (a row = 128 contiguous bytes), then thread <-> pixel sums its nine taps from the tile.
constexpr int TS_T = 16;
__global__ void __launch_bounds__(256) tapsum3x3_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                                                        float* __restrict__ out, int C, int H, int W) {
  MDM_PDL_ENTER();
  extern __shared__ float4 ts_sm[];   // [(TS_T+2)^2][8] float4, chunk index XOR-swizzled by the pixel (bank spread)
  const int tiles_w = (W + TS_T - 1) / TS_T, tiles_h = (H + TS_T - 1) / TS_T;
  int t = blockIdx.x;
  const int tw = t % tiles_w; t /= tiles_w;
  const int th = t % tiles_h; const long long n = t / tiles_h;
  const int h0 = th * TS_T, w0 = tw * TS_T;
  constexpr int HP = TS_T + 2;
  for (int i = threadIdx.x; i < HP * HP * 8; i += blockDim.x) {
    const int q = i & 7, pp = i >> 3;
    const int hh = h0 + pp / HP - 1, ww = w0 + pp % HP - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(reinterpret_cast<const float4*>(z + ((n * H + hh) * W + ww) * 32) + q);
    ts_sm[pp * 8 + (q ^ (pp & 7))] = v;
  }
  __syncthreads();
  const int lh = threadIdx.x / TS_T, lw = threadIdx.x % TS_T;
  const int h = h0 + lh, w = w0 + lw;
  if (h >= H || w >= W) return;
  const float* sm = reinterpret_cast<const float*>(ts_sm);
  float acc[PL_MAXC];
#pragma unroll
  for (int c = 0; c < PL_MAXC; ++c) acc[c] = (bias && c < C) ? bias[c] : 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int pp = (lh + tap / 3) * HP + lw + tap % 3;    // (+1 halo, -1 tap offset)
#pragma unroll
    for (int c = 0; c < PL_MAXC; ++c)
      if (c < C) {
        const int col = tap * C + c;
        acc[c] += sm[(pp * 8 + ((col >> 2) ^ (pp & 7))) * 4 + (col & 3)];
      }
  }
  for (int c = 0; c < C; ++c) out[((n * C + c) * H + h) * W + w] = acc[c];
}

// =============================================================================================
// nearest 2x upsample (fwd), its adjoint (sum of the 2x2 block), zero insertion (stride-2 dgrad)
// =============================================================================================
__global__ void upsample2x_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                  int H, int W, int C, long long total_vec) {
  MDM_PDL_ENTER();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int V = C / 8;
  const int cv = (int)(i % V);
  long long p = i / V;  // output pixel (n, ho, wo)
  const int wo = (int)(p % (2 * W)); p /= (2 * W);
  const int ho = (int)(p % (2 * H)); const long long n = p / (2 * H);
  const uint4 v = *reinterpret_cast<const uint4*>(x + ((n * H + ho / 2) * W + wo / 2) * ldx + cv * 8);
  *reinterpret_cast<uint4*>(y + ((n * 2 * H + ho) * 2 * W + wo) * ldy + cv * 8) = v;
}

__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dy, long long ldy, bf16* __restrict__ dx, long long ldx,
                                      int H, int W, int C, long long total_vec) {
  MDM_PDL_ENTER();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int V = C / 4;
  const int cv = (int)(i % V);
  long long p = i / V;  // input pixel (n, h, w)
  const int w = (int)(p % W); p /= W;
  const int h = (int)(p % H); const long long n = p / H;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const float4 v = ld4(dy + ((n * 2 * H + 2 * h + a) * 2 * W + 2 * w + b) * ldy + cv * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  st4(dx + ((n * H + h) * W + w) * ldx + cv * 4, s);
}

__global__ void zero_insert2x_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                     int H, int W, int C, long long total_vec) {
  MDM_PDL_ENTER();
  // y[n][2h][2w] = x[n][h][w], zero elsewhere; y is [N][2H][2W][C]
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int V = C / 8;
  const int cv = (int)(i % V);
  long long p = i / V;
  const int wo = (int)(p % (2 * W)); p /= (2 * W);
  const int ho = (int)(p % (2 * H)); const long long n = p / (2 * H);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (((ho | wo) & 1) == 0) v = *reinterpret_cast<const uint4*>(x + ((n * H + ho / 2) * W + wo / 2) * ldx + cv * 8);
  *reinterpret_cast<uint4*>(y + ((n * 2 * H + ho) * 2 * W + wo) * ldy + cv * 8) = v;
}

// parity weights of the fused nearest-2x upsample + 3x3 convolution (igemm.cu: up2x): out[2a+b][co][2u+v][ci] = sum of the
// 3x3 taps (r, s) of w[co][3r+s][ci] whose upsampled position lands on low-resolution offset (a-1+u, b-1+v)
__global__ void up2x_weights_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cout, int cin) {
  MDM_PDL_ENTER();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (co, ci)
  if (i >= (long long)cout * cin) return;
  const int co = (int)(i / cin), ci = (int)(i % cin);
  float t[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) t[k] = w[((long long)co * 9 + k) * cin + ci];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          // rows: a = 0 -> u = 0: {0}, u = 1: {1, 2};  a = 1 -> u = 0: {0, 1}, u = 1: {2}   (columns alike)
          const int r0 = a == 0 ? (u == 0 ? 0 : 1) : (u == 0 ? 0 : 2), r1 = a == 0 ? (u == 0 ? 0 : 2) : (u == 0 ? 1 : 2);
          const int s0 = b == 0 ? (v == 0 ? 0 : 1) : (v == 0 ? 0 : 2), s1 = b == 0 ? (v == 0 ? 0 : 2) : (v == 0 ? 1 : 2);
          float acc = 0.f;
          for (int r = r0; r <= r1; ++r)
            for (int s_ = s0; s_ <= s1; ++s_) acc += t[r * 3 + s_];
          out[(((long long)(2 * a + b) * cout + co) * 4 + (2 * u + v)) * cin + ci] = __float2bfloat16(acc);
        }
}

// =============================================================================================
// K4: attention core for L <= 1024 tokens, head_dim 8 (heads = C/8).  qkv: [N*L][3C] bf16.
// (The q/k/v/out projections, >99% of the attention FLOPs, are tcgen05 GEMMs in igemm.cu; the
// L x L x 8 core is far too small for a tensor-core tile.)
// =============================================================================================
constexpr int ATT_D = 8;
// Forward: one CTA = HPC adjacent heads of one sample (their q / k / v slices are contiguous: HPC x 16 bytes per token, so
// every global access is a full 16-byte vector of a >= 64-byte segment); K and V of those heads live in shared memory as
// fp32 [token][head][8] (a warp reads one 128-byte line per key, broadcast inside a head).  One thread owns QPT
// consecutive queries of one head: every K / V vector it reads from shared memory feeds QPT queries (the first version,
// one query per thread, was bound by shared-memory reads: ncu short-scoreboard / MIO stalls, 31 % FMA pipe), KC scores
// per query in registers, so the soft-max costs ONE exp2 per key (scale * log2 e folded into q) and one rescale per KC
// keys.  fp32 FMA bound: 32 L^2 flops per (sample, head).
template <int KC, int QPT>
__global__ void __launch_bounds__(256, 2) attention_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int L, int C, int HPC,
                                                            float scale_log2e) {
  MDM_PDL_ENTER();
  extern __shared__ __align__(16) float att_sm[];  // K[L][HPC][8], V[L][HPC][8]
  float* Ks = att_sm;
  float* Vs = att_sm + (size_t)L * HPC * ATT_D;
  const int h0 = blockIdx.x * HPC, n = blockIdx.y;
  const bf16* base = qkv + (long long)n * L * 3 * C + h0 * ATT_D;
  for (int i = threadIdx.x; i < L * HPC; i += blockDim.x) {        // i = token * HPC + head: 16-byte pieces, coalesced
    const int j = i / HPC, h = i - j * HPC;
    const bf16* row = base + (long long)j * 3 * C + h * ATT_D;
    const uint4 kq = *reinterpret_cast<const uint4*>(row + C);
    const uint4 vq = *reinterpret_cast<const uint4*>(row + 2 * C);
    const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kq);
    const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vq);
    float4 a, b;
    float2 t0 = __bfloat1622float2(k2[0]), t1 = __bfloat1622float2(k2[1]), t2 = __bfloat1622float2(k2[2]), t3 = __bfloat1622float2(k2[3]);
    a = make_float4(t0.x, t0.y, t1.x, t1.y); b = make_float4(t2.x, t2.y, t3.x, t3.y);
    *reinterpret_cast<float4*>(Ks + (size_t)i * ATT_D) = a;
    *reinterpret_cast<float4*>(Ks + (size_t)i * ATT_D + 4) = b;
    t0 = __bfloat1622float2(v2[0]); t1 = __bfloat1622float2(v2[1]); t2 = __bfloat1622float2(v2[2]); t3 = __bfloat1622float2(v2[3]);
    a = make_float4(t0.x, t0.y, t1.x, t1.y); b = make_float4(t2.x, t2.y, t3.x, t3.y);
    *reinterpret_cast<float4*>(Vs + (size_t)i * ATT_D) = a;
    *reinterpret_cast<float4*>(Vs + (size_t)i * ATT_D + 4) = b;
  }
  __syncthreads();
  const int groups = L / QPT;                                        // L % QPT == 0 (checked by the host)
  for (int r = threadIdx.x; r < groups * HPC; r += blockDim.x) {     // r = query group * HPC + head
    const int g = r / HPC, h = r - g * HPC;
    const int i0 = g * QPT;
    float q[QPT][ATT_D], acc[QPT][ATT_D], m[QPT], l[QPT];
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      const uint4 qq = *reinterpret_cast<const uint4*>(base + (long long)(i0 + a) * 3 * C + h * ATT_D);
      const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qq);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float2 t = __bfloat1622float2(q2[d]);
        q[a][2 * d] = t.x * scale_log2e;
        q[a][2 * d + 1] = t.y * scale_log2e;
      }
      m[a] = -INFINITY;
      l[a] = 0.f;
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) acc[a][d] = 0.f;
    }
    const float* Kh = Ks + h * ATT_D;
    const float* Vh = Vs + h * ATT_D;
    const int pitch = HPC * ATT_D;
    for (int j0 = 0; j0 < L; j0 += KC) {
      float s[QPT][KC], cm[QPT];
#pragma unroll
      for (int a = 0; a < QPT; ++a) cm[a] = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < KC; ++jj) {
        const float4 ka = *reinterpret_cast<const float4*>(Kh + (size_t)(j0 + jj) * pitch);
        const float4 kb = *reinterpret_cast<const float4*>(Kh + (size_t)(j0 + jj) * pitch + 4);
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          float v = q[a][0] * ka.x;
          v = fmaf(q[a][1], ka.y, v); v = fmaf(q[a][2], ka.z, v); v = fmaf(q[a][3], ka.w, v);
          v = fmaf(q[a][4], kb.x, v); v = fmaf(q[a][5], kb.y, v); v = fmaf(q[a][6], kb.z, v); v = fmaf(q[a][7], kb.w, v);
          s[a][jj] = v;
          cm[a] = fmaxf(cm[a], v);
        }
      }
      float mn[QPT];
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        mn[a] = fmaxf(m[a], cm[a]);
        const float corr = exp2f(m[a] - mn[a]);       // first chunk: exp2(-inf) = 0
        l[a] *= corr;
#pragma unroll
        for (int d = 0; d < ATT_D; ++d) acc[a][d] *= corr;
        m[a] = mn[a];
      }
#pragma unroll
      for (int jj = 0; jj < KC; ++jj) {
        const float4 va = *reinterpret_cast<const float4*>(Vh + (size_t)(j0 + jj) * pitch);
        const float4 vb = *reinterpret_cast<const float4*>(Vh + (size_t)(j0 + jj) * pitch + 4);
#pragma unroll
        for (int a = 0; a < QPT; ++a) {
          const float p = exp2f(s[a][jj] - mn[a]);
          l[a] += p;
          acc[a][0] = fmaf(p, va.x, acc[a][0]); acc[a][1] = fmaf(p, va.y, acc[a][1]); acc[a][2] = fmaf(p, va.z, acc[a][2]); acc[a][3] = fmaf(p, va.w, acc[a][3]);
          acc[a][4] = fmaf(p, vb.x, acc[a][4]); acc[a][5] = fmaf(p, vb.y, acc[a][5]); acc[a][6] = fmaf(p, vb.z, acc[a][6]); acc[a][7] = fmaf(p, vb.w, acc[a][7]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      const float inv = 1.0f / l[a];
      uint4 o;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int d = 0; d < 4; ++d) o2[d] = __floats2bfloat162_rn(acc[a][2 * d] * inv, acc[a][2 * d + 1] * inv);
      *reinterpret_cast<uint4*>(out + ((long long)n * L + i0 + a) * C + (h0 + h) * ATT_D) = o;
    }
  }
}

// Backward of the attention core, no atomics.  One CTA = HPC adjacent heads of one sample; q, k, v, dO of the CTA's heads
// live in shared memory as fp32 [token][head][8] (a warp reads 32-byte rows, the threads of a warp that share a head
// read the same address).  A thread owns QPT (token, head) rows -- tokens i, i + L / QPT, ... -- so every staged row it
// reads feeds QPT rows of arithmetic (the first version, one row per thread, waited on shared memory: ncu short-scoreboard
// stall 14.9 per issue).  Phase A (rows as QUERIES): one pass for the softmax statistics and the output row
// (D_i = dO_i . O_i), one pass for dq_i = sum_j ds_ij k_j with ds_ij = p_ij (dO_i . v_j - D_i) scale.  Phase B (rows as
// KEYS): dk_j = sum_i ds_ij q_i and dv_j = sum_i p_ij dO_i from the staged (m_i, 1 / l_i, D_i).  72 FMAs + 3 exp2 per
// (query, key, head); round 1's kernel did 16 shared-memory float atomics (CAS loops) per pair.
template <int QPT>
__global__ void __launch_bounds__(1024 / QPT) attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                    bf16* __restrict__ dqkv, int L, int C, int HPC, float scale) {
  MDM_PDL_ENTER();
  extern __shared__ float att_sm[];  // K, V, Q, dO: T x 8 each; statistics: T x 4
  const int T = L * HPC;             // rows of this CTA; blockDim.x == T / QPT
  float* Ks = att_sm;
  float* Vs = Ks + (size_t)T * ATT_D;
  float* Qs = Vs + (size_t)T * ATT_D;
  float* Gs = Qs + (size_t)T * ATT_D;
  float4* St = reinterpret_cast<float4*>(Gs + (size_t)T * ATT_D);
  const int t = threadIdx.x;
  const int h = t % HPC, i0 = t / HPC, istep = L / QPT;   // this thread's tokens: i0 + a * istep
  const int n = blockIdx.y;
  const long long col = (long long)(blockIdx.x * HPC + h) * ATT_D;
  float q[QPT][ATT_D], g[QPT][ATT_D];
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    const int i = i0 + a * istep;
    const int r = i * HPC + h;
    const bf16* row = qkv + ((long long)n * L + i) * 3 * C + col;
    float k[ATT_D], v[ATT_D];
    unpack8(ldg16(row), q[a]);
    unpack8(ldg16(row + C), k);
    unpack8(ldg16(row + 2 * C), v);
    unpack8(ldg16(dout + ((long long)n * L + i) * C + col), g[a]);
    float4* d;
    d = reinterpret_cast<float4*>(Ks + (size_t)r * ATT_D); d[0] = make_float4(k[0], k[1], k[2], k[3]); d[1] = make_float4(k[4], k[5], k[6], k[7]);
    d = reinterpret_cast<float4*>(Vs + (size_t)r * ATT_D); d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
    d = reinterpret_cast<float4*>(Qs + (size_t)r * ATT_D); d[0] = make_float4(q[a][0], q[a][1], q[a][2], q[a][3]); d[1] = make_float4(q[a][4], q[a][5], q[a][6], q[a][7]);
    d = reinterpret_cast<float4*>(Gs + (size_t)r * ATT_D); d[0] = make_float4(g[a][0], g[a][1], g[a][2], g[a][3]); d[1] = make_float4(g[a][4], g[a][5], g[a][6], g[a][7]);
  }
  __syncthreads();
  const float sl2 = scale * 1.4426950408889634f;   // scores in the exp2 domain
  auto dot8 = [](const float (&a)[ATT_D], const float4& x, const float4& y) {   // two chains of four
    float r0 = a[0] * x.x, r1 = a[4] * y.x;
    r0 = fmaf(a[1], x.y, r0); r1 = fmaf(a[5], y.y, r1);
    r0 = fmaf(a[2], x.z, r0); r1 = fmaf(a[6], y.z, r1);
    r0 = fmaf(a[3], x.w, r0); r1 = fmaf(a[7], y.w, r1);
    return r0 + r1;
  };
  auto axpy8 = [](float (&acc)[ATT_D], float w, const float4& x, const float4& y) {
    acc[0] = fmaf(w, x.x, acc[0]); acc[1] = fmaf(w, x.y, acc[1]); acc[2] = fmaf(w, x.z, acc[2]); acc[3] = fmaf(w, x.w, acc[3]);
    acc[4] = fmaf(w, y.x, acc[4]); acc[5] = fmaf(w, y.y, acc[5]); acc[6] = fmaf(w, y.z, acc[6]); acc[7] = fmaf(w, y.w, acc[7]);
  };
  const float4* K4 = reinterpret_cast<const float4*>(Ks) + 2 * h;   // row j of this head: K4[2 j HPC], K4[2 j HPC + 1]
  const float4* V4 = reinterpret_cast<const float4*>(Vs) + 2 * h;
  const float4* Q4 = reinterpret_cast<const float4*>(Qs) + 2 * h;
  const float4* G4 = reinterpret_cast<const float4*>(Gs) + 2 * h;
  const int rs = 2 * HPC;              // float4 stride between tokens
  // ---- phase A1: softmax statistics and D_i
  float m[QPT], inv[QPT], D[QPT];
#pragma unroll
  for (int a = 0; a < QPT; ++a) {
    m[a] = -INFINITY;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) q[a][d] *= sl2;
  }
  {
    float l[QPT], o[QPT][ATT_D];
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      l[a] = 0.f;
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) o[a][d] = 0.f;
    }
#pragma unroll 2
    for (int j = 0; j < L; ++j) {
      const float4 ka = K4[j * rs], kb = K4[j * rs + 1], va = V4[j * rs], vb = V4[j * rs + 1];
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        const float sc = dot8(q[a], ka, kb);
        if (sc > m[a]) {               // rare after the first keys
          const float corr = exp2f(m[a] - sc);
          l[a] *= corr;
#pragma unroll
          for (int d = 0; d < ATT_D; ++d) o[a][d] *= corr;
          m[a] = sc;
        }
        const float pj = exp2f(sc - m[a]);
        l[a] += pj;
        axpy8(o[a], pj, va, vb);
      }
    }
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      float dd = 0.f;
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) dd = fmaf(g[a][d], o[a][d], dd);
      inv[a] = 1.0f / l[a];
      D[a] = dd * inv[a];
      St[(i0 + a * istep) * HPC + h] = make_float4(m[a], inv[a], D[a], 0.f);
    }
  }
  // ---- phase A2: dq
  {
    float dq[QPT][ATT_D];
#pragma unroll
    for (int a = 0; a < QPT; ++a)
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) dq[a][d] = 0.f;
#pragma unroll 2
    for (int j = 0; j < L; ++j) {
      const float4 ka = K4[j * rs], kb = K4[j * rs + 1], va = V4[j * rs], vb = V4[j * rs + 1];
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        const float pj = exp2f(dot8(q[a], ka, kb) - m[a]) * inv[a];
        const float dp = dot8(g[a], va, vb);
        axpy8(dq[a], pj * (dp - D[a]) * scale, ka, kb);
      }
    }
#pragma unroll
    for (int a = 0; a < QPT; ++a)
      *reinterpret_cast<uint4*>(dqkv + ((long long)n * L + i0 + a * istep) * 3 * C + col) = pack8(dq[a]);
  }
  __syncthreads();                     // every row's statistics are staged
  // ---- phase B: this thread's tokens as KEY rows (q / g registers are reused for k * sl2 / v)
  {
    float dk[QPT][ATT_D], dv[QPT][ATT_D];
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      const int i = i0 + a * istep;
      const float4 ka = K4[i * rs], kb = K4[i * rs + 1], va = V4[i * rs], vb = V4[i * rs + 1];
      q[a][0] = ka.x * sl2; q[a][1] = ka.y * sl2; q[a][2] = ka.z * sl2; q[a][3] = ka.w * sl2;
      q[a][4] = kb.x * sl2; q[a][5] = kb.y * sl2; q[a][6] = kb.z * sl2; q[a][7] = kb.w * sl2;
      g[a][0] = va.x; g[a][1] = va.y; g[a][2] = va.z; g[a][3] = va.w; g[a][4] = vb.x; g[a][5] = vb.y; g[a][6] = vb.z; g[a][7] = vb.w;
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) dk[a][d] = dv[a][d] = 0.f;
    }
#pragma unroll 2
    for (int ii = 0; ii < L; ++ii) {
      const float4 qa = Q4[ii * rs], qb = Q4[ii * rs + 1], ga = G4[ii * rs], gb = G4[ii * rs + 1];
      const float4 st = St[ii * HPC + h];
#pragma unroll
      for (int a = 0; a < QPT; ++a) {
        const float pj = exp2f(dot8(q[a], qa, qb) - st.x) * st.y;
        const float dp = dot8(g[a], ga, gb);
        axpy8(dk[a], pj * (dp - st.z) * scale, qa, qb);
        axpy8(dv[a], pj, ga, gb);
      }
    }
#pragma unroll
    for (int a = 0; a < QPT; ++a) {
      bf16* o = dqkv + ((long long)n * L + i0 + a * istep) * 3 * C + col;
      *reinterpret_cast<uint4*>(o + C) = pack8(dk[a]);
      *reinterpret_cast<uint4*>(o + 2 * C) = pack8(dv[a]);
    }
  }
}

// =============================================================================================
// small elementwise / reductions
// =============================================================================================
// diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): [cos(t f_k), sin(t f_k)], f_k = exp(-ln(1e4) k / half)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ out, int N, int dim) {
  MDM_PDL_ENTER();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * dim) return;
  const int n = i / dim, k = i % dim, half = dim / 2;
  const int kk = k < half ? k : k - half;
  const float f = expf(-9.210340371976184f * (float)kk / (float)half);
  const float a = t[n] * f;
  out[i] = __float2bfloat16(k < half ? cosf(a) : sinf(a));
}

// y = silu(x) elementwise; x fp32 [rows][C] -> y bf16  (time-embedding MLP; tiny)
__global__ void silu_f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  MDM_PDL_ENTER();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16(siluf_(x[i]));
}
// dx = dy * silu'(x): dy fp32, x fp32 -> dx bf16
__global__ void silu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, bf16* __restrict__ dx, long long n) {
  MDM_PDL_ENTER();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = __float2bfloat16(dy[i] * dsiluf_(x[i]));
}

// out[c] += sum over rows of dy[row][c]   (bias gradients).  thread <-> 8 channels (16-byte loads), CTA <-> a
// run of rows, shared-memory reduction across the CTA's row threads, one global atomic per channel per CTA.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, long long ld, float* __restrict__ out,
                                                     float* __restrict__ out2, long long rows, int C, int L, int R,
                                                     int rows_per_cta) {
  MDM_PDL_ENTER();
  extern __shared__ float cs_sm[];   // [Cb]: this CTA's column block (blockIdx.y, up to 2048 channels)
  const int c_base = blockIdx.y * 2048;
  const int Cb = min(2048, C - c_base);
  out += c_base;
  if (out2) out2 += c_base;
  C = Cb;
  for (int c = threadIdx.x; c < C; c += blockDim.x) cs_sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x % L, r = threadIdx.x / L;
  if (r < R && lane * 8 < Cb) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(r0 + (long long)rows_per_cta, rows);
    const bf16* base = dy + c_base + lane * 8;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (long long p = r0 + r; p < r1; p += (long long)R * 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pp = p + (long long)u * R;
        v[u] = pp < r1 ? ldg16(base + pp * ld) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&cs_sm[lane * 8 + j], s[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = cs_sm[c];
    atomicAdd(out + c, v);
    if (out2) atomicAdd(out2 + c, v);
  }
}

// out[n][c] (fp32, ld_out) = sum over the HW pixels of sample n of dy[(n*HW+p)][c]; optional dbias[c] += same
__global__ void sample_colsum_kernel(const bf16* __restrict__ dy, long long ld, float* __restrict__ out, long long ld_out,
                                     float* __restrict__ dbias, int HW, int C) {
  MDM_PDL_ENTER();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (c >= C) return;
  float s = 0.f;
  const bf16* p = dy + (long long)n * HW * ld + c;
  for (int i = 0; i < HW; ++i) s += __bfloat162float(p[(long long)i * ld]);
  out[(long long)n * ld_out + c] = s;
  if (dbias) atomicAdd(dbias + c, s);
}

// fused residual + MSE (trainer_masked.py:126-140 / trainer_masked_mean_shift.py:142-159):
//   recon = (x_in + net) - shift ; loss = mean(w_b * (recon - x0)^2) ; dnet = 2 w_b (recon - x0) / numel
__global__ void mse_residual_kernel(const float* __restrict__ x_in, const float* __restrict__ net,
                                    const float* __restrict__ shift, const float* __restrict__ x0,
                                    const float* __restrict__ weight, float* __restrict__ dnet,
                                    float* __restrict__ recon_out, float* __restrict__ partial, long long per_sample,
                                    long long total, float inv_total) {
  MDM_PDL_ENTER();
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float r = __fadd_rn(x_in[i], net[i]);
    if (shift) r = __fsub_rn(r, shift[i]);
    const float d = r - x0[i];
    const float w = weight ? weight[i / per_sample] : 1.0f;
    s += w * d * d;
    if (dnet) dnet[i] = 2.0f * w * d * inv_total;
    if (recon_out) recon_out[i] = r;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void mse_finalize_kernel(const float* __restrict__ partial, int n, float inv_total, float* __restrict__ loss) {
  MDM_PDL_ENTER();
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *loss = s * inv_total;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  MDM_PDL_ENTER();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    st4(y + i, v);
  } else {
    for (long long j = i; j < n; ++j) y[j] = __float2bfloat16(x[j]);
  }
}

}  // namespace mdm

using namespace mdm;

#define GRID1D(n, t) (unsigned)(((n) + (t) - 1) / (t))

extern "C" {

int64_t mdm_gn_ws_floats(int N, int HW, int C, int G) {
  GnGeom g; int nc;
  if (gn_geom(g, HW, C, G, &nc, N, gn_slots(false))) return 0;
  int nc2 = nc;
  if (gn_geom(g, HW, C, G, &nc2, N, gn_slots(true))) return 0;       // forward and backward families chunk differently
  if (nc2 > nc) nc = nc2;
  return (int64_t)N * nc * GN_MAX_GROUPS * 2;
}

int mdm_gn_silu_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                    float* stats, float* ws, int N, int HW, int C, int G, float eps, int silu, void* stream) {
  MDM_CHECK_ARG(x && y && gamma && beta && ws, "gn_silu_fwd: NULL pointer");
  MDM_CHECK_ARG(ld_x % 8 == 0 && ld_y % 8 == 0, "gn_silu_fwd: channel strides must be multiples of 8");
  MDM_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), "gn_silu_fwd: pointers must be 16-byte aligned");
  GnGeom g; int nc;
  int rc = gn_geom(g, HW, C, G, &nc, N, gn_slots(false));
  if (rc) return rc;
  if ((long long)HW * C <= 32768) {   // small map: one CTA per sample, single pass over registers
    const int per_thread = (HW + g.R - 1) / g.R;
    const size_t sm = (size_t)2 * C * sizeof(float);
    if (per_thread <= 4) launch_pdl(gn_small_fwd_kernel<4>, dim3(N), dim3(GN_THREADS), sm, as_stream(stream), (const bf16*)x, ld_x, (bf16*)y, ld_y, gamma, beta, stats, eps, silu, g);
    else launch_pdl(gn_small_fwd_kernel<16>, dim3(N), dim3(GN_THREADS), sm, as_stream(stream), (const bf16*)x, ld_x, (bf16*)y, ld_y, gamma, beta, stats, eps, silu, g);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  if (const int K = gn_cluster_size(g, GN_CL_FWD_V)) {   // medium map: one cluster per sample, one pass
    launch_cluster(gn_cluster_fwd_kernel<GN_CL_FWD_V>, N * K, K, gn_cluster_smem(C, false), as_stream(stream), (const bf16*)x, ld_x, (bf16*)y, ld_y, gamma, beta, stats, eps, silu, g, K);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  dim3 grid(nc, N);
  launch_pdl(gn_stats_kernel<8>, dim3(grid), dim3(GN_THREADS), (size_t)2 * C * sizeof(float), as_stream(stream), (const bf16*)x, ld_x, ws, g);
  MDM_LAUNCH_CHECK();
  launch_pdl(gn_apply_kernel<4>, dim3(grid), dim3(GN_THREADS), 0, as_stream(stream), (const bf16*)x, ld_x, (bf16*)y, ld_y, gamma, beta, ws, stats, eps, silu, g, (const float*)nullptr, 0, (const float*)nullptr);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_gn_silu_fwd_q(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                      float* stats, const float* qa, int qa_quads, const float* qb, int N, int HW, int C, int G,
                      float eps, int silu, void* stream) {
  MDM_CHECK_ARG(x && y && gamma && beta && qa, "gn_silu_fwd_q: NULL pointer");
  MDM_CHECK_ARG(ld_x % 8 == 0 && ld_y % 8 == 0, "gn_silu_fwd_q: channel strides must be multiples of 8");
  GnGeom g; int nc;
  int rc = gn_geom(g, HW, C, G, &nc, N, gn_slots(false));
  if (rc) return rc;
  MDM_CHECK_ARG(g.cpg % 4 == 0, "gn_silu_fwd_q: C/G = %d must be a multiple of 4 (quad sums)", g.cpg);
  MDM_CHECK_ARG(qa_quads > 0 && qa_quads <= C / 4 && (qa_quads == C / 4 || qb != nullptr), "gn_silu_fwd_q: bad quad split %d of %d", qa_quads, C / 4);
  dim3 grid(nc, N);
  launch_pdl(gn_apply_kernel<4>, dim3(grid), dim3(GN_THREADS), 0, as_stream(stream), (const bf16*)x, ld_x, (bf16*)y, ld_y, gamma, beta,
             (const float*)nullptr, stats, eps, silu, g, qa, qa_quads, qb);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_gn_coef_q(const float* qa, int qa_quads, const float* qb, const float* gamma, const float* beta, float* coef,
                  float* stats, int N, int HW, int C, int G, float eps, void* stream) {
  MDM_CHECK_ARG(qa && gamma && beta && coef && N > 0 && G > 0 && G <= GN_MAX_GROUPS && C % G == 0 && (C / G) % 4 == 0,
                "gn_coef_q: bad arguments (C=%d G=%d)", C, G);
  MDM_CHECK_ARG(qa_quads > 0 && qa_quads <= C / 4 && (qa_quads == C / 4 || qb != nullptr), "gn_coef_q: bad quad split %d of %d", qa_quads, C / 4);
  launch_pdl(gn_coef_q_kernel, dim3(N), dim3(256), 0, as_stream(stream), qa, qa_quads, qb, gamma, beta, coef, stats, HW, C, G, eps);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_gn_fwd_kind(int N, int HW, int C, int G) {
  GnGeom g; int nc;
  if (gn_geom(g, HW, C, G, &nc, N, gn_slots(false))) return -1;
  if ((long long)HW * C <= 32768) return 0;
  return gn_cluster_size(g, GN_CL_FWD_V) ? 1 : 2;
}

int mdm_gn_silu_bwd(const void* x, long long ld_x, const void* dy, long long ld_dy, const void* add, long long ld_add,
                    const void* add2, long long ld_add2, void* dx, long long ld_dx, const float* gamma, const float* beta,
                    const float* stats, float* dgamma, float* dbeta, float* ws, float* colsum, long long ld_colsum,
                    float* dbias, int N, int HW, int C, int G, int silu, void* stream) {
  MDM_CHECK_ARG(x && dy && dx && gamma && beta && stats && ws, "gn_silu_bwd: NULL pointer");
  MDM_CHECK_ARG(ld_x % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0 && ld_add % 8 == 0 && ld_add2 % 8 == 0,
                "gn_silu_bwd: channel strides must be multiples of 8");
  GnGeom g; int nc;
  int rc = gn_geom(g, HW, C, G, &nc, N, gn_slots(true));
  if (rc) return rc;
  if ((long long)HW * C <= 16384) {   // small map: one CTA per sample, x and dy stay in registers across both phases
    const int per_thread = (HW + g.R - 1) / g.R;
    const size_t sm = (size_t)2 * C * sizeof(float);
    if (per_thread <= 2) launch_pdl(gn_small_bwd_kernel<2>, dim3(N), dim3(GN_THREADS), sm, as_stream(stream), (const bf16*)x, ld_x, (const bf16*)dy, ld_dy, (const bf16*)add, ld_add, (const bf16*)add2, ld_add2, (bf16*)dx, ld_dx, gamma, beta, stats, dgamma, dbeta, silu, colsum, ld_colsum, dbias, g);
    else launch_pdl(gn_small_bwd_kernel<8>, dim3(N), dim3(GN_THREADS), sm, as_stream(stream), (const bf16*)x, ld_x, (const bf16*)dy, ld_dy, (const bf16*)add, ld_add, (const bf16*)add2, ld_add2, (bf16*)dx, ld_dx, gamma, beta, stats, dgamma, dbeta, silu, colsum, ld_colsum, dbias, g);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  // measured (bench_gn.py, B200): the single-pass cluster backward is SLOWER than the two-pass kernels (its second phase
  // waits on the add / add2 loads with few warps per SM): opt-in until those loads are staged too
  const char* clb = getenv("MDM_GN_CLUSTER_BWD");
  const int K_bwd = (clb && atoi(clb)) ? gn_cluster_size(g, GN_CL_BWD_V) : 0;
  if (const int K = K_bwd) {   // medium map: one cluster per sample, x and dy read once
    launch_cluster(gn_cluster_bwd_kernel<GN_CL_BWD_V>, N * K, K, gn_cluster_smem(C, true), as_stream(stream), (const bf16*)x, ld_x, (const bf16*)dy, ld_dy, (const bf16*)add, ld_add, (const bf16*)add2, ld_add2, (bf16*)dx, ld_dx, gamma, beta, stats, dgamma, dbeta, silu, colsum, ld_colsum, dbias, g, K);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  dim3 grid(nc, N);
  const size_t sm1 = (size_t)2 * C * sizeof(float), sm2 = (size_t)C * sizeof(float);
  launch_pdl(gn_bwd_stats_kernel<4>, dim3(grid), dim3(GN_THREADS), sm1, as_stream(stream), (const bf16*)x, ld_x, (const bf16*)dy, ld_dy, gamma, beta, stats, ws, dgamma, dbeta, silu, g);
  MDM_LAUNCH_CHECK();
#define GN_BWD_APPLY(A1, A2)                                                                                              \
  launch_pdl(gn_bwd_apply_kernel<2, A1, A2>, dim3(grid), dim3(GN_THREADS), sm2, as_stream(stream),                                            \
      (const bf16*)x, ld_x, (const bf16*)dy, ld_dy, (const bf16*)add, ld_add, (const bf16*)add2, ld_add2, (bf16*)dx, ld_dx, \
      gamma, beta, stats, ws, silu, colsum, ld_colsum, dbias, g)
  if (add && add2) GN_BWD_APPLY(true, true);
  else if (add) GN_BWD_APPLY(true, false);
  else if (add2) GN_BWD_APPLY(false, true);
  else GN_BWD_APPLY(false, false);
#undef GN_BWD_APPLY
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_conv_in_fwd(const float* img, const float* w, const float* bias, void* y, long long ld_y, int N, int C, int H,
                    int W, int cout, void* stream) {
  MDM_CHECK_ARG(img && w && y, "conv_in_fwd: NULL pointer");
  MDM_CHECK_ARG(C >= 1 && C <= PL_MAXC, "conv_in_fwd: image channels must be 1..4 (got %d)", C);
  dim3 grid(H * ((W + PL_TW - 1) / PL_TW), N);
  launch_pdl(planar_to_nhwc_conv_kernel<0>, dim3(grid), dim3(128), 0, as_stream(stream), img, w, bias, (bf16*)y, ld_y, C, H, W, cout, 0);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_conv_in_wgrad(const float* img, const void* dy, long long ld_dy, float* dw, float* dbias, int N, int C, int H,
                      int W, int cout, void* stream) {
  MDM_CHECK_ARG(img && dy && dw, "conv_in_wgrad: NULL pointer");
  MDM_CHECK_ARG(C >= 1 && C <= PL_MAXC && cout <= 1024, "conv_in_wgrad: C in 1..4, cout <= 1024");
  const int jobs = H * ((W + PL_TW - 1) / PL_TW);
  const int rows_per_cta = 8;
  dim3 grid((jobs + rows_per_cta - 1) / rows_per_cta, N);
  const int th = ((cout + 31) / 32) * 32;
  launch_pdl(planar_wgrad_kernel<0>, dim3(grid), dim3(th), 0, as_stream(stream), img, (const bf16*)dy, ld_dy, dw, dbias, C, H, W, cout, 0, rows_per_cta);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_conv_out_fwd(const void* x, long long ld_x, const float* w, const float* bias, float* y, int N, int C, int H,
                     int W, int cin, void* stream) {
  MDM_CHECK_ARG(x && w && y, "conv_out_fwd: NULL pointer");
  MDM_CHECK_ARG(C >= 1 && C <= PL_MAXC && cin % 8 == 0, "conv_out_fwd: C in 1..4, cin %% 8");
  const size_t smem = (size_t)9 * cin * 4 * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    MDM_CUDA(cudaFuncSetAttribute(nhwc_to_planar_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid(GRID1D((long long)H * W, 128), N);
  launch_pdl(nhwc_to_planar_conv_kernel, dim3(grid), dim3(128), smem, as_stream(stream), (const bf16*)x, ld_x, w, bias, y, C, H, W, cin);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

/* backward of the last conv: dx (NHWC bf16) = dgrad(dy planar fp32); dw[C][cin][3][3], dbias[C] accumulated */
int mdm_conv_out_bwd(const void* x, long long ld_x, const float* w, const float* dy, void* dx, long long ld_dx, float* dw,
                     float* dbias, int N, int C, int H, int W, int cin, void* stream) {
  MDM_CHECK_ARG(x && w && dy, "conv_out_bwd: NULL pointer");
  MDM_CHECK_ARG(C >= 1 && C <= PL_MAXC && cin <= 1024, "conv_out_bwd: C in 1..4, cin <= 1024");
  cudaStream_t st = as_stream(stream);
  if (dx) {
    dim3 grid(H * ((W + PL_TW - 1) / PL_TW), N);
    launch_pdl(planar_to_nhwc_conv_kernel<1>, dim3(grid), dim3(128), 0, st, dy, w, nullptr, (bf16*)dx, ld_dx, C, H, W, cin, 1);
    MDM_LAUNCH_CHECK();
  }
  if (dw) {
    const int jobs = H * ((W + PL_TW - 1) / PL_TW);
    const int rows_per_cta = 8;
    dim3 grid((jobs + rows_per_cta - 1) / rows_per_cta, N);
    const int th = ((cin + 31) / 32) * 32;
    launch_pdl(planar_wgrad_kernel<1>, dim3(grid), dim3(th), 0, st, dy, (const bf16*)x, ld_x, dw, nullptr, C, H, W, cin, 1, rows_per_cta);
    MDM_LAUNCH_CHECK();
  }
  if (dbias) {
    dim3 grid(8, N * C);
    launch_pdl(planar_sum_kernel, dim3(grid), dim3(256), 0, st, dy, dbias, C, (long long)H * W);
    MDM_LAUNCH_CHECK();
  }
  return MDM_OK;
}

int mdm_im2col3x3(const float* img, void* out, int N, int C, int H, int W, int flip, void* stream) {
  MDM_CHECK_ARG(img && out && C >= 1 && C <= PL_MAXC, "im2col3x3: bad arguments (C=%d)", C);
  const long long total = (long long)N * H * W * ((9 * C + 7) / 8);
#define MDM_IM2COL(CC) launch_pdl(im2col3x3_kernel<CC>, dim3(GRID1D(total, 256)), dim3(256), 0, as_stream(stream), img, (bf16*)out, H, W, flip, total)
  switch (C) {
    case 1: MDM_IM2COL(1); break;
    case 2: MDM_IM2COL(2); break;
    case 3: MDM_IM2COL(3); break;
    default: MDM_IM2COL(4); break;
  }
#undef MDM_IM2COL
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_tapsum3x3(const float* z, const float* bias, float* out, int N, int C, int H, int W, void* stream) {
  MDM_CHECK_ARG(z && out && C >= 1 && C <= 3, "tapsum3x3: bad arguments (C=%d; 9*C must fit 32 columns)", C);
  const long long tiles = (long long)N * ((H + TS_T - 1) / TS_T) * ((W + TS_T - 1) / TS_T);
  const size_t sm = (size_t)(TS_T + 2) * (TS_T + 2) * 8 * sizeof(float4);   // 41472 B
  launch_pdl(tapsum3x3_kernel, dim3((unsigned)tiles), dim3(256), sm, as_stream(stream), z, bias, out, C, H, W);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_upsample2x_fwd(const void* x, long long ld_x, void* y, long long ld_y, int N, int H, int W, int C, void* stream) {
  MDM_CHECK_ARG(x && y && C % 8 == 0, "upsample2x_fwd: bad arguments");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  launch_pdl(upsample2x_kernel, dim3(GRID1D(total, 256)), dim3(256), 0, as_stream(stream), (const bf16*)x, ld_x, (bf16*)y, ld_y, H, W, C, total);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_upsample2x_bwd(const void* dy, long long ld_dy, void* dx, long long ld_dx, int N, int H, int W, int C, void* stream) {
  MDM_CHECK_ARG(dy && dx && C % 4 == 0, "upsample2x_bwd: bad arguments");
  const long long total = (long long)N * H * W * (C / 4);
  launch_pdl(upsample2x_bwd_kernel, dim3(GRID1D(total, 256)), dim3(256), 0, as_stream(stream), (const bf16*)dy, ld_dy, (bf16*)dx, ld_dx, H, W, C, total);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_zero_insert2x(const void* x, long long ld_x, void* y, long long ld_y, int N, int H, int W, int C, void* stream) {
  MDM_CHECK_ARG(x && y && C % 8 == 0, "zero_insert2x: bad arguments");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  launch_pdl(zero_insert2x_kernel, dim3(GRID1D(total, 256)), dim3(256), 0, as_stream(stream), (const bf16*)x, ld_x, (bf16*)y, ld_y, H, W, C, total);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_up2x_weights(const float* w32, void* out_bf16, int cout, int cin, void* stream) {
  MDM_CHECK_ARG(w32 && out_bf16 && cout > 0 && cin > 0, "up2x_weights: bad arguments");
  const long long total = (long long)cout * cin;
  launch_pdl(up2x_weights_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, as_stream(stream), w32, (bf16*)out_bf16, cout, cin);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_attention_fwd(const void* qkv, void* out, int N, int L, int C, void* stream) {
  MDM_CHECK_ARG(qkv && out && C % ATT_D == 0 && L >= 1 && L <= 1024, "attention_fwd: bad arguments (L=%d C=%d)", L, C);
  MDM_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "attention_fwd: pointers must be 16-byte aligned");
  const int heads = C / ATT_D;
  // queries per thread: 4 when L allows; heads per CTA: 256 (query group, head) rows per CTA when the sample has that
  // many, at least 4 heads (64-byte segments), a divisor of the head count, K + V within 64 KB of shared memory
  const int qpt = (L % 4 == 0) ? 4 : 1;
  int hpc = 256 * qpt / L;
  if (hpc < 4) hpc = 4;
  if (hpc > heads) hpc = heads;
  while (hpc > 1 && (heads % hpc != 0 || (size_t)2 * L * hpc * ATT_D * sizeof(float) > 64 * 1024)) hpc >>= 1;
  if (heads % hpc != 0) hpc = 1;
  const size_t smem = (size_t)2 * L * hpc * ATT_D * sizeof(float);
  MDM_CHECK_ARG(smem <= 96 * 1024, "attention_fwd: K + V of one head do not fit shared memory (L=%d)", L);
  static bool attr_set = false;
  if (!attr_set) {
    MDM_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MDM_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MDM_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  int th = (L / qpt) * hpc;
  th = th < 32 ? 32 : (th > 256 ? 256 : ((th + 31) / 32) * 32);
  dim3 grid(heads / hpc, N);
  const float sl2 = 1.4426950408889634f / sqrtf((float)ATT_D);
  cudaStream_t st = as_stream(stream);
  if (qpt == 4 && L % 8 == 0) launch_pdl(attention_fwd_kernel<8, 4>, dim3(grid), dim3(th), smem, st, (const bf16*)qkv, (bf16*)out, L, C, hpc, sl2);
  else if (qpt == 4) launch_pdl(attention_fwd_kernel<4, 4>, dim3(grid), dim3(th), smem, st, (const bf16*)qkv, (bf16*)out, L, C, hpc, sl2);
  else launch_pdl(attention_fwd_kernel<1, 1>, dim3(grid), dim3(th), smem, st, (const bf16*)qkv, (bf16*)out, L, C, hpc, sl2);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_attention_bwd(const void* qkv, const void* dout, void* dqkv, int N, int L, int C, void* stream) {
  MDM_CHECK_ARG(qkv && dout && dqkv && C % ATT_D == 0 && L >= 1 && L <= 1024, "attention_bwd: bad arguments");
  MDM_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)dout % 16 == 0) && ((uintptr_t)dqkv % 16 == 0), "attention_bwd: pointers must be 16-byte aligned");
  const int heads = C / ATT_D;
  // heads per CTA: one thread per (token, head), at most 1024; few heads per CTA on long sequences (the rows of a warp
  // that share a head share their shared-memory reads), many on short ones (enough threads per CTA)
  int hpc = 1024 / L;
  const int want = L >= 32 ? 8 : 64;
  if (hpc > want) hpc = want;
  if (hpc > heads) hpc = heads;
  if (hpc < 1) hpc = 1;
  while (heads % hpc != 0) --hpc;
  const int rows = L * hpc;
  const int qpt = (L % 2 == 0 && rows >= 128) ? 2 : 1;    // rows per thread
  const size_t smem = (size_t)rows * (4 * ATT_D + 4) * sizeof(float);
  static size_t smem_set[2] = {0, 0};
  if (smem > 48 * 1024 && smem > smem_set[qpt - 1]) {
    if (qpt == 2) MDM_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else MDM_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[qpt - 1] = smem;
  }
  dim3 grid(heads / hpc, N);
  const float sc = 1.0f / sqrtf((float)ATT_D);
  if (qpt == 2) launch_pdl(attention_bwd_kernel<2>, dim3(grid), dim3(rows / 2), smem, as_stream(stream), (const bf16*)qkv, (const bf16*)dout, (bf16*)dqkv, L, C, hpc, sc);
  else launch_pdl(attention_bwd_kernel<1>, dim3(grid), dim3(rows), smem, as_stream(stream), (const bf16*)qkv, (const bf16*)dout, (bf16*)dqkv, L, C, hpc, sc);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_timestep_embedding(const float* t, void* out, int N, int dim, void* stream) {
  MDM_CHECK_ARG(t && out && dim % 2 == 0, "timestep_embedding: bad arguments");
  launch_pdl(timestep_embedding_kernel, dim3(GRID1D(N * dim, 256)), dim3(256), 0, as_stream(stream), t, (bf16*)out, N, dim);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_silu_fwd(const float* x, void* y, int64_t n, void* stream) {
  MDM_CHECK_ARG(x && y, "silu_fwd: NULL pointer");
  launch_pdl(silu_f32_to_bf16_kernel, dim3(GRID1D(n, 256)), dim3(256), 0, as_stream(stream), x, (bf16*)y, n);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_silu_bwd(const float* x, const float* dy, void* dx, int64_t n, void* stream) {
  MDM_CHECK_ARG(x && dy && dx, "silu_bwd: NULL pointer");
  launch_pdl(silu_bwd_kernel, dim3(GRID1D(n, 256)), dim3(256), 0, as_stream(stream), x, dy, (bf16*)dx, n);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_colsum(const void* dy, long long ld, float* out, float* out2, int64_t rows, int C, void* stream) {
  MDM_CHECK_ARG(dy && out, "colsum: NULL pointer");
  MDM_CHECK_ARG(C % 8 == 0 && ld % 8 == 0 && ((uintptr_t)dy % 16 == 0), "colsum: C %% 8, ld %% 8, 16-byte aligned (C=%d ld=%lld)", C, ld);
  const int Cb = C < 2048 ? C : 2048;
  const int L = Cb / 8, R = 256 / L;
  // ~64 KB of bf16 per CTA, at least one pass of R rows
  long long rpc = (32768 + Cb - 1) / Cb;
  rpc = ((rpc + R - 1) / R) * R;
  if (rpc > rows) rpc = rows;
  const dim3 blocks((unsigned)((rows + rpc - 1) / rpc), (unsigned)((C + 2047) / 2048));
  launch_pdl(colsum_kernel, dim3(blocks), dim3(256), (size_t)Cb * sizeof(float), as_stream(stream), (const bf16*)dy, ld, out, out2, rows, C, L, R, (int)rpc);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_sample_colsum(const void* dy, long long ld, float* out, long long ld_out, float* dbias, int N, int HW, int C, void* stream) {
  MDM_CHECK_ARG(dy && out, "sample_colsum: NULL pointer");
  dim3 grid(GRID1D(C, 128), N);
  launch_pdl(sample_colsum_kernel, dim3(grid), dim3(128), 0, as_stream(stream), (const bf16*)dy, ld, out, ld_out, dbias, HW, C);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_mse_residual(const float* x_in, const float* net, const float* shift, const float* x0, const float* weight,
                     float* dnet, float* recon, float* loss, float* ws, int64_t per_sample, int64_t total, void* stream) {
  MDM_CHECK_ARG(x_in && net && x0 && loss && ws, "mse_residual: NULL pointer");
  const int blocks = (int)((total + 256 * 8 - 1) / (256 * 8) < 1024 ? (total + 256 * 8 - 1) / (256 * 8) : 1024);
  const float inv = 1.0f / (float)total;
  launch_pdl(mse_residual_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x_in, net, shift, x0, weight, dnet, recon, ws, per_sample, total, inv);
  MDM_LAUNCH_CHECK();
  launch_pdl(mse_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), ws, blocks, inv, loss);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
  MDM_CHECK_ARG(x && y, "cast: NULL pointer");
  launch_pdl(f32_to_bf16_kernel, dim3(GRID1D((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), x, (bf16*)y, n);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

}  // extern "C"
