// Device-resident mt19937 stream that reproduces torch's CPU generator bit for bit.
//
// Replaces the CPU draws of the reference (scheduler.py:282,288,294,434,440,446,496,507,620,
// 658,675,703-707): `torch.FloatTensor(..).uniform_/normal_`, `torch.randperm`.  The word ->
// value transforms are those of ATen's CPU kernels.
//
// mt19937 is one serial recurrence, x[k+624] = x[k+397] ^ twist(x[k], x[k+1]).  Inside a CTA the 624-word block
// lives in shared memory; the usual in-place update needs three dependent phases per block (words 0-226, 227-453,
// 454-623) -- here every new word is written as a function of the OLD block only (substituting the recurrence into
// itself up to three times), so a block costs one barrier: 624 threads each evaluate <= 4 twists, write the new word
// to the other half of a double buffer, temper it and emit it (1.3 words/ns per CTA).
//
// ACROSS CTAs the stream is cut into pieces of W = blocks_per_cta * 624 words: the recurrence is linear over GF(2),
// so CTA c obtains the block that starts (c * W - 624) words ahead as a fixed GF(2) polynomial of the transition
// applied to the current state (csrc/mt_jump.cu: y[J + m] = XOR_{i : g_i = 1} y[i + m]): it generates 33 blocks
// (20592 words) of the sequence into shared memory, XORs the ~10000 shifted copies its polynomial selects (LDS.128
// over a sliding register window, 4 output words per thread), and streams its W words from there.  All CTAs read the
// OLD state; the CTA that owns the last word stages the advanced state behind it and a one-CTA follow-up kernel
// commits it, so consecutive calls on one stream continue the sequence with no host round trip.  Draws below
// MDM_RNG_PAR_MIN_WORDS use one CTA (a jump costs ~0.1 ms).  A 256x1x128x128 threshold mask: 3.3 ms -> ~0.15 ms.
//
// The stream is data independent: the host side launches these kernels on a side stream so they
// overlap the denoiser of the previous step; the data-dependent work (fill + composite, K1/K5)
// runs at HBM rate in degrade.cu from the one-byte-per-pixel masks written here.
#include <math.h>

#include "common.cuh"

namespace mdm {

constexpr int MT_N = 624;
constexpr int MT_THREADS = 640;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// word k of the next block as a function of the current block s[0..623] only
__device__ __forceinline__ uint32_t mt_next_word(const uint32_t* s, int k) {
  if (k < 227) return s[k + 397] ^ mt_twist(s[k], s[k + 1]);
  if (k < 454) return s[k + 170] ^ mt_twist(s[k - 227], s[k - 226]) ^ mt_twist(s[k], s[k + 1]);
  const uint32_t hi = (k == 623) ? (s[397] ^ mt_twist(s[0], s[1])) : s[k + 1];
  return s[k - 57] ^ mt_twist(s[k - 454], s[k - 453]) ^ mt_twist(s[k - 227], s[k - 226]) ^
         mt_twist(s[k], hi);
}

constexpr int MT_DEG = 19937;
constexpr int MT_SEQ_BLOCKS = 33;                 // 33 * 624 = 20592 >= 19968 + 623 words of sequence for the jump
constexpr int MT_STAGE_OFF = 640;                 // rng[640 .. 1264]: staged advanced state (parallel draws)
constexpr int MT_JUMP_GROUPS = 4, MT_JUMP_Q = 156;   // 4 groups x 156 threads x 4 output words = 624 words
constexpr size_t MT_SMEM_SERIAL = 2 * MT_N * sizeof(uint32_t);
constexpr size_t MT_SMEM_PAR = (MT_SEQ_BLOCKS * MT_N + MT_N + MT_JUMP_GROUPS * MT_N) * sizeof(uint32_t);

// One stream, many CTAs.  Window k of the untempered sequence = y[624 k + 1 .. 624 k + 624] (one word past torch's
// block boundaries: y[0] of a freshly seeded state is not a sequence word, every later one is).  CTA c emits the words
// y[pos + c W .. pos + (c + 1) W) of the draw; it starts at window c W / 624 - 1 (c >= 1: from its jump polynomial;
// c = 0: window 0 straight from the state).
template <class Emit>
__global__ void __launch_bounds__(MT_THREADS, 1)
mt_stream_kernel(uint32_t* rng, const uint32_t* __restrict__ polys, int blocks_per_cta, int poly_stride, int64_t n, Emit emit) {
  MDM_PDL_ENTER();
  extern __shared__ __align__(16) uint32_t mt_sm[];
  uint32_t* seq = mt_sm;                           // [33][624] during the jump; windows 0 / 1 afterwards
  const int k = threadIdx.x;
  const int c = blockIdx.x;
  const int64_t W = (int64_t)blocks_per_cta * MT_N;
  const int64_t pos = (int64_t)rng[MT_N];
  const int64_t P = pos + n;                       // stream position after this draw
  const uint32_t y0 = rng[0];
  // window 0: y[1 .. 624] = state words 1 .. 623 and the first word of the next block
  if (k < MT_N) seq[k] = (k < MT_N - 1) ? rng[k + 1] : (rng[397] ^ mt_twist(rng[0], rng[1]));
  __syncthreads();
  int64_t win = 0;                                 // index of the window in st[cur]
  uint32_t* st[2] = {seq, seq + MT_N};
  int cur = 0;
  if (c > 0) {
    uint32_t* poly = mt_sm + MT_SEQ_BLOCKS * MT_N;
    uint32_t* part = poly + MT_N;
    // the table holds x^{(j B - 1) 624} mod phi for j = 1 ..; a CTA that owns poly_stride * B blocks uses every
    // poly_stride-th entry (blocks_per_cta is then poly_stride * B)
    const uint32_t* g = polys + ((size_t)c * poly_stride - 1) * MT_N;
    if (k < MT_N) poly[k] = g[k];
    for (int t = 1; t < MT_SEQ_BLOCKS; ++t) {
      if (k < MT_N) seq[t * MT_N + k] = mt_next_word(seq + (t - 1) * MT_N, k);
      __syncthreads();
    }
    // out[m] = XOR_{i : g_i} seq[i + m]: thread (grp, q) covers polynomial words [grp * 156, grp * 156 + 156) for the
    // four outputs m = 4 q .. 4 q + 3; per polynomial word one 36-word register window (nine LDS.128)
    if (k < MT_JUMP_GROUPS * MT_JUMP_Q) {
      const int grp = k / MT_JUMP_Q, q = k - grp * MT_JUMP_Q;
      uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int wi = grp * MT_JUMP_Q; wi < (grp + 1) * MT_JUMP_Q; ++wi) {
        const uint32_t pw = poly[wi];
        if (pw == 0) continue;
        const uint4* src = reinterpret_cast<const uint4*>(seq + wi * 32 + 4 * q);
        uint32_t r[36];
#pragma unroll
        for (int v = 0; v < 9; ++v) {
          const uint4 u = src[v];
          r[4 * v] = u.x; r[4 * v + 1] = u.y; r[4 * v + 2] = u.z; r[4 * v + 3] = u.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const uint32_t msk = 0u - ((pw >> j) & 1u);
          a0 ^= r[j] & msk; a1 ^= r[j + 1] & msk; a2 ^= r[j + 2] & msk; a3 ^= r[j + 3] & msk;
        }
      }
      *reinterpret_cast<uint4*>(part + grp * MT_N + 4 * q) = make_uint4(a0, a1, a2, a3);
    }
    __syncthreads();
    uint32_t w = 0;
    if (k < MT_N) w = part[k] ^ part[MT_N + k] ^ part[2 * MT_N + k] ^ part[3 * MT_N + k];
    __syncthreads();
    if (k < MT_N) seq[k] = w;
    __syncthreads();
    win = (int64_t)c * blocks_per_cta - 1;
  }
  const int64_t lo = pos + (int64_t)c * W;                       // y-index range [lo, hi) of this CTA
  const int64_t hi = min(lo + W, P);
  const bool last_cta = hi == P;
  if (c == 0 && pos == 0 && k == 0 && n > 0) emit(0, mt_temper(y0));      // y[0] itself (never inside a window)
  // torch block holding the final position: the last CTA runs on until it holds window b_final and its predecessor
  const int64_t b_final = P == 0 ? 0 : (P + MT_N - 1) / MT_N - 1;
  while (true) {
    if (k < MT_N) {
      const int64_t idx = win * MT_N + 1 + k;
      if (idx >= lo && idx < hi) emit(idx - pos, mt_temper(st[cur][k]));
    }
    const bool more_out = (win + 1) * MT_N + 1 < hi;
    const bool more_state = last_cta && win < b_final;
    if (!more_out && !more_state) break;
    if (k < MT_N) st[cur ^ 1][k] = mt_next_word(st[cur], k);
    __syncthreads();
    cur ^= 1;
    ++win;
  }
  if (last_cta) {
    // advanced torch state: key[0] = y[624 b], key[j] = y[624 b + j] = window_b[j - 1]; position P - 624 b.
    // Single-CTA draws write it in place; multi-CTA draws stage it (the other CTAs may not have read rng yet).
    uint32_t* dst = gridDim.x == 1 ? rng : rng + MT_STAGE_OFF;
    __syncthreads();
    if (k < MT_N) {
      uint32_t v;
      if (k == 0) v = b_final == 0 ? y0 : st[cur ^ 1][MT_N - 1];
      else v = st[cur][k - 1];
      dst[k] = v;
    }
    if (k == 0) dst[MT_N] = (uint32_t)(P - b_final * MT_N);
  }
}

__global__ void __launch_bounds__(MT_THREADS, 1) mt_commit_kernel(uint32_t* rng) {
  MDM_PDL_ENTER();
  if (threadIdx.x <= MT_N) rng[threadIdx.x] = rng[MT_STAGE_OFF + threadIdx.x];
}

struct EmitNone {
  __device__ void operator()(int64_t, uint32_t) const {}
};
struct EmitRaw {
  uint32_t* out;
  __device__ void operator()(int64_t i, uint32_t w) const { out[i] = w; }
};
struct EmitUniform {
  float* out;
  float scale, from;
  __device__ void operator()(int64_t i, uint32_t w) const {
    const float u = __fmul_rn((float)(w & 0xFFFFFFu), 5.9604644775390625e-08f);  // * 2^-24, exact
    out[i] = __fadd_rn(__fmul_rn(u, scale), from);
  }
};
struct EmitRandint {
  int64_t* out;
  uint32_t range;
  int64_t lo;
  __device__ void operator()(int64_t i, uint32_t w) const { out[i] = (int64_t)(w % range) + lo; }
};
struct EmitThreshold {
  const double* ratio;
  uint8_t* mask;
  const double* ratio2;
  uint8_t* mask2;
  uint32_t per_sample;
  __device__ void operator()(int64_t i, uint32_t w) const {
    const uint32_t b = (uint32_t)i / per_sample;
    const int64_t k = (int64_t)(w & 0xFFFFFFu);
    // fp32 uniform vs float64 ratio (torch promotes to f64): k*2^-24 > ratio <=> k > floor(ratio*2^24)
    mask[i] = k > (int64_t)floor(ratio[b] * 16777216.0) ? 1 : 0;
    if (mask2) mask2[i] = k > (int64_t)floor(ratio2[b] * 16777216.0) ? 1 : 0;
  }
};

// parallel generation: table lent by the caller (mdm_rng_enable_parallel)
static const uint32_t* g_polys = nullptr;
static int g_npolys = 0, g_blocks_per_cta = 0, g_polys_dev = -1, g_par_stride = 1;

template <class Emit>
struct EmitShifted {   // a later launch of a draw that exceeds the table: indices continue where the previous one stopped
  Emit e;
  int64_t off;
  __device__ void operator()(int64_t i, uint32_t w) const { e(i + off, w); }
};

template <class Emit>
static int launch_stream(uint32_t* rng, int64_t n, Emit e, void* stream) {
  MDM_CHECK_ARG(rng != nullptr, "rng state is NULL");
  MDM_CHECK_ARG(n >= 0, "negative draw count");
  if (n == 0) return MDM_OK;
  int dev = -1;
  const bool par = g_polys != nullptr && n >= MDM_RNG_PAR_MIN_WORDS && cudaGetDevice(&dev) == cudaSuccess && dev == g_polys_dev;
  if (!par) {
    launch_pdl(mt_stream_kernel<Emit>, dim3(1), dim3(MT_THREADS), MT_SMEM_SERIAL, as_stream(stream), rng,
               (const uint32_t*)nullptr, 1 << 24, 1, n, e);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    MDM_CUDA(cudaFuncSetAttribute(mt_stream_kernel<Emit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM_PAR));
    MDM_CUDA(cudaFuncSetAttribute(mt_stream_kernel<EmitShifted<Emit>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM_PAR));
    attr_set = true;
  }
  // CTAs of a launch: every CTA pays the same jump (~0.3 ms of one SM) whatever it then generates, so a draw that
  // runs NEXT TO other kernels (the sampler's masks under the denoiser) is cheaper in SM time with fewer, longer pieces.
  // mdm_rng_set_par_stride(s) / MDM_RNG_PAR_STRIDE = s: pieces of s x blocks_per_cta blocks (default 1: shortest wall
  // time of the draw alone; the sampler's loop asks for 8: 34.77 -> 34.38 ms per c4 denoising step).
  const char* sv = getenv("MDM_RNG_PAR_STRIDE");
  int stride = sv ? atoi(sv) : g_par_stride;
  if (stride < 1) stride = 1;
  if (stride > g_npolys) stride = g_npolys;
  const int bpc = g_blocks_per_cta * stride;
  const int64_t W = (int64_t)bpc * MT_N;
  const int64_t per_launch = W * (g_npolys / stride + 1);       // CTA 0 needs no polynomial
  for (int64_t done = 0; done < n; done += per_launch) {
    const int64_t m = n - done < per_launch ? n - done : per_launch;
    const int ctas = (int)((m + W - 1) / W);
    if (done == 0)
      launch_pdl(mt_stream_kernel<Emit>, dim3(ctas), dim3(MT_THREADS), MT_SMEM_PAR, as_stream(stream), rng, g_polys, bpc, stride, m, e);
    else
      launch_pdl(mt_stream_kernel<EmitShifted<Emit>>, dim3(ctas), dim3(MT_THREADS), MT_SMEM_PAR, as_stream(stream), rng, g_polys,
                 bpc, stride, m, EmitShifted<Emit>{e, done});
    MDM_LAUNCH_CHECK();
    if (ctas > 1) {
      launch_pdl(mt_commit_kernel, dim3(1), dim3(MT_THREADS), 0, as_stream(stream), rng);
      MDM_LAUNCH_CHECK();
    }
  }
  return MDM_OK;
}

// ---- normal_: in-place 16-wide Box-Muller over a buffer of uniforms ---------------------------
__global__ void boxmuller16_kernel(float* data, int64_t npairs, float mean, float std,
                                   const double* ratio, int64_t per_sample) {
  MDM_PDL_ENTER();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const int64_t i0 = (p >> 3) * 16 + (p & 7);
  const float u1 = __fsub_rn(1.0f, data[i0]);
  const float u2 = data[i0 + 8];
  const float radius = sqrtf(__fmul_rn(-2.0f, logf(u1)));
  const float theta = (float)(6.283185307179586476925286766559 * (double)u2);
  float s, c;
  sincosf(theta, &s, &c);
  float a = __fadd_rn(__fmul_rn(__fmul_rn(radius, c), std), mean);
  float b = __fadd_rn(__fmul_rn(__fmul_rn(radius, s), std), mean);
  if (ratio) {  // `random * ratio` with a float64 ratio, then .to(float32)
    const double r0 = ratio[i0 / per_sample];
    const double r1 = ratio[(i0 + 8) / per_sample];
    a = (float)((double)a * r0);
    b = (float)((double)b * r1);
  }
  data[i0] = a;
  data[i0 + 8] = b;
}

// ---- randperm(HW)[:count] -> mask: forward Fisher-Yates, one CTA per sample ---------------------
constexpr int FY_ZCHUNK = 8192;
__global__ void __launch_bounds__(256) fy_mask_kernel(const uint32_t* __restrict__ words,
                                                      const int64_t* __restrict__ count,
                                                      uint8_t* __restrict__ mask, int hw) {
  MDM_PDL_ENTER();
  extern __shared__ uint16_t fy_sm[];
  uint16_t* perm = fy_sm;
  uint16_t* z = fy_sm + hw;
  const int b = blockIdx.x, tid = threadIdx.x;
  int64_t c64 = count[b];
  const int cnt = (int)(c64 < 0 ? 0 : (c64 > hw ? hw : c64));
  uint8_t* m = mask + (int64_t)b * hw;
  for (int i = tid; i < hw; i += blockDim.x) {
    perm[i] = (uint16_t)i;
    m[i] = 1;
  }
  const uint32_t* w = words + (int64_t)b * (hw - 1);
  const int niter = cnt < hw - 1 ? cnt : hw - 1;  // swap i fixes perm[i]; later swaps never touch it
  for (int base = 0; base < niter; base += FY_ZCHUNK) {
    __syncthreads();
    const int mcount = niter - base < FY_ZCHUNK ? niter - base : FY_ZCHUNK;
    for (int j = tid; j < mcount; j += blockDim.x) z[j] = (uint16_t)(w[base + j] % (uint32_t)(hw - (base + j)));
    __syncthreads();
    if (tid == 0) {
      for (int j = 0; j < mcount; ++j) {
        const int i = base + j;
        const int t = i + z[j];
        const uint16_t a = perm[i];
        perm[i] = perm[t];
        perm[t] = a;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < cnt; i += blockDim.x) m[perm[i]] = 0;
}

}  // namespace mdm

using namespace mdm;

extern "C" {

int mdm_rng_seed_host(uint32_t* s, uint32_t seed) {
  if (!s) { set_error("state_host is NULL"); return MDM_E_ARG; }
  s[0] = seed;
  for (int j = 1; j < MT_N; ++j) s[j] = 1812433253u * (s[j - 1] ^ (s[j - 1] >> 30)) + (uint32_t)j;
  s[MT_N] = MT_N;
  return MDM_OK;
}

int mdm_rng_set_par_stride(int stride) {
  const int old = g_par_stride;
  g_par_stride = stride < 1 ? 1 : stride;
  return old;
}

int mdm_rng_enable_parallel(const uint32_t* polys_dev, int n_polys, int blocks_per_cta) {
  if (polys_dev == nullptr || n_polys <= 0) {
    g_polys = nullptr;
    g_npolys = g_blocks_per_cta = 0;
    g_polys_dev = -1;
    return MDM_OK;
  }
  MDM_CHECK_ARG(blocks_per_cta >= 2 && blocks_per_cta <= (1 << 16), "enable_parallel: blocks_per_cta out of range");
  int dev = -1;
  MDM_CUDA(cudaGetDevice(&dev));
  g_polys = polys_dev;
  g_npolys = n_polys;
  g_blocks_per_cta = blocks_per_cta;
  g_polys_dev = dev;
  return MDM_OK;
}

int mdm_rng_raw(uint32_t* rng, uint32_t* out, int64_t n, void* stream) {
  MDM_CHECK_ARG(out != nullptr || n == 0, "out is NULL");
  return launch_stream(rng, n, EmitRaw{out}, stream);
}

int mdm_rng_skip(uint32_t* rng, int64_t n, void* stream) { return launch_stream(rng, n, EmitNone{}, stream); }

int mdm_rng_uniform(uint32_t* rng, float* out, int64_t n, float a, float b, void* stream) {
  MDM_CHECK_ARG(out != nullptr || n == 0, "out is NULL");
  return launch_stream(rng, n, EmitUniform{out, b - a, a}, stream);
}

int mdm_rng_randint(uint32_t* rng, int64_t* out, int64_t n, int64_t lo, int64_t hi, void* stream) {
  MDM_CHECK_ARG(out != nullptr || n == 0, "out is NULL");
  MDM_CHECK_ARG(hi > lo && hi - lo < (int64_t)1 << 32, "randint range must be in (0, 2^32)");
  return launch_stream(rng, n, EmitRandint{out, (uint32_t)(hi - lo), lo}, stream);
}

int mdm_rng_normal(uint32_t* rng, float* out, int batch, int64_t per_sample, float mean, float std,
                   const double* ratio, void* stream) {
  const int64_t n = (int64_t)batch * per_sample;
  MDM_CHECK_ARG(batch > 0 && per_sample > 0, "empty normal draw");
  MDM_CHECK_ARG(n % 16 == 0, "normal_: element count %lld is not a multiple of 16 (tail re-draw path not implemented)", (long long)n);
  int rc = launch_stream(rng, n, EmitUniform{out, 1.0f, 0.0f}, stream);
  if (rc) return rc;
  const int64_t npairs = n / 2;
  const int threads = 256;
  const int64_t blocks = (npairs + threads - 1) / threads;
  launch_pdl(boxmuller16_kernel, dim3((unsigned)blocks), dim3(threads), 0, as_stream(stream), out, npairs, mean, std, ratio, per_sample);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_rng_threshold_mask(uint32_t* rng, const double* ratio, uint8_t* mask, const double* ratio2,
                           uint8_t* mask2, int batch, int64_t per_sample, void* stream) {
  MDM_CHECK_ARG(ratio && mask, "ratio/mask is NULL");
  MDM_CHECK_ARG((ratio2 == nullptr) == (mask2 == nullptr), "ratio2 and mask2 must be given together");
  const int64_t n = (int64_t)batch * per_sample;
  MDM_CHECK_ARG(batch > 0 && per_sample > 0 && n < ((int64_t)1 << 32), "mask size out of range");
  return launch_stream(rng, n, EmitThreshold{ratio, mask, ratio2, mask2, (uint32_t)per_sample}, stream);
}

int mdm_rng_randperm_mask(uint32_t* rng, const int64_t* count, uint8_t* mask, uint32_t* words_ws,
                          int batch, int hw, void* stream) {
  MDM_CHECK_ARG(count && mask && words_ws, "NULL argument");
  MDM_CHECK_ARG(batch > 0 && hw >= 2 && hw <= 65536, "randperm mask supports 2 <= H*W <= 65536 (got %d)", hw);
  int rc = launch_stream(rng, (int64_t)batch * (hw - 1), EmitRaw{words_ws}, stream);
  if (rc) return rc;
  const size_t smem = (size_t)(hw + FY_ZCHUNK) * sizeof(uint16_t);
  static bool attr_set = false;
  if (!attr_set) {
    MDM_CUDA(cudaFuncSetAttribute(fy_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (65536 + FY_ZCHUNK) * 2));
    attr_set = true;
  }
  launch_pdl(fy_mask_kernel, dim3(batch), dim3(256), smem, as_stream(stream), words_ws, count, mask, hw);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

}  // extern "C"
