// Device-resident mt19937 stream that reproduces torch's CPU generator bit for bit.
//
// Replaces the CPU draws of the reference (scheduler.py:282,288,294,434,440,446,496,507,620,
// 658,675,703-707): `torch.FloatTensor(..).uniform_/normal_`, `torch.randperm`.  The word ->
// value transforms are those of ATen's CPU kernels.
//
// mt19937 is one serial recurrence, x[k+624] = x[k+397] ^ twist(x[k], x[k+1]), so the stream is
// produced by ONE CTA that keeps the 624-word block in shared memory.  The usual in-place update
// needs three dependent phases per block (words 0-226, 227-453, 454-623).  Here every new word
// is written as a function of the OLD block only (substituting the recurrence into itself up to
// three times), so a block costs one barrier: 624 threads each evaluate <= 4 twists, write the
// new word to the other half of a double buffer, temper it and emit it.
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

template <class Emit>
__global__ void __launch_bounds__(MT_THREADS, 1) mt_stream_kernel(uint32_t* rng, int64_t n, Emit emit) {
  MDM_PDL_ENTER();
  __shared__ uint32_t st[2][MT_N];
  const int k = threadIdx.x;
  int cur = 0;
  if (k < MT_N) st[0][k] = rng[k];
  int pos = (int)rng[MT_N];
  __syncthreads();
  int64_t done = 0;
  if (pos < MT_N && n > 0) {  // words left in the current block
    const int take = (int)min((int64_t)(MT_N - pos), n);
    if (k >= pos && k < pos + take) emit(done + (k - pos), mt_temper(st[0][k]));
    done += take;
    pos += take;
  }
  while (done < n) {
    uint32_t w = 0;
    if (k < MT_N) {
      w = mt_next_word(st[cur], k);
      st[cur ^ 1][k] = w;
    }
    __syncthreads();
    cur ^= 1;
    const int take = (int)min((int64_t)MT_N, n - done);
    if (k < take) emit(done + k, mt_temper(w));
    done += take;
    pos = take;
  }
  if (k < MT_N) rng[k] = st[cur][k];
  if (k == 0) rng[MT_N] = (uint32_t)pos;
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

template <class Emit>
static int launch_stream(uint32_t* rng, int64_t n, Emit e, void* stream) {
  MDM_CHECK_ARG(rng != nullptr, "rng state is NULL");
  MDM_CHECK_ARG(n >= 0, "negative draw count");
  if (n == 0) return MDM_OK;
  launch_pdl(mt_stream_kernel<Emit>, dim3(1), dim3(MT_THREADS), 0, as_stream(stream), rng, n, e);
  MDM_LAUNCH_CHECK();
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
