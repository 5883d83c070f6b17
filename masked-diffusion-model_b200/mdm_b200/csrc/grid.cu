// Visual / evaluation side on the device: image grids built where the samples live, instead of
// the reference's host round trip through `normalize01(_global)` + `torchvision.utils.make_grid`
// (sampler.py:369-417; utils/datautils.py:211-229).
//
//   normalisation 0: none;  1 ("image"): per image (x - min) / (max - min), NaN -> 0;  2 ("global"): one min / max for
//   the whole batch, no NaN handling -- the reference's two helpers, op for op in fp32.
//   grid: torchvision's layout -- nrow images per row, `pad` pixels of `pad_value` around every cell, single-channel
//   images replicated to three channels.
#include <float.h>

#include "common.cuh"

namespace mdm {

// per-image (min, max); grid = images, one CTA each
__global__ void __launch_bounds__(256) image_minmax_kernel(const float* __restrict__ x, int64_t per_image, float* __restrict__ mm) {
  MDM_PDL_ENTER();
  __shared__ float s_lo[8], s_hi[8];
  const float* p = x + (int64_t)blockIdx.x * per_image;
  float lo = INFINITY, hi = -INFINITY;
  bool nan = false;
  for (int64_t i = threadIdx.x; i < per_image; i += blockDim.x) {
    const float v = p[i];
    nan |= isnan(v);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  nan = __syncthreads_or(nan);                 // torch.amax / amin propagate NaN
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
    mm[2 * blockIdx.x] = nan ? NAN : lo;
    mm[2 * blockIdx.x + 1] = nan ? NAN : hi;
  }
}

// one thread per grid pixel and channel
__global__ void image_grid_kernel(const float* __restrict__ x, const float* __restrict__ mm, int norm, int B, int C, int H, int W,
                                  int nrow, int pad, float pad_value, float* __restrict__ out, int Cg, int Hg, int Wg) {
  MDM_PDL_ENTER();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Cg * Hg * Wg) return;
  const int gx = (int)(i % Wg), gy = (int)((i / Wg) % Hg), c = (int)(i / ((int64_t)Wg * Hg));
  const int cw = W + pad, ch = H + pad;
  const int cx = gx / cw, cy = gy / ch, px = gx % cw - pad, py = gy % ch - pad;
  const int k = cy * nrow + cx;
  float v = pad_value;
  if (px >= 0 && py >= 0 && cx < nrow && k < B && gx < nrow * cw && gy < ((B + nrow - 1) / nrow) * ch) {
    const float s = x[(((int64_t)k * C + (C == 1 ? 0 : c)) * H + py) * W + px];
    if (norm == 0) {
      v = s;
    } else {
      float lo, hi;
      if (norm == 1) { lo = mm[2 * k]; hi = mm[2 * k + 1]; }
      else {   // global: reduce the per-image pairs (B is small)
        lo = mm[0]; hi = mm[1];
        for (int b = 1; b < B; ++b) { lo = fminf(lo, mm[2 * b]); hi = fmaxf(hi, mm[2 * b + 1]); if (isnan(mm[2 * b])) lo = hi = NAN; }
      }
      v = __fdiv_rn(__fsub_rn(s, lo), __fsub_rn(hi, lo));
      if (norm == 1 && isnan(v)) v = 0.0f;      // normalize01: nan_to_num(nan=0); normalize01_global has none
    }
  }
  out[i] = v;
}

}  // namespace mdm

using namespace mdm;

extern "C" {

int mdm_image_grid(const float* imgs, int batch, int channels, int H, int W, int nrow, int pad, float pad_value,
                   int normalization, float* minmax_ws /*2 * batch floats*/, float* out, void* stream) {
  MDM_CHECK_ARG(imgs && out && batch >= 1 && channels >= 1 && H >= 1 && W >= 1 && nrow >= 1 && pad >= 0, "image_grid: bad arguments");
  MDM_CHECK_ARG(normalization >= 0 && normalization <= 2 && (normalization == 0 || minmax_ws), "image_grid: normalization 0..2 (needs minmax_ws)");
  const int xmaps = nrow < batch ? nrow : batch;
  const int ymaps = (batch + xmaps - 1) / xmaps;
  const int Cg = channels == 1 ? 3 : channels, Hg = (H + pad) * ymaps + pad, Wg = (W + pad) * xmaps + pad;
  cudaStream_t st = as_stream(stream);
  if (normalization) {
    launch_pdl(image_minmax_kernel, dim3(batch), dim3(256), 0, st, imgs, (int64_t)channels * H * W, minmax_ws);
    MDM_LAUNCH_CHECK();
  }
  const int64_t total = (int64_t)Cg * Hg * Wg;
  launch_pdl(image_grid_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, imgs, (const float*)minmax_ws, normalization,
             batch, channels, H, W, xmaps, pad, pad_value, out, Cg, Hg, Wg);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

}  // extern "C"
