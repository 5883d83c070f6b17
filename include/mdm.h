/* mdm.h -- C ABI of libmdm_sm100.so: the B200 (sm_100a) kernels behind the masked-diffusion
 * hot path of hytae1993/masked-diffusion-model.
 *
 * The reference has no native/FFI boundary: its hot path sits behind
 * Python call signatures.  This header is the *new* native boundary the Python drop-in modules
 * (`masked-diffusion-model_b200/{scheduler,sampler,trainer_masked*}.py`) bind through ctypes.
 * Each group cites the reference code it replaces (paths relative to /root/reference/code).
 *
 * Conventions
 *   - every function returns 0 on success, a negative MDM_E_* code on failure; the message of
 *     the last failure on the calling thread is available from mdm_last_error();
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller (PyTorch)
 *     owns every buffer, including workspaces -- the library never allocates device memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises, so every entry point can be captured into a CUDA graph;
 *   - images are NCHW contiguous (the reference's layout); activations inside the denoiser
 *     are NHWC bf16 with an explicit channel stride `ld` (elements) so that channel slices of
 *     a concat buffer can be passed without copies;
 *   - there is NO CPU fallback: calling a compute entry point without a CUDA device fails.
 */
#ifndef MDM_H_
#define MDM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDM_OK 0
#define MDM_E_ARG (-1)      /* invalid argument */
#define MDM_E_CUDA (-2)     /* CUDA runtime / driver error */
#define MDM_E_UNSUPPORTED (-3)

/* element types of image/activation buffers */
#define MDM_F32 0
#define MDM_BF16 1
#define MDM_U8 2      /* raw image bytes: normalised on load, (u / 255 - 0.5) / 0.5 (utils/mydataset.py:81) */

/* fill value for degraded pixels: scheduler.py:298-317 (`mean_option`) */
#define MDM_FILL_CONST 0            /* float(mean_option)                         */
#define MDM_FILL_DEGRADED_AREA 1    /* mean of img over masked-out pixels         */
#define MDM_FILL_NON_DEGRADED 2     /* -sum(img*m)/sum(1-m) per channel, NaN -> 0 */
/* `mean_area`: scheduler.py:303-309 */
#define MDM_AREA_IMAGE 0
#define MDM_AREA_CHANNEL 1

const char* mdm_last_error(void);
int mdm_version(void);
/* 1 when a CUDA device is usable by this process, 0 otherwise (never raises) */
int mdm_device_available(void);
/* number of kernels this library has launched (or captured into a CUDA graph) in this process */
long long mdm_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Random stream: a device-resident copy of torch's CPU mt19937 engine.
 * Replaces the CPU draws at scheduler.py:282,288,294,434,440,446,496,507,620,658,675,703-707
 * (all `torch.FloatTensor(...).uniform_/normal_` and `torch.randperm` run on the CPU generator
 * in the reference, whatever the model device -- SURVEY.md section 3.2.1 / quirk q15).
 *
 * `rng` points to MDM_RNG_WORDS uint32 on the device: words[0..623] = engine state,
 * words[624] = position inside the current block (624 = regenerate before the next draw).
 * Kernels consume the stream in order and write the advanced state back, so consecutive
 * calls on one stream continue the sequence with no host round trip.
 * ------------------------------------------------------------------------------------------- */
#define MDM_RNG_WORDS 625

/* host helper: init_genrand(seed) == torch.manual_seed(seed) (left_=1 -> position 624) */
int mdm_rng_seed_host(uint32_t* state_host /*[625]*/, uint32_t seed);

/* Parallel stream generation (csrc/mt_jump.cu + csrc/rng.cu).  mt19937 is linear over GF(2): the state J words ahead
 * is g_J(T) s with g_J = x^J mod the characteristic polynomial, so CTA c of a draw can start c * W words into the
 * stream from a precomputed polynomial and the words of one torch-exact stream are produced by many CTAs.
 *   mdm_rng_jump_table_host: host-side precomputation (Berlekamp-Massey for the characteristic polynomial, then
 *     x^{(c * blocks_per_cta - 1) * 624} mod phi for c = 1 .. n_polys; 624 uint32 per polynomial).
 *   mdm_rng_enable_parallel: lends the table (copied to the DEVICE by the caller) to the library.  From then on every
 *     draw of at least MDM_RNG_PAR_MIN_WORDS words is split over ceil(n / (blocks_per_cta * 624)) CTAs and every `rng`
 *     buffer passed to the draw functions must be MDM_RNG_PAR_WORDS uint32 long: words 640 .. 1264 stage the advanced
 *     state (all CTAs read the old state; one tiny follow-up kernel commits the new one).  NULL / 0 disables.
 *   mdm_rng_set_par_stride: pieces of stride * blocks_per_cta blocks per CTA (every stride-th polynomial of the table);
 *     returns the previous value.  Every CTA pays the same jump, so a draw that runs next to other kernels (the masks
 *     of sampler.py:183-199 under the denoiser) costs less SM time with fewer, longer pieces; 1 = shortest draw alone.
 *   mdm_rng_advance_host: host utility, the state after n more draws (same polynomial arithmetic; CPU tests). */
#define MDM_RNG_PAR_WORDS 1280
#define MDM_RNG_PAR_MIN_WORDS 262144
int mdm_rng_jump_table_host(uint32_t* polys_host, int n_polys, int blocks_per_cta);
int mdm_rng_enable_parallel(const uint32_t* polys_dev, int n_polys, int blocks_per_cta);
int mdm_rng_set_par_stride(int stride);
int mdm_rng_advance_host(const uint32_t* state_in_host /*[625]*/, int64_t n, uint32_t* state_out_host /*[625]*/);

/* n raw 32-bit outputs (untransformed) */
int mdm_rng_raw(uint32_t* rng, uint32_t* out, int64_t n, void* stream);
/* advance the stream by n words, nothing written (scheduler.py:703-705: a uniform draw whose
 * value is discarded but whose words are consumed -- quirk q8) */
int mdm_rng_skip(uint32_t* rng, int64_t n, void* stream);
/* FloatTensor(n).uniform_(a,b): float32((w & 0xFFFFFF) * 2^-24) * (b-a) + a */
int mdm_rng_uniform(uint32_t* rng, float* out, int64_t n, float a, float b, void* stream);
/* FloatTensor(n).normal_(mean,std), n % 16 == 0: 16-wide Box-Muller blocks, optionally scaled
 * per sample in float64 (`random * ratio`, scheduler.py:680-684,713-717): out[b, :] =
 * float32(double(normal) * ratio[b]).  ratio may be NULL (no scaling). */
int mdm_rng_normal(uint32_t* rng, float* out, int batch, int64_t per_sample, float mean,
                   float std, const double* ratio, void* stream);
/* thresholding masks, scheduler.py:288-296 / 440-448 / 496-513:
 *   mask[b, j] = (uniform_(0,1)[b, j] > ratio[b])  <=>  (w & 0xFFFFFF) > floor(ratio[b] * 2^24)
 * one byte per element (1 = keep, 0 = degrade).  When ratio2/mask2 are non-NULL the same
 * uniform field is thresholded a second time (degrade_dependent_base_sampling). */
int mdm_rng_threshold_mask(uint32_t* rng, const double* ratio, uint8_t* mask,
                           const double* ratio2, uint8_t* mask2, int batch, int64_t per_sample,
                           void* stream);
/* indexing masks, scheduler.py:279-284 / 431-436: per sample, torch.randperm(HW)[:count[b]]
 * positions are zeroed.  words_ws: workspace of batch*(HW-1) uint32. */
int mdm_rng_randperm_mask(uint32_t* rng, const int64_t* count, uint8_t* mask,
                          uint32_t* words_ws, int batch, int hw, void* stream);
/* torch.randint(lo, hi, (n,)) on the CPU generator: w % (hi-lo) + lo (trainer_masked.py:114
 * when the batch lives on the CPU generator; int64 output) */
int mdm_rng_randint(uint32_t* rng, int64_t* out, int64_t n, int64_t lo, int64_t hi, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1: fill + composite given a mask.  Replaces scheduler.py:298-321 (degrade_training),
 * :450-475 (degrade_independent_base_sampling), :572-598 (degrade_with_mask).
 *   x_t          = (1-m)*fill + m*img            [batch, C, HW] float32
 *   mask_f32     = m as float32                  [batch, mask_ch, HW]     (optional)
 *   degrade_mask = (1-m)*fill + m                [batch, C, HW] float32   (optional)
 *   fill_out     = fill value                    [batch, C] float32       (optional)
 * img: [batch, C, HW] float32 or bf16.  mask: [batch, mask_ch, HW] bytes, mask_ch in {1, C}.
 * ws: workspace of mdm_degrade_ws_floats(batch, C, HW) floats.
 * ------------------------------------------------------------------------------------------- */
int64_t mdm_degrade_ws_floats(int batch, int channels, int hw);
int mdm_degrade(const void* img, int img_dtype, const uint8_t* mask, int mask_ch, int fill_mode,
                float fill_const, int mean_area, float* x_t, float* mask_f32,
                float* degrade_mask, float* fill_out, float* ws, int batch, int channels, int hw,
                void* stream);
/* GPU data-feeding path: the same operator on RAW uint8 images [batch, C, HW].  The normalisation
 * the reference's CPU data pipeline applies -- torchvision ToTensor (u / 255) + Normalize(0.5, 0.5), utils/mydataset.py:81
 * -- is fused into K1's read, op for op in fp32, so x_t is bit-identical to feeding the normalised fp32 image; the
 * normalised image itself is written to x0_out (optional; the loss of the training step reads it).  Host -> device
 * traffic per batch drops 4x. */
int mdm_degrade_u8(const uint8_t* img_u8, const uint8_t* mask, int mask_ch, int fill_mode, float fill_const,
                   int mean_area, float* x_t, float* x0_out, float* mask_f32, float* degrade_mask, float* fill_out,
                   float* ws, int batch, int channels, int hw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5: one restoration-loop update.  Replaces sampler.py:143-152 and :167-216:
 *   s0      = (((x_t + shift) + net) - shift)              (x_t + shift is what the denoiser saw)
 *   D_t     = degrade(s0, mask_t),  D_next = degrade(s0, mask_next)
 *   x_next  = momentum ? x_t + (D_next - D_t) : D_next      (base_momentum / base_sampling)
 *   x_in_next = x_next + shift_next                         (input of the next denoiser call)
 * net: [batch,C,HW] float32 NCHW (denoiser output).  shift / shift_next: float32 with element
 * strides (sb, sc, sp) so (B,1,1,1), (B,3,1,1), (B,1,H,W) and (B,3,H,W) shifts are all
 * addressable; NULL means zero.  `update` = 0 leaves x_t untouched (last iteration, i == 0).
 * s0_out optional.  ws as for mdm_degrade, twice the size.
 * ------------------------------------------------------------------------------------------- */
int mdm_sampler_step(const float* x_t, const float* net,
                     const float* shift, int64_t sb, int64_t sc, int64_t sp,
                     const uint8_t* mask_t, const uint8_t* mask_next, int mask_ch,
                     int fill_mode, float fill_const, int mean_area, int momentum, int update,
                     const float* shift_next, int64_t nb, int64_t nc, int64_t np_,
                     float* x_next, float* x_in_next, float* s0_out, float* ws,
                     int batch, int channels, int hw, void* stream);

/* x_in = x + shift (scheduler.py:757-766 perturb_shift), strides as above */
int mdm_add_shift(const float* x, const float* shift, int64_t sb, int64_t sc, int64_t sp,
                  float* out, int batch, int channels, int hw, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Denoiser (K2/K3/K4): replaces the cuDNN/cuBLAS/ATen calls behind `diffusers.UNet2DModel`
 * (reference utils/model.py:24-32; call sites trainer_masked.py:125, sampler.py:145).
 * Activations: NHWC bf16, channel stride `ld` in elements (multiple of 8) so channel slices of a
 * concat buffer are addressable.  Weights: bf16 packed [cout][k*k][cin] (K-major for fprop).
 * Spatial sizes must be powers of two (32/64/128/256 on this path).
 * ------------------------------------------------------------------------------------------- */
typedef struct mdm_conv_args {
  const void* x;        /* fprop/wgrad: layer input [N][H*s][W*s][cin]; dgrad: dy [N][H][W][cout] */
  long long ld_x;
  int cin;
  const void* w;        /* packed bf16 weights [cout][k*k][cin] */
  void* y;              /* fprop: output [N][H][W][cout]; dgrad: dx [N][H][W][cin]; wgrad: dy (read) */
  long long ld_y;
  int cout;
  int N, H, W;          /* batch and OUTPUT spatial size */
  int ksize, stride;    /* 1|3, 1|2 */
  const float* bias;    /* [cout] or NULL */
  const float* bias2;   /* second bias (fused shortcut), or NULL */
  const float* rowvec;  /* per-sample vector added to every pixel of sample n: [N][ld_rowvec] */
  long long ld_rowvec;
  const void* resid;    /* bf16 residual added in the epilogue, or NULL */
  long long ld_resid;
  int accumulate;       /* add to the existing contents of y */
  float* y_f32;         /* optional fp32 copy of the result, row-major [N*H*W][cout or cin] */
  const void* x2;       /* fprop only: fused 1x1 shortcut segment (second K range) */
  long long ld_x2;
  int cin2;
  const void* w2;       /* bf16 [cout][1][cin2] */
  float* dw;            /* wgrad output fp32 [cout][k*k][cin], accumulated atomically */
  long long w_col0;     /* dgrad/wgrad on a channel slice of a wider packed weight */
  int w_cols;
  float* dbias;         /* wgrad only, optional: dbias[cout] += sum over pixels of dy (the bias gradient), computed by */
  float* dbias2;        /* one extra N=16 MMA against a tile of ones inside the same kernel; dbias2 gets the same sums */
  float* splitk_ws;     /* optional ZEROED fp32 workspace: lets small-M fprop/dgrad layers split K over the SMs */
  long long splitk_ws_floats; /* (needs N*H*W*cout floats; left zeroed again on return) */
  float* qsum;          /* fprop only, optional: GroupNorm statistics of the output fused into the store epilogue:
                         * qsum[n][cout/4][2] += (sum, sum of squares) over the pixels of sample n of every 4-channel
                         * quad (fp32 atomics into a buffer the caller zeroed); needs H*W % 128 == 0.  Consumed by
                         * mdm_gn_silu_fwd_q, which then reads the activation once instead of twice. */
  int up2x;             /* fprop only: nearest-neighbour 2x upsample FUSED into a 3x3 convolution (diffusers Upsample2D:
                         * F.interpolate(scale 2, nearest) + conv; reference equivalent unet6.py:468-470).  The upsampled
                         * image is never built: output pixels of parity (up_a, up_b) = (row % 2, col % 2) see a 2x2
                         * neighbourhood of the LOW-resolution input with pre-summed weights, 4/9 of the FLOPs.
                         * x: [N][H][W][cin] low-resolution input, y: the FULL [N][2H][2W][cout] output (this call writes
                         * one parity class), w: [cout][4][cin] from mdm_up2x_weights for that parity; H, W multiples of 16. */
  int up_a, up_b;
  const float* gn_coef; /* fprop only (inference): GroupNorm + SiLU of the INPUT folded into the operand path.  x is then the
                         * RAW activation and coef[n][cin][2] = (scale, shift) per sample and channel (mdm_gn_coef_q): the
                         * kernel computes silu(x * scale + shift) while it fills its halo tiles, zero outside the map
                         * -- the normalised activation is never written.  3x3, stride 1, H and W multiples of 16, cin a multiple of
                         * 64 and >= 128.  Reference site: models/unet/unet6.py:358-360 (norm -> act -> conv). */
} mdm_conv_args;

/* (scale, shift) table of a GroupNorm site from the quad sums its producers' epilogues accumulated (mdm_conv_args.qsum):
 * coef[n][c] = (gamma_c rstd, beta_c - mean gamma_c rstd); optional stats[n][G][2] = (mean, rstd).  Feeds
 * mdm_conv_args.gn_coef. */
int mdm_gn_coef_q(const float* qa, int qa_quads, const float* qb, const float* gamma, const float* beta, float* coef,
                  float* stats, int N, int HW, int C, int G, float eps, void* stream);
/* parity weights of the fused upsample convolution: w32 fp32 [cout][9][cin] (packed layout) -> out bf16
 * [4 = 2 a + b][cout][4 = 2 u + v][cin], W_ab[u][v] = sum of the 3x3 taps that land on low-resolution offset
 * (a - 1 + u, b - 1 + v): rows {0} | {1, 2} for a = 0, {0, 1} | {2} for a = 1, columns alike. */
int mdm_up2x_weights(const float* w32, void* out_bf16, int cout, int cin, void* stream);
/* tcgen05/TMEM implicit GEMM fed by TMA (csrc/igemm.cu) */
int mdm_conv_fprop(const mdm_conv_args* a, void* stream);
int mdm_conv_dgrad(const mdm_conv_args* a, void* stream);   /* stride-1 layers */
int mdm_conv_wgrad(const mdm_conv_args* a, void* stream);
/* The GEMM kernel is persistent (one CTA per SM, static work list): a caller that runs another long kernel
 * concurrently on a second stream (sampler.py's restoration loop: the serial mt19937 mask / noise generator under
 * the denoiser) reserves n SMs for it, later launches use 148 - n CTAs.  Returns the previous value (n < 0 only
 * queries); process-wide,
 * read at launch time (a captured CUDA graph keeps the grid it was captured with). */
int mdm_reserve_sms(int n);
/* Optional dynamic work distribution of the same kernel (MDM_IGEMM_DYNAMIC=1): the caller lends a ZEROED device
 * buffer of n_ints ints (the current device's; counter pairs are partitioned into 32 per-stream classes so that GEMMs
 * in flight on different streams never share a pair: n_ints >= 128, 4096 recommended); launches whose CTAs run >= 2
 * items then draw their items from an atomic counter in it instead of a static list, so a CTA that shares its SM with
 * another stream's kernel just takes fewer items.  Every launch leaves its counters zeroed.  A 33rd distinct stream
 * falls back to static lists.  NULL / 0 unregisters. */
int mdm_set_sched_workspace(void* zeroed_ints, int n_ints);

/* K2: GroupNorm (+SiLU) forward / backward (csrc/nn_kernels.cu).  x, y, dy, dx: NHWC bf16 with
 * channel strides; stats: [N][G][2] = (mean, rstd) fp32 written by fwd, read by bwd;
 * ws: mdm_gn_ws_floats() floats.  bwd: dx = d/dx [silu](gn(x)) . dy (+ add + add2), dgamma/dbeta
 * accumulated (fp32 atomics); optional colsum[n][c] += sum over the pixels of sample n of the GroupNorm
 * part of dx (row stride ld_colsum) and dbias[c] += the same over all samples -- the gradients of the
 * time-embedding projection and of the bias of the convolution that produced x. */
int64_t mdm_gn_ws_floats(int N, int HW, int C, int G);
int mdm_gn_silu_fwd(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma,
                    const float* beta, float* stats, float* ws, int N, int HW, int C, int G,
                    float eps, int silu, void* stream);
/* forward with precomputed statistics: qa / qb hold the quad sums (see mdm_conv_args.qsum) of channels [0, 4*qa_quads)
 * and [4*qa_quads, C) -- two producers when x is a channel concatenation; qb may be NULL when qa covers all of C.
 * C/G must be a multiple of 4. */
int mdm_gn_silu_fwd_q(const void* x, long long ld_x, void* y, long long ld_y, const float* gamma, const float* beta,
                      float* stats, const float* qa, int qa_quads, const float* qb, int N, int HW, int C, int G,
                      float eps, int silu, void* stream);
/* which forward kernel family mdm_gn_silu_fwd picks: 0 = one CTA per sample, 1 = one cluster per sample,
 * 2 = statistics pass + apply pass (the case fused statistics save a pass) */
int mdm_gn_fwd_kind(int N, int HW, int C, int G);
int mdm_gn_silu_bwd(const void* x, long long ld_x, const void* dy, long long ld_dy, const void* add,
                    long long ld_add, const void* add2, long long ld_add2, void* dx, long long ld_dx, const float* gamma,
                    const float* beta, const float* stats, float* dgamma, float* dbeta, float* ws,
                    float* colsum, long long ld_colsum, float* dbias,
                    int N, int HW, int C, int G, int silu, void* stream);

/* first / last convolution (image channels C <= 4): planar fp32 NCHW image <-> NHWC bf16.
 * w in the checkpoint layout: conv_in [cout][C][3][3], conv_out [C][cin][3][3] (fp32). */
int mdm_conv_in_fwd(const float* img, const float* w, const float* bias, void* y, long long ld_y,
                    int N, int C, int H, int W, int cout, void* stream);
int mdm_conv_in_wgrad(const float* img, const void* dy, long long ld_dy, float* dw, float* dbias,
                      int N, int C, int H, int W, int cout, void* stream);
int mdm_conv_out_fwd(const void* x, long long ld_x, const float* w, const float* bias, float* y,
                     int N, int C, int H, int W, int cin, void* stream);
int mdm_conv_out_bwd(const void* x, long long ld_x, const float* w, const float* dy, void* dx,
                     long long ld_dx, float* dw, float* dbias, int N, int C, int H, int W, int cin,
                     void* stream);

/* first / last convolution as tensor-core GEMMs (round 1b): gather the 3x3 neighbourhood of the C-channel planar fp32
 * image into a [N*H*W][64] bf16 matrix, column tap*C + c (zero padded; flip = taps mirrored, for the adjoint), and
 * the 9-tap scatter-sum that finishes the last conv: out[n][c][h][w] = bias[c] + sum_tap z[(n,h+dh,w+dw)][tap*C+c]
 * with z fp32 [N*H*W][32].  The GEMMs themselves are mdm_conv_fprop / mdm_conv_wgrad with ksize 1.
 * mdm_im2col3x3 writes columns [0, 8*ceil(9C/8)) only: the caller zero-fills the matrix once (the padding columns
 * never change between calls). */
int mdm_im2col3x3(const float* img, void* out_bf16, int N, int C, int H, int W, int flip, void* stream);
int mdm_tapsum3x3(const float* z, const float* bias, float* out, int N, int C, int H, int W, void* stream);

/* nearest 2x upsample, its adjoint, and zero insertion (turns a stride-2 dgrad into a stride-1 one) */
int mdm_upsample2x_fwd(const void* x, long long ld_x, void* y, long long ld_y, int N, int H, int W, int C, void* stream);
int mdm_upsample2x_bwd(const void* dy, long long ld_dy, void* dx, long long ld_dx, int N, int H, int W, int C, void* stream);
int mdm_zero_insert2x(const void* x, long long ld_x, void* y, long long ld_y, int N, int H, int W, int C, void* stream);

/* K4 attention core, head_dim 8, L tokens: qkv [N*L][3C] bf16 -> out [N*L][C] bf16 */
int mdm_attention_fwd(const void* qkv, void* out, int N, int L, int C, void* stream);
int mdm_attention_bwd(const void* qkv, const void* dout, void* dqkv, int N, int L, int C, void* stream);

/* small pieces of the time-embedding path and of the loss */
int mdm_timestep_embedding(const float* t, void* out_bf16, int N, int dim, void* stream);
int mdm_silu_fwd(const float* x, void* y_bf16, int64_t n, void* stream);
int mdm_silu_bwd(const float* x, const float* dy, void* dx_bf16, int64_t n, void* stream);
int mdm_colsum(const void* dy, long long ld, float* out, float* out2, int64_t rows, int C, void* stream);
int mdm_sample_colsum(const void* dy, long long ld, float* out, long long ld_out, float* dbias, int N, int HW, int C, void* stream);
/* recon = (x_in + net) - shift; loss = mean(w_b (recon - x0)^2); dnet = dloss/dnet
 * (trainer_masked.py:126-140, trainer_masked_mean_shift.py:142-159).  ws: 1024 floats. */
int mdm_mse_residual(const float* x_in, const float* net, const float* shift, const float* x0,
                     const float* weight, float* dnet, float* recon, float* loss, float* ws,
                     int64_t per_sample, int64_t total, void* stream);
int mdm_cast_f32_bf16(const float* x, void* y_bf16, int64_t n, void* stream);

/* Fused optimiser tail: global-norm clip + Adam/AdamW/SGD + EMA + bf16 weight mirror over the flat
 * parameter buffer (replaces trainer_masked.py:144-153).
 * mdm_grad_sumsq: out[0] = sum g^2 (ws: 1024 floats).  mdm_adam_ema_step: mode 0 Adam, 1 AdamW,
 * 2 SGD; gnorm_sq (device scalar, may be NULL) feeds torch's clip rule
 * coef = min(1, max_norm / (norm + 1e-6)); grad_scale multiplies g first (1/world for a summed
 * all-reduce); ema / p_bf16 may be NULL. */
int mdm_grad_sumsq(const float* g, int64_t n, float* ws, float* out, void* stream);
/* beta1 / beta2 / bias corrections are DOUBLES: torch derives 1 - beta, lr / bias_c1 and sqrt(bias_c2) from python
 * doubles and rounds once (1.0f - 0.999f is 4.7e-5 away from 0.001f); the kernel uses the same fp32 scalars. */
int mdm_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, int64_t n,
                      const float* gnorm_sq, float lr, double beta1, double beta2, float eps,
                      float weight_decay, double bias_c1, double bias_c2, float max_norm, float ema_decay,
                      float grad_scale, int mode, void* stream);
/* same, with the per-step scalars read from DEVICE memory: hyper[4] = {lr, lr / bias_c1, sqrt(bias_c2),
 * ema_decay} (computed in double on the host, stored as fp32), so the launch can sit inside a replayed CUDA graph
 * while the LR schedule advances. */
int mdm_adam_ema_step_dev(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, int64_t n,
                          const float* gnorm_sq, const float* hyper, double beta1, double beta2, float eps,
                          float weight_decay, float max_norm, float grad_scale, int mode, void* stream);
/* Step statistics without a device synchronisation (replaces the blocking `loss.item()` of trainer_masked.py:160 /
 * trainer_masked_mean_shift.py:172): copies src[0..n) to dst_host_mapped[0..n) -- PINNED HOST memory, device-mapped
 * (cudaHostAlloc under UVA) -- then increments *counter and stores the new count, as int bits, to
 * dst_host_mapped[n].  The host polls that word: the statistics of a step are readable as soon as the forward part of
 * the (captured) step has run, while its backward and optimiser are still in flight.  n <= 31. */
int mdm_publish_stats(const float* src, int n, int* counter, float* dst_host_mapped, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Visual side on the device (csrc/grid.cu): `normalize01` / `normalize01_global`
 * (utils/datautils.py:211-229) + torchvision `make_grid(nrow, padding)` as sampler.py:369-417 composes them.
 * imgs: [batch, C, H, W] fp32; normalization 0 none / 1 per image (NaN -> 0) / 2 global; out: [C' , (H + pad) * rows + pad,
 * (W + pad) * min(nrow, batch) + pad] fp32 with C' = 3 when C == 1 (torchvision replicates grey images).
 * ------------------------------------------------------------------------------------------- */
int mdm_image_grid(const float* imgs, int batch, int channels, int H, int W, int nrow, int pad, float pad_value,
                   int normalization, float* minmax_ws /*2 * batch floats*/, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient all-reduce over NVLink peer memory (csrc/allreduce.cu).  Replaces the NCCL all-reduce that
 * accelerate / DDP run behind `accelerator.backward` (trainer_masked.py:142): ONE kernel per
 * gradient range, capturable inside the training-step graph on a forked stream, 128-thread blocks that co-reside with
 * the persistent GEMM CTAs.  One process per GPU: every rank exports its flat gradient buffer and a zeroed flag array
 * (mdm_p2p_flag_words() uint32) with mdm_ipc_export, exchanges the 64-byte handles + offsets through its process
 * group, and maps the peers' with mdm_ipc_open.  mdm_p2p_allreduce sums buf[offset .. offset + count) over the ranks
 * in rank order (bit-identical on every rank) in place; every rank must issue the same sequence of calls, each rank's
 * calls stream-ordered.  The caller divides by world.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  void* buf[8];    /* flat fp32 gradient buffer of rank p (own or IPC-mapped), p < world */
  void* flag[8];   /* flag array of rank p */
  int rank, world;
} mdm_p2p_comm;
int mdm_p2p_flag_words(void);
int mdm_ipc_export(const void* ptr, void* handle_out /*64 bytes, host*/, int64_t* offset_out /*host*/);
int mdm_ipc_open(const void* handle /*64 bytes, host*/, int64_t offset, void** ptr_out /*host*/);
int mdm_ipc_close(void* ptr /*as returned by mdm_ipc_open*/, int64_t offset /*as passed to it*/);
int mdm_p2p_allreduce(const mdm_p2p_comm* comm /*host*/, int64_t offset, int64_t count, int blocks, void* stream);
/* Copy-engine variant of the same all-reduce, as pieces the caller strings together on its communication stream
 * (mdm_b200/runtime.py: P2PAllReduce.all_reduce_ce): mdm_p2p_barrier(0) -- the peers' gradients are final;
 * (world - 1) x mdm_memcpy_async pulls of this rank's slice from the peers into a local staging area (copy engines, no
 * SM); mdm_reduce_slices sums them in rank order into the local slice; (world - 1) x mdm_memcpy_async pushes of the
 * reduced slice; mdm_p2p_barrier(1) -- every push has landed and nobody reads this rank's buffer any more.  The SMs run
 * two one-warp kernels and one local HBM-rate reduction per range. */
int mdm_p2p_barrier(const mdm_p2p_comm* comm /*host*/, int which, void* stream);
int mdm_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);
int mdm_reduce_slices(float* slice, const float* staging, int64_t stride, int64_t count, int rank, int world, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDM_H_ */
