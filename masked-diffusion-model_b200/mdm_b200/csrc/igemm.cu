// K3: implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05 + TMEM),
// operands staged by TMA -- fprop, dgrad and wgrad of the denoiser's 3x3 / 1x1 convolutions and
// of every Linear (time embedding, attention projections).  bf16 in, fp32 accumulate.
//
// Replaces the cuDNN / cuBLAS calls behind `diffusers.UNet2DModel` (reference utils/model.py:
// 24-32; layer inventory in SURVEY.md Appendix A).
//
// Activations are NHWC bf16 with an explicit channel stride, so a [128 pixel x 64 channel] A
// tile of the implicit GEMM is ONE tiled-TMA box {64 c, bw, bh, bn} of the 4-D tensor
// (c, w, h, n): the filter tap (r, s) is a coordinate offset, the zero padding is TMA's
// out-of-bounds fill, and a stride-2 convolution is the tensor map's element stride.  The box
// lands in shared memory as 128 rows of 128 bytes with the 128-byte swizzle, which is exactly
// the canonical K-major UMMA operand layout -- no im2col buffer exists anywhere.
//
//   fprop : D[pix][co] = sum_tap sum_ci A[pix+tap][ci] * W[co][tap][ci]      A,B K-major
//   dgrad : D[pix][ci] = sum_tap sum_co dY[pix-tap][co] * W[co][tap][ci]     A K-major, B MN-major
//   wgrad : D[co][ci]  = sum_pix dY[pix][co] * X[pix+tap][ci]   (per tap)   A,B MN-major, split-K
//
// PERSISTENT, warp-specialised kernel: one CTA per SM walks a static list of work items (output tile x K split).
// 224 or 352 threads: warp 0 = TMA producer of the A operand, warp 6 = TMA producer of B, warp 1 = MMA issuer
// (+TMEM allocator), warps 2-5 (and 7-10 when a CTA runs >= 5 items) = epilogue.  The three single-thread roles are
// entered through `elect.sync` with a warp index from `__shfl_sync`, so their loops live in uniform registers (the
// operands of UTCHMMA / UTMALDG are uniform registers; see profiles/README.md section 8).  Three pipelines: a shared-
// memory ring (TMA -> MMA), TWO TMEM accumulator stages (MMA -> epilogue: the epilogue of item i overlaps the main
// loop of item i+1) and two 32 KB staging tiles (epilogue -> TMA store / TMA reduce-add; the residual / accumulate
// operand is TMA-loaded into the same staging tile while the main loop runs).
// Work items: 128 x 128 output tiles; 256 x 128 (two A sub-tiles share every B tile) when enough of them exist; for
// 3x3 stride-1 layers on maps that are multiples of 16 x 16, HALO mode: the item is a 16 x 16 pixel patch whose
// zero-padded 18 x 18 x 64-channel input halo is loaded ONCE per channel chunk and the nine taps are nine shifted
// UMMA descriptors into it (2.3x fewer operand bytes from L2).  wgrad items: one dY tile against two (cin tile, tap)
// entries; the first item of a cout tile carries entry 0 alone plus the bias gradient (one N=16 MMA against ones).
// Epilogues:
//   store : + bias (+ second bias) + per-sample time-embedding vector + residual -> bf16 NHWC TMA store; the kStats
//           instantiations also accumulate GroupNorm quad sums (sum, sum of squares per 4 channels) of the output
//   reduce: fp32 TMA reduce-add into global memory -- wgrad (split over pixels) and split-K partials of
//           small-M fprop/dgrad layers (finished by splitk_finalize_kernel).
#include <cuda.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace mdm {

typedef __nv_bfloat16 bf16_t;
constexpr int TILE_M = 128, TILE_N = 128, TILE_K = 64;
constexpr int A_BYTES = TILE_M * TILE_K * 2;   // 16 KB: one [128 pixel x 64 channel] operand tile
constexpr int B_BYTES = TILE_N * TILE_K * 2;   // 16 KB
// halo mode (3x3 stride-1 fprop / dgrad on maps >= 16 x 8): the M tile is an 8 (w) x 16 (h) pixel patch and
// the A ring holds its zero-padded (16 w x 18 h) x 64 channel input halo, loaded ONCE per channel chunk;
// the nine taps are nine shifted views of it (UMMA descriptor start offsets), a 4x cut of the A traffic.
constexpr int HALO_W = 16, HALO_H = 18, PATCH_W = 8, PATCH_H = 16;
constexpr int HALO_BYTES = HALO_W * HALO_H * 128;   // 36 KB
// halo mode with 256-row work items: the M tile is a 16 x 16 patch = two 8-wide sub-tiles sharing ONE 18 x 18 halo
// and every weight tile: (41.5 KB + 9 x 16 KB) per 18 x 256-cycle MMA groups = 10.3 KB / 256 cycles, under the
// ~42.6 B/cycle/SM the L2 can deliver (plain 256-row stages need 24 KB / 256 cycles: the tensor pipe idles > 50 %)
constexpr int HALO2_W = 18, HALO2_H = 18;
constexpr int HALO2_BYTES = HALO2_W * HALO2_H * 128;   // 41472
constexpr int HALO2_SLOT = 41 * 1024;                   // slot pitch, 1024-aligned (swizzle atom)
constexpr int MAX_A_SLOTS = 4, MAX_B_SLOTS = 6;
constexpr int EPI_BYTES = 32768;               // one staging tile: 128 x 128 bf16, or 2 boxes of 128 x 32 fp32
constexpr int IGEMM_THREADS = 352;             // warps 0: TMA A, 1: MMA, 2-5 (+ 7-10): epilogue, 6: TMA B
constexpr int IGEMM_THREADS_NARROW = 224;      // launch without warps 7-10: four epilogue warps (short grids)
constexpr int NORM_WARPS = 8;                  // kNorm: transform-producer warps of the A operand (warps 11 ..)
constexpr int NORM_THREADS = NORM_WARPS * 32;
constexpr int IGEMM_THREADS_NORM = IGEMM_THREADS + NORM_THREADS;
constexpr int RING_BYTES = 2 * HALO_BYTES + 5 * B_BYTES;    // halo: 2 A + 5 B slots (256-row: 2 x 41 KB + 4 B); plain: 4 A + 4 B slots (128 KB)
static_assert(2 * HALO2_SLOT + 4 * B_BYTES <= RING_BYTES, "256-row halo ring must fit");
constexpr int SMEM_EPI_OFF = RING_BYTES;
constexpr int SMEM_BIAS_OFF = SMEM_EPI_OFF + 2 * EPI_BYTES;   // 128 floats
constexpr int SMEM_QACC_OFF = SMEM_BIAS_OFF + 512;            // [4 epilogue warps][64] floats: per-tile quad (sum, sumsq) partials
constexpr int SMEM_BAR_OFF = SMEM_QACC_OFF + 1024;
constexpr int SMEM_ONES_OFF = 3 * (A_BYTES + 2 * B_BYTES);   // mode 1 only (ring = 3 x 48 KB): the unused ring tail
// dynamic work distribution (kDyn): a ring of (this item, next item) pairs published by the A producer
constexpr int WQ_SLOTS = 4;
constexpr int SMEM_WQ_OFF = SMEM_BAR_OFF + 256;   // wq_full[4], wq_empty[4] (mbarriers), wq[4] (int2)
constexpr int IGEMM_SMEM = SMEM_BAR_OFF + 512 + 1024 /*alignment slack*/;
constexpr int TMEM_COLS = 512;                 // two accumulator stages of 256 fp32 columns
// wgrad bias gradient: D2[co][0..15] = sum_pix dY[pix][co] * 1 lives in the (unused) second half of the stage of a
// single-entry work item (the first item of every cout tile), so both stages stay 256 columns wide
constexpr int ONES_BYTES = 8192;               // [64 K-rows][64 bf16] of 1.0: the B operand of that extra MMA
static_assert(4 * A_BYTES + 4 * B_BYTES <= RING_BYTES && SMEM_ONES_OFF + ONES_BYTES <= RING_BYTES, "plain ring / wgrad ring + ones tile must fit");
static_assert(IGEMM_SMEM <= 232448, "shared memory budget");

// ---- raw PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a pipeline bug traps (launch fails) instead of hanging the GPU.
// Debug build (MDM_NVCC_DEFINES=MDM_IGEMM_DEBUG_WAIT python mdm_b200/build.py): the first wait that times out prints
// which barrier (shared-memory address, see the BARS line) and parity it was waiting for and where the transform warps
// of its CTA stood, then every other wait in the grid gives up, so the kernel ends and the printf buffer is flushed --
// how the two-phase aliasing of the operand ring was found.
#ifdef MDM_IGEMM_DEBUG_WAIT
__device__ int g_wait_abort;
__shared__ volatile int g_dbg[4];      // transform warps: {ring use, stage, work item}
#define DBG(i, v) g_dbg[i] = (v)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*(volatile int*)&g_wait_abort) return;
    if (clock64() - t0 > 400000000LL) {
      printf("WAIT TIMEOUT block %d thread %d bar 0x%x parity %u | transform: use %d stage %d item %d\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity, g_dbg[0], g_dbg[1], g_dbg[2]);
      atomicExch(&g_wait_abort, 1);
      return;
    }
  }
}
#else
#define DBG(i, v)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
#endif
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one lane of the (converged) warp.  With `elect.sync` -- instead of `lane == 0` -- and a warp index the compiler can
// prove uniform, ptxas keeps the single-thread TMA / MMA issue loops in UNIFORM registers: the descriptor operands
// of UTCHMMA / UTMALDG are uniform registers, and from a lane-divergent branch every issue paid a
// vote + ELECT + 5x R2UR waterfall (~15 SASS instructions per MMA, more than the MMA's own 64 cycles)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }   // the epilogue warps

__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major tile [rows][64 bf16]: 8-row groups 1024 B apart; k-th 16-element slice = +32 B
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int k) { return smem_desc(base + k * 32, 16, 1024); }
// K-major view into the halo: 8-pixel row groups one halo row (16 pixels = 2048 B) apart, start shifted by whole
// pixels (128 B).  The 128B swizzle is a function of the absolute shared-memory address bits (measured on
// B200: the shifted views read back exactly what TMA wrote with base_offset = 0), so a shifted start needs
// nothing but the new address.
__device__ __forceinline__ uint64_t desc_halo(uint32_t addr) { return smem_desc(addr, 16, HALO_W * 128); }
// MN-major tile: 64-wide MN atoms 8 KB apart (LBO), 8 K-rows per 1024 B group (SBO); k-th slice = +2 groups
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int k) { return smem_desc(base + k * 2048, 8192, 1024); }

// instruction descriptor: bf16 x bf16 -> f32, M = 128, N = 128
__host__ __device__ constexpr uint32_t make_idesc(int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

struct IgemmArgs {
  int mode;   // 0: activation GEMM (fprop / dgrad), 1: wgrad
  int epi;    // 0: bf16 TMA store (+bias/rowvec/C tile), 1: fp32 TMA reduce-add
  int halo;   // mode 0 only: A = input halo per channel chunk, taps = shifted views; M tile = 8 x 16 patch
  int mt;     // 128-row sub-tiles per work item (2: 256 x 128 output tiles)
  int pw, ph, pn;       // wgrad: geometry of the 64-pixel K box
  float* dbias;         // wgrad: bias gradient dbias[co] += sum_pix dY[pix][co] (one extra N=16 MMA against a ones tile)
  float* dbias2;
  int nseg;
  int seg_taps[2];
  int seg_kc[2];
  signed char tap_dh[2][9], tap_dw[2][9], tap_b[2][9];
  int a_stride[2];
  int b_mn_major;
  int H, W;             // spatial extent the M tiles walk over
  int M_total, N_total;
  // persistent schedule: work item w -> (tile, K split)
  int num_work, num_n, splits;
  int pair_work;        // halo mode with 256-row items: work items >= pair_work are TAIL items of one 128-row half each
  int iters_total, iters_per_split;   // mode 0: (tap, k-chunk) iterations; mode 1: 64-pixel chunks
  // epilogue 0
  const float* bias;
  const float* bias2;
  const float* rowvec;
  long long ld_rowvec;
  int rows_per_vec;
  int has_c;            // a [128 x 128] bf16 tile of mapC (residual, or the output itself) is added
  int store_bf16;       // 0: no bf16 output (fp32 copy only)
  float* out_f32;       // optional fp32 copy [M][N_total] (row-major, ld = N_total)
  // GroupNorm statistics of the OUTPUT, fused into the store epilogue: qsum[n][N_total/4][2] += (sum, sum of
  // squares) of every 4-channel quad of sample n (fp32 atomics; the consumer combines cpg/4 quads per group).
  // Needs H*W % 128 == 0: the 128 rows of a tile belong to one sample.
  float* qsum;
  // kNorm instantiations (inference): the A operand is the RAW activation a_raw[N][H][W][a_ld]; the halo is filled by four
  // transform warps that apply GroupNorm + SiLU on the way, y = silu(x * coef[n][c][0] + coef[n][c][1]) (zero outside the map)
  const void* a_raw;
  long long a_ld;
  int a_c;
  const float* gn_coef;
  // dynamic work distribution (kDyn instantiations): sched[0] = items handed out beyond the first one of every CTA,
  // sched[1] = CTAs that have finished; the last CTA zeroes both for the next launch that gets this pair
  int* sched;
  // wgrad
  int taps;             // valid taps
  int ci_total;
  int w_col0;
  int num_co;
};

struct Work {
  int m_tile, n_tile;   // mode 0: pixel tile, cout tile;  mode 1: cout tile, -
  int y0;               // mode 1: first of the (up to two) consecutive (cin tile, tap) entries sharing this item's dY tiles
  int nh;               // mode 1: entries of this item (item 0 of a cout tile: entry 0 alone + the bias gradient);
                        // mode 0: 128-row sub-tiles of this item (a tail item of the halo kernel has one)
  int half0;            // mode 0: first sub-tile of the item (tail items: 0 or 1)
  int it0, nit;         // iteration range
};
// wgrad entry y -> (cin tile, tap index)
__device__ __forceinline__ void wg_entry(const IgemmArgs& a, int y, int& n_tile, int& tap) {
  n_tile = y / a.taps;
  tap = y - n_tile * a.taps;
}

__device__ __forceinline__ Work decode_work(const IgemmArgs& a, int w) {
  Work k;
  const int z = w % a.splits;
  const int t = w / a.splits;
  k.half0 = 0;
  if (a.mode == 0) {
    int tt = t;
    k.nh = a.mt;
    if (a.pair_work > 0 && t >= a.pair_work) {      // tail wave: one half of a 16 x 16 patch per item
      const int sidx = t - a.pair_work;
      tt = a.pair_work + (sidx >> 1);
      k.half0 = sidx & 1;
      k.nh = 1;
    }
    k.n_tile = tt % a.num_n;
    k.m_tile = tt / a.num_n;
    k.y0 = 0;
  } else {
    // items of a cout tile: {entry 0 (+ bias gradient)}, {1, 2}, {3, 4}, ...
    k.m_tile = t % a.num_co;
    const int j = t / a.num_co;
    k.y0 = j == 0 ? 0 : 2 * j - 1;
    k.nh = j == 0 ? 1 : min(2, a.num_n * a.taps - k.y0);
    k.n_tile = 0;
  }
  k.it0 = z * a.iters_per_split;
  k.nit = min(a.iters_per_split, a.iters_total - k.it0);
  return k;
}

// pixel coordinates of an M tile: plain = 128 consecutive pixels of the flattened (n, h, w) raster;
// halo = an 8 (w) x 16 (h) patch of one image
__device__ __forceinline__ void tile_origin(const IgemmArgs& a, bool halo, int m_tile, int& w0, int& h0, int& n0) {
  if (halo) {
    const int tw = a.W / PATCH_W, th = a.H / PATCH_H;
    w0 = (m_tile % tw) * PATCH_W;
    h0 = ((m_tile / tw) % th) * PATCH_H;
    n0 = m_tile / (tw * th);
  } else {
    const int p0 = m_tile * TILE_M;
    w0 = p0 % a.W;
    h0 = (p0 / a.W) % a.H;
    n0 = p0 / (a.W * a.H);
  }
}

// walks the (segment, tap, k-chunk) iteration space of an activation GEMM without integer division in the
// steady state.  plain order: tap-major, k-chunk-minor; halo order (segment 0): k-chunk-major, tap-minor.
template <bool kHalo>
struct IterWalker {
  int seg, tap, kc;
  __device__ __forceinline__ void init(const IgemmArgs& a, int it) {
    const int seg0_total = a.seg_taps[0] * a.seg_kc[0];
    seg = it >= seg0_total ? 1 : 0;
    if (seg) it -= seg0_total;
    if (kHalo && seg == 0) { kc = it / a.seg_taps[0]; tap = it - kc * a.seg_taps[0]; }
    else { tap = it / a.seg_kc[seg]; kc = it - tap * a.seg_kc[seg]; }
  }
  __device__ __forceinline__ bool halo_it() const { return kHalo && seg == 0; }
  __device__ __forceinline__ void next(const IgemmArgs& a) {
    if (kHalo && seg == 0) {
      if (++tap == a.seg_taps[0]) { tap = 0; if (++kc == a.seg_kc[0]) { seg = 1; kc = 0; } }
    } else {
      if (++kc == a.seg_kc[seg]) { kc = 0; if (++tap == a.seg_taps[seg] && seg == 0) { seg = 1; tap = 0; } }
    }
  }
};

// descriptor words.  hi: SBO | version (bit 46) | SWIZZLE_128B (bits 61..63); lo: start address | LBO
constexpr uint32_t DESC_HI_SBO1024 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_HI_HALO = ((uint32_t)(HALO_W * 128) >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_HI_HALO2 = ((uint32_t)(HALO2_W * 128) >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_LO_KMAJOR = (16u >> 4) << 16;     // LBO (unused for swizzled K-major)
constexpr uint32_t DESC_LO_MNMAJOR = (8192u >> 4) << 16;  // LBO = distance between the two 64-wide MN atoms
__device__ __forceinline__ uint64_t make_desc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// ---- work distribution --------------------------------------------------------------------------------------------
// static (kDyn = false): CTA b runs items b, b + grid, b + 2 grid, ... -- every role computes the sequence itself.
// dynamic (kDyn = true): item b first, then whatever an atomic counter hands out.  A persistent CTA that shares its SM
// with a kernel of another stream (the sampler's mt19937 generator, NCCL's channels) then simply takes fewer items,
// instead of making the whole grid wait for it.  The A producer owns the sequence: it fetches two items ahead (the
// atomic's latency hides behind an item's loads) and publishes (item, next item) pairs through a shared-memory ring
// to the B producer, the MMA issuer and the epilogue warps; -1 ends the sequence.
template <bool kDyn>
struct WorkPub;
template <>
struct WorkPub<false> {
  int w, step, n;
  __device__ __forceinline__ WorkPub(const IgemmArgs& a, int first, int step_, uint64_t*, uint64_t*, volatile int2*)
      : w(first - step_), step(step_), n(a.num_work) {}
  __device__ __forceinline__ int next() {
    w += step;
    return w < n ? w : -1;
  }
};
template <>
struct WorkPub<true> {
  int first, w1, w2, step, n, slot;
  uint32_t ph;
  bool started;
  int* ctr;
  uint64_t *full, *empty;
  volatile int2* q;
  __device__ __forceinline__ WorkPub(const IgemmArgs& a, int first_, int step_, uint64_t* full_, uint64_t* empty_, volatile int2* q_)
      : first(first_), w1(-1), w2(-1), step(step_), n(a.num_work), slot(0), ph(0), started(false), ctr(a.sched), full(full_),
        empty(empty_), q(q_) {}
  __device__ __forceinline__ int fetch() {
    const int t = atomicAdd(ctr, 1) + step;
    return t < n ? t : -1;
  }
  __device__ __forceinline__ int next() {
    int w;
    if (!started) {
      started = true;
      w = first < n ? first : -1;
      w1 = w >= 0 ? fetch() : -1;
    } else {
      w = w1;
      w1 = w2;
    }
    w2 = w1 >= 0 ? fetch() : -1;     // needed at the call after this one
    mbar_wait(&empty[slot], ph ^ 1);
    q[slot].x = w;
    q[slot].y = w1;
    mbar_arrive(&full[slot]);        // release: the pair is visible to whoever acquires this phase
    if (++slot == WQ_SLOTS) { slot = 0; ph ^= 1; }
    return w;
  }
};
template <bool kDyn>
struct WorkSub;
template <>
struct WorkSub<false> {
  int w, step, n;
  __device__ __forceinline__ WorkSub(const IgemmArgs& a, int first, int step_, uint64_t*, uint64_t*, volatile int2*, bool)
      : w(first - step_), step(step_), n(a.num_work) {}
  __device__ __forceinline__ int next(int& wn) {
    w += step;
    wn = w + step < n ? w + step : -1;
    return w < n ? w : -1;
  }
};
template <>
struct WorkSub<true> {
  int slot;
  uint32_t ph;
  bool whole_warp;   // all 32 lanes run the loop (epilogue warps): one arrival per warp, after every lane has read
  uint64_t *full, *empty;
  volatile int2* q;
  __device__ __forceinline__ WorkSub(const IgemmArgs&, int, int, uint64_t* full_, uint64_t* empty_, volatile int2* q_, bool whole_warp_)
      : slot(0), ph(0), whole_warp(whole_warp_), full(full_), empty(empty_), q(q_) {}
  __device__ __forceinline__ int next(int& wn) {
    mbar_wait(&full[slot], ph);
    const int w = q[slot].x;
    wn = q[slot].y;
    if (whole_warp) {
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);
    } else {
      mbar_arrive(&empty[slot]);
    }
    if (++slot == WQ_SLOTS) { slot = 0; ph ^= 1; }
    return w;
  }
};

// kMode 0: activation GEMM (fprop / dgrad), 1: wgrad.  kHalo: see HALO_* above.  Template parameters keep the
// single-thread producer / MMA-issue loops minimal: those loops pace the tensor pipe (4 MMAs = 256 cycles per
// iteration), every extra instruction in them showed up 1:1 in the measured throughput.
// Warp roles: 0 = TMA producer of the A operand, 6 = TMA producer of the B operand, 1 = MMA issuer (+TMEM
// allocator), 2..5 = epilogue.
// kMT: 128-row sub-tiles per work item (2 = a 256 x 128 output tile: both halves share every B tile, which cuts the
// L2 -> SM operand traffic per FLOP by 25 %; 3 stages of [B][A0][A1] = 48 KB; the two TMEM stages hold 256 columns each)
// kStats: the store epilogue also accumulates GroupNorm quad sums of the output (IgemmArgs::qsum); a template
// parameter so that the plain instantiations carry none of that code (measured: +4 % on the level-0 convolutions when
// it was a run-time branch -- the store epilogue of a 256 x 128 item is as long as its main loop).
// kNorm (halo kernel, inference): GroupNorm + SiLU of the INPUT folded into the operand path.  Eight extra warps (11-18,
// 608 threads per CTA) fill the halo slots themselves: 16-byte loads of the raw activation (prefetched one chunk ahead),
// normalise + SiLU in registers, swizzled 16-byte stores into the slot (the layout TMA would have produced), proxy fence,
// one arrival on the slot's full barrier.  The normalised activation never exists in HBM: the stand-alone apply pass
// (read x, write a) and the conv's read of a collapse into one read of x.  Measured (256 x 128 x 128, 128 -> 128): the
// folded convolution costs 1.10 x the plain one (+0.10 ms; the apply pass it replaces is 0.33 ms).  The remaining cost
// is SRAM bandwidth: the MMAs of this kernel already read shared memory at its limit, and the register path moves every
// halo byte through the LSU twice (a variant that only signals the barriers runs 6 % FASTER than the plain kernel).
template <int kMode, bool kHalo, int kMT, bool kStats = false, bool kDyn = false, bool kNorm = false>
__global__ void __launch_bounds__(kNorm ? IGEMM_THREADS_NORM : IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
             const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
             const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapD,
             const __grid_constant__ IgemmArgs args) {
  pdl_launch_dependents();   // the next kernel may be scheduled as SMs drain; it blocks in its own pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET from the __shared__ array (not through an integer cast of the generic pointer): the
  // compiler keeps the shared address space, so every epilogue access is LDS / STS instead of a generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kASlots = kHalo ? 2 : 0;
  constexpr int kBSlots = kHalo ? (kMT == 2 ? 4 : 5) : (kMT == 2 ? 3 : 4);
  constexpr int kHaloW = kMT == 2 ? HALO2_W : HALO_W;             // halo row pitch (pixels)
  constexpr int kHaloBytes = kMT == 2 ? HALO2_BYTES : HALO_BYTES;  // bytes of one halo box
  constexpr int kHaloSlot = kMT == 2 ? HALO2_SLOT : HALO_BYTES;    // A-ring slot pitch
  constexpr uint32_t kDescHiHalo = kMT == 2 ? DESC_HI_HALO2 : DESC_HI_HALO;
  // plain stage: mode 0 = [B][A0](A1) (two pixel sub-tiles share the weight tile); mode 1 = [B0](B1)[A] (two
  // (cin tile, tap) entries share the dY tile)
  constexpr int kBSlotBytes = kHalo ? B_BYTES : (kMT + 1) * A_BYTES;
  constexpr int kAOff = kMode == 1 ? kMT * B_BYTES : B_BYTES;
  constexpr int kAccCols = TILE_N * kMT;
  constexpr int kAccStages = 2;
  uint8_t* a_ring = smem;                                   // halo slots (halo mode only)
  uint8_t* b_ring = smem + kASlots * kHaloSlot;
  float* bias_s = reinterpret_cast<float*>(smem + SMEM_BIAS_OFF);
  float* qacc = reinterpret_cast<float*>(smem + SMEM_QACC_OFF);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + SMEM_BAR_OFF);
  uint64_t* a_empty = a_full + MAX_A_SLOTS;
  uint64_t* b_full = a_empty + MAX_A_SLOTS;
  uint64_t* b_empty = b_full + MAX_B_SLOTS;
  uint64_t* tmem_full_bar = b_empty + MAX_B_SLOTS;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint64_t* c_full_bar = tmem_empty_bar + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_full_bar + 2);
  uint64_t* wq_full = reinterpret_cast<uint64_t*>(smem + SMEM_WQ_OFF);   // [WQ_SLOTS] (kDyn)
  uint64_t* wq_empty = wq_full + WQ_SLOTS;
  volatile int2* wq = reinterpret_cast<volatile int2*>(wq_empty + WQ_SLOTS);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < MAX_A_SLOTS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < MAX_B_SLOTS; ++s) {
      mbar_init(&b_full[s], kHalo ? 1 : 2);   // plain stage: the A and the B producer both arrive
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], blockDim.x >= IGEMM_THREADS ? 8 : 4);   // one arrival per epilogue warp
      mbar_init(&c_full_bar[s], 1);
    }
    if (kDyn) {
      for (int s = 0; s < WQ_SLOTS; ++s) {
        mbar_init(&wq_full[s], 1);
        // B producer, MMA issuer, one per epilogue warp (kNorm: + four transform warps)
        mbar_init(&wq_empty[s], 2 + (blockDim.x >= IGEMM_THREADS ? 8 : 4) + (kNorm ? NORM_WARPS : 0));
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapB0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapD) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (kMode == 1 && args.dbias != nullptr && warp >= 2 && warp <= 5) {   // ones tile (bf16 1.0 = 0x3F80) for the bias-gradient MMA
    uint4* o = reinterpret_cast<uint4*>(smem + SMEM_ONES_OFF);
    for (int i = threadIdx.x - 64; i < ONES_BYTES / 16; i += 128) o[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int w_first = blockIdx.x, w_step = gridDim.x;
#ifdef MDM_IGEMM_DEBUG_WAIT
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("BARS a_full 0x%x a_empty 0x%x b_full 0x%x b_empty 0x%x tmem_full 0x%x tmem_empty 0x%x wq_full 0x%x wq_empty 0x%x c_full 0x%x\n",
           smem_u32(a_full), smem_u32(a_empty), smem_u32(b_full), smem_u32(b_empty), smem_u32(tmem_full_bar), smem_u32(tmem_empty_bar),
           smem_u32(wq_full), smem_u32(wq_empty), smem_u32(c_full_bar));
#endif
  pdl_wait();   // barrier init, TMEM allocation and descriptor prefetch above overlapped the previous kernel's tail

  if (warp == 0) {
   if (elect_one()) {
    // ============================== TMA producer: A operand ==================================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t ka[2] = {0u, 0u};   // kNorm: waits done on a_empty[s] (this thread fills only the shortcut patches)
    WorkPub<kDyn> feed(args, w_first, w_step, wq_full, wq_empty, wq);
    for (int w = feed.next(); w >= 0; w = feed.next()) {
      const Work k = decode_work(args, w);
      if (kMode == 0) {
        int w0, h0, n0, w1 = 0, h1 = 0, n1 = 0;
        tile_origin(args, kHalo, k.m_tile * kMT, w0, h0, n0);
        if (kMT == 2) tile_origin(args, kHalo, k.m_tile * kMT + 1, w1, h1, n1);
        IterWalker<kHalo> it;
        it.init(args, k.it0);
        for (int i = 0; i < k.nit; ++i) {
          if (kHalo) {
            if (kNorm) {
              // the transform warps fill the halo chunks, this thread the shortcut patches.  Each filler has its OWN
              // "slot free" barrier per slot (a_empty[s] here, a_empty[2 + s] for the transform warps; the MMA issuer
              // commits to the one whose owner fills the slot next): a role that merely watched a barrier it does not
              // fill could fall two phases behind, which a parity wait cannot tell from zero
              if (it.halo_it()) {
                if (it.tap == 0) sa ^= 1;
              } else {
                mbar_wait(&a_empty[sa], ka[sa] & 1);
                ++ka[sa];
                mbar_expect_tx(&a_full[sa], k.nh * A_BYTES);
                if (k.nh == kMT) {
                  tma_load_4d(&mapA1, a_ring + sa * kHaloSlot, &a_full[sa], it.kc * TILE_K, w0, h0, n0);
                  if (kMT == 2) tma_load_4d(&mapA1, a_ring + sa * kHaloSlot + A_BYTES, &a_full[sa], it.kc * TILE_K, w1, h1, n1);
                } else {
                  tma_load_4d(&mapA1, a_ring + sa * kHaloSlot, &a_full[sa], it.kc * TILE_K, k.half0 ? w1 : w0, k.half0 ? h1 : h0,
                              k.half0 ? n1 : n0);
                }
                sa ^= 1;
              }
            } else if (it.halo_it()) {
              if (it.tap == 0) {   // one halo per channel chunk
                mbar_wait(&a_empty[sa], pa ^ 1);
                mbar_expect_tx(&a_full[sa], kHaloBytes);
                tma_load_4d(&mapA0, a_ring + sa * kHaloSlot, &a_full[sa], it.kc * TILE_K, w0 - 1, h0 - 1, n0);
                if (++sa == kASlots) { sa = 0; pa ^= 1; }
              }
            } else {               // fused 1x1 shortcut segment: a plain patch in a halo slot
              mbar_wait(&a_empty[sa], pa ^ 1);
              mbar_expect_tx(&a_full[sa], k.nh * A_BYTES);
              if (k.nh == kMT) {
                tma_load_4d(&mapA1, a_ring + sa * kHaloSlot, &a_full[sa], it.kc * TILE_K, w0, h0, n0);
                if (kMT == 2) tma_load_4d(&mapA1, a_ring + sa * kHaloSlot + A_BYTES, &a_full[sa], it.kc * TILE_K, w1, h1, n1);
              } else {             // tail item: its one patch goes to the front of the slot
                tma_load_4d(&mapA1, a_ring + sa * kHaloSlot, &a_full[sa], it.kc * TILE_K, k.half0 ? w1 : w0, k.half0 ? h1 : h0,
                            k.half0 ? n1 : n0);
              }
              if (++sa == kASlots) { sa = 0; pa ^= 1; }
            }
          } else {
            const CUtensorMap* mA = it.seg == 0 ? &mapA0 : &mapA1;
            const int st = args.a_stride[it.seg];
            const int dw = args.tap_dw[it.seg][it.tap], dh = args.tap_dh[it.seg][it.tap];
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_expect_tx(&b_full[sb], kMT * A_BYTES);
            tma_load_4d(mA, b_ring + sb * kBSlotBytes + kAOff, &b_full[sb], it.kc * TILE_K, w0 * st + dw, h0 * st + dh, n0);
            if (kMT == 2)
              tma_load_4d(mA, b_ring + sb * kBSlotBytes + kAOff + A_BYTES, &b_full[sb], it.kc * TILE_K, w1 * st + dw,
                          h1 * st + dh, n1);
            if (++sb == kBSlots) { sb = 0; pb ^= 1; }
          }
          it.next(args);
        }
      } else {
        // wgrad: A = dY (MN-major, M = co); 64 pixels (pw x ph x pn box) per iteration
        const int co0 = k.m_tile * TILE_M;
        const int hw = args.W * args.H;
        const int p0 = k.it0 * TILE_K;
        int n0 = p0 / hw;
        const int rem = p0 - n0 * hw;
        int h0 = rem / args.W, w0 = rem - h0 * args.W;
        for (int i = 0; i < k.nit; ++i) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          uint8_t* a_dst = b_ring + sb * kBSlotBytes + kAOff;
          mbar_expect_tx(&b_full[sb], A_BYTES);
          tma_load_4d(&mapA0, a_dst, &b_full[sb], co0, w0, h0, n0);
          tma_load_4d(&mapA0, a_dst + 8192, &b_full[sb], co0 + 64, w0, h0, n0);
          if (++sb == kBSlots) { sb = 0; pb ^= 1; }
          w0 += args.pw;
          if (w0 >= args.W) { w0 = 0; h0 += args.ph; if (h0 >= args.H) { h0 = 0; n0 += args.pn; } }
        }
      }
    }
   }
  } else if (warp == 6) {
   if (elect_one()) {
    // ============================== TMA producer: B operand ==================================
    int sb = 0, wn_unused;
    uint32_t pb = 0;
    WorkSub<kDyn> feed(args, w_first, w_step, wq_full, wq_empty, wq, false);
    for (int w = feed.next(wn_unused); w >= 0; w = feed.next(wn_unused)) {
      const Work k = decode_work(args, w);
      if (kMode == 0) {
        const int ncol0 = k.n_tile * TILE_N;
        IterWalker<kHalo> it;
        it.init(args, k.it0);
        for (int i = 0; i < k.nit; ++i) {
          const CUtensorMap* mB = it.seg == 0 ? &mapB0 : &mapB1;
          const int tb = args.tap_b[it.seg][it.tap];
          mbar_wait(&b_empty[sb], pb ^ 1);
          uint8_t* b_dst = b_ring + sb * kBSlotBytes;
          mbar_expect_tx(&b_full[sb], B_BYTES);
          if (!args.b_mn_major) {
            tma_load_3d(mB, b_dst, &b_full[sb], it.kc * TILE_K, tb, ncol0);
          } else {
            tma_load_3d(mB, b_dst, &b_full[sb], ncol0, tb, it.kc * TILE_K);
            tma_load_3d(mB, b_dst + 8192, &b_full[sb], ncol0 + 64, tb, it.kc * TILE_K);
          }
          if (++sb == kBSlots) { sb = 0; pb ^= 1; }
          it.next(args);
        }
      } else {
        // wgrad: B = X shifted by the tap (MN-major, N = ci); up to kMT (cin tile, tap) entries per item
        int ci0[kMT], dh[kMT], dw[kMT];
        const int nvalid = k.nh;
#pragma unroll
        for (int half = 0; half < kMT; ++half) {
          int nt, tp;
          wg_entry(args, half < nvalid ? k.y0 + half : k.y0, nt, tp);
          ci0[half] = nt * TILE_N;
          dh[half] = args.tap_dh[0][tp];
          dw[half] = args.tap_dw[0][tp];
        }
        const int st = args.a_stride[0];
        const int hw = args.W * args.H;
        const int p0 = k.it0 * TILE_K;
        int n0 = p0 / hw;
        const int rem = p0 - n0 * hw;
        int h0 = rem / args.W, w0 = rem - h0 * args.W;
        for (int i = 0; i < k.nit; ++i) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          uint8_t* b_dst = b_ring + sb * kBSlotBytes;
          mbar_expect_tx(&b_full[sb], nvalid * B_BYTES);
#pragma unroll
          for (int half = 0; half < kMT; ++half) {
            if (half < nvalid) {
              tma_load_4d(&mapB0, b_dst + half * B_BYTES, &b_full[sb], ci0[half], w0 * st + dw[half], h0 * st + dh[half], n0);
              tma_load_4d(&mapB0, b_dst + half * B_BYTES + 8192, &b_full[sb], ci0[half] + 64, w0 * st + dw[half], h0 * st + dh[half], n0);
            }
          }
          if (++sb == kBSlots) { sb = 0; pb ^= 1; }
          w0 += args.pw;
          if (w0 >= args.W) { w0 = 0; h0 += args.ph; if (h0 >= args.H) { h0 = 0; n0 += args.pn; } }
        }
      }
    }
   }
  } else if (warp == 1) {
   if (elect_one()) {
    // ============================== MMA issuer ================================================
    const int a_mn = kMode == 1 ? 1 : 0;
    const int b_mn = kMode == 1 ? 1 : args.b_mn_major;
    const uint32_t idesc = make_idesc(a_mn, b_mn);
    // descriptor low words of slot 0 and the per-16-element K step of each operand
    const uint32_t lo_b0 = ((smem_u32(b_ring) >> 4) & 0x3FFF) | (b_mn ? DESC_LO_MNMAJOR : DESC_LO_KMAJOR);
    const uint32_t lo_a0 = lo_b0 + (kAOff >> 4);   // plain stage: A tile(s) behind the B tile(s) (same LBO class in mode 1 / 0)
    const uint32_t lo_a_plain0 = (kMode == 1 || !b_mn) ? lo_a0 : (((smem_u32(b_ring) + kAOff) >> 4) & 0x3FFF) | DESC_LO_KMAJOR;
    const uint32_t kstep_a = a_mn ? (2048u >> 4) : (32u >> 4);
    const uint32_t kstep_b = b_mn ? (2048u >> 4) : (32u >> 4);
    const uint32_t lo_halo0 = ((smem_u32(a_ring) >> 4) & 0x3FFF) | DESC_LO_KMAJOR;
    int sa = 0, sb = 0, local = 0, wn_unused;
    uint32_t pa = 0, pb = 0;
    WorkSub<kDyn> feed(args, w_first, w_step, wq_full, wq_empty, wq, false);
    for (int w = feed.next(wn_unused); w >= 0; w = feed.next(wn_unused), ++local) {
      const Work k = decode_work(args, w);
      const int acc = local % kAccStages;
      mbar_wait(&tmem_empty_bar[acc], ((local / kAccStages) & 1) ^ 1);   // the epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kAccCols;
      const int nhalf = k.nh;      // wgrad: the second entry of the pair may not exist; halo tail items: one sub-tile
      const uint32_t half_shift = (kMode == 0 && kHalo) ? (uint32_t)k.half0 : 0u;   // tail item: which half of the patch
      uint32_t lo_a = 0;
      int sa_cur = 0, tap = 0;
      int upos = 0, upos_cur = 0;   // kNorm: A-ring uses of this item so far
      const int n_uses = kHalo ? args.seg_kc[0] + (args.nseg > 1 ? args.seg_taps[1] * args.seg_kc[1] : 0) : 0;
      // wgrad work items (ci tile 0, first tap) also accumulate the bias gradient: A = dY tile, B = ones, N = 16
      const bool bias_item = kMode == 1 && args.dbias != nullptr && k.y0 == 0;
      const uint32_t idesc16 = (idesc & ~(0x3Fu << 17)) | ((16u >> 3) << 17);
      const uint32_t lo_ones = ((smem_u32(smem + SMEM_ONES_OFF) >> 4) & 0x3FFF) | DESC_LO_MNMAJOR;
      const int n_halo = kHalo ? args.seg_taps[0] * args.seg_kc[0] : 0;   // halo iterations come first (no K split)
      bool ready = mbar_try_wait(&b_full[sb], pb);
      for (int i = 0; i < k.nit; ++i) {
        const bool halo_it = kHalo && i < n_halo;
        if (kHalo && (!halo_it || tap == 0)) {      // this iteration starts on a fresh A-ring slot
          sa_cur = sa;
          upos_cur = upos++;
          mbar_wait(&a_full[sa], pa);
          lo_a = lo_halo0 + sa * (kHaloSlot >> 4);
          if (++sa == kASlots) { sa = 0; pa ^= 1; }
        }
        if (!ready) mbar_wait(&b_full[sb], pb);
        tcgen05_fence_after();
        const uint32_t lo_b = lo_b0 + sb * (kBSlotBytes >> 4);
        uint32_t lo_at;
        if (kHalo) {
          lo_at = lo_a;
          if (halo_it) lo_at += (uint32_t)(((args.tap_dh[0][tap] + 1) * kHaloW + (args.tap_dw[0][tap] + 1)) * (128 >> 4));
        } else {
          lo_at = lo_a_plain0 + sb * (kBSlotBytes >> 4);
        }
        const uint32_t hi_a = halo_it ? kDescHiHalo : DESC_HI_SBO1024;
        // second 128-row sub-tile: the 8 pixels to the right inside the halo, or the second plain patch
        const uint32_t half_a = halo_it ? (uint32_t)((8 * 128) >> 4) : (uint32_t)(A_BYTES >> 4);
        // look at the next stage's barrier now: its latency hides behind the MMA issue below
        const int sb_cur = sb;
        if (++sb == kBSlots) { sb = 0; pb ^= 1; }
        ready = (i + 1 < k.nit) ? mbar_try_wait(&b_full[sb], pb) : false;
#pragma unroll
        for (int half = 0; half < kMT; ++half) {
          if (half < nhalf) {
            // mode 0: the halves are two A (pixel) sub-tiles against one B; mode 1: two B tiles against one A
            const uint32_t la = lo_at + (kMode == 0 ? (half + (halo_it ? half_shift : 0u)) * half_a : 0);
            const uint32_t lb = lo_b + (kMode == 1 ? half * (B_BYTES >> 4) : 0);
#pragma unroll
            for (int kk = 0; kk < TILE_K / 16; ++kk) {
              umma_bf16(tmem_d + half * TILE_N, make_desc(hi_a, la + kk * kstep_a), make_desc(DESC_HI_SBO1024, lb + kk * kstep_b),
                        idesc, (i | kk) != 0 ? 1u : 0u);
            }
          }
        }
        if (bias_item) {
#pragma unroll
          for (int kk = 0; kk < TILE_K / 16; ++kk)
            umma_bf16(tmem_d + TILE_N, make_desc(hi_a, lo_at + kk * kstep_a),
                      make_desc(DESC_HI_SBO1024, lo_ones + kk * (2048u >> 4)), idesc16, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&b_empty[sb_cur]);
        if (kHalo) {
          if (!halo_it || tap == args.seg_taps[0] - 1) {
            if (kNorm) {   // free the slot towards whoever fills it next (two uses on): a halo chunk -> the transform warps
              const bool next_halo = upos_cur + 2 < args.seg_kc[0] || upos_cur + 2 >= n_uses;
              umma_commit(&a_empty[(next_halo ? 2 : 0) + sa_cur]);
            } else {
              umma_commit(&a_empty[sa_cur]);
            }
          }
          if (halo_it && ++tap == args.seg_taps[0]) tap = 0;
        }
      }
      umma_commit(&tmem_full_bar[acc]);
    }
   }
  } else if (kNorm && warp >= 11) {
    // ============================== transform producers (kNorm): GroupNorm + SiLU on the way into the halo ==========
    // 256 threads; thread t owns the 8-channel group t % 8 of 11 halo pixels (column (t / 8) % 16 of every second halo row,
    // plus a share of the two right-hand columns).  The sixteen-byte loads
    // of the NEXT halo chunk (next channel chunk, or chunk 0 of the next work item) are issued before this role waits
    // for its slot, so the memory latency hides behind the MMAs of the chunk in flight.
    if (kMode == 0 && kHalo && kMT == 2) {
      const int tt = (int)threadIdx.x - IGEMM_THREADS;   // 0 .. NORM_THREADS - 1
      const int c8 = tt & 7;
      const int r0 = tt >> 3;                            // 0 .. 31
      const int col = r0 & 15, rpar = r0 >> 4;           // pieces 0..8: halo pixel (2 u + rpar, col)
      const bf16_t* xin = reinterpret_cast<const bf16_t*>(args.a_raw);
      constexpr int kRowPieces = HALO2_H / 2;            // 9
      constexpr int kPieces = kRowPieces + 2;            // + the two right-hand columns: 36 pixels over 32 + 4 threads
      static_assert(MAX_A_SLOTS >= 4 && NORM_THREADS == 256 && HALO2_W == 18 && HALO2_H == 18, "piece mapping of the transform warps");
      const int KC = args.seg_kc[0];
      const int n_sc = args.nseg > 1 ? args.seg_taps[1] * args.seg_kc[1] : 0;   // shortcut-segment slot uses per item
      uint4 v[kPieces];
      uint32_t inmask = 0;
      const int pitch_px = args.a_ld;                    // elements between horizontally adjacent pixels
      const int pitch_row = args.W * args.a_ld;          // ... vertically adjacent pixels (fits 32 bits: one image row)
      // halo pixel of piece u (row, column) -- compile-time u
      auto piece_px = [&](int u, int& hy, int& hx) -> bool {
        if (u < kRowPieces) { hy = 2 * u + rpar; hx = col; return true; }
        const int pix = r0 + 32 * (u - kRowPieces);      // 0 .. 63, 36 used
        hy = pix >> 1; hx = 16 + (pix & 1);
        return pix < 36;
      };
      // per-thread constants of the whole kernel: element offset of every piece from the halo origin, and its byte
      // offset inside a slot (128 B per halo pixel, 16-byte groups XOR-swizzled by pixel & 7 -- what TMA SWIZZLE_128B writes)
      int rel[kPieces];
      uint32_t sts[kPieces];
#pragma unroll
      for (int u = 0; u < kPieces; ++u) {
        int hy, hx;
        piece_px(u, hy, hx);
        rel[u] = hy * pitch_row + hx * pitch_px + c8 * 8;
        const int row = hy * kHaloW + hx;
        sts[u] = (uint32_t)(row * 128) + (((uint32_t)c8 ^ (uint32_t)(row & 7)) << 4);
      }
      // per item: halo origin (element pointer, may lie outside the map: only in-bounds pieces are dereferenced), the
      // in-bounds mask of this thread's pieces, the sample index (coefficients)
      const bf16_t* nbase = xin;
      uint32_t nmask = 0;
      int n_next = 0;
      auto locate = [&](int item) {
        const Work k = decode_work(args, item);
        int w0, h0, n0;
        tile_origin(args, kHalo, k.m_tile * kMT, w0, h0, n0);
        nbase = xin + (((long long)n0 * args.H + (h0 - 1)) * args.W + (w0 - 1)) * (long long)args.a_ld;
        nmask = 0;
#pragma unroll
        for (int u = 0; u < kPieces; ++u) {
          int hy, hx;
          const bool used = piece_px(u, hy, hx);
          if (used && (unsigned)(h0 - 1 + hy) < (unsigned)args.H && (unsigned)(w0 - 1 + hx) < (unsigned)args.W) nmask |= 1u << u;
        }
        n_next = n0;
      };
      auto issue = [&](const bf16_t* base, uint32_t mask, int kc) {
        const bf16_t* b = base + kc * TILE_K;
#pragma unroll
        for (int u = 0; u < kPieces; ++u)
          if ((mask >> u) & 1u) {
            // streamed once: no L1 allocation (L1 shares its SRAM bandwidth with the operand reads of the MMAs)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(b + rel[u]));
          }
      };
      int sa = 0, wn = -1;
      uint32_t kh[2] = {0u, 0u};                   // waits done on this role's "slot free" barrier a_empty[2 + s]
      WorkSub<kDyn> feed(args, w_first, w_step, wq_full, wq_empty, wq, true);
      int w = feed.next(wn);
      int n0 = 0;
      const bf16_t* base = xin;
      int dbg_use = 0;
      (void)dbg_use;
      if (w >= 0) { locate(w); base = nbase; inmask = nmask; n0 = n_next; issue(base, inmask, 0); }
      while (w >= 0) {
        DBG(2, w);
        if (wn >= 0) locate(wn);                     // the item after this one (its first chunk is prefetched below)
        for (int kc = 0; kc < KC; ++kc) {
          DBG(0, dbg_use++); DBG(1, 1);
          // halved coefficients as f32x2 pairs: silu(z) = (z/2) (1 + tanh(z/2))
          unsigned long long sc2[4], sh2[4];
          {
            const float4* cf = reinterpret_cast<const float4*>(args.gn_coef + ((long long)n0 * args.a_c + kc * TILE_K + c8 * 8) * 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 t = __ldg(cf + j);           // {scale, shift} of channels 2j, 2j + 1
              asm("mov.b64 %0, {%1, %2};" : "=l"(sc2[j]) : "f"(0.5f * t.x), "f"(0.5f * t.z));
              asm("mov.b64 %0, {%1, %2};" : "=l"(sh2[j]) : "f"(0.5f * t.y), "f"(0.5f * t.w));
            }
          }
          mbar_wait(&a_empty[2 + sa], (kh[sa] & 1u) ^ 1u);   // the MMAs that read this slot have retired
          ++kh[sa];
          DBG(1, 2);
          uint8_t* slot = a_ring + sa * kHaloSlot;
#pragma unroll
          for (int u = 0; u < kPieces; ++u) {
            int hy, hx;
            if (piece_px(u, hy, hx)) {
              const uint32_t wv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
              uint32_t ov[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                unsigned long long x2, z2, r2;
                float z0, z1, t0, t1;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "r"(wv[e] << 16), "r"(wv[e] & 0xffff0000u));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(z2) : "l"(x2), "l"(sc2[e]), "l"(sh2[e]));      // z / 2
                asm("mov.b64 {%0, %1}, %2;" : "=f"(z0), "=f"(z1) : "l"(z2));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(z0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(z1));
                asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(t0), "f"(t1));
                asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(r2) : "l"(z2), "l"(x2));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(z0), "=f"(z1) : "l"(r2));
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(z0, z1);
                ov[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              const bool in = (inmask >> u) & 1u;              // zero padding stays zero (the reference pads the NORMALISED map)
              *reinterpret_cast<uint4*>(slot + sts[u]) = make_uint4(in ? ov[0] : 0u, in ? ov[1] : 0u, in ? ov[2] : 0u, in ? ov[3] : 0u);
            }
          }
          fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core (async proxy)
          DBG(1, 3);
          asm volatile("bar.sync 2, %0;" ::"n"(NORM_THREADS) : "memory");
          if (tt == 0) mbar_arrive(&a_full[sa]);
          sa ^= 1;
          // next chunk's loads fly while the MMAs of this one run
          DBG(1, 4);
          if (kc + 1 < KC) issue(base, inmask, kc + 1);
          else if (wn >= 0) { base = nbase; inmask = nmask; n0 = n_next; issue(base, inmask, 0); }
        }
        sa ^= n_sc & 1;                              // the shortcut patches (warp 0 loads them) take the next n_sc slots
        DBG(1, 5);
        w = feed.next(wn);
      }
      DBG(1, 6);
    }
  } else if ((warp >= 2 && warp <= 5) || (warp >= 7 && warp <= 10)) {
    // ============================== epilogue ==================================================
    // FOUR or EIGHT warps (block size 224 / 352, chosen per launch).  With eight, two warps share a TMEM lane quarter
    // (a warp may only touch lanes 32 * (warp % 4) ..): group 0 (warps 2-5) takes the 32-column chunks 0-1 of a tile,
    // group 1 (warps 7-10) chunks 2-3 -- the store epilogue of a 256 x 128 item is as long as its main loop with four
    // warps, which shows once a CTA runs >= 5 items (+10..19 % on those layers); short grids keep four (cheaper launch).
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const bool wide = blockDim.x >= IGEMM_THREADS;
    const int grp = warp >= 7 ? 1 : 0;
    const int cc0 = wide ? grp * 2 : 0, cc1 = wide ? cc0 + 2 : TILE_N / 32;   // this warp's chunks
    const int epi_threads = wide ? 256 : 128;
    const int row = q * 32 + lane;          // row of the 128 x 128 tile held by this thread
    const int et = grp ? (int)threadIdx.x - 224 + 128 : (int)threadIdx.x - 64;   // 0..255
    const uint32_t sw = (uint32_t)(row & 7);
    uint8_t* stg_base = smem + SMEM_EPI_OFF;
    uint32_t v[32];
    int local = 0;
    // C / D tiles: plain = 2-D [pixel][channel] boxes {64, 128}; halo = 4-D NHWC boxes {64, 8, 16, 1}
    auto load_c = [&](int m128, int n_tile, int buf) {   // residual / accumulate tile of one 128 x 128 output tile
      uint8_t* so = stg_base + buf * EPI_BYTES;
      mbar_expect_tx(&c_full_bar[buf], EPI_BYTES);
      if (kHalo) {
        int w0, h0, n0;
        tile_origin(args, kHalo, m128, w0, h0, n0);
        tma_load_4d(&mapC, so, &c_full_bar[buf], n_tile * TILE_N, w0, h0, n0);
        tma_load_4d(&mapC, so + 16384, &c_full_bar[buf], n_tile * TILE_N + 64, w0, h0, n0);
      } else {
        tma_load_2d(&mapC, so, &c_full_bar[buf], n_tile * TILE_N, m128 * TILE_M);
        tma_load_2d(&mapC, so + 16384, &c_full_bar[buf], n_tile * TILE_N + 64, m128 * TILE_M);
      }
    };
    if (args.has_c && et == 0 && w_first < args.num_work) {
      const Work k0 = decode_work(args, w_first);
      load_c(k0.m_tile * kMT + k0.half0, k0.n_tile, 0);
    }
    int hl = 0;   // 128 x 128 tiles finished by this CTA (staging buffer / C-tile barrier parity)
    int wn;       // the item after this one (-1: none): its C tile is prefetched below
    WorkSub<kDyn> feed(args, w_first, w_step, wq_full, wq_empty, wq, true);
    for (int w = feed.next(wn); w >= 0; w = feed.next(wn), ++local) {
      const Work k = decode_work(args, w);
      const int acc = local % kAccStages;
      const uint32_t acc_parity = (local / kAccStages) & 1;
      if (args.epi == 0) {
       for (int half = 0; half < k.nh; ++half, ++hl) {
        const uint32_t tmem_acc = tmem_base + acc * kAccCols + half * TILE_N + ((uint32_t)(q * 32) << 16);
        const int m128 = k.m_tile * kMT + k.half0 + half;
        const int buf = hl & 1;
        uint8_t* stg = stg_base + buf * EPI_BYTES;
        const int ncol0 = k.n_tile * TILE_N;
        int w0, h0, n0;
        tile_origin(args, kHalo, m128, w0, h0, n0);
        long long p;       // flattened pixel of this thread's row
        if (kHalo) p = ((long long)n0 * args.H + h0 + (row >> 3)) * args.W + w0 + (row & 7);
        else p = (long long)m128 * TILE_M + row;
        const bool valid = p < args.M_total;
        const bool col_ok = ncol0 + et < args.N_total;
        if (et < TILE_N)
          bias_s[et] = ((args.bias && col_ok) ? __ldg(args.bias + ncol0 + et) : 0.f) + ((args.bias2 && col_ok) ? __ldg(args.bias2 + ncol0 + et) : 0.f);
        const float* rv = (args.rowvec && valid) ? args.rowvec + (p / args.rows_per_vec) * args.ld_rowvec + ncol0 : nullptr;
        if (args.has_c) mbar_wait(&c_full_bar[buf], (hl >> 1) & 1);
        epi_bar_sync(epi_threads);   // bias_s visible; staging[buf] is free (thread 0 waited for its last TMA store below)
        if (half == 0) mbar_wait(&tmem_full_bar[acc], acc_parity);
        tcgen05_fence_after();
#pragma unroll 1
        for (int cc = cc0; cc < cc1; ++cc) {
          tmem_ld32(tmem_acc + cc * 32, v);
          if (cc == cc1 - 1 && half == k.nh - 1) {      // this warp has read its part of the accumulator: hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          float f[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc * 32 + j4 * 4);
            f[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + b4.x;
            f[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + b4.y;
            f[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + b4.z;
            f[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + b4.w;
          }
          const bool chunk_ok = ncol0 + cc * 32 < args.N_total;   // N_total is a multiple of 32
          if (rv && chunk_ok) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(rv + cc * 32) + j4);
              f[j4 * 4 + 0] += r4.x; f[j4 * 4 + 1] += r4.y; f[j4 * 4 + 2] += r4.z; f[j4 * 4 + 3] += r4.w;
            }
          }
          // staging address of this thread's four 16-byte pieces: box = 64 columns, rows of 128 B, 128B swizzle
          uint8_t* rowp = stg + (cc >> 1) * 16384 + row * 128;
          const uint32_t jb = (uint32_t)(cc & 1) * 4;
          if (args.has_c) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 u = *reinterpret_cast<const uint4*>(rowp + (((jb + j4) ^ sw) << 4));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t = __bfloat1622float2(h2[e]);
                f[j4 * 8 + e * 2] += t.x;
                f[j4 * 8 + e * 2 + 1] += t.y;
              }
            }
          }
          if (kStats) {
            // quad sums of this row's 32 columns, then a transposing butterfly over the warp's 32 rows: 16 shuffles
            // leave the warp total of value (lane >> 1) -- sums of quads 0..7, then their sums of squares -- in each lane
            float t[16];
#pragma unroll
            for (int qd = 0; qd < 8; ++qd) {
              const float a0 = f[qd * 4], a1 = f[qd * 4 + 1], a2 = f[qd * 4 + 2], a3 = f[qd * 4 + 3];
              const bool ok = valid && chunk_ok;
              t[qd] = ok ? (a0 + a1) + (a2 + a3) : 0.f;
              t[8 + qd] = ok ? fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, a3 * a3))) : 0.f;
            }
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
            float u8[8], u4[4], u2[2];
#pragma unroll
            for (int i = 0; i < 8; ++i) u8[i] = (b4 ? t[8 + i] : t[i]) + __shfl_xor_sync(0xffffffffu, b4 ? t[i] : t[8 + i], 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) u4[i] = (b3 ? u8[4 + i] : u8[i]) + __shfl_xor_sync(0xffffffffu, b3 ? u8[i] : u8[4 + i], 8);
#pragma unroll
            for (int i = 0; i < 2; ++i) u2[i] = (b2 ? u4[2 + i] : u4[i]) + __shfl_xor_sync(0xffffffffu, b2 ? u4[i] : u4[2 + i], 4);
            float u1 = (b1 ? u2[1] : u2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? u2[0] : u2[1], 2);
            u1 += __shfl_xor_sync(0xffffffffu, u1, 1);
            if (!(lane & 1)) {
              const int idx = lane >> 1;                      // 0..7: sum of quad idx, 8..15: sum of squares of quad idx - 8
              qacc[q * 64 + (cc * 8 + (idx & 7)) * 2 + (idx >> 3)] = u1;   // this warp's slot: written once per tile, no atomics
            }
          }
          if (args.out_f32 && valid && chunk_ok) {
            float4* o = reinterpret_cast<float4*>(args.out_f32 + p * (long long)args.N_total + ncol0 + cc * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) o[j4] = make_float4(f[j4 * 4], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]);
          }
          if (args.store_bf16) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint4 u;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[j4 * 8 + e * 2], f[j4 * 8 + e * 2 + 1]);
              *reinterpret_cast<uint4*>(rowp + (((jb + j4) ^ sw) << 4)) = u;
            }
          }
        }
        fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
        epi_bar_sync(epi_threads);
        if (kStats && et < 64) {
          const int quad = (ncol0 >> 2) + (et >> 1);
          if (quad < (args.N_total >> 2))
            atomicAdd(args.qsum + ((long long)n0 * (args.N_total >> 2) + quad) * 2 + (et & 1),
                      (qacc[et] + qacc[64 + et]) + (qacc[128 + et] + qacc[192 + et]));
        }
        if (et == 0 && args.store_bf16) {
          if (kHalo) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"((uint64_t)&mapD), "r"(smem_u32(stg)), "r"(ncol0), "r"(w0), "r"(h0), "r"(n0) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"((uint64_t)&mapD), "r"(smem_u32(stg + 16384)), "r"(ncol0 + 64), "r"(w0), "r"(h0), "r"(n0) : "memory");
          } else {
            tma_store_2d(&mapD, stg, ncol0, m128 * TILE_M);
            tma_store_2d(&mapD, stg + 16384, ncol0 + 64, m128 * TILE_M);
          }
          bulk_commit();
          bulk_wait_read<1>();   // every store but the one just issued has read its smem: the OTHER tile is free
          if (args.has_c) {      // prefetch the C tile of the next 128 x 128 tile into the other staging buffer
            if (half + 1 < k.nh) {
              load_c(m128 + 1, k.n_tile, buf ^ 1);
            } else if (wn >= 0) {
              const Work kn = decode_work(args, wn);
              load_c(kn.m_tile * kMT + kn.half0, kn.n_tile, buf ^ 1);
            }
          }
        }
       }
      } else {
        // fp32 reduce-add: per (sub-)tile 4 chunks of 32 columns -> 4 boxes {32 fp32, 128 rows} in the two staging tiles
        const int nhalf = kMode == 1 ? k.nh : 1;
        for (int half = 0; half < nhalf; ++half) {
          const uint32_t tmem_acc = tmem_base + acc * kAccCols + half * TILE_N + ((uint32_t)(q * 32) << 16);
          int x0, y0;
          if (kMode == 1) {
            int nt, tp;
            wg_entry(args, k.y0 + half, nt, tp);
            x0 = args.tap_b[0][tp] * args.ci_total + args.w_col0 + nt * TILE_N;
            y0 = k.m_tile * TILE_M;
          } else {
            x0 = k.n_tile * TILE_N;
            y0 = k.m_tile * TILE_M;
          }
          if (et == 0) bulk_wait_read<0>();   // the previous reductions have read the staging tiles
          epi_bar_sync(epi_threads);
          if (half == 0) {
            mbar_wait(&tmem_full_bar[acc], acc_parity);
            tcgen05_fence_after();
            if (kMode == 1 && args.dbias != nullptr && k.y0 == 0 && grp == 0) {
              uint32_t b0, b1, b2, b3, b4, b5, b6, b7;
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3), "=r"(b4), "=r"(b5), "=r"(b6), "=r"(b7)
                           : "r"(tmem_base + acc * kAccCols + TILE_N + ((uint32_t)(q * 32) << 16)) : "memory");
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              const int co = k.m_tile * TILE_M + row;
              if (co < args.M_total) {
                atomicAdd(args.dbias + co, __uint_as_float(b0));
                if (args.dbias2) atomicAdd(args.dbias2 + co, __uint_as_float(b0));
              }
            }
          }
#pragma unroll 1
          for (int cc = cc0; cc < cc1; ++cc) {
            tmem_ld32(tmem_acc + cc * 32, v);
            if (cc == cc1 - 1 && half == nhalf - 1) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            uint8_t* rowp = stg_base + cc * 16384 + row * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<uint4*>(rowp + (((uint32_t)j4 ^ sw) << 4)) = make_uint4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
          }
          fence_proxy_async();
          epi_bar_sync(epi_threads);
          if (et == 0) {
#pragma unroll
            for (int cc = 0; cc < TILE_N / 32; ++cc) tma_reduce_add_2d(&mapD, stg_base + cc * 16384, x0 + cc * 32, y0);
            bulk_commit();
          }
        }
      }
    }
    // the staging tiles must stay valid until the bulk stores have READ them; completion of the global writes
    // themselves is implied by grid completion (what the next kernel / griddepcontrol.wait orders against)
    if (et == 0) bulk_wait_read<0>();
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  if (kDyn && threadIdx.x == 0) {   // this CTA has fetched its last item: the last CTA to get here re-arms the counters
    __threadfence();
    if (atomicAdd(args.sched + 1, 1) == (int)gridDim.x - 1) {
      args.sched[0] = 0;
      args.sched[1] = 0;
      __threadfence();
    }
  }
}

// split-K epilogue of small-M fprop / dgrad layers: ws[M][N] fp32 partial sums -> + bias (+bias2) + rowvec
// + residual (or the old output when accumulating) -> bf16 (and/or fp32) ; ws is re-zeroed for its next user.
__global__ void splitk_finalize_kernel(float* __restrict__ ws, int M, int N, const float* __restrict__ bias,
                                       const float* __restrict__ bias2, const float* __restrict__ rowvec, long long ld_rowvec,
                                       int rows_per_vec, const __nv_bfloat16* resid, long long ld_resid,
                                       __nv_bfloat16* out, long long ld_out, float* __restrict__ out_f32) {
  MDM_PDL_ENTER();
  // one thread per float4 (no loop: parallelism hides the latency), 32-bit index math
  const uint32_t n4 = (uint32_t)N >> 2;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)M * n4) return;
  const uint32_t r = i / n4;
  const int c = (int)(i - r * n4) * 4;
  float4* wp = reinterpret_cast<float4*>(ws + (size_t)r * N + c);
  float4 a = *wp;
  *wp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c)); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  if (bias2) { const float4 b = __ldg(reinterpret_cast<const float4*>(bias2 + c)); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  if (rowvec) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(rowvec + (size_t)(r / (uint32_t)rows_per_vec) * ld_rowvec + c));
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  if (resid) {
    const uint2 u = *reinterpret_cast<const uint2*>(resid + (size_t)r * ld_resid + c);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 t0 = __bfloat1622float2(h2[0]), t1 = __bfloat1622float2(h2[1]);
    a.x += t0.x; a.y += t0.y; a.z += t1.x; a.w += t1.y;
  }
  if (out) {
    uint2 u;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
    h2[0] = __floats2bfloat162_rn(a.x, a.y);
    h2[1] = __floats2bfloat162_rn(a.z, a.w);
    *reinterpret_cast<uint2*>(out + (size_t)r * ld_out + c) = u;
  }
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)r * N + c) = a;
}

// ---- host side: tensor maps ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// activation map: bf16 NHWC [N][H][W][C] with channel stride ld; box {64, bw*stride, bh*stride, bn}
static int make_act_map(CUtensorMap* m, const void* ptr, long long ld, int C, int W, int H, int N, int bw, int bh, int bn, int stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return MDM_E_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(act) failed: %d (ptr=%p ld=%lld C=%d W=%d H=%d N=%d box=%d,%d,%d stride=%d)", (int)r, ptr, ld, C, W, H, N, bw, bh, bn, stride);
    return MDM_E_CUDA;
  }
  return MDM_OK;
}

// parity view of a full-resolution NHWC tensor [N][2H][2W][C]: the pixels (2h + a, 2w + b), addressed as (c, w, h, n)
// with doubled pixel / row strides -- the store target of the fused upsample convolution
static int make_parity_map(CUtensorMap* m, const void* ptr, long long ld, int C, int W, int H, int N, int a, int b, int bw, int bh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return MDM_E_CUDA; }
  const char* base = (const char*)ptr + ((long long)a * 2 * W + b) * ld * 2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)2 * ld * 2, (cuuint64_t)2 * (2 * W) * ld * 2, (cuuint64_t)(2 * H) * (2 * W) * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(parity) failed: %d (ptr=%p ld=%lld C=%d W=%d H=%d N=%d)", (int)r, ptr, ld, C, W, H, N);
    return MDM_E_CUDA;
  }
  return MDM_OK;
}

// weight map: bf16 [rows][taps][cols]; box {64 cols, 1 tap, box_rows}
static int make_w_map(CUtensorMap* m, const void* ptr, int cols, int taps, int rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return MDM_E_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)taps, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * taps * 2};
  cuuint32_t box[3] = {64, 1, (cuuint32_t)box_rows};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weight) failed: %d (cols=%d taps=%d rows=%d)", (int)r, cols, taps, rows);
    return MDM_E_CUDA;
  }
  return MDM_OK;
}

// epilogue map: row-major 2-D [rows][cols] with row stride ld (elements); bf16 box {64, 128} or fp32 box {32, 128},
// both 128 bytes wide with the 128B swizzle (the staging tiles are written in that layout)
static int make_tile_map(CUtensorMap* m, const void* ptr, bool f32, long long cols, long long rows, long long ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return MDM_E_CUDA; }
  const int es_bytes = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(tile) failed: %d (ptr=%p f32=%d cols=%lld rows=%lld ld=%lld)", (int)r, ptr, (int)f32, cols, rows, ld);
    return MDM_E_CUDA;
  }
  return MDM_OK;
}

// split `count` pixels (a power of two <= 128) of an [N][H][W] raster into a TMA box (bw, bh, bn)
static void pixel_box(int count, int H, int W, int* bw, int* bh, int* bn) {
  *bw = W < count ? W : count;
  int rest = count / *bw;
  *bh = H < rest ? H : rest;
  *bn = rest / *bh;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

template <int kMode, bool kHalo, int kMT, bool kStats>
static cudaError_t set_smem_attr() {
  cudaError_t e = cudaFuncSetAttribute(igemm_kernel<kMode, kHalo, kMT, kStats, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, IGEMM_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(igemm_kernel<kMode, kHalo, kMT, kStats, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, IGEMM_SMEM);
}
static int ensure_smem_attr() {
  static bool done = false;
  if (!done) {
    MDM_CUDA((set_smem_attr<0, false, 1, false>()));
    MDM_CUDA((set_smem_attr<0, false, 2, false>()));
    MDM_CUDA((set_smem_attr<0, true, 1, false>()));
    MDM_CUDA((set_smem_attr<0, true, 2, false>()));
    MDM_CUDA((set_smem_attr<1, false, 2, false>()));
    MDM_CUDA((set_smem_attr<0, true, 2, true>()));
    MDM_CUDA((set_smem_attr<0, false, 2, true>()));
    MDM_CUDA((set_smem_attr<0, false, 1, true>()));
    MDM_CUDA(cudaFuncSetAttribute(igemm_kernel<0, true, 2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, IGEMM_SMEM));
    MDM_CUDA(cudaFuncSetAttribute(igemm_kernel<0, true, 2, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, IGEMM_SMEM));
    MDM_CUDA(cudaFuncSetAttribute(igemm_kernel<0, true, 2, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, IGEMM_SMEM));
    MDM_CUDA(cudaFuncSetAttribute(igemm_kernel<0, true, 2, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, IGEMM_SMEM));
    done = true;
  }
  return MDM_OK;
}

// taps of a k x k filter that can touch a valid input pixel (a 3x3 filter on a 1x1 map only has its centre):
// out extent H x W, input extent (H*stride) x (W*stride), input coordinate = out*stride + d
static void fill_taps(IgemmArgs& a, int seg, int ksize, bool flip, int H, int W, int stride) {
  const int half = ksize / 2;
  int n = 0;
  for (int r = 0; r < ksize; ++r)
    for (int s = 0; s < ksize; ++s) {
      const int dh = flip ? half - r : r - half, dw = flip ? half - s : s - half;
      bool okh = false, okw = false;
      for (int h = 0; h < H && !okh; ++h) okh = h * stride + dh >= 0 && h * stride + dh < H * stride;
      for (int w = 0; w < W && !okw; ++w) okw = w * stride + dw >= 0 && w * stride + dw < W * stride;
      if (!okh || !okw) continue;
      a.tap_dh[seg][n] = (signed char)dh;
      a.tap_dw[seg][n] = (signed char)dw;
      a.tap_b[seg][n] = (signed char)(r * ksize + s);
      ++n;
    }
  a.seg_taps[seg] = n;
}


static int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// persistent grid: one CTA per SM -- minus the SMs the caller reserved for kernels that run concurrently on another
// stream (mdm_reserve_sms: the sampler's serial mt19937 mask / noise generator shares the GPU with the denoiser, and a
// static work list makes the whole GEMM wait for the CTA that shares its SM: 45.3 -> 42.5 ms per denoising step at
// 256x3x128x128 with one SM left free); MDM_IGEMM_MAX_CTAS caps it from the environment
static std::atomic<int> g_reserved_sms{0};
static int igemm_grid_cap() {
  static const int max_ctas = env_flag("MDM_IGEMM_MAX_CTAS", kNumSMs);
  int cap = max_ctas < kNumSMs && max_ctas > 0 ? max_ctas : kNumSMs;
  const int r = g_reserved_sms.load(std::memory_order_relaxed);
  if (r > 0 && kNumSMs - r < cap) cap = kNumSMs - r;
  return cap < 1 ? 1 : cap;
}

// dynamic work distribution: pairs of ints from a caller-allocated, zeroed device buffer (mdm_set_sched_workspace),
// handed out round-robin so that GEMMs in flight on different streams never share a pair; every launch leaves its pair
// zeroed again.  MDM_IGEMM_DYNAMIC: 0 = static lists, 1 (default) = dynamic wherever a CTA runs >= 2 items, 2 = every
// launch (tests).  Measured (B200, one gpurun call): training step 3x32x32 b128 8.12 ms dynamic vs 8.10 static (noise);
// sampling 256x3x128x128 with the mt19937 generator co-running and NO reserved SM 42.5 ms per step dynamic vs 44.5
// static (43.2 static with one SM reserved).
static int* g_sched_base = nullptr;
static int g_sched_pairs = 0, g_sched_dev = -1;
// Pairs are partitioned by STREAM: launches of one stream are ordered (the last CTA re-arms the pair before the next
// kernel of that stream can draw from it), launches of different streams -- the wgrad side stream, the sampler's
// generator stream, graph branches captured from them -- may run concurrently and must never share a pair.  Every
// stream seen gets its own class of pairs (round robin inside the class); when the classes run out the launch falls
// back to its static work list.
constexpr int SCHED_CLASSES = 32;
static std::mutex g_sched_mu;
static void* g_sched_streams[SCHED_CLASSES];
static unsigned g_sched_seq[SCHED_CLASSES];
static int g_sched_nstreams = 0;
static int* sched_pair_for(void* stream) {
  std::lock_guard<std::mutex> lk(g_sched_mu);
  const int per = g_sched_pairs / SCHED_CLASSES;
  if (per < 1) return nullptr;
  int cls = -1;
  for (int i = 0; i < g_sched_nstreams; ++i)
    if (g_sched_streams[i] == stream) { cls = i; break; }
  if (cls < 0) {
    if (g_sched_nstreams == SCHED_CLASSES) return nullptr;
    cls = g_sched_nstreams++;
    g_sched_streams[cls] = stream;
    g_sched_seq[cls] = 0;
  }
  return g_sched_base + 2 * (cls * per + (int)(g_sched_seq[cls]++ % (unsigned)per));
}

template <int kMode, bool kHalo, int kMT, bool kStats>
static void launch_variant(bool dyn, int grid, int block, cudaStream_t st, const CUtensorMap& mA0, const CUtensorMap& mB0,
                           const CUtensorMap& mA1, const CUtensorMap& mB1, const CUtensorMap& mC, const CUtensorMap& mD,
                           const IgemmArgs& a) {
  if (dyn) launch_pdl(igemm_kernel<kMode, kHalo, kMT, kStats, true>, dim3(grid), dim3(block), IGEMM_SMEM, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else launch_pdl(igemm_kernel<kMode, kHalo, kMT, kStats, false>, dim3(grid), dim3(block), IGEMM_SMEM, st, mA0, mB0, mA1, mB1, mC, mD, a);
}

static int launch_igemm(const CUtensorMap& mA0, const CUtensorMap& mB0, const CUtensorMap& mA1, const CUtensorMap& mB1,
                        const CUtensorMap& mC, const CUtensorMap& mD, IgemmArgs& a, void* stream) {
  const int cap = igemm_grid_cap();
  const int grid = a.num_work < cap ? a.num_work : cap;
  // eight epilogue warps once every CTA runs at least five work items (MDM_IGEMM_EPI8_MIN_ITEMS work items in total)
  static const int epi8_min = env_flag("MDM_IGEMM_EPI8_MIN_ITEMS", 5 * kNumSMs);
  const int IGEMM_BLOCK = a.num_work >= epi8_min ? IGEMM_THREADS : IGEMM_THREADS_NARROW;
  cudaStream_t st = as_stream(stream);
  const int dyn_enabled = env_flag("MDM_IGEMM_DYNAMIC", 1);   // read per call: the tests switch modes
  a.sched = nullptr;
  if (dyn_enabled && g_sched_base != nullptr && (dyn_enabled >= 2 || a.num_work >= 2 * grid)) {
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev == g_sched_dev) a.sched = sched_pair_for(stream);
  }
  const bool dyn = a.sched != nullptr;
  if (a.gn_coef) {   // GroupNorm + SiLU folded into the operand path: eight epilogue + eight transform warps (608 threads)
    if (a.qsum && dyn) launch_pdl(igemm_kernel<0, true, 2, true, true, true>, dim3(grid), dim3(IGEMM_THREADS_NORM), IGEMM_SMEM, st, mA0, mB0, mA1, mB1, mC, mD, a);
    else if (a.qsum) launch_pdl(igemm_kernel<0, true, 2, true, false, true>, dim3(grid), dim3(IGEMM_THREADS_NORM), IGEMM_SMEM, st, mA0, mB0, mA1, mB1, mC, mD, a);
    else if (dyn) launch_pdl(igemm_kernel<0, true, 2, false, true, true>, dim3(grid), dim3(IGEMM_THREADS_NORM), IGEMM_SMEM, st, mA0, mB0, mA1, mB1, mC, mD, a);
    else launch_pdl(igemm_kernel<0, true, 2, false, false, true>, dim3(grid), dim3(IGEMM_THREADS_NORM), IGEMM_SMEM, st, mA0, mB0, mA1, mB1, mC, mD, a);
    MDM_LAUNCH_CHECK();
    return MDM_OK;
  }
  if (a.mode == 1) launch_variant<1, false, 2, false>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else if (a.qsum && a.halo) launch_variant<0, true, 2, true>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else if (a.qsum && a.mt == 2) launch_variant<0, false, 2, true>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else if (a.qsum) launch_variant<0, false, 1, true>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else if (a.halo && a.mt == 2) launch_variant<0, true, 2, false>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else if (a.halo) launch_variant<0, true, 1, false>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else if (a.mt == 2) launch_variant<0, false, 2, false>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  else launch_variant<0, false, 1, false>(dyn, grid, IGEMM_BLOCK, st, mA0, mB0, mA1, mB1, mC, mD, a);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

// can this layer run in halo mode?  3x3, stride 1, bf16 output only.  returns 2: 16 x 16 patches (256-row work
// items), 1: 8 x 16 patches, 0: plain stages.
// MDM_IGEMM_HALO: 0 = never, 1 = 8 x 16 patches wherever possible, 2 (default) = 16 x 16 patches wherever the map is a
// multiple of 16 x 16 and at least 96 work items exist.  Measured (bench_conv.py, B200, after the issue loops moved
// to uniform registers): 16 x 16 patches beat plain 256-row stages on every eligible shape by 0..19 % -- e.g.
// 128x32x32 128->128 fprop 920 vs 824 TFLOP/s, 256->128 dgrad 1316 vs 1163; 64x128x128 128->128 1225 vs 1029 | 1334
// vs 1157; 256->128 1357 vs 1161 | 1382 vs 1193 -- because they move 2.3x fewer operand bytes from L2 to the SMs.
static int halo_ok(const mdm_conv_args* c, const void* out, int n_total, int k_chunks) {
  (void)k_chunks;
  const int enabled = env_flag("MDM_IGEMM_HALO", 2);                   // read per call: the tests switch modes
  const int force = env_flag("MDM_IGEMM_HALO_FORCE", 0);               // tests: every eligible layer, however small
  if (c->up2x || c->gn_coef) return 2;                                 // (checked by the caller: H, W multiples of 16)
  if (!enabled || c->ksize != 3 || c->stride != 1 || out == nullptr || c->y_f32 != nullptr) return 0;
  if (enabled == 1) return (c->H % PATCH_H == 0 && c->W % PATCH_W == 0) ? 1 : 0;
  if (c->H % 16 == 0 && c->W % 16 == 0) {
    const int num_n = (n_total + TILE_N - 1) / TILE_N;
    const long long items = (long long)c->N * (c->H / 16) * (c->W / 16) * num_n;
    if (force || items >= 96) return 2;
  }
  return 0;
}

// fprop / dgrad share this: build the operand / epilogue maps, choose a K split for small grids, launch,
// finish split-K partials.  act = the A-side activation (x for fprop, dy for dgrad) with `act_c` channels.
static int run_activation_gemm(IgemmArgs& a, const mdm_conv_args* c, const void* act, long long ld_act, int act_c,
                               int act_stride, const CUtensorMap& mB0, const CUtensorMap* mB1, void* out, long long ld_out,
                               void* stream) {
  int rc;
  CUtensorMap mA0, mA1, mC, mD;
  int halo_kind = halo_ok(c, out, a.N_total, a.seg_kc[0]);   // 0: plain, 1: 8 x 16 patches, 2: 16 x 16 patches (256-row items)
  if (a.qsum && halo_kind == 1) halo_kind = 0;                 // (no statistics instantiation of the 8 x 16 patch kernel)
  a.halo = halo_kind ? 1 : 0;
  a.num_n = (a.N_total + TILE_N - 1) / TILE_N;   // a narrow last tile is clipped by the TMA store / guarded in the epilogue
  int num_m = a.halo ? c->N * (c->H / PATCH_H) * (c->W / PATCH_W) : (a.M_total + TILE_M - 1) / TILE_M;
  // 256-row tiles once they still fill every SM (both halves share each weight tile: 25 % less operand traffic)
  static const int mt2_enabled = env_flag("MDM_IGEMM_M256", 1);
  a.mt = 1;
  if (halo_kind == 2) {
    a.mt = 2;
    num_m /= 2;
  } else if (mt2_enabled && !a.halo && out != nullptr && ((num_m + 1) / 2) * a.num_n >= env_flag("MDM_IGEMM_M256_MIN", kNumSMs)) {
    a.mt = 2;
    num_m = (num_m + 1) / 2;
  }
  int bw, bh, bn;
  if (a.halo) { bw = PATCH_W; bh = PATCH_H; bn = 1; }
  else pixel_box(128, c->H, c->W, &bw, &bh, &bn);
  if (halo_kind == 2) rc = make_act_map(&mA0, act, ld_act, act_c, c->W, c->H, c->N, HALO2_W, HALO2_H, 1, 1);
  else if (a.halo) rc = make_act_map(&mA0, act, ld_act, act_c, c->W, c->H, c->N, HALO_W, HALO_H, 1, 1);
  else rc = make_act_map(&mA0, act, ld_act, act_c, c->W * act_stride, c->H * act_stride, c->N, bw, bh, bn, act_stride);
  if (rc) return rc;
  mA1 = mA0;
  if (mB1) {   // fused 1x1 shortcut segment: plain tiles (a patch in halo mode)
    rc = make_act_map(&mA1, c->x2, c->ld_x2, c->cin2, c->W, c->H, c->N, bw, bh, bn, 1);
    if (rc) return rc;
  }
  const int tiles = num_m * a.num_n;
  a.iters_total = a.seg_taps[0] * a.seg_kc[0] + (a.nseg > 1 ? a.seg_taps[1] * a.seg_kc[1] : 0);
  int splits = 1;
  const long long ws_need = (long long)a.M_total * a.N_total;
  if (!a.halo && !a.qsum && c->splitk_ws && c->splitk_ws_floats >= ws_need && tiles * 2 <= kNumSMs && a.iters_total >= 8) {
    static const int min_iters = env_flag("MDM_IGEMM_SPLITK_MIN_ITERS", 4);   // K iterations (64 channels x 1 tap) per split, at least
    splits = kNumSMs / tiles;
    if (splits > a.iters_total / min_iters) splits = a.iters_total / min_iters;
    if (splits < 1) splits = 1;
  }
  a.iters_per_split = (a.iters_total + splits - 1) / splits;
  a.splits = (a.iters_total + a.iters_per_split - 1) / a.iters_per_split;
  a.num_work = tiles * a.splits;
  a.pair_work = 0;
  // halo mode with 256-row items: a last partial wave of at most half the SMs is re-cut into single 128-row items
  // (two per patch), so it costs half a wave instead of a whole one (128x32x32: 512 items = 3.46 waves -> 3.46 + ...
  // 444 pair items + 136 single items = 3.5 waves instead of 4)
  static const int tail_split = env_flag("MDM_IGEMM_TAIL_SPLIT", 1);
  if (tail_split && halo_kind == 2 && a.splits == 1) {
    const int grid_cap = igemm_grid_cap();
    const int rem = tiles % grid_cap;
    if (tiles > grid_cap && rem > 0 && 2 * rem <= grid_cap) {
      a.pair_work = tiles - rem;
      a.num_work = a.pair_work + 2 * rem;
    }
  }
  const void* c_ptr = c->accumulate ? out : c->resid;
  const long long c_ld = c->accumulate ? ld_out : c->ld_resid;
  const CUtensorMap& mB1r = mB1 ? *mB1 : mB0;
  if (a.splits == 1) {
    a.epi = 0;
    a.store_bf16 = out != nullptr ? 1 : 0;
    a.has_c = (c_ptr != nullptr && out != nullptr) ? 1 : 0;
    mD = mA0;   // placeholder when nothing is TMA-stored (fp32-only output of the tiny Linear layers)
    if (out) {
      if (c->up2x) rc = make_parity_map(&mD, out, ld_out, a.N_total, c->W, c->H, c->N, c->up_a, c->up_b, PATCH_W, PATCH_H);
      else if (a.halo) rc = make_act_map(&mD, out, ld_out, a.N_total, c->W, c->H, c->N, PATCH_W, PATCH_H, 1, 1);
      else rc = make_tile_map(&mD, out, false, a.N_total, a.M_total, ld_out);
      if (rc) return rc;
    }
    mC = mD;
    if (a.has_c) {
      if (a.halo) rc = make_act_map(&mC, c_ptr, c_ld, a.N_total, c->W, c->H, c->N, PATCH_W, PATCH_H, 1, 1);
      else rc = make_tile_map(&mC, c_ptr, false, a.N_total, a.M_total, c_ld);
      if (rc) return rc;
    }
    return launch_igemm(mA0, mB0, mA1, mB1r, mC, mD, a, stream);
  }
  // split-K: fp32 partials reduce-added into the (zero) workspace, then one elementwise pass
  a.epi = 1;
  a.has_c = 0;
  rc = make_tile_map(&mD, c->splitk_ws, true, a.N_total, a.M_total, a.N_total);
  if (rc) return rc;
  mC = mD;
  rc = launch_igemm(mA0, mB0, mA1, mB1r, mC, mD, a, stream);
  if (rc) return rc;
  const long long total4 = ws_need / 4;
  const int blocks = (int)((total4 + 255) / 256);
  launch_pdl(splitk_finalize_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), c->splitk_ws, a.M_total, a.N_total, a.bias, a.bias2, a.rowvec,
                                                                a.ld_rowvec, a.rows_per_vec, (const __nv_bfloat16*)c_ptr, c_ld,
                                                                (__nv_bfloat16*)out, ld_out, c->y_f32);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

}  // namespace mdm

using namespace mdm;

extern "C" {

int mdm_set_sched_workspace(void* zeroed_ints, int n_ints) {
  if (zeroed_ints == nullptr || n_ints < 2) {   // unregister: static work lists everywhere
    std::lock_guard<std::mutex> lk(g_sched_mu);
    g_sched_base = nullptr;
    g_sched_pairs = 0;
    g_sched_dev = -1;
    g_sched_nstreams = 0;
    return MDM_OK;
  }
  int dev = -1;
  MDM_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_sched_mu);
  g_sched_nstreams = 0;
  g_sched_base = static_cast<int*>(zeroed_ints);
  g_sched_pairs = n_ints / 2;
  g_sched_dev = dev;
  return MDM_OK;
}

int mdm_reserve_sms(int n) {
  if (n < 0) return g_reserved_sms.load(std::memory_order_relaxed);   // query
  if (n > kNumSMs - 1) n = kNumSMs - 1;
  return g_reserved_sms.exchange(n, std::memory_order_relaxed);
}

int mdm_conv_fprop(const mdm_conv_args* c, void* stream) {
  MDM_CHECK_ARG(c && c->x && c->w && (c->y || c->y_f32), "conv_fprop: NULL pointer");
  MDM_CHECK_ARG(c->ksize == 1 || c->ksize == 3, "conv_fprop: ksize must be 1 or 3");
  MDM_CHECK_ARG(c->stride == 1 || c->stride == 2, "conv_fprop: stride must be 1 or 2");
  MDM_CHECK_ARG(c->cin % 64 == 0 && c->cout % 32 == 0, "conv_fprop: cin %% 64 / cout %% 32 (got %d, %d)", c->cin, c->cout);
  MDM_CHECK_ARG(is_pow2(c->H) && is_pow2(c->W), "conv_fprop: output H, W must be powers of two (got %d x %d)", c->H, c->W);
  MDM_CHECK_ARG(c->ld_x % 8 == 0 && c->ld_y % 8 == 0, "conv_fprop: channel strides must be multiples of 8");
  int rc = ensure_smem_attr();
  if (rc) return rc;
  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 0;
  a.nseg = 1;
  if (c->up2x) {
    // fused nearest-2x upsample: parity (up_a, up_b) of the output = a 2x2 convolution of the low-resolution input with
    // offsets {up_a - 1, up_a} x {up_b - 1, up_b} (always inside the 18 x 18 halo) and pre-summed weights [cout][4][cin]
    MDM_CHECK_ARG(c->ksize == 3 && c->stride == 1 && c->y && !c->y_f32 && !c->x2 && !c->resid && !c->accumulate && !c->rowvec,
                  "conv_fprop(up2x): plain 3x3 stride-1 layer with a bf16 output only");
    MDM_CHECK_ARG(c->H % 16 == 0 && c->W % 16 == 0 && (c->up_a | c->up_b) >= 0 && c->up_a <= 1 && c->up_b <= 1,
                  "conv_fprop(up2x): low-resolution H, W must be multiples of 16 (got %d x %d), parity in {0,1}^2", c->H, c->W);
    int n = 0;
    for (int u = 0; u < 2; ++u)
      for (int v = 0; v < 2; ++v, ++n) {
        a.tap_dh[0][n] = (signed char)(c->up_a - 1 + u);
        a.tap_dw[0][n] = (signed char)(c->up_b - 1 + v);
        a.tap_b[0][n] = (signed char)(u * 2 + v);
      }
    a.seg_taps[0] = 4;
  } else {
    fill_taps(a, 0, c->ksize, false, c->H, c->W, c->stride);
  }
  a.seg_kc[0] = c->cin / 64;
  a.a_stride[0] = c->stride;
  a.H = c->H; a.W = c->W;
  a.M_total = c->N * c->H * c->W;
  a.N_total = c->cout;
  a.bias = c->bias; a.bias2 = c->bias2; a.rowvec = c->rowvec; a.ld_rowvec = c->ld_rowvec; a.rows_per_vec = c->H * c->W;
  a.out_f32 = c->y_f32;
  if (c->gn_coef) {
    MDM_CHECK_ARG(c->ksize == 3 && c->stride == 1 && c->y && !c->y_f32 && !c->up2x && c->H % 16 == 0 && c->W % 16 == 0 &&
                      c->cin % 64 == 0 && c->cin >= 128,
                  "conv_fprop(gn_coef): 3x3 stride-1 layer with a bf16 output on a map that is a multiple of 16 x 16, cin a "
                  "multiple of 64 and >= 128");
    a.gn_coef = c->gn_coef;
    a.a_raw = c->x;
    a.a_ld = c->ld_x;
    a.a_c = c->cin;
  }
  if (c->qsum) {
    MDM_CHECK_ARG(c->y != nullptr && ((long long)c->H * c->W) % 128 == 0,
                  "conv_fprop: fused GroupNorm statistics need a bf16 output and H*W %% 128 == 0 (got %d x %d)", c->H, c->W);
    a.qsum = c->qsum;
  }
  CUtensorMap mB0, mB1;
  rc = make_w_map(&mB0, c->w, c->cin, c->up2x ? 4 : c->ksize * c->ksize, c->cout, 128);
  if (rc) return rc;
  if (c->x2) {  // fused 1x1 shortcut: extra K segment on a second activation / weight pair
    MDM_CHECK_ARG(c->w2 && c->cin2 % 64 == 0 && c->ld_x2 % 8 == 0, "conv_fprop: bad shortcut segment");
    a.nseg = 2;
    fill_taps(a, 1, 1, false, c->H, c->W, 1);
    a.seg_kc[1] = c->cin2 / 64;
    a.a_stride[1] = 1;
    rc = make_w_map(&mB1, c->w2, c->cin2, 1, c->cout, 128);
    if (rc) return rc;
  }
  return run_activation_gemm(a, c, c->x, c->ld_x, c->cin, c->stride, mB0, c->x2 ? &mB1 : nullptr, c->y, c->ld_y, stream);
}

// dgrad of a stride-1 conv: dx[pix][ci] (+)= sum_tap sum_co dy[pix - tap][co] * w[co][tap][ci]
int mdm_conv_dgrad(const mdm_conv_args* c, void* stream) {
  MDM_CHECK_ARG(c && c->x && c->w && (c->y || c->y_f32), "conv_dgrad: NULL pointer");
  MDM_CHECK_ARG(c->ksize == 1 || c->ksize == 3, "conv_dgrad: ksize must be 1 or 3");
  MDM_CHECK_ARG(c->stride == 1, "conv_dgrad: stride-2 layers go through zero insertion first");
  MDM_CHECK_ARG(c->cout % 64 == 0 && c->cin % 128 == 0, "conv_dgrad: cout %% 64 / cin %% 128 (got %d, %d)", c->cout, c->cin);
  MDM_CHECK_ARG(is_pow2(c->H) && is_pow2(c->W), "conv_dgrad: H, W must be powers of two");
  MDM_CHECK_ARG(c->ld_x % 8 == 0 && c->ld_y % 8 == 0, "conv_dgrad: channel strides must be multiples of 8");
  int rc = ensure_smem_attr();
  if (rc) return rc;
  // here x = dy [pix][cout] (input), y = dx [pix][cin] (output); w = [cout][taps][cin_total]
  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 0;
  a.nseg = 1;
  fill_taps(a, 0, c->ksize, true, c->H, c->W, 1);
  a.seg_kc[0] = c->cout / 64;
  a.a_stride[0] = 1;
  a.b_mn_major = 1;
  a.H = c->H; a.W = c->W;
  a.M_total = c->N * c->H * c->W;
  a.N_total = c->cin;
  a.out_f32 = c->y_f32;
  a.rows_per_vec = 1;
  CUtensorMap mB0;
  // weight viewed as [K = co][tap][N = ci] with ci contiguous: box {64 ci, 1, 64 co}
  // w_cols: full row length of the packed weight (>= cin when cin is a slice); w_col0 offsets the slice
  rc = make_w_map(&mB0, (const __nv_bfloat16*)c->w + c->w_col0, c->w_cols ? c->w_cols : c->cin, c->ksize * c->ksize, c->cout, 64);
  if (rc) return rc;
  return run_activation_gemm(a, c, c->x, c->ld_x, c->cout, 1, mB0, nullptr, c->y, c->ld_y, stream);
}

// wgrad: dw[co][tap][ci] += sum_pix dy[pix][co] * x[pix*stride + tap][ci]   (fp32, TMA reduce-add, split over pixels)
int mdm_conv_wgrad(const mdm_conv_args* c, void* stream) {
  MDM_CHECK_ARG(c && c->x && c->y && c->dw, "conv_wgrad: NULL pointer");
  MDM_CHECK_ARG(c->ksize == 1 || c->ksize == 3, "conv_wgrad: ksize must be 1 or 3");
  MDM_CHECK_ARG(c->stride == 1 || c->stride == 2, "conv_wgrad: stride must be 1 or 2");
  MDM_CHECK_ARG(c->cin % 64 == 0 && c->cout % 64 == 0, "conv_wgrad: cin, cout %% 64 (got %d, %d)", c->cin, c->cout);
  MDM_CHECK_ARG(is_pow2(c->H) && is_pow2(c->W), "conv_wgrad: H, W must be powers of two");
  MDM_CHECK_ARG(c->w_col0 % 4 == 0, "conv_wgrad: w_col0 must be a multiple of 4");
  int rc = ensure_smem_attr();
  if (rc) return rc;
  // x = layer input [N][H*s][W*s][cin], y = dy [N][H][W][cout]
  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 1;
  a.epi = 1;
  fill_taps(a, 0, c->ksize, false, c->H, c->W, c->stride);
  a.taps = a.seg_taps[0];
  a.a_stride[0] = c->stride;
  a.H = c->H; a.W = c->W;
  a.M_total = c->cout;
  a.N_total = c->cin;
  a.ci_total = c->w_cols ? c->w_cols : c->cin;
  a.w_col0 = (int)c->w_col0;
  a.dbias = c->dbias;
  a.dbias2 = c->dbias2;
  a.num_co = (c->cout + TILE_M - 1) / TILE_M;   // partial tiles: TMA zero-fills the loads and clips the reduce-add
  a.num_n = (c->cin + TILE_N - 1) / TILE_N;
  const long long pixels = (long long)c->N * c->H * c->W;
  a.iters_total = (int)((pixels + 63) / 64);
  // work item = (cout tile, up to two consecutive (cin tile, tap) entries sharing the dY tiles, K split); the first
  // item of every cout tile holds entry 0 alone plus the bias gradient.  The K (pixel) split is chosen by a cost
  // model of the static persistent schedule: waves x max(main loop, epilogue) -- never one item over a wave.
  a.mt = 2;
  const int tiles = a.num_co * (1 + (a.num_n * a.taps) / 2);
  {
    long long best = -1;
    int best_per = a.iters_total;
    const int max_split = a.iters_total;
    for (int split = 1; split <= max_split; ++split) {
      const int per = (a.iters_total + split - 1) / split;
      const int sp = (a.iters_total + per - 1) / per;
      if (sp != split) continue;                     // same schedule as a smaller split count
      if (per < 4 && split > 1) break;
      const long long waves = ((long long)tiles * sp + kNumSMs - 1) / kNumSMs;
      const long long item = per * 512LL > 4000 ? per * 512LL : 4000;   // cycles: 8 MMAs per chunk vs the reduce-add epilogue
      const long long cost = waves * item + 16LL * sp;
      if (best < 0 || cost < best) { best = cost; best_per = per; }
      if (waves > 8) break;
    }
    a.iters_per_split = best_per;
    a.splits = (a.iters_total + best_per - 1) / best_per;
    a.num_work = tiles * a.splits;
  }
  int pw, ph, pn;
  pixel_box(64, c->H, c->W, &pw, &ph, &pn);
  a.pw = pw; a.ph = ph; a.pn = pn;
  CUtensorMap mA0, mB0, mD;
  rc = make_act_map(&mA0, c->y, c->ld_y, c->cout, c->W, c->H, c->N, pw, ph, pn, 1);
  if (rc) return rc;
  rc = make_act_map(&mB0, c->x, c->ld_x, c->cin, c->W * c->stride, c->H * c->stride, c->N, pw, ph, pn, c->stride);
  if (rc) return rc;
  const long long row = (long long)c->ksize * c->ksize * a.ci_total;
  rc = make_tile_map(&mD, c->dw, true, row, c->cout, row);
  if (rc) return rc;
  a.halo = 0;
  return launch_igemm(mA0, mB0, mA0, mB0, mD, mD, a, stream);
}

}  // extern "C"
