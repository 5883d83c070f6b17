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
// One CTA = one 128x128 output tile; 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer
// (+TMEM allocator), warps 2-5 = epilogue (TMEM -> registers -> bias / time-embedding /
// residual -> bf16 NHWC store, or fp32 atomic accumulation for wgrad).  A 4-stage mbarrier ring
// connects producer and issuer; tcgen05.commit releases stages and publishes the accumulator.
#include <cuda.h>

#include "common.cuh"

namespace mdm {

constexpr int TILE_M = 128, TILE_N = 128, TILE_K = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = TILE_M * TILE_K * 2;   // 16 KB
constexpr int B_BYTES = TILE_N * TILE_K * 2;   // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int IGEMM_THREADS = 192;
constexpr int IGEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 128;

// ---- raw PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// bounded wait: a pipeline bug traps (launch fails) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
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
// MN-major tile: 64-wide MN atoms 8 KB apart (LBO), 8 K-rows per 1024 B group (SBO); k-th slice = +2 groups
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int k) { return smem_desc(base + k * 2048, 8192, 1024); }

// instruction descriptor: bf16 x bf16 -> f32, M = 128, N = 128
__host__ __device__ constexpr uint32_t make_idesc(int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

struct IgemmArgs {
  int mode;  // 0: activation GEMM (fprop / dgrad), 1: wgrad
  int nseg;
  int seg_taps[2];
  int seg_kc[2];
  signed char tap_dh[2][9], tap_dw[2][9], tap_b[2][9];
  int a_stride[2];
  int b_mn_major;
  int H, W;            // spatial extent the M tiles walk over
  int M_total, N_total;
  // epilogue (mode 0)
  __nv_bfloat16* out;
  long long ld_out;
  const float* bias;
  const float* bias2;
  const float* rowvec;
  long long ld_rowvec;
  int rows_per_vec;
  const __nv_bfloat16* resid;
  long long ld_resid;
  int accumulate;
  float* out_f32;       // optional fp32 copy [M][N_total] (row-major, ld = N_total)
  // wgrad (mode 1)
  float* dw;
  long long ld_dw;      // taps * ci_total
  int ci_total;
  int taps;
  int kchunks_total;    // 64-pixel chunks
  int kchunks_per_split;
  int pw, ph;           // 64-pixel box geometry: pw*ph*pn = 64
};

__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
             const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
             const __grid_constant__ IgemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapB0) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- iteration space -------------------------------------------------------------------------
  int total_it;
  int kc_begin = 0;
  if (args.mode == 0) {
    total_it = args.seg_taps[0] * args.seg_kc[0] + (args.nseg > 1 ? args.seg_taps[1] * args.seg_kc[1] : 0);
  } else {
    kc_begin = blockIdx.z * args.kchunks_per_split;
    int kc_end = min(kc_begin + args.kchunks_per_split, args.kchunks_total);
    total_it = max(kc_end - kc_begin, 0);
  }

  if (warp == 0 && lane == 0) {
    // ============================== TMA producer ==============================================
    if (args.mode == 0) {
      const int p0 = blockIdx.x * TILE_M;
      const int w0 = p0 % args.W, h0 = (p0 / args.W) % args.H, n0 = p0 / (args.W * args.H);
      const int ncol0 = blockIdx.y * TILE_N;
      int it = 0;
      for (int seg = 0; seg < args.nseg; ++seg) {
        const CUtensorMap* mA = seg == 0 ? &mapA0 : &mapA1;
        const CUtensorMap* mB = seg == 0 ? &mapB0 : &mapB1;
        const int st = args.a_stride[seg];
        for (int tap = 0; tap < args.seg_taps[seg]; ++tap) {
          const int dh = args.tap_dh[seg][tap], dw = args.tap_dw[seg][tap], tb = args.tap_b[seg][tap];
          for (int kc = 0; kc < args.seg_kc[seg]; ++kc, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* a_dst = smem + s * STAGE_BYTES;
            uint8_t* b_dst = a_dst + A_BYTES;
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_4d(mA, a_dst, &full_bar[s], kc * TILE_K, w0 * st + dw, h0 * st + dh, n0);
            if (!args.b_mn_major) {
              tma_load_3d(mB, b_dst, &full_bar[s], kc * TILE_K, tb, ncol0);
            } else {
              tma_load_3d(mB, b_dst, &full_bar[s], ncol0, tb, kc * TILE_K);
              tma_load_3d(mB, b_dst + 8192, &full_bar[s], ncol0 + 64, tb, kc * TILE_K);
            }
          }
        }
      }
    } else {
      // wgrad: A = dY (MN-major, M = co), B = X shifted by the tap (MN-major, N = ci)
      const int co0 = blockIdx.x * TILE_M;
      const int ci_tile = blockIdx.y / args.taps, tap = blockIdx.y % args.taps;
      const int ci0 = ci_tile * TILE_N;
      const int dh = args.tap_dh[0][tap], dw = args.tap_dw[0][tap];
      const int st = args.a_stride[0];
      for (int it = 0; it < total_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int p0 = (kc_begin + it) * TILE_K;  // 64 pixels of dY
        const int w0 = p0 % args.W, h0 = (p0 / args.W) % args.H, n0 = p0 / (args.W * args.H);
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_4d(&mapA0, a_dst, &full_bar[s], co0, w0, h0, n0);
        tma_load_4d(&mapA0, a_dst + 8192, &full_bar[s], co0 + 64, w0, h0, n0);
        tma_load_4d(&mapB0, b_dst, &full_bar[s], ci0, w0 * st + dw, h0 * st + dh, n0);
        tma_load_4d(&mapB0, b_dst + 8192, &full_bar[s], ci0 + 64, w0 * st + dw, h0 * st + dh, n0);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ============================== MMA issuer ================================================
    const int a_mn = args.mode == 1 ? 1 : 0;
    const int b_mn = args.mode == 1 ? 1 : args.b_mn_major;
    const uint32_t idesc = make_idesc(a_mn, b_mn);
    for (int it = 0; it < total_it; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
      for (int k = 0; k < TILE_K / 16; ++k) {
        const uint64_t da = a_mn ? desc_mnmajor(a_base, k) : desc_kmajor(a_base, k);
        const uint64_t db = b_mn ? desc_mnmajor(b_base, k) : desc_kmajor(b_base, k);
        umma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
      }
      umma_commit(&empty_bar[s]);
    }
    umma_commit(tmem_full_bar);
  } else if (warp >= 2) {
    // ============================== epilogue ==================================================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    if (total_it > 0) {
      mbar_wait(tmem_full_bar, 0);
      tcgen05_fence_after();
    }
    uint32_t v[32];
    if (args.mode == 0) {
      const long long p = (long long)blockIdx.x * TILE_M + row;
      const bool valid = p < args.M_total;
      const int ncol0 = blockIdx.y * TILE_N;
      const float* rv = (args.rowvec && valid) ? args.rowvec + (p / args.rows_per_vec) * args.ld_rowvec : nullptr;
#pragma unroll 1
      for (int cc = 0; cc < TILE_N / 32; ++cc) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 32, v);
        if (!valid) continue;
        const int col = ncol0 + cc * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (args.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(args.bias + col + j);
        }
        if (args.bias2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(args.bias2 + col + j);
        }
        if (rv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(rv + col + j);
        }
        if (args.resid) {
          const uint4* r4 = reinterpret_cast<const uint4*>(args.resid + p * args.ld_resid + col);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 u = r4[j4];
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 t = __bfloat1622float2(h2[e]);
              f[j4 * 8 + e * 2] += t.x;
              f[j4 * 8 + e * 2 + 1] += t.y;
            }
          }
        }
        __nv_bfloat16* optr = args.out ? args.out + p * args.ld_out + col : nullptr;
        if (args.accumulate && optr) {
          const uint4* o4 = reinterpret_cast<const uint4*>(optr);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 u = o4[j4];
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 t = __bfloat1622float2(h2[e]);
              f[j4 * 8 + e * 2] += t.x;
              f[j4 * 8 + e * 2 + 1] += t.y;
            }
          }
        }
        if (optr) {
          uint4* o4 = reinterpret_cast<uint4*>(optr);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[j4 * 8 + e * 2], f[j4 * 8 + e * 2 + 1]);
            o4[j4] = u;
          }
        }
        if (args.out_f32) {
          float4* o = reinterpret_cast<float4*>(args.out_f32 + p * (long long)args.N_total + col);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) o[j4] = make_float4(f[j4 * 4], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]);
        }
      }
    } else if (total_it > 0) {
      const int ci_tile = blockIdx.y / args.taps, tap = blockIdx.y % args.taps;
      const long long co = (long long)blockIdx.x * TILE_M + row;
      const bool valid = co < args.M_total;
      float* dst = args.dw + co * args.ld_dw + (long long)tap * args.ci_total + ci_tile * TILE_N;
#pragma unroll 1
      for (int cc = 0; cc < TILE_N / 32; ++cc) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 32, v);
        if (!valid) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (ci_tile * TILE_N + cc * 32 + j < args.N_total) atomicAdd(dst + cc * 32 + j, __uint_as_float(v[j]));
        }
      }
    }
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
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

// split `count` pixels (a power of two <= 128) of an [N][H][W] raster into a TMA box (bw, bh, bn)
static void pixel_box(int count, int H, int W, int* bw, int* bh, int* bn) {
  *bw = W < count ? W : count;
  int rest = count / *bw;
  *bh = H < rest ? H : rest;
  *bn = rest / *bh;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static int ensure_smem_attr() {
  static bool done = false;
  if (!done) {
    MDM_CUDA(cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IGEMM_SMEM));
    done = true;
  }
  return MDM_OK;
}

static void fill_taps(IgemmArgs& a, int seg, int ksize, bool flip) {
  a.seg_taps[seg] = ksize * ksize;
  const int half = ksize / 2;
  for (int r = 0; r < ksize; ++r)
    for (int s = 0; s < ksize; ++s) {
      const int t = r * ksize + s;
      a.tap_dh[seg][t] = (signed char)(flip ? half - r : r - half);
      a.tap_dw[seg][t] = (signed char)(flip ? half - s : s - half);
      a.tap_b[seg][t] = (signed char)t;
    }
}

}  // namespace mdm

using namespace mdm;

extern "C" {

int mdm_conv_fprop(const mdm_conv_args* c, void* stream) {
  MDM_CHECK_ARG(c && c->x && c->w && (c->y || c->y_f32), "conv_fprop: NULL pointer");
  MDM_CHECK_ARG(c->ksize == 1 || c->ksize == 3, "conv_fprop: ksize must be 1 or 3");
  MDM_CHECK_ARG(c->stride == 1 || c->stride == 2, "conv_fprop: stride must be 1 or 2");
  MDM_CHECK_ARG(c->cin % 64 == 0 && c->cout % 128 == 0, "conv_fprop: cin %% 64 / cout %% 128 (got %d, %d)", c->cin, c->cout);
  MDM_CHECK_ARG(is_pow2(c->H) && is_pow2(c->W), "conv_fprop: output H, W must be powers of two (got %d x %d)", c->H, c->W);
  MDM_CHECK_ARG(c->ld_x % 8 == 0 && c->ld_y % 8 == 0, "conv_fprop: channel strides must be multiples of 8");
  int rc = ensure_smem_attr();
  if (rc) return rc;
  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 0;
  a.nseg = 1;
  fill_taps(a, 0, c->ksize, false);
  a.seg_kc[0] = c->cin / 64;
  a.a_stride[0] = c->stride;
  a.H = c->H; a.W = c->W;
  a.M_total = c->N * c->H * c->W;
  a.N_total = c->cout;
  a.out = (__nv_bfloat16*)c->y; a.ld_out = c->ld_y;
  a.bias = c->bias; a.bias2 = c->bias2; a.rowvec = c->rowvec; a.ld_rowvec = c->ld_rowvec; a.rows_per_vec = c->H * c->W;
  a.resid = (const __nv_bfloat16*)c->resid; a.ld_resid = c->ld_resid;
  a.accumulate = c->accumulate;
  a.out_f32 = c->y_f32;
  int bw, bh, bn;
  pixel_box(128, c->H, c->W, &bw, &bh, &bn);
  CUtensorMap mA0, mB0, mA1, mB1;
  rc = make_act_map(&mA0, c->x, c->ld_x, c->cin, c->W * c->stride, c->H * c->stride, c->N, bw, bh, bn, c->stride);
  if (rc) return rc;
  rc = make_w_map(&mB0, c->w, c->cin, c->ksize * c->ksize, c->cout, 128);
  if (rc) return rc;
  mA1 = mA0; mB1 = mB0;
  if (c->x2) {  // fused 1x1 shortcut: extra K segment on a second activation / weight pair
    MDM_CHECK_ARG(c->w2 && c->cin2 % 64 == 0 && c->ld_x2 % 8 == 0, "conv_fprop: bad shortcut segment");
    a.nseg = 2;
    fill_taps(a, 1, 1, false);
    a.seg_kc[1] = c->cin2 / 64;
    a.a_stride[1] = 1;
    rc = make_act_map(&mA1, c->x2, c->ld_x2, c->cin2, c->W, c->H, c->N, bw, bh, bn, 1);
    if (rc) return rc;
    rc = make_w_map(&mB1, c->w2, c->cin2, 1, c->cout, 128);
    if (rc) return rc;
  }
  dim3 grid((a.M_total + TILE_M - 1) / TILE_M, c->cout / TILE_N, 1);
  igemm_kernel<<<grid, IGEMM_THREADS, IGEMM_SMEM, as_stream(stream)>>>(mA0, mB0, mA1, mB1, a);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
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
  fill_taps(a, 0, c->ksize, true);
  a.seg_kc[0] = c->cout / 64;
  a.a_stride[0] = 1;
  a.b_mn_major = 1;
  a.H = c->H; a.W = c->W;
  a.M_total = c->N * c->H * c->W;
  a.N_total = c->cin;
  a.out = (__nv_bfloat16*)c->y; a.ld_out = c->ld_y;
  a.resid = (const __nv_bfloat16*)c->resid; a.ld_resid = c->ld_resid;
  a.accumulate = c->accumulate;
  a.out_f32 = c->y_f32;
  int bw, bh, bn;
  pixel_box(128, c->H, c->W, &bw, &bh, &bn);
  CUtensorMap mA0, mB0;
  rc = make_act_map(&mA0, c->x, c->ld_x, c->cout, c->W, c->H, c->N, bw, bh, bn, 1);
  if (rc) return rc;
  // weight viewed as [K = co][tap][N = ci] with ci contiguous: box {64 ci, 1, 64 co}
  // w_cols: full row length of the packed weight (>= cin when cin is a slice); w_col0 offsets the slice
  rc = make_w_map(&mB0, (const __nv_bfloat16*)c->w + c->w_col0, c->w_cols ? c->w_cols : c->cin, c->ksize * c->ksize, c->cout, 64);
  if (rc) return rc;
  dim3 grid((a.M_total + TILE_M - 1) / TILE_M, c->cin / TILE_N, 1);
  igemm_kernel<<<grid, IGEMM_THREADS, IGEMM_SMEM, as_stream(stream)>>>(mA0, mB0, mA0, mB0, a);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

// wgrad: dw[co][tap][ci] += sum_pix dy[pix][co] * x[pix*stride + tap][ci]   (fp32, atomic split-K)
int mdm_conv_wgrad(const mdm_conv_args* c, void* stream) {
  MDM_CHECK_ARG(c && c->x && c->y && c->dw, "conv_wgrad: NULL pointer");
  MDM_CHECK_ARG(c->ksize == 1 || c->ksize == 3, "conv_wgrad: ksize must be 1 or 3");
  MDM_CHECK_ARG(c->stride == 1 || c->stride == 2, "conv_wgrad: stride must be 1 or 2");
  MDM_CHECK_ARG(c->cin % 128 == 0 && c->cout % 128 == 0, "conv_wgrad: cin, cout %% 128 (got %d, %d)", c->cin, c->cout);
  MDM_CHECK_ARG(is_pow2(c->H) && is_pow2(c->W), "conv_wgrad: H, W must be powers of two");
  int rc = ensure_smem_attr();
  if (rc) return rc;
  // x = layer input [N][H*s][W*s][cin], y = dy [N][H][W][cout]
  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 1;
  fill_taps(a, 0, c->ksize, false);
  a.taps = c->ksize * c->ksize;
  a.a_stride[0] = c->stride;
  a.H = c->H; a.W = c->W;
  a.M_total = c->cout;
  a.N_total = c->cin;
  a.dw = c->dw;
  a.ci_total = c->w_cols ? c->w_cols : c->cin;
  a.ld_dw = (long long)a.taps * a.ci_total;
  a.dw += c->w_col0;
  const long long pixels = (long long)c->N * c->H * c->W;
  a.kchunks_total = (int)((pixels + 63) / 64);
  // split K so that the grid fills the machine (~2 waves), at least 4 chunks per CTA
  const int tiles = (c->cout / TILE_M) * (c->cin / TILE_N) * a.taps;
  int split = (2 * kNumSMs + tiles - 1) / tiles;
  if (split < 1) split = 1;
  int per = (a.kchunks_total + split - 1) / split;
  if (per < 4) per = 4;
  if (per > a.kchunks_total) per = a.kchunks_total;
  split = (a.kchunks_total + per - 1) / per;
  a.kchunks_per_split = per;
  int pw, ph, pn;
  pixel_box(64, c->H, c->W, &pw, &ph, &pn);
  CUtensorMap mA0, mB0;
  rc = make_act_map(&mA0, c->y, c->ld_y, c->cout, c->W, c->H, c->N, pw, ph, pn, 1);
  if (rc) return rc;
  rc = make_act_map(&mB0, c->x, c->ld_x, c->cin, c->W * c->stride, c->H * c->stride, c->N, pw, ph, pn, c->stride);
  if (rc) return rc;
  dim3 grid(c->cout / TILE_M, (c->cin / TILE_N) * a.taps, split);
  igemm_kernel<<<grid, IGEMM_THREADS, IGEMM_SMEM, as_stream(stream)>>>(mA0, mB0, mA0, mB0, a);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

}  // extern "C"
