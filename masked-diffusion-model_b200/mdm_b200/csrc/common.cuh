// Shared helpers for libmdm_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../../include/mdm.h"

namespace mdm {

void set_error(const char* fmt, ...);

#define MDM_CHECK_ARG(cond, ...)             \
  do {                                       \
    if (!(cond)) {                           \
      ::mdm::set_error(__VA_ARGS__);         \
      return MDM_E_ARG;                      \
    }                                        \
  } while (0)

#define MDM_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::mdm::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MDM_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)

// every kernel launch of the library passes through here: count it (mdm_launch_count()) and check it
extern long long g_launch_count;
#define MDM_LAUNCH_CHECK()              \
  do {                                  \
    ++::mdm::g_launch_count;            \
    MDM_CUDA(cudaGetLastError());       \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: every kernel of this library is launched with the stream-serialization
// attribute relaxed and opens with MDM_PDL_ENTER(): `launch_dependents` lets the NEXT kernel of the stream be
// scheduled (its launch latency and prologue overlap this kernel's tail), `wait` blocks until the PREVIOUS grid
// has completed and its writes are visible -- so no global memory may be touched before it.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define MDM_PDL_ENTER()               \
  do {                                \
    ::mdm::pdl_launch_dependents();   \
    ::mdm::pdl_wait();                \
  } while (0)

template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  // MDM_PDL: 0 = never, 1 = every launch, 2 (default) = only launches of at most MDM_PDL_MAX_CTAS CTAs: the
  // launch-latency-bound low-resolution layers, whose grids leave most SMs idle anyway.  Measured on B200 (3x32x32
  // batch 128, captured step): off 8.84 ms, <= 64 CTAs 8.59, <= 148..1200 CTAs 8.33, every launch 8.39.
  static const int pdl_on = [] { const char* v = getenv("MDM_PDL"); return v ? atoi(v) : 2; }();
  static const unsigned pdl_max = [] { const char* v = getenv("MDM_PDL_MAX_CTAS"); return v ? (unsigned)atoi(v) : 300u; }();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_on == 1 || (pdl_on == 2 && grid.x * grid.y * grid.z <= pdl_max)) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);   // errors surface through MDM_LAUNCH_CHECK (cudaGetLastError)
}
#endif

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in every thread. `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float ld_as_float(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}

}  // namespace mdm
