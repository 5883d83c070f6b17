// Data-parallel gradient all-reduce over NVLink peer memory (training = data parallel, one
// exchange step per iteration; the reference gets it from accelerate / DDP's bucketed NCCL all-reduce behind
// `accelerator.backward`, trainer_masked.py:142).
//
// One kernel, launched on a forked stream INSIDE the captured training step: every rank maps every peer's flat fp32
// gradient buffer (CUDA IPC, one process per GPU) and
//   1. waits until the peers' gradients of the range are final            (flag exchange, system-scope release/acquire)
//   2. sums ITS 1/world slice over all ranks with peer loads in rank order (so every rank would compute the same bits)
//      and PUSHES the sum into every rank's buffer (posted peer stores: the all-gather costs no round trips)
//   3. waits until every peer has pushed its slice: the buffer is complete, and nobody reads it any more (the peers
//      may overwrite theirs after the kernel).
// Synchronisation is PER BLOCK: block b of every rank works on the same chunk of every slice, so block b only ever
// waits for block b of its peers -- no grid-wide barrier, no requirement that all blocks be resident, no dead-lock
// with the compute kernels that share the SMs.  Blocks are small (128 threads, ~100 registers, no shared memory) and fit next to a
// resident implicit-GEMM CTA (which owns the SM's shared memory and TMEM but only 45 K registers), unlike NCCL's
// channels, which wait for whole SMs: measured in round 1, the NCCL all-reduce stayed fully exposed (+0.9 ms per step at
// 8 GPUs).  The range finished by each backward segment is reduced while the next segment computes.
// Traffic per GPU: (world - 1) / world x bytes pulled + the same pushed over NVLink, 16 bytes per thread and access,
// 16 peer loads in flight per thread.
#include <cuda.h>

#include "common.cuh"

namespace mdm {

constexpr int AR_MAX_RANKS = 8;
constexpr int AR_MAX_BLOCKS = 128;
constexpr int AR_THREADS = 128;
constexpr int AR_PHASES = 2;
// flag words per rank: [AR_PHASES][AR_MAX_BLOCKS][AR_MAX_RANKS] uint32, then per-block epochs [AR_MAX_BLOCKS], then the
// stand-alone barriers of the copy-engine variant: [AR_BARRIERS][AR_MAX_RANKS] flags + [AR_BARRIERS] epochs
constexpr int AR_BARRIERS = 4;
constexpr int AR_BAR_OFF = AR_PHASES * AR_MAX_BLOCKS * AR_MAX_RANKS + AR_MAX_BLOCKS;
constexpr int AR_FLAG_WORDS = AR_BAR_OFF + AR_BARRIERS * AR_MAX_RANKS + AR_BARRIERS;

struct P2PComm {
  float* buf[AR_MAX_RANKS];       // every rank's flat gradient buffer, mapped into this process
  uint32_t* flag[AR_MAX_RANKS];   // every rank's flag array
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {   // peer (or peer-written) memory: never from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// block b of this rank tells block b of every peer that it reached `phase` of all-reduce number `epoch`, then waits for
// the same word from every peer.  Bounded: a peer that never arrives traps the kernel instead of hanging the GPU.
__device__ __forceinline__ void block_exchange(const P2PComm& c, int phase, uint32_t epoch) {
  __syncthreads();                 // every thread's stores of the previous phase are issued ...
  const int slot = (phase * AR_MAX_BLOCKS + blockIdx.x) * AR_MAX_RANKS;
  if (threadIdx.x < c.world) {
    __threadfence_system();        // ... and ordered before the flag
    st_release_sys(c.flag[threadIdx.x] + slot + c.rank, epoch);
    const uint32_t* mine = c.flag[c.rank] + slot + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > 40000000000LL) __trap();   // ~20 s
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void st_peer(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// W ranks, U elements (float4) per thread in flight: W * U peer loads are issued before the first add -- a peer load
// takes microseconds, the kernel lives on bytes in flight (measured with one element per thread: 90 GB/s on 2 GPUs).
template <int W, int U>
__global__ void __launch_bounds__(AR_THREADS) p2p_allreduce_kernel(const P2PComm c, long long offset, long long count) {
  MDM_PDL_ENTER();
  __shared__ uint32_t s_epoch;
  uint32_t* epochs = c.flag[c.rank] + AR_PHASES * AR_MAX_BLOCKS * AR_MAX_RANKS;
  if (threadIdx.x == 0) s_epoch = ++epochs[blockIdx.x];
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int r = c.rank;
  const long long n4 = count >> 2;                               // float4 elements of the range (count % 4 == 0)
  const long long slice = (n4 + W - 1) / W;                      // per rank
  const long long chunk = (slice + gridDim.x - 1) / gridDim.x;   // per block inside a slice
  const long long s0 = (long long)r * slice;                     // this rank reduces slice r
  const long long e0 = s0 + (long long)blockIdx.x * chunk;
  const long long e1 = min(min(e0 + chunk, s0 + slice), n4);
  float4* bufs[W];
#pragma unroll
  for (int p = 0; p < W; ++p) bufs[p] = reinterpret_cast<float4*>(c.buf[p] + offset);

  block_exchange(c, 0, epoch);     // the peers' gradients of this range are final
  // ---- reduce slice r in rank order (every rank would compute the same bits), push the sums to every rank -----------
  for (long long i = e0 + threadIdx.x; i < e1; i += (long long)AR_THREADS * U) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = i + (long long)u * AR_THREADS;
      if (e < e1) {
#pragma unroll
        for (int p = 0; p < W; ++p) v[u][p] = ld_peer(bufs[p] + e);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = i + (long long)u * AR_THREADS;
      if (e < e1) {
        float4 acc = v[u][0];
#pragma unroll
        for (int p = 1; p < W; ++p) { acc.x += v[u][p].x; acc.y += v[u][p].y; acc.z += v[u][p].z; acc.w += v[u][p].w; }
#pragma unroll
        for (int d = 0; d < W; ++d) {       // own copy first, then the peers, nearest neighbour first
          int p = r + d;
          if (p >= W) p -= W;
          st_peer(bufs[p] + e, acc);
        }
      }
    }
  }
  block_exchange(c, 1, epoch);     // every rank has pushed its slice: this buffer is complete, and nobody reads it any more
}

// ---- copy-engine variant -------------------------------------------------------------------------------------------
// The bulk of the bytes moves through the COPY ENGINES (cudaMemcpyAsync nodes between IPC-mapped buffers, no SM
// involved): pull the peers' copies of this rank's slice into a local staging area, reduce locally (HBM-rate, ~70 us for
// the largest range), push the reduced slice to every peer.  The SMs only run two one-warp barriers and the local
// reduction -- measured: the SM-driven kernel above slows the co-running backward in proportion to its block count.
__global__ void __launch_bounds__(32) p2p_barrier_kernel(const P2PComm c, int which) {
  MDM_PDL_ENTER();
  uint32_t* epochs = c.flag[c.rank] + AR_BAR_OFF + AR_BARRIERS * AR_MAX_RANKS;
  uint32_t epoch = 0;
  if (threadIdx.x == 0) epoch = ++epochs[which];
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const int slot = AR_BAR_OFF + which * AR_MAX_RANKS;
  if (threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(c.flag[threadIdx.x] + slot + c.rank, epoch);
    const uint32_t* mine = c.flag[c.rank] + slot + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > 40000000000LL) __trap();
    }
  }
}

// slice[e] = sum over ranks in rank order of (p == rank ? slice[e] : staging[k(p)][e]); staging holds the peers' copies
// in the order p = rank + 1, rank + 2, ... (mod world), `stride` floats apart
template <int W>
__global__ void __launch_bounds__(256) reduce_slices_kernel(float* __restrict__ slice, const float* __restrict__ staging,
                                                            long long stride, long long n4, int rank) {
  MDM_PDL_ENTER();
  float4* mine = reinterpret_cast<float4*>(slice);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    float4 v[W];
#pragma unroll
    for (int p = 0; p < W; ++p) {
      int k = p - rank - 1;
      if (k < 0) k += W;                 // position of rank p's copy in the staging area (p != rank)
      v[p] = (p == rank) ? mine[e] : *reinterpret_cast<const float4*>(staging + (long long)k * stride + 4 * e);
    }
    float4 acc = v[0];
#pragma unroll
    for (int p = 1; p < W; ++p) { acc.x += v[p].x; acc.y += v[p].y; acc.z += v[p].z; acc.w += v[p].w; }
    mine[e] = acc;
  }
}

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

}  // namespace mdm

using namespace mdm;

extern "C" {

int mdm_p2p_flag_words(void) { return AR_FLAG_WORDS; }

// IPC plumbing (one process per GPU): handle of the cudaMalloc allocation that contains `ptr` + the offset of `ptr`
// inside it; the peer opens the handle in ITS device context (peer access is enabled lazily by the runtime).
int mdm_ipc_export(const void* ptr, void* handle_out /*64 bytes*/, int64_t* offset_out) {
  MDM_CHECK_ARG(ptr && handle_out && offset_out, "ipc_export: NULL argument");
  static GetAddressRangeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
      set_error("ipc_export: cuMemGetAddressRange entry point not available");
      return MDM_E_CUDA;
    }
    fn = reinterpret_cast<GetAddressRangeFn>(p);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) { set_error("ipc_export: cuMemGetAddressRange failed"); return MDM_E_CUDA; }
  cudaIpcMemHandle_t h;
  MDM_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  *offset_out = (int64_t)((CUdeviceptr)ptr - base);
  return MDM_OK;
}

int mdm_ipc_open(const void* handle /*64 bytes*/, int64_t offset, void** ptr_out) {
  MDM_CHECK_ARG(handle && ptr_out, "ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* base = nullptr;
  MDM_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = (char*)base + offset;
  return MDM_OK;
}

static int fill_comm(P2PComm& c, const mdm_p2p_comm* comm) {
  MDM_CHECK_ARG(comm && comm->world >= 2 && comm->world <= AR_MAX_RANKS && comm->rank >= 0 && comm->rank < comm->world,
                "p2p: bad communicator");
  for (int p = 0; p < AR_MAX_RANKS; ++p) {
    c.buf[p] = (float*)comm->buf[p < comm->world ? p : 0];
    c.flag[p] = (uint32_t*)comm->flag[p < comm->world ? p : 0];
    MDM_CHECK_ARG(c.buf[p] && c.flag[p], "p2p: NULL peer pointer");
  }
  c.rank = comm->rank;
  c.world = comm->world;
  return MDM_OK;
}

// copy-engine variant, pieces (the caller strings them together on its communication stream, see runtime.P2PAllReduce):
int mdm_p2p_barrier(const mdm_p2p_comm* comm, int which, void* stream) {
  P2PComm c;
  int rc = fill_comm(c, comm);
  if (rc) return rc;
  MDM_CHECK_ARG(which >= 0 && which < AR_BARRIERS, "p2p_barrier: 0 <= which < %d", AR_BARRIERS);
  launch_pdl(p2p_barrier_kernel, dim3(1), dim3(32), 0, as_stream(stream), c, which);
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  MDM_CHECK_ARG(dst && src && bytes > 0, "memcpy_async: bad arguments");
  MDM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, as_stream(stream)));
  return MDM_OK;
}

int mdm_reduce_slices(float* slice, const float* staging, int64_t stride, int64_t count, int rank, int world, void* stream) {
  MDM_CHECK_ARG(slice && staging && count > 0 && count % 4 == 0 && stride % 4 == 0 && world >= 2 && world <= AR_MAX_RANKS && rank >= 0 && rank < world,
                "reduce_slices: bad arguments");
  const long long n4 = count / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  cudaStream_t st = as_stream(stream);
  const long long str = stride;
  switch (world) {
    case 2: launch_pdl(reduce_slices_kernel<2>, dim3(blocks), dim3(256), 0, st, slice, staging, str, n4, rank); break;
    case 3: launch_pdl(reduce_slices_kernel<3>, dim3(blocks), dim3(256), 0, st, slice, staging, str, n4, rank); break;
    case 4: launch_pdl(reduce_slices_kernel<4>, dim3(blocks), dim3(256), 0, st, slice, staging, str, n4, rank); break;
    case 5: launch_pdl(reduce_slices_kernel<5>, dim3(blocks), dim3(256), 0, st, slice, staging, str, n4, rank); break;
    case 6: launch_pdl(reduce_slices_kernel<6>, dim3(blocks), dim3(256), 0, st, slice, staging, str, n4, rank); break;
    case 7: launch_pdl(reduce_slices_kernel<7>, dim3(blocks), dim3(256), 0, st, slice, staging, str, n4, rank); break;
    default: launch_pdl(reduce_slices_kernel<8>, dim3(blocks), dim3(256), 0, st, slice, staging, str, n4, rank); break;
  }
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

int mdm_ipc_close(void* ptr, int64_t offset) {
  MDM_CHECK_ARG(ptr, "ipc_close: NULL pointer");
  MDM_CUDA(cudaIpcCloseMemHandle((char*)ptr - offset));
  return MDM_OK;
}

// SUM all-reduce of buf[offset .. offset + count) (floats; offset and count multiples of 4) across `world` ranks.
// Every rank must launch the same sequence of calls with the same (offset, count, blocks); calls on one rank must be
// stream-ordered (one communication stream).  The caller divides by world (the fused optimiser folds it in).
int mdm_p2p_allreduce(const mdm_p2p_comm* comm, int64_t offset, int64_t count, int blocks, void* stream) {
  MDM_CHECK_ARG(offset >= 0 && count > 0 && offset % 4 == 0 && count % 4 == 0, "p2p_allreduce: offset / count must be multiples of 4 floats");
  MDM_CHECK_ARG(blocks >= 1 && blocks <= AR_MAX_BLOCKS, "p2p_allreduce: 1 <= blocks <= %d", AR_MAX_BLOCKS);
  P2PComm c;
  int rc = fill_comm(c, comm);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const long long off = offset, cnt = count;
  switch (c.world) {
    case 2: launch_pdl(p2p_allreduce_kernel<2, 8>, dim3(blocks), dim3(AR_THREADS), 0, st, c, off, cnt); break;
    case 3: launch_pdl(p2p_allreduce_kernel<3, 5>, dim3(blocks), dim3(AR_THREADS), 0, st, c, off, cnt); break;
    case 4: launch_pdl(p2p_allreduce_kernel<4, 4>, dim3(blocks), dim3(AR_THREADS), 0, st, c, off, cnt); break;
    case 5: launch_pdl(p2p_allreduce_kernel<5, 3>, dim3(blocks), dim3(AR_THREADS), 0, st, c, off, cnt); break;
    case 6: launch_pdl(p2p_allreduce_kernel<6, 3>, dim3(blocks), dim3(AR_THREADS), 0, st, c, off, cnt); break;
    case 7: launch_pdl(p2p_allreduce_kernel<7, 2>, dim3(blocks), dim3(AR_THREADS), 0, st, c, off, cnt); break;
    default: launch_pdl(p2p_allreduce_kernel<8, 2>, dim3(blocks), dim3(AR_THREADS), 0, st, c, off, cnt); break;
  }
  MDM_LAUNCH_CHECK();
  return MDM_OK;
}

}  // extern "C"
