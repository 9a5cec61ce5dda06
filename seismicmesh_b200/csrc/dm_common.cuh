// Shared device helpers: warp/block scans and reductions, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DM_CUDA_TRY(x)                      \
  do {                                      \
    cudaError_t e__ = (x);                  \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

#define DM_LAUNCH_CHECK()                       \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

// Programmatic dependent launch (sm_90+): the six kernels of an iteration form a dependent chain of
// short launches; each kernel lets its successor be scheduled as soon as all of its own blocks have
// started and then waits for its predecessor's results, which takes the launch latency and the
// block ramp-up of kernel k+1 off the critical path.  DM_PDL=0 builds plain stream-ordered launches.
#ifndef DM_PDL
#define DM_PDL 1
#endif

namespace dm {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void pdl_prologue() {
#if DM_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// `smem`: dynamic shared memory in bytes (the caller has raised the kernel's limit beyond 48 KB)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain_smem(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                            cudaStream_t st, Args&&... args) {
#if DM_PDL
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
#else
  kern<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
  return cudaGetLastError();
#endif
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t st,
                                       Args&&... args) {
  return launch_chain_smem(kern, grid, block, 0, st, static_cast<Args&&>(args)...);
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// arrival counter with release / acquire ordering at device scope: what the arriving thread wrote
// before is visible to whoever sees its increment, and the thread that sees the last arrival may read
// what the others wrote -- a lighter fence than __threadfence() (fence.sc) around a relaxed atomic
__device__ __forceinline__ int arrive_acq_rel(int32_t* counter) {
  int old;
  asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
  return old;
}

__device__ __forceinline__ int warp_inclusive_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// Exclusive scan of one int per thread across the block (blockDim.x multiple of 32, <= 1024).
// `total` receives the block sum in every thread. smem: >= 33 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int& total, int* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int inc = warp_inclusive_scan(v);
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < nw ? smem[lane] : 0;
    const int winc = warp_inclusive_scan(w);
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  const int off = smem[wid];
  total = smem[32];
  __syncthreads();  // smem reusable by the caller
  return off + inc - v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(FULL, v, d);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_down_sync(FULL, v, d));
  return v;
}

// Deterministic block reductions (fixed shuffle tree). Result valid in thread 0. smem >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = lane < nw ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double v, double* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = lane < nw ? smem[lane] : 0.0;
    r = warp_max(r);
  }
  __syncthreads();
  return r;
}

}  // namespace dm
