// Single-pass exclusive scan of int32 (decoupled look-back, one launch).
//
// Scratch: `desc` = one 64-bit descriptor per 2048-element tile, `ticket` = one int.  Both must be
// ZERO on entry (exclusive_scan below zeroes them itself).
// Descriptor word: bits 63..62 = status (0 invalid, 1 tile aggregate, 2 inclusive prefix),
// low 32 bits = value.  Tiles are handed out through the ticket so that a tile's predecessors
// are always already running -> the look-back spin cannot deadlock.
#pragma once
#include "dm_device.cuh"

namespace dm {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

static inline int64_t scan_tiles(int64_t n) { return n <= 0 ? 1 : cdiv(n, SCAN_TILE); }
// descriptors + ticket, rounded up
static inline size_t scan_scratch_bytes(int64_t n) { return align256((size_t)(scan_tiles(n) + 2) * 8); }

__device__ __forceinline__ void scan_load_tile(const int32_t* in, int64_t n, int64_t base, int (&v)[SCAN_ITEMS]) {
  const int64_t i0 = base + (int64_t)threadIdx.x * SCAN_ITEMS;
  if (i0 + SCAN_ITEMS <= n && ((reinterpret_cast<uintptr_t>(in + i0) & 15) == 0)) {
    const int4 a = *reinterpret_cast<const int4*>(in + i0);
    const int4 b = *reinterpret_cast<const int4*>(in + i0 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (i0 + k < n) ? in[i0 + k] : 0;
  }
}

// out[i] = sum_{j<i} in[j] for i in [0,n], i.e. out[n] = total.  in == out allowed.
__global__ void __launch_bounds__(SCAN_THREADS) scan_lookback_kernel(const int32_t* in, int32_t* out, int64_t n,
                                                                     unsigned long long* desc, int* ticket) {
  __shared__ int sm[33];
  __shared__ int s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = (int64_t)tile * SCAN_TILE;
  int v[SCAN_ITEMS];
  scan_load_tile(in, n, base, v);
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
  int total;
  const int ex = block_exclusive_scan(s, total, sm);
  if (threadIdx.x == 0) {
    int prefix = 0;
    if (tile == 0) {
      atomicExch(desc, (2ull << 62) | (unsigned)total);
    } else {
      atomicExch(desc + tile, (1ull << 62) | (unsigned)total);
      int j = tile - 1;
      while (true) {
        const unsigned long long d = *reinterpret_cast<volatile unsigned long long*>(desc + j);
        const unsigned st = (unsigned)(d >> 62);
        if (st == 0) continue;  // predecessor has not published yet
        prefix += (int)(unsigned)(d & 0xffffffffull);
        if (st == 2) break;
        --j;
      }
      atomicExch(desc + tile, (2ull << 62) | (unsigned)(prefix + total));
    }
    s_prefix = prefix;
  }
  __syncthreads();
  int run = s_prefix + ex;
  const int64_t i0 = base + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (i0 + k < n) out[i0 + k] = run;
    run += v[k];
  }
  // the last tile (it contains index n-1, or is the single tile of an empty input) owns out[n]
  if (threadIdx.x == 0 && tile == (int)gridDim.x - 1) out[n] = s_prefix + total;
}

// desc/ticket must already be zero.
static inline int scan_launch(const int32_t* in, int32_t* out, int64_t n, unsigned long long* desc, int* ticket,
                              cudaStream_t st) {
  scan_lookback_kernel<<<(unsigned)scan_tiles(n), SCAN_THREADS, 0, st>>>(in, out, n, desc, ticket);
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

// stand-alone scan: zeroes its scratch first (one memset + one kernel)
static inline int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, void* scratch, size_t scratch_bytes,
                                 cudaStream_t st) {
  if (n < 0) return DM_ERR_ARG;
  const size_t need = scan_scratch_bytes(n);
  if (scratch_bytes < need) return DM_ERR_WORKSPACE;
  cudaError_t e = cudaMemsetAsync(scratch, 0, need, st);
  if (e != cudaSuccess) return (int)e;
  unsigned long long* desc = static_cast<unsigned long long*>(scratch);
  int* ticket = reinterpret_cast<int*>(desc + scan_tiles(n) + 1);
  return scan_launch(in, out, n, desc, ticket, st);
}

}  // namespace dm
