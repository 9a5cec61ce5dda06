// libdistmesh_b200: hand-written sm_100a kernels for the DistMesh force iteration of
// SeismicMesh (generation/mesh_generator.py:460-527) behind the C ABI of
// include/distmesh_b200.h.  Compiled with -fmad=false: fp64 everywhere, no FMA contraction.
//
// Data layout in HBM (all caller-owned): p (N,dim) f64 AoS, t (T,dim+1) i32 AoS; per-iteration
// scratch carved from one workspace (DmPlan).  Every kernel is gather/scan/segmented-sum work
// bound by HBM bandwidth; nothing here is a dense contraction, so tensor cores are not used.
#include <cuda_runtime.h>
#include <float.h>
#include <stdio.h>
#include <string.h>

#include "dm_common.cuh"
#include "dm_sdf.cuh"

using namespace dm;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// =============================================================================================
// exclusive scan (int32): reduce -> scan of block sums -> apply
// =============================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ void scan_load_tile(const int32_t* __restrict__ in, int64_t n, int64_t base,
                                               int (&v)[SCAN_ITEMS]) {
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

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                   int32_t* __restrict__ block_sums) {
  __shared__ int sm[33];
  int v[SCAN_ITEMS];
  scan_load_tile(in, n, (int64_t)blockIdx.x * SCAN_TILE, v);
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
  int total;
  block_exclusive_scan(s, total, sm);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block; exclusive scan of block_sums[0..nb) in place, block_sums[nb] = grand total
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ block_sums, int64_t nb) {
  __shared__ int sm[33];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, total, sm);
    const int carry = carry_s;
    if (i < nb) block_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[nb] = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t* in, int32_t* out, int64_t n,
                                                                  const int32_t* __restrict__ block_sums,
                                                                  int64_t nb) {
  __shared__ int sm[33];
  int v[SCAN_ITEMS];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  scan_load_tile(in, n, base, v);
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
  int total;
  int run = block_exclusive_scan(s, total, sm) + block_sums[blockIdx.x];
  const int64_t i0 = base + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (i0 + k < n) out[i0 + k] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[nb];
}

int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, void* scratch, size_t scratch_bytes,
                   cudaStream_t st) {
  if (n < 0) return DM_ERR_ARG;
  const int64_t nb = n == 0 ? 1 : cdiv(n, SCAN_TILE);
  if (scratch_bytes < (size_t)(nb + 1) * sizeof(int32_t)) return DM_ERR_WORKSPACE;
  int32_t* sums = static_cast<int32_t*>(scratch);
  scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums);
  scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nb);
  scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, sums, nb);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

// =============================================================================================
// point loads
// =============================================================================================
template <int DIM>
__device__ __forceinline__ void load_pt(const double* __restrict__ p, int64_t i, double& x0, double& x1,
                                        double& x2) {
  if (DIM == 2) {
    const double2 v = *reinterpret_cast<const double2*>(p + 2 * i);  // 16-B aligned rows
    x0 = v.x;
    x1 = v.y;
    x2 = 0.0;
  } else {
    const double* q = p + 3 * i;
    x0 = q[0];
    x1 = q[1];
    x2 = q[2];
  }
}
template <int DIM>
__device__ __forceinline__ void store_pt(double* __restrict__ p, int64_t i, double x0, double x1, double x2) {
  if (DIM == 2) {
    *reinterpret_cast<double2*>(p + 2 * i) = make_double2(x0, x1);
  } else {
    double* q = p + 3 * i;
    q[0] = x0;
    q[1] = x1;
    q[2] = x2;
  }
}

template <int DIM>
__device__ __forceinline__ void load_cell(const int32_t* __restrict__ t, int64_t c, int (&v)[4]) {
  if (DIM == 3) {
    const int4 q = *reinterpret_cast<const int4*>(t + 4 * c);  // 16-B aligned rows
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    const int32_t* q = t + 3 * c;
    v[0] = q[0]; v[1] = q[1]; v[2] = q[2]; v[3] = 0;
  }
}

__device__ __forceinline__ void cswap(int& a, int& b) {
  const int lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}
template <int DIM>
__device__ __forceinline__ void sort_cell(int (&v)[4]) {
  if (DIM == 2) {
    cswap(v[0], v[1]);
    cswap(v[1], v[2]);
    cswap(v[0], v[1]);
  } else {
    cswap(v[0], v[1]);
    cswap(v[2], v[3]);
    cswap(v[0], v[2]);
    cswap(v[1], v[3]);
    cswap(v[1], v[2]);
  }
}

// centroid p[t].sum(1)/(dim+1), vertices added in order (mesh_generator.py:737)
template <int DIM>
__device__ __forceinline__ void cell_centroid(const double* __restrict__ p, const int (&v)[4], double& c0,
                                              double& c1, double& c2) {
  double a0, a1, a2, b0, b1, b2;
  load_pt<DIM>(p, v[0], a0, a1, a2);
#pragma unroll
  for (int k = 1; k <= DIM; ++k) {
    load_pt<DIM>(p, v[k], b0, b1, b2);
    a0 = a0 + b0;
    a1 = a1 + b1;
    a2 = a2 + b2;
  }
  c0 = a0 / (double)(DIM + 1);
  c1 = a1 / (double)(DIM + 1);
  c2 = a2 / (double)(DIM + 1);
}

// =============================================================================================
// elementwise kernels: fd, fh, centroids, cull
// =============================================================================================
template <int DIM>
__global__ void sdf_eval_kernel(const double* __restrict__ prog, const double* __restrict__ x, int64_t M,
                                double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double x0, x1, x2;
  load_pt<DIM>(x, i, x0, x1, x2);
  out[i] = sdf_eval(prog, DIM, x0, x1, x2);
}

template <int DIM>
__global__ void size_eval_kernel(const DmSizeFn f, const double* __restrict__ x, int64_t M,
                                 double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double x0, x1, x2;
  load_pt<DIM>(x, i, x0, x1, x2);
  out[i] = size_eval(f, x0, x1, x2);
}

template <int DIM>
__global__ void centroid_kernel(const double* __restrict__ p, const int32_t* __restrict__ t, int64_t T,
                                double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  int v[4];
  load_cell<DIM>(t, c, v);
  double c0, c1, c2;
  cell_centroid<DIM>(p, v, c0, c1, c2);
  store_pt<DIM>(out, c, c0, c1, c2);
}

// mode 0: evaluate fd on the centroid and write keep ; 1: keep given ; 2: all kept
// cnt != NULL: also count raw bars per min vertex (3 | 2 | 1 per sorted cell vertex)
template <int DIM>
__global__ void cull_count_kernel(const double* __restrict__ prog, const double* __restrict__ p,
                                  const int32_t* __restrict__ t, int64_t T, double geps, int mode,
                                  uint8_t* __restrict__ keep, int32_t* __restrict__ cnt) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  int v[4];
  load_cell<DIM>(t, c, v);
  bool k = true;
  if (mode == 0) {
    double c0, c1, c2;
    cell_centroid<DIM>(p, v, c0, c1, c2);
    k = sdf_eval(prog, DIM, c0, c1, c2) < -geps;
    keep[c] = k ? 1 : 0;
  } else if (mode == 1) {
    k = keep[c] != 0;
  }
  if (k && cnt != nullptr) {
    sort_cell<DIM>(v);
    if (DIM == 2) {
      atomicAdd(cnt + v[0], 2);
      atomicAdd(cnt + v[1], 1);
    } else {
      atomicAdd(cnt + v[0], 3);
      atomicAdd(cnt + v[1], 2);
      atomicAdd(cnt + v[2], 1);
    }
  }
}

// bucket_end[] holds bucket starts on entry (exclusive scan of the counts); each cell reserves its
// slots with one atomic per min vertex, so on exit bucket_end[v] is the END of bucket v.
template <int DIM>
__global__ void bar_fill_kernel(const int32_t* __restrict__ t, int64_t T, const uint8_t* __restrict__ keep,
                                int32_t* __restrict__ bucket_end, int32_t* __restrict__ raw) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  if (keep != nullptr && keep[c] == 0) return;
  int v[4];
  load_cell<DIM>(t, c, v);
  sort_cell<DIM>(v);
  if (DIM == 2) {
    int s = atomicAdd(bucket_end + v[0], 2);
    raw[s] = v[1];
    raw[s + 1] = v[2];
    s = atomicAdd(bucket_end + v[1], 1);
    raw[s] = v[2];
  } else {
    int s = atomicAdd(bucket_end + v[0], 3);
    raw[s] = v[1];
    raw[s + 1] = v[2];
    raw[s + 2] = v[3];
    s = atomicAdd(bucket_end + v[1], 2);
    raw[s] = v[2];
    raw[s + 1] = v[3];
    s = atomicAdd(bucket_end + v[2], 1);
    raw[s] = v[3];
  }
}

// =============================================================================================
// per-vertex sort + unique of the raw buckets
// =============================================================================================
constexpr int SU_THREADS = 128;
constexpr int SU_CAP = 8192;  // ints of shared staging per block (32 KB)

// in-place insertion into a sorted unique prefix; returns the unique count
__device__ __forceinline__ int sort_unique_segment(int32_t* seg, int n, int32_t* __restrict__ lcnt) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const int x = seg[i];
    int j = m;
    while (j > 0 && seg[j - 1] > x) --j;
    if (j > 0 && seg[j - 1] == x) continue;
    for (int k = m; k > j; --k) seg[k] = seg[k - 1];
    seg[j] = x;
    ++m;
    atomicAdd(lcnt + x, 1);  // lower-neighbour degree of x
  }
  return m;
}

__global__ void __launch_bounds__(SU_THREADS) sort_unique_kernel(const int32_t* __restrict__ bucket_end,
                                                                 int32_t* raw, int64_t N,
                                                                 int32_t* __restrict__ ucnt,
                                                                 int32_t* __restrict__ lcnt) {
  __shared__ int32_t stage[SU_CAP];
  const int64_t v0 = (int64_t)blockIdx.x * SU_THREADS;
  const int64_t v1 = min(v0 + (int64_t)SU_THREADS, N);
  const int bstart = v0 > 0 ? bucket_end[v0 - 1] : 0;
  const int bend = bucket_end[v1 - 1];
  const int span = bend - bstart;
  const bool staged = span <= SU_CAP;
  if (staged) {
    for (int i = threadIdx.x; i < span; i += SU_THREADS) stage[i] = raw[bstart + i];
    __syncthreads();
  }
  const int64_t v = v0 + threadIdx.x;
  if (v >= N) return;
  const int s = v > 0 ? bucket_end[v - 1] : 0;
  const int n = bucket_end[v] - s;
  int32_t* seg = staged ? (stage + (s - bstart)) : (raw + s);
  const int m = sort_unique_segment(seg, n, lcnt);
  ucnt[v] = m;
  if (staged)
    for (int j = 0; j < m; ++j) raw[s + j] = seg[j];
}

// copy unique heads to the compact CSR and scatter the transposed (lower-neighbour) entries.
// lrow[] holds list starts on entry and list ENDS on exit (same trick as bar_fill_kernel).
__global__ void compact_transpose_kernel(const int32_t* __restrict__ bucket_end, const int32_t* __restrict__ raw,
                                         const int32_t* __restrict__ rowptr, int64_t N,
                                         int32_t* __restrict__ col, int32_t* __restrict__ lrow,
                                         unsigned long long* __restrict__ low) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int s = v > 0 ? bucket_end[v - 1] : 0;
  const int base = rowptr[v];
  const int m = rowptr[v + 1] - base;
  for (int j = 0; j < m; ++j) {
    const int w = raw[s + j];
    col[base + j] = w;
    const int slot = atomicAdd(lrow + w, 1);
    low[slot] = ((unsigned long long)(unsigned)v << 32) | (unsigned)(base + j);
  }
}

// order each vertex's lower-neighbour list by neighbour id (== by bar id)
__global__ void lower_sort_kernel(const int32_t* __restrict__ lrow_end, unsigned long long* low, int64_t N) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int s = v > 0 ? lrow_end[v - 1] : 0;
  const int n = lrow_end[v] - s;
  unsigned long long* seg = low + s;
  for (int i = 1; i < n; ++i) {
    const unsigned long long x = seg[i];
    int j = i;
    while (j > 0 && seg[j - 1] > x) {
      seg[j] = seg[j - 1];
      --j;
    }
    seg[j] = x;
  }
}

__global__ void bars_pairs_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                  int64_t N, int32_t* __restrict__ pairs) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int b = rowptr[v], e = rowptr[v + 1];
  for (int k = b; k < e; ++k) {
    pairs[2 * (int64_t)k] = (int32_t)v;
    pairs[2 * (int64_t)k + 1] = col[k];
  }
}

// =============================================================================================
// bar pass: h at midpoints, partial sums of L^d and h^d
// =============================================================================================
constexpr int BP_THREADS = 256;

template <int DIM>
__device__ __forceinline__ double bar_length(double a0, double a1, double a2, double b0, double b1, double b2,
                                             double& d0, double& d1, double& d2) {
  d0 = a0 - b0;
  d1 = a1 - b1;
  d2 = a2 - b2;
  double s = d0 * d0 + d1 * d1;
  if (DIM == 3) s = s + d2 * d2;
  double L = sqrt(s);
  if (L == 0.0) L = DBL_EPSILON;  // mesh_generator.py:698
  return L;
}

template <int DIM, int MODE>  // MODE 0: evaluate fh and store hbar ; 1: hbar given ; 2: midpoints out
__global__ void __launch_bounds__(BP_THREADS) bar_pass_kernel(const DmSizeFn f, const double* __restrict__ p,
                                                              const int32_t* __restrict__ rowptr,
                                                              const int32_t* __restrict__ col, int64_t N,
                                                              double* __restrict__ hbar,
                                                              double* __restrict__ partials,
                                                              double* __restrict__ mid) {
  __shared__ double sm[32];
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double sL = 0.0, sH = 0.0;
  if (v < N) {
    const int b = rowptr[v], e = rowptr[v + 1];
    if (e > b) {
      double a0, a1, a2;
      load_pt<DIM>(p, v, a0, a1, a2);
      for (int k = b; k < e; ++k) {
        const int w = col[k];
        double b0, b1, b2, d0, d1, d2;
        load_pt<DIM>(p, w, b0, b1, b2);
        // midpoint p[edges].sum(1)/2 (mesh_generator.py:699)
        const double m0 = (a0 + b0) / 2, m1 = (a1 + b1) / 2, m2 = (a2 + b2) / 2;
        if (MODE == 2) {
          store_pt<DIM>(mid, k, m0, m1, m2);
          continue;
        }
        const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
        double h;
        if (MODE == 0) {
          h = size_eval(f, m0, m1, m2);
          hbar[k] = h;
        } else {
          h = hbar[k];
        }
        if (DIM == 2) {
          sL += L * L;
          sH += h * h;
        } else {
          sL += L * L * L;
          sH += h * h * h;
        }
      }
    }
  }
  if (MODE == 2) return;
  const double bl = block_sum(sL, sm);
  const double bh = block_sum(sH, sm);
  if (threadIdx.x == 0) {
    partials[2 * (int64_t)blockIdx.x] = bl;
    partials[2 * (int64_t)blockIdx.x + 1] = bh;
  }
}

// single block: fixed-order reduction of the block partials, scale = (sum L^d / sum h^d)^(1/d)
__global__ void __launch_bounds__(1024) scale_kernel(const double* __restrict__ partials, int64_t nb, int dim,
                                                     double* __restrict__ scalars) {
  __shared__ double sm[32];
  double sL = 0.0, sH = 0.0;
  for (int64_t i = threadIdx.x; i < nb; i += 1024) {
    sL += partials[2 * i];
    sH += partials[2 * i + 1];
  }
  const double tl = block_sum(sL, sm);
  const double th = block_sum(sH, sm);
  if (threadIdx.x == 0) {
    scalars[0] = tl;
    scalars[1] = th;
    const double r = tl / th;
    scalars[2] = dim == 2 ? sqrt(r) : pow(r, 1.0 / 3.0);  // ** (1.0 / dim), mesh_generator.py:700
  }
}

// =============================================================================================
// vertex update: deterministic gather of bar forces + update + projection + max|F|
// =============================================================================================
struct Levels {
  const double* prog[DM_MAX_LEVELS];
  int n;
};

template <int DIM>
__global__ void __launch_bounds__(BP_THREADS) vertex_update_kernel(
    const double* __restrict__ p, double* __restrict__ p_out, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ col, const int32_t* __restrict__ lrow_end,
    const unsigned long long* __restrict__ low, const double* __restrict__ hbar,
    const double* __restrict__ scalars, int64_t N, Levels lv, double L0mult, double delta_t, double deps,
    double h0, int64_t nfix, const uint8_t* __restrict__ fixed, double* __restrict__ Ftot,
    double* __restrict__ partials) {
  __shared__ double sm[32];
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double f2 = 0.0;
  if (v < N) {
    const double scale = scalars[2];
    double a0, a1, a2;
    load_pt<DIM>(p, v, a0, a1, a2);
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    // lower neighbours u < v, ascending: the reference subtracts Fvec of bar (u,v) from row v;
    // -(F/L*(p[u]-p[v])) == (F/L)*(p[v]-p[u]) exactly.
    {
      const int s = v > 0 ? lrow_end[v - 1] : 0, e = lrow_end[v];
      for (int k = s; k < e; ++k) {
        const unsigned long long key = low[k];
        const int u = (int)(key >> 32);
        const int bar = (int)(key & 0xffffffffu);
        double b0, b1, b2, d0, d1, d2;
        load_pt<DIM>(p, u, b0, b1, b2);
        const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
        double F = hbar[bar] * L0mult * scale - L;
        if (F < 0) F = 0;
        const double q = F / L;
        F0 = F0 + q * d0;
        F1 = F1 + q * d1;
        if (DIM == 3) F2 = F2 + q * d2;
      }
    }
    // upper neighbours w > v, ascending
    {
      const int s = rowptr[v], e = rowptr[v + 1];
      for (int k = s; k < e; ++k) {
        const int w = col[k];
        double b0, b1, b2, d0, d1, d2;
        load_pt<DIM>(p, w, b0, b1, b2);
        const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
        double F = hbar[k] * L0mult * scale - L;
        if (F < 0) F = 0;
        const double q = F / L;
        F0 = F0 + q * d0;
        F1 = F1 + q * d1;
        if (DIM == 3) F2 = F2 + q * d2;
      }
    }
    if (v < nfix || (fixed != nullptr && fixed[v])) {  // Ftot[ifix] = 0 (mesh_generator.py:499)
      F0 = 0.0;
      F1 = 0.0;
      F2 = 0.0;
    }
    if (Ftot != nullptr) store_pt<DIM>(Ftot, v, F0, F1, F2);
    f2 = F0 * F0 + F1 * F1;
    if (DIM == 3) f2 = f2 + F2 * F2;
    // p += delta_t * Ftot (mesh_generator.py:502)
    double x0 = a0 + delta_t * F0, x1 = a1 + delta_t * F1, x2 = a2 + delta_t * F2;
    for (int l = 0; l < lv.n; ++l) sdf_project(lv.prog[l], DIM, deps, h0, l, x0, x1, x2);
    store_pt<DIM>(p_out, v, x0, x1, x2);
  }
  const double bm = block_max(f2, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = bm;
}

__global__ void __launch_bounds__(1024) maxdp_kernel(const double* __restrict__ partials, int64_t nb,
                                                     double delta_t, double* __restrict__ scalars) {
  __shared__ double sm[32];
  double m = 0.0;
  for (int64_t i = threadIdx.x; i < nb; i += 1024) m = fmax(m, partials[i]);
  const double t = block_max(m, sm);
  if (threadIdx.x == 0) {
    scalars[3] = t;
    scalars[4] = delta_t * sqrt(t);  // mesh_generator.py:514
  }
}

template <int DIM>
__global__ void project_kernel(const double* __restrict__ prog, double* __restrict__ p, int64_t N, double deps,
                               double h0, int level_idx) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  double x0, x1, x2;
  load_pt<DIM>(p, v, x0, x1, x2);
  if (sdf_project(prog, DIM, deps, h0, level_idx, x0, x1, x2)) store_pt<DIM>(p, v, x0, x1, x2);
}

// _improve_level_set_newton (mesh_generator.py:741-759): alpha = 1,1,1/2,1/6,1/24
template <int DIM>
__global__ void level_set_newton_kernel(const double* __restrict__ prog, double* __restrict__ p,
                                        const int32_t* __restrict__ bid, int64_t nb, double deps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const int64_t v = bid[i];
  double x0, x1, x2;
  load_pt<DIM>(p, v, x0, x1, x2);
  double alpha = 1.0;
  for (int it = 0; it < 5; ++it) {
    const double d = sdf_eval(prog, DIM, x0, x1, x2);
    const double g0 = (sdf_eval(prog, DIM, x0 + deps, x1, x2) - d) / deps;
    const double g1 = (sdf_eval(prog, DIM, x0, x1 + deps, x2) - d) / deps;
    double g2 = 0.0;
    double s = 0.0 + g0 * g0;
    s = s + g1 * g1;
    if (DIM == 3) {
      g2 = (sdf_eval(prog, DIM, x0, x1, x2 + deps) - d) / deps;
      s = s + g2 * g2;
    }
    if (s < deps) s = deps;
    x0 = x0 - alpha * (d * g0 / s);
    x1 = x1 - alpha * (d * g1 / s);
    if (DIM == 3) x2 = x2 - alpha * (d * g2 / s);
    alpha = alpha / (double)(it + 1);
  }
  store_pt<DIM>(p, v, x0, x1, x2);
}

// =============================================================================================
// cell compaction, sliver kernels, halo selection
// =============================================================================================
__global__ void flags_to_int_kernel(const uint8_t* __restrict__ keep, int64_t T, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) out[i] = keep[i] ? 1 : 0;
}
template <int C>
__global__ void compact_cells_kernel(const int32_t* __restrict__ t, const uint8_t* __restrict__ keep,
                                     const int32_t* __restrict__ pos, int64_t T, int32_t* __restrict__ t_out,
                                     int32_t* __restrict__ T_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *T_out = pos[T];
  if (i >= T || !keep[i]) return;
  const int64_t o = pos[i];
#pragma unroll
  for (int k = 0; k < C; ++k) t_out[C * o + k] = t[C * i + k];
}

__global__ void dihedral_kernel(const double* __restrict__ p, const int32_t* __restrict__ t, int64_t T,
                                double min_dh, double max_dh, double* __restrict__ angles,
                                uint8_t* __restrict__ flags) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  int v[4];
  load_cell<3>(t, c, v);
  double P[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) load_pt<3>(p, v[k], P[k][0], P[k][1], P[k][2]);
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double a = dihedral_angle(P, i);
    if (angles != nullptr) angles[6 * c + i] = a;
    bad = bad || (a < min_dh) || (a > max_dh);
  }
  if (flags != nullptr) flags[c] = bad ? 1 : 0;
}

__global__ void circumsphere_grad_kernel(const double* __restrict__ p, const int32_t* __restrict__ t,
                                         const int32_t* __restrict__ ele, int64_t S_, double* __restrict__ grad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S_) return;
  const int64_t c = ele != nullptr ? ele[i] : i;
  int v[4];
  load_cell<3>(t, c, v);
  double P[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) load_pt<3>(p, v[k], P[k][0], P[k][1], P[k][2]);
  double g[3];
  circumsphere_grad(P[0], P[1], P[2], P[3], g);
  grad[3 * i] = g[0];
  grad[3 * i + 1] = g[1];
  grad[3 * i + 2] = g[2];
}

// fancy-index `p[move] += ...` keeps the LAST sliver for a repeated vertex (mesh_generator.py:274)
__global__ void sliver_winner_kernel(const int32_t* __restrict__ t, const int32_t* __restrict__ ele, int64_t S_,
                                     int32_t* __restrict__ winner) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S_) return;
  atomicMax(winner + t[4 * (int64_t)ele[i]], (int)i);
}
// phase 1: displacement of every winning sliver from the PRE-update positions
__global__ void sliver_delta_kernel(const double* __restrict__ p, const int32_t* __restrict__ t,
                                    const int32_t* __restrict__ ele, int64_t S_, double step_h0,
                                    const int32_t* __restrict__ winner, double* __restrict__ delta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S_) return;
  const int64_t c = ele[i];
  int v[4];
  load_cell<3>(t, c, v);
  if (winner[v[0]] != (int)i) return;
  double P[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) load_pt<3>(p, v[k], P[k][0], P[k][1], P[k][2]);
  double g[3];
  circumsphere_grad(P[0], P[1], P[2], P[3], g);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (isinf(g[k])) g[k] = 1.0;  // mesh_generator.py:254
  // np.sum(np.abs(g)**2, axis=-1) ** 0.5  (:257)
  const double nrm = sqrt(fabs(g[0]) * fabs(g[0]) + fabs(g[1]) * fabs(g[1]) + fabs(g[2]) * fabs(g[2]));
#pragma unroll
  for (int k = 0; k < 3; ++k) delta[3 * i + k] = step_h0 * (g[k] / nrm);
}
// phase 2: apply
__global__ void sliver_apply_kernel(double* __restrict__ p, const int32_t* __restrict__ t,
                                    const int32_t* __restrict__ ele, int64_t S_,
                                    const int32_t* __restrict__ winner, const double* __restrict__ delta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S_) return;
  const int64_t v0 = t[4 * (int64_t)ele[i]];
  if (winner[v0] != (int)i) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) p[3 * v0 + k] = p[3 * v0 + k] + delta[3 * i + k];
}

// circumball of each cell vs the padded slab boxes of the rank below / above
// (migration/cpp/cpputils.cpp:85-200, 247-383).  boxes: [min(dim), max(dim)] x 2.
struct HaloBoxes {
  double lo[2][3];
  double hi[2][3];
  int has[2];
};
template <int DIM>
__global__ void halo_select_kernel(const double* __restrict__ p, const int32_t* __restrict__ t, int64_t T,
                                   HaloBoxes hb, uint8_t* __restrict__ flags) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  int v[4];
  load_cell<DIM>(t, c, v);
  double P[4][3];
#pragma unroll
  for (int k = 0; k <= DIM; ++k) load_pt<DIM>(p, v[k], P[k][0], P[k][1], P[k][2]);
  double mn[3], mx[3];
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    mn[j] = P[0][j];
    mx[j] = P[0][j];
#pragma unroll
    for (int k = 1; k <= DIM; ++k) {
      mn[j] = fmin(mn[j], P[k][j]);
      mx[j] = fmax(mx[j], P[k][j]);
    }
  }
  // circumcentre relative to vertex 0
  double cc[3] = {0, 0, 0}, r2;
  bool degenerate = false;
  if (DIM == 2) {
    const double ax = P[1][0] - P[0][0], ay = P[1][1] - P[0][1];
    const double bx = P[2][0] - P[0][0], by = P[2][1] - P[0][1];
    const double det = 2.0 * (ax * by - ay * bx);
    degenerate = det == 0.0;
    const double a2 = ax * ax + ay * ay, b2 = bx * bx + by * by;
    cc[0] = (by * a2 - ay * b2) / det;
    cc[1] = (ax * b2 - bx * a2) / det;
  } else {
    const double a[3] = {P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]};
    const double b[3] = {P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2]};
    const double cv[3] = {P[3][0] - P[0][0], P[3][1] - P[0][1], P[3][2] - P[0][2]};
    const double a2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    const double b2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    const double c2 = cv[0] * cv[0] + cv[1] * cv[1] + cv[2] * cv[2];
    const double bxc[3] = {b[1] * cv[2] - b[2] * cv[1], b[2] * cv[0] - b[0] * cv[2], b[0] * cv[1] - b[1] * cv[0]};
    const double cxa[3] = {cv[1] * a[2] - cv[2] * a[1], cv[2] * a[0] - cv[0] * a[2], cv[0] * a[1] - cv[1] * a[0]};
    const double axb[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double det = 2.0 * (a[0] * bxc[0] + a[1] * bxc[1] + a[2] * bxc[2]);
    degenerate = det == 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) cc[j] = (a2 * bxc[j] + b2 * cxa[j] + c2 * axb[j]) / det;
  }
  if (degenerate) return;  // the reference skips collinear / coplanar cells
  r2 = cc[0] * cc[0] + cc[1] * cc[1] + cc[2] * cc[2];
#pragma unroll
  for (int j = 0; j < DIM; ++j) cc[j] += P[0][j];
  unsigned f = 0;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (!hb.has[s]) continue;
    bool overlap = true;
    double d2 = 0.0;
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      overlap = overlap && !(mx[j] < hb.lo[s][j] || mn[j] > hb.hi[s][j]);
      const double d = cc[j] < hb.lo[s][j] ? hb.lo[s][j] - cc[j] : (cc[j] > hb.hi[s][j] ? cc[j] - hb.hi[s][j] : 0.0);
      d2 += d * d;
    }
    if (overlap && d2 <= r2) f |= (1u << s);
  }
  if (f) {
#pragma unroll
    for (int k = 0; k <= DIM; ++k) {
      // byte-wide OR through a 32-bit atomic on the containing word
      uint8_t* addr = flags + v[k];
      unsigned* word = reinterpret_cast<unsigned*>(reinterpret_cast<uintptr_t>(addr) & ~uintptr_t(3));
      const unsigned shift = (unsigned)(reinterpret_cast<uintptr_t>(addr) & 3) * 8;
      atomicOr(word, f << shift);
    }
  }
}

inline unsigned nblk(int64_t n, int threads) { return (unsigned)(n <= 0 ? 1 : cdiv(n, threads)); }

// Optional per-kernel CUDA-event timing, active only inside dm_force_iteration_profiled on the
// calling thread (bench.py's roofline table).  No effect on the normal path.
constexpr int PROF_MAX = 32;
struct Prof {
  cudaEvent_t ev[PROF_MAX + 1];
  const char* name[PROF_MAX];
  int n;
};
thread_local Prof* tl_prof = nullptr;
inline void mark(const char* name, cudaStream_t st) {
  Prof* pr = tl_prof;
  if (pr && pr->n < PROF_MAX) {
    cudaEventRecord(pr->ev[pr->n + 1], st);
    pr->name[pr->n++] = name;
  }
}

int check_prog_host_side(const double* prog) { return prog == nullptr ? DM_ERR_ARG : DM_OK; }

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* dm_version(void) { return "distmesh_b200 0.1 (sm_100a)"; }

size_t dm_scan_scratch_bytes(int64_t n) {
  const int64_t nb = n <= 0 ? 1 : cdiv(n, SCAN_TILE);
  return align256((size_t)(nb + 1) * sizeof(int32_t));
}

int dm_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* scratch, size_t scratch_bytes,
                          void* stream) {
  if (!in || !out || !scratch) return DM_ERR_ARG;
  return exclusive_scan(in, out, n, scratch, scratch_bytes, S(stream));
}

int dm_sdf_eval(const double* prog, const double* x, int64_t M, int dim, double* out, void* stream) {
  if (check_prog_host_side(prog) || M < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if (M == 0) return DM_OK;  // empty input: pointers may be NULL
  if (!x || !out) return DM_ERR_ARG;
  if (dim == 2)
    sdf_eval_kernel<2><<<nblk(M, 256), 256, 0, S(stream)>>>(prog, x, M, out);
  else
    sdf_eval_kernel<3><<<nblk(M, 256), 256, 0, S(stream)>>>(prog, x, M, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_size_eval(const DmSizeFn* f, const double* x, int64_t M, double* out, void* stream) {
  if (!f || M < 0 || (f->dim != 2 && f->dim != 3)) return DM_ERR_ARG;
  if (f->kind == DM_SIZE_EXTERNAL) return DM_ERR_ARG;
  if (f->kind == DM_SIZE_GRID)
    for (int k = 0; k < f->dim; ++k)
      if (f->n[k] < 2 || !f->axis[k]) return DM_ERR_ARG;
  if (M == 0) return DM_OK;  // empty input: pointers may be NULL
  if (!x || !out) return DM_ERR_ARG;
  if (f->dim == 2)
    size_eval_kernel<2><<<nblk(M, 256), 256, 0, S(stream)>>>(*f, x, M, out);
  else
    size_eval_kernel<3><<<nblk(M, 256), 256, 0, S(stream)>>>(*f, x, M, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_centroids(const double* p, const int32_t* t, int64_t T, int dim, double* out, void* stream) {
  if (!p || !t || !out || T < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  if (dim == 2)
    centroid_kernel<2><<<nblk(T, 256), 256, 0, S(stream)>>>(p, t, T, out);
  else
    centroid_kernel<3><<<nblk(T, 256), 256, 0, S(stream)>>>(p, t, T, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_cull_cells(const double* prog, const double* p, const int32_t* t, int64_t T, int dim, double geps,
                  uint8_t* keep, void* stream) {
  if (!prog || !p || !t || !keep || T < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  if (dim == 2)
    cull_count_kernel<2><<<nblk(T, 256), 256, 0, S(stream)>>>(prog, p, t, T, geps, 0, keep, nullptr);
  else
    cull_count_kernel<3><<<nblk(T, 256), 256, 0, S(stream)>>>(prog, p, t, T, geps, 0, keep, nullptr);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

size_t dm_compact_scratch_bytes(int64_t T) {
  return align256((size_t)(T + 1) * sizeof(int32_t)) + dm_scan_scratch_bytes(T);
}

int dm_compact_cells(const int32_t* t, const uint8_t* keep, int64_t T, int dim, int32_t* t_out,
                     int32_t* T_out_dev, void* scratch, size_t scratch_bytes, void* stream) {
  if (!t || !keep || !t_out || !T_out_dev || !scratch || T < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if (scratch_bytes < dm_compact_scratch_bytes(T)) return DM_ERR_WORKSPACE;
  int32_t* pos = static_cast<int32_t*>(scratch);
  char* tmp = static_cast<char*>(scratch) + align256((size_t)(T + 1) * sizeof(int32_t));
  cudaStream_t st = S(stream);
  flags_to_int_kernel<<<nblk(T, 256), 256, 0, st>>>(keep, T, pos);
  int rc = exclusive_scan(pos, pos, T, tmp, dm_scan_scratch_bytes(T), st);
  if (rc) return rc;
  if (dim == 2)
    compact_cells_kernel<3><<<nblk(T, 256), 256, 0, st>>>(t, keep, pos, T, t_out, T_out_dev);
  else
    compact_cells_kernel<4><<<nblk(T, 256), 256, 0, st>>>(t, keep, pos, T, t_out, T_out_dev);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_dihedral(const double* p, const int32_t* t, int64_t T, double min_dh, double max_dh, double* angles,
                uint8_t* flags, void* stream) {
  if (!p || !t || T < 0) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  dihedral_kernel<<<nblk(T, 128), 128, 0, S(stream)>>>(p, t, T, min_dh, max_dh, angles, flags);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_circumsphere_grad(const double* p, const int32_t* t, const int32_t* ele, int64_t S_, double* grad,
                         void* stream) {
  if (!p || !t || !grad || S_ < 0) return DM_ERR_ARG;
  if (S_ == 0) return DM_OK;
  circumsphere_grad_kernel<<<nblk(S_, 128), 128, 0, S(stream)>>>(p, t, ele, S_, grad);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_sliver_perturb(double* p, int64_t N, const int32_t* t, const int32_t* ele, int64_t S_, double step_h0,
                      int32_t* winner, double* delta, void* stream) {
  if (!p || !t || !ele || !winner || !delta || S_ < 0 || N < 0) return DM_ERR_ARG;
  if (S_ == 0) return DM_OK;
  cudaStream_t st = S(stream);
  DM_CUDA_TRY(cudaMemsetAsync(winner, 0xff, (size_t)N * sizeof(int32_t), st));  // -1
  sliver_winner_kernel<<<nblk(S_, 128), 128, 0, st>>>(t, ele, S_, winner);
  sliver_delta_kernel<<<nblk(S_, 128), 128, 0, st>>>(p, t, ele, S_, step_h0, winner, delta);
  sliver_apply_kernel<<<nblk(S_, 128), 128, 0, st>>>(p, t, ele, S_, winner, delta);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_level_set_newton(const double* prog, double* p, const int32_t* bid, int64_t nb, int dim, double deps,
                        void* stream) {
  if (!prog || !p || !bid || nb < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if (nb == 0) return DM_OK;
  if (dim == 2)
    level_set_newton_kernel<2><<<nblk(nb, 128), 128, 0, S(stream)>>>(prog, p, bid, nb, deps);
  else
    level_set_newton_kernel<3><<<nblk(nb, 128), 128, 0, S(stream)>>>(prog, p, bid, nb, deps);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_project_points(const double* prog, double* p, int64_t N, int dim, double deps, double h0, int level_idx,
                      void* stream) {
  if (!prog || !p || N < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if (N == 0) return DM_OK;
  if (dim == 2)
    project_kernel<2><<<nblk(N, 256), 256, 0, S(stream)>>>(prog, p, N, deps, h0, level_idx);
  else
    project_kernel<3><<<nblk(N, 256), 256, 0, S(stream)>>>(prog, p, N, deps, h0, level_idx);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
static size_t plan_layout(DmPlan* pl, int64_t N, int64_t T, int dim, char* base) {
  const int nb = dim == 2 ? 3 : 6;
  const int64_t K = (int64_t)nb * T;
  const int64_t nblocks = cdiv(N > 0 ? N : 1, BP_THREADS);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* ptr = base ? base + off : nullptr;
    off += align256(bytes);
    return ptr;
  };
  char* keep = take((size_t)(T > 0 ? T : 1));
  char* bucket = take((size_t)(N + 1) * 4);
  char* raw = take((size_t)(K > 0 ? K : 1) * 4);
  char* rowptr = take((size_t)(N + 1) * 4);
  char* col = take((size_t)(K > 0 ? K : 1) * 4);
  char* lrow = take((size_t)(N + 1) * 4);
  char* low = take((size_t)(K > 0 ? K : 1) * 8);
  char* hbar = take((size_t)(K > 0 ? K : 1) * 8);
  char* partials = take((size_t)(2 * nblocks) * 8);
  char* scalars = take(8 * 8);
  char* counters = take(8 * 4);
  const size_t scan_bytes = dm_scan_scratch_bytes(N + 1);
  char* scan_tmp = take(scan_bytes);
  if (pl) {
    pl->N = N;
    pl->T = T;
    pl->dim = dim;
    pl->nb = nb;
    pl->K = K;
    pl->keep = reinterpret_cast<uint8_t*>(keep);
    pl->bucket_end = reinterpret_cast<int32_t*>(bucket);
    pl->raw = reinterpret_cast<int32_t*>(raw);
    pl->rowptr = reinterpret_cast<int32_t*>(rowptr);
    pl->col = reinterpret_cast<int32_t*>(col);
    pl->lrowptr = reinterpret_cast<int32_t*>(lrow);
    pl->low = reinterpret_cast<uint64_t*>(low);
    pl->hbar = reinterpret_cast<double*>(hbar);
    pl->partials = reinterpret_cast<double*>(partials);
    pl->scalars = reinterpret_cast<double*>(scalars);
    pl->counters = reinterpret_cast<int32_t*>(counters);
    pl->scan_tmp = scan_tmp;
    pl->scan_tmp_bytes = scan_bytes;
  }
  return off;
}

size_t dm_plan_bytes(int64_t N, int64_t T, int dim) {
  if (N < 0 || T < 0 || (dim != 2 && dim != 3)) return 0;
  return plan_layout(nullptr, N, T, dim, nullptr);
}

int dm_plan_init(DmPlan* plan, int64_t N, int64_t T, int dim, void* ws, size_t ws_bytes) {
  if (!plan || !ws || N <= 0 || T < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if ((int64_t)(dim == 2 ? 3 : 6) * T >= (int64_t)INT32_MAX || N >= (int64_t)INT32_MAX) return DM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return DM_ERR_ARG;
  if (ws_bytes < plan_layout(nullptr, N, T, dim, nullptr)) return DM_ERR_WORKSPACE;
  plan_layout(plan, N, T, dim, static_cast<char*>(ws));
  return DM_OK;
}

int dm_stage_cull_count(const DmPlan* pl, const double* prog, const double* p, const int32_t* t, double geps,
                        int use_keep, void* stream) {
  if (!pl || (!t && pl->T > 0) || (use_keep && prog && !p)) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  DM_CUDA_TRY(cudaMemsetAsync(pl->bucket_end, 0, (size_t)(pl->N + 1) * 4, st));
  mark("memset_counts", st);
  if (pl->T == 0) return DM_OK;
  const int mode = !use_keep ? 2 : (prog ? 0 : 1);
  if (pl->dim == 2)
    cull_count_kernel<2><<<nblk(pl->T, 256), 256, 0, st>>>(prog, p, t, pl->T, geps, mode, pl->keep, pl->bucket_end);
  else
    cull_count_kernel<3><<<nblk(pl->T, 256), 256, 0, st>>>(prog, p, t, pl->T, geps, mode, pl->keep, pl->bucket_end);
  mark("cull_count", st);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_stage_build_bars(const DmPlan* pl, const int32_t* t, int use_keep, void* stream) {
  if (!pl || (!t && pl->T > 0)) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  const int64_t N = pl->N, T = pl->T;
  int rc = exclusive_scan(pl->bucket_end, pl->bucket_end, N, pl->scan_tmp, pl->scan_tmp_bytes, st);
  if (rc) return rc;
  mark("scan_bucket(3 kernels)", st);
  const uint8_t* keep = use_keep ? pl->keep : nullptr;
  if (T > 0) {
    if (pl->dim == 2)
      bar_fill_kernel<2><<<nblk(T, 256), 256, 0, st>>>(t, T, keep, pl->bucket_end, pl->raw);
    else
      bar_fill_kernel<3><<<nblk(T, 256), 256, 0, st>>>(t, T, keep, pl->bucket_end, pl->raw);
  }
  mark("bar_fill", st);
  DM_CUDA_TRY(cudaMemsetAsync(pl->lrowptr, 0, (size_t)(N + 1) * 4, st));
  mark("memset_lower", st);
  sort_unique_kernel<<<nblk(N, SU_THREADS), SU_THREADS, 0, st>>>(pl->bucket_end, pl->raw, N, pl->rowptr, pl->lrowptr);
  mark("sort_unique", st);
  rc = exclusive_scan(pl->rowptr, pl->rowptr, N, pl->scan_tmp, pl->scan_tmp_bytes, st);
  if (rc) return rc;
  rc = exclusive_scan(pl->lrowptr, pl->lrowptr, N, pl->scan_tmp, pl->scan_tmp_bytes, st);
  if (rc) return rc;
  DM_CUDA_TRY(cudaMemcpyAsync(pl->counters, pl->rowptr + N, 4, cudaMemcpyDeviceToDevice, st));
  mark("scan_rowptrs(6 kernels)", st);
  compact_transpose_kernel<<<nblk(N, 256), 256, 0, st>>>(pl->bucket_end, pl->raw, pl->rowptr, N, pl->col,
                                                         pl->lrowptr,
                                                         reinterpret_cast<unsigned long long*>(pl->low));
  mark("compact_transpose", st);
  lower_sort_kernel<<<nblk(N, 256), 256, 0, st>>>(pl->lrowptr, reinterpret_cast<unsigned long long*>(pl->low), N);
  mark("lower_sort", st);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_bars_pairs(const DmPlan* pl, int32_t* pairs, void* stream) {
  if (!pl || !pairs) return DM_ERR_ARG;
  bars_pairs_kernel<<<nblk(pl->N, 256), 256, 0, S(stream)>>>(pl->rowptr, pl->col, pl->N, pairs);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_bar_midpoints(const DmPlan* pl, const double* p, double* mid, void* stream) {
  if (!pl || !p || !mid) return DM_ERR_ARG;
  DmSizeFn f;
  memset(&f, 0, sizeof(f));
  const unsigned nb_ = nblk(pl->N, BP_THREADS);
  if (pl->dim == 2)
    bar_pass_kernel<2, 2><<<nb_, BP_THREADS, 0, S(stream)>>>(f, p, pl->rowptr, pl->col, pl->N, nullptr, nullptr, mid);
  else
    bar_pass_kernel<3, 2><<<nb_, BP_THREADS, 0, S(stream)>>>(f, p, pl->rowptr, pl->col, pl->N, nullptr, nullptr, mid);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_stage_bar_pass(const DmPlan* pl, const double* p, const DmSizeFn* f, void* stream) {
  if (!pl || !p || !f) return DM_ERR_ARG;
  if (f->kind != DM_SIZE_EXTERNAL && f->kind != DM_SIZE_CONST) {
    if (f->kind != DM_SIZE_GRID || f->dim != pl->dim || !f->grid) return DM_ERR_ARG;
    for (int k = 0; k < f->dim; ++k)
      if (f->n[k] < 2 || !f->axis[k]) return DM_ERR_ARG;
  }
  cudaStream_t st = S(stream);
  const unsigned nb_ = nblk(pl->N, BP_THREADS);
  const bool ext = f->kind == DM_SIZE_EXTERNAL;
  if (pl->dim == 2) {
    if (ext)
      bar_pass_kernel<2, 1><<<nb_, BP_THREADS, 0, st>>>(*f, p, pl->rowptr, pl->col, pl->N, pl->hbar, pl->partials, nullptr);
    else
      bar_pass_kernel<2, 0><<<nb_, BP_THREADS, 0, st>>>(*f, p, pl->rowptr, pl->col, pl->N, pl->hbar, pl->partials, nullptr);
  } else {
    if (ext)
      bar_pass_kernel<3, 1><<<nb_, BP_THREADS, 0, st>>>(*f, p, pl->rowptr, pl->col, pl->N, pl->hbar, pl->partials, nullptr);
    else
      bar_pass_kernel<3, 0><<<nb_, BP_THREADS, 0, st>>>(*f, p, pl->rowptr, pl->col, pl->N, pl->hbar, pl->partials, nullptr);
  }
  mark("bar_pass", st);
  scale_kernel<<<1, 1024, 0, st>>>(pl->partials, (int64_t)nb_, pl->dim, pl->scalars);
  mark("scale", st);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_stage_vertex_update(const DmPlan* pl, const double* p, double* p_out, const double* const* progs,
                           int nlevels, double L0mult, double delta_t, double deps, double h0, int64_t nfix,
                           const uint8_t* fixed, double* Ftot, void* stream) {
  if (!pl || !p || !p_out || p == p_out || nlevels < 0 || nlevels > DM_MAX_LEVELS) return DM_ERR_ARG;
  if (nlevels > 0 && !progs) return DM_ERR_ARG;
  Levels lv;
  memset(&lv, 0, sizeof(lv));
  lv.n = nlevels;
  for (int l = 0; l < nlevels; ++l) {
    if (!progs[l]) return DM_ERR_ARG;
    lv.prog[l] = progs[l];
  }
  cudaStream_t st = S(stream);
  const unsigned nb_ = nblk(pl->N, BP_THREADS);
  const unsigned long long* low = reinterpret_cast<const unsigned long long*>(pl->low);
  if (pl->dim == 2)
    vertex_update_kernel<2><<<nb_, BP_THREADS, 0, st>>>(p, p_out, pl->rowptr, pl->col, pl->lrowptr, low, pl->hbar,
                                                        pl->scalars, pl->N, lv, L0mult, delta_t, deps, h0, nfix,
                                                        fixed, Ftot, pl->partials);
  else
    vertex_update_kernel<3><<<nb_, BP_THREADS, 0, st>>>(p, p_out, pl->rowptr, pl->col, pl->lrowptr, low, pl->hbar,
                                                        pl->scalars, pl->N, lv, L0mult, delta_t, deps, h0, nfix,
                                                        fixed, Ftot, pl->partials);
  mark("vertex_update", st);
  maxdp_kernel<<<1, 1024, 0, st>>>(pl->partials, (int64_t)nb_, delta_t, pl->scalars);
  mark("maxdp", st);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_force_iteration_profiled(const DmPlan* pl, const double* const* progs, int nlevels, const DmSizeFn* f,
                                const double* p, const int32_t* t, double* p_out, double geps, double L0mult,
                                double delta_t, double deps, double h0, int64_t nfix, const uint8_t* fixed,
                                void* stream, float* ms_host, char* names_host, int names_stride, int cap,
                                int* n_host) {
  if (!ms_host || !names_host || !n_host || cap <= 0 || names_stride <= 1) return DM_ERR_ARG;
  Prof pr;
  pr.n = 0;
  for (int i = 0; i <= PROF_MAX; ++i) DM_CUDA_TRY(cudaEventCreate(&pr.ev[i]));
  cudaStream_t st = S(stream);
  cudaEventRecord(pr.ev[0], st);
  tl_prof = &pr;
  const int rc = dm_force_iteration(pl, progs, nlevels, f, p, t, p_out, geps, L0mult, delta_t, deps, h0, nfix,
                                    fixed, nullptr, stream);
  tl_prof = nullptr;
  cudaError_t e = cudaStreamSynchronize(st);
  int n = 0;
  if (rc == 0 && e == cudaSuccess) {
    for (; n < pr.n && n < cap; ++n) {
      cudaEventElapsedTime(&ms_host[n], pr.ev[n], pr.ev[n + 1]);
      strncpy(names_host + (size_t)n * names_stride, pr.name[n], names_stride - 1);
      names_host[(size_t)n * names_stride + names_stride - 1] = 0;
    }
  }
  *n_host = n;
  for (int i = 0; i <= PROF_MAX; ++i) cudaEventDestroy(pr.ev[i]);
  if (rc) return rc;
  return (int)e;
}

int dm_force_iteration(const DmPlan* pl, const double* const* progs, int nlevels, const DmSizeFn* f,
                       const double* p, const int32_t* t, double* p_out, double geps, double L0mult,
                       double delta_t, double deps, double h0, int64_t nfix, const uint8_t* fixed, double* Ftot,
                       void* stream) {
  if (!pl || !progs || nlevels < 1 || !f || f->kind == DM_SIZE_EXTERNAL) return DM_ERR_ARG;
  int rc = dm_stage_cull_count(pl, progs[0], p, t, geps, 1, stream);
  if (rc) return rc;
  rc = dm_stage_build_bars(pl, t, 1, stream);
  if (rc) return rc;
  rc = dm_stage_bar_pass(pl, p, f, stream);
  if (rc) return rc;
  return dm_stage_vertex_update(pl, p, p_out, progs, nlevels, L0mult, delta_t, deps, h0, nfix, fixed, Ftot, stream);
}

int dm_halo_select(const double* p, const int32_t* t, int64_t T, int64_t N, int dim, const double* boxes,
                   int has_below, int has_above, uint8_t* flags, void* stream) {
  if (!p || !t || !boxes || !flags || T < 0 || N < 0 || (dim != 2 && dim != 3)) return DM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(flags) & 3) != 0) return DM_ERR_ARG;
  HaloBoxes hb;
  memset(&hb, 0, sizeof(hb));
  hb.has[0] = has_below;
  hb.has[1] = has_above;
  for (int s = 0; s < 2; ++s)
    for (int j = 0; j < dim; ++j) {
      hb.lo[s][j] = boxes[s * 2 * dim + j];
      hb.hi[s][j] = boxes[s * 2 * dim + dim + j];
    }
  cudaStream_t st = S(stream);
  // flags buffer is padded by the caller to a multiple of 4 bytes
  DM_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)((N + 3) / 4 * 4), st));
  if (T == 0) return DM_OK;
  if (dim == 2)
    halo_select_kernel<2><<<nblk(T, 128), 128, 0, st>>>(p, t, T, hb, flags);
  else
    halo_select_kernel<3><<<nblk(T, 128), 128, 0, st>>>(p, t, T, hb, flags);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

}  // extern "C"
