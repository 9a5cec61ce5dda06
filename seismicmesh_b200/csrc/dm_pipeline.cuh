// The force-iteration pipeline (stages A-D of include/distmesh_b200.h).
//
//   A  cull_count        centroid + fused SDF program -> keep flag ; incident-cell count per vertex
//      scan              (dm_scan.cuh) counts -> incidence offsets
//   B  inc_fill          vertex -> incident kept cells (integer slot claims; order fixed up in B2)
//      adjacency_build   per vertex: gather incident cells, de-duplicate the neighbour ids in a
//                        thread-private shared-memory hash, sort, write the sorted neighbour row.
//                        Row = [lower neighbours ascending | upper neighbours ascending]; the upper
//                        parts of all rows, in vertex order, ARE the reference's sorted unique bar
//                        list (unique_edges, geometry/cpp/fast_geometry.cpp:30-77).
//   C  bar_pass          L, fh(midpoint), sum L^d, sum h^d over unique bars, fixed-order reduction,
//                        scale = ((sum L^d)/(sum h^d))^(1/d)   (last block finishes the reduction)
//   D  vertex_update     per-vertex gather of bar forces in the reference's COO accumulation order,
//                        pfix, p += dt*F, Newton projection per level, max|F| (last block reduces)
#pragma once
#include "dm_device.cuh"

namespace dm {

constexpr int PL_THREADS = 256;  // cull / fill / bar pass / vertex update
constexpr int AB_THREADS = 128;  // adjacency build (shared-memory hash: H slots per thread)

__device__ __forceinline__ int inc_start(const int32_t* __restrict__ inc_end, int64_t v) {
  return v > 0 ? inc_end[v - 1] : 0;
}

// ---------------------------------------------------------------------------------------------
// A: cull + count.  mode 0: evaluate fd on the centroid, write keep ; 1: keep given ; 2: all kept
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(PL_THREADS) cull_count_kernel(const double* __restrict__ prog,
                                                                const double* __restrict__ p,
                                                                const int32_t* __restrict__ t, int64_t T, double geps,
                                                                int mode, uint8_t* __restrict__ keep,
                                                                int32_t* __restrict__ cnt,
                                                                int32_t* __restrict__ counters) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool k = false;
  if (c < T) {
    int v[4];
    load_cell<DIM>(t, c, v);
    k = true;
    if (mode == 0) {
      double c0, c1, c2;
      cell_centroid<DIM>(p, v, c0, c1, c2);
      k = sdf_eval(prog, DIM, c0, c1, c2) < -geps;
      keep[c] = k ? 1 : 0;
    } else if (mode == 1) {
      k = keep[c] != 0;
    }
    if (k && cnt != nullptr) {
#pragma unroll
      for (int j = 0; j <= DIM; ++j) atomicAdd(cnt + v[j], 1);
    }
  }
  if (counters != nullptr) {
    const int nk = __syncthreads_count(k);
    if (threadIdx.x == 0 && nk) atomicAdd(counters + 1, nk);
  }
}

// B1: inc_end[] holds list starts on entry; each kept cell claims one slot per vertex, so on exit
// inc_end[v] is the END of v's incidence list.
template <int DIM>
__global__ void __launch_bounds__(PL_THREADS) inc_fill_kernel(const int32_t* __restrict__ t, int64_t T,
                                                              const uint8_t* __restrict__ keep,
                                                              int32_t* __restrict__ inc_end,
                                                              int32_t* __restrict__ inc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  if (keep != nullptr && keep[c] == 0) return;
  int v[4];
  load_cell<DIM>(t, c, v);
#pragma unroll
  for (int j = 0; j <= DIM; ++j) inc[atomicAdd(inc_end + v[j], 1)] = (int32_t)c;
}

// ---------------------------------------------------------------------------------------------
// B2: adjacency rows
// ---------------------------------------------------------------------------------------------
template <int DIM>
struct HashCfg {
  static constexpr int LOGH = DIM == 3 ? 6 : 5;
  static constexpr int H = 1 << LOGH;       // slots per vertex (3-D: 64, 2-D: 32)
  static constexpr int MAXLOAD = H - H / 4 - 4;  // beyond this the vertex takes the slow path
};

// thread-private open-addressing set; slot k of thread tid lives at col[k * AB_THREADS]
// (bank = tid % 32 for every k -> conflict free).  ~80 % of the candidates are duplicates that
// hit on the first probe, so that is the straight-line path; new keys / collisions branch out.
template <int DIM>
__device__ __forceinline__ void hash_insert(int32_t* col, int x, int& m) {
  constexpr int LOGH = HashCfg<DIM>::LOGH, H = HashCfg<DIM>::H;
  unsigned h = ((unsigned)x * 2654435761u) >> (32 - LOGH);
  int cur = col[h * AB_THREADS];
  if (cur != x) {
    bool dup = false;
    while (cur != -1 && !dup) {
      h = (h + 1) & (H - 1);
      cur = col[h * AB_THREADS];
      dup = cur == x;
    }
    if (!dup) {
      col[h * AB_THREADS] = x;
      ++m;
    }
  }
}

// the DIM candidates one incident cell contributes to vertex v: every OTHER position of the cell.
// Branch-free: the position holding v is replaced by the last id.  A cell that repeats v (never
// produced by a Delaunay code, but legal input for unique_edges) yields the self bar (v,v),
// exactly like the reference's pair list: reported through `selfbar`.
template <int DIM>
__device__ __forceinline__ void cell_candidates(const int (&ids)[4], int v, int (&x)[3], bool& selfbar) {
  const int last = ids[DIM];
  int self = last == v ? 1 : 0;
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    const bool isv = ids[j] == v;
    self += isv ? 1 : 0;
    x[j] = isv ? last : ids[j];
  }
  selfbar = selfbar || self >= 2;
}

// the neighbours a cell contributes to vertex v: every OTHER position of the cell (a repeated
// vertex id yields the self bar (v,v), exactly like the reference's pair list)
template <int DIM, typename F>
__device__ __forceinline__ void for_each_neighbour(const int (&ids)[4], int v, F&& fn) {
  int self = 0;
#pragma unroll
  for (int j = 0; j <= DIM; ++j) {
    if (ids[j] == v)
      ++self;
    else
      fn(ids[j]);
  }
  if (self >= 2) fn(v);
}

// slow path for a vertex whose neighbour set does not fit the shared hash (hub vertices):
// sorted-unique insertion straight into its global adjacency row (capacity DIM * #incident cells)
template <int DIM>
__device__ int adjacency_row_slow(const int32_t* __restrict__ t, const int32_t* __restrict__ inc, int s, int e,
                                  int v, int32_t* row) {
  int m = 0;
  for (int k = s; k < e; ++k) {
    int ids[4];
    load_cell<DIM>(t, inc[k], ids);
    for_each_neighbour<DIM>(ids, v, [&](int x) {
      int j = m;
      while (j > 0 && row[j - 1] > x) --j;
      if (j > 0 && row[j - 1] == x) return;
      for (int q = m; q > j; --q) row[q] = row[q - 1];
      row[j] = x;
      ++m;
    });
  }
  return m;
}

template <int DIM>
__global__ void __launch_bounds__(AB_THREADS) adjacency_build_kernel(const int32_t* __restrict__ t,
                                                                     const int32_t* __restrict__ inc_end,
                                                                     const int32_t* __restrict__ inc, int64_t N,
                                                                     int32_t* __restrict__ adj,
                                                                     int32_t* __restrict__ deg,
                                                                     int32_t* __restrict__ nlow,
                                                                     int32_t* __restrict__ counters) {
  constexpr int H = HashCfg<DIM>::H;
  constexpr int G = 4;  // incident cells gathered per batch (memory-level parallelism)
  __shared__ int32_t tab[H * AB_THREADS];
  __shared__ int sm[33];
  __shared__ int heavy[AB_THREADS];
  __shared__ int nheavy;
  __shared__ int hcount[3];
  const int tid = threadIdx.x;
  const int64_t v = (int64_t)blockIdx.x * AB_THREADS + tid;
  int m = 0, lo = 0;
  if (tid == 0) nheavy = 0;
  __syncthreads();
  if (v < N) {
    const int s = inc_start(inc_end, v), e = inc_end[v];
    int32_t* col = tab + tid;
#pragma unroll
    for (int k = 0; k < H; ++k) col[k * AB_THREADS] = -1;
    bool ok = true, selfbar = false;
    for (int k = s; k < e && ok; k += G) {
      int c[G];
      int ids[G][4];
#pragma unroll
      for (int u = 0; u < G; ++u) c[u] = k + u < e ? inc[k + u] : -1;
#pragma unroll
      for (int u = 0; u < G; ++u)
        if (c[u] >= 0) load_cell<DIM>(t, c[u], ids[u]);
#pragma unroll
      for (int u = 0; u < G; ++u) {
        if (c[u] >= 0) {
          int x[3];
          cell_candidates<DIM>(ids[u], (int)v, x, selfbar);
#pragma unroll
          for (int j = 0; j < DIM; ++j) hash_insert<DIM>(col, x[j], m);
        }
      }
      ok = m <= HashCfg<DIM>::MAXLOAD;  // at most G*DIM = 12 new keys per batch: the table never fills
    }
    if (ok && selfbar) hash_insert<DIM>(col, (int)v, m);
    int32_t* row = adj + (int64_t)DIM * s;
    if (ok) {
      // compact the occupied slots to the top of the column, then insertion-sort them
      int j = 0;
#pragma unroll 8
      for (int k = 0; k < H; ++k) {
        const int x = col[k * AB_THREADS];
        if (x != -1) {
          col[j * AB_THREADS] = x;
          ++j;
        }
      }
      for (int i = 1; i < m; ++i) {
        const int x = col[i * AB_THREADS];
        int q = i;
        while (q > 0 && col[(q - 1) * AB_THREADS] > x) {
          col[q * AB_THREADS] = col[(q - 1) * AB_THREADS];
          --q;
        }
        col[q * AB_THREADS] = x;
      }
      for (int i = 0; i < m; ++i) {
        const int x = col[i * AB_THREADS];
        row[i] = x;
        lo += (x < (int)v) ? 1 : 0;
      }
    } else {
      // neighbour set too large for the private hash (hull / hub vertices): queue the vertex for
      // the cooperative path below
      heavy[atomicAdd(&nheavy, 1)] = tid;
      m = 0;
    }
    if (ok) {
      deg[v] = m;
      nlow[v] = lo;
    }
  }
  __syncthreads();
  // ---- heavy vertices: the whole block de-duplicates one vertex at a time by rank counting ----
  const int nh = nheavy;
  int heavy_bars = 0;
  for (int ih = 0; ih < nh; ++ih) {
    const int64_t hv = (int64_t)blockIdx.x * AB_THREADS + heavy[ih];
    const int s = inc_start(inc_end, hv), e = inc_end[hv];
    const int n = DIM * (e - s);
    int32_t* row = adj + (int64_t)DIM * s;
    __syncthreads();  // tab / counters free
    if (tid == 0) {
      hcount[0] = 0;
      hcount[1] = 0;
      hcount[2] = 0;
    }
    if (2 * (n + 1) <= H * AB_THREADS) {
      int32_t* val = tab;
      int32_t* first = tab + (n + 1);
      __syncthreads();
      for (int k = s + tid; k < e; k += AB_THREADS) {
        int ids[4], x[3];
        bool sb = false;
        load_cell<DIM>(t, inc[k], ids);
        cell_candidates<DIM>(ids, (int)hv, x, sb);
#pragma unroll
        for (int j = 0; j < DIM; ++j) val[(k - s) * DIM + j] = x[j];
        if (sb) hcount[2] = 1;
      }
      __syncthreads();
      const int nn = n + (hcount[2] ? 1 : 0);
      if (tid == 0 && hcount[2]) val[n] = (int)hv;
      __syncthreads();
      for (int i = tid; i < nn; i += AB_THREADS) {
        const int x = val[i];
        int f = 1;
        for (int j = 0; j < i; ++j)
          if (val[j] == x) {
            f = 0;
            break;
          }
        first[i] = f;
      }
      __syncthreads();
      for (int i = tid; i < nn; i += AB_THREADS) {
        if (!first[i]) continue;
        const int x = val[i];
        int r = 0;
        for (int j = 0; j < nn; ++j) r += (first[j] && val[j] < x) ? 1 : 0;
        row[r] = x;
        atomicAdd(&hcount[0], 1);
        if (x < (int)hv) atomicAdd(&hcount[1], 1);
      }
      __syncthreads();
    } else {
      __syncthreads();
      if (tid == 0) {  // beyond the cooperative capacity: sequential last resort
        const int mm = adjacency_row_slow<DIM>(t, inc, s, e, (int)hv, row);
        int ll = 0;
        for (int i = 0; i < mm; ++i) ll += (row[i] < (int)hv) ? 1 : 0;
        hcount[0] = mm;
        hcount[1] = ll;
      }
      __syncthreads();
    }
    if (tid == 0) {
      deg[hv] = hcount[0];
      nlow[hv] = hcount[1];
      heavy_bars += hcount[0] - hcount[1];
    }
  }
  int total;
  block_exclusive_scan(m - lo + heavy_bars, total, sm);  // unique bars owned by this block's vertices
  if (tid == 0 && total) atomicAdd(counters, total);
}

// bar ids: rowptr[v] = #upper neighbours (scanned afterwards)
__global__ void upper_count_kernel(const int32_t* __restrict__ deg, const int32_t* __restrict__ nlow, int64_t N,
                                   int32_t* __restrict__ rowptr) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) rowptr[v] = deg[v] - nlow[v];
}

template <int DIM>
__global__ void bars_pairs_kernel(const int32_t* __restrict__ inc_end, const int32_t* __restrict__ adj,
                                  const int32_t* __restrict__ deg, const int32_t* __restrict__ nlow,
                                  const int32_t* __restrict__ rowptr, int64_t N, int32_t* __restrict__ pairs) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int32_t* row = adj + (int64_t)DIM * inc_start(inc_end, v);
  const int lo = nlow[v], m = deg[v];
  int64_t e = rowptr[v];
  for (int j = lo; j < m; ++j, ++e) {
    pairs[2 * e] = (int32_t)v;
    pairs[2 * e + 1] = row[j];
  }
}

// h of every unique bar in bar order (diagnostics / parity tests): hmode as in bar_pass_kernel
template <int DIM>
__global__ void bar_sizes_kernel(const DmSizeFn f, const int32_t* __restrict__ inc_end,
                                 const int32_t* __restrict__ deg, const int32_t* __restrict__ nlow,
                                 const int32_t* __restrict__ rowptr, const double* __restrict__ hslot,
                                 const double* __restrict__ hbar, int64_t N, double* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int64_t base = (int64_t)DIM * inc_start(inc_end, v);
  const int lo = nlow[v], m = deg[v];
  int64_t e = rowptr[v];
  for (int j = lo; j < m; ++j, ++e)
    out[e] = f.kind == DM_SIZE_CONST ? f.hconst : (f.kind == DM_SIZE_GRID ? hslot[base + j] : hbar[e]);
}

// bar id of the bar (u, v), u < v: position of v among u's upper neighbours
template <int DIM>
__device__ __forceinline__ int bar_id_of(const int32_t* __restrict__ inc_end, const int32_t* __restrict__ adj,
                                         const int32_t* __restrict__ deg, const int32_t* __restrict__ nlow,
                                         const int32_t* __restrict__ rowptr, int u, int v) {
  const int32_t* row = adj + (int64_t)DIM * inc_start(inc_end, u);
  const int lo = nlow[u], m = deg[u];
  for (int j = lo; j < m; ++j)
    if (row[j] == v) return rowptr[u] + (j - lo);
  return -1;
}

// ---------------------------------------------------------------------------------------------
// C: bar pass.  HMODE 0: constant h ; 1: gridded fh (h stored at the upper directed slot) ;
//               2: h per bar id supplied by the caller ; 3: write bar midpoints only
// ---------------------------------------------------------------------------------------------
template <int DIM, int HMODE>
__global__ void __launch_bounds__(PL_THREADS) bar_pass_kernel(
    const DmSizeFn f, const double* __restrict__ p, const int32_t* __restrict__ inc_end,
    const int32_t* __restrict__ adj, const int32_t* __restrict__ deg, const int32_t* __restrict__ nlow,
    const int32_t* __restrict__ rowptr, int64_t N, double* __restrict__ hslot, const double* __restrict__ hbar,
    double* __restrict__ mid, double* partials, int32_t* done, double* scalars) {
  __shared__ double sm[32];
  __shared__ bool s_last;
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double sL = 0.0, sH = 0.0;
  if (v < N) {
    const int lo = nlow[v], m = deg[v];
    if (m > lo) {
      const int64_t base = (int64_t)DIM * inc_start(inc_end, v);
      const int32_t* row = adj + base;
      double a0, a1, a2;
      load_pt<DIM>(p, v, a0, a1, a2);
      for (int j = lo; j < m; ++j) {
        const int w = row[j];
        double b0, b1, b2, d0, d1, d2;
        load_pt<DIM>(p, w, b0, b1, b2);
        // midpoint p[edges].sum(1)/2 (mesh_generator.py:699)
        const double m0 = (a0 + b0) / 2, m1 = (a1 + b1) / 2, m2 = (a2 + b2) / 2;
        if (HMODE == 3) {
          store_pt<DIM>(mid, rowptr[v] + (j - lo), m0, m1, m2);
          continue;
        }
        const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
        double h;
        if (HMODE == 0) {
          h = f.hconst;
        } else if (HMODE == 1) {
          h = size_eval(f, m0, m1, m2);
          hslot[base + j] = h;
        } else {
          h = hbar[rowptr[v] + (j - lo)];
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
  if (HMODE == 3) return;
  const double bl = block_sum(sL, sm);
  const double bh = block_sum(sH, sm);
  if (threadIdx.x == 0) {
    partials[2 * (int64_t)blockIdx.x] = bl;
    partials[2 * (int64_t)blockIdx.x + 1] = bh;
    __threadfence();
    s_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {  // the last block to finish reduces the partials in a FIXED order
    __threadfence();
    double tL = 0.0, tH = 0.0;
    for (int64_t i = threadIdx.x; i < (int64_t)gridDim.x; i += PL_THREADS) {
      tL += __ldcg(partials + 2 * i);
      tH += __ldcg(partials + 2 * i + 1);
    }
    const double rl = block_sum(tL, sm);
    const double rh = block_sum(tH, sm);
    if (threadIdx.x == 0) {
      scalars[0] = rl;
      scalars[1] = rh;
      const double r = rl / rh;
      scalars[2] = DIM == 2 ? sqrt(r) : pow(r, 1.0 / 3.0);  // ** (1.0 / dim), mesh_generator.py:700
    }
  }
}

// ---------------------------------------------------------------------------------------------
// D: vertex update
// ---------------------------------------------------------------------------------------------
struct Levels {
  const double* prog[DM_MAX_LEVELS];
  int n;
};

template <int DIM, int HMODE>
__global__ void __launch_bounds__(PL_THREADS) vertex_update_kernel(
    const DmSizeFn f, const double* __restrict__ p, double* __restrict__ p_out,
    const int32_t* __restrict__ inc_end, const int32_t* __restrict__ adj, const int32_t* __restrict__ deg,
    const int32_t* __restrict__ nlow, const int32_t* __restrict__ rowptr, const double* __restrict__ hslot,
    const double* __restrict__ hbar, const double* scalars_in, int64_t N, Levels lv, double L0mult, double delta_t,
    double deps, double h0, int64_t nfix, const uint8_t* __restrict__ fixed, double* __restrict__ Ftot,
    double* partials, int32_t* done, double* scalars) {
  __shared__ double sm[32];
  __shared__ bool s_last;
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double f2 = 0.0;
  if (v < N) {
    const double scale = __ldcg(scalars_in + 2);
    const double k0 = L0mult;
    double a0, a1, a2;
    load_pt<DIM>(p, v, a0, a1, a2);
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    const int lo = nlow[v], m = deg[v];
    const int64_t base = (int64_t)DIM * inc_start(inc_end, v);
    const int32_t* row = adj + base;
    // The reference accumulates bar by bar (coo_matrix.toarray): row v receives -Fvec of its
    // lower bars (u,v), u ascending, then +Fvec of its upper bars (v,w), w ascending.  The row is
    // sorted, and -(F/L*(p[u]-p[v])) == (F/L)*(p[v]-p[u]) exactly, so one ascending sweep with
    // d = p[v]-p[nbr] reproduces the reference's sum bit for bit.
    for (int j = 0; j < m; ++j) {
      const int w = row[j];
      double b0, b1, b2, d0, d1, d2;
      load_pt<DIM>(p, w, b0, b1, b2);
      const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
      double h;
      if (HMODE == 0) {
        h = f.hconst;
      } else if (HMODE == 1) {
        if (j >= lo)
          h = hslot[base + j];
        else  // lower bar: same midpoint bits as its owner computed -> same h bits
          h = size_eval(f, (b0 + a0) / 2, (b1 + a1) / 2, (b2 + a2) / 2);
      } else {
        const int e = j >= lo ? rowptr[v] + (j - lo) : bar_id_of<DIM>(inc_end, adj, deg, nlow, rowptr, w, (int)v);
        h = hbar[e];
      }
      double Fs = h * k0 * scale - L;  // L0 - L  (mesh_generator.py:700-702)
      if (Fs < 0) Fs = 0;
      const double q = Fs / L;
      F0 = F0 + q * d0;
      F1 = F1 + q * d1;
      if (DIM == 3) F2 = F2 + q * d2;
    }
    if (v < nfix || (fixed != nullptr && fixed[v])) {  // Ftot[ifix] = 0 (mesh_generator.py:499)
      F0 = 0.0;
      F1 = 0.0;
      F2 = 0.0;
    }
    if (Ftot != nullptr) store_pt<DIM>(Ftot, v, F0, F1, F2);
    f2 = F0 * F0 + F1 * F1;
    if (DIM == 3) f2 = f2 + F2 * F2;
    // p += delta_t * Ftot (mesh_generator.py:502), then one Newton projection per level (:505-506)
    double x0 = a0 + delta_t * F0, x1 = a1 + delta_t * F1, x2 = a2 + delta_t * F2;
    for (int l = 0; l < lv.n; ++l) sdf_project(lv.prog[l], DIM, deps, h0, l, x0, x1, x2);
    store_pt<DIM>(p_out, v, x0, x1, x2);
  }
  const double bm = block_max(f2, sm);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = bm;
    __threadfence();
    s_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double mx = 0.0;
    for (int64_t i = threadIdx.x; i < (int64_t)gridDim.x; i += PL_THREADS) mx = fmax(mx, __ldcg(partials + i));
    const double r = block_max(mx, sm);
    if (threadIdx.x == 0) {
      scalars[3] = r;
      scalars[4] = delta_t * sqrt(r);  // maxdp, mesh_generator.py:514
    }
  }
}

}  // namespace dm
