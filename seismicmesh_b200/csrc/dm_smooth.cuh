// Termination path on the device: Laplacian smoothing of a 2-D mesh as ONE linear solve (replaces
// geometry.laplacian2_fixed_point, SeismicMesh/geometry/utils.py:494-547, which assembles the matrix with
// SciPy and solves it with pyamg's Ruge-Stuben AMG).
//
// The reference's matrix has, for every triangle edge (a, b), +1 on the two diagonal entries and -1 on the
// two off-diagonal ones, and identity rows for the boundary vertices (the vertices of edges that belong to
// one triangle only).  An interior vertex has a closed fan of triangles around it, so each of its edges
// belongs to two triangles: its row reads  2 * (deg(v) * x_v - sum of its neighbours' x) = 0,  i.e. the
// vertex sits at the average of its neighbours; and a vertex is a boundary vertex exactly when it has MORE
// neighbours than incident triangles (an open fan has one more).  Both facts come for free from the
// neighbour rows of stage B and a count of incident cells, so no matrix is assembled: the interior block
// deg * I - Adj is symmetric positive definite and is solved by Jacobi-preconditioned conjugate gradients,
// the two coordinates side by side, two launches per iteration, all scalars (alpha, beta, the dot products:
// fixed-order two-level reductions) staying on the device.
#pragma once
#include "dm_pipeline.cuh"

namespace dm {

constexpr int LS_THREADS = 256;
// scalars: [0,1] r.z  [2,3] p.Ap  [4,5] beta  [6,7] r.r  [8,9] sum (deg x)^2 (the scale the residual is compared with)
constexpr int LS_SCALARS = 16;

struct LapWork {
  int32_t* ntri;      // (N) incident cells
  uint8_t* interior;  // (N)
  double2 *r, *z, *p0, *p1, *Ap;
  double* partials;   // (blocks, 4)
  double* sc;         // LS_SCALARS
  int32_t* done;      // 2 counters
};

__global__ void lap_count_kernel(const int32_t* __restrict__ t, int64_t T, int32_t* __restrict__ ntri) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  for (int j = 0; j < 3; ++j) atomicAdd(ntri + t[3 * c + j], 1);
}

// the last block to arrive adds the per-block partial sums in block order: `nq` quantities per block
__device__ __forceinline__ bool lap_reduce(const double (&q)[4], int nq, double* partials, int32_t* done, double* s_red,
                                           double (&tot)[4]) {
  __shared__ bool s_last;
  for (int k = 0; k < nq; ++k) {
    const double b = block_sum(q[k], s_red);
    if (threadIdx.x == 0) partials[4 * (int64_t)blockIdx.x + k] = b;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  for (int k = 0; k < nq; ++k) {
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < (int64_t)gridDim.x; i += LS_THREADS) a += __ldcg(partials + 4 * i + k);
    tot[k] = block_sum(a, s_red);
  }
  if (threadIdx.x == 0) *done = 0;
  return threadIdx.x == 0;
}

// interior flags, r = b - A x, z = r / deg, p = z ; r.z and the scale
__global__ void __launch_bounds__(LS_THREADS) lap_init_kernel(const Rows<2> R, const double2* __restrict__ x, int64_t N,
                                                             LapWork w) {
  __shared__ double s_red[32];
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double q[4] = {0.0, 0.0, 0.0, 0.0};
  if (v < N) {
    const int2 dg = R.degs[v];
    const int m = dg.x;
    const bool inner = m > 0 && m <= w.ntri[v];  // closed fan(s): as many neighbours as triangles
    w.interior[v] = inner ? 1 : 0;
    double2 r = make_double2(0.0, 0.0), z = r;
    if (inner) {
      const int32_t* row = R.row(v, m);
      double s0 = 0.0, s1 = 0.0;
      for (int j = 0; j < m; ++j) {
        const double2 y = x[row[j]];
        s0 += y.x;
        s1 += y.y;
      }
      const double2 xv = x[v];
      const double d = (double)m;
      r = make_double2(s0 - d * xv.x, s1 - d * xv.y);
      z = make_double2(r.x / d, r.y / d);
      q[0] = r.x * z.x;
      q[1] = r.y * z.y;
      q[2] = (d * xv.x) * (d * xv.x);
      q[3] = (d * xv.y) * (d * xv.y);
    }
    w.r[v] = r;
    w.z[v] = z;
    w.p0[v] = make_double2(0.0, 0.0);  // "previous direction": p = z + beta * 0 in the first iteration
  }
  double tot[4];
  if (lap_reduce(q, 4, w.partials, w.done, s_red, tot)) {
    w.sc[0] = tot[0];
    w.sc[1] = tot[1];
    w.sc[4] = 0.0;
    w.sc[5] = 0.0;
    w.sc[6] = tot[0];  // (not r.r, but only used for the first convergence look, which follows an update)
    w.sc[7] = tot[1];
    w.sc[8] = tot[2];
    w.sc[9] = tot[3];
  }
}

// p_new = z + beta * p_old (also for the neighbours, on the fly) ; Ap = deg * p_new - sum of the neighbours' p_new ; p.Ap
__global__ void __launch_bounds__(LS_THREADS) lap_ap_kernel(const Rows<2> R, int64_t N, LapWork w, const double2* __restrict__ po,
                                                           double2* __restrict__ pn) {
  __shared__ double s_red[32];
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double b0 = __ldcg(w.sc + 4), b1 = __ldcg(w.sc + 5);
  double q[4] = {0.0, 0.0, 0.0, 0.0};
  if (v < N) {
    double2 pv = make_double2(0.0, 0.0), ap = pv;
    if (w.interior[v]) {
      const int m = R.degs[v].x;
      const int32_t* row = R.row(v, m);
      const double2 zv = w.z[v], ov = po[v];
      pv = make_double2(zv.x + b0 * ov.x, zv.y + b1 * ov.y);
      double s0 = 0.0, s1 = 0.0;
      for (int j = 0; j < m; ++j) {  // z and p are zero on boundary vertices
        const int u = row[j];
        const double2 zu = w.z[u], ou = po[u];
        s0 += zu.x + b0 * ou.x;
        s1 += zu.y + b1 * ou.y;
      }
      const double d = (double)m;
      ap = make_double2(d * pv.x - s0, d * pv.y - s1);
      q[0] = pv.x * ap.x;
      q[1] = pv.y * ap.y;
    }
    pn[v] = pv;
    w.Ap[v] = ap;
  }
  double tot[4];
  if (lap_reduce(q, 2, w.partials, w.done + 1, s_red, tot)) {
    w.sc[2] = tot[0];
    w.sc[3] = tot[1];
  }
}

// x += alpha p ; r -= alpha Ap ; z = r / deg ; beta = (r.z)_new / (r.z)_old
__global__ void __launch_bounds__(LS_THREADS) lap_update_kernel(const Rows<2> R, int64_t N, LapWork w, const double2* __restrict__ pn,
                                                               double2* __restrict__ x) {
  __shared__ double s_red[32];
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double rz0 = __ldcg(w.sc + 0), rz1 = __ldcg(w.sc + 1), pap0 = __ldcg(w.sc + 2), pap1 = __ldcg(w.sc + 3);
  const double a0 = pap0 > 0.0 ? rz0 / pap0 : 0.0, a1 = pap1 > 0.0 ? rz1 / pap1 : 0.0;
  double q[4] = {0.0, 0.0, 0.0, 0.0};
  if (v < N && w.interior[v]) {
    const double d = (double)R.degs[v].x;
    const double2 pv = pn[v], ap = w.Ap[v];
    double2 xv = x[v], r = w.r[v];
    xv.x += a0 * pv.x;
    xv.y += a1 * pv.y;
    r.x -= a0 * ap.x;
    r.y -= a1 * ap.y;
    const double2 z = make_double2(r.x / d, r.y / d);
    x[v] = xv;
    w.r[v] = r;
    w.z[v] = z;
    q[0] = r.x * z.x;
    q[1] = r.y * z.y;
    q[2] = r.x * r.x;
    q[3] = r.y * r.y;
  }
  double tot[4];
  if (lap_reduce(q, 4, w.partials, w.done, s_red, tot)) {
    w.sc[4] = rz0 > 0.0 ? tot[0] / rz0 : 0.0;
    w.sc[5] = rz1 > 0.0 ? tot[1] / rz1 : 0.0;
    w.sc[0] = tot[0];
    w.sc[1] = tot[1];
    w.sc[6] = tot[2];
    w.sc[7] = tot[3];
  }
}

}  // namespace dm
