// Stand-alone kernels: fd / fh evaluation, centroids, projection, level-set Newton, cell
// compaction, sliver kernels, halo selection.
#pragma once
#include "dm_device.cuh"

namespace dm {

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

// cell records of a 3-D size grid (DmSizeFn::cells): thread per cell, 64 B written contiguously
__global__ void size_cells_kernel(const double* __restrict__ grid, int n0, int n1, int n2, double* __restrict__ cells) {
  const int64_t nc = (int64_t)(n0 - 1) * (n1 - 1) * (n2 - 1);
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int i2 = (int)(c % (n2 - 1)), i1 = (int)((c / (n2 - 1)) % (n1 - 1)), i0 = (int)(c / ((int64_t)(n2 - 1) * (n1 - 1)));
  const double* g = grid + ((int64_t)i0 * n1 + i1) * n2 + i2;
  double v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = g[((int64_t)(k >> 2) * n1 + ((k >> 1) & 1)) * n2 + (k & 1)];
  stg256(cells + c * 8, v[0], v[1], v[2], v[3]);
  stg256(cells + c * 8 + 4, v[4], v[5], v[6], v[7]);
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

// One sliver_removal pass in one kernel: keep test of the cell (fd(centroid) < -geps,
// mesh_generator.py:734-738) + dihedral bound test of the kept cells (:532-540); flags[c] = 1 for a
// kept cell with an angle out of bounds.  Cell ids stay those of the UNcompacted list: the order of
// the flagged cells is the order the reference sees after its order-preserving cull.
// cos of dihedral angle i from the UNNORMALISED edge vectors: with u1, u2, u3 the three edges from one end of
// the tetrahedron's edge i, cos = ((u2.u3)|u1|^2 - (u1.u2)(u1.u3)) / (|u1 x u2| |u1 x u3|) -- the same quantity
// dihedral_angle feeds to acos, without its three normalisations (nine divisions, three square roots) and with
// one reciprocal square root.  Equal to it up to rounding (~1e-15): good for a SCREEN, not for the answer.
__device__ __forceinline__ double dihedral_cos_screen(const double P[4][3], int i) {
  const int e[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
  const int i0 = e[i][0], i1 = e[i][1], i2 = e[5 - i][0], i3 = e[5 - i][1];
  double u1[3], u2[3], u3[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    u1[j] = P[i1][j] - P[i0][j];
    u2[j] = P[i2][j] - P[i0][j];
    u3[j] = P[i3][j] - P[i0][j];
  }
  const double n11 = u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2];
  const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  const double d13 = u1[0] * u3[0] + u1[1] * u3[1] + u1[2] * u3[2];
  const double d23 = u2[0] * u3[0] + u2[1] * u3[1] + u2[2] * u3[2];
  const double a0 = u1[1] * u2[2] - u1[2] * u2[1], a1 = u1[2] * u2[0] - u1[0] * u2[2], a2 = u1[0] * u2[1] - u1[1] * u2[0];
  const double b0 = u1[1] * u3[2] - u1[2] * u3[1], b1 = u1[2] * u3[0] - u1[0] * u3[2], b2 = u1[0] * u3[1] - u1[1] * u3[0];
  const double aa = a0 * a0 + a1 * a1 + a2 * a2, bb = b0 * b0 + b1 * b1 + b2 * b2;
  return (d23 * n11 - d12 * d13) * rsqrt(aa * bb);
}

// cos_hi = cos(min_dh), cos_lo = cos(max_dh): an angle whose screened cosine lies inside (cos_lo, cos_hi) by
// SCREEN_MARGIN is inside (min_dh, max_dh) whatever the last bits of the reference formula and of acos say
// (near the 10 / 170 degree bounds a cosine margin of 1e-9 is 6e-9 rad, against ~1e-15 of rounding); every other
// angle -- 3 % of them on the ball -- is computed with the reference's own operation order and compared as the
// reference compares it, so the flags are the reference's.
constexpr double SCREEN_MARGIN = 1e-9;
constexpr int SF_THREADS = 128;
__global__ void __launch_bounds__(SF_THREADS) sliver_flags_kernel(const double* __restrict__ prog, const double* __restrict__ p,
                                                                 const int32_t* __restrict__ t, int64_t T, double geps,
                                                                 double min_dh, double max_dh, double cos_lo, double cos_hi,
                                                                 uint8_t* __restrict__ keep, uint8_t* __restrict__ flags) {
  // The angles that need the reference formula are few and scattered over the lanes (a warp would run the whole
  // formula for one lane's sake), so they are QUEUED per block -- (cell, angle) pairs in shared memory -- and
  // worked off densely afterwards, one pair per thread.
  __shared__ unsigned short s_q[6 * SF_THREADS];
  __shared__ int s_n;
  __shared__ unsigned char s_bad[SF_THREADS];
  const int tid = threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + tid;
  if (tid == 0) s_n = 0;
  s_bad[tid] = 0;
  __syncthreads();
  bool kept = false;
  if (c < T) {
    int v[4];
    load_cell<3>(t, c, v);
    double P[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) load_pt<3>(p, v[k], P[k][0], P[k][1], P[k][2]);
    // centroid p[t].sum(1)/4, vertices added in order (mesh_generator.py:737)
    double c0 = P[0][0], c1 = P[0][1], c2 = P[0][2];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      c0 = c0 + P[k][0];
      c1 = c1 + P[k][1];
      c2 = c2 + P[k][2];
    }
    kept = sdf_eval(prog, 3, c0 / 4.0, c1 / 4.0, c2 / 4.0) < -geps;
    if (kept) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double cs = dihedral_cos_screen(P, i);
        if (!(cs > cos_lo + SCREEN_MARGIN && cs < cos_hi - SCREEN_MARGIN))  // (a NaN is queued as well)
          s_q[atomicAdd(&s_n, 1)] = (unsigned short)(tid << 3 | i);
      }
    }
    if (keep != nullptr) keep[c] = kept ? 1 : 0;
  }
  __syncthreads();
  const int nq = s_n;
  for (int e = tid; e < nq; e += SF_THREADS) {
    const int cell = s_q[e] >> 3, i = s_q[e] & 7;
    int v[4];
    load_cell<3>(t, (int64_t)blockIdx.x * blockDim.x + cell, v);
    double Q[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) load_pt<3>(p, v[k], Q[k][0], Q[k][1], Q[k][2]);
    const double a = dihedral_angle(Q, i);  // the reference's operation order
    if ((a < min_dh) || (a > max_dh)) s_bad[cell] = 1;
  }
  __syncthreads();
  if (c < T) flags[c] = s_bad[tid];
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

// Which vertex of a tetrahedron sits in column 0 decides which vertex the sliver perturbation moves
// (mesh_generator.py:234,245-274: "vertex 0 of every sliver, last write wins"); the reference takes
// whatever order CGAL hands out.  Our triangulation stage makes the choice explicit: among the
// vertices that are well inside the domain (key = fd(p[v]) < thresh) one is picked by a hash of the
// cell (so neighbouring slivers do not all move the same vertex), and when there is none the one with
// the smallest key (the most interior).  Boundary vertices are therefore only moved when a cell has
// nothing else to offer.  Applied as an EVEN permutation (orientation is preserved).
__global__ void cells_lead_interior_kernel(const double* __restrict__ key, int32_t* __restrict__ t, int64_t T,
                                           double thresh) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  int4 q = *reinterpret_cast<const int4*>(t + 4 * c);
  const int v[4] = {q.x, q.y, q.z, q.w};
  double k[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) k[j] = key[v[j]];
  int cnt = 0, amin = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    cnt += k[j] < thresh ? 1 : 0;
    if (k[j] < k[amin]) amin = j;  // first minimum (np.argmin)
  }
  int sel = amin;
  if (cnt > 0) {  // the pick-th interior column, pick = (sum of the ids) mod (number of interior columns)
    const int pick = (int)(((int64_t)v[0] + v[1] + v[2] + v[3]) % cnt);
    int seen = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = k[j] < thresh;
      if (in && seen == pick) sel = j;
      seen += in ? 1 : 0;
    }
  }
  // even permutations that bring column `sel` to the front: (0 1 2 3), (1 2 0 3), (2 0 1 3), (3 1 0 2)
  if (sel == 1)
    q = make_int4(v[1], v[2], v[0], v[3]);
  else if (sel == 2)
    q = make_int4(v[2], v[0], v[1], v[3]);
  else if (sel == 3)
    q = make_int4(v[3], v[1], v[0], v[2]);
  *reinterpret_cast<int4*>(t + 4 * c) = q;
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


// ---------------------------------------------------------------------------------------------
// Sizing preprocessing, the elementwise chain of get_sizing_function_from_segy in ONE pass over the
// velocity grid (sizing/mesh_size_function.py): wavelength sizing h = vp / (freq * wl) (:411-426),
// optionally the minimum with a gradient-sizing field (:429-450, computed by the caller), the hmin / hmax
// clamp (:180-181) and the CFL bound (:453-468).  Same operations in the same order as the NumPy
// expressions (no FMA contraction), so the grid is bit-identical to the reference's.
// ---------------------------------------------------------------------------------------------
struct SizingParams {
  double inv_denom;   // unused (kept 0): the division is by freq * wl, as NumPy does it
  double freq_wl;     // freq * wl ; <= 0: no wavelength sizing (h = 99999)
  double hmin, hmax;
  double dt;          // 0: no CFL bound
  double dimf;        // (double) dim
  double cr_lim;      // cr_max / (dim * space_order)
  double dim_cr_lim;  // dim * cr_lim
};
__global__ void sizing_elementwise_kernel(const double* __restrict__ vp, const double* __restrict__ h_gr, int64_t n,
                                          SizingParams q, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = vp[i];
  // np.minimum(h_wl, h_gr) with the reference's 99999 stand-ins for a term that is switched off; with both
  // switched off the grid keeps its initial value hmin (:162-166)
  double h = q.freq_wl > 0.0 ? v / q.freq_wl : (h_gr != nullptr ? 99999.0 : q.hmin);
  h = fmin(h, h_gr != nullptr ? h_gr[i] : (q.freq_wl > 0.0 ? 99999.0 : h));
  if (h < q.hmin) h = q.hmin;
  if (h > q.hmax) h = q.hmax;
  if (q.dt != 0.0) {
    const double cr_old = (v * q.dt) / (q.dimf * h);
    if (cr_old > q.cr_lim) h = (v * q.dt) / q.dim_cr_lim;
  }
  out[i] = h;
}

// ---------------------------------------------------------------------------------------------
// gradient limiting of a gridded size function (sizing/cpp/FastHJ.cpp:63-157, c_limgrad): the
// reference relaxes node pairs of the 6-edge stencil in a sequential active-set sweep until no
// pair differs by more than delta (+ ftol).  The operator only ever lowers values and, up to the ftol band
// of its acceptance test, any relaxation order converges to the same fixed point (min-plus closure over
// grid paths).  Here every sweep is a JACOBI step (read one buffer, write the other), so the result is
// the same from run to run -- an in-place sweep is a chaotic Gauss-Seidel whose result can differ by about
// ftol per node between runs, enough to flip a rejection-sampling decision of the initial points.
// ---------------------------------------------------------------------------------------------
__global__ void limgrad_sweep_kernel(const double* __restrict__ src, double* __restrict__ dst, int n0, int n1, int n2,
                                     double delta, double ftol, int32_t* __restrict__ changed) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k2 = (int)(i % n2), k1 = (int)((i / n2) % n1), k0 = (int)(i / ((int64_t)n2 * n1));
  const int64_t s1 = n2, s0 = (int64_t)n1 * n2;
  const double v = src[i];
  double m = v;
  if (k2 > 0) m = fmin(m, src[i - 1]);
  if (k2 < n2 - 1) m = fmin(m, src[i + 1]);
  if (k1 > 0) m = fmin(m, src[i - s1]);
  if (k1 < n1 - 1) m = fmin(m, src[i + s1]);
  if (k0 > 0) m = fmin(m, src[i - s0]);
  if (k0 < n0 - 1) m = fmin(m, src[i + s0]);
  const double cand = m + delta;  // FastHJ.cpp:138,147
  double out = v;
  if (v > cand + ftol) {
    out = cand;
    *changed = 1;
  }
  dst[i] = out;
}

// ---------------------------------------------------------------------------------------------
// halo push: the rows of p listed in idx go straight into a neighbour GPU's ghost buffer through
// its peer mapping (NVLink stores; dst is a device pointer into the PEER's memory).  Replaces
// pack + ncclSend/ncclRecv for the per-iteration ghost exchange (migration.exchange,
// migration/migration.py:148-183).
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void halo_push_kernel(const double* __restrict__ p, const int32_t* __restrict__ idx, int64_t n,
                                 double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x0, x1, x2;
  load_pt<DIM>(p, idx[i], x0, x1, x2);
  store_pt<DIM>(dst, i, x0, x1, x2);
}


// ---------------------------------------------------------------------------------------------
// Halo exchange as TWO launches per iteration (round 2): `halo_push2_kernel` stores the rows this rank
// exports to rank-1 and rank+1 straight into the GHOST ROWS of the neighbours' position buffers (peer
// mappings of their symmetric-memory buffers: NVLink stores, no staging buffer, no unpack copy) and the
// last block to finish raises a stamp on each neighbour; `halo_wait_kernel` holds the stream until both
// neighbours' stamps for this step have arrived.  Every block makes its stores visible system-wide
// (fence.sc.sys) before it arrives on the local counter; the last arrival (acquire) publishes the stamp
// with a system-scope release store, the waiter reads it with a system-scope acquire load.
// ---------------------------------------------------------------------------------------------
struct HaloPush {
  const int32_t* idx[2];   // local rows exported below / above
  int64_t n[2];
  double* dst[2];          // first ghost row of this rank's block in the neighbour's buffer (peer address)
  unsigned long long* flag[2];  // the neighbour's "arrived from above" / "arrived from below" stamp (peer address)
};
template <int DIM>
__global__ void halo_push2_kernel(const double* __restrict__ p, HaloPush h, unsigned long long stamp, int32_t* done) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int side = i < h.n[0] ? 0 : 1;
  const int64_t k = side == 0 ? i : i - h.n[0];
  if (k < h.n[side]) {
    double x0, x1, x2;
    load_pt<DIM>(p, h.idx[side][k], x0, x1, x2);
    store_pt<DIM>(h.dst[side], k, x0, x1, x2);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && arrive_acq_rel(done) == (int)gridDim.x - 1) {
    *done = 0;
    for (int s = 0; s < 2; ++s)
      if (h.flag[s] != nullptr)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(h.flag[s]), "l"(stamp) : "memory");
  }
}

// One thread per awaited neighbour spins until its stamp reaches `stamp` (stamps only grow).  The spin is
// bounded (~2 s of SM clock): a neighbour that never arrives raises *err instead of hanging the GPU.
__global__ void halo_wait_kernel(const unsigned long long* f0, const unsigned long long* f1, unsigned long long stamp,
                                 int32_t* err) {
  const unsigned long long* f = threadIdx.x == 0 ? f0 : f1;
  if (f == nullptr) return;
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    if (v >= stamp) break;
    if (clock64() - t0 > 4000000000LL) {
      atomicExch(err, 1);
      break;
    }
    __nanosleep(64);
  }
}

// ---------------------------------------------------------------------------------------------
// np.pad on the device (domain extension of the sizing grid, sizing/mesh_size_function.py:526-587, which calls
// np.pad with mode "edge", "constant" or "linear_ramp").  NumPy pads one axis after the other, axis 0 first, and
// while it pads axis k its region of interest is the full (already padded) extent along the axes before k and the
// ORIGINAL extent along the axes after k (numpy/lib/_arraypad_impl.py, _view_roi); the pads of axis k are computed
// from the edge planes of that region.  One launch per stage here, the same order, the same arithmetic:
// linear_ramp is np.linspace(end_value, edge, width, endpoint=False): j * ((edge - end) / width) + end -- or, when
// ANY element of the edge plane equals the end value (a zero step somewhere), (j / width) * (edge - end) + end for
// the whole plane (numpy/_core/function_base.py linspace: `any_step_zero`); hence the flag pass before each fill.
// ---------------------------------------------------------------------------------------------
struct PadGeom {
  int n[3];   // padded shape (2-D: n[2] = 1)
  int lo[3];  // original area [lo, hi) along every axis
  int hi[3];
};

__global__ void pad_copy_kernel(const double* __restrict__ in, double* __restrict__ out, PadGeom g) {
  const int64_t m0 = g.hi[0] - g.lo[0], m1 = g.hi[1] - g.lo[1], m2 = g.hi[2] - g.lo[2];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m0 * m1 * m2) return;
  const int64_t i2 = i % m2, i1 = (i / m2) % m1, i0 = i / (m2 * m1);
  out[((i0 + g.lo[0]) * g.n[1] + (i1 + g.lo[1])) * g.n[2] + (i2 + g.lo[2])] = in[i];
}

// extent of the region of interest of stage `axis` along axis q
__device__ __forceinline__ void pad_roi(const PadGeom& g, int axis, int q, int& b, int& e) {
  if (q < axis) {
    b = 0;
    e = g.n[q];
  } else {
    b = g.lo[q];
    e = g.hi[q];
  }
}

// flags[0] / flags[1] = 1 if some element of the lower / upper edge plane equals end_lo / end_hi
__global__ void pad_flags_kernel(const double* __restrict__ out, PadGeom g, int axis, double end_lo, double end_hi,
                                 int32_t* __restrict__ flags) {
  int b[3], e[3];
  for (int q = 0; q < 3; ++q) pad_roi(g, axis, q, b[q], e[q]);
  const int qa = axis == 0 ? 1 : 0, qb = axis == 2 ? 1 : 2;  // the two other axes
  const int64_t na = e[qa] - b[qa], nb = e[qb] - b[qb];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na * nb) return;
  int idx[3];
  idx[qa] = b[qa] + (int)(i / nb);
  idx[qb] = b[qb] + (int)(i % nb);
  idx[axis] = g.lo[axis];
  if (out[((int64_t)idx[0] * g.n[1] + idx[1]) * g.n[2] + idx[2]] == end_lo) flags[0] = 1;
  idx[axis] = g.hi[axis] - 1;
  if (out[((int64_t)idx[0] * g.n[1] + idx[1]) * g.n[2] + idx[2]] == end_hi) flags[1] = 1;
}

// mode 0 edge, 1 constant (= end value), 2 linear_ramp
__global__ void pad_fill_kernel(double* __restrict__ out, PadGeom g, int axis, int mode, double end_lo, double end_hi,
                                const int32_t* __restrict__ flags) {
  int b[3], e[3];
  for (int q = 0; q < 3; ++q) pad_roi(g, axis, q, b[q], e[q]);
  const int qa = axis == 0 ? 1 : 0, qb = axis == 2 ? 1 : 2;
  const int64_t na = e[qa] - b[qa], nb = e[qb] - b[qb];
  const int wl = g.lo[axis], wr = g.n[axis] - g.hi[axis];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na * nb * (wl + wr)) return;
  // the pad index runs fastest when the padded axis is the last one (coalesced stores), else the last other axis does
  int64_t r = i;
  int jj, ia, ib;
  if (axis == 2) {
    jj = (int)(r % (wl + wr));
    r /= (wl + wr);
    ib = (int)(r % nb);
    ia = (int)(r / nb);
  } else {
    ib = (int)(r % nb);
    r /= nb;
    jj = (int)(r % (wl + wr));
    ia = (int)(r / (wl + wr));
  }
  int idx[3];
  idx[qa] = b[qa] + ia;
  idx[qb] = b[qb] + ib;
  const bool lower = jj < wl;
  const int W = lower ? wl : wr;
  const int j = lower ? jj : wr - 1 - (jj - wl);  // position on the ramp counted from the OUTER end
  idx[axis] = lower ? g.lo[axis] : g.hi[axis] - 1;
  const double edge = out[((int64_t)idx[0] * g.n[1] + idx[1]) * g.n[2] + idx[2]];
  const double endv = lower ? end_lo : end_hi;
  double v;
  if (mode == 0) {
    v = edge;
  } else if (mode == 1) {
    v = endv;
  } else {
    const double delta = edge - endv;
    const double step = delta / (double)W;
    v = flags[lower ? 0 : 1] ? ((double)j / (double)W) * delta : (double)j * step;
    v = v + endv;
  }
  idx[axis] = lower ? j : g.n[axis] - 1 - j;
  out[((int64_t)idx[0] * g.n[1] + idx[1]) * g.n[2] + idx[2]] = v;
}

// ---------------------------------------------------------------------------------------------
// scipy.ndimage.uniform_filter along ONE axis (mode "reflect", origin 0), as the `grad=` option of the sizing
// preprocessing calls it (sizing/mesh_size_function.py:428-448).  SciPy's uniform_filter1d keeps a running
// sum per line -- tmp = sum of the first `size` extended samples; out[0] = tmp / size; then
// tmp += ext[i + size - 1] - ext[i - 1]; out[i] = tmp / size -- and uniform_filter applies it axis after axis;
// the same recurrence here (one thread per line, so the rounding sequence is SciPy's), bit-identical
// (restated and checked against SciPy on the CPU: oracle.uniform_filter1d_lines).
// ---------------------------------------------------------------------------------------------
__global__ void uniform_filter_axis_kernel(const double* __restrict__ in, double* __restrict__ out, int n0, int n1, int n2,
                                           int axis, int size, int square) {
  const int n[3] = {n0, n1, n2};
  const int len = n[axis];
  const int64_t lines = (int64_t)n0 * n1 * n2 / len;
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= lines) return;
  // line l: position along the two other axes; consecutive threads walk the last remaining axis
  int64_t base, stride;
  if (axis == 2) {
    base = l * n2;
    stride = 1;
  } else if (axis == 1) {
    base = (l / n2) * (int64_t)n1 * n2 + l % n2;
    stride = n2;
  } else {
    base = l;
    stride = (int64_t)n1 * n2;
  }
  const int left = size / 2;
  // extended sample e (0 <= e < len + size - 1) -> reflected index into the line: (d c b a | a b c d | d c b a)
  auto ext = [&](int e) {
    int i = e - left;
    if (i < 0) i = -i - 1;
    if (i >= len) i = 2 * len - 1 - i;
    const double x = in[base + (int64_t)i * stride];
    return square ? x * x : x;  // (the filter of vp**2 without a squared copy of the model)
  };
  double tmp = 0.0;
  for (int k = 0; k < size; ++k) tmp += ext(k);
  out[base] = tmp / (double)size;
  for (int i = 1; i < len; ++i) {
    tmp += ext(i + size - 1) - ext(i - 1);
    out[base + (int64_t)i * stride] = tmp / (double)size;
  }
}

// h_gr = grad / ((var / vmax - vmin_scaled) + 0.10), var = sqr_mean - mean * mean (mesh_size_function.py:441-448);
// pass 0 leaves var in `out` (the caller takes its extrema), pass 1 finishes in place
__global__ void variance_size_kernel(const double* __restrict__ mean, const double* __restrict__ sqr_mean, int64_t n, int pass,
                                     double vmax, double vmin_scaled, double grad, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pass == 0) {
    const double m = mean[i];
    out[i] = sqr_mean[i] - m * m;
  } else {
    double v = out[i] / vmax;
    v = v - vmin_scaled;
    out[i] = grad / (v + 0.10);
  }
}

// vp[i] < thresh -> value, in place; *count += how many (the water layer of a shear-velocity model,
// sizing/mesh_size_function.py:148-159)
__global__ void replace_below_kernel(double* __restrict__ a, int64_t n, double thresh, double value, int do_replace,
                                     unsigned long long* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool hit = i < n && a[i] < thresh;
  if (hit && do_replace) a[i] = value;
  const unsigned m = __ballot_sync(FULL, hit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

}  // namespace dm
