// Device helpers shared by all kernels: point / cell loads, centroid, bar length.
#pragma once
#include <float.h>

#include "dm_common.cuh"
#include "dm_sdf.cuh"

namespace dm {

// 256-bit read-only load (LDG.E.ENL2.256 on sm_100a): one request per gathered 3-D point
__device__ __forceinline__ void ldg256(const double* q, double& a, double& b, double& c, double& d) {
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(q));
}
__device__ __forceinline__ void stg256(double* q, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(q), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// PAD: 3-D points come from the plan's padded copy (4 doubles = 32 B per point, 32-B aligned: a
// gather is ONE 256-bit request touching one sector instead of three 8-B requests over two)
template <int DIM, bool PAD = false>
__device__ __forceinline__ void load_pt(const double* __restrict__ p, int64_t i, double& x0, double& x1,
                                        double& x2) {
  if (DIM == 2) {
    const double2 v = *reinterpret_cast<const double2*>(p + 2 * i);  // 16-B aligned rows
    x0 = v.x;
    x1 = v.y;
    x2 = 0.0;
  } else if (PAD) {
    double w;
    ldg256(p + 4 * i, x0, x1, x2, w);
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

// centroid p[t].sum(1)/(dim+1), vertices added in order (mesh_generator.py:737)
template <int DIM, bool PAD = false>
__device__ __forceinline__ void cell_centroid(const double* __restrict__ p, const int (&v)[4], double& c0,
                                              double& c1, double& c2) {
  double q[DIM + 1][3];
#pragma unroll
  for (int k = 0; k <= DIM; ++k) load_pt<DIM, PAD>(p, v[k], q[k][0], q[k][1], q[k][2]);  // all gathers in flight
  double a0 = q[0][0], a1 = q[0][1], a2 = q[0][2];
#pragma unroll
  for (int k = 1; k <= DIM; ++k) {
    a0 = a0 + q[k][0];
    a1 = a1 + q[k][1];
    a2 = a2 + q[k][2];
  }
  c0 = a0 / (double)(DIM + 1);
  c1 = a1 / (double)(DIM + 1);
  c2 = a2 / (double)(DIM + 1);
}

// barvec = a - b ; L = sqrt(sum barvec^2) ; L==0 -> eps   (mesh_generator.py:696-698)
template <int DIM>
__device__ __forceinline__ double bar_length(double a0, double a1, double a2, double b0, double b1, double b2,
                                             double& d0, double& d1, double& d2) {
  d0 = a0 - b0;
  d1 = a1 - b1;
  d2 = a2 - b2;
  double s = d0 * d0 + d1 * d1;
  if (DIM == 3) s = s + d2 * d2;
  double L = sqrt(s);
  if (L == 0.0) L = DBL_EPSILON;
  return L;
}

static inline unsigned nblk(int64_t n, int threads) { return (unsigned)(n <= 0 ? 1 : cdiv(n, threads)); }
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace dm
