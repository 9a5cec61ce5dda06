// SDF program interpreter + mesh-size interpolation: the per-point arithmetic of the hot path.
//
// Pure functions, usable from device code and (for the host-side unit tests of the arithmetic
// only, tests/hostsim) from plain C++.  Arithmetic order follows the reference's NumPy
// expressions term by term (cited per primitive) and the file is compiled with -fmad=false, so
// un-rotated trees are bit-identical to the reference's `.eval`.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/distmesh_b200.h"

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#define DM_LDG(ptr) (*(ptr))
#define DM_LDG_KEEP(ptr) ldg_keep(ptr)
#else
#define DM_HD inline
#define DM_LDG(ptr) (*(ptr))
#define DM_LDG_KEEP(ptr) ldg_keep(ptr)
#endif

#define DM_SDF_STACK 8
#define DM_SDF_PSTACK 2

namespace dm {

// a node of an interpolation axis: a few KB that every lookup of every thread reads -- kept in L1 ahead of
// the streaming traffic (the corner records below, the position gathers)
DM_HD double ldg_keep(const double* q) {
#if defined(__CUDA_ARCH__)
  double v;
  asm("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(q));
  return v;
#else
  return *q;
#endif
}

// eight consecutive doubles of a 64-B aligned record: two 256-bit read-only loads on the device.  A record
// is used once (the 3-D grid is far larger than any cache): it does not displace anything in L1.
DM_HD void dm_load8(const double* c, double (&v)[8]) {
#if defined(__CUDA_ARCH__)
  asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(c));
  asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7]) : "l"(c + 4));
#else
  for (int k = 0; k < 8; ++k) v[k] = c[k];
#endif
}

DM_HD double dmin(double a, double b) { return a < b ? a : b; }  // std::min(a,b) semantics
DM_HD double dmax(double a, double b) { return a > b ? a : b; }

// numpy.mod (floored modulo), signed_distance_functions.py:293
DM_HD double floormod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  return r;
}

// signed_distance_functions.py:89-128 (_manipulate): translate back, rotate back (x, then y,
// then z; 2-D: single angle), scale back.
DM_HD void sdf_transform(const double* __restrict__ ins, int flags, int dim, double& x0, double& x1,
                         double& x2) {
  if (flags & DM_TF_TRANSLATE) {
    x0 = x0 - ins[8];
    x1 = x1 - ins[9];
    if (dim == 3) x2 = x2 - ins[10];
  }
  if (dim == 2) {
    if (flags & DM_TF_ROT0) {
      const double c = ins[11], s = ins[12];
      const double a = c * x0 + s * x1;
      const double b = -s * x0 + c * x1;
      x0 = a;
      x1 = b;
    }
  } else {
    if (flags & DM_TF_ROT0) {  // Rx^T
      const double c = ins[11], s = ins[12];
      const double a = c * x1 + s * x2;
      const double b = -s * x1 + c * x2;
      x1 = a;
      x2 = b;
    }
    if (flags & DM_TF_ROT1) {  // Ry^T
      const double c = ins[13], s = ins[14];
      const double a = c * x0 - s * x2;
      const double b = s * x0 + c * x2;
      x0 = a;
      x2 = b;
    }
    if (flags & DM_TF_ROT2) {  // Rz^T
      const double c = ins[15], s = ins[16];
      const double a = c * x0 + s * x1;
      const double b = -s * x0 + c * x1;
      x0 = a;
      x1 = b;
    }
  }
  if (flags & DM_TF_STRETCH) {  // vx = (v.x) v ; x = vx/alpha + (x - vx)   (:111-120)
    const double v0 = ins[17], v1 = ins[18], v2 = ins[19], alpha = ins[20];
    double dot = x0 * v0;
    dot = dot + x1 * v1;
    if (dim == 3) dot = dot + x2 * v2;
    const double w0 = dot * v0, w1 = dot * v1, w2 = dot * v2;
    x0 = w0 / alpha + (x0 - w0);
    x1 = w1 / alpha + (x1 - w1);
    if (dim == 3) x2 = w2 / alpha + (x2 - w2);
  }
}

DM_HD double sdf_rect(const double* __restrict__ prm, double x0, double x1) {
  // fast_geometry.cpp:175-182
  const double s1 = -prm[2] + x1, s2 = prm[3] - x1, s3 = -prm[0] + x0, s4 = prm[1] - x0;
  return -dmin(dmin(dmin(s1, s2), s3), s4);
}

DM_HD double sdf_cube(const double* __restrict__ prm, double x0, double x1, double x2) {
  // fast_geometry.cpp:119-131
  const double s1 = -prm[4] + x2, s2 = prm[5] - x2, s3 = -prm[2] + x1, s4 = prm[3] - x1;
  const double s5 = -prm[0] + x0, s6 = prm[1] - x0;
  return -dmin(dmin(dmin(dmin(dmin(s1, s2), s3), s4), s5), s6);
}

DM_HD double sdf_primitive(int op, const double* __restrict__ ins, double x0, double x1, double x2) {
  const double* prm = ins + 2;
  switch (op) {
    case DM_OP_DISK: {  // :596-598
      const double a = x0 - prm[0], b = x1 - prm[1];
      return sqrt(a * a + b * b) - prm[2];
    }
    case DM_OP_BALL: {  // :601-603
      const double a = x0 - prm[0], b = x1 - prm[1], c = x2 - prm[2];
      return sqrt(a * a + b * b + c * c) - prm[3];
    }
    case DM_OP_RECT:
      return sdf_rect(prm, x0, x1);
    case DM_OP_CUBE:
      return sdf_cube(prm, x0, x1, x2);
    case DM_OP_TORUS: {  // :536-540
      const double q0 = sqrt(fabs(x0) * fabs(x0) + fabs(x2) * fabs(x2)) - prm[0];
      return sqrt(fabs(q0) * fabs(q0) + fabs(x1) * fabs(x1)) - prm[1];
    }
    case DM_OP_PRISM: {  // :557-563 (literal 0.866025, signed x1)
      const double a = fabs(x2) - prm[1];
      const double b = dmax(fabs(x0) * 0.866025 + x1 * 0.5, -x1) - prm[0] * 0.5;
      return dmax(a, b);
    }
    case DM_OP_CYLINDER: {  // :583-590 ; prm = (r, h/2)
      const double l = sqrt(fabs(x0) * fabs(x0) + fabs(x2) * fabs(x2));
      const double d0 = fabs(l) - prm[0];
      const double d1 = fabs(x1) - prm[1];
      const double m0 = dmax(d0, 0.0), m1 = dmax(d1, 0.0);
      return dmin(dmax(d0, d1), 0.0) + sqrt(fabs(m0) * fabs(m0) + fabs(m1) * fabs(m1));
    }
    default:
      return 0.0;
  }
}

// Evaluate a lowered SDF program at one point.
DM_HD double sdf_eval(const double* __restrict__ prog, int dim, double x0, double x1, double x2) {
  const int n = (int)DM_LDG(prog);
  const double* ins = prog + 1;
  if (n == 1) {  // the common case (Disk / Ball / Rectangle / Cube): no stack traffic
    const int op = (int)ins[0];
    sdf_transform(ins, (int)ins[1], dim, x0, x1, x2);
    return sdf_primitive(op, ins, x0, x1, x2);
  }
  double st[DM_SDF_STACK];
  double ps[DM_SDF_PSTACK][3];
  int sp = 0, pp = 0;
  for (int i = 0; i < n; ++i, ins += DM_SDF_WORDS) {
    const int op = (int)ins[0];
    if (op < DM_OP_UNION) {
      double y0 = x0, y1 = x1, y2 = x2;
      sdf_transform(ins, (int)ins[1], dim, y0, y1, y2);
      st[sp++] = sdf_primitive(op, ins, y0, y1, y2);
    } else if (op < DM_OP_REPEAT_BEGIN) {
      const double b = st[--sp];
      const double a = st[--sp];
      const double k = ins[2];
      double r;
      switch (op) {
        case DM_OP_UNION:
          r = dmin(a, b);
          break;
        case DM_OP_INTER:
          r = dmax(a, b);
          break;
        case DM_OP_DIFF:
          r = dmax(a, -b);
          break;
        case DM_OP_SUNION: {  // :332-334
          const double h = dmax(k - fabs(a - b), 0.0);
          r = dmin(a, b) - (h * h * 0.25) / k;
          break;
        }
        case DM_OP_SINTER: {  // :372-374
          const double h = dmax(k - fabs(a - b), 0.0);
          r = dmax(a, b) + h * h * 0.25 / k;
          break;
        }
        default: {  // DM_OP_SDIFF :412-414
          const double h = dmax(k - fabs(-a - b), 0.0);
          r = dmax(-a, b) + (h * h * 0.25) / k;
          break;
        }
      }
      st[sp++] = r;
    } else if (op == DM_OP_REPEAT_BEGIN) {
      ps[pp][0] = x0;
      ps[pp][1] = x1;
      ps[pp][2] = x2;
      ++pp;
      const double P0 = ins[2], P1 = ins[3], P2 = ins[4];
      x0 = floormod(x0 + 0.5 * P0, P0) - 0.5 * P0;
      x1 = floormod(x1 + 0.5 * P1, P1) - 0.5 * P1;
      x2 = floormod(x2 + 0.5 * P2, P2) - 0.5 * P2;
    } else {  // DM_OP_REPEAT_END
      --pp;
      x0 = ps[pp][0];
      x1 = ps[pp][1];
      x2 = ps[pp][2];
      const double v = st[--sp];
      st[sp++] = dmax(v, sdf_cube(ins + 2, x0, x1, x2));
    }
  }
  return st[0];
}

// One forward-difference Newton step towards the zero level set
// (_project_points_back_newton, mesh_generator.py:762-784). Returns true if the point moved.
DM_HD bool sdf_project(const double* __restrict__ prog, int dim, double deps, double h0, int level_idx,
                       double& x0, double& x1, double& x2) {
  const double d = sdf_eval(prog, dim, x0, x1, x2);
  const bool out = level_idx == 0 ? (d > 0.0) : (d > 0.0 && d < h0 / 1.5);
  if (!out) return false;
  const double g0 = (sdf_eval(prog, dim, x0 + deps, x1, x2) - d) / deps;
  const double g1 = (sdf_eval(prog, dim, x0, x1 + deps, x2) - d) / deps;
  double g2v = 0.0;
  double s = 0.0 + g0 * g0;
  s = s + g1 * g1;
  if (dim == 3) {
    g2v = (sdf_eval(prog, dim, x0, x1, x2 + deps) - d) / deps;
    s = s + g2v * g2v;
  }
  if (s < deps) s = deps;
  x0 = x0 - d * g0 / s;
  x1 = x1 - d * g1 / s;
  if (dim == 3) x2 = x2 - d * g2v / s;
  return true;
}

// ---------------------------------------------------------------------------------------------
// gridded fh: scipy RegularGridInterpolator(linear, bounds_error=False, fill_value=None)
// find_indices: largest i with axis[i] <= x, clamped to [0, n-2]  (== clip(searchsorted(right)-1))
// ---------------------------------------------------------------------------------------------
// What grid_find needs of an axis besides the nodes themselves: its first node and the float32 scale of the
// uniform-spacing guess.  Loop invariant: a kernel builds it once per thread (grid_guess) instead of reading
// the axis ends and dividing for every lookup.
struct GridGuess {
  double a0[3];
  float scale[3];
  int cells32;  // the corner-record index of a 3-D grid fits 32 bits (it nearly always does): one 32-bit product chain
};
DM_HD GridGuess grid_guess(const DmSizeFn& f) {
  GridGuess g;
  for (int k = 0; k < 3; ++k) {
    // (an axis that does not exist reads axis 0 instead: every load below is from a valid address whether or
    //  not the compiler turns the guard into a select)
    const int kk = k < f.dim ? k : 0;
    const double* ax = f.axis[kk];
    const int n = f.n[kk];
    const double a0 = DM_LDG(ax), a1 = DM_LDG(ax + n - 1);
    g.a0[k] = a0;
    g.scale[k] = (float)(n - 1) / (float)(a1 - a0);
  }
  g.cells32 = f.dim == 3 && (int64_t)(f.n[0] - 1) * (f.n[1] - 1) * (f.n[2] - 1) < ((int64_t)1 << 31) ? 1 : 0;
  return g;
}

// -> i, lo = ax[i], hi = ax[i+1]
DM_HD int grid_find(const double* __restrict__ ax, int n, double x, double a0, float scale, double& lo, double& hi) {
  // uniform-spacing guess, then fix up against the ACTUAL (float32-rounded) axis.  The guess only
  // has to land within a node or two, so it is computed in float32 (a handful of instructions
  // instead of a float64 division); the loops below make the result exact whatever the guess.
  // Clamp to [0, n-2] without branches (a NaN guess becomes 0: fmaxf returns the other operand).
  const float g = fminf(fmaxf((float)(x - a0) * scale, 0.0f), (float)(n - 2));
  int i = (int)g;
  lo = DM_LDG_KEEP(ax + i);
  hi = DM_LDG_KEEP(ax + i + 1);
  if ((x < lo && i > 0) || (x >= hi && i < n - 2)) {  // the guess is off (a node boundary, or far outside): walk
    while (i > 0 && x < lo) {
      --i;
      hi = lo;
      lo = DM_LDG_KEEP(ax + i);
    }
    while (i < n - 2 && x >= hi) {
      ++i;
      lo = hi;
      hi = DM_LDG_KEEP(ax + i + 1);
    }
  }
  return i;
}

DM_HD double size_eval(const DmSizeFn& f, const GridGuess& gg, double x0, double x1, double x2) {
  if (f.kind == DM_SIZE_CONST) return f.hconst;
  double a00, a01, a10, a11;
  const int i0 = grid_find(f.axis[0], f.n[0], x0, gg.a0[0], gg.scale[0], a00, a01);
  const int i1 = grid_find(f.axis[1], f.n[1], x1, gg.a0[1], gg.scale[1], a10, a11);
  const double y0 = (x0 - a00) / (a01 - a00);
  const double y1 = (x1 - a10) / (a11 - a10);
  if (f.dim == 2) {
    // scipy _rgi_cython.evaluate_linear_2d accumulation order
    const int64_t n1 = f.n[1];
    const double* g = f.grid + (int64_t)i0 * n1 + i1;
    const double v00 = DM_LDG(g), v01 = DM_LDG(g + 1), v10 = DM_LDG(g + n1), v11 = DM_LDG(g + n1 + 1);
    double out = v00 * (1 - y0) * (1 - y1);
    out = out + v01 * (1 - y0) * y1;
    out = out + v10 * y0 * (1 - y1);
    out = out + v11 * y0 * y1;
    return out;
  }
  double a20, a21;
  const int i2 = grid_find(f.axis[2], f.n[2], x2, gg.a0[2], gg.scale[2], a20, a21);
  const double y2 = (x2 - a20) / (a21 - a20);
  const int64_t n1 = f.n[1], n2 = f.n[2];
  double cv[8];  // corner values in itertools.product order
  if (f.cells != nullptr) {
    const double* c = gg.cells32 ? f.cells + (size_t)(((unsigned)i0 * (unsigned)(f.n[1] - 1) + (unsigned)i1) * (unsigned)(f.n[2] - 1) + (unsigned)i2) * 8
                                 : f.cells + (((int64_t)i0 * (n1 - 1) + i1) * (n2 - 1) + i2) * 8;
    dm_load8(c, cv);
  } else {
    const double* g = f.grid + ((int64_t)i0 * n1 + i1) * n2 + i2;
#pragma unroll
    for (int k = 0; k < 8; ++k) cv[k] = DM_LDG(g + ((int64_t)(k >> 2) * n1 + ((k >> 1) & 1)) * n2 + (k & 1));
  }
  // scipy _rgi._evaluate_linear: corners in itertools.product order, weight ((1*w0)*w1)*w2,
  // value = value + v*weight starting from 0.0
  double out = 0.0;
  const double w0[2] = {1 - y0, y0}, w1[2] = {1 - y1, y1}, w2[2] = {1 - y2, y2};
#pragma unroll
  for (int c0 = 0; c0 < 2; ++c0)
#pragma unroll
    for (int c1 = 0; c1 < 2; ++c1)
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const double w = ((1.0 * w0[c0]) * w1[c1]) * w2[c2];
        out = out + cv[c0 * 4 + c1 * 2 + c2] * w;
      }
  return out;
}

DM_HD double size_eval(const DmSizeFn& f, double x0, double x1, double x2) {
  if (f.kind == DM_SIZE_CONST) return f.hconst;
  return size_eval(f, grid_guess(f), x0, x1, x2);
}

// ---------------------------------------------------------------------------------------------
// sliver helpers
// ---------------------------------------------------------------------------------------------
// fast_geometry.cpp:351-418 : dihedral angle i (edge table {2,3},{1,3},{1,2},{0,3},{0,2},{0,1})
DM_HD double dihedral_angle(const double P[4][3], int i) {
  const int e[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
  const int i0 = e[i][0], i1 = e[i][1], i2 = e[5 - i][0], i3 = e[5 - i][1];
  double v1[3], v2[3], v3[3];
  for (int j = 0; j < 3; ++j) {
    v1[j] = P[i1][j] - P[i0][j];
    v2[j] = P[i2][j] - P[i0][j];
    v3[j] = P[i3][j] - P[i0][j];
  }
  auto nrm = [](const double* u) {
    double a = 0.;
    for (int j = 0; j < 3; ++j) a += u[j] * u[j];
    return sqrt(a);
  };
  auto dot = [](const double* a, const double* b) {
    double s = 0.0;
    for (int j = 0; j < 3; ++j) s = s + a[j] * b[j];
    return s;
  };
  const double n1 = nrm(v1), n2 = nrm(v2), n3 = nrm(v3);
  for (int j = 0; j < 3; ++j) {
    v1[j] /= n1;
    v2[j] /= n2;
    v3[j] /= n3;
  }
  const double d23 = dot(v2, v3), d12 = dot(v1, v2), d13 = dot(v1, v3);
  double c12[3], c13[3];
  c12[0] = v1[1] * v2[2] - v1[2] * v2[1];
  c12[1] = v1[2] * v2[0] - v1[0] * v2[2];
  c12[2] = v1[0] * v2[1] - v1[1] * v2[0];
  c13[0] = v1[1] * v3[2] - v1[2] * v3[1];
  c13[1] = v1[2] * v3[0] - v1[0] * v3[2];
  c13[2] = v1[0] * v3[1] - v1[1] * v3[0];
  const double cphi = (d23 - d12 * d13) / (nrm(c12) * nrm(c13));
  return acos(cphi);
}

// fast_geometry.cpp:580-666 : gradient of the circumradius wrt vertex 0
DM_HD void circumsphere_grad(const double* p0, const double* p1, const double* p2, const double* p3,
                             double* g) {
  const double x1 = p0[0] - p3[0], y1 = p0[1] - p3[1], z1 = p0[2] - p3[2];
  const double x2 = p1[0] - p3[0], y2 = p1[1] - p3[1], z2 = p1[2] - p3[2];
  const double x3 = p2[0] - p3[0], y3 = p2[1] - p3[1], z3 = p2[2] - p3[2];
  const double sq1 = x1 * x1 + y1 * y1 + z1 * z1;
  const double sq2 = x2 * x2 + y2 * y2 + z2 * z2;
  const double sq3 = x3 * x3 + y3 * y3 + z3 * z3;
  const double dax = y2 * z3 - y3 * z2, day = z2 * x3 - x2 * z3, daz = x2 * y3 - x3 * y2;
  const double dDx_dx = -2.0 * x1 * dax;
  const double dDx_dy = -2.0 * y1 * dax + sq2 * z3 - sq3 * z2;
  const double dDx_dz = -2.0 * z1 * dax - sq2 * y3 + sq3 * y2;
  const double dDy_dx = -2.0 * x1 * day - sq2 * z3 + sq3 * z2;
  const double dDy_dy = -2.0 * y1 * day;
  const double dDy_dz = -2.0 * z1 * day + sq2 * x3 - sq3 * x2;
  const double dDz_dx = -2.0 * x1 * daz + sq2 * y3 - sq3 * y2;
  const double dDz_dy = -2.0 * y1 * daz - sq2 * x3 + sq3 * x2;
  const double dDz_dz = -2.0 * z1 * daz;
  const double a = x1 * dax + y1 * day + z1 * daz;
  const double Dx = -sq1 * dax + y1 * (sq2 * z3 - sq3 * z2) - z1 * (sq2 * y3 - sq3 * y2);
  const double Dy = -sq1 * day - x1 * (sq2 * z3 - sq3 * z2) + z1 * (sq2 * x3 - sq3 * x2);
  const double Dz = -sq1 * daz + x1 * (sq2 * y3 - sq3 * y2) - y1 * (sq2 * x3 - sq3 * x2);
  const double ssq = Dx * Dx + Dy * Dy + Dz * Dz;
  g[0] = (Dx * dDx_dx + Dy * dDy_dx + Dz * dDz_dx) / (2.0 * a * a) - (dax * ssq) / (2.0 * a * a * a);
  g[1] = (Dx * dDx_dy + Dy * dDy_dy + Dz * dDz_dy) / (2.0 * a * a) - (day * ssq) / (2.0 * a * a * a);
  g[2] = (Dx * dDx_dz + Dy * dDy_dz + Dz * dDz_dz) / (2.0 * a * a) - (daz * ssq) / (2.0 * a * a * a);
}

}  // namespace dm
