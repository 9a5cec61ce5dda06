// TEST INFRASTRUCTURE: runs the per-point arithmetic of seismicmesh_b200/csrc/dm_sdf.cuh on the
// host (plain g++, -ffp-contract=off) so that the SDF program lowering and the interpreter /
// interpolation arithmetic can be unit-tested in the CPU-only container.  Never shipped, never
// used by the product path; the CUDA kernels compile the SAME header.
#include "../../seismicmesh_b200/csrc/dm_sdf.cuh"

extern "C" {

void hs_sdf_eval(const double* prog, const double* x, long M, int dim, double* out) {
  for (long i = 0; i < M; ++i) {
    const double* q = x + (long)dim * i;
    out[i] = dm::sdf_eval(prog, dim, q[0], q[1], dim == 3 ? q[2] : 0.0);
  }
}

void hs_size_eval(const DmSizeFn* f, const double* x, long M, double* out) {
  for (long i = 0; i < M; ++i) {
    const double* q = x + (long)f->dim * i;
    out[i] = dm::size_eval(*f, q[0], q[1], f->dim == 3 ? q[2] : 0.0);
  }
}

void hs_project(const double* prog, double* p, long N, int dim, double deps, double h0, int level) {
  for (long i = 0; i < N; ++i) {
    double* q = p + (long)dim * i;
    double x0 = q[0], x1 = q[1], x2 = dim == 3 ? q[2] : 0.0;
    if (dm::sdf_project(prog, dim, deps, h0, level, x0, x1, x2)) {
      q[0] = x0;
      q[1] = x1;
      if (dim == 3) q[2] = x2;
    }
  }
}

void hs_dihedral(const double* p, const int* t, long T, double* angles) {
  for (long c = 0; c < T; ++c) {
    double P[4][3];
    for (int k = 0; k < 4; ++k)
      for (int j = 0; j < 3; ++j) P[k][j] = p[3 * (long)t[4 * c + k] + j];
    for (int i = 0; i < 6; ++i) angles[6 * c + i] = dm::dihedral_angle(P, i);
  }
}

void hs_circumsphere_grad(const double* p, const int* t, const int* ele, long S, double* g) {
  for (long i = 0; i < S; ++i) {
    const long c = ele[i];
    dm::circumsphere_grad(p + 3 * (long)t[4 * c], p + 3 * (long)t[4 * c + 1], p + 3 * (long)t[4 * c + 2],
                          p + 3 * (long)t[4 * c + 3], g + 3 * i);
  }
}
}
