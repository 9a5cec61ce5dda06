// libdistmesh_b200: hand-written sm_100a kernels for the DistMesh force iteration of
// SeismicMesh (generation/mesh_generator.py:460-527) behind the C ABI of
// include/distmesh_b200.h.  Compiled with -fmad=false: fp64 everywhere, no FMA contraction.
//
// Data layout in HBM (all caller-owned): p (N,dim) f64 AoS, t (T,dim+1) i32 AoS; per-iteration
// scratch carved from one workspace (DmPlan).  Every kernel is gather/scan/segmented-sum work
// bound by HBM bandwidth; nothing here is a dense contraction, so tensor cores are not used.
//
//   dm_sdf.cuh       per-point arithmetic (SDF interpreter, fh interpolation, sliver formulas)
//   dm_scan.cuh      single-pass look-back scan
//   dm_pipeline.cuh  stages A-D of the force iteration
//   dm_aux.cuh       stand-alone kernels (fd/fh eval, projection, compaction, sliver, halo)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dm_aux.cuh"
#include "dm_pipeline.cuh"
#include "dm_scan.cuh"
#include "dm_smooth.cuh"
#include "dm_tiles.cuh"

using namespace dm;

namespace {

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

inline bool bad_dim(int dim) { return dim != 2 && dim != 3; }

int check_size_fn(const DmSizeFn* f, int dim) {
  if (!f) return DM_ERR_ARG;
  if (f->kind == DM_SIZE_CONST || f->kind == DM_SIZE_EXTERNAL) return DM_OK;
  if (f->kind != DM_SIZE_GRID || f->dim != dim || !f->grid) return DM_ERR_ARG;
  for (int k = 0; k < f->dim; ++k)
    if (f->n[k] < 2 || !f->axis[k]) return DM_ERR_ARG;
  return DM_OK;
}

template <int DIM>
size_t plan_layout_dim(DmPlan* pl, int64_t N, int64_t T, char* base) {
  constexpr int CAP = PCfg<DIM>::CAP, RS = PCfg<DIM>::RS;
  const size_t esz = sizeof(typename PCfg<DIM>::entry_t);
  const int64_t K = (int64_t)DIM * (DIM + 1) * T;
  const int64_t K1 = K > 0 ? K : 1, T1 = T > 0 ? T : 1;
  const int64_t heap_ints = K1 + 4 * (N + 1);
  // bar sums: one per adjacency block + one per group of RG blocks + one per heavy vertex (dm_pipeline.cuh)
  // (the leaves are the 32-vertex warps of rows_kernel; the lane-group kernel of round 1 has fewer blocks)
  const int64_t nbm = cdiv(N > 0 ? N : 1, 32);
  const int64_t ngrp = cdiv(nbm, RG);
  const int64_t nblocks = nbm + ngrp + (N + 2) + 8;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* ptr = base ? base + off : nullptr;
    off += align256(bytes);
    return ptr;
  };
  char* keep = take((size_t)T1);
  // ---- zero region (contiguous): cnt | sync | counters | gdone : zeroed by the prep kernel of stage A
  const size_t z0 = off;
  char* cnt = take((size_t)(N + 1) * 4);
  char* sync = take(8 * 4);
  char* counters = take(8 * 4);
  char* gdone = take((size_t)(ngrp + 1) * 4);
  const size_t zbytes = off - z0;
  // ----
  // third layout: per-vertex buckets ; fourth layout (dm_tiles.cuh, the default): per-tile record lists.
  // One region, sized for either.
  const size_t bucket_bytes = (size_t)(N + 1) * CAP * esz;
  const size_t trec_bytes = (size_t)cdiv(N + 1, TL_R) * TCfg<DIM>::CAPT * sizeof(int4);
  char* bucket = take(bucket_bytes > trec_bytes ? bucket_bytes : trec_bytes);
  char* ovf_v = take((size_t)(DIM + 1) * T1 * 4);
  char* ovf_e = take((size_t)(DIM + 1) * T1 * sizeof(int4));  // tile layout: 16-B records in 2-D as well
  char* hv = take((size_t)(N + 1) * 4);
  char* adj = take((size_t)(N + 1) * RS * 4);
  char* heap = take((size_t)heap_ints * 4);
  char* degs = take((size_t)(N + 1) * 8);
  char* rowptr = take((size_t)(N + 1) * 4);
  char* hslot = take((size_t)((N + 1) * RS + heap_ints) * 8);
  char* hbar = take((size_t)(K1 / 2 + 1) * 8);
  char* partials = take((size_t)(2 * nblocks) * 8);
  char* scalars = take(8 * 8);
  char* p4 = take(DIM == 3 ? (size_t)(N + 1) * 32 : 0);
  char* esc = take((size_t)(N + 1) * 4);
  const size_t scan_bytes = scan_scratch_bytes(N + 1);
  char* scan_tmp = take(scan_bytes);
  if (pl) {
    pl->N = N;
    pl->T = T;
    pl->dim = DIM;
    pl->_pad0 = 0;
    pl->K = K;
    pl->keep = reinterpret_cast<uint8_t*>(keep);
    pl->zero_base = cnt;
    pl->zero_bytes = zbytes;
    pl->cnt = reinterpret_cast<int32_t*>(cnt);
    pl->sync = reinterpret_cast<int32_t*>(sync);
    pl->counters = reinterpret_cast<int32_t*>(counters);
    pl->gdone = reinterpret_cast<int32_t*>(gdone);
    pl->bucket = bucket;
    pl->ovf_v = reinterpret_cast<int32_t*>(ovf_v);
    pl->ovf_e = ovf_e;
    pl->hv = reinterpret_cast<int32_t*>(hv);
    pl->adj = reinterpret_cast<int32_t*>(adj);
    pl->heap = reinterpret_cast<int32_t*>(heap);
    pl->degs = reinterpret_cast<int32_t*>(degs);
    pl->rowptr = reinterpret_cast<int32_t*>(rowptr);
    pl->hslot = reinterpret_cast<double*>(hslot);
    pl->hbar = reinterpret_cast<double*>(hbar);
    pl->partials = reinterpret_cast<double*>(partials);
    pl->scalars = reinterpret_cast<double*>(scalars);
    pl->p4 = DIM == 3 ? reinterpret_cast<double*>(p4) : nullptr;
    pl->esc = reinterpret_cast<int32_t*>(esc);
    pl->scan_tmp = scan_tmp;
    pl->scan_tmp_bytes = scan_bytes;
    pl->n_rows = N;
    pl->layout = DM_LAYOUT_AUTO;
  }
  return off;
}

size_t plan_layout(DmPlan* pl, int64_t N, int64_t T, int dim, char* base) {
  return dim == 2 ? plan_layout_dim<2>(pl, N, T, base) : plan_layout_dim<3>(pl, N, T, base);
}

template <int DIM>
Rows<DIM> rows_of(const DmPlan* pl) {
  Rows<DIM> R;
  R.adj = pl->adj;
  R.heap = pl->heap;
  R.degs = reinterpret_cast<const int2*>(pl->degs);
  R.N = pl->N;
  return R;
}

template <int DIM>
int launch_bar_pass(const DmPlan* pl, const double* p, const DmSizeFn& f, int hmode, double* mid, cudaStream_t st) {
  const unsigned nb = nblk(pl->n_rows, PL_THREADS);
#define DM_BP(H)                                                                                              \
  bar_pass_kernel<DIM, H><<<nb, PL_THREADS, 0, st>>>(f, p, rows_of<DIM>(pl), pl->rowptr, pl->n_rows, pl->hslot, \
                                                     pl->hbar, mid, pl->partials, pl->sync + 1, pl->scalars)
  switch (hmode) {
    case 0: DM_BP(0); break;
    case 1: DM_BP(1); break;
    case 2: DM_BP(2); break;
    default: DM_BP(3); break;
  }
#undef DM_BP
  return (int)cudaGetLastError();
}

// pad: gather neighbour positions from the plan's padded copy (made by stage A from this same p)
template <int DIM>
int launch_vertex_update(const DmPlan* pl, const double* p, bool pad, double* p_out, const Levels& lv,
                         const DmSizeFn& f, int hmode, double L0mult, double delta_t, double deps, double h0,
                         int64_t nfix, const uint8_t* fixed, double* Ftot, cudaStream_t st) {
  const unsigned nb = nblk(pl->n_rows, VU_THREADS);  // vertices without a row (ghost copies) are not updated
  const double* pg = pad ? pl->p4 : p;
  const bool fuse = lv.n > 0 && pl->n_rows < DM_FUSE_PROJECT_BELOW;  // (see DM_FUSE_PROJECT_BELOW)
#define DM_VU_(H, P, F)                                                                                             \
  launch_chain(vertex_update_kernel<DIM, H, P, F>, nb, VU_THREADS, st, f, p, pg, p_out, rows_of<DIM>(pl), pl->rowptr, \
               pl->hslot, pl->hbar, pl->scalars, pl->n_rows, lv, L0mult, delta_t, deps, h0, nfix, fixed, Ftot,        \
               pl->partials, pl->sync + 2, pl->scalars, pl->esc, pl->counters + 5)
#define DM_VU(H, P)      \
  do {                   \
    if (fuse)            \
      DM_VU_(H, P, true); \
    else                 \
      DM_VU_(H, P, false); \
  } while (0)
  if (DIM == 3 && pad) {
    switch (hmode) {
      case 0: DM_VU(0, true); break;
      case 1: DM_VU(1, true); break;
      default: return DM_ERR_ARG;
    }
  } else {
    switch (hmode) {
      case 0: DM_VU(0, false); break;
      case 1: DM_VU(1, false); break;
      default: DM_VU(2, false); break;
    }
  }
#undef DM_VU
#undef DM_VU_
  mark("vertex_update+maxdp", st);
  if (lv.n > 0 && !fuse) {  // Newton projection of the listed (escaped) vertices
    launch_chain(project_list_kernel<DIM>, PJ_BLOCKS, PJ_THREADS, st, lv, deps, h0, pl->esc, pl->counters + 5,
                 pl->sync + 4, p_out);
    mark("project_escaped", st);
  }
  return (int)cudaGetLastError();
}

// Stages A + B: which layout (include/distmesh_b200.h, DM_LAYOUT_*).  A pure function of the plan, so the
// stages of one iteration agree; DM_TILES=0 / 1 in the environment overrides the plan (comparison runs, and
// the GPU suite is run both ways).
constexpr int64_t TILES_FROM_ROWS = 200000;
static bool use_tiles(const DmPlan* pl) {
  static const int forced = [] {
    const char* e = getenv("DM_TILES");
    return !e ? -1 : (e[0] == '0' ? 0 : 1);
  }();
  if (forced >= 0) return forced == 1;
  if (pl->layout == DM_LAYOUT_AUTO) return pl->n_rows >= TILES_FROM_ROWS;
  return pl->layout == DM_LAYOUT_TILES;
}

// (per call, not once per process: the attribute belongs to the CURRENT device's copy of the kernel)
template <int DIM, int BAR>
static cudaError_t tile_smem_ready(size_t bytes) {
  return cudaFuncSetAttribute(tile_rows_kernel<DIM, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int DIM>
static int stage_cull_scatter(const DmPlan* pl, const double* prog, const double* p, const int32_t* t, double geps,
                              int mode, cudaStream_t st, int64_t cell0 = 0, int64_t ncells = -1) {
  typedef typename PCfg<DIM>::entry_t entry_t;
  if (ncells < 0) ncells = pl->T;
  // cells [cell0, cell0 + ncells): t points at the first of them
  const unsigned nb = nblk(ncells, DM_CS_THREADS);
  const double* pc = DIM == 3 ? pl->p4 : p;  // 3-D: the padded copy made by the prep kernel
  if (use_tiles(pl)) {
    launch_chain(cull_bin_kernel<DIM, DIM == 3, DM_CB_CPT>, nblk(ncells, DM_CS_THREADS * DM_CB_CPT), DM_CS_THREADS, st, prog, pc, t,
                 ncells, geps, mode, pl->keep + cell0,
                 pl->cnt, static_cast<int4*>(pl->bucket), pl->ovf_v, static_cast<int4*>(pl->ovf_e), pl->counters,
                 (int)pl->n_rows);
    mark("cull_bin", st);
    return (int)cudaGetLastError();
  }
  launch_chain(cull_scatter_kernel<DIM, DIM == 3>, nb, DM_CS_THREADS, st, prog, pc, t, ncells, geps, mode,
               pl->keep + cell0, pl->cnt, static_cast<entry_t*>(pl->bucket), pl->ovf_v,
               static_cast<entry_t*>(pl->ovf_e), pl->hv, pl->counters, (int)pl->n_rows);
  mark("cull_scatter", st);
  return (int)cudaGetLastError();
}

// bar: -1 rows only (staged path: a separate bar pass follows) ; otherwise f->kind (0 const, 1 grid):
// the bar pass and the reduction to the scale are fused into the adjacency kernel
template <int DIM>
static int stage_adjacency(const DmPlan* pl, int bar, const DmSizeFn* f, const double* p, cudaStream_t st) {
  typedef typename PCfg<DIM>::entry_t entry_t;
  const int64_t N = pl->N;
  int2* degs = reinterpret_cast<int2*>(pl->degs);
  const entry_t* bucket = static_cast<const entry_t*>(pl->bucket);
  const entry_t* ovf_e = static_cast<const entry_t*>(pl->ovf_e);
  DmSizeFn fz;
  memset(&fz, 0, sizeof(fz));
  const DmSizeFn& ff = f ? *f : fz;
  const double* pp = (DIM == 3 && p) ? pl->p4 : p;  // bar pass gathers from the padded copy
  if (use_tiles(pl)) {
    const unsigned nb = nblk(cdiv(pl->n_rows, TL_R), TL_WPB);
    constexpr size_t smem = (size_t)TL_WPB * tile_warp_ints<DIM>() * sizeof(int32_t);
#define DM_TILE_LAUNCH(B)                                                                                          \
  DM_CUDA_TRY((tile_smem_ready<DIM, B>(smem)));                                                                    \
  launch_chain_smem(tile_rows_kernel<DIM, B>, nb, TL_THREADS, smem, st, pl->cnt, static_cast<const int4*>(pl->bucket), \
                    pl->ovf_v, static_cast<const int4*>(pl->ovf_e), N, pl->n_rows, pl->adj, pl->heap, degs,         \
                    pl->counters, ff, pp, pl->hslot, pl->partials, pl->gdone, pl->sync + 3, pl->scalars);           \
  mark("tile_rows", st)
    switch (bar) {
      case 0: DM_TILE_LAUNCH(0); break;
      case 1: DM_TILE_LAUNCH(1); break;
      default: DM_TILE_LAUNCH(-1); break;
    }
#undef DM_TILE_LAUNCH
    return (int)cudaGetLastError();
  }
  constexpr int VPB = AB_THREADS / PCfg<DIM>::G;
  const unsigned nb = nblk(pl->n_rows, VPB);
#define DM_ADJ(B)                                                                                                   \
  launch_chain(adjacency_kernel<DIM, B>, nb + HV_BLOCKS, AB_THREADS, st, pl->cnt, bucket, pl->ovf_v, ovf_e, N,       \
               pl->n_rows, pl->adj, pl->heap, degs, pl->hv, pl->counters, ff, pp, pl->hslot, pl->partials, pl->gdone, \
               pl->sync + 3, pl->scalars);                                                                          \
  mark("adjacency", st)
  switch (bar) {
    case 0: DM_ADJ(0); break;
    case 1: DM_ADJ(1); break;
    default: DM_ADJ(-1); break;
  }
#undef DM_ADJ
  return (int)cudaGetLastError();
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* dm_version(void) { return "distmesh_b200 0.3 (sm_100a)"; }

size_t dm_scan_scratch_bytes(int64_t n) { return scan_scratch_bytes(n); }

int dm_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* scratch, size_t scratch_bytes,
                          void* stream) {
  if (!in || !out || !scratch) return DM_ERR_ARG;
  return exclusive_scan(in, out, n, scratch, scratch_bytes, S(stream));
}

int dm_sdf_eval(const double* prog, const double* x, int64_t M, int dim, double* out, void* stream) {
  if (!prog || M < 0 || bad_dim(dim)) return DM_ERR_ARG;
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
  if (!f || M < 0 || bad_dim(f->dim) || f->kind == DM_SIZE_EXTERNAL) return DM_ERR_ARG;
  if (check_size_fn(f, f->dim)) return DM_ERR_ARG;
  if (M == 0) return DM_OK;  // empty input: pointers may be NULL
  if (!x || !out) return DM_ERR_ARG;
  if (f->dim == 2)
    size_eval_kernel<2><<<nblk(M, 256), 256, 0, S(stream)>>>(*f, x, M, out);
  else
    size_eval_kernel<3><<<nblk(M, 256), 256, 0, S(stream)>>>(*f, x, M, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_size_build_cells(const DmSizeFn* f, double* cells, void* stream) {
  if (!f || !cells || f->kind != DM_SIZE_GRID || f->dim != 3 || check_size_fn(f, 3)) return DM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(cells) & 63) != 0) return DM_ERR_ARG;
  const int64_t nc = (int64_t)(f->n[0] - 1) * (f->n[1] - 1) * (f->n[2] - 1);
  size_cells_kernel<<<nblk(nc, 256), 256, 0, S(stream)>>>(f->grid, f->n[0], f->n[1], f->n[2], cells);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_centroids(const double* p, const int32_t* t, int64_t T, int dim, double* out, void* stream) {
  if (T < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  if (!p || !t || !out) return DM_ERR_ARG;
  if (dim == 2)
    centroid_kernel<2><<<nblk(T, 256), 256, 0, S(stream)>>>(p, t, T, out);
  else
    centroid_kernel<3><<<nblk(T, 256), 256, 0, S(stream)>>>(p, t, T, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_cull_cells(const double* prog, const double* p, const int32_t* t, int64_t T, int dim, double geps,
                  uint8_t* keep, void* stream) {
  if (!prog || T < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  if (!p || !t || !keep) return DM_ERR_ARG;
  if (dim == 2)
    cull_scatter_kernel<2><<<nblk(T, DM_CS_THREADS), DM_CS_THREADS, 0, S(stream)>>>(
        prog, p, t, T, geps, 0, keep, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
  else
    cull_scatter_kernel<3><<<nblk(T, DM_CS_THREADS), DM_CS_THREADS, 0, S(stream)>>>(
        prog, p, t, T, geps, 0, keep, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

size_t dm_compact_scratch_bytes(int64_t T) {
  return align256((size_t)(T + 1) * sizeof(int32_t)) + scan_scratch_bytes(T);
}

int dm_compact_cells(const int32_t* t, const uint8_t* keep, int64_t T, int dim, int32_t* t_out,
                     int32_t* T_out_dev, void* scratch, size_t scratch_bytes, void* stream) {
  if (!T_out_dev || !scratch || T < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if (T > 0 && (!t || !keep || !t_out)) return DM_ERR_ARG;
  if (scratch_bytes < dm_compact_scratch_bytes(T)) return DM_ERR_WORKSPACE;
  int32_t* pos = static_cast<int32_t*>(scratch);
  char* tmp = static_cast<char*>(scratch) + align256((size_t)(T + 1) * sizeof(int32_t));
  cudaStream_t st = S(stream);
  if (T > 0) flags_to_int_kernel<<<nblk(T, 256), 256, 0, st>>>(keep, T, pos);
  int rc = exclusive_scan(pos, pos, T, tmp, scan_scratch_bytes(T), st);
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
  if (T < 0) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  if (!p || !t) return DM_ERR_ARG;
  dihedral_kernel<<<nblk(T, 128), 128, 0, S(stream)>>>(p, t, T, min_dh, max_dh, angles, flags);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_sliver_flags(const double* prog, const double* p, const int32_t* t, int64_t T, double geps, double min_dh,
                    double max_dh, uint8_t* keep, uint8_t* flags, void* stream) {
  if (!prog || T < 0) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  if (!p || !t || !flags) return DM_ERR_ARG;
  // the screen's bounds: cos is decreasing on [0, pi]; bounds outside it switch the screen off on that side
  const double cos_hi = (min_dh > 0.0 && min_dh < 3.14159) ? cos(min_dh) : 2.0;
  const double cos_lo = (max_dh > 0.0 && max_dh < 3.14159) ? cos(max_dh) : -2.0;
  const bool screen = min_dh < max_dh && cos_lo < cos_hi;
  sliver_flags_kernel<<<nblk(T, SF_THREADS), SF_THREADS, 0, S(stream)>>>(prog, p, t, T, geps, min_dh, max_dh, screen ? cos_lo : 2.0,
                                                           screen ? cos_hi : -2.0, keep, flags);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_circumsphere_grad(const double* p, const int32_t* t, const int32_t* ele, int64_t S_, double* grad,
                         void* stream) {
  if (S_ < 0) return DM_ERR_ARG;
  if (S_ == 0) return DM_OK;
  if (!p || !t || !grad) return DM_ERR_ARG;
  circumsphere_grad_kernel<<<nblk(S_, 128), 128, 0, S(stream)>>>(p, t, ele, S_, grad);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_sliver_perturb(double* p, int64_t N, const int32_t* t, const int32_t* ele, int64_t S_, double step_h0,
                      int32_t* winner, double* delta, void* stream) {
  if (S_ < 0 || N < 0) return DM_ERR_ARG;
  if (S_ == 0) return DM_OK;
  if (!p || !t || !ele || !winner || !delta) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  DM_CUDA_TRY(cudaMemsetAsync(winner, 0xff, (size_t)N * sizeof(int32_t), st));  // -1
  sliver_winner_kernel<<<nblk(S_, 128), 128, 0, st>>>(t, ele, S_, winner);
  sliver_delta_kernel<<<nblk(S_, 128), 128, 0, st>>>(p, t, ele, S_, step_h0, winner, delta);
  sliver_apply_kernel<<<nblk(S_, 128), 128, 0, st>>>(p, t, ele, S_, winner, delta);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_cells_lead_interior(const double* key, int32_t* t, int64_t T, double thresh, void* stream) {
  if (T < 0) return DM_ERR_ARG;
  if (T == 0) return DM_OK;
  if (!key || !t) return DM_ERR_ARG;
  cells_lead_interior_kernel<<<nblk(T, 256), 256, 0, S(stream)>>>(key, t, T, thresh);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_level_set_newton(const double* prog, double* p, const int32_t* bid, int64_t nb, int dim, double deps,
                        void* stream) {
  if (!prog || nb < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if (nb == 0) return DM_OK;
  if (!p || !bid) return DM_ERR_ARG;
  if (dim == 2)
    level_set_newton_kernel<2><<<nblk(nb, 128), 128, 0, S(stream)>>>(prog, p, bid, nb, deps);
  else
    level_set_newton_kernel<3><<<nblk(nb, 128), 128, 0, S(stream)>>>(prog, p, bid, nb, deps);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_project_points(const double* prog, double* p, int64_t N, int dim, double deps, double h0, int level_idx,
                      void* stream) {
  if (!prog || N < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if (N == 0) return DM_OK;
  if (!p) return DM_ERR_ARG;
  if (dim == 2)
    project_kernel<2><<<nblk(N, 256), 256, 0, S(stream)>>>(prog, p, N, deps, h0, level_idx);
  else
    project_kernel<3><<<nblk(N, 256), 256, 0, S(stream)>>>(prog, p, N, deps, h0, level_idx);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

// ---------------------------------------------------------------------------------------------
// plan + stages
// ---------------------------------------------------------------------------------------------
size_t dm_plan_bytes(int64_t N, int64_t T, int dim) {
  if (N < 0 || T < 0 || bad_dim(dim)) return 0;
  return plan_layout(nullptr, N, T, dim, nullptr);
}

int dm_plan_init(DmPlan* plan, int64_t N, int64_t T, int dim, void* ws, size_t ws_bytes) {
  if (!plan || !ws || N <= 0 || T < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if ((int64_t)dim * (dim + 1) * T >= (int64_t)INT32_MAX || N >= (int64_t)INT32_MAX) return DM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return DM_ERR_ARG;
  if (ws_bytes < plan_layout(nullptr, N, T, dim, nullptr)) return DM_ERR_WORKSPACE;
  plan_layout(plan, N, T, dim, static_cast<char*>(ws));
  return DM_OK;
}

int dm_plan_set_rows(DmPlan* plan, int64_t n_rows) {
  if (!plan || n_rows < 1 || n_rows > plan->N) return DM_ERR_ARG;
  plan->n_rows = n_rows;
  return DM_OK;
}

int dm_plan_set_layout(DmPlan* plan, int layout) {
  if (!plan || layout < DM_LAYOUT_AUTO || layout > DM_LAYOUT_TILES) return DM_ERR_ARG;
  plan->layout = layout;
  return DM_OK;
}

int dm_stage_prep(const DmPlan* pl, const double* p, void* stream) {
  if (!pl) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  {  // zero [cnt | sync | counters | gdone] and (3-D) refresh the padded point copy, one launch
    const int64_t zq = (int64_t)(pl->zero_bytes / 16);
    const bool pad = pl->dim == 3 && p != nullptr;
    const int64_t n = pad && pl->N > zq ? pl->N : zq;
    // head of the chain: a plain stream-ordered launch (whatever precedes it -- a copy, somebody
    // else's kernel -- completes first); its successors are launched programmatically
    prep_kernel<<<nblk(n, PL_THREADS), PL_THREADS, 0, st>>>(static_cast<int4*>(pl->zero_base), zq, p,
                                                            pad ? pl->p4 : nullptr, pl->N);
    mark("prep(zero+pad)", st);
  }
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_stage_cull_chunk(const DmPlan* pl, const double* prog, const double* p, const int32_t* t_chunk, int64_t cell0,
                        int64_t ncells, double geps, int use_keep, void* stream) {
  if (!pl || cell0 < 0 || ncells < 0 || cell0 + ncells > pl->T || (!t_chunk && ncells > 0)) return DM_ERR_ARG;
  if (use_keep && prog && !p) return DM_ERR_ARG;
  if (ncells == 0) return DM_OK;
  const int mode = !use_keep ? 2 : (prog ? 0 : 1);
  cudaStream_t st = S(stream);
  return pl->dim == 2 ? stage_cull_scatter<2>(pl, prog, p, t_chunk, geps, mode, st, cell0, ncells)
                      : stage_cull_scatter<3>(pl, prog, p, t_chunk, geps, mode, st, cell0, ncells);
}

int dm_stage_cull_count(const DmPlan* pl, const double* prog, const double* p, const int32_t* t, double geps,
                        int use_keep, void* stream) {
  if (!pl || (!t && pl->T > 0) || (use_keep && prog && !p)) return DM_ERR_ARG;
  const int rc = dm_stage_prep(pl, p, stream);
  if (rc) return rc;
  return dm_stage_cull_chunk(pl, prog, p, t, 0, pl->T, geps, use_keep, stream);
}

int dm_stage_build_adjacency(const DmPlan* pl, void* stream) {
  if (!pl) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  return pl->dim == 2 ? stage_adjacency<2>(pl, -1, nullptr, nullptr, st) : stage_adjacency<3>(pl, -1, nullptr, nullptr, st);
}

int dm_stage_bar_index(const DmPlan* pl, void* stream) {
  if (!pl) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  upper_count_kernel<<<nblk(pl->N, 256), 256, 0, st>>>(reinterpret_cast<const int2*>(pl->degs), pl->N, pl->n_rows, pl->rowptr);
  return exclusive_scan(pl->rowptr, pl->rowptr, pl->N, pl->scan_tmp, pl->scan_tmp_bytes, st);
}

int dm_bars_pairs(const DmPlan* pl, int32_t* pairs, void* stream) {
  if (!pl || !pairs) return DM_ERR_ARG;
  if (pl->dim == 2)
    bars_pairs_kernel<2><<<nblk(pl->n_rows, 256), 256, 0, S(stream)>>>(rows_of<2>(pl), pl->rowptr, pl->n_rows, pairs);
  else
    bars_pairs_kernel<3><<<nblk(pl->n_rows, 256), 256, 0, S(stream)>>>(rows_of<3>(pl), pl->rowptr, pl->n_rows, pairs);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_bar_sizes(const DmPlan* pl, const DmSizeFn* f, double* out, void* stream) {
  if (!pl || !f || !out) return DM_ERR_ARG;
  if (pl->dim == 2)
    bar_sizes_kernel<2><<<nblk(pl->n_rows, 256), 256, 0, S(stream)>>>(*f, rows_of<2>(pl), pl->rowptr, pl->hslot, pl->hbar,
                                                                      pl->n_rows, out);
  else
    bar_sizes_kernel<3><<<nblk(pl->n_rows, 256), 256, 0, S(stream)>>>(*f, rows_of<3>(pl), pl->rowptr, pl->hslot, pl->hbar,
                                                                      pl->n_rows, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_bar_midpoints(const DmPlan* pl, const double* p, double* mid, void* stream) {
  if (!pl || !p || !mid) return DM_ERR_ARG;
  DmSizeFn f;
  memset(&f, 0, sizeof(f));
  cudaStream_t st = S(stream);
  return pl->dim == 2 ? launch_bar_pass<2>(pl, p, f, 3, mid, st) : launch_bar_pass<3>(pl, p, f, 3, mid, st);
}

int dm_stage_bar_pass(const DmPlan* pl, const double* p, const DmSizeFn* f, void* stream) {
  if (!pl || !p || check_size_fn(f, pl->dim)) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  const int rc = pl->dim == 2 ? launch_bar_pass<2>(pl, p, *f, f->kind, nullptr, st)
                              : launch_bar_pass<3>(pl, p, *f, f->kind, nullptr, st);
  mark("bar_pass+scale", st);
  return rc;
}

static int vertex_update_impl(const DmPlan* pl, const double* p, bool pad, double* p_out, const double* const* progs,
                              int nlevels, const DmSizeFn* f, double L0mult, double delta_t, double deps, double h0,
                              int64_t nfix, const uint8_t* fixed, double* Ftot, cudaStream_t st) {
  Levels lv;
  memset(&lv, 0, sizeof(lv));
  lv.n = nlevels;
  for (int l = 0; l < nlevels; ++l) {
    if (!progs[l]) return DM_ERR_ARG;
    lv.prog[l] = progs[l];
  }
  const int rc = pl->dim == 2 ? launch_vertex_update<2>(pl, p, false, p_out, lv, *f, f->kind, L0mult, delta_t, deps, h0,
                                                        nfix, fixed, Ftot, st)
                              : launch_vertex_update<3>(pl, p, pad, p_out, lv, *f, f->kind, L0mult, delta_t, deps, h0,
                                                        nfix, fixed, Ftot, st);
  return rc;
}

int dm_stage_vertex_update(const DmPlan* pl, const double* p, double* p_out, const double* const* progs,
                           int nlevels, const DmSizeFn* f, double L0mult, double delta_t, double deps, double h0,
                           int64_t nfix, const uint8_t* fixed, double* Ftot, void* stream) {
  if (!pl || !p || !p_out || p == p_out || nlevels < 0 || nlevels > DM_MAX_LEVELS) return DM_ERR_ARG;
  if ((nlevels > 0 && !progs) || check_size_fn(f, pl->dim)) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  return vertex_update_impl(pl, p, false, p_out, progs, nlevels, f, L0mult, delta_t, deps, h0, nfix, fixed, Ftot,
                            st);
}

int dm_force_iteration(const DmPlan* pl, const double* const* progs, int nlevels, const DmSizeFn* f,
                       const double* p, const int32_t* t, double* p_out, double geps, double L0mult,
                       double delta_t, double deps, double h0, int64_t nfix, const uint8_t* fixed, double* Ftot,
                       void* stream) {
  if (!pl || !progs || nlevels < 1 || nlevels > DM_MAX_LEVELS || !f || f->kind == DM_SIZE_EXTERNAL) return DM_ERR_ARG;
  if (!p || !p_out || p == p_out || check_size_fn(f, pl->dim)) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  (void)st;
  // A: cull + scatter ; B+C: adjacency rows with the bar pass fused in ; D: vertex update
  const int rc = dm_stage_cull_count(pl, progs[0], p, t, geps, 1, stream);
  if (rc) return rc;
  return dm_force_iteration_tail(pl, progs, nlevels, f, p, p_out, L0mult, delta_t, deps, h0, nfix, fixed, Ftot, stream);
}

int dm_force_iteration_tail(const DmPlan* pl, const double* const* progs, int nlevels, const DmSizeFn* f,
                            const double* p, double* p_out, double L0mult, double delta_t, double deps, double h0,
                            int64_t nfix, const uint8_t* fixed, double* Ftot, void* stream) {
  if (!pl || !progs || nlevels < 1 || nlevels > DM_MAX_LEVELS || !f || f->kind == DM_SIZE_EXTERNAL) return DM_ERR_ARG;
  if (!p || !p_out || p == p_out || check_size_fn(f, pl->dim)) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  int rc = pl->dim == 2 ? stage_adjacency<2>(pl, f->kind, f, p, st) : stage_adjacency<3>(pl, f->kind, f, p, st);
  if (rc) return rc;
  return vertex_update_impl(pl, p, true, p_out, progs, nlevels, f, L0mult, delta_t, deps, h0, nfix, fixed, Ftot, st);
}

int dm_force_iteration_reuse(const DmPlan* pl, const double* const* progs, int nlevels, const DmSizeFn* f,
                             const double* p, double* p_out, double L0mult, double delta_t, double deps, double h0,
                             int64_t nfix, const uint8_t* fixed, double* Ftot, void* stream) {
  if (!pl || !progs || nlevels < 1 || nlevels > DM_MAX_LEVELS || !f || f->kind == DM_SIZE_EXTERNAL) return DM_ERR_ARG;
  if (!p || !p_out || p == p_out || check_size_fn(f, pl->dim)) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  // C + D on the rows of the last dm_force_iteration / stage B: the cell list has not changed
  int rc = pl->dim == 2 ? launch_bar_pass<2>(pl, p, *f, f->kind, nullptr, st)
                        : launch_bar_pass<3>(pl, p, *f, f->kind, nullptr, st);
  mark("bar_pass+scale", st);
  if (rc) return rc;
  return vertex_update_impl(pl, p, false, p_out, progs, nlevels, f, L0mult, delta_t, deps, h0, nfix, fixed, Ftot, st);
}

int dm_stage_displacement(const DmPlan* pl, const double* p, const double* p_ref, const DmSizeFn* f, void* stream) {
  if (!pl || !p || !p_ref) return DM_ERR_ARG;
  if (f && (f->kind == DM_SIZE_EXTERNAL || check_size_fn(f, pl->dim))) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  const unsigned nb = nblk(pl->N, PL_THREADS);
  DmSizeFn fz;
  memset(&fz, 0, sizeof(fz));
  const DmSizeFn& ff = f ? *f : fz;
#define DM_DISP(D_, R_) \
  displacement_kernel<D_, R_><<<nb, PL_THREADS, 0, st>>>(ff, p, p_ref, pl->N, pl->partials, pl->sync + 5, pl->scalars)
  if (pl->dim == 2) {
    if (f) DM_DISP(2, 1); else DM_DISP(2, 0);
  } else {
    if (f) DM_DISP(3, 1); else DM_DISP(3, 0);
  }
#undef DM_DISP
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

static size_t lap_layout(LapWork* w, int64_t N, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* ptr = base ? base + off : nullptr;
    off += align256(bytes);
    return ptr;
  };
  const int64_t nb = nblk(N, LS_THREADS);
  char* ntri = take((size_t)N * 4);
  char* done = take(2 * 4);
  char* interior = take((size_t)N);
  char* v[5];
  for (int k = 0; k < 5; ++k) v[k] = take((size_t)N * 16);
  char* partials = take((size_t)nb * 4 * 8);
  char* sc = take(LS_SCALARS * 8);
  if (w) {
    w->ntri = reinterpret_cast<int32_t*>(ntri);
    w->done = reinterpret_cast<int32_t*>(done);
    w->interior = reinterpret_cast<uint8_t*>(interior);
    w->r = reinterpret_cast<double2*>(v[0]);
    w->z = reinterpret_cast<double2*>(v[1]);
    w->p0 = reinterpret_cast<double2*>(v[2]);
    w->p1 = reinterpret_cast<double2*>(v[3]);
    w->Ap = reinterpret_cast<double2*>(v[4]);
    w->partials = reinterpret_cast<double*>(partials);
    w->sc = reinterpret_cast<double*>(sc);
  }
  return off;
}

size_t dm_laplacian_work_bytes(int64_t N) { return N > 0 ? lap_layout(nullptr, N, nullptr) : 0; }

int dm_laplacian_smooth(const DmPlan* pl, const int32_t* t, int64_t T, double* x, void* work, size_t work_bytes, double rtol,
                        int max_iter, int* iters_host, double* resid_host, void* stream) {
  if (!pl || pl->dim != 2 || pl->n_rows != pl->N || T < 0 || (!t && T > 0) || !x || !work || max_iter < 0 || !(rtol >= 0.0))
    return DM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(work) & 255) != 0) return DM_ERR_ARG;
  const int64_t N = pl->N;
  if (work_bytes < lap_layout(nullptr, N, nullptr)) return DM_ERR_WORKSPACE;
  LapWork w;
  lap_layout(&w, N, static_cast<char*>(work));
  cudaStream_t st = S(stream);
  const Rows<2> R = rows_of<2>(pl);
  double2* xs = reinterpret_cast<double2*>(x);
  const unsigned nb = nblk(N, LS_THREADS);
  // [ntri | done] are contiguous in the layout
  DM_CUDA_TRY(cudaMemsetAsync(w.ntri, 0, (size_t)(reinterpret_cast<char*>(w.interior) - reinterpret_cast<char*>(w.ntri)), st));
  if (T > 0) lap_count_kernel<<<nblk(T, 256), 256, 0, st>>>(t, T, w.ntri);
  lap_init_kernel<<<nb, LS_THREADS, 0, st>>>(R, xs, N, w);
  DM_LAUNCH_CHECK();
  const int look = 32;  // iterations between two looks at the residual
  int it = 0;
  double sc[4] = {0.0, 0.0, 0.0, 0.0};  // r.r (2), scale (2)
  bool conv = false;
  while (!conv && it < max_iter) {
    for (int k = 0; k < look && it < max_iter; ++k, ++it) {
      double2* po = (it & 1) ? w.p1 : w.p0;
      double2* pn = (it & 1) ? w.p0 : w.p1;
      lap_ap_kernel<<<nb, LS_THREADS, 0, st>>>(R, N, w, po, pn);
      lap_update_kernel<<<nb, LS_THREADS, 0, st>>>(R, N, w, pn, xs);
    }
    DM_LAUNCH_CHECK();
    DM_CUDA_TRY(cudaMemcpyAsync(sc, w.sc + 6, sizeof(sc), cudaMemcpyDeviceToHost, st));
    DM_CUDA_TRY(cudaStreamSynchronize(st));
    conv = sc[0] <= rtol * rtol * sc[2] && sc[1] <= rtol * rtol * sc[3];
  }
  if (iters_host) *iters_host = it;
  if (resid_host) {  // relative residuals of the two coordinates
    resid_host[0] = sc[2] > 0.0 ? sqrt(sc[0] / sc[2]) : 0.0;
    resid_host[1] = sc[3] > 0.0 ? sqrt(sc[1] / sc[3]) : 0.0;
  }
  return conv || max_iter == 0 ? DM_OK : DM_ERR_WORKSPACE;  // not converged within max_iter
}

int dm_size_from_velocity(const double* vp, const double* h_gr, int64_t n, int dim, double freq, double wl, double hmin,
                          double hmax, double dt, double cr_max, double space_order, double* out, void* stream) {
  if (n < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if (n == 0) return DM_OK;
  if (!vp || !out) return DM_ERR_ARG;
  SizingParams q;
  memset(&q, 0, sizeof(q));
  q.freq_wl = wl > 0.0 ? freq * wl : 0.0;
  q.hmin = hmin;
  q.hmax = hmax;
  q.dimf = (double)dim;
  const bool cfl = !(cr_max == 0.0 || dt == 0.0 || space_order == 0.0);
  q.dt = cfl ? dt : 0.0;
  q.cr_lim = cfl ? cr_max / ((double)dim * space_order) : 0.0;
  q.dim_cr_lim = (double)dim * q.cr_lim;
  sizing_elementwise_kernel<<<nblk(n, 256), 256, 0, S(stream)>>>(vp, h_gr, n, q, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_pad(const double* in, double* out, int dim, const int64_t* shape_host, const int64_t* before_host,
           const int64_t* after_host, int mode, double end_before, double end_after, int32_t* flags_dev, void* stream) {
  if (!in || !out || in == out || bad_dim(dim) || !shape_host || !before_host || !after_host || mode < 0 || mode > 2 || !flags_dev)
    return DM_ERR_ARG;
  PadGeom g;
  int64_t total = 1, inner = 1;
  for (int q = 0; q < 3; ++q) {
    const int64_t m = q < dim ? shape_host[q] : 1, bq = q < dim ? before_host[q] : 0, aq = q < dim ? after_host[q] : 0;
    if (m < 1 || bq < 0 || aq < 0 || m + bq + aq > INT32_MAX) return DM_ERR_ARG;
    g.n[q] = (int)(m + bq + aq);
    g.lo[q] = (int)bq;
    g.hi[q] = (int)(bq + m);
    total *= g.n[q];
    inner *= m;
  }
  (void)total;
  cudaStream_t st = S(stream);
  pad_copy_kernel<<<nblk(inner, 256), 256, 0, st>>>(in, out, g);
  for (int axis = 0; axis < dim; ++axis) {
    const int64_t w = (int64_t)g.lo[axis] + (g.n[axis] - g.hi[axis]);
    if (w == 0) continue;
    int64_t plane = 1;
    for (int q = 0; q < 3; ++q)
      if (q != axis) plane *= q < axis ? g.n[q] : g.hi[q] - g.lo[q];
    if (mode == 2) {
      DM_CUDA_TRY(cudaMemsetAsync(flags_dev, 0, 2 * sizeof(int32_t), st));
      pad_flags_kernel<<<nblk(plane, 256), 256, 0, st>>>(out, g, axis, end_before, end_after, flags_dev);
    }
    pad_fill_kernel<<<nblk(plane * w, 256), 256, 0, st>>>(out, g, axis, mode, end_before, end_after, flags_dev);
  }
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_uniform_filter(const double* in, double* out, double* tmp, int64_t n0, int64_t n1, int64_t n2, const int* size_host,
                      int ndim, int square_input, void* stream) {
  if (!in || !out || !tmp || in == out || in == tmp || out == tmp || !size_host || (ndim != 2 && ndim != 3)) return DM_ERR_ARG;
  if (n0 < 1 || n1 < 1 || n2 < 1 || n0 * n1 * n2 > (int64_t)INT32_MAX * 64 || n0 > INT32_MAX || n1 > INT32_MAX || n2 > INT32_MAX) return DM_ERR_ARG;
  const int64_t n[3] = {n0, n1, n2};
  for (int a = 0; a < ndim; ++a)
    if (size_host[a] < 1 || size_host[a] > n[a]) return DM_ERR_ARG;  // (SciPy's multiple reflections of very short lines are not reproduced)
  cudaStream_t st = S(stream);
  // axis after axis, ping-pong between tmp and out so that the last axis lands in `out`
  const double* src = in;
  for (int a = 0; a < ndim; ++a) {
    double* dst = ((ndim - 1 - a) % 2 == 0) ? out : tmp;
    const int64_t lines = n0 * n1 * n2 / n[a];
    uniform_filter_axis_kernel<<<nblk(lines, 128), 128, 0, st>>>(src, dst, (int)n0, (int)n1, (int)n2, a, size_host[a],
                                                                  a == 0 && square_input ? 1 : 0);
    src = dst;
  }
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_variance_size(const double* mean, const double* sqr_mean, int64_t n, int pass, double vmax, double vmin_scaled, double grad,
                     double* out, void* stream) {
  if (n < 0 || pass < 0 || pass > 1 || !out || (pass == 0 && (!mean || !sqr_mean))) return DM_ERR_ARG;
  if (n == 0) return DM_OK;
  variance_size_kernel<<<nblk(n, 256), 256, 0, S(stream)>>>(mean, sqr_mean, n, pass, vmax, vmin_scaled, grad, out);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_replace_below(double* a, int64_t n, double thresh, double value, int do_replace, unsigned long long* count_dev,
                     void* stream) {
  if (n < 0 || !count_dev) return DM_ERR_ARG;
  if (n == 0) return DM_OK;
  if (!a) return DM_ERR_ARG;
  replace_below_kernel<<<nblk(n, 256), 256, 0, S(stream)>>>(a, n, thresh, value, do_replace, count_dev);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_limgrad(double* f, double* tmp, int64_t n0, int64_t n1, int64_t n2, double delta, double ftol, int max_sweeps,
               int32_t* changed_dev, int* sweeps_host, void* stream) {
  if (!f || !tmp || f == tmp || !changed_dev || n0 < 1 || n1 < 1 || n2 < 1 || max_sweeps < 0 || !(delta >= 0.0)) return DM_ERR_ARG;
  if (n0 > INT32_MAX || n1 > INT32_MAX || n2 > INT32_MAX) return DM_ERR_ARG;
  cudaStream_t st = S(stream);
  const int64_t n = n0 * n1 * n2;
  const int batch = 32;  // sweeps between two looks at the convergence flag (even: the result is back in f)
  int done = 0, flag = 1;
  double *src = f, *dst = tmp;
  while (flag && done < max_sweeps) {
    DM_CUDA_TRY(cudaMemsetAsync(changed_dev, 0, sizeof(int32_t), st));
    for (int b = 0; b < batch && done < max_sweeps; ++b, ++done) {
      limgrad_sweep_kernel<<<nblk(n, 256), 256, 0, st>>>(src, dst, (int)n0, (int)n1, (int)n2, delta, ftol, changed_dev);
      double* sw = src;
      src = dst;
      dst = sw;
    }
    DM_LAUNCH_CHECK();
    DM_CUDA_TRY(cudaMemcpyAsync(&flag, changed_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DM_CUDA_TRY(cudaStreamSynchronize(st));
  }
  if (src != f) DM_CUDA_TRY(cudaMemcpyAsync(f, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (sweeps_host) *sweeps_host = done;
  return flag ? DM_ERR_WORKSPACE : DM_OK;  // not converged within max_sweeps
}

int dm_halo_push(const double* p, const int32_t* idx, int64_t n, int dim, double* dst_peer, void* stream) {
  if (n < 0 || bad_dim(dim)) return DM_ERR_ARG;
  if (n == 0) return DM_OK;
  if (!p || !idx || !dst_peer) return DM_ERR_ARG;
  if (dim == 2)
    halo_push_kernel<2><<<nblk(n, 256), 256, 0, S(stream)>>>(p, idx, n, dst_peer);
  else
    halo_push_kernel<3><<<nblk(n, 256), 256, 0, S(stream)>>>(p, idx, n, dst_peer);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_halo_push2(const double* p, int dim, const int32_t* idx_below, int64_t n_below, double* dst_below,
                  unsigned long long* flag_below, const int32_t* idx_above, int64_t n_above, double* dst_above,
                  unsigned long long* flag_above, unsigned long long stamp, int32_t* done_dev, void* stream) {
  if (!p || bad_dim(dim) || n_below < 0 || n_above < 0 || !done_dev) return DM_ERR_ARG;
  if ((n_below > 0 && (!idx_below || !dst_below)) || (n_above > 0 && (!idx_above || !dst_above))) return DM_ERR_ARG;
  HaloPush h;
  h.idx[0] = idx_below;
  h.idx[1] = idx_above;
  h.n[0] = n_below;
  h.n[1] = n_above;
  h.dst[0] = dst_below;
  h.dst[1] = dst_above;
  h.flag[0] = flag_below;
  h.flag[1] = flag_above;
  const unsigned nb = nblk(n_below + n_above, 256);
  if (dim == 2)
    halo_push2_kernel<2><<<nb, 256, 0, S(stream)>>>(p, h, stamp, done_dev);
  else
    halo_push2_kernel<3><<<nb, 256, 0, S(stream)>>>(p, h, stamp, done_dev);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_halo_wait(const unsigned long long* flag_from_below, const unsigned long long* flag_from_above,
                 unsigned long long stamp, int32_t* err_dev, void* stream) {
  if (!err_dev) return DM_ERR_ARG;
  if (!flag_from_below && !flag_from_above) return DM_OK;
  halo_wait_kernel<<<1, 2, 0, S(stream)>>>(flag_from_below, flag_from_above, stamp, err_dev);
  DM_LAUNCH_CHECK();
  return DM_OK;
}

int dm_halo_select(const double* p, const int32_t* t, int64_t T, int64_t N, int dim, const double* boxes,
                   int has_below, int has_above, uint8_t* flags, void* stream) {
  if (!p || !boxes || !flags || T < 0 || N < 0 || bad_dim(dim) || (!t && T > 0)) return DM_ERR_ARG;
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
  // the flags buffer is padded by the caller to a multiple of 4 bytes
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
