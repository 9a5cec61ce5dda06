/*
 * distmesh_b200.h -- C ABI of libdistmesh_b200.so (hand-written sm_100a CUDA kernels).
 *
 * Drop-in boundary for ONE hot path of krober10nd/SeismicMesh: the DistMesh force-iteration
 * loop inside generate_mesh (SeismicMesh/generation/mesh_generator.py:460-527) and the sibling
 * loop of sliver_removal (:204-286).  Delaunay retriangulation stays on the host.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - coordinates are float64, row-major (N, dim); connectivity is int32, row-major (T, dim+1)
 *     (the reference's native modules use C `double` / `int`, geometry/cpp/fast_geometry.cpp);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no hidden allocation, no global state: all scratch lives in a caller-provided workspace
 *     (query its size with dm_plan_bytes), so one host thread per device can call concurrently;
 *   - return value: 0 = ok, <0 = DM_ERR_* (argument / capacity errors), >0 = cudaError_t.
 *   - all arithmetic is IEEE fp64 without FMA contraction (nvcc -fmad=false) so that stages
 *     whose NumPy counterpart is elementwise are bit-identical to the reference.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef DISTMESH_B200_H
#define DISTMESH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_OK 0
#define DM_ERR_ARG (-1)       /* bad argument (dim, null pointer, negative size) */
#define DM_ERR_WORKSPACE (-2) /* workspace too small */
#define DM_ERR_PROGRAM (-3)   /* malformed SDF program */

#define DM_MAX_LEVELS 8

/* ---------------------------------------------------------------------------------------------
 * SDF program: the reference evaluates a tree of Python objects with one NumPy pass per node
 * (SeismicMesh/geometry/signed_distance_functions.py:283-623).  Here the tree is lowered on the
 * host to a postfix program (float64 words in device memory) that ONE kernel evaluates per point.
 *
 *   prog[0]            = number of instructions n
 *   prog[1 + 24*i ...] = instruction i, 24 float64 words:
 *       [0] opcode   [1] transform flags   [2..7] parameters   [8..10] translation
 *       [11..16] cos,sin of the x / y / z rotation angles   [17..19] unit stretch vector
 *       [20] stretch factor alpha   [21..23] reserved
 * ------------------------------------------------------------------------------------------- */
#define DM_SDF_WORDS 24
enum {
  DM_OP_DISK = 1,      /* params xc,yc,r            (:429-444, :596-598) */
  DM_OP_BALL = 2,      /* params xc,yc,zc,r         (:450-468, :601-603) */
  DM_OP_RECT = 3,      /* params x1,x2,y1,y2        (:474-489; fast_geometry.cpp:165-185) */
  DM_OP_CUBE = 4,      /* params x1,x2,y1,y2,z1,z2  (:495-516; fast_geometry.cpp:104-134) */
  DM_OP_TORUS = 5,     /* params r1,r2              (:522-540) */
  DM_OP_PRISM = 6,     /* params b,h                (:546-563) */
  DM_OP_CYLINDER = 7,  /* params r,h/2              (:569-590) */
  DM_OP_UNION = 16,    /* pops b,a ; pushes min(a,b)                        (:336-339) */
  DM_OP_SUNION = 17,   /* params k ; smooth union                           (:332-334) */
  DM_OP_INTER = 18,    /* max(a,b)                                          (:376-379) */
  DM_OP_SINTER = 19,   /* params k ; smooth intersection                    (:372-374) */
  DM_OP_DIFF = 20,     /* max(a,-b)                                         (:416-420) */
  DM_OP_SDIFF = 21,    /* params k ; max(-a,b)+h^2/4k on the reversed list  (:412-414,421-423) */
  DM_OP_REPEAT_BEGIN = 24, /* params Px,Py,Pz : push point, x <- floormod(x+P/2,P)-P/2 (:292-293) */
  DM_OP_REPEAT_END = 25    /* params bbox6    : pop point, v <- max(v, cube(x))        (:294) */
};
#define DM_TF_TRANSLATE 1
#define DM_TF_ROT0 2 /* 2-D rotation, or rotation about x in 3-D */
#define DM_TF_ROT1 4 /* about y */
#define DM_TF_ROT2 8 /* about z */
#define DM_TF_STRETCH 16

/* ---------------------------------------------------------------------------------------------
 * Mesh-size function fh (SeismicMesh/sizing/size_function.py:1-12; gridded interpolant built at
 * sizing/mesh_size_function.py:391-408 with float32-rounded axes, :514-523).
 * ------------------------------------------------------------------------------------------- */
enum {
  DM_SIZE_CONST = 0,    /* scalar edge_length (mesh_generator.py:574-585) */
  DM_SIZE_GRID = 1,     /* scipy RegularGridInterpolator(linear, extrapolating) semantics */
  DM_SIZE_EXTERNAL = 2  /* h per bar supplied by the caller (opaque Python callable) */
};
typedef struct DmSizeFn {
  int32_t kind;
  int32_t dim;
  int32_t n[3];          /* nodes per axis */
  int32_t _pad;
  const double *axis[3]; /* DEVICE: the ACTUAL axis vectors (float64 of the float32 linspace) */
  const double *grid;    /* DEVICE: (n0,n1[,n2]) float64, C order */
  double hconst;         /* DM_SIZE_CONST */
  const double *cells;   /* DEVICE, optional (NULL = off), 3-D only: the 8 corner values of every grid
                            cell stored contiguously, ((i0*(n1-1)+i1)*(n2-1)+i2)*8 + c0*4+c1*2+c2, built by
                            dm_size_build_cells: a trilinear lookup then reads ONE aligned 64-B record
                            instead of 4 sectors in 4 DRAM pages.  Same values, same arithmetic. */
} DmSizeFn;

/* ---------------------------------------------------------------------------------------------
 * Stand-alone stage kernels (parity-testable one by one)
 * ------------------------------------------------------------------------------------------- */

/* fd(x): replaces `domain.eval(x)` (signed_distance_functions.py, all classes) and the natives
 * drectangle_fast / dblock_fast (fast_geometry.cpp:187,136).  x (M,dim) -> out (M). */
int dm_sdf_eval(const double *prog, const double *x, int64_t M, int dim, double *out, void *stream);

/* fh(x) on a grid: replaces SizeFunction.eval -> RegularGridInterpolator.__call__
 * (size_function.py:11-12).  x (M,dim) -> out (M). */
int dm_size_eval(const DmSizeFn *fh_host, const double *x, int64_t M, double *out, void *stream);

/* fills fh->cells (caller-allocated, (n0-1)*(n1-1)*(n2-1)*8 float64, 64-B aligned) from fh->grid. */
int dm_size_build_cells(const DmSizeFn *fh_host, double *cells, void *stream);

/* centroids p[t].sum(1)/(dim+1) (mesh_generator.py:737) -> out (T,dim); for opaque-callable fd. */
int dm_centroids(const double *p, const int32_t *t, int64_t T, int dim, double *out, void *stream);

/* keep[i] = fd(centroid_i) < -geps   (mesh_generator.py:734-738, _remove_triangles_outside). */
int dm_cull_cells(const double *prog, const double *p, const int32_t *t, int64_t T, int dim,
                  double geps, uint8_t *keep, void *stream);

/* order-preserving compaction t[keep] -> t_out ; *T_out_dev receives the kept count.
 * scratch: (T+1) int32 + dm_scan_scratch_bytes(T+1). */
int dm_compact_cells(const int32_t *t, const uint8_t *keep, int64_t T, int dim, int32_t *t_out,
                     int32_t *T_out_dev, void *scratch, size_t scratch_bytes, void *stream);
size_t dm_compact_scratch_bytes(int64_t T);

/* 6 dihedral angles per tet, cell-major (replaces _fast_geometry.calc_dihedral_angles,
 * fast_geometry.cpp:351-452) and the out-of-bounds test of mesh_generator.py:532-540.
 * angles (6T) may be NULL; flags (T) u8 = 1 if any angle < min_dh or > max_dh. */
int dm_dihedral(const double *p, const int32_t *t, int64_t T, double min_dh, double max_dh,
                double *angles, uint8_t *flags, void *stream);

/* one sliver_removal pass in one kernel (mesh_generator.py:204-243): keep[c] = fd(centroid) < -geps
 * (may be NULL), flags[c] = kept AND some dihedral angle < min_dh or > max_dh.  Cell ids are those of
 * the uncompacted list t (the order of flagged cells equals the reference's after its cull). */
int dm_sliver_flags(const double *prog, const double *p, const int32_t *t, int64_t T, double geps,
                    double min_dh, double max_dh, uint8_t *keep, uint8_t *flags, void *stream);

/* gradient of the circumsphere radius wrt vertex 0 of the listed tets (replaces
 * _fast_geometry.calc_circumsphere_grad, fast_geometry.cpp:580-703).  ele (S) int32 cell ids
 * (NULL = all T cells, S=T) -> grad (S,3). */
int dm_circumsphere_grad(const double *p, const int32_t *t, const int32_t *ele, int64_t S,
                         double *grad, void *stream);

/* sliver perturbation (mesh_generator.py:245-274): p[t[ele,0]] += step*h0*unit(grad), inf->1,
 * last sliver wins for a repeated vertex.  winner (N) int32 and delta (S,3) f64 are scratch.
 * p updated in place (all gradients are taken from the pre-update positions). */
int dm_sliver_perturb(double *p, int64_t N, const int32_t *t, const int32_t *ele, int64_t S,
                      double step_h0, int32_t *winner, double *delta, void *stream);

/* Column order of the tetrahedra handed to the sliver loop.  The reference moves "vertex 0 of every
 * sliver" (mesh_generator.py:234,245-274) and takes whatever vertex order CGAL's cells have
 * (generation/cpp/delaunay_class3.cpp get_finite_cells); here the triangulation stage decides it:
 * key (N) = fd at the vertices; column 0 becomes, among the vertices with key < thresh (well inside
 * the domain), the one picked by (sum of the cell's ids) mod (their count), and the vertex with the
 * smallest key when there is none.  Even permutation of the columns, in place; t (T,4). */
int dm_cells_lead_interior(const double *key, int32_t *t, int64_t T, double thresh, void *stream);

/* 5 damped Newton steps onto the zero level set for the listed vertices (replaces
 * _improve_level_set_newton, mesh_generator.py:741-759).  p updated in place. */
int dm_level_set_newton(const double *prog, double *p, const int32_t *bid, int64_t nb, int dim,
                        double deps, void *stream);


/* ---------------------------------------------------------------------------------------------
 * The force iteration (mesh_generator.py:482, 497-521) as a plan over a caller workspace
 *
 * Per-iteration device structures (all int32 / float64, carved from the workspace).  The layout is
 * built around the observation that every scattered 4..16-byte access costs one 128-byte line
 * wavefront: per-vertex structures are fixed-stride, line-aligned rows.
 *   keep     (T)            cull flags
 *   cnt      (N+1)          kept cells incident to each vertex                       [tiles: records per tile]
 *   bucket   (N*CAP)        per vertex: the OTHER vertex ids of each incident kept cell
 *                           (3-D: CAP=48 entries of 4 ints; 2-D: CAP=16 entries of 2 ints)
 *                           [tiles: per tile of 32 vertices a list of 16-B records {other ids, vertex & 31},
 *                            3-D: 1280 records per tile, 2-D: 320]
 *   ovf_v/e  ((dim+1)*T)    spill records (vertex, entry) of buckets that overflowed  [tiles: (tile, record)]
 *   hv       (N)            vertices whose bucket spilled (hull / hub vertices): listed by stage A, their rows
 *                           are built one block per vertex by the first blocks of stage B's grid
 *   adj      (N*RS)         sorted unique neighbour ids of vertex v at adj[v*RS ...] (RS = 32 ints
 *                           in 3-D = one 128-B line, 16 in 2-D); degs[v] = {deg, nlow}: the first
 *                           nlow are < v.  deg > RS: the row lives in `heap` at offset adj[v*RS].
 *                           The upper parts (>= v) of all rows in vertex order ARE the reference's
 *                           sorted unique (E,2) bars.
 *   rowptr   (N+1)          bar ids = exclusive scan of deg-nlow (built on demand)
 *   hslot    (N*RS + heap)  gridded fh at the midpoint of bar (v,w), stored at its upper slot
 *   hbar     (K/2)          fh per bar id, for DM_SIZE_EXTERNAL
 *   p4       (N*4)          3-D: padded copy of p (32-B rows) refreshed by stage A
 *   esc      (N)            vertices that left a level set in stage D (projected by its second kernel)
 * ------------------------------------------------------------------------------------------- */
typedef struct DmPlan {
  int64_t N, T;      /* vertices, cells handed over by the host Delaunay */
  int32_t dim, _pad0;
  int64_t K;         /* dim*(dim+1)*T : directed (vertex, neighbour) candidates */
  uint8_t *keep;
  void *zero_base;   /* [cnt | sync | counters | gdone]: zeroed by stage A's prep kernel */
  size_t zero_bytes;
  int32_t *cnt;
  int32_t *sync;     /* [1] bar-pass blocks done [2] update blocks done [3] bar-sum arrivals (adjacency) [4] projection blocks done [5] displacement blocks done */
  int32_t *gdone;    /* adjacency blocks done, per reduction group of 128 blocks (zeroed with cnt) */
  int32_t *counters; /* [0]=E unique bars [1]=reserved [2]=spill records [3]=heavy vertices
                        [4]=heap cursor (ints) [5]=escaped vertices (stage D) */
  void *bucket;
  int32_t *ovf_v;
  void *ovf_e;
  int32_t *hv;
  int32_t *adj;
  int32_t *heap;
  int32_t *degs;     /* (N) pairs {deg, nlow} */
  int32_t *rowptr;
  double *hslot;
  double *hbar;
  double *partials;  /* per-block partial sums / maxima */
  double *scalars;   /* [0]=sum L^d [1]=sum h^d [2]=scale [3]=max|F|^2 [4]=maxdp [5]=max displacement (ttol test) */
  double *p4;        /* 3-D: (N,4) padded copy of p made by stage A (32-B rows: one 256-bit gather) */
  int32_t *esc;      /* (N) vertices that left a level set in stage D, projected by its second kernel */
  void *scan_tmp;    /* scratch of the on-demand scans */
  size_t scan_tmp_bytes;
  int64_t n_rows;    /* vertices [0, n_rows) get neighbour rows, bar sums, forces and an update (default N);
                        the others are only NEIGHBOURS: the ghost copies of a slab (dm_plan_set_rows) */
  int64_t layout;    /* DM_LAYOUT_*: how stages A + B pass the kept cells to the vertices (dm_plan_set_layout) */
} DmPlan;

/* Stages A + B exist in two layouts with identical outputs (rows are sorted sets):
 *   DM_LAYOUT_BUCKETS  every kept cell is pushed into a fixed-capacity bucket of each of its vertices; a lane
 *                      group per vertex de-duplicates and sorts its bucket (many short-lived warps: the faster
 *                      one on small meshes and with a 3-D gridded fh);
 *   DM_LAYOUT_TILES    every kept cell leaves one 16-B record per vertex in the list of the vertex's TILE of 32
 *                      consecutive ids (`bucket` then holds the lists, `cnt` their lengths); one warp per tile
 *                      builds the 32 rows (fewer scattered accesses and fewer instructions: the faster one on
 *                      large meshes);
 *   DM_LAYOUT_AUTO     tiles from 200 000 rows on (the default). */
#define DM_LAYOUT_AUTO 0
#define DM_LAYOUT_BUCKETS 1
#define DM_LAYOUT_TILES 2
int dm_plan_set_layout(DmPlan *plan_host, int layout);

size_t dm_plan_bytes(int64_t N, int64_t T, int dim);
/* carve `ws` (device, 256-B aligned, >= dm_plan_bytes) into *plan (host struct). */
int dm_plan_init(DmPlan *plan_host, int64_t N, int64_t T, int dim, void *ws, size_t ws_bytes);

/* Multi-GPU slabs (mesh_generator.py:715-731): with the local vertices ordered [owned | ghosts], only the
 * owned ones need rows / forces / an update -- the ghosts are somebody else's vertices and their new
 * positions arrive through the halo exchange.  n_rows = number of owned vertices (1..N).  The bar sums
 * (force scale) then run over the bars whose smaller LOCAL id is owned, i.e. every bar with an owned end. */
int dm_plan_set_rows(DmPlan *plan_host, int64_t n_rows);

/* stage A: keep flags (fd on centroids) + scatter of every kept cell to its vertices' buckets.
 * prog == NULL: plan->keep was filled by the caller (opaque fd), only count.
 * use_keep == 0: every cell is kept (plain _get_edges(t) semantics, mesh_generator.py:680-688). */
int dm_stage_cull_count(const DmPlan *plan_host, const double *prog, const double *p,
                        const int32_t *t, double geps, int use_keep, void *stream);
/* Stage A in pieces, for callers that stream the cell list in: dm_stage_prep (zero the counters, pad
 * the points) once, then dm_stage_cull_chunk on cells [cell0, cell0+ncells) as they arrive (t_chunk
 * points at the first of them; cells are independent, so chunks may be processed in any order), then
 * dm_force_iteration_tail (stages B-D).  dm_stage_cull_count == prep + one chunk with all cells. */
int dm_stage_prep(const DmPlan *plan_host, const double *p, void *stream);
int dm_stage_cull_chunk(const DmPlan *plan_host, const double *prog, const double *p,
                        const int32_t *t_chunk, int64_t cell0, int64_t ncells, double geps,
                        int use_keep, void *stream);
/* stage B: sorted unique neighbour rows (replaces _fast_geometry.unique_edges,
 * fast_geometry.cpp:30-77, bit-exact: see dm_bars_pairs) from the kept cells stage A scattered. */
int dm_stage_build_adjacency(const DmPlan *plan_host, void *stream);
/* bar ids (rowptr) for dm_bars_pairs / dm_bar_midpoints / DM_SIZE_EXTERNAL. */
int dm_stage_bar_index(const DmPlan *plan_host, void *stream);
/* (E,2) int32 pairs in the reference's order (needs dm_stage_bar_index). */
int dm_bars_pairs(const DmPlan *plan_host, int32_t *pairs, void *stream);
/* bar midpoints (E,dim) for an opaque fh (mesh_generator.py:699; needs dm_stage_bar_index). */
int dm_bar_midpoints(const DmPlan *plan_host, const double *p, double *mid, void *stream);
/* h of every unique bar of the last bar pass, in bar order (E) -- diagnostics / parity tests
 * (the reference's `hedges`, mesh_generator.py:699; needs dm_stage_bar_index). */
int dm_bar_sizes(const DmPlan *plan_host, const DmSizeFn *fh_host, double *out, void *stream);
/* stage C: h at bar midpoints + the global scale ((sum L^d)/(sum h^d))^(1/d)
 * (mesh_generator.py:696-700).  DM_SIZE_EXTERNAL: plan->hbar already holds h per bar id. */
int dm_stage_bar_pass(const DmPlan *plan_host, const double *p, const DmSizeFn *fh_host,
                      void *stream);
/* stage D: forces gathered per vertex in the reference's accumulation order, pfix mask, update,
 * Newton projection per level, max|F| (mesh_generator.py:701-712, 499-521).
 * progs_host: nlevels device program pointers (0 levels = no projection, opaque fd);
 * nfix: the first nfix vertices are fixed; fixed (N) u8 optional extra mask (may be NULL);
 * Ftot (N,dim) optional output (may be NULL). p_out must not alias p. */
int dm_stage_vertex_update(const DmPlan *plan_host, const double *p, double *p_out,
                           const double *const *progs_host, int nlevels, const DmSizeFn *fh_host,
                           double L0mult, double delta_t, double deps, double h0, int64_t nfix,
                           const uint8_t *fixed, double *Ftot, void *stream);
/* projection only (for opaque fd the host does it; for tests): p in place, level idx semantics
 * of _project_points_back_newton (mesh_generator.py:762-784). */
int dm_project_points(const double *prog, double *p, int64_t N, int dim, double deps, double h0,
                      int level_idx, void *stream);

/* A+B+C+D in one call: one DistMesh force iteration on (p, t). */
int dm_force_iteration(const DmPlan *plan_host, const double *const *progs_host, int nlevels,
                       const DmSizeFn *fh_host, const double *p, const int32_t *t, double *p_out,
                       double geps, double L0mult, double delta_t, double deps, double h0,
                       int64_t nfix, const uint8_t *fixed, double *Ftot, void *stream);

/* B+C+D after stage A has been run (in one piece or in chunks). */
int dm_force_iteration_tail(const DmPlan *plan_host, const double *const *progs_host, int nlevels,
                            const DmSizeFn *fh_host, const double *p, double *p_out, double L0mult,
                            double delta_t, double deps, double h0, int64_t nfix,
                            const uint8_t *fixed, double *Ftot, void *stream);

/* C+D only, on the rows left by the last dm_force_iteration (or stage B) of this plan: one force
 * iteration WITHOUT retriangulation (the cell list is unchanged, only p moved).  This is the step
 * DistMesh takes while the `ttol` displacement test does not fire (Persson & Strang; north_star
 * item 5); the reference itself retriangulates every iteration (mesh_generator.py:460-470), so the
 * host driver uses it only when the caller opts in (generate_mesh(..., ttol=...)). */
int dm_force_iteration_reuse(const DmPlan *plan_host, const double *const *progs_host, int nlevels,
                             const DmSizeFn *fh_host, const double *p, double *p_out, double L0mult,
                             double delta_t, double deps, double h0, int64_t nfix,
                             const uint8_t *fixed, double *Ftot, void *stream);
/* the `ttol` test: plan->scalars[5] = max_v |p[v] - p_ref[v]|_2 / h_v (p_ref = positions at the last
 * retriangulation).  fh_host == NULL: h_v = 1 (absolute displacement, compare with ttol * h0 as DistMesh
 * does); otherwise h_v = fh(p[v]) (constant or gridded; DM_SIZE_EXTERNAL is not accepted): displacement
 * in units of the LOCAL mesh size, which is what makes the test usable on graded meshes. */
int dm_stage_displacement(const DmPlan *plan_host, const double *p, const double *p_ref,
                          const DmSizeFn *fh_host, void *stream);

/* dm_force_iteration with a CUDA event recorded after every kernel (measurement only, used by
 * bench.py for the per-kernel roofline table; synchronises the stream).  ms_host[i] and the
 * NUL-terminated names_host[i*names_stride ...] describe the i-th timed span; *n_host = count. */
int dm_force_iteration_profiled(const DmPlan *plan_host, const double *const *progs_host, int nlevels,
                                const DmSizeFn *fh_host, const double *p, const int32_t *t,
                                double *p_out, double geps, double L0mult, double delta_t,
                                double deps, double h0, int64_t nfix, const uint8_t *fixed,
                                void *stream, float *ms_host, char *names_host, int names_stride,
                                int cap, int *n_host);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU slab halo (replaces migration/cpp/cpputils.cpp where_to2/3 :85-200,:247-383 and
 * migration.enqueue, migration.py:116-145): flag vertices whose incident-cell circumball touches
 * the slab box below / above.  boxes_host: 2 boxes x (2*dim) doubles (min..., max...);
 * flags (N) u8: bit0 = export below, bit1 = export above.
 * ------------------------------------------------------------------------------------------- */
int dm_halo_select(const double *p, const int32_t *t, int64_t T, int64_t N, int dim,
                   const double *boxes_host, int has_below, int has_above, uint8_t *flags,
                   void *stream);

/* ghost push over peer memory: dst_peer[i] = p[idx[i]] for i < n, where dst_peer points into a
 * NEIGHBOUR GPU's ghost buffer (a peer-mapped / symmetric-memory address; NVLink stores).  One kernel
 * instead of pack + ncclSend/ncclRecv (+ unpack) for migration.exchange (migration.py:148-183); the
 * caller orders it with the peer through its own signal (parallel.PeerHalo). */
int dm_halo_push(const double *p, const int32_t *idx, int64_t n, int dim, double *dst_peer,
                 void *stream);

/* The whole per-iteration ghost exchange (migration.exchange, migration.py:148-183) as TWO launches.
 * dm_halo_push2: the rows of p listed in idx_below / idx_above go straight into the GHOST ROWS of the
 * neighbours' position buffers (dst_*: peer-mapped addresses of the first ghost row this rank fills in the
 * neighbour's buffer), then `stamp` is stored, with system-scope release, to flag_below / flag_above (peer
 * addresses of the neighbour's arrival stamps; NULL = no neighbour on that side).  done_dev: one zeroed int32
 * of LOCAL device scratch.  dm_halo_wait: holds the stream until this rank's own stamps (written by the
 * neighbours) have reached `stamp`; stamps must grow from step to step.  A neighbour that does not arrive
 * within ~2 s of SM clock sets *err_dev = 1 instead of hanging the device. */
int dm_halo_push2(const double *p, int dim, const int32_t *idx_below, int64_t n_below, double *dst_below,
                  unsigned long long *flag_below, const int32_t *idx_above, int64_t n_above,
                  double *dst_above, unsigned long long *flag_above, unsigned long long stamp,
                  int32_t *done_dev, void *stream);
int dm_halo_wait(const unsigned long long *flag_from_below, const unsigned long long *flag_from_above,
                 unsigned long long stamp, int32_t *err_dev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Sizing preprocessing, elementwise chain in one pass (replaces the NumPy expressions of
 * get_sizing_function_from_segy, sizing/mesh_size_function.py:411-426 wavelength sizing, :180-181 clamp,
 * :453-468 CFL bound): out[i] = clamp(min(vp[i] / (freq*wl), h_gr[i]), hmin, hmax), then the CFL bound
 * (vp*dt)/(dim*cr_lim) where (vp*dt)/(dim*out) > cr_lim = cr_max/(dim*space_order).  wl <= 0: no wavelength
 * term; h_gr NULL: no gradient term; dt, cr_max or space_order == 0: no CFL bound.  Bit-identical to NumPy.
 * ------------------------------------------------------------------------------------------- */
int dm_size_from_velocity(const double *vp, const double *h_gr, int64_t n, int dim, double freq, double wl,
                          double hmin, double hmax, double dt, double cr_max, double space_order,
                          double *out, void *stream);

/* Sizing preprocessing: a[i] < thresh -> value in place (do_replace != 0) and *count_dev += the number of such
 * entries (one zeroed uint64 of device memory): the water layer of a shear-velocity model
 * (sizing/mesh_size_function.py:148-159: vp[vp < 1e-3] = vp_water), without a pass over the model on the host. */
int dm_replace_below(double *a, int64_t n, double thresh, double value, int do_replace,
                     unsigned long long *count_dev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Sizing preprocessing, the `grad=` option: windowed variance of the velocity model (replaces the two
 * scipy.ndimage.uniform_filter calls and the NumPy expressions of sizing/mesh_size_function.py:428-448).
 * dm_uniform_filter: out = uniform_filter(in, size) with SciPy's defaults (mode "reflect", origin 0), axis after
 * axis with SciPy's running-sum recurrence per line, bit-identical; in (n0,n1,n2) C order (n2 = 1 and ndim = 2
 * for a 2-D grid), size_host[ndim] window lengths (<= the axis lengths), tmp: n0*n1*n2 float64 of scratch;
 * square_input != 0: the filter of in*in (the reference's uniform_filter(vp**2)) without a squared copy.
 * dm_variance_size: pass 0: out = sqr_mean - mean*mean ; pass 1 (in place on out, with vmax = max(out) and
 * vmin_scaled = min(out)/vmax taken by the caller): out = grad / ((out / vmax - vmin_scaled) + 0.10).
 * ------------------------------------------------------------------------------------------- */
int dm_uniform_filter(const double *in, double *out, double *tmp, int64_t n0, int64_t n1, int64_t n2,
                      const int *size_host, int ndim, int square_input, void *stream);
int dm_variance_size(const double *mean, const double *sqr_mean, int64_t n, int pass, double vmax,
                     double vmin_scaled, double grad, double *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Sizing preprocessing: domain extension of the gridded size function (replaces np.pad as called by
 * get_sizing_function_from_segy, sizing/mesh_size_function.py:526-587) with NumPy's semantics and arithmetic:
 * in (shape) -> out (shape + before + after), C order, dim 2 or 3; axes padded one after the other, axis 0 first,
 * each from the edge planes of NumPy's region of interest.  mode 0 = "edge", 1 = "constant" (end_before /
 * end_after are the constant values), 2 = "linear_ramp" (end values; np.linspace(end, edge, width,
 * endpoint=False) including its any-zero-step rule).  Bit-identical to np.pad.  flags_dev: two int32 of scratch.
 * ------------------------------------------------------------------------------------------- */
int dm_pad(const double *in, double *out, int dim, const int64_t *shape_host, const int64_t *before_host,
           const int64_t *after_host, int mode, double end_before, double end_after, int32_t *flags_dev,
           void *stream);

/* ---------------------------------------------------------------------------------------------
 * Sizing preprocessing: gradient limiting of a gridded size function in place (replaces
 * _FastHJ.limgrad, sizing/cpp/FastHJ.cpp:63-190, called from _enforce_gradation_sizing,
 * sizing/mesh_size_function.py:471-496).  f (n0,n1,n2) float64 C order (n2 = 1 in 2-D);
 * delta = elen*dfdx ; ftol = min(f)*sqrt(1e-9) as in the reference.  Relaxes to the reference's
 * fixed point (unique up to ftol) with Jacobi sweeps between f and tmp (n0*n1*n2 float64 of device scratch),
 * so the result is deterministic; it ends up in f.  changed_dev: one int32 of device scratch.  Synchronises
 * the stream every 32 sweeps; *sweeps_host = sweeps run.  DM_ERR_WORKSPACE if max_sweeps did not suffice.
 * ------------------------------------------------------------------------------------------- */
int dm_limgrad(double *f, double *tmp, int64_t n0, int64_t n1, int64_t n2, double delta, double ftol,
               int max_sweeps, int32_t *changed_dev, int *sweeps_host, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Termination path: Laplacian smoothing of a 2-D mesh as one linear solve (replaces
 * geometry.laplacian2_fixed_point, geometry/utils.py:494-547: SciPy assembly + pyamg Ruge-Stuben solve).
 * Every interior vertex goes to the average of its neighbours; the boundary vertices -- those with more
 * neighbours than incident triangles, i.e. the vertices of edges that belong to one triangle only
 * (get_boundary_vertices, geometry/utils.py:399-417) -- stay where they are.  Needs the neighbour rows of
 * stage B for the cells t (dm_stage_cull_count(..., use_keep = 0) + dm_stage_build_adjacency on this plan).
 * Matrix-free Jacobi-preconditioned conjugate gradients on the interior block, both coordinates at once, two
 * launches per iteration, scalars on the device; the host looks at the residual every 32 iterations and
 * stops when |r| <= rtol * |diag(A) x| for both coordinates.  x (N,2): in = the vertices, out = the solution.
 * work: dm_laplacian_work_bytes(N) of device scratch, 256-B aligned.  DM_ERR_WORKSPACE if max_iter did not
 * suffice.  *iters_host = iterations run, resid_host[2] = the relative residuals reached.
 * ------------------------------------------------------------------------------------------- */
size_t dm_laplacian_work_bytes(int64_t N);
int dm_laplacian_smooth(const DmPlan *plan_host, const int32_t *t, int64_t T, double *x, void *work,
                        size_t work_bytes, double rtol, int max_iter, int *iters_host, double *resid_host,
                        void *stream);

/* utilities */
size_t dm_scan_scratch_bytes(int64_t n);
/* exclusive scan of int32 in[0..n) -> out[0..n], out[n] = total (in == out allowed) */
int dm_exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, void *scratch,
                          size_t scratch_bytes, void *stream);
const char *dm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DISTMESH_B200_H */
