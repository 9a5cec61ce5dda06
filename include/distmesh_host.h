/*
 * distmesh_host.h -- C ABI of libdistmesh_host.so: the HOST side of the retriangulation step (2-D and 3-D).
 *
 * BASELINE.json north_star keeps the Delaunay retriangulation of every DistMesh iteration on the
 * host.  The reference does it with CGAL behind two pybind11 classes
 * (SeismicMesh/generation/cpp/delaunay_class.cpp:33-117, delaunay_class3.cpp): the Python loop
 * hands the coordinates over as a Python list (`dt.insert(p.ravel().tolist())`,
 * mesh_generator.py:466), CGAL spatially re-sorts them, and `get_finite_vertices` /
 * `get_finite_cells` copy a RENUMBERED vertex array and the cells back (:480-481).
 *
 * This library replaces that with a stateless triangulator on raw buffers that keeps the caller's
 * vertex numbering, so device-resident coordinates never have to be
 * permuted: `points` can be the pinned staging buffer the device loop downloads into, `cells`
 * the pinned buffer it uploads from.
 *
 *   - every pointer is a HOST pointer; float64 row-major coordinates, int32 row-major cells;
 *   - no global state, re-entrant; the only allocation is internal scratch freed before return;
 *   - Output order: cells GROUPED BY THEIR SMALLEST VERTEX ID (groups ascending, cells of a group in
 *     lexicographic order of their sorted ids); the column order inside a cell is the triangulator's own and
 *     is NOT canonicalised.  The grouping is what the device pipeline's warp-aggregated slot claims and
 *     position gathers like best (csrc/host/dm_cell_order.h); an unbiased column 0 is what the reference's
 *     sliver loop needs ("vertex 0 of every sliver", mesh_generator.py:234,245-274).  Orientation is NOT
 *     normalised: the loop body never uses it, and the reference fixes it at termination (fix_mesh);
 *   - results are exact Delaunay triangulations: the orientation / in-circle predicates run a
 *     floating-point filter and fall back to exact expansion arithmetic, so co-circular and
 *     collinear inputs (the initial lattice, boundary vertices projected onto straight edges) are
 *     handled without tolerances; for points in general position the cell SET equals the one any
 *     other correct Delaunay code (CGAL, Qhull) returns.
 */
#ifndef DISTMESH_HOST_H
#define DISTMESH_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMH_OK 0
#define DMH_ERR_ARG (-1)      /* null pointer / negative size / a coordinate that is not finite */
#define DMH_ERR_CAPACITY (-2) /* `cells` too small (see dmh_delaunay2d_max_cells) */

/* "distmesh_host <version>" */
const char* dmh_version(void);

/* Upper bound of the number of triangles of a 2-D Delaunay triangulation of N points (2N - 5,
 * at least 1): the capacity to give `cells`. */
int64_t dmh_delaunay2d_max_cells(int64_t N);

/* Delaunay triangulation of `points` (N,2).  Replaces DelaunayTriangulation.insert +
 * get_finite_cells (generation/cpp/delaunay_class.cpp:45-62, 99-117) with the vertex ids = input
 * rows.  Writes *T_out triangles to `cells` (capacity `cap` rows) in the order described under "Output
 * order" above.  Input rows
 * that are in no triangle are counted in *duplicates_out (exact duplicates of an earlier row, which
 * keeps the cells; CGAL and Qhull leave those out as well) and *lost_out (everything else: all rows
 * when the input is collinear or has fewer than 3 distinct points, in which case *T_out = 0; a row
 * whose insertion-order tie was lost to rounding -- callers should retriangulate such an input with
 * another code).  Either pointer may be NULL.
 * Sweep-hull construction (points inserted by distance from a seed circumcentre, advancing convex
 * front, Lawson flips), O(N log N). */
int dmh_delaunay2d(const double* points, int64_t N, int32_t* cells, int64_t cap, int64_t* T_out,
                   int64_t* duplicates_out, int64_t* lost_out);

/* Capacity to give `cells` of dmh_delaunay3d for N points: 8 N + 64 tetrahedra, generous for mesh-like
 * point sets (about 6.7 N); a pathological input can need more, see DMH_ERR_CAPACITY below. */
int64_t dmh_delaunay3d_max_cells(int64_t N);

/* Delaunay triangulation of `points` (N,3).  Replaces DelaunayTriangulation3.insert +
 * get_finite_cells (generation/cpp/delaunay_class3.cpp) with the vertex ids = input rows.  Writes
 * *T_out tetrahedra to `cells` (see "Output order" above); with
 * DMH_ERR_CAPACITY nothing is written and *T_out is the capacity that is needed.  *duplicates_out
 * counts rows left out as exact duplicates of another row (ONE copy is in the cells, not necessarily
 * the first), *lost_out rows left out for any other reason: all N when there are not four affinely
 * independent points (then *T_out = 0), non-zero also when an internal consistency check failed --
 * callers should retriangulate such an input with another code.  Incremental Bowyer-Watson in a
 * biased randomised insertion order (Morton curve within a round), ghost tetrahedra on the hull. */
int dmh_delaunay3d(const double* points, int64_t N, int32_t* cells, int64_t cap, int64_t* T_out,
                   int64_t* duplicates_out, int64_t* lost_out);

/* The same triangulation built by `threads` host threads (<= 0: DM_HOST_THREADS if set, else the cores
 * this process may run on, at most 16; dmh_delaunay3d is this call with threads = 1).  The reference
 * has no counterpart: CGAL's insert behind delaunay_class3.cpp is single-threaded, and the 3-D
 * retriangulation is where its time-to-mesh goes (docs/source/usrman/overview.rst:116).
 * The points of a large round of the insertion order are divided among the boxes of a k-d tree, one
 * thread each; a thread inserts a point only while the walk to it
 * and its cavity consist of tetrahedra with all vertices in the thread's own box, and otherwise hands
 * the point to the next pass (other cutting planes) and finally to the serial code.  No locks: no two
 * threads ever modify or delete the same tetrahedron; tetrahedra with vertices of several boxes are
 * frozen during a pass (only their neighbour slots facing a cavity are updated, each by one thread).
 * For points in general position the cell set is the one dmh_delaunay3d returns; the output order is
 * the same canonical order. */
int dmh_delaunay3d_mt(const double* points, int64_t N, int32_t* cells, int64_t cap, int64_t* T_out,
                      int64_t* duplicates_out, int64_t* lost_out, int threads);

/* A 3-D triangulation that stays around between calls.  dmh_dt3_build triangulates `points` (N,3) with
 * `threads` threads (as dmh_delaunay3d_mt) and returns a handle (NULL on error, *rc_out = the code);
 * dmh_dt3_cells writes its cells (conventions of dmh_delaunay3d; ids = rows in the order the points were
 * given, first the N of the build, then every inserted batch); dmh_dt3_insert adds `more` (M,3) to the SAME
 * triangulation by serial incremental insertion; dmh_dt3_points = rows so far; dmh_dt3_free releases it.
 * This is the reference's use of its CGAL object in the slab-parallel loop: the owned vertices are
 * triangulated, the cells decide which vertices a neighbour needs, the ghost vertices received in exchange
 * are INSERTED (`dt.insert`, mesh_generator.py:466 with :715-731) and the cells read again -- one
 * construction and a short insertion per iteration instead of two constructions.  A handle whose start was
 * degenerate (fewer than four affinely independent points) is rebuilt from all points at the next insert. */
void* dmh_dt3_build(const double* points, int64_t N, int threads, int* rc_out);
int dmh_dt3_insert(void* handle, const double* more, int64_t M);
int64_t dmh_dt3_points(void* handle);
int dmh_dt3_cells(void* handle, int32_t* cells, int64_t cap, int64_t* T_out, int64_t* duplicates_out, int64_t* lost_out);
void dmh_dt3_free(void* handle);

/* Sorted unique rows of an int32 table `rows` (n, k), k = 2, 3 or 4, ids in [0, N): the ids of every row
 * are sorted ascending, the rows put in lexicographic order and equal rows collapsed, in place; the
 * *n_unique rows left are at the front, `counts` (n entries, may be NULL) holds how often each occurred.
 * Replaces geometry.unique_rows of the row-wise sorted list (SeismicMesh/geometry/utils.py:141-172) where
 * the reference's termination path and its boundary queries use it on cells, facets and edges (fix_mesh
 * :204-246, get_boundary_edges / get_boundary_facets :310-361: a facet that occurs once is on the boundary).
 * `threads` <= 0: 8. */
int dmh_sort_unique_rows_i32(int32_t* rows, int64_t n, int k, int64_t N, int32_t* counts, int64_t* n_unique, int threads);

/* The exact predicates, exported for the tests.
 * orient3d > 0: (a, b, c, d) positively oriented; insphere > 0: e strictly inside the sphere through a positively oriented (a, b, c, d). */
double dmh_orient3d(const double* a, const double* b, const double* c, const double* d);
double dmh_insphere(const double* a, const double* b, const double* c, const double* d, const double* e);

/* The two 2-D predicates.
 * orient2d > 0: a, b, c counter-clockwise; incircle > 0: d strictly inside the circle through the
 * counter-clockwise a, b, c.  Only the SIGN is meaningful (exact, including 0). */
double dmh_orient2d(const double* a, const double* b, const double* c);
double dmh_incircle(const double* a, const double* b, const double* c, const double* d);

#ifdef __cplusplus
}
#endif
#endif
