// The force-iteration pipeline (stages A-D of include/distmesh_b200.h), third layout ("buckets"; stages A + B
// also exist in a fourth, "tiles": dm_tiles.cuh, the choice is a plan setting).
//
// What bounds these kernels on B200 is not DRAM bytes.  Every scattered 4..16-B access (gather, store or
// atomic) costs about one L1 wavefront however few bytes it moves, so the layout makes every per-vertex
// structure one aligned 128-B line (or a few), read / written by a lane group in one wavefront; beyond that,
// round 2 found stage A bound by the latency of its dependent chain (cell ids -> gathers -> SDF -> claims ->
// stores), stage B by instruction issue and shared-memory wavefronts, stage D by its gathers and the fp64
// square roots / divisions:
//
//   0  prep              zero the per-iteration counters; 3-D: padded point copy p4 (one 256-bit gather per point)
//   A  cull_scatter      one thread per cell: centroid + fused SDF program -> keep flag; each kept
//                        cell appends its OTHER vertex ids to the fixed-capacity bucket of each of
//                        its vertices.  Slots are claimed with warp-aggregated atomics (one per run
//                        of equal ids in consecutive lanes: host Delaunay codes emit cells grouped
//                        around vertices); a full bucket spills to a global overflow list and lists
//                        its vertex as heavy.
//   B  adjacency         one lane group (8 lanes in 3-D, 4 in 2-D) per vertex: coalesced read of
//                        the bucket, de-duplication in a group-private shared-memory hash (plain
//                        LDS/STS, write-then-verify, no shared atomics), rank sort, ONE 128-B
//                        store of the sorted neighbour row.  Row = [lower neighbours ascending |
//                        upper neighbours ascending]; the upper parts of all rows, in vertex
//                        order, ARE the reference's sorted unique bar list (unique_edges,
//                        geometry/cpp/fast_geometry.cpp:30-77).  Fused in: the bar pass over the
//                        row's upper bars (L, fh(midpoint), sum L^d, sum h^d) and the fixed-order
//                        reduction to scale = ((sum L^d)/(sum h^d))^(1/d).  The first HV_BLOCKS
//                        blocks of the grid build the rows of the heavy vertices (bucket spilled:
//                        hull / hub vertices), one block per vertex, beside the main blocks.
//   C  bar_pass          stand-alone bar pass of the staged path (opaque fh)
//   D  vertex_update     per-vertex gather of bar forces in the reference's COO accumulation order,
//                        pfix, p += dt*F, max|F| (last block reduces); vertices that left a level
//                        set are listed and get their Newton projection from project_list_kernel
#pragma once
#include "dm_device.cuh"

namespace dm {

constexpr int PL_THREADS = 256;  // cull / bar pass / vertex update
constexpr int AB_THREADS = 256;  // adjacency: 256 / G vertices per block
constexpr int HV_BLOCKS = 148;   // heavy-vertex blocks at the head of the adjacency grid; they loop over the heavy list
constexpr int HV_SMEM = 4608;    // heavy path: candidates sorted in shared memory up to this many ints (the main
                                 // path's tables + lists: 2 * 32 * 72 in 3-D, 2 * 64 * 36 in 2-D)
// tuning knobs (resident blocks per SM the register allocation is held to)
#ifndef DM_CS_MINB
#define DM_CS_MINB 8
#endif
#ifndef DM_CS_THREADS
#define DM_CS_THREADS 128
#endif
#ifndef DM_VU_THREADS
#define DM_VU_THREADS 128
#endif
#ifndef DM_HASH_MAXSTEPS
#define DM_HASH_MAXSTEPS 12
#endif
#ifndef DM_AB_MINB
#define DM_AB_MINB 5
#endif
#ifndef DM_VU_MINB
#define DM_VU_MINB 8
#endif
// The Newton projection of the escaped vertices inside vertex_update (no list, no project_list launch) pays on
// the small configurations, where a launch is a fifth of the step (disk h0=0.01 39.0 -> 36.8 us, EAGE-shaped hmin
// 150 117 -> 112 us, BP2004-shaped hmin 75 56.3 -> 53.3 us), and costs on the large ones, where the escaped lanes
// hold their warps (ball h0=0.02: vertex_update 75 -> 87 us for a 13 us launch): chosen by the number of rows.
#ifndef DM_FUSE_PROJECT_BELOW
#define DM_FUSE_PROJECT_BELOW 200000
#endif
#ifndef DM_AB_BARRIER
#define DM_AB_BARRIER 0  // 1: block totals of stage B through a block barrier instead of an arrival counter
#endif
// Gridded fh: the bar pass leaves h at EVERY slot of a row, lower neighbours included, so that the vertex
// update reads h of all its bars from its own row and never interpolates (its per-vertex loop is latency
// bound; the bar pass spreads the interpolations of a row over a lane group).  Every bar is therefore
// interpolated twice, once per end, with the same midpoint bits and hence the same h.  Measured in round 2
// and rejected: evaluating h once per bar and handing it to the other end through a search of the owner's
// row -- inside the vertex update (+9 % on that kernel), in a split lower / upper loop (+7..35 %), or in a
// mirror kernel of its own after the row build (0.130 ms against 0.087 ms for the second interpolation on
// the EAGE-shaped workload): five dependent gathers cost more than three divisions and a 64-B record.
constexpr bool H_ALL_SLOTS = true;

template <int DIM>
struct PCfg;
template <>
struct PCfg<3> {
  static constexpr int CAP = 48;  // bucket capacity: incident cells per vertex (mean 24, p99.9 42)
  static constexpr int RS = 32;   // fixed adjacency row: 32 ints = one 128-B line (mean degree 14)
  static constexpr int G = 8;     // lanes per vertex
  static constexpr int LOGH = 6;  // group hash slots
  typedef int4 entry_t;           // the 3 other vertices of an incident cell (+ pad)
};
template <>
struct PCfg<2> {
  static constexpr int CAP = 16;
  static constexpr int RS = 16;
  static constexpr int G = 4;
  static constexpr int LOGH = 5;
  typedef int2 entry_t;  // the 2 other vertices of an incident triangle
};

__host__ __device__ __forceinline__ int64_t round4(int64_t x) { return (x + 3) & ~(int64_t)3; }

// adjacency rows: deg <= RS -> fixed row at adj[v*RS]; else in the heap at offset adj[v*RS]
template <int DIM>
struct Rows {
  const int32_t* __restrict__ adj;
  const int32_t* __restrict__ heap;
  const int2* __restrict__ degs;  // {deg, nlow}
  int64_t N;
  __device__ __forceinline__ const int32_t* row(int64_t v, int deg) const {
    const int32_t* r = adj + v * PCfg<DIM>::RS;
    return deg <= PCfg<DIM>::RS ? r : heap + (uint32_t)r[0];
  }
  // index of the row's first slot in the per-slot arrays (hslot)
  __device__ __forceinline__ int64_t slot_base(int64_t v, int deg) const {
    return deg <= PCfg<DIM>::RS ? v * PCfg<DIM>::RS : N * PCfg<DIM>::RS + (uint32_t)adj[v * PCfg<DIM>::RS];
  }
};

// ---------------------------------------------------------------------------------------------
// prep: zero the per-iteration counters and (3-D) make the padded point copy p4 (N,4)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_THREADS) prep_kernel(int4* __restrict__ zero, int64_t zquads,
                                                         const double* __restrict__ p, double* __restrict__ p4,
                                                         int64_t N) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < zquads) zero[i] = make_int4(0, 0, 0, 0);
  if (p4 != nullptr && i < N) {
    const double* q = p + 3 * i;
    stg256(p4 + 4 * i, q[0], q[1], q[2], 0.0);
  }
}

// ---------------------------------------------------------------------------------------------
// A: cull + scatter.  mode 0: evaluate fd on the centroid, write keep ; 1: keep given ; 2: all kept
// ---------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ typename PCfg<DIM>::entry_t others_of(const int (&ids)[4], int j);
template <>
__device__ __forceinline__ int4 others_of<3>(const int (&ids)[4], int j) {
  return make_int4(ids[j == 0 ? 1 : 0], ids[j <= 1 ? 2 : 1], ids[j <= 2 ? 3 : 2], 0);
}
template <>
__device__ __forceinline__ int2 others_of<2>(const int (&ids)[4], int j) {
  return make_int2(ids[j == 0 ? 1 : 0], ids[j <= 1 ? 2 : 1]);
}

// Lanes that claim a slot of the same vertex in the same step share one atomic when they are
// consecutive (equal-id runs: host Delaunay codes emit cells grouped around vertices).
// ~33 M scattered accesses on the ball h0=0.02 mesh (position gathers, slot claims, entry stores);
// l1tex__data_pipe_lsu_wavefronts is its top ncu metric, but removing most of them (dm_tiles.cuh) did not
// make stage A faster: the chain of dependent accesses is what it waits for.  Measured and rejected: claiming the
// slots before the cull decision is known + prefetching the next cell in a grid-stride loop (the
// longer live ranges spill and cost more than the overlap gains), more resident blocks, 4-byte entries.
template <int DIM, bool PAD = false>
__global__ void __launch_bounds__(DM_CS_THREADS, DM_CS_MINB) cull_scatter_kernel(
    const double* __restrict__ prog, const double* __restrict__ p, const int32_t* __restrict__ t, int64_t T,
    double geps, int mode, uint8_t* __restrict__ keep, int32_t* __restrict__ cnt,
    typename PCfg<DIM>::entry_t* __restrict__ bucket, int32_t* __restrict__ ovf_v,
    typename PCfg<DIM>::entry_t* __restrict__ ovf_e, int32_t* __restrict__ hv, int32_t* __restrict__ counters,
    int n_rows) {
  pdl_prologue();
  constexpr int CAP = PCfg<DIM>::CAP;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool k = false;
  int ids[4] = {0, 0, 0, 0};
  if (c < T) {
    load_cell<DIM>(t, c, ids);
    k = true;
    if (mode == 0) {
      double c0, c1, c2;
      cell_centroid<DIM, PAD>(p, ids, c0, c1, c2);
      k = sdf_eval(prog, DIM, c0, c1, c2) < -geps;
      keep[c] = k ? 1 : 0;
    } else if (mode == 1) {
      k = keep[c] != 0;
    }
  }
  if (cnt == nullptr) return;
  // The host triangulators hand the cells over grouped by their smallest id but with their own column
  // order (dm_cell_order.h; the sliver loop needs an unbiased column 0).  The centroid above was summed
  // in that order (bit-exactness); for the scatter the ids are sorted in registers, so that column j of
  // consecutive lanes holds equal ids again and the slot claims below merge.
  {
    auto cx = [](int& a, int& b) {
      const int lo_ = min(a, b), hi_ = max(a, b);
      a = lo_;
      b = hi_;
    };
    if (DIM == 3) {
      cx(ids[0], ids[1]);
      cx(ids[2], ids[3]);
      cx(ids[0], ids[2]);
      cx(ids[1], ids[3]);
      cx(ids[1], ids[2]);
    } else {
      cx(ids[0], ids[1]);
      cx(ids[1], ids[2]);
      cx(ids[0], ids[1]);
    }
  }
  // Slot claims of the DIM+1 vertices are independent: issue all the atomics first and only then
  // wait for them, so a warp pays ONE L2 round trip instead of DIM+1.
  const unsigned lt = (1u << lane) - 1u;
  int base[DIM + 1], rank[DIM + 1], leader[DIM + 1];
#pragma unroll
  for (int j = 0; j <= DIM; ++j) {
    // (vertices >= n_rows are ghost copies: nobody builds their rows, so nothing is pushed to them)
    const bool kj = k && ids[j] < n_rows;
    const int vj = kj ? ids[j] : -1 - lane;  // a culled lane never continues a run
    const int prev = __shfl_up_sync(FULL, vj, 1);
    const unsigned heads = __ballot_sync(FULL, lane == 0 || vj != prev);
    const int start = 31 - __clz(heads & (lt | (1u << lane)));
    const unsigned above = heads & ~(lt | (1u << lane));
    const int end = above ? __ffs(above) - 1 : 32;
    leader[j] = start;
    rank[j] = lane - start;
    base[j] = 0;
    if (kj && lane == start) base[j] = atomicAdd(cnt + ids[j], end - start);
  }
#pragma unroll
  for (int j = 0; j <= DIM; ++j) {
    const int slot = __shfl_sync(FULL, base[j], leader[j]) + rank[j];
    if (k && ids[j] < n_rows) {
      const typename PCfg<DIM>::entry_t e = others_of<DIM>(ids, j);
      if (slot < CAP) {
        bucket[(int64_t)ids[j] * CAP + slot] = e;
      } else {  // bucket full: spill (hull / hub vertices); the first spill lists the vertex as heavy
        const int o = atomicAdd(counters + 2, 1);
        ovf_v[o] = ids[j];
        ovf_e[o] = e;
        if (slot == CAP) hv[atomicAdd(counters + 3, 1)] = ids[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// B: adjacency rows
// ---------------------------------------------------------------------------------------------
constexpr int HASH_EMPTY = -1;
constexpr int HASH_MAXSTEPS = DM_HASH_MAXSTEPS;  // probe sequence longer than this -> vertex goes the heavy way

template <int LOGH>
__device__ __forceinline__ unsigned hash_slot(int x) {
  return ((unsigned)x * 2654435761u) >> (32 - LOGH);
}

// NC candidates per lane, all 32 lanes in lockstep.  Group-private open-addressing set in shared
// memory with plain loads / stores: a lane that finds its slot empty writes its key, the warp
// synchronises, and every lane re-reads the slot: whoever finds its own key there is done (it
// either won the slot or lost it to an equal key), everyone else moves to the next slot.  Slots
// only ever go EMPTY -> key, so a key is stored at the first slot of its probe sequence that was
// empty when it arrived and later equal keys find it before they find an empty slot.
// The kernel is bound by shared-memory wavefronts (random banks: ~3.5 per warp access) and by
// issue slots, so the body keeps no per-candidate flag: a finished candidate is parked on the
// lane's private word behind the table (tab[H + lg], one bank per lane of the warp, so parked
// accesses never conflict) with its key replaced by the parked word's content, and from then on
// it "hits" there for free.  Lanes without a candidate are handed a copy of a real key of the same
// vertex (a duplicate insert is a no-op).
constexpr int HASH_PARKED = -2;
template <int NC, int LOGH>
__device__ __forceinline__ void hash_insert_lockstep(int32_t* tab, int (&x)[NC], unsigned park, bool& punt) {
  constexpr unsigned HM = (1u << LOGH) - 1u;
  unsigned h[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) h[c] = hash_slot<LOGH>(x[c]);
  // A step (round 2, as in dm_tiles.cuh): every unfinished key reads its slot ONCE; whoever finds its own key is
  // done, whoever finds the slot empty writes its key; after the warp barrier only the writers look again (they
  // either won the slot or lost it -- to an equal key: done, to another: next slot).  No second barrier: a slot
  // that is being verified was written before the barrier, so it is not EMPTY any more and nobody writes it in
  // the next step.
  int steps = 0;
  bool any;
  do {
    int v[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) v[c] = tab[h[c]];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (v[c] == HASH_EMPTY) tab[h[c]] = x[c];
    __syncwarp();
    any = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      bool hit = v[c] == x[c];
      if (v[c] == HASH_EMPTY) hit = tab[h[c]] == x[c];
      const unsigned nxt = (h[c] + 1u) & HM;
      any = any | !hit;
      h[c] = hit ? park : nxt;
      x[c] = hit ? HASH_PARKED : x[c];
    }
  } while (__any_sync(FULL, any) && ++steps <= HASH_MAXSTEPS);
  punt = punt | any;  // (only possible when the step limit ended the loop) table (nearly) full: the heavy path takes the vertex
}

// L^d and h^d of the bar (a, p[w]); GRID: h is interpolated at the midpoint and stored at `hout`.
// upper == false (GRID only): the slot is a LOWER neighbour of the row's vertex -- h is evaluated and stored
// for the vertex update to read (same midpoint bits as the bar's owner forms, hence the same h), but the
// bar belongs to its other end and is not added to the sums.
template <int DIM, bool GRID>
__device__ __forceinline__ void bar_terms_at(const DmSizeFn& f, const GridGuess& gg, double a0, double a1, double a2,
                                             double b0, double b1, double b2, double* hout, double& sL, double& sH,
                                             bool upper = true) {
  double d0, d1, d2;
  double h = f.hconst;
  if (GRID) {
    // midpoint p[edges].sum(1)/2 (mesh_generator.py:699)
    h = size_eval(f, gg, (a0 + b0) / 2, (a1 + b1) / 2, (a2 + b2) / 2);
    *hout = h;
    if (!upper) return;
  }
  const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
  if (DIM == 2) {
    sL += L * L;
    sH += h * h;
  } else {
    sL += L * L * L;
    sH += h * h * h;
  }
}
template <int DIM, bool GRID>
__device__ __forceinline__ void bar_terms(const DmSizeFn& f, const GridGuess& gg, const double* __restrict__ pp,
                                          double a0, double a1, double a2, int w, double* hout, double& sL, double& sH,
                                          bool upper = true) {
  double b0, b1, b2;
  load_pt<DIM, true>(pp, w, b0, b1, b2);  // pp: the plan's padded point copy
  bar_terms_at<DIM, GRID>(f, gg, a0, a1, a2, b0, b1, b2, hout, sL, sH, upper);
}

// ---- heavy vertices: one block per vertex -------------------------------------------------------
// ascending compare-exchange network (bitonic with the "flip" first step, so every comparator
// points the same way): positions >= n behave as +inf and never move, hence any n works in place.
template <typename PTR>
__device__ __forceinline__ void block_sort_ascending(PTR a, int n) {
  for (int k = 2; (k >> 1) < n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < n) {
          const int x = a[i], y = a[l];
          if (x > y) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int HV_RANK = 768;  // up to this many candidates: one-pass rank sort instead of the network
constexpr int RG = 128;       // adjacency blocks per reduction group (second level of the bar sums)

// Where the bar sums of the adjacency kernel live in `partials` (pairs {sum L^d, sum h^d}):
//   [0, nb)            one per main block            (nb = blocks of VPB vertices)
//   [nb, nb + ng)      one per group of RG blocks    (ng = ceil(nb / RG))
//   [nb + ng, ...)     one per heavy vertex, at the RANK of the vertex id among the heavy vertices
// Every level is added in index order by one warp (lane-strided, fixed shuffle tree), so the scale
// ((sum L^d)/(sum h^d))^(1/d) (mesh_generator.py:700) is deterministic although blocks finish in any
// order: the last block of a group adds the group, the last arrival overall adds groups + heavy.
template <int DIM>
__device__ __forceinline__ void final_scale(const double* partials, int64_t nb, int ng, int nheavy, double* scalars) {
  const int lane = threadIdx.x & 31;
  double tL = 0.0, tH = 0.0;
  for (int g = lane; g < ng; g += 32) {
    tL += __ldcg(partials + 2 * (nb + g));
    tH += __ldcg(partials + 2 * (nb + g) + 1);
  }
  tL = warp_sum(tL);
  tH = warp_sum(tH);
  double hL = 0.0, hH = 0.0;
  for (int i = lane; i < nheavy; i += 32) {
    hL += __ldcg(partials + 2 * (nb + ng + i));
    hH += __ldcg(partials + 2 * (nb + ng + i) + 1);
  }
  hL = warp_sum(hL);
  hH = warp_sum(hH);
  if (lane == 0) {
    const double rl = tL + hL, rh = tH + hH;
    scalars[0] = rl;
    scalars[1] = rh;
    const double r = rl / rh;
    scalars[2] = DIM == 2 ? sqrt(r) : pow(r, 1.0 / 3.0);  // ** (1.0 / dim), mesh_generator.py:700
  }
}

// A vertex whose bucket overflowed (hull / hub vertices; listed by stage A): the whole block gathers
// its star from the bucket and the spill list, sorts + de-duplicates it and writes the row.
template <int DIM, int BAR, int THREADS = AB_THREADS>
__device__ __noinline__ void heavy_vertex(int ih, int nheavy, int novf, int32_t* s_val, int* s_scan, double* s_dbl,
                                             int* s_base, int* s_pos, const int32_t* __restrict__ cnt,
                                             const typename PCfg<DIM>::entry_t* __restrict__ bucket,
                                             const int32_t* __restrict__ ovf_v,
                                             const typename PCfg<DIM>::entry_t* __restrict__ ovf_e, int64_t N,
                                             int32_t* __restrict__ adj, int32_t* __restrict__ heap,
                                             int2* __restrict__ degs, const int32_t* __restrict__ hv,
                                             int32_t* __restrict__ counters, const DmSizeFn& f,
                                             const double* __restrict__ pp, double* __restrict__ hslot,
                                             double* hpart) {
  constexpr int CAP = PCfg<DIM>::CAP, RS = PCfg<DIM>::RS;
  const int tid = threadIdx.x;
  const int v = hv[ih];
  const int nc = cnt[v];
  const int nb = nc < CAP ? nc : CAP;  // entries in the bucket; the other nc - nb were spilled
  const int n = DIM * nc;              // candidates
  __syncthreads();
  if (tid == 0) {
    *s_base = atomicAdd(counters + 4, (int)round4(n));
    *s_pos = 0;
  }
  __syncthreads();
  int32_t* reg = heap + *s_base;  // candidates, later the row, live here
  const bool in_smem = n <= HV_SMEM;
  for (int i = tid; i < nb; i += THREADS) {
    const typename PCfg<DIM>::entry_t e = bucket[(int64_t)v * CAP + i];
    const int* ei = reinterpret_cast<const int*>(&e);
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      if (in_smem)
        s_val[i * DIM + c] = ei[c];
      else
        reg[i * DIM + c] = ei[c];
    }
  }
  for (int k = tid; k < novf; k += THREADS) {
    if (ovf_v[k] == v) {
      const int i = nb + atomicAdd(s_pos, 1);
      const typename PCfg<DIM>::entry_t e = ovf_e[k];
      const int* ei = reinterpret_cast<const int*>(&e);
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        if (in_smem)
          s_val[i * DIM + c] = ei[c];
        else
          reg[i * DIM + c] = ei[c];
      }
    }
  }
  // rank of v among the heavy vertices: where its bar sums go (see final_scale)
  int vrank = 0;
  for (int j0 = 0; j0 < nheavy; j0 += THREADS) {
    const int j = j0 + tid;
    vrank += __syncthreads_count(j < nheavy && hv[j] < v);
  }
  __syncthreads();
  int U = 0, lo = 0;
  if (n <= HV_RANK) {
    // one pass: candidate i is the first of its value if no equal value precedes it, and its
    // row position is the number of DISTINCT smaller values = #{j : val[j] < val[i], j first}.
    constexpr int PER = HV_RANK / THREADS;
    bool first[PER];
    int x[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = tid + u * THREADS;
      first[u] = i < n;
      x[u] = i < n ? s_val[i] : 0;
    }
#pragma unroll 8  // the shared-memory loads of eight steps in flight (latency, not issue, bounds this sweep)
    for (int j = 0; j < n; ++j) {
      const int y = s_val[j];
#pragma unroll
      for (int u = 0; u < PER; ++u) first[u] = first[u] && !(y == x[u] && j < tid + u * THREADS);
    }
    __syncthreads();
    int32_t* s_first = s_val + HV_RANK;  // HV_SMEM >= 2 * HV_RANK
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = tid + u * THREADS;
      if (i < n) s_first[i] = first[u] ? 1 : 0;
    }
    __syncthreads();
    int r[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) r[u] = 0;
#pragma unroll 8
    for (int j = 0; j < n; ++j) {
      const int y = s_val[j];
      const int fj = s_first[j];
#pragma unroll
      for (int u = 0; u < PER; ++u) r[u] += (fj && y < x[u]) ? 1 : 0;
    }
    int nf = 0, nl = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      if (first[u]) {
        reg[r[u]] = x[u];
        ++nf;
        nl += x[u] < v ? 1 : 0;
      }
    }
    int tot;
    block_exclusive_scan(nf, tot, s_scan);
    U = tot;
    block_exclusive_scan(nl, tot, s_scan);
    lo = tot;
  } else {
    if (in_smem)
      block_sort_ascending(s_val, n);
    else
      block_sort_ascending(reg, n);
    // unique: chunk by chunk; an element's output position never exceeds its input position and
    // all reads of a chunk complete before its writes, so compaction in place is safe
    for (int c0 = 0; c0 < n; c0 += THREADS) {
      const int i = c0 + tid;
      int x = 0;
      bool first = false;
      if (i < n) {
        x = in_smem ? s_val[i] : reg[i];
        first = i == 0 || (in_smem ? s_val[i - 1] : reg[i - 1]) != x;
      }
      int tot;
      const int off = block_exclusive_scan(first ? 1 : 0, tot, s_scan);  // syncs: reads done
      if (first) reg[U + off] = x;
      const int nl = __syncthreads_count(first && x < v);
      U += tot;
      lo += nl;
    }
  }
  __syncthreads();
  int32_t* row = adj + (int64_t)v * RS;
  int64_t sbase = (int64_t)v * RS;
  if (U <= RS) {
    for (int i = tid; i < U; i += THREADS) row[i] = reg[i];
  } else {
    if (tid == 0) row[0] = *s_base;
    sbase = N * RS + *s_base;
  }
  if (tid == 0) {
    degs[v] = make_int2(U, lo);
    atomicAdd(counters, U - lo);
  }
  if (BAR >= 0) {
    double sL = 0.0, sH = 0.0;
    double a0, a1, a2;
    load_pt<DIM, true>(pp, v, a0, a1, a2);
    GridGuess gg;
    if (BAR == 1) gg = grid_guess(f);
    for (int j = (BAR == 1 && H_ALL_SLOTS ? 0 : lo) + tid; j < U; j += THREADS)
      bar_terms<DIM, BAR == 1>(f, gg, pp, a0, a1, a2, reg[j], hslot + sbase + j, sL, sH, j >= lo);
    const double bl = block_sum(sL, s_dbl);
    const double bh = block_sum(sH, s_dbl);
    if (tid == 0) {
      hpart[2 * vrank] = bl;
      hpart[2 * vrank + 1] = bh;
    }
  }
}

// A neighbour set that does not fit the group table (more than ~H distinct ids within CAP cells:
// not a manifold star, but legal input): the group selects the ids in ascending order straight from
// its bucket, minimum by minimum, into the heap, then writes the row and does its bar pass.  Called
// by the whole warp; groups with punt == false only take part in the shuffles.
struct RowSums {
  int bars;
  double sL, sH;
};
template <int DIM, int BAR>
__device__ __noinline__ RowSums select_row(bool punt, int n, int v,
                                           const typename PCfg<DIM>::entry_t* __restrict__ brow, int64_t N,
                                           int32_t* __restrict__ adj, int32_t* __restrict__ heap,
                                           int2* __restrict__ degs, int32_t* __restrict__ counters,
                                           const DmSizeFn& f, const double* __restrict__ pp,
                                           double* __restrict__ hslot) {
  constexpr int RS = PCfg<DIM>::RS, G = PCfg<DIM>::G;
  RowSums out;
  out.bars = 0;
  out.sL = 0.0;
  out.sH = 0.0;
  typedef typename PCfg<DIM>::entry_t entry_t;
  const int lg = threadIdx.x % G;
  int base = 0;
  if (punt && lg == 0) base = atomicAdd(counters + 4, (int)round4(DIM * n));
  base = __shfl_sync(FULL, base, 0, G);
  const int* bi = reinterpret_cast<const int*>(brow);
  constexpr int EW = sizeof(entry_t) / 4;  // ints per entry (DIM ids + padding)
  int last = -1, Up = 0, lop = 0;
  for (;;) {
    int m = 0x7fffffff;
    if (punt)
      for (int i = lg; i < DIM * n; i += G) {
        const int x = bi[(i / DIM) * EW + i % DIM];
        if (x > last && x < m) m = x;
      }
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(FULL, m, d, G));
    const bool more = punt && m != 0x7fffffff;
    if (!__any_sync(FULL, more)) break;
    if (more) {
      if (lg == 0) heap[base + Up] = m;
      ++Up;
      lop += m < v ? 1 : 0;
      last = m;
    }
  }
  __syncwarp();
  if (!punt) return out;
  const int32_t* srt = heap + base;
  int32_t* row = adj + (int64_t)v * RS;
  int64_t sbase = (int64_t)v * RS;
  if (Up <= RS) {
    if (4 * lg < Up) reinterpret_cast<int4*>(row)[lg] = reinterpret_cast<const int4*>(srt)[lg];
  } else {
    if (lg == 0) row[0] = base;
    sbase = N * RS + base;
  }
  if (lg == 0) {
    degs[v] = make_int2(Up, lop);
    out.bars = Up - lop;
  }
  if (BAR >= 0 && Up > (BAR == 1 && H_ALL_SLOTS ? 0 : lop)) {
    double a0, a1, a2;
    load_pt<DIM, true>(pp, v, a0, a1, a2);
    GridGuess gg;
    if (BAR == 1) gg = grid_guess(f);
    for (int j = (BAR == 1 && H_ALL_SLOTS ? 0 : lop) + lg; j < Up; j += G)
      bar_terms<DIM, BAR == 1>(f, gg, pp, a0, a1, a2, srt[j], hslot + sbase + j, out.sL, out.sH, j >= lop);
  }
  return out;
}

// BAR: -1 rows only ; 0 also accumulate sum L^d, sum h^d of the upper bars with constant h ;
//      1 the same with gridded fh (h stored per upper slot).
// Grid = HV_BLOCKS heavy-vertex blocks (they start first and run beside the main blocks; the heavy
// list is complete when the kernel starts) + one main block per VPB vertices.
// (gridded fh: one resident block less, so that the interpolation does not spill: -2..-10 % on the EAGE-shaped
//  workloads; constant h is 5 % faster with the fifth block)
template <int DIM, int BAR>
__global__ void __launch_bounds__(AB_THREADS, BAR == 1 ? DM_AB_MINB - 1 : DM_AB_MINB) adjacency_kernel(
    const int32_t* __restrict__ cnt, const typename PCfg<DIM>::entry_t* __restrict__ bucket,
    const int32_t* __restrict__ ovf_v, const typename PCfg<DIM>::entry_t* __restrict__ ovf_e, int64_t N, int64_t NR,
    int32_t* __restrict__ adj, int32_t* __restrict__ heap, int2* __restrict__ degs, const int32_t* __restrict__ hv,
    int32_t* __restrict__ counters, const __grid_constant__ DmSizeFn f, const double* __restrict__ pp,
    double* __restrict__ hslot, double* partials, int32_t* gdone, int32_t* total_done, double* scalars) {
  // N: vertices (sizes the per-slot arrays) ; NR <= N: vertices that get a row (the grid covers NR)
  pdl_prologue();
  constexpr int CAP = PCfg<DIM>::CAP, RS = PCfg<DIM>::RS, G = PCfg<DIM>::G, LOGH = PCfg<DIM>::LOGH;
  constexpr int H = 1 << LOGH;
  constexpr int VPB = AB_THREADS / G;  // vertices per block
  constexpr int SPL = H / G;           // table slots per lane (8)
  static_assert(SPL == 8, "compaction below reads two int4 per lane");
  // group stride H + 32/(32/G) ints: the groups of one warp start in different banks, so the
  // broadcast reads of the rank sort (same k in every group) do not conflict
  constexpr int GS = H + G;  // G parking words behind the table / >= 4 ints of padding behind the list;
                             // G == 32 / (groups per warp), so the groups of a warp start in different banks
  static_assert(G >= 4 && 32 % G == 0, "group layout");
  static_assert((GS * 4) % 16 == 0, "int4 access");
  static_assert(2 * VPB * GS <= HV_SMEM && HV_SMEM >= 2 * HV_RANK, "shared buffer");
  __shared__ __align__(16) int32_t s_raw[HV_SMEM];  // main: tables | lists ; heavy block: candidates
  __shared__ int s_scan[33];
  __shared__ double s_dbl[32];
  __shared__ int s_base, s_pos;
  __shared__ int s_arrived;  // warps that have left their totals (see the end of the kernel)
  __shared__ bool s_fin;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t nbm = (int64_t)gridDim.x - HV_BLOCKS;  // main blocks
  const int ng = (int)((nbm + RG - 1) / RG);

  if (blockIdx.x < HV_BLOCKS) {  // ---------------- heavy-vertex block
    const int nheavy = counters[3], novf = counters[2];
    for (int ih = blockIdx.x; ih < nheavy; ih += HV_BLOCKS)
      heavy_vertex<DIM, BAR>(ih, nheavy, novf, s_raw, s_scan, s_dbl, &s_base, &s_pos, cnt, bucket, ovf_v, ovf_e, N, adj,
                             heap, degs, hv, counters, f, pp, hslot, partials + 2 * (nbm + ng));
    if (BAR < 0) return;
    __syncthreads();
    if (tid == 0) s_fin = arrive_acq_rel(total_done) == ng + HV_BLOCKS - 1;
    __syncthreads();
    if (s_fin && tid < 32) final_scale<DIM>(partials, nbm, ng, nheavy, scalars);
    return;
  }

  // ---------------- main block
  const int64_t bidx = (int64_t)blockIdx.x - HV_BLOCKS;
  int32_t* s_tab = s_raw;
  int32_t* s_lst = s_raw + VPB * GS;
  const int lg = tid % G, grp = tid / G;
  const int64_t v = bidx * VPB + grp;
  __shared__ GridGuess s_gg;  // gridded fh: what the index search needs of the axes, once per block
  if (tid == 0) {
    s_arrived = 0;
    if (BAR == 1) s_gg = grid_guess(f);
  }
  __syncthreads();  // the only block barrier: at the very start, where no warp has to wait long
  int32_t* tab = s_tab + grp * GS;
  int32_t* lst = s_lst + grp * GS;

  // ---- empty table
  {
    int4* t4 = reinterpret_cast<int4*>(tab + lg * SPL);
    t4[0] = make_int4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY);
    t4[1] = make_int4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY);
    tab[H + lg] = HASH_PARKED;  // the lane's parking word (see hash_insert_lockstep)
  }
  int n = v < NR ? cnt[v] : 0;
  const bool heavy = n > CAP;  // overflowed bucket: a heavy-vertex block builds this row
  if (heavy) n = 0;
  bool punt = false;
  __syncwarp();

  // ---- insert every candidate of the bucket: each round, lane lg takes EPL entries (lg, lg+G, ...).
  //      The loads of round r+1 are issued before round r is hashed (the bucket comes from DRAM).
  constexpr int EPL = 2;
  typedef typename PCfg<DIM>::entry_t entry_t;
  const entry_t* brow = bucket + v * CAP;
  const int key0 = n > 0 ? reinterpret_cast<const int*>(brow)[0] : 0;  // filler for idle lanes
  entry_t cur[EPL], nxt[EPL];
#pragma unroll
  for (int u = 0; u < EPL; ++u)
    if (lg + u * G < n) cur[u] = brow[lg + u * G];
  for (int i = lg; __any_sync(FULL, i < n); i += EPL * G) {
#pragma unroll
    for (int u = 0; u < EPL; ++u)
      if (i + (EPL + u) * G < n) nxt[u] = brow[i + (EPL + u) * G];
    int x[EPL * DIM];
#pragma unroll
    for (int u = 0; u < EPL; ++u) {
      const bool a = i + u * G < n;
      const int* ei = reinterpret_cast<const int*>(&cur[u]);
#pragma unroll
      for (int c = 0; c < DIM; ++c) x[u * DIM + c] = a ? ei[c] : key0;
    }
    hash_insert_lockstep<EPL * DIM, LOGH>(tab, x, (unsigned)(H + lg), punt);
#pragma unroll
    for (int u = 0; u < EPL; ++u) cur[u] = nxt[u];
  }
  // group-wide punt flag
  const unsigned gsh = (unsigned)(lane - lg);
  const unsigned gmask = (G == 32 ? FULL : ((1u << G) - 1u)) << gsh;
  punt = (__ballot_sync(FULL, punt) & gmask) != 0u;

  // ---- compact the occupied slots (each lane owns SPL consecutive slots)
  int vals[SPL];
  {
    const int4* t4 = reinterpret_cast<const int4*>(tab + lg * SPL);
    const int4 a = t4[0], b = t4[1];
    vals[0] = a.x; vals[1] = a.y; vals[2] = a.z; vals[3] = a.w;
    vals[4] = b.x; vals[5] = b.y; vals[6] = b.z; vals[7] = b.w;
  }
  int c = 0;
#pragma unroll
  for (int k = 0; k < SPL; ++k) c += vals[k] != HASH_EMPTY ? 1 : 0;
  int inc = c;
#pragma unroll
  for (int d = 1; d < G; d <<= 1) {
    const int o = __shfl_up_sync(FULL, inc, d, G);
    if (lg >= d) inc += o;
  }
  int U = __shfl_sync(FULL, inc, G - 1, G);
  if (punt || n == 0) U = 0;  // n == 0: nothing but filler keys went in
  {
    int off = inc - c;
#pragma unroll
    for (int k = 0; k < SPL; ++k)
      if (vals[k] != HASH_EMPTY) lst[off++] = vals[k];
  }
  __syncwarp();

  // ---- rank sort: lst (unsorted, unique, padded with +inf to a multiple of 4) -> tab (ascending);
  //      two elements per lane and four list entries per shared-memory load
  if (lg < 4) lst[U + lg] = 0x7fffffff;
  __syncwarp();
  int lo = 0;
  for (int i = lg; i < U; i += 2 * G) {
    const int x0 = lst[i];
    const bool two = i + G < U;
    const int x1 = two ? lst[i + G] : 0x7fffffff;
    int r0 = 0, r1 = 0;
    for (int k = 0; k < U; k += 4) {
      const int4 q = *reinterpret_cast<const int4*>(lst + k);
      r0 += (q.x < x0 ? 1 : 0) + (q.y < x0 ? 1 : 0) + (q.z < x0 ? 1 : 0) + (q.w < x0 ? 1 : 0);
      r1 += (q.x < x1 ? 1 : 0) + (q.y < x1 ? 1 : 0) + (q.z < x1 ? 1 : 0) + (q.w < x1 ? 1 : 0);
    }
    tab[r0] = x0;
    lo += x0 < (int)v ? 1 : 0;
    if (two) {
      tab[r1] = x1;
      lo += x1 < (int)v ? 1 : 0;
    }
  }
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) lo += __shfl_xor_sync(FULL, lo, d, G);
  __syncwarp();

  // ---- write the row
  int bars = 0;
  double sL = 0.0, sH = 0.0;
  if (v < NR && !heavy && !punt) {
    int32_t* row = adj + v * RS;
    int64_t sbase = v * RS;
    if (U <= RS) {
      // RS ints = G lanes x int4 (3-D: 8 x 16 B = one 128-B line; 2-D: 4 x 16 B)
      static_assert(RS == 4 * G, "one int4 per lane");
      if (4 * lg < U) reinterpret_cast<int4*>(row)[lg] = reinterpret_cast<const int4*>(tab)[lg];
    } else {  // rare: more than RS neighbours but still within the group hash
      int base = 0;
      if (lg == 0) base = atomicAdd(counters + 4, (int)round4(U));
      base = __shfl_sync(gmask, base, 0, G);  // group-uniform branch: only this group's lanes are here
      for (int i = lg; i < U; i += G) heap[base + i] = tab[i];
      if (lg == 0) row[0] = base;
      sbase = N * RS + base;
    }
    if (lg == 0) {
      degs[v] = make_int2(U, lo);
      bars = U - lo;
    }
    if (BAR >= 0 && U > (BAR == 1 && H_ALL_SLOTS ? 0 : lo)) {
      // bar pass (mesh_generator.py:696-700): L^d, h^d of the row's upper bars; gridded fh: h of EVERY slot
      double a0, a1, a2;
      load_pt<DIM, true>(pp, v, a0, a1, a2);
      GridGuess gg;
      if (BAR == 1) gg = s_gg;
      for (int j = (BAR == 1 && H_ALL_SLOTS ? 0 : lo) + lg; j < U; j += G)
        bar_terms<DIM, BAR == 1>(f, gg, pp, a0, a1, a2, tab[j], hslot + sbase + j, sL, sH, j >= lo);
    }
  }
  // ---- a neighbour set that does not fit the group table: rare, out of line.  Warp-uniform branch.
  if (__any_sync(FULL, punt)) {
    const RowSums rs = select_row<DIM, BAR>(punt, n, (int)v, brow, N, adj, heap, degs, counters, f, pp, hslot);
    bars += rs.bars;
    sL += rs.sL;
    sH += rs.sH;
  }
  // ---- block totals without a barrier: every warp leaves its (fixed shuffle tree) sums in shared
  //      memory; the warp that arrives last adds them in warp order and publishes the block's values
  __shared__ double s_wL[AB_THREADS / 32], s_wH[AB_THREADS / 32];
  __shared__ int s_wbars[AB_THREADS / 32];
  const int wid = tid >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) bars += __shfl_down_sync(FULL, bars, d);
  if (BAR >= 0) {
    sL = warp_sum(sL);
    sH = warp_sum(sH);
  }
  int lead = 0;  // lane 0: 1 = this warp closes its reduction group
#if DM_AB_BARRIER
  if (lane == 0) {
    s_wbars[wid] = bars;
    if (BAR >= 0) {
      s_wL[wid] = sL;
      s_wH[wid] = sH;
    }
  }
  __syncthreads();
  if (wid != 0) return;
  if (lane == 0) {
    {
#else
  if (lane == 0) {
    s_wbars[wid] = bars;
    if (BAR >= 0) {
      s_wL[wid] = sL;
      s_wH[wid] = sH;
    }
    __threadfence_block();
    if (atomicAdd(&s_arrived, 1) == AB_THREADS / 32 - 1) {  // the block's last warp
      __threadfence_block();
#endif
      int tb = 0;
      double tL = 0.0, tH = 0.0;
#pragma unroll
      for (int w = 0; w < AB_THREADS / 32; ++w) {
        tb += s_wbars[w];
        if (BAR >= 0) {
          tL += s_wL[w];
          tH += s_wH[w];
        }
      }
      if (tb) atomicAdd(counters, tb);  // unique bars owned by this block's vertices
      if (BAR >= 0) {
        partials[2 * bidx] = tL;
        partials[2 * bidx + 1] = tH;
        const int64_t g = bidx / RG;
        const int gsize = (int)(nbm - g * RG < RG ? nbm - g * RG : RG);
        lead = arrive_acq_rel(gdone + g) == gsize - 1 ? 1 : 0;
      }
    }
  }
  if (BAR < 0) return;
  if (!__shfl_sync(FULL, lead, 0)) return;
  // ---- this warp saw the last block of its group finish: add the group's block sums in block order
  {
    __syncwarp();  // lane 0's acquire orders the other lanes' reads as well
    const int64_t g = bidx / RG;
    const int gsize = (int)(nbm - g * RG < RG ? nbm - g * RG : RG);
    double tL = 0.0, tH = 0.0;
    for (int i = lane; i < gsize; i += 32) {
      tL += __ldcg(partials + 2 * (g * RG + i));
      tH += __ldcg(partials + 2 * (g * RG + i) + 1);
    }
    tL = warp_sum(tL);
    tH = warp_sum(tH);
    int fin = 0;
    if (lane == 0) {
      partials[2 * (nbm + g)] = tL;
      partials[2 * (nbm + g) + 1] = tH;
      fin = arrive_acq_rel(total_done) == ng + HV_BLOCKS - 1 ? 1 : 0;
    }
    if (__shfl_sync(FULL, fin, 0)) {
      __syncwarp();
      final_scale<DIM>(partials, nbm, ng, counters[3], scalars);
    }
  }
}

// bar ids: rowptr[v] = #upper neighbours (scanned afterwards)
__global__ void upper_count_kernel(const int2* __restrict__ degs, int64_t N, int64_t NR, int32_t* __restrict__ rowptr) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) {
    int n = 0;
    if (v < NR) {  // vertices without a row (ghost copies) own no bar
      const int2 d = degs[v];
      n = d.x - d.y;
    }
    rowptr[v] = n;
  }
}

template <int DIM>
__global__ void bars_pairs_kernel(const Rows<DIM> R, const int32_t* __restrict__ rowptr, int64_t N,
                                  int32_t* __restrict__ pairs) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int2 d = R.degs[v];
  const int32_t* row = R.row(v, d.x);
  int64_t e = rowptr[v];
  for (int j = d.y; j < d.x; ++j, ++e) {
    pairs[2 * e] = (int32_t)v;
    pairs[2 * e + 1] = row[j];
  }
}

// h of every unique bar in bar order (diagnostics / parity tests): hmode as in bar_pass_kernel
template <int DIM>
__global__ void bar_sizes_kernel(const DmSizeFn f, const Rows<DIM> R, const int32_t* __restrict__ rowptr,
                                 const double* __restrict__ hslot, const double* __restrict__ hbar, int64_t N,
                                 double* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int2 d = R.degs[v];
  const int64_t base = R.slot_base(v, d.x);
  int64_t e = rowptr[v];
  for (int j = d.y; j < d.x; ++j, ++e)
    out[e] = f.kind == DM_SIZE_CONST ? f.hconst : (f.kind == DM_SIZE_GRID ? hslot[base + j] : hbar[e]);
}

// bar id of the bar (u, v), u < v: position of v among u's upper neighbours
template <int DIM>
__device__ __forceinline__ int bar_id_of(const Rows<DIM>& R, const int32_t* __restrict__ rowptr, int u, int v) {
  const int2 d = R.degs[u];
  const int32_t* row = R.row(u, d.x);
  for (int j = d.y; j < d.x; ++j)
    if (row[j] == v) return rowptr[u] + (j - d.y);
  return -1;
}

// ---------------------------------------------------------------------------------------------
// C: bar pass (stand-alone version for the staged path).  HMODE 0: constant h ; 1: gridded fh (h stored at the upper directed slot) ;
//               2: h per bar id supplied by the caller ; 3: write bar midpoints only
// ---------------------------------------------------------------------------------------------
template <int DIM, int HMODE>
__global__ void __launch_bounds__(PL_THREADS) bar_pass_kernel(const DmSizeFn f, const double* __restrict__ p,
                                                              const Rows<DIM> R, const int32_t* __restrict__ rowptr,
                                                              int64_t N, double* __restrict__ hslot,
                                                              const double* __restrict__ hbar, double* __restrict__ mid,
                                                              double* partials, int32_t* done, double* scalars) {
  __shared__ double sm[32];
  __shared__ bool s_last;
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double sL = 0.0, sH = 0.0;
  if (v < N) {
    const int2 dg = R.degs[v];
    const int lo = dg.y, m = dg.x;
    // gridded fh (HMODE 1): h is left at EVERY slot of the row, lower neighbours included (the vertex update
    // reads it there); the sums run over the upper bars only
    const int first = (HMODE == 1 && H_ALL_SLOTS) ? 0 : lo;
    if (m > first) {
      const int32_t* row = R.row(v, m);
      const int64_t base = R.slot_base(v, m);
      double a0, a1, a2;
      load_pt<DIM>(p, v, a0, a1, a2);
      GridGuess gg;
      if (HMODE == 1) gg = grid_guess(f);
      // rows are 16-B aligned and padded to a multiple of 4 ints: walk them in int4 chunks
      for (int j0 = first & ~3; j0 < m; j0 += 4) {
        const int4 q = *reinterpret_cast<const int4*>(row + j0);
        const int wq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j < first || j >= m) continue;
          const int w = wq[u];
          double b0, b1, b2, d0, d1, d2;
          load_pt<DIM>(p, w, b0, b1, b2);
          // midpoint p[edges].sum(1)/2 (mesh_generator.py:699)
          const double m0 = (a0 + b0) / 2, m1 = (a1 + b1) / 2, m2 = (a2 + b2) / 2;
          if (HMODE == 1 && j < lo) {
            hslot[base + j] = size_eval(f, gg, m0, m1, m2);
            continue;
          }
          if (HMODE == 3) {
            store_pt<DIM>(mid, rowptr[v] + (j - lo), m0, m1, m2);
            continue;
          }
          const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
          double h;
          if (HMODE == 0) {
            h = f.hconst;
          } else if (HMODE == 1) {
            h = size_eval(f, gg, m0, m1, m2);
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
      *done = 0;  // self-resetting: the stage can run again without stage A (row re-use)
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

// One thread per vertex.  The reference accumulates bar by bar (coo_matrix.toarray): row v
// receives -Fvec of its lower bars (u,v), u ascending, then +Fvec of its upper bars (v,w), w
// ascending.  The row is sorted, and -(F/L*(p[u]-p[v])) == (F/L)*(p[v]-p[u]) exactly, so one
// ascending sweep with d = p[v]-p[nbr] reproduces the reference's sum bit for bit.
// The kernel is latency bound (dependent row load -> position gather -> sqrt -> divide per
// neighbour), so the row is consumed four neighbours at a time: their gathers and their
// sqrt / divide chains are independent and in flight together, only the additions into F are
// serial.  Slots past the end of the row are redirected to the vertex itself (a valid address;
// their term is discarded).  (A lane-group-per-vertex variant with the ordered sum done through
// shared memory was measured at 2.4x the time of this one: the serial tail then holds a whole block.)
constexpr int VU_THREADS = DM_VU_THREADS;

// The listed vertices (those that left a level set, see vertex_update_kernel) get the reference's
// sequence of projections, level after level, starting from the updated position in p_out: a vertex
// that is not listed is one no level would have moved, so the result equals projecting everybody.
constexpr int PJ_THREADS = 128;
constexpr int PJ_BLOCKS = 592;
template <int DIM>
__global__ void __launch_bounds__(PJ_THREADS) project_list_kernel(const Levels lv, double deps, double h0,
                                                                  const int32_t* __restrict__ esc,
                                                                  int32_t* esc_count, int32_t* done,
                                                                  double* __restrict__ p_out) {
  pdl_prologue();
  const int n = *reinterpret_cast<volatile int32_t*>(esc_count);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = esc[i];
    double x0, x1, x2;
    load_pt<DIM>(p_out, v, x0, x1, x2);
    for (int l = 0; l < lv.n; ++l) sdf_project(lv.prog[l], DIM, deps, h0, l, x0, x1, x2);
    store_pt<DIM>(p_out, v, x0, x1, x2);
  }
  // the last block to finish empties the list, so stage D can run again without stage A
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(done, 1) == (int)gridDim.x - 1) {
    *esc_count = 0;
    *done = 0;
  }
}

template <int DIM, int HMODE, bool PAD, bool FUSE>
__global__ void __launch_bounds__(VU_THREADS, DM_VU_MINB) vertex_update_kernel(
    const DmSizeFn f, const double* __restrict__ p, const double* __restrict__ pg, double* __restrict__ p_out,
    const Rows<DIM> R, const int32_t* __restrict__ rowptr, const double* __restrict__ hslot,
    const double* __restrict__ hbar, const double* scalars_in, int64_t N, Levels lv, double L0mult, double delta_t,
    double deps, double h0, int64_t nfix, const uint8_t* __restrict__ fixed, double* __restrict__ Ftot,
    double* partials, int32_t* done, double* scalars, int32_t* __restrict__ esc, int32_t* __restrict__ esc_count) {
  pdl_prologue();
  __shared__ double s_wmax[VU_THREADS / 32];
  __shared__ int s_arrived;
  if (threadIdx.x == 0) s_arrived = 0;
  __syncthreads();  // the only block barrier: at the very start, where no warp has to wait
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double f2 = 0.0;
  if (v < N) {
    const double scale = __ldcg(scalars_in + 2);
    const double k0 = L0mult;
    double a0, a1, a2;
    load_pt<DIM, PAD>(pg, v, a0, a1, a2);
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    const int2 dg = R.degs[v];
    const int lo = dg.y, m = dg.x;
    const int32_t* row = R.row(v, m);
    const int64_t base = R.slot_base(v, m);
    // rows are 16-B aligned and padded to a multiple of 4 ints; the next chunk of the row is requested before
    // this chunk's gathers (the row load is otherwise the head of every chunk's dependent chain)
    int4 q = make_int4(0, 0, 0, 0);
    if (m > 0) q = *reinterpret_cast<const int4*>(row);
    for (int j0 = 0; j0 < m; j0 += 4) {
      const int wq[4] = {q.x, q.y, q.z, q.w};
      if (j0 + 4 < m) q = *reinterpret_cast<const int4*>(row + j0 + 4);
      double c0[4], c1[4], c2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        const int w = j < m ? wq[u] : (int)v;
        double b0, b1, b2, d0, d1, d2;
        load_pt<DIM, PAD>(pg, w, b0, b1, b2);
        // (a slot past the end of the row is a bar of length 1, its term discarded: sqrt(0) would send the whole
        //  warp through the square root's special-case subroutine in nearly every chunk)
        if (j >= m) {
          b0 = a0 - 1.0;
          asm("" : "+d"(b0));
        }
        const double L = bar_length<DIM>(a0, a1, a2, b0, b1, b2, d0, d1, d2);
        double h;
        if (HMODE == 0) {
          h = f.hconst;
        } else if (HMODE == 1) {
          // gridded fh: h of every bar (v, w) waits at its slot in v's row -- the bar pass that follows the
          // row build evaluates it for the lower slots as well as the upper ones (same midpoint bits at both
          // ends of a bar, hence the same h), with a lane group per vertex; here it is one coalesced read
          h = j < m ? hslot[base + j] : 0.0;
        } else {
          h = 0.0;
          if (j < m) h = hbar[j >= lo ? rowptr[v] + (j - lo) : bar_id_of<DIM>(R, rowptr, w, (int)v)];
        }
        double Fs = h * k0 * scale - L;  // L0 - L  (mesh_generator.py:700-702)
        if (Fs < 0) Fs = 0;
        // F / L.  A bar longer than L0 carries no force, and a ZERO numerator is the one operand that takes the
        // fp64 division off its fast path (ncu: four warps in five went through the ~70-instruction subroutine
        // because one of their lanes had Fs == 0).  0 / L is +0 exactly, so those lanes divide 1 / L instead and
        // the quotient is replaced by 0: same bits, no subroutine.  (Written with selects on the OPERANDS: a
        // branch around the division is turned back into "divide, then select" by the compiler.)
        double qn = Fs > 0.0 ? Fs : 1.0;
        asm("" : "+d"(qn));  // opaque to the optimiser, or it folds the select back into "Fs / L, then select"
        const double qq = qn / L;
        const double qf = Fs > 0.0 ? qq : 0.0;
        c0[u] = qf * d0;
        c1[u] = qf * d1;
        c2[u] = qf * d2;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < m) {
          F0 = F0 + c0[u];
          F1 = F1 + c1[u];
          if (DIM == 3) F2 = F2 + c2[u];
        }
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
    // p += delta_t * Ftot (mesh_generator.py:502).  The Newton projection per level (:505-506) moves
    // only the few vertices that left a level set; doing it here would leave 1-2 lanes of nearly
    // every warp walking the finite-difference path while 30 wait, so those vertices are only
    // LISTED here (fd evaluated once per level, all lanes) and projected by project_list_kernel.
    double x0 = a0 + delta_t * F0, x1 = a1 + delta_t * F1, x2 = a2 + delta_t * F2;
    bool out = false;
    for (int l = 0; l < lv.n; ++l) {
      const double d = sdf_eval(lv.prog[l], DIM, x0, x1, x2);
      out = out || (l == 0 ? (d > 0.0) : (d > 0.0 && d < h0 / 1.5));
    }
    if (FUSE) {
      // the escaped lanes walk the finite-difference path here, level after level as project_list_kernel does,
      // while the rest of their warp waits
      if (out)
        for (int l = 0; l < lv.n; ++l) sdf_project(lv.prog[l], DIM, deps, h0, l, x0, x1, x2);
      store_pt<DIM>(p_out, v, x0, x1, x2);
    } else {
      store_pt<DIM>(p_out, v, x0, x1, x2);
      if (out) esc[atomicAdd(esc_count, 1)] = (int32_t)v;
    }
  }
  // max |F|^2 without a block barrier (the warps of a block finish at different times: rows differ in length):
  // every warp leaves its maximum in shared memory, the warp that arrives last publishes the block's, and the
  // warp that sees the last block arrive reduces the blocks' maxima (a maximum does not depend on the order).
  constexpr int NW = VU_THREADS / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double wm = warp_max(f2);
  int last = 0;
  if (lane == 0) {
    s_wmax[wid] = wm;
    __threadfence_block();
    if (atomicAdd(&s_arrived, 1) == NW - 1) {
      __threadfence_block();
      double bm = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) bm = fmax(bm, s_wmax[w]);
      partials[blockIdx.x] = bm;
      last = arrive_acq_rel(done) == (int)gridDim.x - 1 ? 1 : 0;
    }
  }
  if (!__shfl_sync(FULL, last, 0)) return;
  __syncwarp();  // lane 0's acquire orders the other lanes' reads as well
  double mx = 0.0;
  for (int64_t i = lane; i < (int64_t)gridDim.x; i += 32) mx = fmax(mx, __ldcg(partials + i));
  mx = warp_max(mx);
  if (lane == 0) {
    scalars[3] = mx;
    scalars[4] = delta_t * sqrt(mx);  // maxdp, mesh_generator.py:514
    *done = 0;
  }
}

// max_v |p[v] - p_ref[v]|_2 / h_v -> scalars[5]: the displacement since the last retriangulation (the
// `ttol` test of DistMesh, Persson & Strang; north_star (5)), measured in units of the LOCAL mesh size:
// h_v = 1 (REL = 0: absolute displacement), or fh(p[v]) for a constant / gridded size function
// (REL = 1).  On a graded mesh the coarse regions move far more than ttol * hmin every iteration, so
// an absolute test retriangulates every time; relative to the local size it behaves like the uniform
// case.  Last block reduces; self-resetting.
template <int DIM, int REL>
__global__ void __launch_bounds__(PL_THREADS) displacement_kernel(const DmSizeFn f, const double* __restrict__ p,
                                                                 const double* __restrict__ p_ref, int64_t N,
                                                                 double* partials, int32_t* done, double* scalars) {
  pdl_prologue();
  __shared__ double sm[32];
  __shared__ bool s_last;
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (v < N) {
    double a0, a1, a2, b0, b1, b2;
    load_pt<DIM>(p, v, a0, a1, a2);
    load_pt<DIM>(p_ref, v, b0, b1, b2);
    const double e0 = a0 - b0, e1 = a1 - b1, e2 = a2 - b2;
    d2 = e0 * e0 + e1 * e1;
    if (DIM == 3) d2 = d2 + e2 * e2;
    if (REL) {
      const double h = f.kind == DM_SIZE_GRID ? size_eval(f, a0, a1, a2) : f.hconst;
      d2 = d2 / (h * h);
    }
  }
  const double bm = block_max(d2, sm);
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
      scalars[5] = sqrt(r);
      *done = 0;
    }
  }
}

}  // namespace dm
