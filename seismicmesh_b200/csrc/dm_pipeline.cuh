// The force-iteration pipeline (stages A-D of include/distmesh_b200.h), third layout.
//
// What bounds these kernels on B200 is not DRAM bytes but the number of distinct 128-B lines the
// LSU / L2 have to touch: every scattered 4..16-B access (gather, store or atomic) costs about one
// L1 wavefront (~1.3 cycles per lane per SM, measured), no matter how few bytes it moves.  The
// layout is therefore chosen to minimise scattered accesses per cell and to make every per-vertex
// structure one aligned 128-B line (or a few), read / written by a lane group in one wavefront:
//
//   A  cull_scatter      one thread per cell: centroid + fused SDF program -> keep flag; each kept
//                        cell appends its OTHER vertex ids to the fixed-capacity bucket of each of
//                        its vertices.  Slots are claimed with warp-aggregated atomics
//                        (match.any: host Delaunay codes emit cells grouped around vertices, so a
//                        warp's 128 claims collapse to ~50 atomics); a full bucket spills to a
//                        global overflow list.
//   B  adjacency         one lane group (8 lanes in 3-D, 4 in 2-D) per vertex: coalesced read of
//                        the bucket, de-duplication in a group-private shared-memory hash (plain
//                        LDS/STS, write-then-verify, no shared atomics), rank sort, ONE 128-B
//                        store of the sorted neighbour row.  Row = [lower neighbours ascending |
//                        upper neighbours ascending]; the upper parts of all rows, in vertex
//                        order, ARE the reference's sorted unique bar list (unique_edges,
//                        geometry/cpp/fast_geometry.cpp:30-77).
//      adjacency_heavy   vertices whose bucket overflowed or whose neighbour set does not fit the
//                        group hash (hull / hub vertices): one block per vertex, sort + unique.
//   C  bar_pass          L, fh(midpoint), sum L^d, sum h^d over unique bars, fixed-order reduction,
//                        scale = ((sum L^d)/(sum h^d))^(1/d)   (last block finishes the reduction)
//   D  vertex_update     per-vertex gather of bar forces in the reference's COO accumulation order,
//                        pfix, p += dt*F, Newton projection per level, max|F| (last block reduces)
#pragma once
#include "dm_device.cuh"

namespace dm {

constexpr int PL_THREADS = 256;  // cull / bar pass / vertex update
constexpr int AB_THREADS = 256;  // adjacency: 256 / G vertices per block
constexpr int HV_THREADS = 256;  // heavy-vertex path: one block per vertex
constexpr int HV_BLOCKS = 296;   // fixed grid (2 per SM); loops over the heavy list
constexpr int HV_SMEM = 6144;    // candidates sorted in shared memory up to this many

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

template <int DIM, bool AGG>
__global__ void __launch_bounds__(PL_THREADS) cull_scatter_kernel(
    const double* __restrict__ prog, const double* __restrict__ p, const int32_t* __restrict__ t, int64_t T,
    double geps, int mode, uint8_t* __restrict__ keep, int32_t* __restrict__ cnt,
    typename PCfg<DIM>::entry_t* __restrict__ bucket, int32_t* __restrict__ ovf_v,
    typename PCfg<DIM>::entry_t* __restrict__ ovf_e, int32_t* __restrict__ counters) {
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
      cell_centroid<DIM>(p, ids, c0, c1, c2);
      k = sdf_eval(prog, DIM, c0, c1, c2) < -geps;
      keep[c] = k ? 1 : 0;
    } else if (mode == 1) {
      k = keep[c] != 0;
    }
  }
  if (cnt != nullptr) {
    const unsigned km = __ballot_sync(FULL, k);
    if (k) {
#pragma unroll
      for (int j = 0; j <= DIM; ++j) {
        const int v = ids[j];
        int slot;
        if (AGG) {
          // all lanes that claim a slot of the same vertex in this step share one atomic
          const unsigned m = __match_any_sync(km, v);
          const int leader = __ffs(m) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(cnt + v, __popc(m));
          base = __shfl_sync(m, base, leader);
          slot = base + __popc(m & ((1u << lane) - 1u));
        } else {
          slot = atomicAdd(cnt + v, 1);
        }
        const typename PCfg<DIM>::entry_t e = others_of<DIM>(ids, j);
        if (slot < CAP) {
          bucket[(int64_t)v * CAP + slot] = e;
        } else {  // bucket full: spill (hull / hub vertices)
          const int o = atomicAdd(counters + 2, 1);
          ovf_v[o] = v;
          ovf_e[o] = e;
        }
      }
    }
  }
  if (counters != nullptr) {
    const int nk = __syncthreads_count(k);
    if (threadIdx.x == 0 && nk) atomicAdd(counters + 1, nk);
  }
}

// ---------------------------------------------------------------------------------------------
// B: adjacency rows
// ---------------------------------------------------------------------------------------------
constexpr int HASH_EMPTY = -1;
constexpr int HASH_MAXSTEPS = 12;  // probe sequence longer than this -> vertex goes the heavy way

template <int LOGH>
__device__ __forceinline__ unsigned hash_slot(int x) {
  return ((unsigned)x * 2654435761u) >> (32 - LOGH);
}

// One candidate per lane, all 32 lanes in lockstep.  Group-private open-addressing set in shared
// memory with plain loads / stores: a lane that finds its slot empty writes its key, the warp
// synchronises, and every lane re-reads the slot: whoever finds its own key there is done (it
// either won the slot or lost it to an equal key), everyone else moves to the next slot.  Slots
// only ever go EMPTY -> key, so a key is stored at the first slot of its probe sequence that was
// empty when it arrived and later equal keys find it before they find an empty slot.
template <int NC, int LOGH>
__device__ __forceinline__ void hash_insert_lockstep(int32_t* tab, const int (&x)[NC], bool act, bool& punt) {
  constexpr unsigned HM = (1u << LOGH) - 1u;
  unsigned h[NC];
  bool pend[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    h[c] = hash_slot<LOGH>(x[c]);
    pend[c] = act;
  }
  int steps = 0;
  while (true) {
    bool any = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) any = any || pend[c];
    if (!__any_sync(FULL, any)) break;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (pend[c] && tab[h[c]] == HASH_EMPTY) tab[h[c]] = x[c];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (pend[c]) {
        if (tab[h[c]] == x[c])
          pend[c] = false;
        else
          h[c] = (h[c] + 1u) & HM;
      }
    __syncwarp();
    if (++steps > HASH_MAXSTEPS) {  // table (nearly) full: give up, the heavy path takes the vertex
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        punt = punt || pend[c];
        pend[c] = false;
      }
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(AB_THREADS) adjacency_kernel(const int32_t* __restrict__ cnt,
                                                               const typename PCfg<DIM>::entry_t* __restrict__ bucket,
                                                               int64_t N, int32_t* __restrict__ adj,
                                                               int32_t* __restrict__ heap, int2* __restrict__ degs,
                                                               int32_t* __restrict__ hv, int32_t* __restrict__ counters) {
  constexpr int CAP = PCfg<DIM>::CAP, RS = PCfg<DIM>::RS, G = PCfg<DIM>::G, LOGH = PCfg<DIM>::LOGH;
  constexpr int H = 1 << LOGH;
  constexpr int VPB = AB_THREADS / G;  // vertices per block
  constexpr int SPL = H / G;           // table slots per lane (8)
  static_assert(SPL == 8, "compaction below reads two int4 per lane");
  // group stride H + 32/(32/G) ints: the groups of one warp start in different banks, so the
  // broadcast reads of the rank sort (same k in every group) do not conflict
  constexpr int GS = H + 32 / (32 / G);
  static_assert((GS * 4) % 16 == 0, "int4 access");
  __shared__ __align__(16) int32_t s_tab[VPB * GS];
  __shared__ __align__(16) int32_t s_lst[VPB * GS];
  __shared__ int s_red[33];
  const int tid = threadIdx.x, lane = tid & 31;
  const int lg = tid % G, grp = tid / G;
  const int64_t v = (int64_t)blockIdx.x * VPB + grp;
  int32_t* tab = s_tab + grp * GS;
  int32_t* lst = s_lst + grp * GS;

  // ---- empty table
  {
    int4* t4 = reinterpret_cast<int4*>(tab + lg * SPL);
    t4[0] = make_int4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY);
    t4[1] = make_int4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY);
  }
  int n = v < N ? cnt[v] : 0;
  bool punt = n > CAP;  // overflowed bucket: part of the star lives in the spill list
  if (punt) n = 0;
  __syncwarp();

  // ---- insert every candidate of the bucket: lane lg takes entries lg, lg+G, ...
  const typename PCfg<DIM>::entry_t* brow = bucket + v * CAP;
  for (int i = lg; __any_sync(FULL, i < n); i += G) {
    const bool act = i < n;
    int x[DIM];
    if (act) {
      const typename PCfg<DIM>::entry_t e = brow[i];
      x[0] = e.x;
      x[1] = e.y;
      if (DIM == 3) x[DIM - 1] = reinterpret_cast<const int*>(&e)[DIM - 1];
    } else {
#pragma unroll
      for (int c = 0; c < DIM; ++c) x[c] = 0;
    }
    hash_insert_lockstep<DIM, LOGH>(tab, x, act, punt);
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
  if (punt) U = 0;
  {
    int off = inc - c;
#pragma unroll
    for (int k = 0; k < SPL; ++k)
      if (vals[k] != HASH_EMPTY) lst[off++] = vals[k];
  }
  __syncwarp();

  // ---- rank sort: lst (unsorted, unique) -> tab (ascending)
  int lo = 0;
  for (int i = lg; i < U; i += G) {
    const int x = lst[i];
    int r = 0;
    for (int k = 0; k < U; ++k) r += lst[k] < x ? 1 : 0;
    tab[r] = x;
    lo += x < (int)v ? 1 : 0;
  }
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) lo += __shfl_xor_sync(FULL, lo, d, G);
  __syncwarp();

  // ---- write the row
  int bars = 0;
  if (v < N) {
    if (punt) {
      if (lg == 0) hv[atomicAdd(counters + 3, 1)] = (int32_t)v;
    } else {
      int32_t* row = adj + v * RS;
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
      }
      if (lg == 0) {
        degs[v] = make_int2(U, lo);
        bars = U - lo;
      }
    }
  }
  int total;
  block_exclusive_scan(bars, total, s_red);  // unique bars owned by this block's vertices
  if (tid == 0 && total) atomicAdd(counters, total);
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

template <int DIM>
__global__ void __launch_bounds__(HV_THREADS) adjacency_heavy_kernel(
    const int32_t* __restrict__ cnt, const typename PCfg<DIM>::entry_t* __restrict__ bucket,
    const int32_t* __restrict__ ovf_v, const typename PCfg<DIM>::entry_t* __restrict__ ovf_e, int64_t N,
    int32_t* __restrict__ adj, int32_t* __restrict__ heap, int2* __restrict__ degs, const int32_t* __restrict__ hv,
    int32_t* __restrict__ counters) {
  constexpr int CAP = PCfg<DIM>::CAP, RS = PCfg<DIM>::RS;
  __shared__ int32_t s_val[HV_SMEM];
  __shared__ int s_scan[33];
  __shared__ int s_base, s_pos;
  const int tid = threadIdx.x;
  const int nheavy = counters[3];
  const int novf = counters[2];
  for (int ih = blockIdx.x; ih < nheavy; ih += gridDim.x) {
    const int v = hv[ih];
    const int nc = cnt[v];
    const int nb = nc < CAP ? nc : CAP;  // entries in the bucket; the other nc - nb were spilled
    const int n = DIM * nc;              // candidates
    __syncthreads();
    if (tid == 0) {
      s_base = atomicAdd(counters + 4, (int)round4(n));
      s_pos = 0;
    }
    __syncthreads();
    int32_t* reg = heap + s_base;  // candidates, later the row, live here
    const bool in_smem = n <= HV_SMEM;
#ifdef DM_DEBUG_HEAVY
    if (tid == 0) printf("heavy ih=%d v=%d nc=%d nb=%d n=%d base=%d novf=%d nheavy=%d\n", ih, v, nc, nb, n, s_base, novf, nheavy);
#endif
    // gather the candidates
    for (int i = tid; i < nb; i += HV_THREADS) {
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
    if (nc > CAP) {
      for (int k = tid; k < novf; k += HV_THREADS) {
        if (ovf_v[k] == v) {
          const int i = nb + atomicAdd(&s_pos, 1);
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
    }
    __syncthreads();
    if (in_smem)
      block_sort_ascending(s_val, n);
    else
      block_sort_ascending(reg, n);
#ifdef DM_DEBUG_HEAVY
    if (tid == 0) printf("heavy ih=%d sorted\n", ih);
#endif
    // unique: chunk by chunk; an element's output position never exceeds its input position and
    // all reads of a chunk complete before its writes, so compaction in place is safe
    int U = 0, lo = 0;
    for (int c0 = 0; c0 < n; c0 += HV_THREADS) {
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
    __syncthreads();
    int32_t* row = adj + (int64_t)v * RS;
    if (U <= RS) {
      for (int i = tid; i < U; i += HV_THREADS) row[i] = reg[i];
    } else if (tid == 0) {
      row[0] = s_base;
    }
    if (tid == 0) {
      degs[v] = make_int2(U, lo);
      atomicAdd(counters, U - lo);
#ifdef DM_DEBUG_HEAVY
      printf("heavy ih=%d done U=%d lo=%d\n", ih, U, lo);
#endif
    }
  }
}

// bar ids: rowptr[v] = #upper neighbours (scanned afterwards)
__global__ void upper_count_kernel(const int2* __restrict__ degs, int64_t N, int32_t* __restrict__ rowptr) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) {
    const int2 d = degs[v];
    rowptr[v] = d.x - d.y;
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
// C: bar pass.  HMODE 0: constant h ; 1: gridded fh (h stored at the upper directed slot) ;
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
    if (m > lo) {
      const int32_t* row = R.row(v, m);
      const int64_t base = R.slot_base(v, m);
      double a0, a1, a2;
      load_pt<DIM>(p, v, a0, a1, a2);
      // rows are 16-B aligned and padded to a multiple of 4 ints: walk them in int4 chunks
      for (int j0 = lo & ~3; j0 < m; j0 += 4) {
        const int4 q = *reinterpret_cast<const int4*>(row + j0);
        const int wq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j < lo || j >= m) continue;
          const int w = wq[u];
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
    const DmSizeFn f, const double* __restrict__ p, double* __restrict__ p_out, const Rows<DIM> R,
    const int32_t* __restrict__ rowptr, const double* __restrict__ hslot, const double* __restrict__ hbar,
    const double* scalars_in, int64_t N, Levels lv, double L0mult, double delta_t, double deps, double h0,
    int64_t nfix, const uint8_t* __restrict__ fixed, double* __restrict__ Ftot, double* partials, int32_t* done,
    double* scalars) {
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
    const int2 dg = R.degs[v];
    const int lo = dg.y, m = dg.x;
    const int32_t* row = R.row(v, m);
    const int64_t base = R.slot_base(v, m);
    // The reference accumulates bar by bar (coo_matrix.toarray): row v receives -Fvec of its
    // lower bars (u,v), u ascending, then +Fvec of its upper bars (v,w), w ascending.  The row is
    // sorted, and -(F/L*(p[u]-p[v])) == (F/L)*(p[v]-p[u]) exactly, so one ascending sweep with
    // d = p[v]-p[nbr] reproduces the reference's sum bit for bit.
    for (int j0 = 0; j0 < m; j0 += 4) {
      const int4 q = *reinterpret_cast<const int4*>(row + j0);
      const int wq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j >= m) break;
        const int w = wq[u];
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
          const int e = j >= lo ? rowptr[v] + (j - lo) : bar_id_of<DIM>(R, rowptr, w, (int)v);
          h = hbar[e];
        }
        double Fs = h * k0 * scale - L;  // L0 - L  (mesh_generator.py:700-702)
        if (Fs < 0) Fs = 0;
        const double qf = Fs / L;
        F0 = F0 + qf * d0;
        F1 = F1 + qf * d1;
        if (DIM == 3) F2 = F2 + qf * d2;
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
