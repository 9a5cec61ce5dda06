// Stages A + B of the force iteration, fourth layout (round 2): records binned by TILES of 32 vertices.
//
// The third layout (dm_pipeline.cuh) pushes every kept cell into a bucket of each of its vertices (one global
// atomic per run of equal ids, ~65 per warp of 32 cells, and a 16-B entry store on ~128 different lines per
// warp) and builds every row with a lane group of its own (hash, compaction, rank sort, a reduction per 4
// vertices: 149 M warp instructions on the ball h0 = 0.02).  Here the unit of hand-over is a TILE of 32
// consecutive vertex ids:
//
//   A  cull_bin        one thread per TWO cells: centroid + fused SDF program -> keep flag (as before); each kept
//                      cell appends one 16-B RECORD per vertex -- {the other vertex ids, vertex & 31} -- to
//                      the record list of the vertex's tile.  List positions are claimed per (column, tile)
//                      with __match_any_sync: the host Delaunay codes emit cells grouped around vertices, so
//                      a warp of 32 cells touches ~5 tiles (~10 atomics and ~12 contiguous runs of stores per
//                      warp, against 65 and 128).  A full list spills to a global list.  The kernel turned out
//                      to be bound by the latency of its dependent chain, not by those accesses (DESIGN.md
//                      section 4): hence the two cells per thread.
//   B  tile_rows       one WARP per tile, no block barrier before the final totals.  The record batches
//                      stream into a shared-memory ring (cp.async, 4 batches ahead); every lane takes one
//                      record and inserts its ids into the hash set of the record's vertex -- 32 sets of 64
//                      slots per warp in shared memory, plain loads and stores in warp lockstep (probe; a lane
//                      that found the slot empty writes; warp barrier; only the writers verify).  Then lane l
//                      IS vertex l of the tile: it compacts its own set in place (odd stride: every lane on its
//                      own bank), sorts it with a compile-time bitonic network on 32 registers, counts its
//                      lower neighbours, and the warp writes the 32 rows with coalesced 128-B stores.  The bar
//                      pass (L, fh(midpoint), sum L^d, sum h^d) runs over the rows where they lie, a lane group
//                      per vertex, two passes in flight, and there is ONE reduction per 32 vertices.  Vertices
//                      with more than RS neighbours: a complete set (it holds up to 64) is sorted where it
//                      lies by the whole warp; an overflowed one is rebuilt exactly from the tile's records --
//                      gathered into the warp's shared memory (the heap for hubs), sorted, de-duplicated.
//
// Same outputs as the third layout (adj / heap / degs / hslot / counters[0] / scalars[0..2]); rows are sets,
// sorted, so they are bit-identical (tests/test_gpu_parity.py::test_tile_layout_equals_bucket_layout); the bar
// sums are added in a different (fixed) order.  95 M warp instructions on the same mesh, 156 us against 191 us.
#pragma once
#include <limits.h>

#include "dm_pipeline.cuh"

namespace dm {

#ifndef DM_TL_MINB
#define DM_TL_MINB 6
#endif
#ifndef DM_TL_WPB
#define DM_TL_WPB 4
#endif
#ifndef DM_TL_LOGH3
#define DM_TL_LOGH3 6
#endif
#ifndef DM_TL_LOGH2
#define DM_TL_LOGH2 5
#endif
#ifndef DM_TL_MAXSTEPS
#define DM_TL_MAXSTEPS 64
#endif

constexpr int TL_R = 32;  // vertices per tile = lanes of the warp that builds their rows
constexpr int TL_WPB = DM_TL_WPB;
constexpr int TL_THREADS = 32 * TL_WPB;

template <int DIM>
struct TCfg;
template <>
struct TCfg<3> {
  static constexpr int RS = PCfg<3>::RS;
  static constexpr int LOGH = DM_TL_LOGH3;  // slots of a vertex's hash set (mean degree 14)
  static constexpr int CAPT = 40 * TL_R;    // records per tile (mean 24.8 per vertex, worst tile of the ball 28)
  static constexpr int G = 8;               // lanes per vertex in the bar pass
};
template <>
struct TCfg<2> {
  static constexpr int RS = PCfg<2>::RS;
  static constexpr int LOGH = DM_TL_LOGH2;
  static constexpr int CAPT = 10 * TL_R;  // mean 6 triangles per vertex
  static constexpr int G = 4;
};
#ifndef DM_TL_STAGES
#define DM_TL_STAGES 4
#endif
constexpr int TL_STAGES = DM_TL_STAGES;  // record batches (32 x 16 B) in flight per warp (cp.async ring)
template <int DIM>
__host__ __device__ constexpr int tile_sets_ints() {
  return TL_R * ((1 << TCfg<DIM>::LOGH) + 1) + TL_R;  // 32 sets at an odd stride + one parking word per lane
}
template <int DIM>
__host__ __device__ constexpr int tile_warp_ints() {
  return tile_sets_ints<DIM>() + TL_STAGES * TL_R * 4;  // + the record ring
}

// 16-B asynchronous copy global -> shared (LDGSTS): the record batches of a tile stream into the warp's ring
// without passing through registers, TL_STAGES batches ahead of the one being hashed
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// A: cull + records.  mode as in cull_scatter_kernel.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ int4 record_of(const int (&ids)[4], int j) {
  if (DIM == 3) return make_int4(ids[j == 0 ? 1 : 0], ids[j <= 1 ? 2 : 1], ids[j <= 2 ? 3 : 2], ids[j] & (TL_R - 1));
  return make_int4(ids[j == 0 ? 1 : 0], ids[j <= 1 ? 2 : 1], 0, ids[j] & (TL_R - 1));
}

// CPT cells per thread (cells c, c + blockDim.x, ...): the kernel is bound by the LATENCY of its dependent
// chain -- cell ids (DRAM) -> position gathers (L2) -> SDF -> position claims (L2 atomics, hot counters) ->
// stores -- at the occupancy its registers allow, so every phase is written for all CPT cells of a thread
// before the next one starts: their chains are independent and overlap.
#ifndef DM_CB_CPT
#define DM_CB_CPT 2
#endif
#ifndef DM_CB_MINB
#define DM_CB_MINB 6
#endif
template <int DIM, bool PAD, int CPT>
__global__ void __launch_bounds__(DM_CS_THREADS, DM_CB_MINB) cull_bin_kernel(
    const double* __restrict__ prog, const double* __restrict__ p, const int32_t* __restrict__ t, int64_t T,
    double geps, int mode, uint8_t* __restrict__ keep, int32_t* __restrict__ tcnt, int4* __restrict__ trec,
    int32_t* __restrict__ ovf_v, int4* __restrict__ ovf_e, int32_t* __restrict__ counters, int n_rows) {
  pdl_prologue();
  constexpr int CAPT = TCfg<DIM>::CAPT;
  const int64_t c0 = (int64_t)blockIdx.x * (blockDim.x * CPT) + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool k[CPT];
  int ids[CPT][4];
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    const int64_t c = c0 + (int64_t)u * blockDim.x;
    k[u] = c < T;
    ids[u][0] = ids[u][1] = ids[u][2] = ids[u][3] = 0;
    if (k[u]) load_cell<DIM>(t, c, ids[u]);
  }
  if (mode == 0) {
    double cc[CPT][3];
#pragma unroll
    for (int u = 0; u < CPT; ++u)  // (a cell past the end gathers vertex 0: a valid address)
      cell_centroid<DIM, PAD>(p, ids[u], cc[u][0], cc[u][1], cc[u][2]);  // summed in the cell's own column order
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
      const int64_t c = c0 + (int64_t)u * blockDim.x;
      if (k[u]) {
        k[u] = sdf_eval(prog, DIM, cc[u][0], cc[u][1], cc[u][2]) < -geps;
        keep[c] = k[u] ? 1 : 0;
      }
    }
  } else if (mode == 1) {
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
      const int64_t c = c0 + (int64_t)u * blockDim.x;
      if (k[u]) k[u] = keep[c] != 0;
    }
  }
  // ids sorted in registers: column j of the lanes of a warp then holds ids of few tiles
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    auto cx = [](int& a, int& b) {
      const int lo_ = min(a, b), hi_ = max(a, b);
      a = lo_;
      b = hi_;
    };
    if (DIM == 3) {
      cx(ids[u][0], ids[u][1]);
      cx(ids[u][2], ids[u][3]);
      cx(ids[u][0], ids[u][2]);
      cx(ids[u][1], ids[u][3]);
      cx(ids[u][1], ids[u][2]);
    } else {
      cx(ids[u][0], ids[u][1]);
      cx(ids[u][1], ids[u][2]);
      cx(ids[u][0], ids[u][1]);
    }
  }
  // one list-position claim per (cell slot, column, tile) of the warp; all claims issued before any is waited on
  const unsigned lt = (1u << lane) - 1u;
  int base[CPT][DIM + 1], rl[CPT][DIM + 1];  // rl: rank | leader << 8
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
#pragma unroll
    for (int j = 0; j <= DIM; ++j) {
      // (vertices >= n_rows are ghost copies: nobody builds their rows, so they get no records)
      const bool kj = k[u] && ids[u][j] < n_rows;
      const int key = kj ? ids[u][j] / TL_R : -1 - lane;  // a lane without a record is alone in its class
      const unsigned m = __match_any_sync(FULL, key);
      const int leader = __ffs(m) - 1;
      rl[u][j] = __popc(m & lt) | (leader << 8);
      base[u][j] = 0;
      if (kj && lane == leader) base[u][j] = atomicAdd(tcnt + key, __popc(m));
    }
  }
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
#pragma unroll
    for (int j = 0; j <= DIM; ++j) {
      const int slot = __shfl_sync(FULL, base[u][j], rl[u][j] >> 8) + (rl[u][j] & 0xff);
      if (k[u] && ids[u][j] < n_rows) {
        const int tile = ids[u][j] / TL_R;
        const int4 rec = record_of<DIM>(ids[u], j);
        if (slot < CAPT) {
          trec[(int64_t)tile * CAPT + slot] = rec;
        } else {  // list full (hub vertices): the record goes to the global spill list with its tile
          const int o = atomicAdd(counters + 2, 1);
          ovf_v[o] = tile;
          ovf_e[o] = rec;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// B: rows of a tile
// ---------------------------------------------------------------------------------------------
// One record per lane, all 32 lanes in lockstep: the record's NC ids go into the set at tab[tb ...].  Slots
// only ever go EMPTY -> key.  A step: every unfinished key reads its slot; whoever finds its own key is done,
// whoever finds the slot empty writes its key; after the warp barrier only the writers look again (they
// either won the slot or lost it -- to an equal key: done, to another: next slot).  A finished key is parked
// on the lane's private word behind the sets (it "finds itself" there without conflicts from then on).
// No second barrier: a slot that is being verified was written before the barrier, so it is not EMPTY any
// more and nobody writes it in the next step.
template <int NC, int LOGH, bool PARTIAL>
__device__ __forceinline__ void tile_insert_lockstep(int32_t* tab, unsigned park, const int4& rec, bool valid,
                                                     unsigned& pmask) {
  // PARTIAL: some lanes of this batch hold no record (`valid` false; the last batch of a list, the spill scan)
  constexpr unsigned HM = (1u << LOGH) - 1u;
  constexpr int TS = (1 << LOGH) + 1;
  constexpr int MAXS = DM_TL_MAXSTEPS < (1 << LOGH) ? DM_TL_MAXSTEPS : (1 << LOGH);
  const unsigned tb = (unsigned)rec.w * TS;
  int x[NC];
  x[0] = rec.x;
  x[1] = rec.y;
  if (NC == 3) x[NC - 1] = rec.z;
  unsigned h[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    h[c] = tb + hash_slot<LOGH>(x[c]);
    if (PARTIAL && !valid) {
      h[c] = park;
      x[c] = HASH_PARKED;
    }
  }
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
      const unsigned nxt = tb + ((h[c] - tb + 1u) & HM);
      any = any | !hit;
      h[c] = hit ? park : nxt;
      x[c] = hit ? HASH_PARKED : x[c];
    }
  } while (__any_sync(FULL, any) && ++steps < MAXS);
  // a key that walked the whole set without finding a place (the loop only ends early when every lane is done):
  // the set is full, the vertex is rebuilt exactly afterwards
  if (any) pmask |= 1u << rec.w;
}

// ascending sort of RS registers: bitonic network, every index static
template <int RS>
__device__ __forceinline__ void sort_registers(int (&r)[RS]) {
#pragma unroll
  for (int k = 2; k <= RS; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < RS; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const int a = r[i], b = r[l];
          const bool up = (i & k) == 0;
          r[i] = up ? min(a, b) : max(a, b);
          r[l] = up ? max(a, b) : min(a, b);
        }
      }
    }
  }
}

// in place, any n, one warp; `a` in global or shared memory
__device__ __forceinline__ void warp_sort_ascending(int32_t* a, int n) {
  const int lane = threadIdx.x & 31;
  for (int k = 2; (k >> 1) < n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < n) {
          const int x = a[i], y = a[l];
          if (x > y) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
      __syncwarp();
    }
  }
}

// The records of one tile: its list and, if the list overflowed, the tile's entries of the global spill list.
struct TileRecords {
  const int4* recs;
  int n;  // records in the list
  const int32_t* ovf_v;
  const int4* ovf_e;
  int novf;  // spill records to scan (0 unless this tile's list overflowed)
  int tile;
};

// A vertex with more than RS neighbours (hull / hub vertices), handled by the whole warp.
//   nset >= 0: its hash set held them all -- `setp` (the vertex's own set memory, compacted) has the nset
//              distinct ids: sorted in place;
//   nset <  0: the set overflowed or a probe sequence got too long: the candidates are gathered again from
//              the tile's records into s_buf (the warp's shared-memory region, free by now) or, when they do
//              not fit, into the heap, sorted there and de-duplicated.
// Either way the row goes to the heap when it is longer than RS.
template <int DIM, int BAR>
__device__ __noinline__ void tile_big_vertex(const TileRecords& tr, int32_t* s_buf, int32_t* setp, int nset, int vloc,
                                             int64_t v, int64_t N, int32_t* __restrict__ adj, int32_t* __restrict__ heap,
                                             int2* __restrict__ degs, int32_t* __restrict__ counters,
                                             const DmSizeFn& f, const double* __restrict__ pp,
                                             double* __restrict__ hslot, int& bars, double& sL, double& sH) {
  constexpr int RS = TCfg<DIM>::RS;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int32_t* reg;
  int U = 0, lo = 0, base = 0;
  bool in_heap = false;
  if (nset >= 0) {
    reg = setp;
    U = nset;
    warp_sort_ascending(reg, U);
    for (int i0 = 0; i0 < U; i0 += 32) lo += __popc(__ballot_sync(FULL, i0 + lane < U && reg[i0 + lane] < (int)v));
    if (U > RS) {
      if (lane == 0) base = atomicAdd(counters + 4, (int)round4(U));
      base = __shfl_sync(FULL, base, 0);
    }
  } else {
    int nc = 0;
    for (int i0 = 0; i0 < tr.n; i0 += 32) {
      const int i = i0 + lane;
      nc += __popc(__ballot_sync(FULL, i < tr.n && tr.recs[i].w == vloc));
    }
    for (int i0 = 0; i0 < tr.novf; i0 += 32) {
      const int i = i0 + lane;
      nc += __popc(__ballot_sync(FULL, i < tr.novf && tr.ovf_v[i] == tr.tile && tr.ovf_e[i].w == vloc));
    }
    if (lane == 0) base = atomicAdd(counters + 4, (int)round4(DIM * nc));
    base = __shfl_sync(FULL, base, 0);
    const int n = DIM * nc;
    in_heap = n > TL_R * ((1 << TCfg<DIM>::LOGH) + 1);
    reg = in_heap ? heap + base : s_buf;
    int off = 0;
    for (int i0 = 0; i0 < tr.n; i0 += 32) {
      const int i = i0 + lane;
      int4 r = make_int4(0, 0, 0, -1);
      if (i < tr.n) r = tr.recs[i];
      const unsigned b = __ballot_sync(FULL, r.w == vloc);
      if (r.w == vloc) {
        int32_t* q = reg + (off + __popc(b & lt)) * DIM;
        q[0] = r.x;
        q[1] = r.y;
        if (DIM == 3) q[DIM - 1] = r.z;
      }
      off += __popc(b);
    }
    for (int i0 = 0; i0 < tr.novf; i0 += 32) {
      const int i = i0 + lane;
      int4 r = make_int4(0, 0, 0, -1);
      if (i < tr.novf && tr.ovf_v[i] == tr.tile) r = tr.ovf_e[i];
      const unsigned b = __ballot_sync(FULL, r.w == vloc);
      if (r.w == vloc) {
        int32_t* q = reg + (off + __popc(b & lt)) * DIM;
        q[0] = r.x;
        q[1] = r.y;
        if (DIM == 3) q[DIM - 1] = r.z;
      }
      off += __popc(b);
    }
    __syncwarp();
    warp_sort_ascending(reg, n);
    // unique, chunk by chunk: an element's output position never exceeds its input position and all reads
    // of a chunk complete before its writes, so compaction in place is safe
    for (int c0 = 0; c0 < n; c0 += 32) {
      const int i = c0 + lane;
      int x = 0;
      bool first = false;
      if (i < n) {
        x = reg[i];
        first = i == 0 || reg[i - 1] != x;
      }
      const unsigned b = __ballot_sync(FULL, first);
      lo += __popc(__ballot_sync(FULL, first && x < (int)v));
      __syncwarp();
      if (first) reg[U + __popc(b & lt)] = x;
      U += __popc(b);
      __syncwarp();
    }
  }
  int32_t* row = adj + v * RS;
  int64_t sbase = v * RS;
  if (U <= RS) {
    for (int i = lane; i < U; i += 32) row[i] = reg[i];
  } else {
    if (!in_heap)
      for (int i = lane; i < U; i += 32) heap[base + i] = reg[i];
    if (lane == 0) row[0] = base;
    sbase = N * RS + base;
  }
  if (lane == 0) {
    degs[v] = make_int2(U, lo);
    bars += U - lo;
  }
  if (BAR >= 0) {
    double a0, a1, a2;
    load_pt<DIM, true>(pp, v, a0, a1, a2);
    GridGuess gg;
    if (BAR == 1) gg = grid_guess(f);
    for (int j = (BAR == 1 && H_ALL_SLOTS ? 0 : lo) + lane; j < U; j += 32)
      bar_terms<DIM, BAR == 1>(f, gg, pp, a0, a1, a2, reg[j], hslot + sbase + j, sL, sH, j >= lo);
  }
  __syncwarp();  // the next vertex re-uses s_buf
}

// BAR as in adjacency_kernel.  Grid: one warp per tile of 32 vertices, TL_WPB warps per block.
template <int DIM, int BAR>
__global__ void __launch_bounds__(TL_THREADS, DM_TL_MINB) tile_rows_kernel(
    const int32_t* __restrict__ tcnt, const int4* __restrict__ trec, const int32_t* __restrict__ ovf_v,
    const int4* __restrict__ ovf_e, int64_t N, int64_t NR, int32_t* __restrict__ adj, int32_t* __restrict__ heap,
    int2* __restrict__ degs, int32_t* __restrict__ counters, const __grid_constant__ DmSizeFn f,
    const double* __restrict__ pp, double* __restrict__ hslot, double* partials, int32_t* gdone, int32_t* total_done,
    double* scalars) {
  pdl_prologue();
  constexpr int RS = TCfg<DIM>::RS, LOGH = TCfg<DIM>::LOGH, CAPT = TCfg<DIM>::CAPT, G = TCfg<DIM>::G;
  constexpr int H = 1 << LOGH, TS = H + 1;
  static_assert(H >= RS, "a set holds at least a fixed row");
  extern __shared__ __align__(16) int32_t s_dyn[];
  __shared__ double s_wL[TL_WPB], s_wH[TL_WPB];
  __shared__ int s_wbars[TL_WPB];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t nbm = gridDim.x;
  const int ng = (int)((nbm + RG - 1) / RG);
  const int64_t bidx = blockIdx.x;
  const int64_t ntiles = (NR + TL_R - 1) / TL_R;
  const int64_t tile = bidx * TL_WPB + wid;
  int32_t* tab = s_dyn + wid * tile_warp_ints<DIM>();
  const unsigned park = (unsigned)(TL_R * TS + lane);

  int bars = 0;
  double sL = 0.0, sH = 0.0;
  if (tile < ntiles) {  // warp-uniform
    for (int i = lane; i < TL_R * TS; i += 32) tab[i] = HASH_EMPTY;
    tab[park] = HASH_PARKED;
    __syncwarp();
    TileRecords tr;
    const int cnt = tcnt[tile];
    tr.recs = trec + tile * CAPT;
    tr.n = cnt < CAPT ? cnt : CAPT;
    tr.ovf_v = ovf_v;
    tr.ovf_e = ovf_e;
    tr.novf = cnt > CAPT ? counters[2] : 0;
    tr.tile = (int)tile;

    // ---- every record's ids into the set of the record's vertex; the next record is loaded before
    //      this one is hashed (the lists come from DRAM / L2)
    unsigned pmask = 0u;
    {
      // (the copies do not wait for the count: all CAPT records of a list are valid memory, what lies past
      //  the count is simply not used).  Every lane copies and reads back its own 16 bytes.
      static_assert(tile_sets_ints<DIM>() % 4 == 0, "ring alignment");
      int4* ring = reinterpret_cast<int4*>(tab + tile_sets_ints<DIM>());
      const int n = tr.n;
#pragma unroll
      for (int sg = 0; sg < TL_STAGES; ++sg) {
        cp_async16(ring + sg * TL_R + lane, tr.recs + (sg * TL_R + lane < CAPT ? sg * TL_R + lane : lane));
        cp_async_commit();
      }
      int slot = 0;
      for (int b = 0; b < n; b += 32) {
        cp_async_wait<TL_STAGES - 1>();
        const int4 cur = ring[slot * TL_R + lane];
        if (b + 32 <= n)
          tile_insert_lockstep<DIM, LOGH, false>(tab, park, cur, true, pmask);
        else
          tile_insert_lockstep<DIM, LOGH, true>(tab, park, cur, b + lane < n, pmask);
        const int ia = b + TL_STAGES * TL_R + lane;
        cp_async16(ring + slot * TL_R + lane, tr.recs + (ia < CAPT ? ia : lane));
        cp_async_commit();
        slot = slot + 1 == TL_STAGES ? 0 : slot + 1;
      }
      cp_async_wait<0>();
      for (int i0 = 0; i0 < tr.novf; i0 += 32) {  // this tile's list overflowed: its spill records
        const int i = i0 + lane;
        const bool mine = i < tr.novf && ovf_v[i] == tr.tile;
        if (!__any_sync(FULL, mine)) continue;
        int4 r = make_int4(0, 0, 0, 0);
        if (mine) r = ovf_e[i];
        tile_insert_lockstep<DIM, LOGH, true>(tab, park, r, mine, pmask);
      }
    }
    __syncwarp();
    const unsigned punted = __reduce_or_sync(FULL, pmask);

    // ---- lane l is vertex l of the tile: compact its set in place (writes never pass the reads), sort it
    const int64_t v0 = tile * TL_R;
    const int64_t v = v0 + lane;
    int32_t* my = tab + lane * TS;
    int U = 0;
#pragma unroll
    for (int s0 = 0; s0 < H; s0 += 8) {
      int q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = my[s0 + u];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (q[u] != HASH_EMPTY) my[U++] = q[u];
    }
    // more than RS neighbours (hull / hub vertices) or an incomplete set: the warp deals with them at the end
    const bool failed = ((punted >> lane) & 1u) != 0u;
    const bool over = U > RS || failed;
    const int Uset = U;
    if (over || v >= NR) U = 0;
    int r[RS];
#pragma unroll
    for (int i = 0; i < RS; ++i) r[i] = i < U ? my[i] : INT_MAX;
    sort_registers<RS>(r);
    int lo = 0;
#pragma unroll
    for (int i = 0; i < RS; ++i) {
      lo += r[i] < (int)v ? 1 : 0;
      if (!over) my[i] = r[i];  // (a big vertex keeps its compacted set where it is)
    }
    if (v < NR && !over) {
      degs[v] = make_int2(U, lo);
      bars = U - lo;
    }
    __syncwarp();
    // ---- the rows of the tile, 128 B per store instruction
    {
      constexpr int RPI = 32 / RS;  // rows per instruction
#pragma unroll 4
      for (int k0 = 0; k0 < TL_R; k0 += RPI) {
        const int k = k0 + lane / RS, col = lane % RS;
        const int Uk = __shfl_sync(FULL, U, k);
        if (col < ((Uk + 3) & ~3)) adj[(v0 + k) * RS + col] = tab[k * TS + col];
      }
    }
    // ---- bar pass (mesh_generator.py:696-700): L^d, h^d of the rows' upper bars; gridded fh: h of EVERY slot
    if (BAR >= 0) {
      // a lane group per vertex, VPP vertices per pass, two passes in flight (their position gathers are
      // independent; the pass is bound by their latency)
      constexpr int VPP = 32 / G;
      constexpr int NPS = TL_R / VPP;
      static_assert(NPS % 2 == 0, "passes are taken two at a time");
      const int lg = lane % G;
      GridGuess gg;
      if (BAR == 1) gg = grid_guess(f);
#pragma unroll 1
      for (int ps = 0; ps < NPS; ps += 2) {
        int vl[2], Uv[2], lov[2], j0[2], w[2];
        double a[2][3], b[2][3];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          vl[u] = (ps + u) * VPP + lane / G;
          Uv[u] = __shfl_sync(FULL, U, vl[u]);
          lov[u] = __shfl_sync(FULL, lo, vl[u]);
          j0[u] = ((BAR == 1 && H_ALL_SLOTS) ? 0 : lov[u]) + lg;
          // a lane without a bar gathers the vertex itself (a valid address; the term is discarded)
          w[u] = j0[u] < Uv[u] ? tab[vl[u] * TS + j0[u]] : (int)(v0 + vl[u] < N ? v0 + vl[u] : 0);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          load_pt<DIM, true>(pp, v0 + vl[u] < N ? v0 + vl[u] : 0, a[u][0], a[u][1], a[u][2]);
          load_pt<DIM, true>(pp, w[u], b[u][0], b[u][1], b[u][2]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (j0[u] < Uv[u]) {
            const int64_t vv = v0 + vl[u];
            bar_terms_at<DIM, BAR == 1>(f, gg, a[u][0], a[u][1], a[u][2], b[u][0], b[u][1], b[u][2], hslot + vv * RS + j0[u], sL,
                                        sH, j0[u] >= lov[u]);
            for (int j = j0[u] + G; j < Uv[u]; j += G)  // rows longer than the lane group (rare in 3-D)
              bar_terms<DIM, BAR == 1>(f, gg, pp, a[u][0], a[u][1], a[u][2], tab[vl[u] * TS + j], hslot + vv * RS + j, sL, sH,
                                       j >= lov[u]);
          }
        }
      }
    }
    // ---- the vertices whose set did not fit, one after the other
    // ---- the vertices with more than RS neighbours, one after the other: first those whose set is complete
    //      (sorted where it lies), then those that have to be rebuilt from the records (in the sets' memory)
    unsigned todo = __ballot_sync(FULL, over && v < NR);
    if (todo) {
      const unsigned rebuilt = __ballot_sync(FULL, failed);
      __syncwarp();  // the bar pass has read the rows
      for (unsigned m = todo & ~rebuilt; m; m &= m - 1) {
        const int vl = __ffs(m) - 1;
        tile_big_vertex<DIM, BAR>(tr, tab, tab + vl * TS, __shfl_sync(FULL, Uset, vl), vl, v0 + vl, N, adj, heap, degs, counters,
                                  f, pp, hslot, bars, sL, sH);
      }
      for (unsigned m = todo & rebuilt; m; m &= m - 1) {
        const int vl = __ffs(m) - 1;
        tile_big_vertex<DIM, BAR>(tr, tab, nullptr, -1, vl, v0 + vl, N, adj, heap, degs, counters, f, pp, hslot, bars, sL, sH);
      }
    }
  }

  // ---- block totals in warp order, then the two-level fixed-order reduction of adjacency_kernel
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) bars += __shfl_down_sync(FULL, bars, d);
  if (BAR >= 0) {
    sL = warp_sum(sL);
    sH = warp_sum(sH);
  }
  if (lane == 0) {
    s_wbars[wid] = bars;
    if (BAR >= 0) {
      s_wL[wid] = sL;
      s_wH[wid] = sH;
    }
  }
  __syncthreads();
  if (wid != 0) return;
  int lead = 0;
  if (lane == 0) {
    int tb = 0;
    double tL = 0.0, tH = 0.0;
#pragma unroll
    for (int w = 0; w < TL_WPB; ++w) {
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
  if (BAR < 0) return;
  if (!__shfl_sync(FULL, lead, 0)) return;
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
      fin = arrive_acq_rel(total_done) == ng - 1 ? 1 : 0;
    }
    if (__shfl_sync(FULL, fin, 0)) {
      __syncwarp();
      final_scale<DIM>(partials, nbm, ng, 0, scalars);
    }
  }
}

}  // namespace dm
