// Stage B, alternative design (round 2 experiment; OFF by default, DM_ROWS=1 selects it): one THREAD per
// vertex with a private hash set, instead of the 8-lane group per vertex of adjacency_kernel.
//
// A warp takes a TILE of 32 consecutive vertices, lane = vertex:
//   * the lane walks its own bucket (stage A's output) and inserts every id into its PRIVATE table
//     tab[slot * 32 + lane]: the bank of every access is the lane id, nothing is shared between lanes,
//     no atomic, no __syncwarp, no verify pass;
//   * the table column is compacted in place, sorted by a compile-time bitonic network on REGISTERS
//     (which also yields nlow = #{ids < v}), and the 32 rows leave as coalesced 16-B stores;
//   * the bar pass over the row's upper part follows per lane; rows with more than RS ids are built
//     by the whole warp (slow_row), bucket overflows by the heavy-vertex blocks, as before.
// Parity: the complete GPU suite passes with it (103 tests, gpurun_out r2b / r2d).
// Measured on B200 (ball h0 = 0.02, profiles/r2d_*): 84 M warp instructions against 142 M for the
// lane-group kernel, but only 21 of 64 warps resident (8 KB of table per warp, 80 registers) and
// IPC 0.41 against 0.72: 207 us against 191 us; gridded fh: 100 us against 56 us (EAGE-shaped), since the
// per-lane bar pass serialises seven trilinear lookups where the group spreads them over 8 lanes.
// The cost is SIMT divergence: a warp executes the probe / insert path of an id whenever ANY of its
// 32 lanes needs it (94 % of the id rounds), so the 80 % of ids that are duplicates save nothing.
// Kept as the record of that measurement; adjacency_kernel stays the product path.
#pragma once
#include "dm_pipeline.cuh"

namespace dm {

#ifndef DM_ROWS_WPB
#define DM_ROWS_WPB 4  // warps (= tiles of 32 vertices) per block
#endif
#ifndef DM_ROWS_MINB
#define DM_ROWS_MINB 6  // resident blocks per SM the register allocation is held to
#endif
constexpr int ROWS_WPB = DM_ROWS_WPB;
constexpr int ROWS_THREADS = 32 * ROWS_WPB;

template <int DIM>
struct RowsCfg {
  static constexpr int RS = PCfg<DIM>::RS;
  static constexpr int H = 2 * RS;            // private table slots: never more than RS + 1 keys -> load <= 1/2
  static constexpr int WARP_INTS = H * 32;    // tab[slot * 32 + lane]: 8 KB per warp in 3-D, 4 KB in 2-D
};

template <int NN>
__device__ __forceinline__ void bitonic_regs(int (&r)[NN]) {
#pragma unroll
  for (int k = 2; k <= NN; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < NN; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const int a = r[i], b = r[l];
          const int mn = min(a, b), mx = max(a, b);
          const bool up = (i & k) == 0;
          r[i] = up ? mn : mx;
          r[l] = up ? mx : mn;
        }
      }
    }
  }
}

// 16-byte read-only load of bucket words
__device__ __forceinline__ int4 ldg_int4(const int4* q) {
  int4 v;
  asm("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(q));
  return v;
}

// A vertex with more than RS distinct neighbours: the whole warp selects the ids in ascending order
// straight from the bucket (n entries of DIM ids), minimum by minimum, into the heap.
template <int DIM, int BAR>
__device__ __noinline__ void slow_row(int v, int n, const typename PCfg<DIM>::entry_t* __restrict__ brow, int64_t N,
                                      int32_t* __restrict__ adj, int32_t* __restrict__ heap, int2* __restrict__ degs,
                                      int32_t* __restrict__ counters, const DmSizeFn& f, const double* __restrict__ pp,
                                      double* __restrict__ hslot, int& bars, double& sL, double& sH) {
  constexpr int RS = PCfg<DIM>::RS;
  typedef typename PCfg<DIM>::entry_t entry_t;
  constexpr int EW = sizeof(entry_t) / 4;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(counters + 4, (int)round4(DIM * n));
  base = __shfl_sync(FULL, base, 0);
  const int* bi = reinterpret_cast<const int*>(brow);
  int last = -1, U = 0, lo = 0;
  for (;;) {
    int m = 0x7fffffff;
    for (int i = lane; i < DIM * n; i += 32) {
      const int x = bi[(i / DIM) * EW + i % DIM];
      if (x > last && x < m) m = x;
    }
    m = __reduce_min_sync(FULL, m);
    if (m == 0x7fffffff) break;
    if (lane == 0) heap[base + U] = m;
    ++U;
    lo += m < v ? 1 : 0;
    last = m;
  }
  __syncwarp();
  const int32_t* srt = heap + base;
  int32_t* row = adj + (int64_t)v * RS;
  int64_t sbase = (int64_t)v * RS;
  if (U <= RS) {  // (only when called for a table that merely looked full)
    for (int i = lane; i < U; i += 32) row[i] = srt[i];
  } else {
    if (lane == 0) row[0] = base;
    sbase = N * RS + base;
  }
  if (lane == 0) {
    degs[v] = make_int2(U, lo);
    bars += U - lo;
  }
  if (BAR >= 0 && U > (BAR == 1 ? 0 : lo)) {
    double a0, a1, a2;
    load_pt<DIM, true>(pp, v, a0, a1, a2);
    for (int j = (BAR == 1 ? 0 : lo) + lane; j < U; j += 32)
      bar_terms<DIM, BAR == 1>(f, pp, a0, a1, a2, srt[j], hslot + sbase + j, sL, sH, j >= lo);
  }
}

// Grid = HV_BLOCKS heavy-vertex blocks + one block per ROWS_WPB tiles of 32 vertices.  The main warps
// never synchronise with each other: a warp is a "block" of the reduction tree of final_scale
// (partials[w] per warp, groups of RG warps).
// Dynamic shared memory: max(ROWS_WPB * RowsCfg::WARP_INTS, HV_SMEM) ints.
template <int DIM, int BAR>
__global__ void __launch_bounds__(ROWS_THREADS, DM_ROWS_MINB) rows_kernel(
    const int32_t* __restrict__ cnt, const typename PCfg<DIM>::entry_t* __restrict__ bucket,
    const int32_t* __restrict__ ovf_v, const typename PCfg<DIM>::entry_t* __restrict__ ovf_e, int64_t N,
    int32_t* __restrict__ adj, int32_t* __restrict__ heap, int2* __restrict__ degs, const int32_t* __restrict__ hv,
    int32_t* __restrict__ counters, const __grid_constant__ DmSizeFn f, const double* __restrict__ pp,
    double* __restrict__ hslot, double* partials, int32_t* gdone, int32_t* total_done, double* scalars) {
  pdl_prologue();
  typedef RowsCfg<DIM> RC;
  constexpr int CAP = PCfg<DIM>::CAP, RS = RC::RS, H = RC::H;
  constexpr unsigned HM = H - 1;
  constexpr int LOGH = DIM == 3 ? 6 : 5;
  static_assert((1 << LOGH) == H, "table size");
  static_assert(HV_RANK % ROWS_THREADS == 0, "heavy_vertex: candidates per thread");
  extern __shared__ __align__(16) int32_t s_dyn[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t nbm = ((int64_t)gridDim.x - HV_BLOCKS) * ROWS_WPB;  // main WARPS: the leaves of the reduction tree
  const int ng = (int)((nbm + RG - 1) / RG);

  if (blockIdx.x < HV_BLOCKS) {  // ---------------- heavy-vertex block (round-1 code, block size as a parameter)
    __shared__ int s_scan[33];
    __shared__ double s_dbl[32];
    __shared__ int s_base, s_pos;
    __shared__ bool s_fin;
    const int nheavy = counters[3], novf = counters[2];
    for (int ih = blockIdx.x; ih < nheavy; ih += HV_BLOCKS)
      heavy_vertex<DIM, BAR, ROWS_THREADS>(ih, nheavy, novf, s_dyn, s_scan, s_dbl, &s_base, &s_pos, cnt, bucket, ovf_v,
                                           ovf_e, N, adj, heap, degs, hv, counters, f, pp, hslot,
                                           partials + 2 * (nbm + ng));
    if (BAR < 0) return;
    __syncthreads();
    if (tid == 0) s_fin = arrive_acq_rel(total_done) == ng + HV_BLOCKS - 1;
    __syncthreads();
    if (s_fin && tid < 32) final_scale<DIM>(partials, nbm, ng, nheavy, scalars);
    return;
  }

  // ---------------- main warp: tile of 32 vertices, lane = vertex
  const int64_t widx = ((int64_t)blockIdx.x - HV_BLOCKS) * ROWS_WPB + wid;
  int32_t* tab = s_dyn + wid * RC::WARP_INTS + lane;  // this lane's column: slot s at tab[s * 32]
  const int64_t v = widx * 32 + lane;
#pragma unroll
  for (int s = 0; s < H; ++s) tab[s * 32] = HASH_EMPTY;
  int n = v < N ? cnt[v] : 0;
  const bool heavy = n > CAP;  // overflowed bucket: a heavy-vertex block builds this row
  if (heavy) n = 0;

  // ---- insert every id of the bucket into the private set
  typedef typename PCfg<DIM>::entry_t entry_t;
  const entry_t* brow = bucket + (v < N ? v : 0) * CAP;
  constexpr int EPW = 16 / sizeof(entry_t);  // entries per 16-B word: 1 (3-D) or 2 (2-D)
  constexpr int NX = EPW * DIM;              // ids per word
  const int4* bw = reinterpret_cast<const int4*>(brow);
  const int nw = (n + EPW - 1) / EPW;
  const int maxw = __reduce_max_sync(FULL, nw);
  int k = 0;  // distinct ids so far (> RS: overflow, the warp builds the row afterwards)
  // one 16-B bucket word = NX ids.  About four ids in five are duplicates of an id that is already in
  // the set, most of them at the first slot of their probe sequence: the NX first probes are issued
  // together (independent loads), and only the ids they do not settle walk the probe loop.
  auto insert_word = [&](const int4& q, int i) {
    int x[NX];
    if constexpr (DIM == 3) {
      x[0] = q.x; x[1] = q.y; x[2] = q.z;
    } else {
      // the second entry of the word may lie past the end of an odd-length bucket: repeat the first
      const bool two = 2 * i + 1 < n;
      x[0] = q.x; x[1] = q.y;
      x[2] = two ? q.z : q.x;
      x[3] = two ? q.w : q.y;
    }
    unsigned h[NX];
    int y[NX];
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      h[c] = hash_slot<LOGH>(x[c]);
      y[c] = tab[h[c] * 32];
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      if (y[c] != x[c] && k <= RS) {
        unsigned hh = h[c];
        for (;;) {  // (re-reads the first slot: an earlier id of this word may have taken it)
          const int z = tab[hh * 32];
          if (z == x[c]) break;
          if (z == HASH_EMPTY) {
            tab[hh * 32] = x[c];
            ++k;
            break;
          }
          hh = (hh + 1u) & HM;
        }
      }
    }
  };
  // the bucket comes from DRAM / L2 (lane-strided 16-B loads): PF words in flight ahead of the hashing
  constexpr int PF = 4;
  int4 cur[PF], nxt[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) {
    cur[u] = make_int4(0, 0, 0, 0);
    nxt[u] = make_int4(0, 0, 0, 0);
    if (u < nw) cur[u] = ldg_int4(bw + u);
  }
  for (int i0 = 0; i0 < maxw; i0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (i0 + PF + u < nw) nxt[u] = ldg_int4(bw + i0 + PF + u);
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (i0 + u < nw) insert_word(cur[u], i0 + u);
#pragma unroll
    for (int u = 0; u < PF; ++u) cur[u] = nxt[u];
  }
  const bool over = k > RS;  // more than RS distinct neighbours

  // ---- compact the column in place (the write position never passes the read position) ...
  {
    int w = 0;
#pragma unroll 8
    for (int s = 0; s < H; ++s) {
      const int y = tab[s * 32];
      if (y != HASH_EMPTY) {
        tab[w * 32] = y;
        ++w;
      }
    }
  }
  // ---- ... sort it on registers, count the lower neighbours, put the sorted row back
  int lo = 0;
  {
    int r[RS];
#pragma unroll
    for (int j = 0; j < RS; ++j) r[j] = j < k ? tab[j * 32] : 0x7fffffff;
    bitonic_regs<RS>(r);
#pragma unroll
    for (int j = 0; j < RS; ++j) {
      tab[j * 32] = r[j];
      lo += r[j] < (int)v ? 1 : 0;
    }
  }
  const int U = (heavy || over || v >= N) ? 0 : k;
  const bool mine = v < N && !heavy && !over;
  if (mine) degs[v] = make_int2(U, lo);
  __syncwarp();

  // ---- the 32 rows leave as coalesced 16-B stores: G lanes per row, 32 / G rows per pass
  {
    constexpr int G = RS / 4;  // int4 per row: 8 (3-D) or 4 (2-D)
    const int lg = lane % G, sub = lane / G;
    const int32_t* col0 = tab - lane;  // column of lane 0
    const int64_t v0 = v - lane;
#pragma unroll
    for (int pass = 0; pass < G; ++pass) {
      const int u = pass * (32 / G) + sub;  // row (= lane of its owner) handled by this lane group
      const int Uu = __shfl_sync(FULL, U, u);
      if (4 * lg < Uu) {
        const int32_t* src = col0 + u + (4 * lg) * 32;
        reinterpret_cast<int4*>(adj + (v0 + u) * RS)[lg] = make_int4(src[0], src[32], src[64], src[96]);
      }
    }
  }

  // ---- bar pass over the upper neighbours (mesh_generator.py:696-700)
  int bars = mine ? U - lo : 0;
  double sL = 0.0, sH = 0.0;
  if (BAR >= 0 && mine && U > (BAR == 1 ? 0 : lo)) {
    double a0, a1, a2;
    load_pt<DIM, true>(pp, v, a0, a1, a2);
    const int64_t sbase = v * RS;
    for (int j = (BAR == 1 ? 0 : lo); j < U; ++j)
      bar_terms<DIM, BAR == 1>(f, pp, a0, a1, a2, tab[j * 32], hslot + sbase + j, sL, sH, j >= lo);
  }
  // ---- vertices with more than RS distinct neighbours: the whole warp, one vertex at a time
  unsigned todo = __ballot_sync(FULL, over && v < N && !heavy);
  while (todo) {
    const int u = __ffs(todo) - 1;
    todo &= todo - 1;
    const int nu = __shfl_sync(FULL, n, u);
    const int64_t vu = v - lane + u;
    int b2 = 0;
    double l2 = 0.0, h2 = 0.0;
    slow_row<DIM, BAR>((int)vu, nu, bucket + vu * CAP, N, adj, heap, degs, counters, f, pp, hslot, b2, l2, h2);
    bars += b2;
    sL += l2;
    sH += h2;
  }

  // ---- warp totals (fixed shuffle tree) -> partials[widx]; the warp that completes a group of RG warps
  //      adds the group in warp order, the last arrival overall adds groups + heavy vertices
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) bars += __shfl_down_sync(FULL, bars, d);
  if (BAR >= 0) {
    sL = warp_sum(sL);
    sH = warp_sum(sH);
  }
  int lead = 0;  // lane 0: 1 = this warp closes its reduction group
  if (lane == 0) {
    if (bars) atomicAdd(counters, bars);  // unique bars owned by this warp's vertices
    if (BAR >= 0) {
      partials[2 * widx] = sL;
      partials[2 * widx + 1] = sH;
      const int64_t g = widx / RG;
      const int gsize = (int)(nbm - g * RG < RG ? nbm - g * RG : RG);
      lead = arrive_acq_rel(gdone + g) == gsize - 1 ? 1 : 0;
    }
  }
  if (BAR < 0) return;
  if (!__shfl_sync(FULL, lead, 0)) return;
  {
    __syncwarp();
    const int64_t g = widx / RG;
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

}  // namespace dm
