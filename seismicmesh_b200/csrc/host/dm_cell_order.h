// Output order of the host triangulators: vertex ids ascending within a cell, cells in lexicographic
// order.  This is the order the device pipeline likes best (dm_pipeline.cuh, stage A): consecutive
// cells share their first vertices, so the warp-aggregated slot claims merge (2.4 atomics per cell
// instead of 3.7 in creation order on the ball h0 = 0.03 mesh) and the position gathers of a warp
// fall into few lines (10 distinct 128-B lines per warp instead of 58; Qhull's own order: 2.4 / 23).
// Orientation is therefore NOT normalised (the loop body never uses it; the reference fixes it once,
// at termination, in fix_mesh).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace dmx {

template <int K>
inline void order_cells(int32_t* cells, int64_t T, int64_t N) {
  if (T <= 0) return;
  for (int64_t c = 0; c < T; ++c) std::sort(cells + K * c, cells + K * c + K);
  // counting sort by the first (smallest) id, then a small sort inside every group
  std::vector<int64_t> start(N + 1, 0);
  for (int64_t c = 0; c < T; ++c) ++start[cells[K * c] + 1];
  for (int64_t v = 0; v < N; ++v) start[v + 1] += start[v];
  std::vector<int32_t> tmp((size_t)(K * T));
  {
    std::vector<int64_t> pos(start.begin(), start.end() - 1);
    for (int64_t c = 0; c < T; ++c) {
      const int64_t o = pos[cells[K * c]]++;
      for (int k = 0; k < K; ++k) tmp[K * o + k] = cells[K * c + k];
    }
  }
  struct Cell {
    int32_t v[K];
  };
  Cell* g = reinterpret_cast<Cell*>(tmp.data());
  for (int64_t v = 0; v < N; ++v)
    if (start[v + 1] - start[v] > 1)
      std::sort(g + start[v], g + start[v + 1], [](const Cell& a, const Cell& b) {
        for (int k = 1; k < K; ++k)
          if (a.v[k] != b.v[k]) return a.v[k] < b.v[k];
        return false;
      });
  std::copy(tmp.begin(), tmp.end(), cells);
}

}  // namespace dmx
