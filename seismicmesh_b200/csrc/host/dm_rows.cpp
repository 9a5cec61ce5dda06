// Sorted unique rows of an integer table behind include/distmesh_host.h -- what the termination path does to
// cell / facet / edge lists again and again: `geometry.unique_rows` of the row-wise sorted list
// (SeismicMesh/geometry/utils.py:141-172, called from fix_mesh :204-246, get_boundary_edges / _facets
// :310-361 and through them from get_boundary_vertices :364-382, i.e. from every sliver_removal pass with
// preserve=True).  The reference sorts byte views of the rows; NumPy's lexsort of 1.6 M facets was half of
// the 3-D termination time here.  Same machinery as the triangulators' output order (dm_cell_order.h):
// bucket the rows by their smallest id in two cache-friendly passes, sort the few rows of every id, then
// collapse equal neighbours.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "distmesh_host.h"
#include "dm_cell_order.h"

namespace {

// a row without its smallest id: (s1, s2) in `hi`, s3 in `lo` (zero where the row is shorter)
struct Rest {
  uint64_t hi, lo;
};
inline bool operator<(const Rest& a, const Rest& b) { return a.hi != b.hi ? a.hi < b.hi : a.lo < b.lo; }

template <int K>
int64_t sort_unique(int32_t* rows, int64_t n, int64_t N, int nth, int32_t* counts) {
  if (n < 100000) nth = 1;
  int shift = 0;
  while ((N >> shift) > 1024) ++shift;
  const int64_t B = (N >> shift) + 1;  // buckets of consecutive smallest ids, small enough for the cache
  std::unique_ptr<Rest[]> k1(new Rest[(size_t)n]), k2(new Rest[(size_t)n]);
  std::unique_ptr<int32_t[]> first(new int32_t[(size_t)n]);  // smallest id of the row at each slot of k1
  std::vector<int64_t> start(N + 1, 0);               // rows per smallest id, then the first slot of its group
  std::vector<int64_t> cursor((size_t)(nth * B), 0);  // rows per (thread, bucket), then the thread's write position
  dmx::order_threads(nth, [&](int j) {
    int64_t* mine = cursor.data() + (size_t)j * B;
    for (int64_t r = n * j / nth; r < n * (j + 1) / nth; ++r) {
      std::sort(rows + K * r, rows + K * r + K);
      const int32_t m = rows[K * r];
      ++mine[m >> shift];
      if (nth > 1)
        __atomic_fetch_add(&start[m], (int64_t)1, __ATOMIC_RELAXED);
      else
        ++start[m];
    }
  });
  {
    int64_t at = 0;
    for (int64_t v = 0; v <= N; ++v) {
      const int64_t k = start[v];
      start[v] = at;
      at += k;
    }
    for (int64_t b = 0; b < B; ++b) {
      int64_t pos = start[std::min<int64_t>(N, b << shift)];
      for (int j = 0; j < nth; ++j) {
        const int64_t k = cursor[(size_t)j * B + b];
        cursor[(size_t)j * B + b] = pos;
        pos += k;
      }
    }
  }
  dmx::order_threads(nth, [&](int j) {
    int64_t* mine = cursor.data() + (size_t)j * B;
    for (int64_t r = n * j / nth; r < n * (j + 1) / nth; ++r) {
      const int32_t* s = rows + K * r;
      Rest key;
      key.hi = K >= 3 ? ((uint64_t)(uint32_t)s[1] << 32 | (uint32_t)s[2]) : (uint64_t)(uint32_t)s[1];
      key.lo = K >= 4 ? (uint64_t)(uint32_t)s[K - 1] : 0;
      const int64_t at = mine[s[0] >> shift]++;
      k1[(size_t)at] = key;
      first[(size_t)at] = s[0];
    }
  });
  std::vector<int64_t> bcut(nth + 1, B);
  bcut[0] = 0;
  {
    int j = 1;
    for (int64_t b = 0; b < B && j < nth; ++b)
      while (j < nth && start[std::min<int64_t>(N, b << shift)] >= n * j / nth) bcut[j++] = b;
  }
  dmx::order_threads(nth, [&](int j) {
    std::vector<int64_t> pos;
    for (int64_t b = bcut[j]; b < bcut[j + 1]; ++b) {
      const int64_t v0 = std::min<int64_t>(N, b << shift), v1 = std::min<int64_t>(N, (b + 1) << shift);
      pos.assign(start.begin() + v0, start.begin() + v1);
      for (int64_t c = start[v0]; c < start[v1]; ++c) k2[(size_t)pos[first[(size_t)c] - v0]++] = k1[(size_t)c];
      for (int64_t v = v0; v < v1; ++v) {
        if (start[v + 1] - start[v] > 1) std::sort(k2.get() + start[v], k2.get() + start[v + 1]);
        for (int64_t c = start[v]; c < start[v + 1]; ++c) {
          int32_t* out = rows + K * c;
          const Rest& key = k2[(size_t)c];
          out[0] = (int32_t)v;
          if (K >= 3) {
            out[1] = (int32_t)(key.hi >> 32);
            out[2] = (int32_t)(key.hi & 0xffffffffu);
          } else {
            out[1] = (int32_t)key.hi;
          }
          if (K >= 4) out[3] = (int32_t)key.lo;
        }
      }
    }
  });
  int64_t m = 0;
  for (int64_t r = 0; r < n; ++r) {
    if (m > 0 && std::memcmp(rows + K * (m - 1), rows + K * r, K * sizeof(int32_t)) == 0) {
      if (counts != nullptr) ++counts[m - 1];
      continue;
    }
    if (m != r) std::memcpy(rows + K * m, rows + K * r, K * sizeof(int32_t));
    if (counts != nullptr) counts[m] = 1;
    ++m;
  }
  return m;
}

}  // namespace

extern "C" {

int dmh_sort_unique_rows_i32(int32_t* rows, int64_t n, int k, int64_t N, int32_t* counts, int64_t* n_unique, int threads) {
  if (n < 0 || n_unique == nullptr || (n > 0 && rows == nullptr) || N < 0 || n > (int64_t)std::numeric_limits<int32_t>::max())
    return DMH_ERR_ARG;
  *n_unique = 0;
  for (int64_t i = 0; i < n * (int64_t)k; ++i)
    if (rows[i] < 0 || rows[i] >= N) return DMH_ERR_ARG;
  const int nth = std::max(1, std::min(threads <= 0 ? 8 : threads, 32));
  switch (k) {
    case 2: *n_unique = sort_unique<2>(rows, n, N, nth, counts); break;
    case 3: *n_unique = sort_unique<3>(rows, n, N, nth, counts); break;
    case 4: *n_unique = sort_unique<4>(rows, n, N, nth, counts); break;
    default: return DMH_ERR_ARG;
  }
  return DMH_OK;
}

}  // extern "C"
