// Host Delaunay triangulator behind include/distmesh_host.h (3-D).
//
// The reference retriangulates with CGAL's Delaunay_triangulation_3 on every DistMesh iteration
// (SeismicMesh/generation/cpp/delaunay_class3.cpp behind mesh_generator.py:466-481); this is the
// from-scratch replacement on raw buffers that keeps the caller's vertex numbering.
//
// Construction: incremental Bowyer-Watson.  The points are inserted in rounds of growing size (a
// biased randomised insertion order: each round is a random sample eight times larger than the one
// before) and along a Morton curve within a round, so the tetrahedron containing a new point is a
// short walk from the one created last.  The tetrahedra whose open circumball contains the point
// (found by a flood fill from the located one) are replaced by the fan joining the point to the
// boundary facets of that cavity.  The convex hull is closed by "ghost" tetrahedra that share one
// vertex at infinity (id -1): a ghost is in conflict with a point strictly beyond its hull facet,
// or in the plane of the facet when the finite tetrahedron behind it is in conflict.
//
// Exactness: orient3d / insphere run a floating-point filter and fall back to exact arithmetic on
// expansions (dm_exact.h).  With exact signs and STRICT conflicts the cavity is star-shaped around
// the new point even for co-spherical input (a facet in the plane of the point has the point inside
// the circumballs on both of its sides or on neither), so no tolerance is needed; the fan is
// nevertheless checked (every new finite tetrahedron positively oriented) and a violation is
// reported as `lost`, never papered over.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <system_error>
#include <thread>
#include <utility>
#include <vector>
#if defined(__linux__)
#include <sched.h>
#endif

#include "distmesh_host.h"
#include "dm_cell_order.h"
#include "dm_exact.h"

namespace {

using namespace dmx;

constexpr double EPS = 1.1102230246251565e-16;  // 2^-53
constexpr double O3D_BOUND = (7.0 + 56.0 * EPS) * EPS;
constexpr double ISP_BOUND = (16.0 + 224.0 * EPS) * EPS;

inline XV xd(double a, double b) { return XV(ex_diff(a, b)); }

// ---- tie path for exactly representable differences (the common case: nearby points whose
//      coordinate differences need no tail): everything on the stack, sizes known at compile time
inline Ex<2> prod2(double a, double b) {
  Ex<2> r;
  double hi, lo;
  two_prod(a, b, hi, lo);
  r.n = 0;
  if (lo != 0.0) r.c[r.n++] = lo;
  if (hi != 0.0 || r.n == 0) r.c[r.n++] = hi;
  return r;
}
template <int A>
inline Ex<2 * A> scaled(const Ex<A>& x, double b) {
  Ex<2 * A> r;
  r.n = ex_scale(x.c, x.n, b, r.c);
  return r;
}
// u*v - w*z
inline Ex<4> cross2(double u, double v, double w, double z) {
  Ex<4> r;
  ex_sub(prod2(u, v), prod2(w, z), r);
  return r;
}
// x*p + s*y*q + z*r with s = +-1 (the 3 x 3 minors)
inline Ex<24> minor3(double x, const Ex<4>& p, double sy, const Ex<4>& q, double z, const Ex<4>& r) {
  Ex<16> t;
  ex_add(scaled(p, x), scaled(q, sy), t);
  Ex<24> out;
  ex_add(t, scaled(r, z), out);
  return out;
}
inline Ex<6> lift3(double x, double y, double z) {
  Ex<4> t;
  ex_add(prod2(x, x), prod2(y, y), t);
  Ex<6> out;
  ex_add(t, prod2(z, z), out);
  return out;
}

inline bool exact_diff(double a, double b, double& d) {
  double lo;
  two_diff(a, b, d, lo);
  return lo == 0.0;
}

double orient3d_exact(const double* a, const double* b, const double* c, const double* d) {
  {
    double v[9];
    bool ok = true;
    for (int k = 0; k < 3; ++k)
      ok = ok & exact_diff(a[k], d[k], v[k]) & exact_diff(b[k], d[k], v[3 + k]) & exact_diff(c[k], d[k], v[6 + k]);
    if (ok) {
      const double adx = v[0], ady = v[1], adz = v[2], bdx = v[3], bdy = v[4], bdz = v[5], cdx = v[6], cdy = v[7], cdz = v[8];
      const Ex<24> det = minor3(adz, cross2(bdx, cdy, cdx, bdy), bdz, cross2(cdx, ady, adx, cdy), cdz, cross2(adx, bdy, bdx, ady));
      return ex_sign(det);
    }
  }
  const XV adx = xd(a[0], d[0]), ady = xd(a[1], d[1]), adz = xd(a[2], d[2]);
  const XV bdx = xd(b[0], d[0]), bdy = xd(b[1], d[1]), bdz = xd(b[2], d[2]);
  const XV cdx = xd(c[0], d[0]), cdy = xd(c[1], d[1]), cdz = xd(c[2], d[2]);
  const XV bc = xv_sub(xv_mul(bdx, cdy), xv_mul(cdx, bdy));
  const XV ca = xv_sub(xv_mul(cdx, ady), xv_mul(adx, cdy));
  const XV ab = xv_sub(xv_mul(adx, bdy), xv_mul(bdx, ady));
  return xv_add(xv_add(xv_mul(adz, bc), xv_mul(bdz, ca)), xv_mul(cdz, ab)).sign();
}

// > 0: (a, b, c, d) is positively oriented (the convention every stored tetrahedron satisfies)
inline double orient3d(const double* a, const double* b, const double* c, const double* d) {
  const double adx = a[0] - d[0], bdx = b[0] - d[0], cdx = c[0] - d[0];
  const double ady = a[1] - d[1], bdy = b[1] - d[1], cdy = c[1] - d[1];
  const double adz = a[2] - d[2], bdz = b[2] - d[2], cdz = c[2] - d[2];
  const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
  const double cdxady = cdx * ady, adxcdy = adx * cdy;
  const double adxbdy = adx * bdy, bdxady = bdx * ady;
  const double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
  const double permanent = (std::fabs(bdxcdy) + std::fabs(cdxbdy)) * std::fabs(adz) +
                           (std::fabs(cdxady) + std::fabs(adxcdy)) * std::fabs(bdz) +
                           (std::fabs(adxbdy) + std::fabs(bdxady)) * std::fabs(cdz);
  const double bound = O3D_BOUND * permanent;
  if (det > bound || -det > bound) return det;
  return orient3d_exact(a, b, c, d);
}

double insphere_exact(const double* a, const double* b, const double* c, const double* d, const double* e) {
  {
    double v[12];
    bool ok = true;
    for (int k = 0; k < 3; ++k)
      ok = ok & exact_diff(a[k], e[k], v[k]) & exact_diff(b[k], e[k], v[3 + k]) & exact_diff(c[k], e[k], v[6 + k]) &
           exact_diff(d[k], e[k], v[9 + k]);
    if (ok) {
      const double aex = v[0], aey = v[1], aez = v[2], bex = v[3], bey = v[4], bez = v[5];
      const double cex = v[6], cey = v[7], cez = v[8], dex = v[9], dey = v[10], dez = v[11];
      const Ex<4> ab = cross2(aex, bey, bex, aey), bc = cross2(bex, cey, cex, bey), cd = cross2(cex, dey, dex, cey);
      const Ex<4> da = cross2(dex, aey, aex, dey), ac = cross2(aex, cey, cex, aey), bd = cross2(bex, dey, dex, bey);
      const Ex<24> abc = minor3(aez, bc, -bez, ac, cez, ab), bcd = minor3(bez, cd, -cez, bd, dez, bc);
      const Ex<24> cda = minor3(cez, da, dez, ac, aez, cd), dab = minor3(dez, ab, aez, bd, bez, da);
      static thread_local Ex<288> t1, t2, t3, t4;
      static thread_local Ex<576> s1, s2;
      static thread_local Ex<1152> det;
      ex_mul(abc, lift3(dex, dey, dez), t1);
      ex_mul(dab, lift3(cex, cey, cez), t2);
      ex_mul(cda, lift3(bex, bey, bez), t3);
      ex_mul(bcd, lift3(aex, aey, aez), t4);
      ex_sub(t1, t2, s1);
      ex_sub(t3, t4, s2);
      ex_add(s1, s2, det);
      return ex_sign(det);
    }
  }
  const XV aex = xd(a[0], e[0]), aey = xd(a[1], e[1]), aez = xd(a[2], e[2]);
  const XV bex = xd(b[0], e[0]), bey = xd(b[1], e[1]), bez = xd(b[2], e[2]);
  const XV cex = xd(c[0], e[0]), cey = xd(c[1], e[1]), cez = xd(c[2], e[2]);
  const XV dex = xd(d[0], e[0]), dey = xd(d[1], e[1]), dez = xd(d[2], e[2]);
  const XV ab = xv_sub(xv_mul(aex, bey), xv_mul(bex, aey));
  const XV bc = xv_sub(xv_mul(bex, cey), xv_mul(cex, bey));
  const XV cd = xv_sub(xv_mul(cex, dey), xv_mul(dex, cey));
  const XV da = xv_sub(xv_mul(dex, aey), xv_mul(aex, dey));
  const XV ac = xv_sub(xv_mul(aex, cey), xv_mul(cex, aey));
  const XV bd = xv_sub(xv_mul(bex, dey), xv_mul(dex, bey));
  const XV abc = xv_add(xv_sub(xv_mul(aez, bc), xv_mul(bez, ac)), xv_mul(cez, ab));
  const XV bcd = xv_add(xv_sub(xv_mul(bez, cd), xv_mul(cez, bd)), xv_mul(dez, bc));
  const XV cda = xv_add(xv_add(xv_mul(cez, da), xv_mul(dez, ac)), xv_mul(aez, cd));
  const XV dab = xv_add(xv_add(xv_mul(dez, ab), xv_mul(aez, bd)), xv_mul(bez, da));
  auto lift = [](const XV& x, const XV& y, const XV& z) {
    return xv_add(xv_add(xv_mul(x, x), xv_mul(y, y)), xv_mul(z, z));
  };
  const XV al = lift(aex, aey, aez), bl = lift(bex, bey, bez), cl = lift(cex, cey, cez), dl = lift(dex, dey, dez);
  return xv_add(xv_sub(xv_mul(dl, abc), xv_mul(cl, dab)), xv_sub(xv_mul(bl, cda), xv_mul(al, bcd))).sign();
}

// > 0: e strictly inside the sphere through a, b, c, d, for a positively oriented (a, b, c, d)
inline double insphere(const double* a, const double* b, const double* c, const double* d, const double* e) {
  const double aex = a[0] - e[0], bex = b[0] - e[0], cex = c[0] - e[0], dex = d[0] - e[0];
  const double aey = a[1] - e[1], bey = b[1] - e[1], cey = c[1] - e[1], dey = d[1] - e[1];
  const double aez = a[2] - e[2], bez = b[2] - e[2], cez = c[2] - e[2], dez = d[2] - e[2];
  const double aexbey = aex * bey, bexaey = bex * aey, ab = aexbey - bexaey;
  const double bexcey = bex * cey, cexbey = cex * bey, bc = bexcey - cexbey;
  const double cexdey = cex * dey, dexcey = dex * cey, cd = cexdey - dexcey;
  const double dexaey = dex * aey, aexdey = aex * dey, da = dexaey - aexdey;
  const double aexcey = aex * cey, cexaey = cex * aey, ac = aexcey - cexaey;
  const double bexdey = bex * dey, dexbey = dex * bey, bd = bexdey - dexbey;
  const double abc = aez * bc - bez * ac + cez * ab;
  const double bcd = bez * cd - cez * bd + dez * bc;
  const double cda = cez * da + dez * ac + aez * cd;
  const double dab = dez * ab + aez * bd + bez * da;
  const double al = aex * aex + aey * aey + aez * aez, bl = bex * bex + bey * bey + bez * bez;
  const double cl = cex * cex + cey * cey + cez * cez, dl = dex * dex + dey * dey + dez * dez;
  const double det = (dl * abc - cl * dab) + (bl * cda - al * bcd);
  // the same expression with every term replaced by its magnitude
  const double Aab = std::fabs(aexbey) + std::fabs(bexaey), Abc = std::fabs(bexcey) + std::fabs(cexbey);
  const double Acd = std::fabs(cexdey) + std::fabs(dexcey), Ada = std::fabs(dexaey) + std::fabs(aexdey);
  const double Aac = std::fabs(aexcey) + std::fabs(cexaey), Abd = std::fabs(bexdey) + std::fabs(dexbey);
  const double fa = std::fabs(aez), fb = std::fabs(bez), fc = std::fabs(cez), fd = std::fabs(dez);
  const double permanent = dl * (fa * Abc + fb * Aac + fc * Aab) + cl * (fd * Aab + fa * Abd + fb * Ada) +
                           bl * (fc * Ada + fd * Aac + fa * Acd) + al * (fb * Acd + fc * Abd + fd * Abc);
  const double bound = ISP_BOUND * permanent;
  if (det > bound || -det > bound) return det;
  return insphere_exact(a, b, c, d, e);
}

// exact test: are a, b, c collinear (all three coordinate-plane projections have zero area)
bool collinear(const double* a, const double* b, const double* c) {
  for (int k = 0; k < 3; ++k) {
    const int i = k, j = (k + 1) % 3;
    const XV l = xv_mul(xd(a[i], c[i]), xd(b[j], c[j])), r = xv_mul(xd(a[j], c[j]), xd(b[i], c[i]));
    if (xv_sub(l, r).sign() != 0.0) return false;
  }
  return true;
}

inline uint64_t spread3(uint64_t x) {  // 21 bits -> every third bit
  x &= 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}

constexpr int32_t INF = -1, DEAD = -2;
constexpr int MAX_THREADS = 32;      // (stamps of thread j are j + 1 + k * <threads of the call + 1>: unique over the whole run)
constexpr int64_t CHUNK = 512;       // tetrahedron slots a thread claims at a time
constexpr int PAR_PASSES = 6;
constexpr int64_t PAR_MIN_ROUND = 8000;   // rounds smaller than this are inserted by one thread
inline int64_t env_or(const char* name, int64_t dflt) {
  const char* e = std::getenv(name);
  return e != nullptr && std::atoll(e) > 0 ? std::atoll(e) : dflt;
}
// rows a box should hold at least: per round (number of boxes) and per pass (when to stop)
inline int64_t par_round_rows() {
  static const int64_t v = env_or("DM_HOST_ROUND_ROWS", 2000);
  return v;
}
inline int par_passes() {
  static const int v = (int)env_or("DM_HOST_PASSES", PAR_PASSES);
  return v;
}
inline int64_t par_pass_rows() {
  static const int64_t v = env_or("DM_HOST_PASS_ROWS", 1500);
  return v;
}

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline bool trace_on() {
  static const bool on = std::getenv("DM_HOST_TRACE") != nullptr;
  return on;
}

template <class F>
void run_threads(int nth, F&& fn) {
  if (nth <= 1) {
    fn(0);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nth - 1);
  int started = 1;
  try {
    for (int j = 1; j < nth; ++j) {
      th.emplace_back([&fn, j] { fn(j); });
      ++started;
    }
  } catch (const std::system_error&) {
    // the process may not have another thread: the pieces that got none run here, one after the other
    // (no piece ever waits for another)
  }
  fn(0);
  for (int j = started; j < nth; ++j) fn(j);
  for (auto& t : th) t.join();
}

struct Facet {
  int32_t t, k, nb;
};
struct Slot {
  uint64_t key;
  int32_t t, j, tag;
};

// what one inserting thread owns
struct Ctx {
  std::vector<int32_t> freelist, stack, cav, newt, deferred;
  std::vector<Facet> bnd;
  std::vector<Slot> table;
  int32_t stamp = 0, stamp_next = 1, last = -1;
  int64_t dups = 0, lost = 0;
  bool failed = false, exhausted = false;
  int me = 0;
};

// The triangulation.  Serial phases may grow the arrays; during a parallel phase their size is fixed
// and every thread touches only tetrahedra whose finite vertices all belong to its own partition
// (see insert<true>), so no two threads ever read or write the same tetrahedron.
struct Delaunay3 {
  const double* P = nullptr;  // coordinates in insertion order
  int64_t n = 0;
  struct alignas(32) Tet {
    int32_t v[4];  // vertices (INF: the vertex at infinity; v[0] == DEAD: slot unused)
    int32_t n[4];  // neighbours, opposite the vertex of the same slot
  };
  std::vector<Tet> T;           // (vertices and neighbours of a tetrahedron share half a cache line)
  std::vector<int32_t> mark;    // +stamp: in the cavity of that insertion, -stamp: tested and not in conflict
  std::vector<uint8_t> owner;   // partition of every vertex during a parallel phase
  std::atomic<int64_t> top{0};  // slots handed out so far
  int32_t stamp_stride = 1;     // number of inserting contexts (a small stride keeps the stamps inside int32: at most
                                // ~4 attempts per point and context even when every pass hands everything back)
  bool parallel = false;

  const double* pt(int32_t v) const { return P + 3 * (int64_t)v; }
  bool is_dead(int64_t t) const { return T[t].v[0] == DEAD; }
  int64_t slots() const { return top.load(std::memory_order_relaxed); }

  void grow(int64_t want) {
    if ((int64_t)T.size() >= want) return;
    const int64_t sz = std::max<int64_t>(want, (int64_t)T.size() * 3 / 2);
    T.resize(sz, Tet{{DEAD, DEAD, DEAD, DEAD}, {0, 0, 0, 0}});
    mark.resize(sz, 0);
  }
  // make sure the thread has `need` free slots; false: the arrays are full and may not grow now
  bool reserve_slots(Ctx& c, int64_t need) {
    while ((int64_t)c.freelist.size() < need) {
      const int64_t k = std::max<int64_t>(CHUNK, need - (int64_t)c.freelist.size());
      const int64_t at = top.fetch_add(k, std::memory_order_relaxed);
      if (at + k > (int64_t)T.size()) {
        if (parallel) {
          top.fetch_sub(k, std::memory_order_relaxed);
          return false;
        }
        grow(at + k);
      }
      for (int64_t t = at + k - 1; t >= at; --t) c.freelist.push_back((int32_t)t);  // lowest slot on top
    }
    return true;
  }
  int32_t new_tet(Ctx& c) {  // after reserve_slots
    const int32_t t = c.freelist.back();
    c.freelist.pop_back();
    T[t].v[0] = 0;
    mark[t] = 0;
    return t;
  }
  int inf_slot(int32_t t) const {
    const int32_t* v = T[t].v;
    return v[0] == INF ? 0 : v[1] == INF ? 1 : v[2] == INF ? 2 : v[3] == INF ? 3 : -1;
  }
  // every finite vertex of t in partition `me` (the vertex at infinity belongs to everybody: two
  // tetrahedra of different partitions never share a facet, which has at least two finite vertices)
  bool own(int32_t t, int me) const {
    const int32_t* v = T[t].v;
    for (int k = 0; k < 4; ++k)
      if (v[k] != INF && owner[v[k]] != me) return false;
    return true;
  }
  // orientation of tetrahedron t with the vertex of slot k replaced by the point p
  double orient_with(int32_t t, int k, const double* p) const {
    const int32_t* v = T[t].v;
    const double* q[4] = {k == 0 ? p : pt(v[0]), k == 1 ? p : pt(v[1]), k == 2 ? p : pt(v[2]), k == 3 ? p : pt(v[3])};
    return orient3d(q[0], q[1], q[2], q[3]);
  }
  double insphere_of(int32_t t, const double* p) const {
    const int32_t* v = T[t].v;
    return insphere(pt(v[0]), pt(v[1]), pt(v[2]), pt(v[3]), p);
  }
  bool conflict(int32_t t, const double* p) const {
    const int ki = inf_slot(t);
    if (ki < 0) return insphere_of(t, p) > 0.0;
    const double o = orient_with(t, ki, p);  // > 0: strictly beyond the hull facet
    if (o != 0.0) return o > 0.0;
    return insphere_of(T[t].n[ki], p) > 0.0;  // in the facet's plane: as the tetrahedron behind it
  }

  // neighbours of the first five tetrahedra by matching faces
  void link_all() {
    const int64_t m = slots();
    for (int64_t t = 0; t < m; ++t)
      for (int k = 0; k < 4; ++k) {
        if (is_dead(t)) continue;
        int32_t f[3];
        int c = 0;
        for (int j = 0; j < 4; ++j)
          if (j != k) f[c++] = T[t].v[j];
        std::sort(f, f + 3);
        for (int64_t u = 0; u < m; ++u) {
          if (u == t || is_dead(u)) continue;
          for (int kk = 0; kk < 4; ++kk) {
            int32_t g[3];
            int d = 0;
            for (int j = 0; j < 4; ++j)
              if (j != kk) g[d++] = T[u].v[j];
            std::sort(g, g + 3);
            if (f[0] == g[0] && f[1] == g[1] && f[2] == g[2]) T[t].n[k] = (int32_t)u;
          }
        }
      }
  }

  void init(Ctx& cx, int32_t a, int32_t b, int32_t c, int32_t d) {
    if (orient3d(pt(a), pt(b), pt(c), pt(d)) < 0.0) std::swap(a, b);
    reserve_slots(cx, 5);
    const int32_t t0 = new_tet(cx);
    int32_t* v = T[t0].v;
    v[0] = a, v[1] = b, v[2] = c, v[3] = d;
    for (int k = 0; k < 4; ++k) {  // ghost behind face k: the vertex at infinity in slot k, orientation flipped
      const int32_t g = new_tet(cx);
      int32_t* w = T[g].v;
      const int32_t* s = T[t0].v;
      for (int j = 0; j < 4; ++j) w[j] = s[j];
      w[k] = INF;
      const int x = k == 0 ? 1 : 0, y = k <= 1 ? 2 : 1;  // two of the other slots
      std::swap(w[x], w[y]);
    }
    link_all();
    cx.last = t0;
  }

  void match_face(Ctx& c, int32_t t, int j, int32_t u, int32_t w) {
    if (u > w) std::swap(u, w);
    const uint64_t key = (uint64_t)(u + 1) * (uint64_t)(n + 2) + (uint64_t)(w + 1);
    const size_t mask = c.table.size() - 1;
    size_t h = (size_t)(key * 0x9E3779B97F4A7C15ULL >> 20) & mask;
    for (;;) {
      Slot& s = c.table[h];
      if (s.tag != c.stamp) {
        s = Slot{key, t, j, c.stamp};
        return;
      }
      if (s.key == key) {
        T[t].n[j] = s.t;
        T[s.t].n[s.j] = t;
        return;
      }
      h = (h + 1) & mask;
    }
  }

  // Insert row i.  PAR: the walk and the cavity may only contain tetrahedra of the thread's own box (all
  // finite vertices in it); as soon as one of them reaches another tetrahedron the point is given back
  // (false) before anything has been modified.  The ring of tetrahedra around the cavity may contain
  // tetrahedra with vertices of several boxes -- those are frozen during the phase.
  template <bool PAR>
  bool insert(Ctx& c, int32_t i) {
    const double* p = pt(i);
    // ---- locate: walk from the tetrahedron created last
    int32_t t = c.last;
    if (PAR && (t < 0 || !own(t, c.me))) return false;
    const int64_t limit = PAR ? 4096 : 4 * slots() + 64;
    bool found = false;
    for (int64_t steps = 0; steps < limit; ++steps) {
      int32_t nxt = -1;
      const int ki = inf_slot(t);
      if (ki >= 0) {
        if (conflict(t, p)) {
          found = true;
          break;
        }
        nxt = T[t].n[ki];
      } else {
        for (int kk = 0; kk < 4; ++kk) {
          const int k = (kk + (int)(steps & 3)) & 3;
          if (orient_with(t, k, p) < 0.0) {
            nxt = T[t].n[k];
            if (!PAR || own(nxt, c.me)) break;  // (any face that separates t from p will do)
          }
        }
        if (nxt < 0) {
          found = true;
          break;
        }
      }
      if (PAR && !own(nxt, c.me)) return false;
      t = nxt;
    }
    if (!found) {  // the walk did not settle: exhaustive search
      if (PAR) return false;
      for (int64_t u = 0; u < slots() && !found; ++u)
        if (!is_dead(u) && conflict((int32_t)u, p)) t = (int32_t)u, found = true;
      if (!found) {
        ++c.lost;
        return true;
      }
    }
    if (inf_slot(t) < 0) {
      const int32_t* v = T[t].v;
      for (int k = 0; k < 4; ++k) {
        const double* q = pt(v[k]);
        if (q[0] == p[0] && q[1] == p[1] && q[2] == p[2]) {  // exact duplicate of an earlier row
          ++c.dups;
          return true;
        }
      }
      if (!conflict(t, p)) {
        ++c.lost;
        return true;
      }
    }
    // ---- cavity: flood fill over the tetrahedra in conflict
    c.stamp = c.stamp_next;
    c.stamp_next += stamp_stride;
    const int32_t stamp = c.stamp;
    c.cav.clear();
    c.bnd.clear();
    c.stack.clear();
    c.stack.push_back(t);
    mark[t] = stamp;
    while (!c.stack.empty()) {
      const int32_t cc = c.stack.back();
      c.stack.pop_back();
      c.cav.push_back(cc);
      for (int k = 0; k < 4; ++k) {
        const int32_t nb = T[cc].n[k];
        if (mark[nb] == stamp) continue;
        if (PAR && !own(nb, c.me)) {
          // A tetrahedron with vertices of several boxes: nobody deletes it or moves its vertices in this
          // phase, so it can be tested; it cannot join the cavity, but it may sit in the ring around it
          // (only the neighbour slot facing the cavity is written, see below).  Never marked: no thread
          // writes the mark of a tetrahedron that is not its own.
          if (conflict(nb, p)) return false;
          c.bnd.push_back(Facet{cc, k, nb});
          continue;
        }
        if (mark[nb] != -stamp) {
          if (conflict(nb, p)) {
            mark[nb] = stamp;
            c.stack.push_back(nb);
            continue;
          }
          mark[nb] = -stamp;
        }
        c.bnd.push_back(Facet{cc, k, nb});
      }
    }
    // ---- the fan: one new tetrahedron per boundary facet
    if (!reserve_slots(c, (int64_t)c.bnd.size())) {
      c.exhausted = true;
      return false;
    }
    size_t want = 64;
    while (want < 8 * c.bnd.size()) want <<= 1;
    if (c.table.size() < want) c.table.assign(want, Slot{0, 0, 0, 0});
    int32_t fin = -1;
    c.newt.clear();
    for (const Facet& f : c.bnd) {
      const int32_t nt = new_tet(c);
      int32_t* w = T[nt].v;
      const int32_t* s = T[f.t].v;
      for (int j = 0; j < 4; ++j) w[j] = s[j];
      w[f.k] = i;
      T[nt].n[f.k] = f.nb;
      // (relaxed atomics: in a parallel phase the ring tetrahedron may belong to no box, and then another
      //  thread may be looking for ITS slot in it at the same time -- each of them only ever writes the
      //  slot that faces its own cavity)
      int32_t* back = T[f.nb].n;
      for (int m = 0; m < 4; ++m)
        if (__atomic_load_n(&back[m], __ATOMIC_RELAXED) == f.t) {
          __atomic_store_n(&back[m], nt, __ATOMIC_RELAXED);
          break;
        }
      if (w[0] != INF && w[1] != INF && w[2] != INF && w[3] != INF) {
        if (!(orient3d(pt(w[0]), pt(w[1]), pt(w[2]), pt(w[3])) > 0.0)) c.failed = true;  // cavity not star-shaped
        fin = nt;
      }
      for (int j = 0; j < 4; ++j) {
        if (j == f.k) continue;
        int32_t e[2];
        int cnt = 0;
        for (int m = 0; m < 4; ++m)
          if (m != j && m != f.k) e[cnt++] = w[m];
        match_face(c, nt, j, e[0], e[1]);
      }
      c.newt.push_back(nt);
    }
    for (int32_t cc : c.cav) {
      T[cc].v[0] = DEAD;
      c.freelist.push_back(cc);
    }
    c.last = fin >= 0 ? fin : c.newt.back();
    return true;
  }
};

int pick_threads(int threads) {
  if (threads <= 0) {
    const char* e = std::getenv("DM_HOST_THREADS");
    threads = e != nullptr ? std::atoi(e) : 0;
  }
  if (threads <= 0) {
    threads = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) threads = std::min(threads > 0 ? threads : 1 << 20, CPU_COUNT(&set));
#endif
    if (threads >= 8) --threads;  // (one core left to the caller's other threads: measured faster on a 16-core box)
    threads = std::min(threads, 16);
  }
  return std::max(1, std::min(threads, MAX_THREADS - 1));
}

// Partition of space into `cnt` boxes holding about the same number of points each: a k-d tree over
// a sample of the points, every node cutting `cnt` boxes into a and cnt - a at the a / cnt quantile of
// the node's axis.  Boxes are convex, so a straight walk between two points of a box stays in it.
struct KdTree {
  struct Node {
    int axis;
    double plane;
    int left, right;  // >= 0: node, < 0: -(box + 1)
  };
  std::vector<Node> nodes;
  int root = -1;
  int build(const double* P, int32_t* lo, int32_t* hi, int cnt, int first, int axis, int rule) {
    if (cnt <= 1 || hi - lo < 2) return -(first + 1);
    int a = rule == 0 ? cnt / 2 : rule == 1 ? (3 * cnt + 4) / 8 : (5 * cnt + 3) / 8;
    a = std::max(1, std::min(cnt - 1, a));
    // cut across the axis whose turn it is unless the box is thin in it (less than 0.6 of its longest
    // extent: a cut there would leave two plates that are mostly border)
    {
      double ext[3];
      for (int k = 0; k < 3; ++k) {
        double mn = P[3 * (int64_t)*lo + k], mx = mn;
        for (const int32_t* x = lo; x < hi; ++x) mn = std::min(mn, P[3 * (int64_t)*x + k]), mx = std::max(mx, P[3 * (int64_t)*x + k]);
        ext[k] = mx - mn;
      }
      const double longest = std::max(ext[0], std::max(ext[1], ext[2]));
      for (int tries = 0; tries < 3 && ext[axis] < 0.6 * longest; ++tries) axis = (axis + 1) % 3;
    }
    int32_t* mid = lo + (hi - lo) * (int64_t)a / cnt;
    std::nth_element(lo, mid, hi, [&](int32_t x, int32_t y) { return P[3 * (int64_t)x + axis] < P[3 * (int64_t)y + axis]; });
    const int me = (int)nodes.size();
    nodes.push_back(Node{axis, P[3 * (int64_t)*mid + axis], 0, 0});
    const int l = build(P, lo, mid, a, first, (axis + 1) % 3, rule);
    const int r = build(P, mid, hi, cnt - a, first + a, (axis + 1) % 3, rule);
    nodes[me].left = l;
    nodes[me].right = r;
    return me;
  }
  int box_of(const double* x) const {
    int ref = root;
    while (ref >= 0) {
      const Node& nd = nodes[ref];
      ref = x[nd.axis] < nd.plane ? nd.left : nd.right;
    }
    return -ref - 1;
  }
};

// One round of the insertion order by several threads.  `pending`: rows of the round in Morton order.
// Every pass cuts space into boxes along other planes, so that what a pass had to give back because it
// touched a border of its box is inside a box of the next one; what is left after the last pass goes to
// the serial code.
// `batch`: the points to insert are not spread like the points already in (ghost layers added to a slab): the
// boxes are then cut from a sample of the pending points themselves, so that every thread gets its share.
void parallel_round(Delaunay3& D, std::vector<Ctx>& ctx, std::vector<int32_t>& pending, int nth, bool batch = false) {
  const int64_t n = D.n;
  if ((int64_t)D.owner.size() < n) D.owner.resize(n, 0);
  std::vector<int32_t> sample;
  if (batch) {
    const int64_t stride = std::max<int64_t>(1, (int64_t)pending.size() / 32768);
    for (int64_t v = 0; v < (int64_t)pending.size(); v += stride) sample.push_back(pending[(size_t)v]);
  } else {
    const int64_t stride = std::max<int64_t>(1, n / 32768);
    for (int64_t v = 0; v < n; v += stride) sample.push_back((int32_t)v);
  }
  int most = nth;  // boxes of a pass (halved whenever a pass inserted less than a fifth of its points: boxes too small)
  for (int pass = 0; pass < par_passes(); ++pass) {
    if ((int64_t)pending.size() < 2 * par_pass_rows()) break;
    const int want = (int)std::max<int64_t>(2, std::min<int64_t>(most, (int64_t)pending.size() / par_pass_rows()));
    // (first axis, cutting rule, number of boxes: no two passes cut along the same planes)
    const int parts = std::min<int>((int)ctx.size(), want + (pass / 3) % 2);
    KdTree kd;
    kd.root = kd.build(D.P, sample.data(), sample.data() + sample.size(), parts, 0, pass % 3, (pass + pass / 3) % 3);
    run_threads(nth, [&](int j) {
      const int64_t lo = n * j / nth, hi = n * (j + 1) / nth;
      for (int64_t v = lo; v < hi; ++v) D.owner[v] = (uint8_t)kd.box_of(D.pt((int32_t)v));
    });
    // every thread: its rows of `pending` (in order) and a tetrahedron of its box to start from
    const int64_t m = D.slots();
    std::vector<std::vector<int32_t>> mine(parts);
    run_threads(parts, [&](int j) {
      Ctx& c = ctx[j];
      c.me = j;
      c.last = -1;
      c.deferred.clear();
      c.exhausted = false;
      for (int32_t r : pending)
        if (D.owner[r] == j) mine[j].push_back(r);
      if (mine[j].empty()) return;
      const int64_t from = m * j / parts;
      for (int64_t s = 0; s < m; ++s) {
        const int64_t t = from + s < m ? from + s : from + s - m;
        if (!D.is_dead(t) && D.own((int32_t)t, j)) {
          c.last = (int32_t)t;
          break;
        }
      }
    });
    D.parallel = true;
    const double tp0 = now_s();
    const size_t before = pending.size();
    run_threads(parts, [&](int j) {
      Ctx& c = ctx[j];
      for (int32_t r : mine[j])
        if (c.failed || c.exhausted || !D.insert<true>(c, r)) c.deferred.push_back(r);
    });
    D.parallel = false;
    // what is left, back in the order of the round
    std::vector<int32_t> next;
    bool stop = false;
    for (int j = 0; j < parts; ++j) {
      next.insert(next.end(), ctx[j].deferred.begin(), ctx[j].deferred.end());
      stop = stop || ctx[j].failed || ctx[j].exhausted;
    }
    std::sort(next.begin(), next.end());
    const bool poor = 5 * next.size() > 4 * pending.size();  // less than a fifth inserted: the boxes have become too small
    pending.swap(next);
    if (trace_on()) std::fprintf(stderr, "[dmh3d] pass %d: %d parts, %zu -> %zu pending, %.3f s\n", pass, parts, before, pending.size(), now_s() - tp0);
    if (stop || (poor && parts <= 2)) break;
    if (poor) most = std::max(2, parts / 2);
  }
}

// A triangulation that stays around between calls (dmh_dt3_*): the points in insertion order, the map back
// to the caller's rows, the tetrahedra and the per-thread insertion state.
struct Dt3 {
  std::vector<double> sorted;
  std::vector<int32_t> ids;
  Delaunay3 D;
  std::vector<Ctx> ctx;
  int nth = 1;
  int64_t N = 0;
  bool built = false, failed = false;
  double t_begin = 0.0, t_sorted = 0.0;
};

// insertion of all N points; S.built stays false when there are not four affinely independent points
int dt3_build(Dt3& S, const double* points, int64_t N, int threads) {
  int nth = pick_threads(threads);
  if (N < PAR_MIN_ROUND) nth = 1;
  S.nth = nth;
  S.N = N;
  const double t_begin = now_s();

  // ---- insertion order: rounds of growing size, Morton order within a round
  double lo[3], hi[3];
  for (int k = 0; k < 3; ++k) lo[k] = std::numeric_limits<double>::infinity(), hi[k] = -lo[k];
  for (int64_t i = 0; i < N; ++i)
    for (int k = 0; k < 3; ++k) {
      const double v = points[3 * i + k];
      if (!(v == v) || std::fabs(v) == std::numeric_limits<double>::infinity()) return DMH_ERR_ARG;
      lo[k] = std::min(lo[k], v);
      hi[k] = std::max(hi[k], v);
    }
  struct Item {
    uint64_t key;
    int32_t id;
  };
  std::vector<Item> items(N);
  run_threads(nth, [&](int j) {
    for (int64_t i = N * j / nth; i < N * (j + 1) / nth; ++i) {
      uint64_t q[3];
      for (int k = 0; k < 3; ++k) {
        const double w = hi[k] - lo[k];
        q[k] = w > 0.0 ? (uint64_t)std::min(2097151.0, (points[3 * i + k] - lo[k]) / w * 2097152.0) : 0;
      }
      const uint64_t morton = spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2;
      // round: 0 for 7 points in 8, 1 for 7 in 64, ... (a fixed hash of the row number: deterministic)
      uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ULL;
      h ^= h >> 29;
      h *= 0xBF58476D1CE4E5B9ULL;
      h ^= h >> 32;
      int round = 0;
      while (round < 20 && (h & 7) == 0) {
        ++round;
        h >>= 3;
      }
      items[i] = Item{((uint64_t)(20 - round) << 58) | (morton >> 6), (int32_t)i};  // later rounds sort last
    }
  });
  auto less = [](const Item& a, const Item& b) { return a.key != b.key ? a.key < b.key : a.id < b.id; };
  if (nth > 1) {  // sorted runs, then pairwise merges
    std::vector<int64_t> edge(nth + 1);
    for (int j = 0; j <= nth; ++j) edge[j] = N * j / nth;
    run_threads(nth, [&](int j) { std::sort(items.begin() + edge[j], items.begin() + edge[j + 1], less); });
    for (int w = 1; w < nth; w *= 2)
      run_threads((nth + 2 * w - 1) / (2 * w), [&](int g) {
        const int a = g * 2 * w, mid = std::min(nth, a + w), b = std::min(nth, a + 2 * w);
        if (mid < b) std::inplace_merge(items.begin() + edge[a], items.begin() + edge[mid], items.begin() + edge[b], less);
      });
  } else {
    std::sort(items.begin(), items.end(), less);
  }
  std::vector<double>& sorted = S.sorted;
  std::vector<int32_t>& ids = S.ids;
  sorted.assign(3 * N, 0.0);
  ids.assign(N, 0);
  run_threads(nth, [&](int j) {
    for (int64_t r = N * j / nth; r < N * (j + 1) / nth; ++r) {
      ids[r] = items[r].id;
      for (int k = 0; k < 3; ++k) sorted[3 * r + k] = points[3 * (int64_t)items[r].id + k];
    }
  });
  // where the rounds start (the round number is in the top bits of the key)
  std::vector<int64_t> round_start;
  if (nth > 1) {
    round_start.push_back(0);
    for (int64_t r = 1; r < N; ++r)
      if ((items[r].key >> 58) != (items[r - 1].key >> 58)) round_start.push_back(r);
    round_start.push_back(N);
  }
  items.clear();
  items.shrink_to_fit();

  const double t_sorted = now_s();
  Delaunay3& D = S.D;
  D.P = sorted.data();
  D.n = N;
  D.grow(7 * N + 64 + (nth > 1 ? (int64_t)nth * 4 * CHUNK + N : 0));
  std::vector<Ctx>& ctx = S.ctx;
  ctx.assign(nth + 1, Ctx());  // (a pass with shifted bounds has one partition more)
  D.stamp_stride = nth + 1;
  for (int j = 0; j <= nth; ++j) ctx[j].stamp_next = j + 1;
  Ctx& c0 = ctx[0];
  // ---- four affinely independent points to start from
  int32_t a = 0, b = -1, c = -1, d = -1;
  for (int64_t r = 1; r < N && b < 0; ++r)
    if (D.pt((int32_t)r)[0] != D.pt(a)[0] || D.pt((int32_t)r)[1] != D.pt(a)[1] || D.pt((int32_t)r)[2] != D.pt(a)[2]) b = (int32_t)r;
  for (int64_t r = 1; r < N && b >= 0 && c < 0; ++r)
    if (r != b && !collinear(D.pt(a), D.pt(b), D.pt((int32_t)r))) c = (int32_t)r;
  for (int64_t r = 1; r < N && c >= 0 && d < 0; ++r)
    if (r != b && r != c && orient3d(D.pt(a), D.pt(b), D.pt(c), D.pt((int32_t)r)) != 0.0) d = (int32_t)r;
  if (d < 0) return DMH_OK;  // fewer than four affinely independent points: no cell, every row lost
  S.built = true;
  D.init(c0, a, b, c, d);
  bool failed = false;
  auto serial = [&](int64_t r) {
    if (r == a || r == b || r == c || r == d) return;
    D.insert<false>(c0, (int32_t)r);
    failed = failed || c0.failed;
  };
  if (nth == 1) {
    for (int64_t r = 1; r < N && !failed; ++r) serial(r);
  } else {
    D.owner.assign(N, 0);
    for (size_t g = 0; g + 1 < round_start.size() && !failed; ++g) {
      const int64_t r0 = round_start[g], r1 = round_start[g + 1];
      const int use = (int)std::min<int64_t>(nth, (r1 - r0) / par_round_rows());
      if (r1 - r0 < PAR_MIN_ROUND || use < 2 || r0 < 64) {
        for (int64_t r = std::max<int64_t>(r0, 1); r < r1 && !failed; ++r) serial(r);
        continue;
      }
      std::vector<int32_t> pending;
      pending.reserve(r1 - r0);
      for (int64_t r = r0; r < r1; ++r)
        if (r != a && r != b && r != c && r != d) pending.push_back((int32_t)r);
      const double tr0 = now_s();
      parallel_round(D, ctx, pending, use);
      const double tr1 = now_s();
      for (int j = 0; j <= nth; ++j) failed = failed || ctx[j].failed;
      // the serial code starts its walk from a live tetrahedron
      if (c0.last < 0 || D.is_dead(c0.last)) {
        c0.last = -1;
        for (int64_t t = 0; t < D.slots() && c0.last < 0; ++t)
          if (!D.is_dead(t)) c0.last = (int32_t)t;
      }
      for (size_t x = 0; x < pending.size() && !failed; ++x) serial(pending[x]);
      if (trace_on()) std::fprintf(stderr, "[dmh3d] round of %ld: parallel %.3f s, serial rest (%zu) %.3f s\n", (long)(r1 - r0), tr1 - tr0, pending.size(), now_s() - tr1);
    }
  }
  S.failed = failed;
  S.t_begin = t_begin;
  S.t_sorted = t_sorted;
  return DMH_OK;
}

// the finite tetrahedra in the caller's numbering and the canonical output order
int dt3_extract(Dt3& S, int32_t* cells, int64_t cap, int64_t* T_out, int64_t* duplicates_out, int64_t* lost_out) {
  const double t_built = now_s();
  Delaunay3& D = S.D;
  std::vector<Ctx>& ctx = S.ctx;
  const std::vector<int32_t>& ids = S.ids;
  const int nth = S.nth;
  const int64_t N = S.N;
  const bool failed = S.failed;
  const double t_begin = S.t_begin, t_sorted = S.t_sorted;
  int64_t dups = 0, lost = 0;
  for (int j = 0; j <= nth; ++j) dups += ctx[j].dups, lost += ctx[j].lost;
  if (failed) lost += 1;
  // ---- the finite tetrahedra, in the caller's numbering
  const int64_t m = D.slots();
  std::vector<int64_t> cnt(nth + 1, 0);
  run_threads(nth, [&](int j) {
    int64_t k = 0;
    for (int64_t t = m * j / nth; t < m * (j + 1) / nth; ++t)
      if (!D.is_dead(t) && D.inf_slot((int32_t)t) < 0) ++k;
    cnt[j + 1] = k;
  });
  for (int j = 0; j < nth; ++j) cnt[j + 1] += cnt[j];
  const int64_t T = cnt[nth];
  *T_out = T;
  if (duplicates_out != nullptr) *duplicates_out = dups;
  if (lost_out != nullptr) *lost_out = lost;
  if (T > cap) return DMH_ERR_CAPACITY;
  run_threads(nth, [&](int j) {
    int64_t o = cnt[j];
    for (int64_t t = m * j / nth; t < m * (j + 1) / nth; ++t) {
      if (D.is_dead(t) || D.inf_slot((int32_t)t) >= 0) continue;
      for (int k = 0; k < 4; ++k) cells[4 * o + k] = ids[D.T[t].v[k]];
      ++o;
    }
  });
  const double t_out = now_s();
  dmx::order_cells<4>(cells, T, N, nth);
  if (trace_on())
    std::fprintf(stderr, "[dmh3d] N=%ld threads=%d: sort %.3f, build %.3f, extract %.3f, order %.3f s\n", (long)N, nth, t_sorted - t_begin,
                 t_built - t_sorted, t_out - t_built, now_s() - t_out);
  return DMH_OK;
}

// M more points (rows N .. N + M - 1 of the caller) into the triangulation that stands: serial insertion along
// a Morton curve of their own.  What the reference does with the ghost vertices of a slab: `dt.insert` on the
// CGAL triangulation that already holds the owned ones (mesh_generator.py:466, 715-731).
int dt3_insert(Dt3& S, const double* more, int64_t M) {
  if (M == 0) return DMH_OK;
  double lo[3], hi[3];
  for (int k = 0; k < 3; ++k) lo[k] = std::numeric_limits<double>::infinity(), hi[k] = -lo[k];
  for (int64_t i = 0; i < M; ++i)
    for (int k = 0; k < 3; ++k) {
      const double v = more[3 * i + k];
      if (!(v == v) || std::fabs(v) == std::numeric_limits<double>::infinity()) return DMH_ERR_ARG;
      lo[k] = std::min(lo[k], v);
      hi[k] = std::max(hi[k], v);
    }
  const int64_t N = S.N;
  std::vector<std::pair<uint64_t, int32_t>> order((size_t)M);
  for (int64_t i = 0; i < M; ++i) {
    uint64_t q[3];
    for (int k = 0; k < 3; ++k) {
      const double w = hi[k] - lo[k];
      q[k] = w > 0.0 ? (uint64_t)std::min(2097151.0, (more[3 * i + k] - lo[k]) / w * 2097152.0) : 0;
    }
    order[(size_t)i] = {spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2, (int32_t)i};
  }
  std::sort(order.begin(), order.end());
  S.sorted.reserve(S.sorted.size() + 3 * M);
  for (int64_t r = 0; r < M; ++r) {
    const int64_t i = order[(size_t)r].second;
    for (int k = 0; k < 3; ++k) S.sorted.push_back(more[3 * i + k]);
    S.ids.push_back((int32_t)(N + i));
  }
  Delaunay3& D = S.D;
  D.P = S.sorted.data();
  D.n = N + M;
  S.N = N + M;
  Ctx& c0 = S.ctx[0];
  if (c0.last < 0 || D.is_dead(c0.last)) {
    c0.last = -1;
    for (int64_t t = 0; t < D.slots() && c0.last < 0; ++t)
      if (!D.is_dead(t)) c0.last = (int32_t)t;
  }
  // a large batch goes through the passes of a round (boxes cut from the batch itself), the rest serially
  std::vector<int32_t> pending;
  const int use = (int)std::min<int64_t>(S.nth, M / par_round_rows());
  if (M >= PAR_MIN_ROUND && use >= 2) {
    // one point in sixteen first, serially: a layer of points beyond a flat side of the hull starts with points
    // that see the whole side (their cavity spans every box); once the side is broken up cavities are local
    for (int64_t r = N; r < N + M; ++r) {
      if ((r - N) % 16 == 0)
        D.insert<false>(c0, (int32_t)r);
      else
        pending.push_back((int32_t)r);
    }
    D.grow(D.slots() + 8 * M + (int64_t)S.nth * 4 * CHUNK);  // (the arrays may not grow during a parallel phase)
    parallel_round(D, S.ctx, pending, use, true);
    for (size_t j = 0; j < S.ctx.size(); ++j) S.failed = S.failed || S.ctx[j].failed;
    if (c0.last < 0 || D.is_dead(c0.last)) {
      c0.last = -1;
      for (int64_t t = 0; t < D.slots() && c0.last < 0; ++t)
        if (!D.is_dead(t)) c0.last = (int32_t)t;
    }
    for (size_t x = 0; x < pending.size() && !S.failed && !c0.failed; ++x) D.insert<false>(c0, pending[x]);
  } else {
    for (int64_t r = N; r < N + M && !c0.failed; ++r) D.insert<false>(c0, (int32_t)r);
  }
  S.failed = S.failed || c0.failed;
  return DMH_OK;
}

}  // namespace

extern "C" {

int64_t dmh_delaunay3d_max_cells(int64_t N) { return N < 4 ? 1 : 8 * N + 64; }

int dmh_delaunay3d_mt(const double* points, int64_t N, int32_t* cells, int64_t cap, int64_t* T_out,
                      int64_t* duplicates_out, int64_t* lost_out, int threads) {
  if (N < 0 || cap < 0 || T_out == nullptr || (N > 0 && points == nullptr) || (cap > 0 && cells == nullptr) ||
      N > (int64_t)std::numeric_limits<int32_t>::max() / 64)
    return DMH_ERR_ARG;
  *T_out = 0;
  if (duplicates_out != nullptr) *duplicates_out = 0;
  if (lost_out != nullptr) *lost_out = N;
  if (N < 4) return DMH_OK;
  Dt3 S;
  const int rc = dt3_build(S, points, N, threads);
  if (rc != DMH_OK || !S.built) return rc;
  return dt3_extract(S, cells, cap, T_out, duplicates_out, lost_out);
}

int dmh_delaunay3d(const double* points, int64_t N, int32_t* cells, int64_t cap, int64_t* T_out,
                   int64_t* duplicates_out, int64_t* lost_out) {
  return dmh_delaunay3d_mt(points, N, cells, cap, T_out, duplicates_out, lost_out, 1);
}

/* ---- a triangulation that stays around: build, read the cells, add points, read the cells again ---- */
namespace {
struct Dt3Handle {
  std::unique_ptr<Dt3> s;
  std::vector<double> all;  // every point so far, in the caller's order (a degenerate start is redone from these)
  int threads = 0;
  int64_t N = 0;
  int rebuild() {
    s.reset(new Dt3);
    s->N = N;
    return N >= 4 ? dt3_build(*s, all.data(), N, threads) : DMH_OK;
  }
};
}  // namespace

void* dmh_dt3_build(const double* points, int64_t N, int threads, int* rc_out) {
  int rc = DMH_OK;
  Dt3Handle* H = nullptr;
  if (N < 0 || (N > 0 && points == nullptr) || N > (int64_t)std::numeric_limits<int32_t>::max() / 64) {
    rc = DMH_ERR_ARG;
  } else {
    H = new Dt3Handle;
    H->all.assign(points, points + 3 * N);
    H->N = N;
    H->threads = threads;
    rc = H->rebuild();
    if (rc != DMH_OK) {
      delete H;
      H = nullptr;
    }
  }
  if (rc_out != nullptr) *rc_out = rc;
  return H;
}

int dmh_dt3_insert(void* handle, const double* more, int64_t M) {
  Dt3Handle* H = static_cast<Dt3Handle*>(handle);
  if (H == nullptr || M < 0 || (M > 0 && more == nullptr) || H->N + M > (int64_t)std::numeric_limits<int32_t>::max() / 64)
    return DMH_ERR_ARG;
  if (M == 0) return DMH_OK;
  for (int64_t i = 0; i < 3 * M; ++i)  // (refused before anything is recorded)
    if (!(more[i] == more[i]) || std::fabs(more[i]) == std::numeric_limits<double>::infinity()) return DMH_ERR_ARG;
  H->all.insert(H->all.end(), more, more + 3 * M);
  H->N += M;
  if (!H->s->built || H->s->failed) return H->rebuild();  // nothing sound to add to: start over with everything
  return dt3_insert(*H->s, more, M);
}

int64_t dmh_dt3_points(void* handle) { return handle == nullptr ? -1 : static_cast<Dt3Handle*>(handle)->N; }

int dmh_dt3_cells(void* handle, int32_t* cells, int64_t cap, int64_t* T_out, int64_t* duplicates_out, int64_t* lost_out) {
  Dt3Handle* H = static_cast<Dt3Handle*>(handle);
  if (H == nullptr || cap < 0 || T_out == nullptr || (cap > 0 && cells == nullptr)) return DMH_ERR_ARG;
  *T_out = 0;
  if (duplicates_out != nullptr) *duplicates_out = 0;
  if (lost_out != nullptr) *lost_out = H->N;
  if (!H->s->built) return DMH_OK;
  return dt3_extract(*H->s, cells, cap, T_out, duplicates_out, lost_out);
}

void dmh_dt3_free(void* handle) { delete static_cast<Dt3Handle*>(handle); }

double dmh_orient3d(const double* a, const double* b, const double* c, const double* d) { return orient3d(a, b, c, d); }
double dmh_insphere(const double* a, const double* b, const double* c, const double* d, const double* e) {
  return insphere(a, b, c, d, e);
}

}  // extern "C"
