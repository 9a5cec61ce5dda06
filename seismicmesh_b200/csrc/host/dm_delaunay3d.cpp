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
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

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

constexpr int32_t INF = -1;

struct Delaunay3 {
  const double* P = nullptr;  // coordinates in insertion order
  int64_t n = 0;
  std::vector<int32_t> tv, tn;  // 4 vertices / 4 neighbours (opposite the vertex of the same slot) per tetrahedron
  std::vector<int32_t> mark;    // +stamp: in the current cavity, -stamp: tested and not in conflict
  std::vector<uint8_t> dead;
  std::vector<int32_t> freelist, stack, cav, newt;
  struct Facet {
    int32_t t, k, nb;
  };
  std::vector<Facet> bnd;
  struct Slot {
    uint64_t key;
    int32_t t, j, tag;
  };
  std::vector<Slot> table;
  int32_t stamp = 0, last = 0;
  int64_t dups = 0, lost = 0;
  bool failed = false;

  const double* pt(int32_t v) const { return P + 3 * (int64_t)v; }
  int64_t ntets() const { return (int64_t)dead.size(); }

  int32_t new_tet() {
    int32_t t;
    if (!freelist.empty()) {
      t = freelist.back();
      freelist.pop_back();
    } else {
      t = (int32_t)dead.size();
      tv.resize(tv.size() + 4);
      tn.resize(tn.size() + 4);
      mark.push_back(0);
      dead.push_back(0);
    }
    dead[t] = 0;
    mark[t] = 0;
    return t;
  }
  int inf_slot(int32_t t) const {
    const int32_t* v = &tv[4 * (int64_t)t];
    return v[0] == INF ? 0 : v[1] == INF ? 1 : v[2] == INF ? 2 : v[3] == INF ? 3 : -1;
  }
  // orientation of tetrahedron t with the vertex of slot k replaced by the point p
  double orient_with(int32_t t, int k, const double* p) const {
    const int32_t* v = &tv[4 * (int64_t)t];
    const double* q[4] = {k == 0 ? p : pt(v[0]), k == 1 ? p : pt(v[1]), k == 2 ? p : pt(v[2]), k == 3 ? p : pt(v[3])};
    return orient3d(q[0], q[1], q[2], q[3]);
  }
  double insphere_of(int32_t t, const double* p) const {
    const int32_t* v = &tv[4 * (int64_t)t];
    return insphere(pt(v[0]), pt(v[1]), pt(v[2]), pt(v[3]), p);
  }
  bool conflict(int32_t t, const double* p) const {
    const int ki = inf_slot(t);
    if (ki < 0) return insphere_of(t, p) > 0.0;
    const double o = orient_with(t, ki, p);  // > 0: strictly beyond the hull facet
    if (o != 0.0) return o > 0.0;
    return insphere_of(tn[4 * (int64_t)t + ki], p) > 0.0;  // in the facet's plane: as the tetrahedron behind it
  }

  // neighbours of the first five tetrahedra by matching faces
  void link_all() {
    const int64_t m = ntets();
    for (int64_t t = 0; t < m; ++t)
      for (int k = 0; k < 4; ++k) {
        int32_t f[3];
        int c = 0;
        for (int j = 0; j < 4; ++j)
          if (j != k) f[c++] = tv[4 * t + j];
        std::sort(f, f + 3);
        for (int64_t u = 0; u < m; ++u) {
          if (u == t) continue;
          for (int kk = 0; kk < 4; ++kk) {
            int32_t g[3];
            int d = 0;
            for (int j = 0; j < 4; ++j)
              if (j != kk) g[d++] = tv[4 * u + j];
            std::sort(g, g + 3);
            if (f[0] == g[0] && f[1] == g[1] && f[2] == g[2]) tn[4 * t + k] = (int32_t)u;
          }
        }
      }
  }

  bool init(int32_t a, int32_t b, int32_t c, int32_t d) {
    if (orient3d(pt(a), pt(b), pt(c), pt(d)) < 0.0) std::swap(a, b);
    const int32_t t0 = new_tet();
    int32_t* v = &tv[4 * (int64_t)t0];
    v[0] = a, v[1] = b, v[2] = c, v[3] = d;
    for (int k = 0; k < 4; ++k) {  // ghost behind face k: the vertex at infinity in slot k, orientation flipped
      const int32_t g = new_tet();
      int32_t* w = &tv[4 * (int64_t)g];
      const int32_t* s = &tv[4 * (int64_t)t0];
      for (int j = 0; j < 4; ++j) w[j] = s[j];
      w[k] = INF;
      const int x = k == 0 ? 1 : 0, y = k <= 1 ? 2 : 1;  // two of the other slots
      std::swap(w[x], w[y]);
    }
    link_all();
    last = t0;
    return true;
  }

  void match_face(int32_t t, int j, int32_t u, int32_t w) {
    if (u > w) std::swap(u, w);
    const uint64_t key = (uint64_t)(u + 1) * (uint64_t)(n + 2) + (uint64_t)(w + 1);
    const size_t mask = table.size() - 1;
    size_t h = (size_t)(key * 0x9E3779B97F4A7C15ULL >> 20) & mask;
    for (;;) {
      Slot& s = table[h];
      if (s.tag != stamp) {
        s = Slot{key, t, j, stamp};
        return;
      }
      if (s.key == key) {
        tn[4 * (int64_t)t + j] = s.t;
        tn[4 * (int64_t)s.t + s.j] = t;
        return;
      }
      h = (h + 1) & mask;
    }
  }

  void insert(int32_t i) {
    const double* p = pt(i);
    // ---- locate: walk from the tetrahedron created last
    int32_t t = last;
    const int64_t limit = 4 * ntets() + 64;
    bool found = false;
    for (int64_t steps = 0; steps < limit; ++steps) {
      const int ki = inf_slot(t);
      if (ki >= 0) {
        if (conflict(t, p)) {
          found = true;
          break;
        }
        t = tn[4 * (int64_t)t + ki];
        continue;
      }
      bool moved = false;
      for (int kk = 0; kk < 4; ++kk) {
        const int k = (kk + (int)(steps & 3)) & 3;
        if (orient_with(t, k, p) < 0.0) {
          t = tn[4 * (int64_t)t + k];
          moved = true;
          break;
        }
      }
      if (!moved) {
        found = true;
        break;
      }
    }
    if (!found) {  // the walk did not settle: exhaustive search
      for (int64_t u = 0; u < ntets() && !found; ++u)
        if (!dead[u] && conflict((int32_t)u, p)) t = (int32_t)u, found = true;
      if (!found) {
        ++lost;
        return;
      }
    }
    if (inf_slot(t) < 0) {
      const int32_t* v = &tv[4 * (int64_t)t];
      for (int k = 0; k < 4; ++k) {
        const double* q = pt(v[k]);
        if (q[0] == p[0] && q[1] == p[1] && q[2] == p[2]) {  // exact duplicate of an earlier row
          ++dups;
          return;
        }
      }
      if (!conflict(t, p)) {
        ++lost;
        return;
      }
    }
    // ---- cavity: flood fill over the tetrahedra in conflict
    ++stamp;
    cav.clear();
    bnd.clear();
    stack.clear();
    stack.push_back(t);
    mark[t] = stamp;
    while (!stack.empty()) {
      const int32_t c = stack.back();
      stack.pop_back();
      cav.push_back(c);
      for (int k = 0; k < 4; ++k) {
        const int32_t nb = tn[4 * (int64_t)c + k];
        if (mark[nb] == stamp) continue;
        if (mark[nb] != -stamp) {
          if (conflict(nb, p)) {
            mark[nb] = stamp;
            stack.push_back(nb);
            continue;
          }
          mark[nb] = -stamp;
        }
        bnd.push_back(Facet{c, k, nb});
      }
    }
    // ---- the fan: one new tetrahedron per boundary facet
    size_t want = 64;
    while (want < 8 * bnd.size()) want <<= 1;
    if (table.size() < want) table.assign(want, Slot{0, 0, 0, 0});
    int32_t fin = -1;
    newt.clear();
    for (const Facet& f : bnd) {
      const int32_t nt = new_tet();
      int32_t* w = &tv[4 * (int64_t)nt];
      const int32_t* s = &tv[4 * (int64_t)f.t];
      for (int j = 0; j < 4; ++j) w[j] = s[j];
      w[f.k] = i;
      tn[4 * (int64_t)nt + f.k] = f.nb;
      int32_t* back = &tn[4 * (int64_t)f.nb];
      for (int m = 0; m < 4; ++m)
        if (back[m] == f.t) {
          back[m] = nt;
          break;
        }
      if (w[0] != INF && w[1] != INF && w[2] != INF && w[3] != INF) {
        if (!(orient3d(pt(w[0]), pt(w[1]), pt(w[2]), pt(w[3])) > 0.0)) failed = true;  // cavity not star-shaped
        fin = nt;
      }
      for (int j = 0; j < 4; ++j) {
        if (j == f.k) continue;
        int32_t e[2];
        int c = 0;
        for (int m = 0; m < 4; ++m)
          if (m != j && m != f.k) e[c++] = w[m];
        match_face(nt, j, e[0], e[1]);
      }
      newt.push_back(nt);
    }
    for (int32_t c : cav) {
      dead[c] = 1;
      freelist.push_back(c);
    }
    last = fin >= 0 ? fin : newt.back();
  }
};

}  // namespace

extern "C" {

int64_t dmh_delaunay3d_max_cells(int64_t N) { return N < 4 ? 1 : 8 * N + 64; }

int dmh_delaunay3d(const double* points, int64_t N, int32_t* cells, int64_t cap, int64_t* T_out,
                   int64_t* duplicates_out, int64_t* lost_out) {
  if (N < 0 || cap < 0 || T_out == nullptr || (N > 0 && points == nullptr) || (cap > 0 && cells == nullptr) ||
      N > (int64_t)std::numeric_limits<int32_t>::max() / 64)
    return DMH_ERR_ARG;
  *T_out = 0;
  if (duplicates_out != nullptr) *duplicates_out = 0;
  if (lost_out != nullptr) *lost_out = N;
  if (N < 4) return DMH_OK;

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
  for (int64_t i = 0; i < N; ++i) {
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
  std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key != b.key ? a.key < b.key : a.id < b.id; });
  std::vector<double> sorted(3 * N);
  std::vector<int32_t> ids(N);
  for (int64_t r = 0; r < N; ++r) {
    ids[r] = items[r].id;
    for (int k = 0; k < 3; ++k) sorted[3 * r + k] = points[3 * (int64_t)items[r].id + k];
  }
  items.clear();
  items.shrink_to_fit();

  Delaunay3 D;
  D.P = sorted.data();
  D.n = N;
  D.tv.reserve(4 * (7 * N + 64));
  D.tn.reserve(4 * (7 * N + 64));
  // ---- four affinely independent points to start from
  int32_t a = 0, b = -1, c = -1, d = -1;
  for (int64_t r = 1; r < N && b < 0; ++r)
    if (D.pt((int32_t)r)[0] != D.pt(a)[0] || D.pt((int32_t)r)[1] != D.pt(a)[1] || D.pt((int32_t)r)[2] != D.pt(a)[2]) b = (int32_t)r;
  for (int64_t r = 1; r < N && b >= 0 && c < 0; ++r)
    if (r != b && !collinear(D.pt(a), D.pt(b), D.pt((int32_t)r))) c = (int32_t)r;
  for (int64_t r = 1; r < N && c >= 0 && d < 0; ++r)
    if (r != b && r != c && orient3d(D.pt(a), D.pt(b), D.pt(c), D.pt((int32_t)r)) != 0.0) d = (int32_t)r;
  if (d < 0) return DMH_OK;  // fewer than four affinely independent points: no cell, every row lost
  D.init(a, b, c, d);
  for (int64_t r = 1; r < N; ++r) {
    if (r == b || r == c || r == d) continue;
    D.insert((int32_t)r);
    if (D.failed) break;
  }
  if (D.failed) D.lost += 1;
  int64_t T = 0;
  for (int64_t t = 0; t < D.ntets(); ++t)
    if (!D.dead[t] && D.inf_slot((int32_t)t) < 0) ++T;
  *T_out = T;
  if (duplicates_out != nullptr) *duplicates_out = D.dups;
  if (lost_out != nullptr) *lost_out = D.lost;
  if (T > cap) return DMH_ERR_CAPACITY;
  int64_t o = 0;
  for (int64_t t = 0; t < D.ntets(); ++t) {
    if (D.dead[t] || D.inf_slot((int32_t)t) >= 0) continue;
    for (int k = 0; k < 4; ++k) cells[4 * o + k] = ids[D.tv[4 * t + k]];
    ++o;
  }
  dmx::order_cells<4>(cells, T, N);
  return DMH_OK;
}

double dmh_orient3d(const double* a, const double* b, const double* c, const double* d) { return orient3d(a, b, c, d); }
double dmh_insphere(const double* a, const double* b, const double* c, const double* d, const double* e) {
  return insphere(a, b, c, d, e);
}

}  // extern "C"
