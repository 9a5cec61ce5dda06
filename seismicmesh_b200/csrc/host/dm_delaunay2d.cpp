// Host Delaunay triangulator behind include/distmesh_host.h (2-D).
//
// The reference retriangulates with CGAL on every DistMesh iteration
// (SeismicMesh/generation/cpp/delaunay_class.cpp:45-117 behind mesh_generator.py:466-481); this is
// the from-scratch replacement on raw buffers that keeps the caller's vertex numbering.
//
// Construction: sweep-hull.  A seed triangle near the centre of the bounding box is chosen (the
// point closest to the centre, its nearest neighbour, and the third point giving the smallest
// circumcircle); all other points are inserted in order of distance from the seed circumcentre.
// Every new point lies outside the convex hull of the points inserted so far, so it is connected
// to the hull edges it can see (the hull is a doubly linked list; the first visible edge is found
// through a hash of the hull vertices by pseudo-angle around the centre), and the Delaunay
// property is restored by Lawson flips on a half-edge structure (cell 3*t+k holds vertex k of
// triangle t; half[e] is the opposite half-edge or -1 on the hull).
//
// Exactness: orient2d / incircle evaluate a filtered floating-point determinant and fall back to
// exact arithmetic on floating-point expansions (error-free two_sum / two_prod, sums of
// non-overlapping components) when the filter cannot certify the sign, so ties are recognised as
// ties: flips terminate on co-circular lattices and collinear boundary vertices stay on the hull.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off (the error-free transformations must not be
// contracted or re-associated; never -ffast-math).
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
constexpr double CCW_BOUND = (3.0 + 16.0 * EPS) * EPS;
constexpr double ICC_BOUND = (10.0 + 96.0 * EPS) * EPS;

double orient2d_exact(const double* a, const double* b, const double* c) {
  // (ax-cx)(by-cy) - (ay-cy)(bx-cx) with every difference and product carried exactly
  const Ex<2> acx = ex_diff(a[0], c[0]), acy = ex_diff(a[1], c[1]);
  const Ex<2> bcx = ex_diff(b[0], c[0]), bcy = ex_diff(b[1], c[1]);
  Ex<8> l, r;
  ex_mul(acx, bcy, l);
  ex_mul(acy, bcx, r);
  Ex<16> det;
  ex_sub(l, r, det);
  return ex_sign(det);
}

inline double orient2d(const double* a, const double* b, const double* c) {
  const double dl = (a[0] - c[0]) * (b[1] - c[1]);
  const double dr = (a[1] - c[1]) * (b[0] - c[0]);
  const double det = dl - dr;
  double sum;
  if (dl > 0.0) {
    if (dr <= 0.0) return det;
    sum = dl + dr;
  } else if (dl < 0.0) {
    if (dr >= 0.0) return det;
    sum = -dl - dr;
  } else {
    return det;
  }
  const double bound = CCW_BOUND * sum;
  if (det >= bound || -det >= bound) return det;
  return orient2d_exact(a, b, c);
}

// One term of the in-circle determinant: ((px-dx)^2 + (py-dy)^2) * ((qx-dx)(ry-dy) - (rx-dx)(qy-dy))
struct Rel {
  Ex<2> x, y;
};
void incircle_term(const Rel& p, const Rel& q, const Rel& r, Ex<512>& out) {
  Ex<8> xx, yy, qr, rq;
  ex_mul(p.x, p.x, xx);
  ex_mul(p.y, p.y, yy);
  Ex<16> lift, cross;
  ex_add(xx, yy, lift);
  ex_mul(q.x, r.y, qr);
  ex_mul(r.x, q.y, rq);
  ex_sub(qr, rq, cross);
  ex_mul(lift, cross, out);
}

double incircle_exact(const double* a, const double* b, const double* c, const double* d) {
  // | a-d ; b-d ; c-d | over the columns (x, y, x^2 + y^2): the differences are carried exactly as
  // two-component values, so this is the exact sign of the untranslated 4 x 4 determinant
  const Rel ra{ex_diff(a[0], d[0]), ex_diff(a[1], d[1])};
  const Rel rb{ex_diff(b[0], d[0]), ex_diff(b[1], d[1])};
  const Rel rc{ex_diff(c[0], d[0]), ex_diff(c[1], d[1])};
  static thread_local Ex<512> ta, tb, tc;
  static thread_local Ex<1024> tab;
  static thread_local Ex<1536> det;
  incircle_term(ra, rb, rc, ta);
  incircle_term(rb, rc, ra, tb);
  incircle_term(rc, ra, rb, tc);
  ex_add(ta, tb, tab);
  ex_add(tab, tc, det);
  return ex_sign(det);
}

inline double incircle(const double* a, const double* b, const double* c, const double* d) {
  const double adx = a[0] - d[0], ady = a[1] - d[1];
  const double bdx = b[0] - d[0], bdy = b[1] - d[1];
  const double cdx = c[0] - d[0], cdy = c[1] - d[1];
  const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
  const double alift = adx * adx + ady * ady;
  const double cdxady = cdx * ady, adxcdy = adx * cdy;
  const double blift = bdx * bdx + bdy * bdy;
  const double adxbdy = adx * bdy, bdxady = bdx * ady;
  const double clift = cdx * cdx + cdy * cdy;
  const double det = alift * (bdxcdy - cdxbdy) + blift * (cdxady - adxcdy) + clift * (adxbdy - bdxady);
  const double permanent = (std::fabs(bdxcdy) + std::fabs(cdxbdy)) * alift +
                           (std::fabs(cdxady) + std::fabs(adxcdy)) * blift +
                           (std::fabs(adxbdy) + std::fabs(bdxady)) * clift;
  const double bound = ICC_BOUND * permanent;
  if (det > bound || -det > bound) return det;
  return incircle_exact(a, b, c, d);
}

// ---------------------------------------------------------------------------------------------
// sweep-hull
// ---------------------------------------------------------------------------------------------
struct SweepHull {
  const double* P;
  int64_t n;
  std::vector<int32_t> tri, half;
  int64_t len = 0;  // half-edges in use (3 per triangle)
  std::vector<int32_t> hprev, hnext, htri, hhash, stack;
  std::vector<int32_t> ids;    // rank in the insertion order -> input row
  std::vector<double> sorted;  // coordinates in insertion order
  int64_t hsize = 0;
  double cx = 0.0, cy = 0.0;
  int32_t hstart = 0;

  const double* pt(int64_t i) const { return P + 2 * i; }

  int64_t key(double x, double y) const {
    const double dx = x - cx, dy = y - cy;
    const double den = std::fabs(dx) + std::fabs(dy);
    const double p = den > 0.0 ? dx / den : 0.0;
    const double a = (dy > 0.0 ? 3.0 - p : 1.0 + p) / 4.0;  // monotone in the angle, in [0, 1]
    int64_t k = (int64_t)std::floor(a * (double)hsize);
    if (k < 0) k = 0;
    return k % hsize;
  }
  void link(int32_t a, int32_t b) {
    half[a] = b;
    if (b != -1) half[b] = a;
  }
  int32_t add_triangle(int32_t i0, int32_t i1, int32_t i2, int32_t a, int32_t b, int32_t c) {
    const int32_t t = (int32_t)len;
    tri[t] = i0;
    tri[t + 1] = i1;
    tri[t + 2] = i2;
    link(t, a);
    link(t + 1, b);
    link(t + 2, c);
    len += 3;
    return t;
  }
  // Restore the Delaunay property around half-edge a (and, transitively, around the edges a flip
  // exposes).  Returns the half-edge that now leaves the newest point along the hull side.
  int32_t legalize(int32_t a) {
    stack.clear();
    int32_t ar = 0;
    for (;;) {
      const int32_t b = half[a];
      const int32_t a0 = a - a % 3;
      ar = a0 + (a + 2) % 3;
      if (b == -1) {  // hull edge
        if (stack.empty()) break;
        a = stack.back();
        stack.pop_back();
        continue;
      }
      const int32_t b0 = b - b % 3;
      const int32_t al = a0 + (a + 1) % 3;
      const int32_t bl = b0 + (b + 2) % 3;
      const int32_t p0 = tri[ar], pr = tri[a], pl = tri[al], p1 = tri[bl];
      if (incircle(pt(p0), pt(pr), pt(pl), pt(p1)) > 0.0) {  // p1 strictly inside: flip the edge
        tri[a] = p1;
        tri[b] = p0;
        const int32_t hbl = half[bl];
        // the flip moves the hull half-edge bl (if it is one) to a: repoint its hull vertex, which
        // is the vertex the half-edge starts from (htri[v] always starts at v)
        if (hbl == -1 && htri[p1] == bl) htri[p1] = a;
        link(a, hbl);
        link(b, half[ar]);
        link(ar, bl);
        stack.push_back(b0 + (b + 1) % 3);
      } else {
        if (stack.empty()) break;
        a = stack.back();
        stack.pop_back();
      }
    }
    return ar;
  }

  int64_t dups = 0;  // rows left out as exact duplicates of an earlier row
  int64_t lost = 0;  // rows left out for any other reason (degenerate input, no visible hull edge)

  void run() {
    if (n < 3) {
      lost = n;
      return;
    }
    double minx = std::numeric_limits<double>::infinity(), miny = minx, maxx = -minx, maxy = -minx;
    for (int64_t i = 0; i < n; ++i) {
      minx = std::min(minx, P[2 * i]);
      maxx = std::max(maxx, P[2 * i]);
      miny = std::min(miny, P[2 * i + 1]);
      maxy = std::max(maxy, P[2 * i + 1]);
    }
    const double mx = 0.5 * (minx + maxx), my = 0.5 * (miny + maxy);
    auto d2 = [&](int64_t i, double x, double y) {
      const double dx = P[2 * i] - x, dy = P[2 * i + 1] - y;
      return dx * dx + dy * dy;
    };
    const double inf = std::numeric_limits<double>::infinity();
    // seed: the point closest to the centre, its nearest distinct neighbour, smallest circumcircle
    int64_t i0 = 0, i1 = -1, i2 = -1, r0 = 0, r1 = 0, r2 = 0;
    double best = inf;
    for (int64_t i = 0; i < n; ++i) {
      const double d = d2(i, mx, my);
      if (d < best) best = d, i0 = i;
    }
    best = inf;
    for (int64_t i = 0; i < n; ++i) {
      const double d = d2(i, P[2 * i0], P[2 * i0 + 1]);
      if (d > 0.0 && d < best) best = d, i1 = i;
    }
    if (i1 < 0) {  // all rows coincide
      dups = n - 1;
      lost = 1;
      return;
    }
    double rbest = inf, ccx = 0.0, ccy = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      if (i == i0 || i == i1) continue;
      const double dx = P[2 * i1] - P[2 * i0], dy = P[2 * i1 + 1] - P[2 * i0 + 1];
      const double ex = P[2 * i] - P[2 * i0], ey = P[2 * i + 1] - P[2 * i0 + 1];
      const double bl = dx * dx + dy * dy, cl = ex * ex + ey * ey;
      const double den = dx * ey - dy * ex;
      if (den == 0.0 || cl == 0.0) continue;
      const double s = 0.5 / den;
      const double x = (ey * bl - dy * cl) * s, y = (dx * cl - ex * bl) * s;
      const double r = x * x + y * y;
      if (r < rbest && orient2d(pt(i0), pt(i1), pt(i)) != 0.0) rbest = r, i2 = i, ccx = x, ccy = y;
    }
    if (i2 < 0) {  // collinear input: no triangle
      lost = n;
      return;
    }
    if (orient2d(pt(i0), pt(i1), pt(i2)) < 0.0) std::swap(i1, i2);  // seed counter-clockwise
    cx = P[2 * i0] + ccx;
    cy = P[2 * i0 + 1] + ccy;

    // insertion order: distance from the seed circumcentre; ties by coordinates, so that exact
    // duplicates are neighbours in the order.  From here on the points are addressed by their RANK
    // in this order (coordinates copied in that order: a new point, the hull vertices around it and
    // the triangles it touches were all written recently), and `ids` maps ranks back to input rows.
    struct Item {
      double d, x, y;
      int32_t id;
    };
    std::vector<Item> items(n);
    for (int64_t i = 0; i < n; ++i) items[i] = Item{d2(i, cx, cy), P[2 * i], P[2 * i + 1], (int32_t)i};
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
      if (a.d != b.d) return a.d < b.d;
      if (a.x != b.x) return a.x < b.x;
      if (a.y != b.y) return a.y < b.y;
      return a.id < b.id;
    });
    ids.resize(n);
    sorted.resize(2 * n);
    for (int64_t k = 0; k < n; ++k) {
      ids[k] = items[k].id;
      sorted[2 * k] = items[k].x;
      sorted[2 * k + 1] = items[k].y;
      if (items[k].id == i0) r0 = k;
      if (items[k].id == i1) r1 = k;
      if (items[k].id == i2) r2 = k;
    }
    items.clear();
    items.shrink_to_fit();
    P = sorted.data();
    i0 = r0, i1 = r1, i2 = r2;

    const int64_t maxt = std::max<int64_t>(2 * n - 5, 1);
    tri.assign(3 * maxt, 0);
    half.assign(3 * maxt, -1);
    // Angular hash of the hull vertices.  The front has O(sqrt(n)) vertices, but on an elongated
    // domain the part of it that still advances is squeezed into a narrow range of angles, so the
    // resolution grows with the aspect ratio of the bounding box (a coarse hash lands far from the
    // visible edge and the search below walks the hull: 10 steps per point on the BP2004 shape).
    {
      const double w = maxx - minx, h = maxy - miny;
      const double aspect = (w > 0.0 && h > 0.0) ? std::min(64.0, std::max(w / h, h / w)) : 1.0;
      hsize = std::min<int64_t>(std::max<int64_t>(n, 16), (int64_t)std::ceil(2.0 * aspect * std::sqrt((double)n)));
    }
    hprev.assign(n, 0);
    hnext.assign(n, 0);
    htri.assign(n, 0);
    hhash.assign(hsize, -1);
    stack.reserve(512);

    hstart = (int32_t)i0;
    hnext[i0] = hprev[i2] = (int32_t)i1;
    hnext[i1] = hprev[i0] = (int32_t)i2;
    hnext[i2] = hprev[i1] = (int32_t)i0;
    htri[i0] = 0;
    htri[i1] = 1;
    htri[i2] = 2;
    hhash[key(P[2 * i0], P[2 * i0 + 1])] = (int32_t)i0;
    hhash[key(P[2 * i1], P[2 * i1 + 1])] = (int32_t)i1;
    hhash[key(P[2 * i2], P[2 * i2 + 1])] = (int32_t)i2;
    add_triangle((int32_t)i0, (int32_t)i1, (int32_t)i2, -1, -1, -1);

    double xp = 0.0, yp = 0.0;
    for (int64_t k = 0; k < n; ++k) {
      const int32_t i = (int32_t)k;
      const double x = P[2 * i], y = P[2 * i + 1];
      if (k > 0 && x == xp && y == yp) {  // exact duplicate of the previous row in the order
        ++dups;
        continue;
      }
      xp = x;
      yp = y;
      if (i == i0 || i == i1 || i == i2) continue;

      // a hull vertex near the direction of the new point, then forward to the first visible edge
      int32_t start = 0;
      const int64_t kk = key(x, y);
      for (int64_t j = 0; j < hsize; ++j) {
        start = hhash[(kk + j) % hsize];
        if (start != -1 && start != hnext[start]) break;
      }
      start = hprev[start];
      int32_t e = start, q;
      while (q = hnext[e], orient2d(pt(i), pt(e), pt(q)) >= 0.0) {
        e = q;
        if (e == start) {
          e = -1;
          break;
        }
      }
      if (e == -1) {  // sees no hull edge: not outside the hull (distance ties lost to rounding)
        ++lost;
        continue;
      }
      int32_t t = add_triangle(e, i, hnext[e], -1, -1, htri[e]);
      htri[i] = legalize(t + 2);
      htri[e] = t;
      // forward along the hull while the edges are visible
      int32_t nx = hnext[e];
      while (q = hnext[nx], orient2d(pt(i), pt(nx), pt(q)) < 0.0) {
        t = add_triangle(nx, i, q, htri[i], -1, htri[nx]);
        htri[i] = legalize(t + 2);
        hnext[nx] = nx;  // removed from the hull
        nx = q;
      }
      // backward, if the first visible edge was the one the search started from
      if (e == start) {
        while (q = hprev[e], orient2d(pt(i), pt(q), pt(e)) < 0.0) {
          t = add_triangle(q, i, e, -1, htri[e], htri[q]);
          legalize(t + 2);
          htri[q] = t;
          hnext[e] = e;  // removed from the hull
          e = q;
        }
      }
      hstart = hprev[i] = e;
      hnext[e] = hprev[nx] = i;
      hnext[i] = nx;
      hhash[key(x, y)] = i;
      hhash[key(P[2 * e], P[2 * e + 1])] = e;
    }
  }
};

}  // namespace

extern "C" {

const char* dmh_version(void) { return "distmesh_host 0.2"; }

int64_t dmh_delaunay2d_max_cells(int64_t N) { return N < 3 ? 1 : 2 * N - 5; }

int dmh_delaunay2d(const double* points, int64_t N, int32_t* cells, int64_t cap, int64_t* T_out,
                   int64_t* duplicates_out, int64_t* lost_out) {
  if (N < 0 || cap < 0 || T_out == nullptr || (N > 0 && points == nullptr) || (cap > 0 && cells == nullptr) ||
      N > (int64_t)std::numeric_limits<int32_t>::max() / 6)
    return DMH_ERR_ARG;
  for (int64_t i = 0; i < 2 * N; ++i)
    if (!(points[i] == points[i]) || std::fabs(points[i]) == std::numeric_limits<double>::infinity()) return DMH_ERR_ARG;
  SweepHull s;
  s.P = points;
  s.n = N;
  s.run();
  const int64_t T = s.len / 3;
  *T_out = T;
  if (duplicates_out != nullptr) *duplicates_out = s.dups;
  if (lost_out != nullptr) *lost_out = s.lost;
  if (T > cap) return DMH_ERR_CAPACITY;
  for (int64_t j = 0; j < 3 * T; ++j) cells[j] = s.ids[s.tri[j]];
  dmx::order_cells<3>(cells, T, N);
  return DMH_OK;
}

double dmh_orient2d(const double* a, const double* b, const double* c) { return orient2d(a, b, c); }
double dmh_incircle(const double* a, const double* b, const double* c, const double* d) {
  return incircle(a, b, c, d);
}

}  // extern "C"
