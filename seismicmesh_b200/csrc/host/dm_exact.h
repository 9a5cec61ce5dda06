// Exact arithmetic on floating-point expansions for the geometric predicates of the host
// triangulators (dm_delaunay2d.cpp, dm_delaunay3d.cpp).  Error-free transformations (two_sum,
// two_prod through FMA, two_diff) and sums of non-overlapping components with zero elimination.
// Compile with -ffp-contract=off and never with -ffast-math.
#pragma once
#include <cmath>
#include <vector>

namespace dmx {

inline void two_sum(double a, double b, double& x, double& y) {
  x = a + b;
  const double bv = x - a;
  const double av = x - bv;
  y = (a - av) + (b - bv);
}
inline void two_prod(double a, double b, double& x, double& y) {
  x = a * b;
  y = std::fma(a, b, -x);  // exact residual of the product
}

inline void two_diff(double a, double b, double& x, double& y) {
  x = a - b;
  const double bv = a - x;
  const double av = x + bv;
  y = (a - av) + (bv - b);
}

// A value held exactly as a sum of doubles: non-overlapping components of increasing magnitude,
// zeros eliminated (a zero value is the single component 0.0).  Because zeros are dropped, the cost
// of every operation follows the number of components that are actually needed: differences of
// nearby coordinates are exact in one double, and the whole determinant then stays a few dozen
// components long.
template <int CAP>
struct Ex {
  double c[CAP];
  int n;
};

// h = e + f (linear-time merge by magnitude, then one carry sweep); h must not alias e or f
inline int ex_sum(const double* e, int en, const double* f, int fn, double* h) {
  int ei = 0, fi = 0, hn = 0;
  auto take = [&]() {  // the next component in order of increasing magnitude
    if (fi >= fn || (ei < en && std::fabs(e[ei]) <= std::fabs(f[fi]))) return e[ei++];
    return f[fi++];
  };
  double q = take();
  while (ei < en || fi < fn) {
    double s, r;
    two_sum(q, take(), s, r);
    q = s;
    if (r != 0.0) h[hn++] = r;
  }
  if (q != 0.0 || hn == 0) h[hn++] = q;
  return hn;
}

// h = e * b; h holds up to 2 * en components and must not alias e
inline int ex_scale(const double* e, int en, double b, double* h) {
  int hn = 0;
  double q, lo;
  two_prod(e[0], b, q, lo);
  if (lo != 0.0) h[hn++] = lo;
  for (int i = 1; i < en; ++i) {
    double t, tl, s, r;
    two_prod(e[i], b, t, tl);
    two_sum(q, tl, s, r);
    if (r != 0.0) h[hn++] = r;
    two_sum(t, s, q, r);
    if (r != 0.0) h[hn++] = r;
  }
  if (q != 0.0 || hn == 0) h[hn++] = q;
  return hn;
}

template <int A, int B, int R>
inline void ex_add(const Ex<A>& x, const Ex<B>& y, Ex<R>& r) {
  static_assert(R >= A + B, "capacity");
  r.n = ex_sum(x.c, x.n, y.c, y.n, r.c);
}
template <int A, int B, int R>
inline void ex_sub(const Ex<A>& x, const Ex<B>& y, Ex<R>& r) {
  static_assert(R >= A + B, "capacity");
  Ex<B> m;
  for (int i = 0; i < y.n; ++i) m.c[i] = -y.c[i];
  r.n = ex_sum(x.c, x.n, m.c, y.n, r.c);
}
// r = x * y: the partial products x * y_i are added up one by one
template <int A, int B, int R>
inline void ex_mul(const Ex<A>& x, const Ex<B>& y, Ex<R>& r) {
  static_assert(R >= 2 * A * B, "capacity");
  Ex<R> acc;
  Ex<2 * A> part;
  r.n = ex_scale(x.c, x.n, y.c[0], r.c);
  for (int i = 1; i < y.n; ++i) {
    part.n = ex_scale(x.c, x.n, y.c[i], part.c);
    acc.n = ex_sum(r.c, r.n, part.c, part.n, acc.c);
    for (int k = 0; k < acc.n; ++k) r.c[k] = acc.c[k];
    r.n = acc.n;
  }
}
template <int CAP>
inline double ex_sign(const Ex<CAP>& x) {
  return x.c[x.n - 1];  // the most significant component carries the sign
}
// a - b, exactly
inline Ex<2> ex_diff(double a, double b) {
  Ex<2> r;
  double hi, lo;
  two_diff(a, b, hi, lo);
  r.n = 0;
  if (lo != 0.0) r.c[r.n++] = lo;
  if (hi != 0.0 || r.n == 0) r.c[r.n++] = hi;
  return r;
}

// ---- dynamically sized values (the 3-D predicates: sizes depend on how many components survive)
struct XV {
  std::vector<double> c;
  XV() {}
  explicit XV(double v) : c(1, v) {}
  template <int CAP>
  explicit XV(const Ex<CAP>& e) : c(e.c, e.c + e.n) {}
  double sign() const { return c.back(); }
};
inline XV xv_add(const XV& x, const XV& y) {
  XV r;
  r.c.resize(x.c.size() + y.c.size());
  r.c.resize(ex_sum(x.c.data(), (int)x.c.size(), y.c.data(), (int)y.c.size(), r.c.data()));
  return r;
}
inline XV xv_neg(const XV& x) {
  XV r = x;
  for (double& v : r.c) v = -v;
  return r;
}
inline XV xv_sub(const XV& x, const XV& y) { return xv_add(x, xv_neg(y)); }
inline XV xv_mul(const XV& x, const XV& y) {
  const XV& big = x.c.size() >= y.c.size() ? x : y;
  const XV& small = x.c.size() >= y.c.size() ? y : x;
  XV r, part, acc;
  r.c.resize(2 * big.c.size());
  r.c.resize(ex_scale(big.c.data(), (int)big.c.size(), small.c[0], r.c.data()));
  part.c.resize(2 * big.c.size());
  for (size_t i = 1; i < small.c.size(); ++i) {
    const int pn = ex_scale(big.c.data(), (int)big.c.size(), small.c[i], part.c.data());
    acc.c.resize(r.c.size() + pn);
    acc.c.resize(ex_sum(r.c.data(), (int)r.c.size(), part.c.data(), pn, acc.c.data()));
    r.c.swap(acc.c);
  }
  return r;
}

}  // namespace dmx
