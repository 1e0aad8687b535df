// convexified_map.hpp -- non-convex road polygon -> convex cells shrunk by the collision radius.
//
// Takes the place of the reference's ConvexifiedMap (common/map/convexified_map.{hpp,cpp}: simplify, buffer by -r,
// Voronoi cells of the boundary points, clip, triangulate, greedy convex merge) for callers of MiqpPlanner /
// the C ABI that hand in a whole road polygon.  Same interface (Convert, GetMapConvexPolygons,
// GetIntersectingConvexPolygons, HasValidPolygon, SetMapPolygon), no Boost / BARK: a different, exact decomposition
//
//   1. drop a repeated closing vertex, orient counter-clockwise, drop collinear / duplicate vertices
//      (max_simplify_dist: vertices closer than this to the line through their neighbours are dropped);
//   2. triangulate by ear clipping;
//   3. Hertel-Mehlhorn: remove every diagonal whose removal leaves its two cells convex (at most four times the
//      minimum number of convex pieces);
//   4. shrink: every cell edge that lies on the boundary of the input polygon moves inwards by buffer_radius,
//      the diagonals between cells stay where they are (the cells must keep touching, a point of the car may be in
//      any of them: obstacle_environment_constraints.mod:6-47).  Every point of a shrunk cell keeps the distance
//      buffer_radius from the lines through the boundary edges of its own cell, hence from their end points.
//
// The result differs from the reference's cell layout (any convex cover of the same area is a valid
// MultiEnvironmentConvexPolygon input); the reference pins only success flags for this component
// (common/tests/convexified_map_test.cc:23-208).
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

#include "planner_prep.hpp"

namespace miqp {
namespace common {
namespace map {

using planner::MatrixXd;
using planner::Point2;

typedef std::map<int, MatrixXd> PolygonMap;   // id -> (k, 2) counter-clockwise vertices (miqp::common::geometry::PolygonMap)

class ConvexifiedMap {
 public:
  ConvexifiedMap() {}
  // `params` of the reference (a BARK parameter server) has no counterpart; the numeric arguments keep their meaning
  ConvexifiedMap(const MatrixXd &map_polygon, double buffer_radius, double max_simplify_dist = 0.0, double buffer_reference = 2.0,
                 double buffer_for_merging_tolerance = 1e-9)
      : input_(map_polygon), r_(buffer_radius), simplify_(max_simplify_dist), bufRef_(buffer_reference), tol_(buffer_for_merging_tolerance) {}

  void SetMapPolygon(const MatrixXd &map_polygon) { input_ = map_polygon; cells_.clear(); decomposed_ = false; }
  bool HasValidPolygon() const { return Clean(input_).rows() >= 3; }
  MatrixXd GetMapNonConvexPolygon() const { return input_; }
  PolygonMap GetMapConvexPolygons() const { return cells_; }
  bool IsDecomposed() const { return decomposed_; }

  // shrinks and converts the polygon into convex cells; false if the polygon is degenerate or self-intersecting
  bool Convert() {
    cells_.clear(); decomposed_ = false;
    const MatrixXd v = Clean(input_);
    const int n = v.rows();
    if (n < 3) return false;
    std::vector<std::vector<int>> polys;
    if (planner::IsConvexCcw(v)) {
      std::vector<int> all(n);
      for (int k = 0; k < n; ++k) all[k] = k;
      polys.push_back(all);
    } else {
      if (!Triangulate(v, polys)) return false;
      Merge(v, polys);
    }
    int id = 0;
    double area = 0.0;
    for (const std::vector<int> &poly : polys) {
      MatrixXd cell = ShrinkBoundaryEdges(v, poly);
      if (cell.rows() < 3) continue;        // swallowed by the buffer
      area += std::fabs(planner::SignedArea(cell));
      cells_[id++] = cell;
    }
    decomposed_ = !cells_.empty() && area > 0.0;
    return decomposed_;
  }

  // cells that the reference line, buffered by buffer_reference, touches
  PolygonMap GetIntersectingConvexPolygons(const std::vector<Point2> &reference) const {
    PolygonMap out;
    for (const auto &kv : cells_)
      if (planner::LineBufferTouchesConvex(reference, bufRef_, kv.second)) out.insert(kv);
    return out;
  }

 private:
  static double Cross(const MatrixXd &v, int a, int b, int c) {
    return (v(b, 0) - v(a, 0)) * (v(c, 1) - v(a, 1)) - (v(b, 1) - v(a, 1)) * (v(c, 0) - v(a, 0));
  }
  MatrixXd Clean(const MatrixXd &in) const {
    std::vector<Point2> pts;
    for (int k = 0; k < in.rows(); ++k) {
      if (!pts.empty() && pts.back().x == in(k, 0) && pts.back().y == in(k, 1)) continue;
      pts.push_back({in(k, 0), in(k, 1)});
    }
    if (pts.size() >= 2 && pts.front().x == pts.back().x && pts.front().y == pts.back().y) pts.pop_back();
    double a2 = 0.0;
    for (size_t k = 0; k < pts.size(); ++k) { const Point2 &p = pts[k], &q = pts[(k + 1) % pts.size()]; a2 += p.x * q.y - q.x * p.y; }
    if (a2 < 0) std::reverse(pts.begin(), pts.end());
    // drop vertices that lie (almost) on the line through their neighbours
    bool changed = true;
    while (changed && pts.size() > 3) {
      changed = false;
      for (size_t k = 0; k < pts.size(); ++k) {
        const Point2 &a = pts[(k + pts.size() - 1) % pts.size()], &b = pts[k], &c = pts[(k + 1) % pts.size()];
        const double len = std::hypot(c.x - a.x, c.y - a.y);
        const double d = len > 0 ? std::fabs((c.x - a.x) * (b.y - a.y) - (c.y - a.y) * (b.x - a.x)) / len : std::hypot(b.x - a.x, b.y - a.y);
        if (d <= std::max(simplify_, 1e-12)) { pts.erase(pts.begin() + k); changed = true; break; }
      }
    }
    MatrixXd out((int)pts.size(), 2);
    for (size_t k = 0; k < pts.size(); ++k) { out((int)k, 0) = pts[k].x; out((int)k, 1) = pts[k].y; }
    return out;
  }
  static bool InTriangle(const MatrixXd &v, int a, int b, int c, int p) {
    return Cross(v, a, b, p) >= 0 && Cross(v, b, c, p) >= 0 && Cross(v, c, a, p) >= 0;
  }
  // ear clipping of a simple counter-clockwise polygon
  static bool Triangulate(const MatrixXd &v, std::vector<std::vector<int>> &tris) {
    std::vector<int> idx(v.rows());
    for (int k = 0; k < v.rows(); ++k) idx[k] = k;
    int guard = 0;
    while (idx.size() > 3) {
      bool clipped = false;
      const int m = (int)idx.size();
      for (int k = 0; k < m; ++k) {
        const int a = idx[(k + m - 1) % m], b = idx[k], c = idx[(k + 1) % m];
        if (Cross(v, a, b, c) <= 0) continue;   // reflex or flat corner
        bool empty = true;
        for (int q = 0; q < m && empty; ++q) {
          const int p = idx[q];
          if (p == a || p == b || p == c) continue;
          if (InTriangle(v, a, b, c, p)) empty = false;
        }
        if (!empty) continue;
        tris.push_back({a, b, c});
        idx.erase(idx.begin() + k);
        clipped = true;
        break;
      }
      if (!clipped || ++guard > 100000) return false;   // not a simple polygon
    }
    tris.push_back({idx[0], idx[1], idx[2]});
    return true;
  }
  bool ConvexLoop(const MatrixXd &v, const std::vector<int> &p) const {
    const int m = (int)p.size();
    for (int k = 0; k < m; ++k) if (Cross(v, p[k], p[(k + 1) % m], p[(k + 2) % m]) < -tol_) return false;
    return true;
  }
  // Hertel-Mehlhorn: merge two cells over a shared diagonal whenever the union stays convex
  void Merge(const MatrixXd &v, std::vector<std::vector<int>> &polys) const {
    bool merged = true;
    while (merged) {
      merged = false;
      for (size_t a = 0; a < polys.size() && !merged; ++a)
        for (size_t b = a + 1; b < polys.size() && !merged; ++b) {
          const std::vector<int> &P = polys[a], &Q = polys[b];
          const int np = (int)P.size(), nq = (int)Q.size();
          for (int i = 0; i < np && !merged; ++i)
            for (int j = 0; j < nq && !merged; ++j) {
              // shared edge: P[i] -> P[i+1] equals Q[j+1] -> Q[j] reversed
              if (P[i] != Q[(j + 1) % nq] || P[(i + 1) % np] != Q[j]) continue;
              std::vector<int> u;
              for (int k = 0; k < np; ++k) {
                u.push_back(P[(i + 1 + k) % np]);
                if ((i + 1 + k) % np == i) break;
              }
              // u runs P[i+1] .. P[i]; continue with Q after Q[j+1] = P[i] up to before Q[j] = P[i+1]
              for (int k = 2; k < nq; ++k) u.push_back(Q[(j + k) % nq]);
              if (!ConvexLoop(v, u)) continue;
              polys[a] = u;
              polys.erase(polys.begin() + b);
              merged = true;
            }
        }
    }
  }
  // half-plane clipping of the cell by its boundary edges moved inwards by r_
  MatrixXd ShrinkBoundaryEdges(const MatrixXd &v, const std::vector<int> &poly) const {
    const int n = v.rows(), m = (int)poly.size();
    std::vector<Point2> cell;
    for (int k : poly) cell.push_back({v(k, 0), v(k, 1)});
    for (int k = 0; k < m; ++k) {
      const int a = poly[k], b = poly[(k + 1) % m];
      if ((a + 1) % n != b) continue;   // a diagonal between two cells: stays
      const double ex = v(b, 0) - v(a, 0), ey = v(b, 1) - v(a, 1), len = std::hypot(ex, ey);
      if (len <= 0) continue;
      const double nx = -ey / len, ny = ex / len;                 // inward normal of a counter-clockwise polygon
      const double c = nx * v(a, 0) + ny * v(a, 1) + r_;          // keep nx x + ny y >= c
      std::vector<Point2> out;
      const int cm = (int)cell.size();
      for (int q = 0; q < cm; ++q) {
        const Point2 &p0 = cell[q], &p1 = cell[(q + 1) % cm];
        const double d0 = nx * p0.x + ny * p0.y - c, d1 = nx * p1.x + ny * p1.y - c;
        if (d0 >= 0) out.push_back(p0);
        if ((d0 >= 0) != (d1 >= 0)) { const double t = d0 / (d0 - d1); out.push_back({p0.x + t * (p1.x - p0.x), p0.y + t * (p1.y - p0.y)}); }
      }
      cell.swap(out);
      if (cell.size() < 3) return MatrixXd();
    }
    // drop duplicate points produced by the clipping
    std::vector<Point2> uniq;
    for (const Point2 &p : cell)
      if (uniq.empty() || std::hypot(p.x - uniq.back().x, p.y - uniq.back().y) > 1e-12) uniq.push_back(p);
    if (uniq.size() > 1 && std::hypot(uniq.front().x - uniq.back().x, uniq.front().y - uniq.back().y) <= 1e-12) uniq.pop_back();
    if (uniq.size() < 3) return MatrixXd();
    MatrixXd out((int)uniq.size(), 2);
    for (size_t k = 0; k < uniq.size(); ++k) { out((int)k, 0) = uniq[k].x; out((int)k, 1) = uniq[k].y; }
    if (std::fabs(planner::SignedArea(out)) < 1e-12) return MatrixXd();
    return out;
  }

  MatrixXd input_;
  double r_ = 0.0, simplify_ = 0.0, bufRef_ = 2.0, tol_ = 1e-9;
  PolygonMap cells_;
  bool decomposed_ = false;
};

}  // namespace map
}  // namespace common
}  // namespace miqp
