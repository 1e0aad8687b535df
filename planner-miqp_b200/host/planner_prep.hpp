// planner_prep.hpp -- CPU-side input preparation of the planner facade (a few kB per plan).
//
// What MiqpPlanner needs from the reference's common/ directory, without BARK/boost/Eigen:
//   FittingTables        common/parameter/fitting_polynomial_parameters.hpp:28-90 (selection of the
//                        fitted sin / cos / curvature polynomials; an unknown (R, vmax, vmin)
//                        combination throws std::invalid_argument as at :33-44)
//   ParameterPreparer    common/parameter/parameter_preparer.cpp:37-143 (wedge rays, mean angles,
//                        acc / jerk boxes rotated into every region, in float as the reference)
//   region helpers       common/parameter/regions.cpp:16-127
//   PolyLine + ReferenceTrajectoryGenerator
//                        common/reference/reference_trajectory_generator.cpp:51-148 on a poly-line
//                        centre line with linear interpolation (the BARK spline smoothing is not
//                        reproduced; identical on straight lines, which is what the reference's
//                        own tests pin: common/tests/reference_trajectory_generator_test.cc:22-166)
//   convex polygon helpers (inward offset, separating-axis intersection, point inside)
// The same logic exists in Python (planner-miqp_b200/model_parameters.py) for the scenario
// generators; tests/test_planner_capi.py checks that both produce identical problem data.
#pragma once
#include <cmath>
#include <set>
#include <stdexcept>
#include <utility>
#include <vector>

#include "planner_data.hpp"

namespace miqp {
namespace planner {

// ---------------------------------------------------------------------------------------
struct FittingTableSet { int R, vmax, vmin; std::vector<std::vector<double>> t; };  // t[6][3R]
inline const std::vector<FittingTableSet> &AllFittingTables() {
  static const std::vector<FittingTableSet> all = {
#include "generated/fitting_tables.inc"
  };
  return all;
}

inline const FittingTableSet &SelectFittingTables(int nr_regions, float vmax, float vmin) {
  for (const FittingTableSet &s : AllFittingTables())
    if (s.R == nr_regions && (float)s.vmax == vmax && (float)s.vmin == vmin) return s;
  throw std::invalid_argument("Invalid number of regions or velocity!");
}

inline MatrixXd TableToMatrix(const std::vector<double> &flat, int R) {
  MatrixXd m(R, 3);
  for (int j = 0; j < R; ++j) for (int k = 0; k < 3; ++k) m(j, k) = flat[3 * j + k];
  return m;
}

inline double WrapRadiantTo2Pi(double a) {
  a = std::fmod(a, 2.0 * M_PI);
  if (a < 0) a += 2.0 * M_PI;
  return a;
}

// ---------------------------------------------------------------------------------------
class ParameterPreparer {
 public:
  ParameterPreparer(int nrRegions, float maxVelocityFitting, float minVelocityFitting, float accLonMax,
                    float accLonMin, float jerkLonMax, float accLatMinMax, float jerkLatMinMax)
      : R_(nrRegions), vmax_(maxVelocityFitting), tables_(&SelectFittingTables(nrRegions, maxVelocityFitting, minVelocityFitting)) {
    acc_[0] = accLonMin; acc_[1] = accLonMax; acc_[2] = -accLatMinMax; acc_[3] = accLatMinMax;
    jerk_[0] = -jerkLonMax; jerk_[1] = jerkLonMax; jerk_[2] = -jerkLatMinMax; jerk_[3] = jerkLatMinMax;
    // wedge rays: ray j at angle 2 pi j / R, scaled to vmax; region j lies between ray j and ray j+1
    frac_.resize(R_, 4);
    std::vector<double> cx(R_), cy(R_);
    for (int j = 0; j < R_; ++j) {
      const double alpha = (2.0 * M_PI) * j / R_;   // setLinSpaced(R+1, 0, 2 pi): low + j * (high - low) / R
      cx[j] = (double)vmax_ * std::cos(alpha); cy[j] = (double)vmax_ * std::sin(alpha);
    }
    for (int j = 0; j < R_; ++j) {
      frac_(j, 0) = cx[j]; frac_(j, 1) = cy[j];
      frac_(j, 2) = cx[(j + 1) % R_]; frac_(j, 3) = cy[(j + 1) % R_];
    }
    for (int j = 0; j < R_; ++j) {
      double a1 = WrapRadiantTo2Pi(std::atan2(frac_(j, 1), frac_(j, 0)));
      double a2 = WrapRadiantTo2Pi(std::atan2(frac_(j, 3), frac_(j, 2)));
      if (j + 1 == R_) a2 += 2.0 * M_PI;
      mean_.push_back((a1 + a2) / 2.0);
    }
  }
  const FractionParameters &GetFractionParameters() const { return frac_; }
  const FittingTableSet &Tables() const { return *tables_; }
  float LatAccMax() const { return acc_[3]; }
  int NrRegions() const { return R_; }
  // rows: min_x, max_x, min_y, max_y over the regions
  void AccLimits(std::vector<double> out[4]) const { Limits(acc_, out); }
  void JerkLimits(std::vector<double> out[4]) const { Limits(jerk_, out); }

 private:
  void Limits(const float lim[4], std::vector<double> out[4]) const {
    for (int k = 0; k < 4; ++k) out[k].assign(R_, 0.0);
    for (int j = 0; j < R_; ++j) {
      const float th = (float)mean_[j];             // the reference passes the angle as float
      const double c = std::cos((double)th), s = std::sin((double)th);
      float xs[4], ys[4]; int n = 0;
      for (int a = 1; a >= 0; --a)                   // lon max, lon min
        for (int b = 3; b >= 2; --b) {               // lat max, lat min
          xs[n] = (float)((double)lim[a] * c - (double)lim[b] * s);
          ys[n] = (float)((double)lim[a] * s + (double)lim[b] * c);
          ++n;
        }
      float mnx = xs[0], mxx = xs[0], mny = ys[0], mxy = ys[0];
      for (int k = 1; k < 4; ++k) { mnx = std::min(mnx, xs[k]); mxx = std::max(mxx, xs[k]); mny = std::min(mny, ys[k]); mxy = std::max(mxy, ys[k]); }
      out[0][j] = mnx; out[1][j] = mxx; out[2][j] = mny; out[3][j] = mxy;
    }
  }
  int R_;
  float vmax_;
  const FittingTableSet *tables_;
  float acc_[4], jerk_[4];   // lon min, lon max, lat min, lat max
  FractionParameters frac_;
  std::vector<double> mean_;
};

// all regions whose wedge contains the direction (vx, vy), with the reference's 1e-3 tolerance
inline std::vector<int> CalculateRegionIdx(const FractionParameters &frac, float vx, float vy) {
  const double eps = (double)1e-3f, x = vx, y = vy;
  std::vector<int> out;
  for (int j = 0; j < frac.rows(); ++j) {
    const bool below_ub = frac(j, 2) * y <= frac(j, 3) * x + eps;
    const bool above_lb = frac(j, 0) * y >= frac(j, 1) * x - eps;
    if (below_ub && above_lb) out.push_back(j);
  }
  return out;
}

// widens a contiguous (cyclic) run of possible regions by one region on either side, `expansions` times
inline bool ReserveNeighborRegions(MatrixXi &possible, int car, int expansions) {
  const int s = possible.cols();
  for (int e = 0; e < expansions; ++e) {
    int first = -1, last = -1;
    for (int i = 0; i < s; ++i) {
      if (possible(car, i) != 1) continue;
      if (i >= 1 && possible(car, i - 1) == 0) first = i - 1;
      if (i + 1 < s && possible(car, i + 1) == 0) last = i + 1;
      if (i == 0 && possible(car, s - 1) == 0) first = s - 1;
      if (i == s - 1 && possible(car, 0) == 0) last = 0;
    }
    if (first < 0 || last < 0) return false;
    possible(car, first) = 1; possible(car, last) = 1;
  }
  return true;
}

inline void CalculateRegionCombinations(const std::vector<std::vector<int>> &per_car, std::vector<int> acc,
                                        std::vector<std::vector<int>> &out) {
  if (acc.size() == per_car.size()) { out.push_back(acc); return; }
  for (int r : per_car[acc.size()]) { std::vector<int> n = acc; n.push_back(r); CalculateRegionCombinations(per_car, n, out); }
}

// ---------------------------------------------------------------------------------------
struct Point2 { double x, y; };

class PolyLine {
 public:
  PolyLine() = default;
  // pts: x0, y0, x1, y1, ...; resampled with linear interpolation every interp_inc metres
  PolyLine(const double *pts, int n, double interp_inc) {
    std::vector<Point2> raw(n);
    for (int k = 0; k < n; ++k) raw[k] = {pts[2 * k], pts[2 * k + 1]};
    Init(raw, interp_inc);
  }
  bool Valid() const { return p_.size() >= 2 && s_.back() > 0.0; }
  double Length() const { return s_.empty() ? 0.0 : s_.back(); }
  const std::vector<Point2> &Points() const { return p_; }
  double NearestS(double x, double y) const {
    double best = 1e300, sbest = 0.0;
    for (size_t k = 0; k + 1 < p_.size(); ++k) {
      const double dx = p_[k + 1].x - p_[k].x, dy = p_[k + 1].y - p_[k].y, l2 = dx * dx + dy * dy;
      double t = l2 > 0 ? ((x - p_[k].x) * dx + (y - p_[k].y) * dy) / l2 : 0.0;
      t = std::min(1.0, std::max(0.0, t));
      const double qx = p_[k].x + t * dx - x, qy = p_[k].y + t * dy - y, d = qx * qx + qy * qy;
      if (d < best) { best = d; sbest = s_[k] + t * std::sqrt(l2); }
    }
    return sbest;
  }
  Point2 PointAt(double s) const {
    s = std::min(std::max(s, 0.0), Length());
    const int k = Segment(s);
    const double l = s_[k + 1] - s_[k], t = l > 0 ? (s - s_[k]) / l : 0.0;
    return {p_[k].x + t * (p_[k + 1].x - p_[k].x), p_[k].y + t * (p_[k + 1].y - p_[k].y)};
  }
  double AngleAt(double s) const {
    s = std::min(std::max(s, 0.0), Length());
    const int k = Segment(s);
    return std::atan2(p_[k + 1].y - p_[k].y, p_[k + 1].x - p_[k].x);
  }

 private:
  void Init(const std::vector<Point2> &raw, double inc) {
    std::vector<double> s(raw.size(), 0.0);
    for (size_t k = 1; k < raw.size(); ++k) s[k] = s[k - 1] + std::hypot(raw[k].x - raw[k - 1].x, raw[k].y - raw[k - 1].y);
    p_ = raw; s_ = s;
    if (inc > 0 && raw.size() >= 2 && s.back() > 0) {
      const int n = std::max((int)std::ceil(s.back() / inc), 1);
      std::vector<Point2> q(n + 1);
      size_t seg = 0;
      for (int k = 0; k <= n; ++k) {
        const double sk = (k == n) ? s.back() : s.back() * k / n;
        while (seg + 2 < raw.size() && s[seg + 1] < sk) ++seg;
        const double l = s[seg + 1] - s[seg], t = l > 0 ? (sk - s[seg]) / l : 0.0;
        q[k] = {raw[seg].x + t * (raw[seg + 1].x - raw[seg].x), raw[seg].y + t * (raw[seg + 1].y - raw[seg].y)};
      }
      p_ = q; s_.assign(q.size(), 0.0);
      for (size_t k = 1; k < q.size(); ++k) s_[k] = s_[k - 1] + std::hypot(q[k].x - q[k - 1].x, q[k].y - q[k - 1].y);
    }
  }
  int Segment(double s) const {   // largest k with s_[k] <= s, clamped to a valid segment
    int lo = 0, hi = (int)s_.size() - 1;
    while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (s_[mid] <= s) lo = mid; else hi = mid; }
    return std::min(std::max(lo, 0), (int)p_.size() - 2);
  }
  std::vector<Point2> p_;
  std::vector<double> s_;
};

// rows of a reference trajectory: x, y, heading, speed
struct RefPoint { double x, y, theta, v; };

class ReferenceTrajectoryGenerator {
 public:
  ReferenceTrajectoryGenerator(double dt, int n_points, double interp_inc, double v_des, double delta_s_des)
      : dt_(dt), n_(n_points), inc_(interp_inc), v_des_(v_des), ds_des_(delta_s_des) {}
  void ResetDesiredVelocity(double v, double ds) { v_des_ = v; ds_des_ = ds; }
  const std::vector<RefPoint> &GetLastTrajectory() const { return last_; }
  // x, y, heading, speed of the car now; walks the centre line with a linear speed ramp from the
  // current speed to v_des over delta_s_des metres, and brakes to zero towards the end of the line
  const std::vector<RefPoint> &Generate(double x, double y, double theta, double v0, const PolyLine &line) {
    last_.assign(n_, RefPoint{x, y, theta, v0});
    const double s_start = line.NearestS(x, y), s_end = line.Length();
    double s_des = std::min(s_end, s_start + ds_des_);
    const double vel_0 = (ds_des_ <= 0.0) ? v_des_ : v0;
    double vel_i = vel_0, vel_end = v_des_;
    if (vel_i * n_ * dt_ + s_start > s_end && (vel_end > 0 || s_des > s_end)) { vel_end = 0.0; s_des = s_end; }
    double s_i = s_start;
    for (int i = 1; i < n_; ++i) {
      s_i += vel_i * dt_;
      const Point2 pt = line.PointAt(s_i);
      const double th = line.AngleAt(s_i);
      if ((s_des - s_start) < 1e-2) vel_i = 0.0;
      else if (s_i > s_des) vel_i = vel_end;
      else if (s_i < s_start) vel_i = vel_0;
      else vel_i = (vel_end - vel_0) / (s_des - s_start) * (s_i - s_start) + vel_0;
      last_[i] = RefPoint{pt.x, pt.y, th, vel_i};
    }
    return last_;
  }

 private:
  double dt_; int n_; double inc_, v_des_, ds_des_;
  std::vector<RefPoint> last_;
};

// ---------------------------------------------------------------------------------------
// convex polygon helpers; polygons are (k, 2) vertex matrices without a repeated closing vertex
inline double SignedArea(const MatrixXd &v) {
  double a = 0.0;
  for (int k = 0, n = v.rows(); k < n; ++k) { const int l = (k + 1) % n; a += v(k, 0) * v(l, 1) - v(l, 0) * v(k, 1); }
  return 0.5 * a;
}
inline MatrixXd Reversed(const MatrixXd &v) {
  MatrixXd r(v.rows(), 2);
  for (int k = 0; k < v.rows(); ++k) { r(k, 0) = v(v.rows() - 1 - k, 0); r(k, 1) = v(v.rows() - 1 - k, 1); }
  return r;
}
inline bool IsConvexCcw(const MatrixXd &v, double tol = 1e-9) {
  const int n = v.rows();
  if (n < 3) return false;
  for (int k = 0; k < n; ++k) {
    const int l = (k + 1) % n, m = (k + 2) % n;
    const double cr = (v(l, 0) - v(k, 0)) * (v(m, 1) - v(l, 1)) - (v(l, 1) - v(k, 1)) * (v(m, 0) - v(l, 0));
    if (cr < -tol) return false;
  }
  return true;
}
inline bool PointInConvex(const MatrixXd &v, double x, double y, double tol = 0.0) {
  for (int k = 0, n = v.rows(); k < n; ++k) {
    const int l = (k + 1) % n;
    const double cr = (v(l, 0) - v(k, 0)) * (y - v(k, 1)) - (v(l, 1) - v(k, 1)) * (x - v(k, 0));
    if (cr < -tol) return false;
  }
  return true;
}
// separating-axis test of two convex polygons (either orientation)
inline bool ConvexIntersect(const MatrixXd &a, const MatrixXd &b) {
  const MatrixXd *ps[2] = {&a, &b};
  for (int w = 0; w < 2; ++w) {
    const MatrixXd &p = *ps[w];
    for (int k = 0, n = p.rows(); k < n; ++k) {
      const int l = (k + 1) % n;
      const double nx = p(l, 1) - p(k, 1), ny = -(p(l, 0) - p(k, 0));
      double amin = 1e300, amax = -1e300, bmin = 1e300, bmax = -1e300;
      for (int i = 0; i < a.rows(); ++i) { const double d = nx * a(i, 0) + ny * a(i, 1); amin = std::min(amin, d); amax = std::max(amax, d); }
      for (int i = 0; i < b.rows(); ++i) { const double d = nx * b(i, 0) + ny * b(i, 1); bmin = std::min(bmin, d); bmax = std::max(bmax, d); }
      if (amax < bmin || bmax < amin) return false;
    }
  }
  return true;
}
// inward offset of a convex CCW polygon by r: clip the polygon against every edge moved inwards
inline MatrixXd ShrinkConvexCcw(const MatrixXd &v, double r) {
  std::vector<Point2> poly;
  for (int k = 0; k < v.rows(); ++k) poly.push_back({v(k, 0), v(k, 1)});
  for (int k = 0, n = v.rows(); k < n && !poly.empty(); ++k) {
    const int l = (k + 1) % n;
    const double ex = v(l, 0) - v(k, 0), ey = v(l, 1) - v(k, 1), len = std::hypot(ex, ey);
    if (len == 0) continue;
    const double nx = -ey / len, ny = ex / len;              // inward normal of a CCW polygon
    const double c = nx * v(k, 0) + ny * v(k, 1) + r;        // keep nx x + ny y >= c
    std::vector<Point2> out;
    for (size_t i = 0; i < poly.size(); ++i) {
      const Point2 &p = poly[i], &q = poly[(i + 1) % poly.size()];
      const double dp = nx * p.x + ny * p.y - c, dq = nx * q.x + ny * q.y - c;
      if (dp >= 0) out.push_back(p);
      if ((dp >= 0) != (dq >= 0)) { const double t = dp / (dp - dq); out.push_back({p.x + t * (q.x - p.x), p.y + t * (q.y - p.y)}); }
    }
    poly.swap(out);
  }
  MatrixXd res((int)poly.size(), 2);
  for (size_t k = 0; k < poly.size(); ++k) { res((int)k, 0) = poly[k].x; res((int)k, 1) = poly[k].y; }
  return res;
}
inline double DistPointSegment(double x, double y, double ax, double ay, double bx, double by) {
  const double dx = bx - ax, dy = by - ay, l2 = dx * dx + dy * dy;
  double t = l2 > 0 ? ((x - ax) * dx + (y - ay) * dy) / l2 : 0.0;
  t = std::min(1.0, std::max(0.0, t));
  return std::hypot(ax + t * dx - x, ay + t * dy - y);
}
// does the poly-line through pts, thickened by `buffer`, touch the convex CCW polygon?
inline bool LineBufferTouchesConvex(const std::vector<Point2> &pts, double buffer, const MatrixXd &poly) {
  for (const Point2 &p : pts) {
    if (PointInConvex(poly, p.x, p.y)) return true;
    for (int k = 0, n = poly.rows(); k < n; ++k) {
      const int l = (k + 1) % n;
      if (DistPointSegment(p.x, p.y, poly(k, 0), poly(k, 1), poly(l, 0), poly(l, 1)) <= buffer) return true;
    }
  }
  return false;
}
// Douglas-Peucker simplification of an open poly-line (ConvertToBarkLine applies boost::geometry::simplify)
inline void SimplifyPolyline(const std::vector<Point2> &in, double tol, std::vector<Point2> &out) {
  out.clear();
  if (in.size() < 3 || tol <= 0) { out = in; return; }
  std::vector<char> keep(in.size(), 0);
  keep.front() = keep.back() = 1;
  std::vector<std::pair<int, int>> stack{{0, (int)in.size() - 1}};
  while (!stack.empty()) {
    const auto [a, b] = stack.back(); stack.pop_back();
    double worst = -1.0; int wi = -1;
    for (int k = a + 1; k < b; ++k) {
      const double d = DistPointSegment(in[k].x, in[k].y, in[a].x, in[a].y, in[b].x, in[b].y);
      if (d > worst) { worst = d; wi = k; }
    }
    if (wi >= 0 && worst > tol) { keep[wi] = 1; stack.push_back({a, wi}); stack.push_back({wi, b}); }
  }
  for (size_t k = 0; k < in.size(); ++k) if (keep[k]) out.push_back(in[k]);
}

}  // namespace planner
}  // namespace miqp
