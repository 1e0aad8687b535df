// dat_reader.hpp -- reads an OPL .dat parameter file into ModelParameters.
//
// The reference can run its solver from a .dat file instead of C++ inputs (CplexWrapper::DATFILE,
// src/cplex_wrapper.hpp:63,196-198; the fixtures cplexmodel/*.dat and the parameters_<t>.txt dumps of
// src/cplex_wrapper.cpp:141-155 share the dialect).  Grammar handled here: `name = value;` with value a
// number, `[ ... ]` array (nested, elements separated by blanks and/or commas), `{ <k, x1, y1, x2, y2> ... }`
// edge set, /* */ and // comments.  Polygons come back as vertex matrices (first point of every edge).
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "planner_data.hpp"

namespace miqp {
namespace planner {
namespace datio {

struct Value {   // number | list | set of tuples
  enum Kind { NUM, LIST, SET, TUPLE } kind = NUM;
  double num = 0.0;
  std::vector<Value> items;
};

class Parser {
 public:
  explicit Parser(const std::string &text) : t_(text) {}
  std::map<std::string, Value> ParseAll() {
    std::map<std::string, Value> out;
    for (;;) {
      Skip();
      if (pos_ >= t_.size()) break;
      if (t_[pos_] == ';') { ++pos_; continue; }
      const std::string name = Ident();
      Skip(); Expect('=');
      out[name] = ParseValue();
      Skip();
      if (pos_ < t_.size() && t_[pos_] == ';') ++pos_;
    }
    return out;
  }

 private:
  void Skip() {
    for (;;) {
      while (pos_ < t_.size() && (std::isspace((unsigned char)t_[pos_]) || t_[pos_] == ',')) ++pos_;
      if (pos_ + 1 < t_.size() && t_[pos_] == '/' && t_[pos_ + 1] == '*') {
        const size_t e = t_.find("*/", pos_ + 2);
        pos_ = (e == std::string::npos) ? t_.size() : e + 2;
      } else if (pos_ + 1 < t_.size() && t_[pos_] == '/' && t_[pos_ + 1] == '/') {
        const size_t e = t_.find('\n', pos_);
        pos_ = (e == std::string::npos) ? t_.size() : e + 1;
      } else break;
    }
  }
  void Expect(char c) {
    Skip();
    if (pos_ >= t_.size() || t_[pos_] != c) throw std::runtime_error(std::string("dat: expected '") + c + "' at offset " + std::to_string(pos_));
    ++pos_;
  }
  std::string Ident() {
    const size_t b = pos_;
    while (pos_ < t_.size() && (std::isalnum((unsigned char)t_[pos_]) || t_[pos_] == '_')) ++pos_;
    if (b == pos_) throw std::runtime_error("dat: identifier expected at offset " + std::to_string(pos_));
    return t_.substr(b, pos_ - b);
  }
  Value ParseValue() {
    Skip();
    if (pos_ >= t_.size()) throw std::runtime_error("dat: value expected at end of file");
    const char c = t_[pos_];
    if (c == '[' || c == '{' || c == '<') {
      const char close = (c == '[') ? ']' : (c == '{') ? '}' : '>';
      Value v; v.kind = (c == '[') ? Value::LIST : (c == '{') ? Value::SET : Value::TUPLE;
      ++pos_;
      for (;;) {
        Skip();
        if (pos_ >= t_.size()) throw std::runtime_error("dat: unterminated bracket");
        if (t_[pos_] == close) { ++pos_; break; }
        v.items.push_back(ParseValue());
      }
      return v;
    }
    char *end = nullptr;
    const double x = std::strtod(t_.c_str() + pos_, &end);
    if (end == t_.c_str() + pos_) throw std::runtime_error("dat: number expected at offset " + std::to_string(pos_));
    pos_ = (size_t)(end - t_.c_str());
    Value v; v.num = x;
    return v;
  }
  const std::string &t_;
  size_t pos_ = 0;
};

inline double Num(const std::map<std::string, Value> &m, const char *k, double dflt, bool required = true) {
  const auto it = m.find(k);
  if (it == m.end()) { if (required) throw std::runtime_error(std::string("dat: missing ") + k); return dflt; }
  return it->second.num;
}
inline VectorXd Vec(const std::map<std::string, Value> &m, const char *k) {
  const auto it = m.find(k);
  if (it == m.end()) throw std::runtime_error(std::string("dat: missing ") + k);
  VectorXd v((int)it->second.items.size());
  for (int i = 0; i < v.size(); ++i) v(i) = it->second.items[i].num;
  return v;
}
inline MatrixXd Mat(const std::map<std::string, Value> &m, const char *k) {
  const auto it = m.find(k);
  if (it == m.end()) throw std::runtime_error(std::string("dat: missing ") + k);
  const std::vector<Value> &rows = it->second.items;
  const int r = (int)rows.size(), c = r ? (int)rows[0].items.size() : 0;
  MatrixXd out(r, c);
  for (int i = 0; i < r; ++i) {
    if ((int)rows[i].items.size() != c) throw std::runtime_error(std::string("dat: ragged matrix ") + k);
    for (int j = 0; j < c; ++j) out(i, j) = rows[i].items[j].num;
  }
  return out;
}
// { <k, x1, y1, x2, y2> ... } -> (k, 2) vertex matrix
inline MatrixXd Polygon(const Value &set) {
  MatrixXd v((int)set.items.size(), 2);
  for (int e = 0; e < v.rows(); ++e) {
    const Value &t = set.items[e];
    if (t.items.size() != 5) throw std::runtime_error("dat: edge tuples need 5 entries");
    v(e, 0) = t.items[1].num; v(e, 1) = t.items[2].num;
  }
  return v;
}

inline void ReadParametersDat(const std::string &path, ModelParameters &p) {
  std::ifstream f(path);
  if (!f.good()) throw std::runtime_error("dat: cannot open " + path);
  std::stringstream ss; ss << f.rdbuf();
  const std::string text = ss.str();
  const std::map<std::string, Value> m = Parser(text).ParseAll();
  p = ModelParameters();
  p.NumSteps = (int)Num(m, "NumSteps", 0); p.nr_environments = (int)Num(m, "nr_environments", 0);
  p.nr_regions = (int)Num(m, "nr_regions", 0); p.nr_obstacles = (int)Num(m, "nr_obstacles", 0);
  p.max_lines_obstacles = (int)Num(m, "max_lines_obstacles", 0); p.NumCars = (int)Num(m, "NumCars", 0);
  p.max_solution_time = (float)Num(m, "max_solution_time", 10); p.relative_mip_gap_tolerance = (float)Num(m, "relative_mip_gap_tolerance", 1e-4);
  p.mipdisplay = (int)Num(m, "mipdisplay", 2, false); p.mipemphasis = (int)Num(m, "mipemphasis", 0, false);
  p.relobjdif = (float)Num(m, "relobjdif", 0, false); p.cutpass = (int)Num(m, "cutpass", 0, false);
  p.probe = (int)Num(m, "probe", 0, false); p.repairtries = (int)Num(m, "repairtries", 0, false);
  p.rinsheur = (int)Num(m, "rinsheur", 0, false); p.varsel = (int)Num(m, "varsel", 0, false);
  p.mircuts = (int)Num(m, "mircuts", 0, false); p.parallelmode = (int)Num(m, "parallelmode", 0, false);
  p.ts = (float)Num(m, "ts", 0);
  p.min_vel_x_y = (float)Num(m, "min_vel_x_y", 0); p.max_vel_x_y = (float)Num(m, "max_vel_x_y", 0);
  p.total_min_acc = (float)Num(m, "total_min_acc", 0); p.total_max_acc = (float)Num(m, "total_max_acc", 0);
  p.total_min_jerk = (float)Num(m, "total_min_jerk", 0); p.total_max_jerk = (float)Num(m, "total_max_jerk", 0);
  p.agent_safety_distance = Vec(m, "agent_safety_distance"); p.agent_safety_distance_slack = Vec(m, "agent_safety_distance_slack");
  p.maximum_slack = (float)Num(m, "maximum_slack", 0);
  p.WEIGHTS_POS_X = Vec(m, "WEIGHTS_POS_X"); p.WEIGHTS_VEL_X = Vec(m, "WEIGHTS_VEL_X"); p.WEIGHTS_ACC_X = Vec(m, "WEIGHTS_ACC_X");
  p.WEIGHTS_POS_Y = Vec(m, "WEIGHTS_POS_Y"); p.WEIGHTS_VEL_Y = Vec(m, "WEIGHTS_VEL_Y"); p.WEIGHTS_ACC_Y = Vec(m, "WEIGHTS_ACC_Y");
  p.WEIGHTS_JERK_X = Vec(m, "WEIGHTS_JERK_X"); p.WEIGHTS_JERK_Y = Vec(m, "WEIGHTS_JERK_Y");
  p.WEIGHTS_SLACK = (float)Num(m, "WEIGHTS_SLACK", 0); p.WEIGHTS_SLACK_OBSTACLE = (float)Num(m, "WEIGHTS_SLACK_OBSTACLE", 0);
  p.WheelBase = Vec(m, "WheelBase"); p.CollisionRadius = Vec(m, "CollisionRadius");
  p.IntitialState = Mat(m, "IntitialState");
  p.x_ref = Mat(m, "x_ref"); p.vx_ref = Mat(m, "vx_ref"); p.y_ref = Mat(m, "y_ref"); p.vy_ref = Mat(m, "vy_ref");
  p.acc_limit_params.min_x = Mat(m, "min_acc_x"); p.acc_limit_params.max_x = Mat(m, "max_acc_x");
  p.acc_limit_params.min_y = Mat(m, "min_acc_y"); p.acc_limit_params.max_y = Mat(m, "max_acc_y");
  p.jerk_limit_params.min_x = Mat(m, "min_jerk_x"); p.jerk_limit_params.max_x = Mat(m, "max_jerk_x");
  p.jerk_limit_params.min_y = Mat(m, "min_jerk_y"); p.jerk_limit_params.max_y = Mat(m, "max_jerk_y");
  { const VectorXd r = Vec(m, "initial_region"); p.initial_region.resize(r.size()); for (int c = 0; c < r.size(); ++c) p.initial_region(c) = (int)r(c); }
  { const MatrixXd r = Mat(m, "possible_region"); p.possible_region.resize(r.rows(), r.cols());
    for (int c = 0; c < r.rows(); ++c) for (int j = 0; j < r.cols(); ++j) p.possible_region(c, j) = (int)r(c, j); }
  p.ObstacleConvexPolygon.clear(); p.obstacle_is_soft.clear();
  const auto ito = m.find("ObstacleConvexPolygon");
  if (ito != m.end())
    for (const Value &obst : ito->second.items) {
      std::vector<MatrixXd> steps;
      for (const Value &set : obst.items) steps.push_back(Polygon(set));
      p.ObstacleConvexPolygon.push_back(steps);
    }
  const auto its = m.find("obstacle_is_soft");
  for (size_t o = 0; o < p.ObstacleConvexPolygon.size(); ++o)
    p.obstacle_is_soft.push_back((its != m.end() && o < its->second.items.size()) ? (int)its->second.items[o].num : 0);
  p.MultiEnvironmentConvexPolygon.clear(); p.environmentPolygonIds.clear();
  const auto ite = m.find("MultiEnvironmentConvexPolygon");
  if (ite != m.end())
    for (const Value &set : ite->second.items) {
      p.MultiEnvironmentConvexPolygon.push_back(Polygon(set));
      p.environmentPolygonIds.push_back((PolygonId)p.environmentPolygonIds.size());
    }
  if ((int)p.ObstacleConvexPolygon.size() != p.nr_obstacles || (int)p.MultiEnvironmentConvexPolygon.size() != p.nr_environments)
    throw std::runtime_error("dat: polygon counts do not match nr_obstacles / nr_environments");
  p.fraction_parameters = Mat(m, "fraction_parameters");
  p.minimum_region_change_speed = (float)Num(m, "minimum_region_change_speed", 0);
  p.poly_orientation_params.POLY_SINT_UB = Mat(m, "POLY_SINT_UB"); p.poly_orientation_params.POLY_SINT_LB = Mat(m, "POLY_SINT_LB");
  p.poly_orientation_params.POLY_COSS_UB = Mat(m, "POLY_COSS_UB"); p.poly_orientation_params.POLY_COSS_LB = Mat(m, "POLY_COSS_LB");
  p.poly_curvature_params.POLY_KAPPA_AX_MAX = Mat(m, "POLY_KAPPA_AX_MAX"); p.poly_curvature_params.POLY_KAPPA_AX_MIN = Mat(m, "POLY_KAPPA_AX_MIN");
}

}  // namespace datio
}  // namespace planner
}  // namespace miqp
