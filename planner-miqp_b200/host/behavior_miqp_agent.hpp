// behavior_miqp_agent.hpp -- the planning cycle of the reference's BehaviorMiqpAgent without BARK.
//
// Counterpart of bark::models::behavior::BehaviorMiqpAgent (src/behavior_miqp_agent.hpp:27-342,
// src/behavior_miqp_agent.cpp:37-335): one object per ego vehicle that is called once per simulation step with the observed
// world and keeps the MiqpPlanner alive in between -- environment polygon, ego car, the other agents either as dynamic
// obstacles with predicted occupancies (single-agent planning) or as further cars of the joint plan (multi-agent planning),
// receding-horizon warm start, failure handling (EXPIRED status + the last trajectory), the bicycle-model input of the
// first step.
//
// BARK is not part of this build, so the BARK types are replaced by plain data (the reference reads exactly these fields of
// them): ObservedWorld -> ObservedWorldLite (world time, ego and other agents as pose + speed + box shape + centre line of
// their lane corridor + road polygon), dynamic::Trajectory -> rows {t, x, y, theta, v}, the parameter server -> Params with
// the reference's parameter names in the comments.  Predictions of other agents (the reference rolls BARK behaviour models
// forward, CollectDynOccupancies) are taken from ObservedAgent::prediction if given, else constant velocity along the heading.
#pragma once
#include <array>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "miqp_planner.hpp"

namespace miqp {
namespace planner {

struct ObservedAgent {
  int id = 0;
  double x = 0, y = 0, theta = 0, v = 0, a = 0;        // BARK state X, Y, THETA, VEL (+ acceleration of the last action)
  double length = 4.0, width = 2.0;                    // box shape
  std::vector<double> lane_center;                     // x0, y0, x1, y1, ...: centre line of the agent's lane corridor
  std::vector<std::array<double, 3>> prediction;       // optional: x, y, theta per planning step (N entries)
};

struct ObservedWorldLite {
  double time = 0.0;
  ObservedAgent ego;
  std::vector<ObservedAgent> others;
  MatrixXd road_polygon;                               // (k, 2) road corridor polygon of the ego (may be non-convex)
};

enum class BehaviorStatus { NOT_STARTED_YET = 0, VALID = 1, EXPIRED = 2 };

class BehaviorMiqpAgent {
 public:
  struct Params {
    double desired_velocity = 10.0;               // Miqp::DesiredVelocity
    double delta_s_desired_velocity = 5.0;        // Miqp::DeltaSDesiredVelocity
    bool use_box_as_env = false;                  // Miqp::UseBoxAsEnv
    bool write_debug_files = false;               // Miqp::WriteDebugFiles
    std::string debug_file_path, debug_file_prefix;   // Miqp::DebugFilePath / DebugFilePrefix
    bool multi_agent_planning = false;            // Miqp::MultiAgentPlanning
    bool obstacles_soft = true;                   // Miqp::ObstaclesSoft
    double prediction_error_time_percentage = 0.0;    // Miqp::PredictionErrorTimePercentage
  };
  using Trajectory = std::vector<std::array<double, 5>>;   // rows {t, x, y, theta, v}

  BehaviorMiqpAgent(const Settings &settings, const Params &params)
      : settings_(settings), params_(params), planner_(settings), warmstart_type_(settings.warmstartType) {
    if (params_.write_debug_files) planner_.ActivateDebugFileWrite(params_.debug_file_path, params_.debug_file_prefix);
  }

  // One planning cycle (src/behavior_miqp_agent.cpp:137-335).  Returns the ego trajectory; after a failed plan the status is
  // EXPIRED, the solution time NaN and the last valid trajectory is returned.
  Trajectory Plan(double delta_time, const ObservedWorldLite &w) {
    const double t = w.time;
    if (UpdateEnvironmentPolygon(w)) planner_.UpdateConvexifiedMap(envPoly_);
    planner_.SetDoWarmstart(warmstart_type_);

    double s0[6];
    MiqpPlanner::CarStateToMiqpState((float)w.ego.x, (float)w.ego.y, (float)w.ego.theta, (float)w.ego.v, (float)w.ego.a, s0);
    const PolyLine ref = Line(w.ego);
    if (firstrun_) {
      idx_ego_ = planner_.AddCar(s0, ref, params_.desired_velocity, params_.delta_s_desired_velocity, t, true);
      car_idxs_[w.ego.id] = idx_ego_;
      firstrun_ = false;
    } else {
      planner_.UpdateCar(idx_ego_, s0, ref, t, true);
    }

    if (params_.multi_agent_planning) {
      // cars of the previous step go (last one first: only the last car can be removed), the observed ones come in
      while (planner_.GetNrCars() > 1) planner_.RemoveCar(planner_.GetNrCars() - 1);
      car_idxs_.clear();
      car_idxs_[w.ego.id] = idx_ego_;
      for (const ObservedAgent &o : w.others) {
        double sj[6];
        MiqpPlanner::CarStateToMiqpState((float)o.x, (float)o.y, (float)o.theta, (float)o.v, (float)o.a, sj);
        double vdes = o.v;
        if (std::fabs(params_.prediction_error_time_percentage) > 0.01) vdes *= params_.prediction_error_time_percentage;
        car_idxs_[o.id] = planner_.AddCar(sj, Line(o), vdes, params_.delta_s_desired_velocity, t, true);
      }
    } else {
      // other agents as dynamic obstacles with predicted occupancies; the obstacle set is rebuilt when agents appear or vanish
      // (MiqpPlanner::RemoveObstacle is "not implemented" in the reference as well)
      bool same = obstacle_ids_.size() == w.others.size();
      for (const ObservedAgent &o : w.others) same = same && obstacle_ids_.count(o.id) > 0;
      if (!same) { planner_.RemoveAllObstacles(); obstacle_ids_.clear(); }
      for (const ObservedAgent &o : w.others) {
        const std::vector<std::array<double, 3>> poses = Prediction(o);
        auto it = obstacle_ids_.find(o.id);
        if (it == obstacle_ids_.end()) {
          const int id = planner_.AddObstacle(poses, o.length, o.width, params_.obstacles_soft, false);
          if (id >= 0) obstacle_ids_[o.id] = id;
        } else {
          std::vector<MatrixXd> occ = planner_.CreateMiqpObstacle(poses, o.length, o.width);
          planner_.UpdateObstacle(it->second, occ);
        }
      }
    }

    last_planning_success_ = planner_.Plan(t);
    if (!last_planning_success_) {
      status_ = BehaviorStatus::EXPIRED;
      last_solution_time_ = std::nan("");
      return last_trajectory_;
    }
    Trajectory traj = planner_.GetTrajectory(idx_ego_, t);
    // input of the single-track model for the first step: acceleration and steering angle
    double st1[6];
    planner_.Get2ndOrderStateFromSolution(1, idx_ego_, st1);
    const double planned_vel = std::sqrt(st1[1] * st1[1] + st1[4] * st1[4]);
    const double acc = (planned_vel - w.ego.v) / delta_time;
    double delta = 0.0;
    if (traj.size() >= 2) {
      const double theta_dot = AngleDiff(traj[0][3], traj[1][3]) / delta_time;
      delta = std::atan2(theta_dot * settings_.wheelBase, w.ego.v);
    }
    last_action_ = {acc, delta};
    last_trajectories_all_cars_.clear();
    for (const auto &kv : car_idxs_) last_trajectories_all_cars_.push_back(planner_.GetTrajectory(kv.second, t));
    last_solution_time_ = planner_.GetSolutionProperties().time;
    last_trajectory_ = traj;
    status_ = BehaviorStatus::VALID;
    return traj;
  }

  const Trajectory &GetLastTrajectory() const { return last_trajectory_; }
  const std::vector<Trajectory> &GetLastTrajectoriesAllCars() const { return last_trajectories_all_cars_; }
  std::array<double, 2> GetLastAction() const { return last_action_; }
  double GetLastSolutionTime() const { return last_solution_time_; }
  bool GetLastPlanningSuccess() const { return last_planning_success_; }
  BehaviorStatus GetBehaviorStatus() const { return status_; }
  const std::map<int, int> &GetCarIdxs() const { return car_idxs_; }
  const std::map<int, int> &GetObstacleIds() const { return obstacle_ids_; }
  const MatrixXd &GetEnvironmentPolygon() const { return envPoly_; }
  void SetWarmstartType(MiqpPlannerWarmstartType t) { warmstart_type_ = t; }
  MiqpPlanner &GetPlanner() { return planner_; }
  const Settings &GetSettings() const { return settings_; }
  const Params &GetParams() const { return params_; }

 private:
  PolyLine Line(const ObservedAgent &a) const {
    return PolyLine(a.lane_center.data(), (int)a.lane_center.size() / 2, settings_.refLineInterpInc);
  }
  // x, y, theta per planning step: the given prediction (repeated at its end) or constant velocity along the heading
  std::vector<std::array<double, 3>> Prediction(const ObservedAgent &o) const {
    const int N = settings_.nr_steps;
    std::vector<std::array<double, 3>> p(N);
    double scale = 1.0;
    if (std::fabs(params_.prediction_error_time_percentage) > 0.01) scale = params_.prediction_error_time_percentage;
    for (int i = 0; i < N; ++i) {
      if (!o.prediction.empty()) p[i] = o.prediction[std::min<size_t>(i, o.prediction.size() - 1)];
      else { const double d = o.v * scale * settings_.ts * i; p[i] = {o.x + d * std::cos(o.theta), o.y + d * std::sin(o.theta), o.theta}; }
    }
    return p;
  }
  // CalculateEnvironmentPolygon: true if the polygon changed (or is new)
  bool UpdateEnvironmentPolygon(const ObservedWorldLite &w) {
    MatrixXd poly = w.road_polygon;
    if (params_.use_box_as_env && poly.rows() > 0) {
      double x0 = poly(0, 0), x1 = x0, y0 = poly(0, 1), y1 = y0;
      for (int k = 1; k < poly.rows(); ++k) { x0 = std::min(x0, poly(k, 0)); x1 = std::max(x1, poly(k, 0)); y0 = std::min(y0, poly(k, 1)); y1 = std::max(y1, poly(k, 1)); }
      poly.resize(4, 2);
      poly(0, 0) = x0; poly(0, 1) = y0; poly(1, 0) = x1; poly(1, 1) = y0; poly(2, 0) = x1; poly(2, 1) = y1; poly(3, 0) = x0; poly(3, 1) = y1;
    }
    bool same = poly.rows() == envPoly_.rows();
    for (int k = 0; same && k < poly.rows(); ++k) same = poly(k, 0) == envPoly_(k, 0) && poly(k, 1) == envPoly_(k, 1);
    if (same) return false;
    envPoly_ = poly;
    return true;
  }
  static double AngleDiff(double from, double to) {   // bark::geometry::SignedAngleDiff
    double d = std::fmod(to - from + M_PI, 2.0 * M_PI);
    if (d < 0) d += 2.0 * M_PI;
    return d - M_PI;
  }

  Settings settings_;
  Params params_;
  MiqpPlanner planner_;
  MiqpPlannerWarmstartType warmstart_type_;
  MatrixXd envPoly_;
  bool firstrun_ = true, last_planning_success_ = false;
  int idx_ego_ = -1;
  std::map<int, int> car_idxs_, obstacle_ids_;
  Trajectory last_trajectory_;
  std::vector<Trajectory> last_trajectories_all_cars_;
  std::array<double, 2> last_action_ = {0.0, 0.0};
  double last_solution_time_ = 0.0;
  BehaviorStatus status_ = BehaviorStatus::NOT_STARTED_YET;
};

}  // namespace planner
}  // namespace miqp
