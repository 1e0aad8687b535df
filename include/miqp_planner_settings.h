/*
 * miqp_planner_settings.h -- settings block of the planner C API (B200 backend).
 *
 * Binary layout contract: this struct is passed BY VALUE through NewCMiqpPlannerSettings
 * and must stay field-for-field identical (order, types, the 1000-byte path buffer) to the
 * struct Apollo already compiles against, reference src/miqp_planner_settings.h:28-78.
 * Defaults: reference src/miqp_planner_data.hpp:190-251 (mirrored by
 * miqp::planner::DefaultSettings / ApolloDefaultSettings in host/planner_data.hpp).
 *
 * Fields that only steer CPLEX's search (mipdisplay .. mircuts, parallelMode, cplexModelpath,
 * buffer_cplex_outputs) are accepted and stored so that callers need no change; the device
 * branch and bound has no use for them ("settings that are ignored").
 */
#ifndef MIQP_PLANNER_SETTINGS_HEADER
#define MIQP_PLANNER_SETTINGS_HEADER

#ifndef __cplusplus
#include <stdbool.h>
#endif

/* which MIP start the planner hands to the solver (reference CplexWrapper::WarmstartType) */
enum MiqpPlannerWarmstartType {
  NO_WARMSTART = 0,
  RECEDING_HORIZON_WARMSTART = 1, /* shifted previous solution */
  LAST_SOLUTION_WARMSTART = 2,    /* previous solution as written by the solver */
  BOTH_WARMSTART_STRATEGIES = 3
};

/* CPLEX "Parallel" parameter values; the device search is deterministic in every mode */
enum MiqpPlannerParallelMode { DETERMINISTIC = 0, AUTO = 1, OPPORTUNISTIC = -1 };

struct MiqpPlannerSettings {
  /* --- problem size --- */
  int nr_regions;                        /* orientation wedges R (16, 32 or 64 tables exist) */
  int nr_steps;                          /* horizon N */
  int nr_neighbouring_possible_regions;  /* widening of the possible-region mask */
  float ts;                              /* step length [s] */
  int precision;                         /* reals enter the model rounded to precision-2 decimals */
  /* --- model constants --- */
  float constant_agent_safety_distance_slack;
  float minimum_region_change_speed;
  float lambda;                          /* cost share of the ego car in a joint plan */
  float wheelBase;
  float collisionRadius;
  float slackWeight;
  float slackWeightObstacle;
  float jerkWeight;
  float positionWeight;
  float velocityWeight;
  float acclerationWeight;               /* (sic) spelling is part of the ABI */
  float accLonMaxLimit;
  float accLonMinLimit;
  float jerkLonMaxLimit;
  float accLatMinMaxLimit;
  float jerkLatMinMaxLimit;
  /* --- geometry preparation (CPU) --- */
  float simplificationDistanceMap;
  float simplificationDistanceReferenceLine;
  float bufferReference;
  float buffer_for_merging_tolerance;
  float refLineInterpInc;
  int additionalStepsForReferenceLongerHorizon;
  /* --- solver --- */
  float max_solution_time;               /* time limit [s] */
  float relative_mip_gap_tolerance;      /* stop at |bound - incumbent| / (1e-10 + |incumbent|) <= this */
  int mipdisplay;
  int mipemphasis;
  float relobjdif;
  int cutpass;
  int probe;
  int repairtries;
  int rinsheur;
  int varsel;
  int mircuts;
  char cplexModelpath[1000];
  bool useSos;
  bool useBranchingPriorities;
  enum MiqpPlannerWarmstartType warmstartType;
  enum MiqpPlannerParallelMode parallelMode;
  float max_velocity_fitting;            /* vmax of the fitted polynomial tables (10 or 20) */
  bool buffer_cplex_outputs;
  /* --- obstacle region of interest --- */
  bool obstacle_roi_filter;
  float obstacle_roi_behind_distance;
  float obstacle_roi_front_distance;
  float obstacle_roi_side_distance;
};

#ifndef __cplusplus
typedef struct MiqpPlannerSettings MiqpPlannerSettings;
#endif

#endif /* MIQP_PLANNER_SETTINGS_HEADER */
