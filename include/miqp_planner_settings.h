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
  float constant_agent_safety_distance_slack; /* [m] soft safety distance between cars */
  float minimum_region_change_speed;     /* [m/s] below it the orientation region is frozen */
  float lambda;                          /* cost share of the ego car in a joint plan */
  float wheelBase;                       /* [m] rear axle to front axle */
  float collisionRadius;                 /* [m] radius of the two circles (rear / front axle) */
  float slackWeight;                     /* cost of the agent-to-agent safety slack */
  float slackWeightObstacle;             /* cost of giving up a soft obstacle */
  float jerkWeight;                      /* tracking weights: scaled by lambda (ego) or (1 - lambda) / (cars - 1) */
  float positionWeight;                  /*   position error */
  float velocityWeight;                  /*   velocity error */
  float acclerationWeight;               /* (sic) spelling is part of the ABI */
  float accLonMaxLimit;                  /* [m/s^2] straight-driving limits, rotated into every region */
  float accLonMinLimit;                  /* [m/s^2] (negative) */
  float jerkLonMaxLimit;                 /* [m/s^3] symmetric */
  float accLatMinMaxLimit;               /* [m/s^2] symmetric */
  float jerkLatMinMaxLimit;              /* [m/s^3] symmetric */
  /* --- geometry preparation (CPU) --- */
  float simplificationDistanceMap;       /* [m] Douglas-Peucker tolerance of the road polygon */
  float simplificationDistanceReferenceLine; /* [m] the same for reference lines */
  float bufferReference;                 /* [m] cells within this distance of a reference are kept */
  float buffer_for_merging_tolerance;    /* [m] slack when convex cells are merged */
  float refLineInterpInc;                /* [m] resampling step of reference lines */
  int additionalStepsForReferenceLongerHorizon; /* extra steps of the reference used for the possible regions */
  /* --- solver --- */
  float max_solution_time;               /* time limit [s] */
  float relative_mip_gap_tolerance;      /* stop at |bound - incumbent| / (1e-10 + |incumbent|) <= this */
  int mipdisplay;                        /* CPLEX knobs from here to mircuts: stored, not used */
  int mipemphasis;
  float relobjdif;
  int cutpass;
  int probe;
  int repairtries;
  int rinsheur;
  int varsel;
  int mircuts;
  char cplexModelpath[1000];
  bool useSos;                           /* stored; region choices are multi-way disjunctions anyway */
  bool useBranchingPriorities;           /* stored, not used */
  enum MiqpPlannerWarmstartType warmstartType;
  enum MiqpPlannerParallelMode parallelMode;
  float max_velocity_fitting;            /* vmax of the fitted polynomial tables (10 or 20) */
  bool buffer_cplex_outputs;
  /* --- obstacle region of interest --- */
  bool obstacle_roi_filter;              /* drop obstacles outside a rectangle around the ego car */
  float obstacle_roi_behind_distance;    /* [m] */
  float obstacle_roi_front_distance;     /* [m] */
  float obstacle_roi_side_distance;      /* [m] */
};

#ifndef __cplusplus
typedef struct MiqpPlannerSettings MiqpPlannerSettings;
#endif

#endif /* MIQP_PLANNER_SETTINGS_HEADER */
