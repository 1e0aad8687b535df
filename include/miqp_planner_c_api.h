/*
 * miqp_planner_c_api.h -- the planner-level C ABI (libmiqp_planner_c_api.so), B200 backend.
 *
 * Same 17 entry points, argument order and array conventions as the library Apollo links
 * today (reference src/miqp_planner_c_api.h:21-226, implemented at
 * src/miqp_planner_c_api.cpp:20-260), so the shared object is a drop-in: the planner facade
 * behind the handle prepares ModelParameters on the CPU exactly as before and hands every
 * Plan() to the CUDA branch and bound of libmiqp_b200.so (include/miqp_b200.h) instead of
 * CPLEX/OPL.  Two additions at the end (batched planning, last solve statistics) are new.
 *
 * Trajectory arrays are caller-allocated, row-major [N][TRAJECTORY_SIZE].
 * The two `size` out-parameters are C++ references in the reference header; at ABI level a
 * reference is a pointer, so C callers see `int *` and C++ callers keep `int &`.
 */
#ifndef MIQP_PLANNER_C_API_HEADER
#define MIQP_PLANNER_C_API_HEADER

#include "miqp_planner_settings.h"

/* column index inside one trajectory row */
#define TRAJECTORY_TIME_IDX 0
#define TRAJECTORY_X_IDX 1
#define TRAJECTORY_Y_IDX 2
#define TRAJECTORY_VX_IDX 3
#define TRAJECTORY_VY_IDX 4
#define TRAJECTORY_AX_IDX 5
#define TRAJECTORY_AY_IDX 6
#define TRAJECTORY_UX_IDX 7
#define TRAJECTORY_UY_IDX 8
#define TRAJECTORY_SIZE 9

typedef void *CMiqpPlanner;

#ifdef __cplusplus
#define MIQP_SIZE_OUT int &
extern "C" {
#else
#define MIQP_SIZE_OUT int *
#endif

/* life cycle.  NewCMiqpPlanner uses ApolloDefaultSettings unless the library was built with
 * -DPLANNER_MIQP_CAPI_NO_APOLLO (reference src/miqp_planner_c_api.cpp:20-27). */
CMiqpPlanner NewCMiqpPlanner();
CMiqpPlanner NewCMiqpPlannerSettings(struct MiqpPlannerSettings settings);
void DelCMiqpPlanner(CMiqpPlanner c_miqp_planner);

/* cars.  initial_state_in = {x, vx, ax, y, vy, ay}; ref_in = {x0, y0, x1, y1, ...} with
 * ref_size points.  Returns the car index. */
int AddCarCMiqpPlanner(CMiqpPlanner c_miqp_planner, double initial_state_in[], double ref_in[],
                       const int ref_size, double vDes, double deltaSDes, const double timestep,
                       const bool track_reference_positions);
void UpdateCarCMiqpPlanner(CMiqpPlanner c_miqp_planner, int idx, double initial_state_in[],
                           double ref_in[], const int ref_size, const double timestep,
                           bool track_reference_positions);
void UpdateDesiredVelocityCMiqpPlanner(CMiqpPlanner c_miqp_planner, const int carIdx,
                                       const double vDes, const double deltaSDes);

/* one joint MIQP over all cars; true if a feasible plan exists afterwards */
bool PlanCMiqpPlanner(CMiqpPlanner c_miqp_planner, const double timestep);

/* debug files (parameters_<t>.txt etc.) */
void ActivateDebugFileWriteCMiqpPlanner(CMiqpPlanner c_miqp_planner, char path[], char name[]);

/* getters */
int GetNCMiqpPlanner(CMiqpPlanner c_miqp_planner);
float GetTsCMiqpPlanner(CMiqpPlanner c_miqp_planner);
float GetCollisionRadius(CMiqpPlanner c_miqp_planner);
void GetRawCMiqpTrajectoryCMiqpPlanner(CMiqpPlanner c_miqp_planner, int carIdx, double start_time,
                                       double *trajectory, MIQP_SIZE_OUT size);
void GetRawCLastReferenceTrajectoryCMiqpPlaner(CMiqpPlanner c_miqp_planner, int carIdx,
                                               double start_time, double *trajectory,
                                               MIQP_SIZE_OUT size);

/* road polygon as {x0, y0, x1, y1, ...}; false if it could not be decomposed */
bool UpdateConvexifiedMapCMiqpPlaner(CMiqpPlanner c_miqp_planner, double poly_pts[],
                                     const int poly_size);

/* obstacles: four corner points per time step (arrays of length `size`).  Returns the
 * obstacle id or -1 if the obstacle does not touch the drivable area / region of interest. */
int AddObstacleCMiqpPlanner(CMiqpPlanner c_miqp_planner, double p1_x[], double p1_y[], double p2_x[],
                            double p2_y[], double p3_x[], double p3_y[], double p4_x[], double p4_y[],
                            const int size, bool is_static, bool is_soft);
void UpdateObstacleCMiqpPlanner(CMiqpPlanner c_miqp_planner, int id, double p1_x[], double p1_y[],
                                double p2_x[], double p2_y[], double p3_x[], double p3_y[],
                                double p4_x[], double p4_y[], const int size, bool is_static);
void RemoveAllObstaclesCMiqpPlanner(CMiqpPlanner c_miqp_planner);

/* ---- additions of the B200 backend ---------------------------------------------------- */

/* Multi-scenario dispatch: plans `count` independent planners in ONE device batch (the
 * north-star's batched Plan()).  success[k] receives what PlanCMiqpPlanner would have
 * returned for planners[k].  Returns the number of successful plans, -1 on a device error. */
int PlanBatchCMiqpPlanner(CMiqpPlanner *planners, int count, const double timestep, bool *success);

/* When the same planners are planned again in the same order with RECEDING_HORIZON_WARMSTART, PlanBatchCMiqpPlanner takes the
 * MIP starts from the previous incumbents on the device (shifted by one step there, miqp_b200_batch_upload_replan) instead of the
 * host-side shifted vectors; this counts the batches that did so (process-wide). */
long DeviceWarmstartBatchesCMiqpPlanner(void);

/* SolutionProperties of the last Plan() (reference src/cplex_wrapper.hpp:41-52):
 * out = {objective, gap, time [s], status, nodes, rows, binaries, continuous}. */
void GetSolutionPropertiesCMiqpPlanner(CMiqpPlanner c_miqp_planner, double out[8]);

#ifdef __cplusplus
}
#endif
#endif /* MIQP_PLANNER_C_API_HEADER */
