// dev_problem.cuh -- device-side view of a batch of plans (one DevProb per plan).
//
// A plan is the content of the reference's ModelParameters
// (src/miqp_planner_data.hpp:99-185).  All plans of a batch are packed by the host into
// one double blob and one int blob (single H2D copy each); DevProb holds dimensions,
// scalars and offsets.  Derived tables (edge normals, per-region rows, stage costs, row
// index prefixes) live in the same blobs and are filled on the device by
// prepare_tables_kernel (formulation.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace miqp {

// big-M constants, cplexmodel/parameters.mod:24-32
constexpr double BIGM_JERK = 10.0;
constexpr double BIGM_FRAC = 1000.0;
constexpr double BIGM_FRONT = 100.0;
constexpr double BIGM_ACC = 10.0;
constexpr double BIGM_KAPPA = 1000.0;
constexpr double BIGM_VEL = 100.0;
constexpr double BIGM_ENV = 10000.0;
constexpr double BIGM_OBS = 10000.0;
constexpr double BIGM_AGENTS = 1000.0;

// continuous core blocks, decision_variables.mod:10-26
enum { B_UX = 0, B_UY, B_PX, B_VX, B_AX, B_PY, B_VY, B_AY, B_XFU, B_XFL, B_YFU, B_YFL };
// stage vector of one car
enum { Y_PX = 0, Y_VX, Y_AX, Y_PY, Y_VY, Y_AY, Y_UX, Y_UY };

// decision codes of a node (one byte per disjunction)
constexpr unsigned char UNDEC = 255;
constexpr unsigned char MODE_FROZEN = 254;
constexpr unsigned char OBS_SOFT = 253;

// row families of the big-M model in OPL instantiation order
enum {
  FAM_IC1 = 0,   // initial_conditions.mod:13-27
  FAM_IC2,       // :29-50
  FAM_IC3,       // :52-60
  FAM_DYN,       // model_region_constraints.mod:11-19
  FAM_BOX,       // :22-39
  FAM_REGION,    // :43-113
  FAM_MINSPEED,  // minimum_speed_constraints.mod:9-44
  FAM_ENV,       // obstacle_environment_constraints.mod:9-36
  FAM_OBS,       // :52-97
  FAM_A2A_ZERO,  // agent_collision_constraints.mod:26-36
  FAM_A2A,       // :38-73
  NUM_FAM
};

struct DevProb {
  int N, R, C, O, L, E, K, P;
  int nEnvEdges, maxEnvEdges;
  double ts, c2, c3;
  double min_vel, max_vel, total_min_acc, total_max_acc, total_min_jerk, total_max_jerk;
  double maximum_slack, w_slack, w_slack_obs, vm, gap_tol;

  // ---- offsets into the double blob (absolute) ----
  long o_safety, o_safety_slack;
  long o_w[8];            // pos_x, vel_x, acc_x, pos_y, vel_y, acc_y, jerk_x, jerk_y  [C]
  long o_wb, o_radius;    // [C]
  long o_x0;              // [C][6]
  long o_front0;          // [C][2]  front axle point at step 0 (host: atan2/cos/sin)
  long o_ref[4];          // x_ref, vx_ref, y_ref, vy_ref [C][N]
  long o_lim[8];          // min_acc_x, max_acc_x, min_acc_y, max_acc_y, min_jerk_x, max_jerk_x, min_jerk_y, max_jerk_y [C][R]
  long o_obs_edges;       // [O][N][L][4]
  long o_env_edges;       // [nEnvEdges][4]
  long o_frac;            // [R][4]
  long o_poly[6];         // sint_ub, sint_lb, coss_ub, coss_lb, kappa_max, kappa_min [R][3]
  // derived (device-filled)
  long o_envtab;          // [nEnvEdges][3]  unit edge direction ex,ey and ec = ex*y1 - x1*ey
  long o_obstab;          // [O][N][L][3]
  long o_modetab;         // [R][4][5]  wedge1, wedge2, kappa_max, kappa_min rows: a_vx,a_ax,a_vy,a_ay,rhs (normalised)
  long o_fronttab;        // [C][R][12] wb*poly: fxu[3], fxl[3], fyu[3], fyl[3]
  long o_cost;            // [C][N][16] diag q[8] (=2w) and linear c[8] (=-2w ref) per stage
  long o_wtab;            // [C][N][14] reach of the cost-only LQ problem: Wx (3x3 packed), Wy, var(ux), var(uy)
  double cost_const;      // sum w ref^2 (filled by prepare_tables)

  // ---- offsets into the int blob ----
  long o_initreg;         // [C] 1-based
  long o_possible;        // [C][R]
  long o_obs_nedges;      // [O][N]
  long o_obs_soft;        // [O]
  long o_env_off;         // [E+1]
  long o_alt;             // [C][4R] mode alternatives (j*4+h), host-filled
  long o_nalt;            // [C]
  // derived (device-filled)
  long o_posspre;         // [C][R+1] number of possible regions before j
  long o_obsrowpre;       // [N][O+1] obstacle rows of one car before obstacle o at step i
  long o_obsnnzpre;       // [N][O+1]
  long o_obsstep_rows;    // [N+1] obstacle rows (all cars) before step i
  long o_obsstep_nnz;     // [N+1]

  // ---- column layout (decision_variables.mod order) ----
  int base_nwe, base_ar, base_rcna, base_dcc, base_dcf, base_so, base_sof, base_c2c, base_sv, ncols;

  // ---- big-M model: family bases (device-filled) ----
  long fam_row[NUM_FAM + 1];
  long fam_nnz[NUM_FAM + 1];
  long region_rows_car[9], region_nnz_car[9];  // prefix over cars (C <= 8) of one step's region rows

  // ---- batch-level placement of this plan's outputs ----
  long row_base, nnz_base;   // into batch CSR arrays (assemble)
  long x_base;               // into batch column-vector array
  // ---- branch and bound ----
  int ndec, ndec_pad, off_mode, off_env, off_obs, off_pair;
  int kmax;                  // inequality row slots per stage of a node relaxation
};

struct DevBatch {
  const DevProb *prob;
  double *dblob;
  int *iblob;
  int count;
};

__host__ __device__ inline int col_core(const DevProb &p, int blk, int c, int i) { return (blk * p.C + c) * p.N + i; }
__host__ __device__ inline int col_nwe(const DevProb &p, int k, int c, int e, int i) { return p.base_nwe + ((k * p.C + c) * p.E + e) * p.N + i; }
__host__ __device__ inline int col_ar(const DevProb &p, int c, int i, int j) { return p.base_ar + (c * p.N + i) * p.R + j; }
__host__ __device__ inline int col_rcna(const DevProb &p, int k, int c, int i) { return p.base_rcna + (k * p.C + c) * p.N + i; }
__host__ __device__ inline int col_dcc(const DevProb &p, int c, int o, int i, int e) { return p.base_dcc + ((c * p.O + o) * p.N + i) * p.L + e; }
__host__ __device__ inline int col_dcf(const DevProb &p, int c, int o, int i, int e, int f) { return p.base_dcf + (((c * p.O + o) * p.N + i) * p.L + e) * 4 + f; }
__host__ __device__ inline int col_so(const DevProb &p, int c, int o, int i) { return p.base_so + (c * p.O + o) * p.N + i; }
__host__ __device__ inline int col_sof(const DevProb &p, int c, int o, int i, int f) { return p.base_sof + ((c * p.O + o) * p.N + i) * 4 + f; }
__host__ __device__ inline int col_c2c(const DevProb &p, int k1, int k2, int i, int s) { return p.base_c2c + ((k1 * p.K + k2) * p.N + i) * 16 + s; }
__host__ __device__ inline int col_sv(const DevProb &p, int k1, int k2, int i, int s) { return p.base_sv + ((k1 * p.K + k2) * p.N + i) * 4 + s; }

}  // namespace miqp
