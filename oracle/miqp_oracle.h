/*
 * miqp_oracle.h -- CPU oracle for the planner-miqp MIQP hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (planner-miqp_b200/) never links or calls it.
 *
 * What it restates (reference paths relative to /root/reference):
 *   - the MIQP formulation of cplexmodel/{parameters,initialization,decision_variables,
 *     objective_function,initial_conditions,model_region_constraints,
 *     minimum_speed_constraints,obstacle_environment_constraints,
 *     agent_collision_constraints}.mod, row by row in OPL instantiation order;
 *   - the polygon -> closed edge list convention of
 *     src/model_input_data_source.cpp:150-178;
 *   - the solve that the reference delegates to IBM CPLEX 12.10 (util/deps.bzl:69-92,
 *     call site src/cplex_wrapper.cpp:158-185).  CPLEX is a proprietary third-party
 *     dependency that is absent from the tree; its published algorithm (branch and
 *     bound over convex QP relaxations, terminating on the relative gap
 *     |best_bound - incumbent| / (1e-10 + |incumbent|) <= epgap) is restated here as a
 *     plain-C disjunctive branch and bound with a dense primal-dual interior-point QP.
 *
 * Parity pins (see tests/test_oracle_*.py): problem sizes 12361 rows / 29834 non-zeros /
 * 1240 binaries / 340 continuous and objective 9.57603 of cplexmodel_testcase.dat
 * (test/cplex_wrapper_test.cc:857-876), the full CPLEX solution vector of the same
 * instance (test/cplex_wrapper_test.cc:283-457), 8944 rows of test_sos.dat, plus
 * HiGHS (scipy.optimize.milp) outer-approximation brackets of the optimum generated
 * offline (tests/golden/, script oracle/make_highs_brackets.py).
 */
#ifndef MIQP_ORACLE_H
#define MIQP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Flat problem: content of ModelParameters (src/miqp_planner_data.hpp:99-185), row-major. */
typedef struct OrcProblem {
  int N, R, C, O, L, E;
  double ts;
  double min_vel, max_vel, total_min_acc, total_max_acc, total_min_jerk, total_max_jerk;
  double maximum_slack, w_slack, w_slack_obs, min_region_change_speed;
  double gap_tol, time_limit;
  const double *safety;        /* [N]  agent_safety_distance */
  const double *safety_slack;  /* [N]  agent_safety_distance_slack */
  const double *w_pos_x, *w_vel_x, *w_acc_x, *w_pos_y, *w_vel_y, *w_acc_y, *w_jerk_x, *w_jerk_y; /* [C] */
  const double *wheelbase, *radius;                 /* [C] */
  const double *x0;                                 /* [C][6] x,vx,ax,y,vy,ay */
  const double *x_ref, *vx_ref, *y_ref, *vy_ref;    /* [C][N] */
  const double *min_acc_x, *max_acc_x, *min_acc_y, *max_acc_y;     /* [C][R] */
  const double *min_jerk_x, *max_jerk_x, *min_jerk_y, *max_jerk_y; /* [C][R] */
  const int *initial_region;   /* [C], 1-based */
  const int *possible_region;  /* [C][R] */
  const double *obs_edges;     /* [O][N][L][4] x1,y1,x2,y2 */
  const int *obs_nedges;       /* [O][N] */
  const int *obs_soft;         /* [O] */
  const double *env_edges;     /* [env_off[E]][4] */
  const int *env_off;          /* [E+1] */
  const double *frac;          /* [R][4] */
  const double *poly_sint_ub, *poly_sint_lb, *poly_coss_ub, *poly_coss_lb; /* [R][3] */
  const double *poly_kappa_max, *poly_kappa_min;                           /* [R][3] */
} OrcProblem;

typedef struct OrcSizes {
  int ncols, ncont, nbin;
  long nrows, nnz_struct, nnz; /* nnz = structural minus exact-zero coefficients */
} OrcSizes;

/* column layout (decision_variables.mod order). Blocks 0..11 are the continuous core:
 * u_x,u_y,pos_x,vel_x,acc_x,pos_y,vel_y,acc_y,pos_x_front_UB,pos_x_front_LB,
 * pos_y_front_UB,pos_y_front_LB, each [C][N]. */
typedef struct OrcLayout {
  int C, N, R, O, L, E, K;
  int base_nwe, base_ar, base_rcna, base_dcc, base_dcf, base_so, base_sof, base_c2c, base_sv;
  int ncols;
} OrcLayout;

void orc_layout(const OrcProblem *p, OrcLayout *lay);
/* is_bin[ncols] (may be NULL), lb/ub[ncols] variable bounds as declared */
void orc_col_info(const OrcProblem *p, unsigned char *is_bin, double *lb, double *ub);

void orc_sizes(const OrcProblem *p, OrcSizes *out);

/* Structural CSR in OPL row order with explicit zeros kept.
 * rowptr[nrows+1], cols[nnz_struct], vals[nnz_struct], lo[nrows], hi[nrows].
 * Any pointer may be NULL (skipped). Returns nrows. */
long orc_build_rows(const OrcProblem *p, long *rowptr, int *cols, double *vals, double *lo, double *hi);

/* objective of a full column vector (objective_function.mod:7-19) */
double orc_objective(const OrcProblem *p, const double *x);
/* max violation over all rows, bounds and integrality; worst_row (may be NULL) gets the row
 * index of the worst row violation or -1-col for a bound/integrality violation */
double orc_max_violation(const OrcProblem *p, const double *x, long *worst_row);

/* Given only the 8 trajectory blocks (u_x,u_y,pos_x..acc_y filled in x), derive the front
 * axle variables, every binary and the slacks of a consistent full assignment (the
 * cheapest one).  Returns the resulting max violation. */
double orc_complete_assignment(const OrcProblem *p, double *x);

typedef struct OrcSolveInfo {
  int status;          /* 0 success (incumbent), 1 no solution, 3 time limit without incumbent */
  double objective, best_bound, gap, seconds, max_violation;
  long nodes, qp_solves, qp_iters;
  int proven;          /* 1 if gap <= gap_tol was reached */
  long uncertified;    /* nodes closed without a converged relaxation or a Farkas certificate (their bound stays in best_bound) */
} OrcSolveInfo;

/* Solve the MIQP.  x_out[ncols] receives the incumbent (full column vector).
 * warm (may be NULL) is a full column vector used as MIP start. */
int orc_solve(const OrcProblem *p, const double *warm, double *x_out, OrcSolveInfo *info, int verbose);

/* QP with every discrete decision taken from the binaries in x_bin (full vector; only
 * the binaries are read).  Used to pin the QP solver against the CPLEX golden vector. */
int orc_solve_fixed(const OrcProblem *p, const double *x_bin, double *x_out, double *objective);

#ifdef __cplusplus
}
#endif
#endif
