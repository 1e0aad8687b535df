/*
 * miqp_oracle.c -- model part of the CPU oracle: column layout, OPL-order row
 * instantiation, objective and violation evaluation.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see miqp_oracle.h).
 *
 * Every row family cites the .mod lines it restates (paths relative to
 * /root/reference/cplexmodel/).  Compile with -ffp-contract=off: the CUDA assembly kernel
 * is compared bit for bit against these coefficients.
 */
#include "miqp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* parameters.mod:24-32 */
#define BIGM_JERK 10.0
#define BIGM_FRAC 1000.0
#define BIGM_FRONT 100.0
#define BIGM_ACC 10.0
#define BIGM_KAPPA 1000.0
#define BIGM_VEL 100.0
#define BIGM_ENV 10000.0
#define BIGM_OBS 10000.0
#define BIGM_AGENTS 1000.0

enum { B_UX = 0, B_UY, B_PX, B_VX, B_AX, B_PY, B_VY, B_AY, B_XFU, B_XFL, B_YFU, B_YFL };

void orc_layout(const OrcProblem *p, OrcLayout *l) {
  l->C = p->C; l->N = p->N; l->R = p->R; l->O = p->O; l->L = p->L; l->E = p->E;
  l->K = p->C - 1;
  int C = p->C, N = p->N, R = p->R, O = p->O, L = p->L, E = p->E, K = l->K;
  int b = 12 * C * N;
  l->base_nwe = b;  b += 5 * C * E * N;
  l->base_ar = b;   b += C * N * R;
  l->base_rcna = b; b += 5 * C * N;
  l->base_dcc = b;  b += C * O * N * L;
  l->base_dcf = b;  b += 4 * C * O * N * L;
  l->base_so = b;   b += C * O * N;
  l->base_sof = b;  b += 4 * C * O * N;
  l->base_c2c = b;  b += 16 * K * K * N;
  l->base_sv = b;   b += 4 * K * K * N;
  l->ncols = b;
}

static inline int c_core(const OrcLayout *l, int blk, int c, int i) { return (blk * l->C + c) * l->N + i; }
static inline int c_nwe(const OrcLayout *l, int k, int c, int e, int i) { return l->base_nwe + ((k * l->C + c) * l->E + e) * l->N + i; }
static inline int c_ar(const OrcLayout *l, int c, int i, int j) { return l->base_ar + (c * l->N + i) * l->R + j; }
static inline int c_rcna(const OrcLayout *l, int k, int c, int i) { return l->base_rcna + (k * l->C + c) * l->N + i; }
static inline int c_dcc(const OrcLayout *l, int c, int o, int i, int e) { return l->base_dcc + ((c * l->O + o) * l->N + i) * l->L + e; }
static inline int c_dcf(const OrcLayout *l, int c, int o, int i, int e, int f) { return l->base_dcf + (((c * l->O + o) * l->N + i) * l->L + e) * 4 + f; }
static inline int c_so(const OrcLayout *l, int c, int o, int i) { return l->base_so + (c * l->O + o) * l->N + i; }
static inline int c_sof(const OrcLayout *l, int c, int o, int i, int f) { return l->base_sof + ((c * l->O + o) * l->N + i) * 4 + f; }
static inline int c_c2c(const OrcLayout *l, int k1, int k2, int i, int s) { return l->base_c2c + ((k1 * l->K + k2) * l->N + i) * 16 + s; }
static inline int c_sv(const OrcLayout *l, int k1, int k2, int i, int s) { return l->base_sv + ((k1 * l->K + k2) * l->N + i) * 4 + s; }

void orc_col_info(const OrcProblem *p, unsigned char *is_bin, double *lb, double *ub) {
  OrcLayout l; orc_layout(p, &l);
  for (int k = 0; k < l.ncols; ++k) {
    int bin = (k >= l.base_nwe && k < l.base_so) || (k >= l.base_c2c && k < l.base_sv);
    if (is_bin) is_bin[k] = (unsigned char)bin;
    double lo = -HUGE_VAL, hi = HUGE_VAL;
    if (bin) { lo = 0; hi = 1; }
    else if (k >= l.base_so && k < l.base_c2c) { lo = 0; hi = 1; }          /* decision_variables.mod:46-47 */
    else if (k >= l.base_sv) { lo = 0; hi = p->maximum_slack; }              /* decision_variables.mod:53 */
    if (lb) lb[k] = lo;
    if (ub) ub[k] = hi;
  }
}

/* ------------------------------------------------------------------------------------ */
typedef struct Sink {
  long nrows, nnzs, nnz;
  long *rowptr; int *cols; double *vals; double *lo, *hi;
  const double *x; double maxviol; long worst;
} Sink;

static void emit(Sink *s, double lo, double hi, int n, const int *cols, const double *vals) {
  if (s->rowptr) s->rowptr[s->nrows] = s->nnzs;
  if (s->lo) s->lo[s->nrows] = lo;
  if (s->hi) s->hi[s->nrows] = hi;
  double act = 0.0;
  for (int k = 0; k < n; ++k) {
    if (s->cols) s->cols[s->nnzs + k] = cols[k];
    if (s->vals) s->vals[s->nnzs + k] = vals[k];
    if (vals[k] != 0.0) s->nnz++;
    if (s->x) act += vals[k] * s->x[cols[k]];
  }
  if (s->x) {
    double v = 0.0;
    if (lo - act > v) v = lo - act;
    if (act - hi > v) v = act - hi;
    if (v > s->maxviol) { s->maxviol = v; s->worst = s->nrows; }
  }
  s->nnzs += n;
  s->nrows++;
}

static void emit1(Sink *s, double lo, double hi, int c0, double v0) { emit(s, lo, hi, 1, &c0, &v0); }
static void emit2(Sink *s, double lo, double hi, int c0, double v0, int c1, double v1) {
  int c[2] = {c0, c1}; double v[2] = {v0, v1}; emit(s, lo, hi, 2, c, v);
}

#define INF HUGE_VAL

/* cross-product row of one polygon edge for point (X,Y):
 * (x2-x1)*(Y-y1) - (X-x1)*(y2-y1)  =  dx*Y - dy*X - (dx*y1 - x1*dy)
 * obstacle_environment_constraints.mod:17-27 and :61-65 */
static void edge_terms(const double *e, double *dx, double *dy, double *rhs) {
  *dx = e[2] - e[0];
  *dy = e[3] - e[1];
  double a = *dx * e[1];
  double b = e[0] * *dy;
  *rhs = a - b;
}

static void build(const OrcProblem *p, Sink *s) {
  OrcLayout l; orc_layout(p, &l);
  const int C = p->C, N = p->N, R = p->R, O = p->O, L = p->L, E = p->E, K = l.K;
  const double ts = p->ts;
  const double c2 = 0.5 * (ts * ts);
  const double c3 = (1.0 / 6.0) * ((ts * ts) * ts);
  int cols[80]; double vals[80];
  int *bc = NULL; double *bv = NULL;
  int wide = R; if (L + 1 > wide) wide = L + 1; if (E > wide) wide = E; if (wide < 8) wide = 8;
  bc = (int *)malloc(sizeof(int) * (size_t)wide);
  bv = (double *)malloc(sizeof(double) * (size_t)wide);

  /* ---- initial_conditions.mod:13-27 ---- */
  for (int c = 0; c < C; ++c) {
    const double *x0 = p->x0 + 6 * c;
    double th = atan2(x0[4], x0[1]);           /* initialization.mod:25-29 */
    double ct = cos(th), st = sin(th), wb = p->wheelbase[c];
    emit1(s, x0[0], x0[0], c_core(&l, B_PX, c, 0), 1.0);
    emit1(s, x0[1], x0[1], c_core(&l, B_VX, c, 0), 1.0);
    emit1(s, x0[2], x0[2], c_core(&l, B_AX, c, 0), 1.0);
    emit1(s, x0[3], x0[3], c_core(&l, B_PY, c, 0), 1.0);
    emit1(s, x0[4], x0[4], c_core(&l, B_VY, c, 0), 1.0);
    emit1(s, x0[5], x0[5], c_core(&l, B_AY, c, 0), 1.0);
    double fx = x0[0] + ct * wb, fy = x0[3] + st * wb;
    emit1(s, fx, fx, c_core(&l, B_XFU, c, 0), 1.0);
    emit1(s, fx, fx, c_core(&l, B_XFL, c, 0), 1.0);
    emit1(s, fy, fy, c_core(&l, B_YFU, c, 0), 1.0);
    emit1(s, fy, fy, c_core(&l, B_YFL, c, 0), 1.0);
    emit1(s, 0.0, 0.0, c_core(&l, B_UX, c, N - 1), 1.0);
    emit1(s, 0.0, 0.0, c_core(&l, B_UY, c, N - 1), 1.0);
  }
  /* ---- initial_conditions.mod:29-50 (j outer, c inner) ---- */
  for (int j = 0; j < R; ++j)
    for (int c = 0; c < C; ++c) {
      int ar = c_ar(&l, c, 0, j);
      double on = (j + 1 == p->initial_region[c]) ? 1.0 : 0.0;
      emit1(s, on, on, ar, 1.0);
      int ux = c_core(&l, B_UX, c, 0), uy = c_core(&l, B_UY, c, 0);
      emit2(s, -INF, p->max_jerk_x[c * R + j] + BIGM_JERK, ux, 1.0, ar, BIGM_JERK);
      emit2(s, p->min_jerk_x[c * R + j] - BIGM_JERK, INF, ux, 1.0, ar, -BIGM_JERK);
      emit2(s, -INF, p->max_jerk_y[c * R + j] + BIGM_JERK, uy, 1.0, ar, BIGM_JERK);
      emit2(s, p->min_jerk_y[c * R + j] - BIGM_JERK, INF, uy, 1.0, ar, -BIGM_JERK);
    }
  /* ---- initial_conditions.mod:52-60 ---- */
  for (int c = 0; c < C; ++c)
    for (int k = 0; k < 5; ++k) emit1(s, 0.0, 0.0, c_rcna(&l, k, c, 0), 1.0);

  /* ---- model_region_constraints.mod:11-19 ---- */
  for (int i = 1; i < N; ++i)
    for (int c = 0; c < C; ++c)
      for (int ax = 0; ax < 2; ++ax) {
        int P = ax ? B_PY : B_PX, V = ax ? B_VY : B_VX, A = ax ? B_AY : B_AX, U = ax ? B_UY : B_UX;
        cols[0] = c_core(&l, P, c, i); vals[0] = 1.0;
        cols[1] = c_core(&l, P, c, i - 1); vals[1] = -1.0;
        cols[2] = c_core(&l, V, c, i - 1); vals[2] = -ts;
        cols[3] = c_core(&l, A, c, i - 1); vals[3] = -c2;
        cols[4] = c_core(&l, U, c, i - 1); vals[4] = -c3;
        emit(s, 0.0, 0.0, 5, cols, vals);
        cols[0] = c_core(&l, V, c, i); vals[0] = 1.0;
        cols[1] = c_core(&l, V, c, i - 1); vals[1] = -1.0;
        cols[2] = c_core(&l, A, c, i - 1); vals[2] = -ts;
        cols[3] = c_core(&l, U, c, i - 1); vals[3] = -c2;
        emit(s, 0.0, 0.0, 4, cols, vals);
        cols[0] = c_core(&l, A, c, i); vals[0] = 1.0;
        cols[1] = c_core(&l, A, c, i - 1); vals[1] = -1.0;
        cols[2] = c_core(&l, U, c, i - 1); vals[2] = -ts;
        emit(s, 0.0, 0.0, 3, cols, vals);
      }
  /* ---- model_region_constraints.mod:22-39 (vel_y has no upper bound; vel_x twice) ---- */
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < C; ++c) {
      emit1(s, p->min_vel, INF, c_core(&l, B_VX, c, i), 1.0);
      emit1(s, p->min_vel, INF, c_core(&l, B_VY, c, i), 1.0);
      emit1(s, -INF, p->max_vel, c_core(&l, B_VX, c, i), 1.0);
      emit1(s, -INF, p->max_vel, c_core(&l, B_VX, c, i), 1.0);
      emit1(s, -INF, p->total_max_acc, c_core(&l, B_AX, c, i), 1.0);
      emit1(s, p->total_min_acc, INF, c_core(&l, B_AX, c, i), 1.0);
      emit1(s, -INF, p->total_max_acc, c_core(&l, B_AY, c, i), 1.0);
      emit1(s, p->total_min_acc, INF, c_core(&l, B_AY, c, i), 1.0);
      emit1(s, -INF, p->total_max_jerk, c_core(&l, B_UX, c, i), 1.0);
      emit1(s, p->total_min_jerk, INF, c_core(&l, B_UX, c, i), 1.0);
      emit1(s, -INF, p->total_max_jerk, c_core(&l, B_UY, c, i), 1.0);
      emit1(s, p->total_min_jerk, INF, c_core(&l, B_UY, c, i), 1.0);
    }
  /* ---- model_region_constraints.mod:43-113 ---- */
  for (int i = 1; i < N; ++i)
    for (int c = 0; c < C; ++c) {
      const double wb = p->wheelbase[c];
      const int px = c_core(&l, B_PX, c, i), vx = c_core(&l, B_VX, c, i), axc = c_core(&l, B_AX, c, i);
      const int py = c_core(&l, B_PY, c, i), vy = c_core(&l, B_VY, c, i), ayc = c_core(&l, B_AY, c, i);
      const int ux = c_core(&l, B_UX, c, i), uy = c_core(&l, B_UY, c, i);
      const int rho = c_rcna(&l, 4, c, i);
      for (int j = 0; j < R; ++j) {
        const int ar = c_ar(&l, c, i, j);
        if (p->possible_region[c * R + j] == 1) {
          const double *f = p->frac + 4 * j;
          /* :53-54 wedge */
          cols[0] = vy; vals[0] = f[0]; cols[1] = vx; vals[1] = -f[1]; cols[2] = ar; vals[2] = -BIGM_FRAC; cols[3] = rho; vals[3] = BIGM_FRAC;
          emit(s, -BIGM_FRAC, INF, 4, cols, vals);
          cols[0] = vy; vals[0] = f[2]; cols[1] = vx; vals[1] = -f[3]; cols[2] = ar; vals[2] = BIGM_FRAC; cols[3] = rho; vals[3] = -BIGM_FRAC;
          emit(s, -INF, BIGM_FRAC, 4, cols, vals);
          /* :57-70 front axle box */
          const double *polys[4] = {p->poly_coss_ub + 3 * j, p->poly_coss_lb + 3 * j, p->poly_sint_ub + 3 * j, p->poly_sint_lb + 3 * j};
          const int fcol[4] = {c_core(&l, B_XFU, c, i), c_core(&l, B_XFL, c, i), c_core(&l, B_YFU, c, i), c_core(&l, B_YFL, c, i)};
          for (int q = 0; q < 4; ++q) {
            const double *P = polys[q];
            cols[0] = fcol[q]; vals[0] = 1.0;
            cols[1] = (q < 2) ? px : py; vals[1] = -1.0;
            cols[2] = vx; vals[2] = -(wb * P[1]);
            cols[3] = vy; vals[3] = -(wb * P[2]);
            cols[4] = ar; vals[4] = -BIGM_FRONT;
            emit(s, wb * P[0] - BIGM_FRONT, INF, 5, cols, vals);
            vals[4] = BIGM_FRONT;
            emit(s, -INF, wb * P[0] + BIGM_FRONT, 5, cols, vals);
          }
          /* :73-82 jerk */
          emit2(s, -INF, p->max_jerk_x[c * R + j] + BIGM_JERK, ux, 1.0, ar, BIGM_JERK);
          emit2(s, p->min_jerk_x[c * R + j] - BIGM_JERK, INF, ux, 1.0, ar, -BIGM_JERK);
          emit2(s, -INF, p->max_jerk_y[c * R + j] + BIGM_JERK, uy, 1.0, ar, BIGM_JERK);
          emit2(s, p->min_jerk_y[c * R + j] - BIGM_JERK, INF, uy, 1.0, ar, -BIGM_JERK);
          /* :85-94 acc */
          emit2(s, -INF, p->max_acc_x[c * R + j] + BIGM_ACC, axc, 1.0, ar, BIGM_ACC);
          emit2(s, p->min_acc_x[c * R + j] - BIGM_ACC, INF, axc, 1.0, ar, -BIGM_ACC);
          emit2(s, -INF, p->max_acc_y[c * R + j] + BIGM_ACC, ayc, 1.0, ar, BIGM_ACC);
          emit2(s, p->min_acc_y[c * R + j] - BIGM_ACC, INF, ayc, 1.0, ar, -BIGM_ACC);
          /* :97-104 curvature */
          const double sl = (f[1] + f[3]) / (f[0] + f[2]);
          const double *KX = p->poly_kappa_max + 3 * j, *KN = p->poly_kappa_min + 3 * j;
          cols[0] = ayc; vals[0] = 1.0; cols[1] = vx; vals[1] = -KX[1]; cols[2] = vy; vals[2] = -KX[2];
          cols[3] = axc; vals[3] = -sl; cols[4] = ar; vals[4] = BIGM_KAPPA; cols[5] = rho; vals[5] = -BIGM_KAPPA;
          emit(s, -INF, KX[0] + BIGM_KAPPA, 6, cols, vals);
          cols[0] = ayc; vals[0] = 1.0; cols[1] = vx; vals[1] = -KN[1]; cols[2] = vy; vals[2] = -KN[2];
          cols[3] = axc; vals[3] = -sl; cols[4] = ar; vals[4] = -BIGM_KAPPA; cols[5] = rho; vals[5] = BIGM_KAPPA;
          emit(s, KN[0] - BIGM_KAPPA, INF, 6, cols, vals);
        } else {
          emit1(s, 0.0, 0.0, ar, 1.0); /* :108 */
        }
      }
      for (int j = 0; j < R; ++j) { bc[j] = c_ar(&l, c, i, j); bv[j] = 1.0; }
      emit(s, 1.0, 1.0, R, bc, bv); /* :113 */
    }
  /* ---- minimum_speed_constraints.mod:9-44 (repeated for every j) ---- */
  {
    const double vm = p->min_region_change_speed;
    for (int i = 1; i < N; ++i)
      for (int c = 0; c < C; ++c) {
        const int vx = c_core(&l, B_VX, c, i), vy = c_core(&l, B_VY, c, i);
        const int bxp = c_rcna(&l, 0, c, i), byp = c_rcna(&l, 1, c, i), bxn = c_rcna(&l, 2, c, i), byn = c_rcna(&l, 3, c, i);
        const int rho = c_rcna(&l, 4, c, i);
        for (int j = 0; j < R; ++j) {
          emit2(s, vm, INF, vx, 1.0, bxp, BIGM_VEL);
          emit2(s, -INF, vm + BIGM_VEL, vx, 1.0, bxp, BIGM_VEL);
          emit2(s, -INF, vm + BIGM_VEL, vx, -1.0, bxn, BIGM_VEL);
          emit2(s, vm, INF, vx, -1.0, bxn, BIGM_VEL);
          emit2(s, vm, INF, vy, 1.0, byp, BIGM_VEL);
          emit2(s, -INF, vm + BIGM_VEL, vy, 1.0, byp, BIGM_VEL);
          emit2(s, -INF, vm + BIGM_VEL, vy, -1.0, byn, BIGM_VEL);
          emit2(s, vm, INF, vy, -1.0, byn, BIGM_VEL);
          cols[0] = c_ar(&l, c, i, j); vals[0] = 1.0; cols[1] = c_ar(&l, c, i - 1, j); vals[1] = -1.0; cols[2] = rho; vals[2] = 1.0;
          emit(s, -INF, 1.0, 3, cols, vals);
          vals[2] = -1.0;
          emit(s, -1.0, INF, 3, cols, vals);
          emit2(s, -INF, 0.0, rho, 1.0, bxp, -1.0);
          emit2(s, -INF, 0.0, rho, 1.0, byp, -1.0);
          emit2(s, -INF, 0.0, rho, 1.0, bxn, -1.0);
          emit2(s, -INF, 0.0, rho, 1.0, byn, -1.0);
          cols[0] = rho; vals[0] = 1.0; cols[1] = bxp; vals[1] = -1.0; cols[2] = byp; vals[2] = -1.0; cols[3] = bxn; vals[3] = -1.0; cols[4] = byn; vals[4] = -1.0;
          emit(s, -3.0, INF, 5, cols, vals);
        }
      }
  }
  /* ---- obstacle_environment_constraints.mod:9-36 ---- */
  if (E > 0) {
    for (int i = 0; i < N; ++i)
      for (int c = 0; c < C; ++c) {
        const int px = c_core(&l, B_PX, c, i), py = c_core(&l, B_PY, c, i);
        const int xu = c_core(&l, B_XFU, c, i), xl = c_core(&l, B_XFL, c, i), yu = c_core(&l, B_YFU, c, i), yl = c_core(&l, B_YFL, c, i);
        const int X[5] = {px, xu, xl, xu, xl};
        const int Y[5] = {py, yu, yu, yl, yl};
        for (int e = 0; e < E; ++e)
          for (int ed = p->env_off[e]; ed < p->env_off[e + 1]; ++ed) {
            double dx, dy, rhs; edge_terms(p->env_edges + 4 * ed, &dx, &dy, &rhs);
            for (int k = 0; k < 5; ++k) {
              cols[0] = Y[k]; vals[0] = dx; cols[1] = X[k]; vals[1] = -dy; cols[2] = c_nwe(&l, k, c, e, i); vals[2] = BIGM_ENV;
              emit(s, rhs, INF, 3, cols, vals);
            }
          }
        for (int k = 0; k < 5; ++k) {
          for (int e = 0; e < E; ++e) { bc[e] = c_nwe(&l, k, c, e, i); bv[e] = 1.0; }
          emit(s, -INF, (double)(E - 1), E, bc, bv);
        }
      }
  }
  /* ---- obstacle_environment_constraints.mod:52-97 ---- */
  if (O > 0) {
    for (int i = 0; i < N; ++i)
      for (int c = 0; c < C; ++c) {
        const int px = c_core(&l, B_PX, c, i), py = c_core(&l, B_PY, c, i);
        const int xu = c_core(&l, B_XFU, c, i), xl = c_core(&l, B_XFL, c, i), yu = c_core(&l, B_YFU, c, i), yl = c_core(&l, B_YFL, c, i);
        /* rear, front1=(LB,LB), front2=(UB,LB), front3=(LB,UB), front4=(UB,UB) */
        const int X[5] = {px, xl, xu, xl, xu};
        const int Y[5] = {py, yl, yl, yu, yu};
        for (int o = 0; o < O; ++o) {
          const int ne = p->obs_nedges[o * N + i];
          for (int ed = 0; ed < ne; ++ed) {
            double dx, dy, rhs; edge_terms(p->obs_edges + 4 * ((o * N + i) * L + ed), &dx, &dy, &rhs);
            for (int k = 0; k < 5; ++k) {
              int dcol = (k == 0) ? c_dcc(&l, c, o, i, ed) : c_dcf(&l, c, o, i, ed, k - 1);
              cols[0] = Y[k]; vals[0] = dx; cols[1] = X[k]; vals[1] = -dy; cols[2] = dcol; vals[2] = -BIGM_OBS;
              emit(s, -INF, rhs, 3, cols, vals);
            }
          }
          for (int k = 0; k < 5; ++k) {
            int n = 0;
            for (int ed = 0; ed < ne; ++ed) { bc[n] = (k == 0) ? c_dcc(&l, c, o, i, ed) : c_dcf(&l, c, o, i, ed, k - 1); bv[n] = 1.0; ++n; }
            if (p->obs_soft[o] == 1) { bc[n] = (k == 0) ? c_so(&l, c, o, i) : c_sof(&l, c, o, i, k - 1); bv[n] = -1.0; ++n; }
            emit(s, -INF, (double)(ne - 1), n, bc, bv);
          }
        }
      }
  }
  /* ---- agent_collision_constraints.mod:10-73 ---- */
  if (C > 1) {
    for (int i = 0; i < N; ++i)
      for (int k1 = 1; k1 < K; ++k1)
        for (int k2 = 0; k2 < k1; ++k2) {
          for (int q = 0; q < 4; ++q) emit1(s, 0.0, 0.0, c_sv(&l, k1, k2, i, q), 1.0);
          for (int q = 0; q < 16; ++q) emit1(s, 0.0, 0.0, c_c2c(&l, k1, k2, i, q), 1.0);
        }
    for (int i = 0; i < N; ++i)
      for (int a = 0; a < C - 1; ++a)
        for (int b = a + 1; b < C; ++b) {
          const int k1 = a, k2 = b - 1;
          const double RR = p->radius[a] + p->radius[b];           /* initialization.mod:16-22 */
          const double D = RR + p->safety[i];
          const double Ds = RR + p->safety[i] + p->safety_slack[i];
          const int pxa = c_core(&l, B_PX, a, i), pya = c_core(&l, B_PY, a, i), pxb = c_core(&l, B_PX, b, i), pyb = c_core(&l, B_PY, b, i);
          const int xua = c_core(&l, B_XFU, a, i), xla = c_core(&l, B_XFL, a, i), yua = c_core(&l, B_YFU, a, i), yla = c_core(&l, B_YFL, a, i);
          const int xub = c_core(&l, B_XFU, b, i), xlb = c_core(&l, B_XFL, b, i), yub = c_core(&l, B_YFU, b, i), ylb = c_core(&l, B_YFL, b, i);
#define BB(q) c_c2c(&l, k1, k2, i, (q) - 1)
#define SS(q) c_sv(&l, k1, k2, i, (q) - 1)
#define ROW4(lo_, hi_, ca, va, cb, vb, cc, vc, cd, vd) do { cols[0]=ca; vals[0]=va; cols[1]=cb; vals[1]=vb; cols[2]=cc; vals[2]=vc; cols[3]=cd; vals[3]=vd; emit(s, lo_, hi_, 4, cols, vals);} while (0)
#define ROW3(lo_, hi_, ca, va, cb, vb, cc, vc) do { cols[0]=ca; vals[0]=va; cols[1]=cb; vals[1]=vb; cols[2]=cc; vals[2]=vc; emit(s, lo_, hi_, 3, cols, vals);} while (0)
          /* :41-47 rear/rear:  pa <= pb - (Ds - s) + M b   <=>  pa - pb - s - M b <= -Ds */
          ROW4(-INF, -Ds, pxa, 1.0, pxb, -1.0, SS(1), -1.0, BB(1), -BIGM_AGENTS);
          ROW4(Ds, INF, pxa, 1.0, pxb, -1.0, SS(1), 1.0, BB(2), BIGM_AGENTS);
          ROW4(-INF, -Ds, pya, 1.0, pyb, -1.0, SS(2), -1.0, BB(3), -BIGM_AGENTS);
          ROW4(Ds, INF, pya, 1.0, pyb, -1.0, SS(2), 1.0, BB(4), BIGM_AGENTS);
          ROW4(-INF, 3.0, BB(1), 1.0, BB(2), 1.0, BB(3), 1.0, BB(4), 1.0);
          emit1(s, -INF, p->safety_slack[i], SS(1), 1.0);
          emit1(s, -INF, p->safety_slack[i], SS(2), 1.0);
          /* :50-54 rear a vs front b */
          ROW3(-INF, -D, pxa, 1.0, xlb, -1.0, BB(5), -BIGM_AGENTS);
          ROW3(D, INF, pxa, 1.0, xub, -1.0, BB(6), BIGM_AGENTS);
          ROW3(-INF, -D, pya, 1.0, ylb, -1.0, BB(7), -BIGM_AGENTS);
          ROW3(D, INF, pya, 1.0, yub, -1.0, BB(8), BIGM_AGENTS);
          ROW4(-INF, 3.0, BB(5), 1.0, BB(6), 1.0, BB(7), 1.0, BB(8), 1.0);
          /* :57-61 rear b vs front a */
          ROW3(-INF, -D, pxb, 1.0, xla, -1.0, BB(9), -BIGM_AGENTS);
          ROW3(D, INF, pxb, 1.0, xua, -1.0, BB(10), BIGM_AGENTS);
          ROW3(-INF, -D, pyb, 1.0, yla, -1.0, BB(11), -BIGM_AGENTS);
          ROW3(D, INF, pyb, 1.0, yua, -1.0, BB(12), BIGM_AGENTS);
          ROW4(-INF, 3.0, BB(9), 1.0, BB(10), 1.0, BB(11), 1.0, BB(12), 1.0);
          /* :65-71 front/front:  0 <= xla - (Ds - s) - xub + M b  <=>  xla - xub + s + M b >= Ds */
          ROW4(Ds, INF, xla, 1.0, xub, -1.0, SS(3), 1.0, BB(13), BIGM_AGENTS);
          ROW4(-INF, -Ds, xua, 1.0, xlb, -1.0, SS(3), -1.0, BB(14), -BIGM_AGENTS);
          ROW4(Ds, INF, yla, 1.0, yub, -1.0, SS(4), 1.0, BB(15), BIGM_AGENTS);
          ROW4(-INF, -Ds, yua, 1.0, ylb, -1.0, SS(4), -1.0, BB(16), -BIGM_AGENTS);
          ROW4(-INF, 3.0, BB(13), 1.0, BB(14), 1.0, BB(15), 1.0, BB(16), 1.0);
          emit1(s, -INF, p->safety_slack[i], SS(3), 1.0);
          emit1(s, -INF, p->safety_slack[i], SS(4), 1.0);
#undef BB
#undef SS
#undef ROW4
#undef ROW3
        }
  }
  if (s->rowptr) s->rowptr[s->nrows] = s->nnzs;
  free(bc); free(bv);
}

void orc_sizes(const OrcProblem *p, OrcSizes *out) {
  OrcLayout l; orc_layout(p, &l);
  Sink s; memset(&s, 0, sizeof s);
  build(p, &s);
  out->ncols = l.ncols;
  out->nbin = (l.base_so - l.base_nwe) + (l.base_sv - l.base_c2c);
  out->ncont = l.ncols - out->nbin;
  out->nrows = s.nrows; out->nnz_struct = s.nnzs; out->nnz = s.nnz;
}

long orc_build_rows(const OrcProblem *p, long *rowptr, int *cols, double *vals, double *lo, double *hi) {
  Sink s; memset(&s, 0, sizeof s);
  s.rowptr = rowptr; s.cols = cols; s.vals = vals; s.lo = lo; s.hi = hi;
  build(p, &s);
  return s.nrows;
}

double orc_objective(const OrcProblem *p, const double *x) {
  OrcLayout l; orc_layout(p, &l);
  const int C = p->C, N = p->N, O = p->O, K = l.K;
  double cost = 0.0;
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < C; ++c) {
      double dpx = x[c_core(&l, B_PX, c, i)] - p->x_ref[c * N + i];
      double dvx = x[c_core(&l, B_VX, c, i)] - p->vx_ref[c * N + i];
      double dpy = x[c_core(&l, B_PY, c, i)] - p->y_ref[c * N + i];
      double dvy = x[c_core(&l, B_VY, c, i)] - p->vy_ref[c * N + i];
      double ax = x[c_core(&l, B_AX, c, i)], ay = x[c_core(&l, B_AY, c, i)];
      double ux = x[c_core(&l, B_UX, c, i)], uy = x[c_core(&l, B_UY, c, i)];
      cost += p->w_pos_x[c] * dpx * dpx + p->w_vel_x[c] * dvx * dvx + p->w_acc_x[c] * ax * ax
            + p->w_pos_y[c] * dpy * dpy + p->w_vel_y[c] * dvy * dvy + p->w_acc_y[c] * ay * ay
            + p->w_jerk_x[c] * ux * ux + p->w_jerk_y[c] * uy * uy;
    }
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < C; ++c)
      for (int o = 0; o < O; ++o) {
        double v = x[c_so(&l, c, o, i)];
        cost += p->w_slack_obs * v * v;
        for (int f = 0; f < 4; ++f) { double w = x[c_sof(&l, c, o, i, f)]; cost += p->w_slack_obs * w * w; }
      }
  for (int i = 0; i < N; ++i)
    for (int k1 = 0; k1 < K; ++k1)
      for (int k2 = 0; k2 < K; ++k2)
        for (int q = 0; q < 4; ++q) { double v = x[c_sv(&l, k1, k2, i, q)]; cost += p->w_slack * v * v; }
  return cost;
}

double orc_max_violation(const OrcProblem *p, const double *x, long *worst_row) {
  OrcLayout l; orc_layout(p, &l);
  Sink s; memset(&s, 0, sizeof s);
  s.x = x; s.worst = -1;
  build(p, &s);
  double mv = s.maxviol; long worst = s.worst;
  unsigned char *isb = (unsigned char *)malloc((size_t)l.ncols);
  double *lb = (double *)malloc(sizeof(double) * (size_t)l.ncols), *ub = (double *)malloc(sizeof(double) * (size_t)l.ncols);
  orc_col_info(p, isb, lb, ub);
  for (int k = 0; k < l.ncols; ++k) {
    double v = 0.0;
    if (lb[k] - x[k] > v) v = lb[k] - x[k];
    if (x[k] - ub[k] > v) v = x[k] - ub[k];
    if (isb[k]) { double r = fabs(x[k] - floor(x[k] + 0.5)); if (r > v) v = r; }
    if (!(x[k] == x[k])) v = HUGE_VAL;
    if (v > mv) { mv = v; worst = -1 - k; }
  }
  free(isb); free(lb); free(ub);
  if (worst_row) *worst_row = worst;
  return mv;
}
