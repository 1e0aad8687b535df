#define _POSIX_C_SOURCE 200809L
/*
 * miqp_oracle_bnb.c -- solver part of the CPU oracle.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see miqp_oracle.h).
 *
 * Restates what the reference delegates to CPLEX (src/cplex_wrapper.cpp:158-185): a
 * branch and bound over the binaries of the cplexmodel .mod files that terminates on the CPLEX
 * relative gap |best_bound - incumbent| / (1e-10 + |incumbent|) <= epgap
 * (cplexmodel.mod:8-10) or on the time limit.
 *
 * Algorithm (deliberately different from the CUDA path so that agreement means
 * something): the states are eliminated through the triple integrator
 * (model_region_constraints.mod:11-19), leaving a dense QP in the jerks; a node fixes a
 * subset of the model's disjunctions (region + low-speed mode per car-step, environment
 * polygon per point, separating obstacle edge per point, collision side per pair
 * quadruple); undecided disjunctions contribute no rows (a relaxation of the big-M LP
 * relaxation); branching picks the most violated disjunction of the relaxed optimum.
 * Node QPs are solved by a dense Mehrotra primal-dual interior-point method with a
 * Cholesky factorisation.  A node whose relaxed optimum satisfies one alternative of
 * every disjunction is re-solved with those alternatives enforced and becomes the
 * incumbent.
 */
#include "miqp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define BIGM_OBS_PENALTY_SLACK 1.0 /* slackvarsObstacle = 1 relaxes a soft obstacle completely */

enum { Y_PX = 0, Y_VX, Y_AX, Y_PY, Y_VY, Y_AY, Y_UX, Y_UY };

/* decision codes */
#define UNDEC 255
#define MODE_FROZEN 254
#define OBS_SOFT 253
/* mode value for rho=0: j*4 + h, h in {0:vx>=vm, 1:vy>=vm, 2:vx<=-vm, 3:vy<=-vm} */

typedef struct SRow {
  int c, i, c2, slack; /* c2 = -1 if single car, slack = -1 if none (index into slack vars) */
  double a[8], a2[8], as, rhs;
} SRow;

typedef struct Ctx {
  const OrcProblem *p;
  OrcLayout lay;
  int C, N, R, O, L, E, P; /* P pairs */
  int nu, ns, n;           /* controls, slack vars, total */
  int ndec;                /* bytes per node */
  int off_mode, off_env, off_obs, off_pair;
  double *Sp, *Sv, *Sa;    /* [N][N-1] influence of u_j on state_i */
  double *y0;              /* [C][N][6] free response */
  double *Q, *cvec; double cconst;
  int *nalt_mode; unsigned char *alt_mode; /* per car: list of (j*4+h) alternatives */
  double feas_tol;
  /* statistics */
  long qp_solves, qp_iters;
} Ctx;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static inline int zidx(const Ctx *k, int c, int axis, int j) { return (c * 2 + axis) * (k->N - 1) + j; }
static inline int pair_index(const Ctx *k, int a, int b) { /* a<b */ int idx = 0; for (int x = 0; x < a; ++x) idx += k->C - 1 - x; return idx + (b - a - 1); }
static inline int sidx(const Ctx *k, int pr, int i, int q) { return k->nu + (pr * k->N + i) * 4 + q; }
static inline unsigned char *d_mode(const Ctx *k, unsigned char *d, int c, int i) { return d + k->off_mode + c * k->N + i; }
static inline unsigned char *d_env(const Ctx *k, unsigned char *d, int c, int i, int pt) { return d + k->off_env + (c * k->N + i) * 5 + pt; }
static inline unsigned char *d_obs(const Ctx *k, unsigned char *d, int c, int o, int i, int pt) { return d + k->off_obs + ((c * k->O + o) * k->N + i) * 5 + pt; }
static inline unsigned char *d_pair(const Ctx *k, unsigned char *d, int pr, int i, int q) { return d + k->off_pair + (pr * k->N + i) * 4 + q; }

/* ---------------------------------------------------------------------------------- */
/* dense linear algebra */
static int chol(double *A, int n) { /* lower, in place; returns 0 ok */
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return 1;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return 0;
}
static void chol_solve(const double *Lm, int n, double *b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= Lm[i * n + k] * b[k]; b[i] = s / Lm[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= Lm[k * n + i] * b[k]; b[i] = s / Lm[i * n + i]; }
}

/* ---------------------------------------------------------------------------------- */
typedef struct QP {
  int n, m, cap;
  double *G, *h;      /* m x n row-major */
  int *seg;           /* per row: nseg, then (start,len) x 5  -> 11 ints */
} QP;

static void qp_init(QP *q, int n) { q->n = n; q->m = 0; q->cap = 0; q->G = NULL; q->h = NULL; q->seg = NULL; }
static void qp_free(QP *q) { free(q->G); free(q->h); free(q->seg); }
static double *qp_newrow(QP *q) {
  if (q->m == q->cap) {
    q->cap = q->cap ? q->cap * 2 : 256;
    q->G = (double *)realloc(q->G, sizeof(double) * (size_t)q->cap * q->n);
    q->h = (double *)realloc(q->h, sizeof(double) * (size_t)q->cap);
    q->seg = (int *)realloc(q->seg, sizeof(int) * (size_t)q->cap * 11);
  }
  double *g = q->G + (size_t)q->m * q->n;
  memset(g, 0, sizeof(double) * (size_t)q->n);
  q->seg[q->m * 11] = 0;
  return g;
}
static void qp_addseg(QP *q, int start, int len) {
  if (len <= 0) return;
  int *s = q->seg + q->m * 11;
  /* keep segments sorted by start */
  int ns = s[0], pos = ns;
  while (pos > 0 && s[1 + 2 * (pos - 1)] > start) { s[1 + 2 * pos] = s[1 + 2 * (pos - 1)]; s[2 + 2 * pos] = s[2 + 2 * (pos - 1)]; --pos; }
  s[1 + 2 * pos] = start; s[2 + 2 * pos] = len; s[0] = ns + 1;
}

/* condensed row from a stage row; returns 0 if the row became constant (then *viol holds lhs-rhs) */
static int condense(const Ctx *k, QP *q, const SRow *r, double *constviol) {
  const int N = k->N, Nm = N - 1;
  double *g = qp_newrow(q);
  double h = r->rhs;
  for (int side = 0; side < 2; ++side) {
    int c = side ? r->c2 : r->c;
    if (c < 0) continue;
    const double *a = side ? r->a2 : r->a;
    int i = r->i;
    const double *y0 = k->y0 + (c * N + i) * 6;
    for (int t = 0; t < 6; ++t) h -= a[t] * y0[t];
    for (int axis = 0; axis < 2; ++axis) {
      double ap = a[axis * 3 + 0], av = a[axis * 3 + 1], aa = a[axis * 3 + 2], au = a[6 + axis];
      int len = 0;
      if (ap != 0.0 || av != 0.0 || aa != 0.0) {
        for (int j = 0; j < i; ++j)
          g[zidx(k, c, axis, j)] += ap * k->Sp[i * Nm + j] + av * k->Sv[i * Nm + j] + aa * k->Sa[i * Nm + j];
        len = i;
      }
      if (au != 0.0 && i < Nm) { g[zidx(k, c, axis, i)] += au; len = i + 1; }
      if (len > Nm) len = Nm;
      qp_addseg(q, zidx(k, c, axis, 0), len);
    }
  }
  if (r->slack >= 0 && r->as != 0.0) { g[r->slack] += r->as; qp_addseg(q, r->slack, 1); }
  /* constant row? */
  const int *s = q->seg + q->m * 11;
  double nrm = 0.0;
  for (int a = 0; a < s[0]; ++a) for (int j = 0; j < s[2 + 2 * a]; ++j) { double v = g[s[1 + 2 * a] + j]; nrm += v * v; }
  if (nrm < 1e-24) { if (constviol) *constviol = -h; return 0; }
  q->h[q->m] = h;
  q->m++;
  return 1;
}

/* Farkas certificate of infeasibility of {G z <= h, |z_a| <= B_a}: multipliers lam >= 0 with
 *   h'lam + sum_a |(G'lam)_a| B_a < 0.
 * (Every feasible z has lam'(Gz - h) <= 0, i.e. (G'lam)'z <= h'lam, while (G'lam)'z >= -sum |(G'lam)_a| B_a.)
 * B_a: the jerks are bounded by the global jerk box (model_region_constraints.mod:36-39, rows of every node),
 * the pair slacks by maximum_slack (decision_variables.mod:53).  The multipliers of a diverging
 * interior-point iteration converge to such a certificate when the node is infeasible. */
static int farkas_certificate(const Ctx *k, const QP *q, const double *lam, double *r /* [n] scratch */) {
  const int n = q->n, m = q->m;
  double lmax = 0.0;
  for (int i = 0; i < m; ++i) if (lam[i] > lmax) lmax = lam[i];
  if (!(lmax > 0.0) || !(lmax < HUGE_VAL)) return 0;
  const double sc = 1.0 / lmax;
  memset(r, 0, sizeof(double) * (size_t)n);
  double hl = 0.0, mag = 0.0;
  for (int i = 0; i < m; ++i) {
    const double l = lam[i] * sc;
    if (!(l > 0.0)) continue;
    const double *g = q->G + (size_t)i * n; const int *sg = q->seg + i * 11;
    for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) r[st + j] += g[st + j] * l; }
    hl += q->h[i] * l; mag += fabs(q->h[i]) * l;
  }
  const OrcProblem *p = k->p;
  const double U = fmax(fabs(p->total_min_jerk), fabs(p->total_max_jerk));
  double slackterm = 0.0;
  for (int a = 0; a < n; ++a) slackterm += fabs(r[a]) * (a < k->nu ? U : p->maximum_slack);
  return hl + slackterm < -1e-10 * (mag + slackterm) - 1e-13;
}

/* Dense Mehrotra predictor-corrector.  min 1/2 z'Qz + c'z  s.t. Gz <= h.
 * returns QP_OPTIMAL, QP_INFEASIBLE (Farkas certificate verified), QP_FEASIBLE_POINT (the iteration
 * stalled at a primal feasible point: *obj_out is only an upper bound of the optimum) or QP_UNKNOWN
 * (no convergence, no certificate). */
enum { QP_OPTIMAL = 0, QP_INFEASIBLE = 1, QP_FEASIBLE_POINT = 2, QP_UNKNOWN = 3 };
static int qp_solve(Ctx *k, const QP *q, double *z, double *obj_out, double *lb_out) {
  const int n = q->n, m = q->m;
  const double *Q = k->Q, *cv = k->cvec;
  double *s = (double *)malloc(sizeof(double) * (size_t)(m + 1) * 8);
  double *lam = s + m, *ds = lam + m, *dl = ds + m, *rp = dl + m, *w = rp + m, *t = w + m, *dsa = t + m;
  double *M = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *rd = (double *)malloc(sizeof(double) * (size_t)n * 4);
  double *dz = rd + n, *rhs = dz + n, *dla = NULL;
  dla = (double *)malloc(sizeof(double) * (size_t)(m + 1));
  int status = 2, it, certified = 0; static long nchecks = 0;
  const char *tr_ = getenv("ORC_TRACE"); const int trace2 = tr_ && tr_[0] == '2';
  double cn = 0.0; for (int a = 0; a < n; ++a) if (fabs(cv[a]) > cn) cn = fabs(cv[a]);
  memset(z, 0, sizeof(double) * (size_t)n);
  for (int r = 0; r < m; ++r) { double v = q->h[r]; s[r] = v > 1.0 ? v : 1.0; lam[r] = 1.0; }
  int stall = 0;
  for (it = 0; it < 100; ++it) {
    /* residuals */
    for (int a = 0; a < n; ++a) { double v = cv[a]; const double *Qr = Q + (size_t)a * n; for (int b = 0; b < n; ++b) v += Qr[b] * z[b]; rd[a] = v; }
    double mu = 0.0, rpn = 0.0, rdn = 0.0;
    for (int r = 0; r < m; ++r) {
      const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11;
      double gz = 0.0;
      for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) { gz += g[st + j] * z[st + j]; rd[st + j] += g[st + j] * lam[r]; } }
      rp[r] = gz + s[r] - q->h[r];
      if (fabs(rp[r]) > rpn) rpn = fabs(rp[r]);
      mu += s[r] * lam[r];
    }
    if (m > 0) mu /= m;
    for (int a = 0; a < n; ++a) if (fabs(rd[a]) > rdn) rdn = fabs(rd[a]);
    if (rpn <= 1e-9 && rdn <= 1e-8 * (1.0 + cn) && mu <= 1e-10) { status = 0; break; }
    /* M = Q + G' W G (lower triangle) */
    memcpy(M, Q, sizeof(double) * (size_t)n * n);
    for (int r = 0; r < m; ++r) {
      const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11;
      double wr = lam[r] / s[r]; w[r] = wr;
      for (int a = 0; a < sg[0]; ++a) {
        int sa = sg[1 + 2 * a], la = sg[2 + 2 * a];
        for (int ia = 0; ia < la; ++ia) {
          double gv = wr * g[sa + ia];
          if (gv == 0.0) continue;
          double *Mr = M + (size_t)(sa + ia) * n;
          for (int b = 0; b <= a; ++b) {
            int sb = sg[1 + 2 * b], lb = sg[2 + 2 * b];
            int lim = (b == a) ? ia + 1 : lb;
            for (int ib = 0; ib < lim; ++ib) Mr[sb + ib] += gv * g[sb + ib];
          }
        }
      }
    }
    /* early exit for infeasible nodes: once the multipliers have grown, test them as a Farkas certificate */
    if (it >= 3) { double lm = 0.0; for (int r = 0; r < m; ++r) if (lam[r] > lm) lm = lam[r];
      static double thr = -1.0; if (thr < 0.0) { const char *e = getenv("ORC_FK_THR"); thr = e ? atof(e) : 1e3; }
      if (trace2) fprintf(stderr, "   lm %.3e cn %.3e\n", lm, cn);
      if (lm > thr * (1.0 + cn)) { nchecks++; if (farkas_certificate(k, q, lam, rhs)) { status = 1; certified = 1; break; } } }
    if (chol(M, n)) {
      /* numerically indefinite close to the solution (weights lam/s up to 1e10 on a condensed Hessian): shift the
       * diagonal; the Newton step becomes inexact, the iteration still converges */
      int failed = 1;
      if (mu <= 1e-7 && rpn <= 1e-9) { status = 2; break; }   /* as good as converged: the Lagrangian bound below is tight here */
      for (double shift = 1e-9; failed && shift <= 1e-3; shift *= 100.0) {
        memcpy(M, Q, sizeof(double) * (size_t)n * n);
        for (int r = 0; r < m; ++r) {
          const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11; const double wr = w[r];
          for (int a = 0; a < sg[0]; ++a) { int sa = sg[1 + 2 * a], la = sg[2 + 2 * a];
            for (int ia = 0; ia < la; ++ia) { double gv = wr * g[sa + ia]; if (gv == 0.0) continue; double *Mr = M + (size_t)(sa + ia) * n;
              for (int b = 0; b <= a; ++b) { int sb = sg[1 + 2 * b], lb = sg[2 + 2 * b]; int lim = (b == a) ? ia + 1 : lb; for (int ib = 0; ib < lim; ++ib) Mr[sb + ib] += gv * g[sb + ib]; } } }
        }
        double dmax = 0.0; for (int a = 0; a < n; ++a) if (M[a * n + a] > dmax) dmax = M[a * n + a];
        for (int a = 0; a < n; ++a) M[a * n + a] += shift * dmax;
        failed = chol(M, n);
      }
      if (failed) { status = 2; break; }
    }
    /* predictor: rc = s*lam */
    for (int r = 0; r < m; ++r) t[r] = w[r] * rp[r] - lam[r]; /* (lam*rp - rc)/s */
    for (int a = 0; a < n; ++a) rhs[a] = -rd[a];
    for (int r = 0; r < m; ++r) { const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11; for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) rhs[st + j] -= g[st + j] * t[r]; } }
    memcpy(dz, rhs, sizeof(double) * (size_t)n);
    chol_solve(M, n, dz);
    double aff = 1.0;
    for (int r = 0; r < m; ++r) {
      const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11; double gd = 0.0;
      for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) gd += g[st + j] * dz[st + j]; }
      dsa[r] = -rp[r] - gd;
      dla[r] = -lam[r] - w[r] * dsa[r];
      if (dsa[r] < 0.0) { double a1 = -s[r] / dsa[r]; if (a1 < aff) aff = a1; }
      if (dla[r] < 0.0) { double a1 = -lam[r] / dla[r]; if (a1 < aff) aff = a1; }
    }
    double mu_aff = 0.0;
    for (int r = 0; r < m; ++r) mu_aff += (s[r] + aff * dsa[r]) * (lam[r] + aff * dla[r]);
    if (m > 0) mu_aff /= m;
    double sigma = (mu > 0.0) ? pow(mu_aff / mu, 3.0) : 0.0;
    if (sigma > 1.0) sigma = 1.0;
    /* corrector */
    for (int r = 0; r < m; ++r) { double rc = s[r] * lam[r] + dsa[r] * dla[r] - sigma * mu; t[r] = (lam[r] * rp[r] - rc) / s[r]; }
    for (int a = 0; a < n; ++a) rhs[a] = -rd[a];
    for (int r = 0; r < m; ++r) { const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11; for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) rhs[st + j] -= g[st + j] * t[r]; } }
    memcpy(dz, rhs, sizeof(double) * (size_t)n);
    chol_solve(M, n, dz);
    double alpha = 1.0;
    for (int r = 0; r < m; ++r) {
      const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11; double gd = 0.0;
      for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) gd += g[st + j] * dz[st + j]; }
      ds[r] = -rp[r] - gd;
      double rc = s[r] * lam[r] + dsa[r] * dla[r] - sigma * mu;
      dl[r] = -(rc + lam[r] * ds[r]) / s[r];
      if (ds[r] < 0.0) { double a1 = -s[r] / ds[r]; if (a1 < alpha) alpha = a1; }
      if (dl[r] < 0.0) { double a1 = -lam[r] / dl[r]; if (a1 < alpha) alpha = a1; }
    }
    alpha *= 0.995; if (alpha > 1.0) alpha = 1.0;
    for (int a = 0; a < n; ++a) z[a] += alpha * dz[a];
    for (int r = 0; r < m; ++r) { s[r] += alpha * ds[r]; lam[r] += alpha * dl[r]; }
    if (alpha < 1e-6) { if (++stall >= 5) { status = 2; break; } } else stall = 0;
    double lmax = 0.0; for (int r = 0; r < m; ++r) if (lam[r] > lmax) lmax = lam[r];
    if (trace2) fprintf(stderr, "   it %d alpha %.3e aff %.3e sigma %.3e mu %.3e rpn %.3e rdn %.3e lmax %.3e\n", it, alpha, aff, sigma, mu, rpn, rdn, lmax);
    if (lmax > 1e13) { status = 1; break; }
  }
  k->qp_iters += it;
  k->qp_solves++;
  if (status != 0 && !certified) {
    /* not converged: a primal feasible point is still usable (as an upper bound); otherwise the node is
     * infeasible only with a certificate */
    double worst = 0.0;
    for (int r = 0; r < m; ++r) {
      const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11; double gz = 0.0;
      for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) gz += g[st + j] * z[st + j]; }
      if (gz - q->h[r] > worst) worst = gz - q->h[r];
    }
    if (worst <= 1e-7) status = QP_FEASIBLE_POINT;
    else status = farkas_certificate(k, q, lam, rhs) ? QP_INFEASIBLE : QP_UNKNOWN;
  }
  double lagr = -HUGE_VAL;
  if (status == QP_FEASIBLE_POINT) {
    /* Lagrangian lower bound from the multipliers of the stalled iteration.  For every feasible z':
     *   f(z') >= L(z', lam) >= L(z, lam) + grad_z L(z, lam)'(z' - z)      (L convex in z, lam >= 0)
     * with L(z, lam) = f(z) - lam'(h - Gz) and |z'_a - z_a| bounded through the box of the variables
     * (global jerk box, pair slacks in [0, maximum_slack]). */
    const OrcProblem *p = k->p;
    for (int a = 0; a < n; ++a) { double v = cv[a]; const double *Qr = Q + (size_t)a * n; for (int b = 0; b < n; ++b) v += Qr[b] * z[b]; rd[a] = v; }
    double f = 0.0; for (int a = 0; a < n; ++a) f += z[a] * 0.5 * (rd[a] + cv[a]);
    double comp = 0.0;
    for (int r = 0; r < m; ++r) {
      const double *g = q->G + (size_t)r * n; const int *sg = q->seg + r * 11; double gz = 0.0; const double l = lam[r] > 0.0 ? lam[r] : 0.0;
      for (int a = 0; a < sg[0]; ++a) { int st = sg[1 + 2 * a], ln = sg[2 + 2 * a]; for (int j = 0; j < ln; ++j) { gz += g[st + j] * z[st + j]; rd[st + j] += g[st + j] * l; } }
      comp += l * (q->h[r] - gz);
    }
    double corr = 0.0;
    for (int a = 0; a < n; ++a) {
      const double lo = (a < k->nu) ? p->total_min_jerk : 0.0, hi = (a < k->nu) ? p->total_max_jerk : p->maximum_slack;
      const double range = fmax(hi - z[a], z[a] - lo);
      corr += fabs(rd[a]) * (range > 0.0 ? range : 0.0);
    }
    lagr = f - comp - corr + k->cconst;
    lagr -= 1e-12 * (fabs(f) + fabs(comp) + corr);
  }
  if (getenv("ORC_TRACE")) fprintf(stderr, "[qp] m %d iters %d status %d certified_early %d checks %ld\n", m, it, status, certified, nchecks);
  if (status == QP_OPTIMAL || status == QP_FEASIBLE_POINT) {
    double o = 0.0;
    for (int a = 0; a < n; ++a) { double v = 0.0; const double *Qr = Q + (size_t)a * n; for (int b = 0; b < n; ++b) v += Qr[b] * z[b]; o += z[a] * (0.5 * v + cv[a]); }
    *obj_out = o + k->cconst;
    if (lb_out) *lb_out = (status == QP_OPTIMAL) ? *obj_out : lagr;
  }
  free(s); free(M); free(rd); free(dla);
  return status;
}

/* ---------------------------------------------------------------------------------- */
/* stage-space geometry helpers */

/* trajectory from controls; traj[c][i][8] */
static void simulate(const Ctx *k, const double *z, double *traj) {
  const OrcProblem *p = k->p; const int N = k->N, C = k->C;
  const double ts = p->ts, c2 = 0.5 * (ts * ts), c3 = (1.0 / 6.0) * ((ts * ts) * ts);
  for (int c = 0; c < C; ++c) {
    double *y = traj + (size_t)c * N * 8;
    const double *x0 = p->x0 + 6 * c;
    for (int t = 0; t < 6; ++t) y[t] = x0[t];
    for (int i = 0; i < N; ++i) {
      double *yi = y + i * 8;
      yi[Y_UX] = (i < N - 1) ? z[zidx(k, c, 0, i)] : 0.0;
      yi[Y_UY] = (i < N - 1) ? z[zidx(k, c, 1, i)] : 0.0;
      if (i + 1 < N) {
        double *yn = yi + 8;
        for (int ax = 0; ax < 2; ++ax) {
          double P = yi[ax * 3], V = yi[ax * 3 + 1], A = yi[ax * 3 + 2], U = yi[6 + ax];
          yn[ax * 3] = P + ts * V + c2 * A + c3 * U;
          yn[ax * 3 + 1] = V + ts * A + c2 * U;
          yn[ax * 3 + 2] = A + ts * U;
        }
      }
    }
  }
}

/* front axle affine maps of region j for car c: X = px + fx[0] + fx[1]*vx + fx[2]*vy */
static void front_maps(const OrcProblem *p, int c, int j, double fxu[3], double fxl[3], double fyu[3], double fyl[3]) {
  double wb = p->wheelbase[c];
  for (int t = 0; t < 3; ++t) {
    fxu[t] = wb * p->poly_coss_ub[3 * j + t];
    fxl[t] = wb * p->poly_coss_lb[3 * j + t];
    fyu[t] = wb * p->poly_sint_ub[3 * j + t];
    fyl[t] = wb * p->poly_sint_lb[3 * j + t];
  }
}

/* point pt in {0 rear, 1 (xU,yU), 2 (xL,yU), 3 (xU,yL), 4 (xL,yL)} as affine map of y:
 * X = sum ax[t]*y[t] + x0c ; Y = sum ay[t]*y[t] + y0c   (only px,vx,py,vy used) */
static void point_map(const Ctx *k, int c, int j, int pt, double ax[8], double *xc, double ay[8], double *yc) {
  memset(ax, 0, 8 * sizeof(double)); memset(ay, 0, 8 * sizeof(double));
  ax[Y_PX] = 1.0; ay[Y_PY] = 1.0; *xc = 0.0; *yc = 0.0;
  if (pt == 0) return;
  double fxu[3], fxl[3], fyu[3], fyl[3];
  front_maps(k->p, c, j, fxu, fxl, fyu, fyl);
  const double *fx = (pt == 1 || pt == 3) ? fxu : fxl;
  const double *fy = (pt == 1 || pt == 2) ? fyu : fyl;
  *xc = fx[0]; ax[Y_VX] = fx[1]; ax[Y_VY] = fx[2];
  *yc = fy[0]; ay[Y_VX] = fy[1]; ay[Y_VY] = fy[2];
}

static void srow_clear(SRow *r, int c, int i) { memset(r, 0, sizeof *r); r->c = c; r->i = i; r->c2 = -1; r->slack = -1; }
static void srow_normalize(SRow *r) {
  double n = 0.0;
  for (int t = 0; t < 8; ++t) n += r->a[t] * r->a[t] + r->a2[t] * r->a2[t];
  n += r->as * r->as;
  if (n <= 0.0) return;
  n = 1.0 / sqrt(n);
  for (int t = 0; t < 8; ++t) { r->a[t] *= n; r->a2[t] *= n; }
  r->as *= n; r->rhs *= n;
}
static double srow_eval(const SRow *r, const double *traj, const double *zs, int N) {
  const double *y = traj + ((size_t)r->c * N + r->i) * 8;
  double v = -r->rhs;
  for (int t = 0; t < 8; ++t) v += r->a[t] * y[t];
  if (r->c2 >= 0) { const double *y2 = traj + ((size_t)r->c2 * N + r->i) * 8; for (int t = 0; t < 8; ++t) v += r->a2[t] * y2[t]; }
  if (r->slack >= 0) v += r->as * zs[r->slack];
  return v; /* > 0 means violated */
}

/* edge row: sign=+1 -> cross(P) <= 0 (obstacle, chosen edge); sign=-1 -> cross(P) >= 0 (environment) */
static void edge_row(const Ctx *k, SRow *r, int c, int i, int j, int pt, const double *e, int sign) {
  srow_clear(r, c, i);
  double ax[8], ay[8], xc, yc;
  point_map(k, c, j, pt, ax, &xc, ay, &yc);
  double dx = e[2] - e[0], dy = e[3] - e[1];
  /* cross = dx*(Y - y1) - (X - x1)*dy */
  /* scaled by the edge length only: the row value is the signed distance (in metres) of
   * the point from the edge line, for the rear point and the front corners alike */
  double len = sqrt(dx * dx + dy * dy);
  if (len > 0.0) { dx /= len; dy /= len; }
  double ec = dx * e[1] - e[0] * dy;
  for (int t = 0; t < 8; ++t) r->a[t] = sign * (dx * ay[t] - dy * ax[t]);
  r->rhs = -sign * (dx * yc - dy * xc - ec);
}

/* rows of a rho=0 mode (j,h) that go beyond simple bounds: wedge (2), curvature (2), half-plane (1) */
static int mode_rows(const Ctx *k, SRow *out, int c, int i, int j, int h) {
  const OrcProblem *p = k->p; const double *f = p->frac + 4 * j; int n = 0;
  SRow *r;
  r = &out[n++]; srow_clear(r, c, i); r->a[Y_VY] = -f[0]; r->a[Y_VX] = f[1]; r->rhs = 0.0; srow_normalize(r);  /* f1 vy >= f2 vx */
  r = &out[n++]; srow_clear(r, c, i); r->a[Y_VY] = f[2]; r->a[Y_VX] = -f[3]; r->rhs = 0.0; srow_normalize(r);  /* f3 vy <= f4 vx */
  double sl = (f[1] + f[3]) / (f[0] + f[2]);
  const double *KX = p->poly_kappa_max + 3 * j, *KN = p->poly_kappa_min + 3 * j;
  r = &out[n++]; srow_clear(r, c, i); r->a[Y_AY] = 1.0; r->a[Y_VX] = -KX[1]; r->a[Y_VY] = -KX[2]; r->a[Y_AX] = -sl; r->rhs = KX[0]; srow_normalize(r);
  r = &out[n++]; srow_clear(r, c, i); r->a[Y_AY] = -1.0; r->a[Y_VX] = KN[1]; r->a[Y_VY] = KN[2]; r->a[Y_AX] = sl; r->rhs = -KN[0]; srow_normalize(r);
  double vm = p->min_region_change_speed;
  r = &out[n++]; srow_clear(r, c, i);
  switch (h) {
    case 0: r->a[Y_VX] = -1.0; r->rhs = -vm; break; /* vx >= vm  (b_xp = 0) */
    case 1: r->a[Y_VY] = -1.0; r->rhs = -vm; break; /* vy >= vm  (b_yp = 0) */
    case 2: r->a[Y_VX] = 1.0; r->rhs = -vm; break;  /* vx <= -vm (b_xn = 0) */
    default: r->a[Y_VY] = 1.0; r->rhs = -vm; break; /* vy <= -vm (b_yn = 0) */
  }
  return n;
}

/* bound rows of a stage: lo[8], hi[8] (infinite = none) */
static void stage_bounds(const Ctx *k, int c, int i, int jeff, int frozen, double lo[8], double hi[8]) {
  const OrcProblem *p = k->p; const int R = k->R;
  for (int t = 0; t < 8; ++t) { lo[t] = -HUGE_VAL; hi[t] = HUGE_VAL; }
  lo[Y_VX] = p->min_vel; hi[Y_VX] = p->max_vel; lo[Y_VY] = p->min_vel; /* vel_y has no upper bound */
  lo[Y_AX] = p->total_min_acc; hi[Y_AX] = p->total_max_acc; lo[Y_AY] = p->total_min_acc; hi[Y_AY] = p->total_max_acc;
  lo[Y_UX] = p->total_min_jerk; hi[Y_UX] = p->total_max_jerk; lo[Y_UY] = p->total_min_jerk; hi[Y_UY] = p->total_max_jerk;
  if (jeff >= 0) {
    int q = c * R + jeff;
#define TIGHT_LO(t, v) if ((v) > lo[t]) lo[t] = (v)
#define TIGHT_HI(t, v) if ((v) < hi[t]) hi[t] = (v)
    TIGHT_LO(Y_UX, p->min_jerk_x[q]); TIGHT_HI(Y_UX, p->max_jerk_x[q]);
    TIGHT_LO(Y_UY, p->min_jerk_y[q]); TIGHT_HI(Y_UY, p->max_jerk_y[q]);
    if (i > 0) {
      TIGHT_LO(Y_AX, p->min_acc_x[q]); TIGHT_HI(Y_AX, p->max_acc_x[q]);
      TIGHT_LO(Y_AY, p->min_acc_y[q]); TIGHT_HI(Y_AY, p->max_acc_y[q]);
    }
  }
  if (frozen) {
    double vm = p->min_region_change_speed;
    TIGHT_LO(Y_VX, -vm); TIGHT_HI(Y_VX, vm); TIGHT_LO(Y_VY, -vm); TIGHT_HI(Y_VY, vm);
  }
}

/* effective region per (c,i) from decisions; -1 unknown */
static void effective_regions(const Ctx *k, const unsigned char *dec, int *jeff) {
  for (int c = 0; c < k->C; ++c) {
    int jp = k->p->initial_region[c] - 1;
    jeff[c * k->N] = jp;
    for (int i = 1; i < k->N; ++i) {
      unsigned char m = dec[k->off_mode + c * k->N + i];
      int j;
      if (m == UNDEC) j = -1;
      else if (m == MODE_FROZEN) j = jp;
      else j = m >> 2;
      jeff[c * k->N + i] = j; jp = j;
    }
  }
}

/* pair quadruple rows. q: 0 rear/rear, 1 rear a/front b, 2 rear b/front a, 3 front/front; side 0..3 */
static void pair_row(const Ctx *k, SRow *r, int a, int b, int i, int q, int side, int ja, int jb) {
  const OrcProblem *p = k->p;
  srow_clear(r, a, i); r->c2 = b;
  double RR = p->radius[a] + p->radius[b];
  double D = RR + p->safety[i], Ds = D + p->safety_slack[i];
  int pr = pair_index(k, a, b);
  double fxu_a[3] = {0}, fxl_a[3] = {0}, fyu_a[3] = {0}, fyl_a[3] = {0}, fxu_b[3] = {0}, fxl_b[3] = {0}, fyu_b[3] = {0}, fyl_b[3] = {0};
  if (i == 0) { /* initial_conditions.mod:20-23: front box collapsed onto the heading */
    double tha = atan2(p->x0[6 * a + 4], p->x0[6 * a + 1]), thb = atan2(p->x0[6 * b + 4], p->x0[6 * b + 1]);
    fxu_a[0] = fxl_a[0] = cos(tha) * p->wheelbase[a]; fyu_a[0] = fyl_a[0] = sin(tha) * p->wheelbase[a];
    fxu_b[0] = fxl_b[0] = cos(thb) * p->wheelbase[b]; fyu_b[0] = fyl_b[0] = sin(thb) * p->wheelbase[b];
  } else {
    if (ja >= 0) front_maps(p, a, ja, fxu_a, fxl_a, fyu_a, fyl_a);
    if (jb >= 0) front_maps(p, b, jb, fxu_b, fxl_b, fyu_b, fyl_b);
  }
  double cst = 0.0; /* lhs: expr + cst <= 0 */
#define ADD_REAR(arr, axis, sg) arr[(axis) ? Y_PY : Y_PX] += (sg)
#define ADD_FRONT(arr, axis, f, sg) do { arr[(axis) ? Y_PY : Y_PX] += (sg); arr[Y_VX] += (sg) * f[1]; arr[Y_VY] += (sg) * f[2]; cst += (sg) * f[0]; } while (0)
  int axis = side >> 1; /* sides 0,1 -> x ; 2,3 -> y */
  int first = (side & 1) == 0; /* first: "a below b" form */
  if (q == 0) {
    /* pa <= pb - (Ds - s)  |  pa >= pb + (Ds - s) */
    if (first) { ADD_REAR(r->a, axis, 1.0); ADD_REAR(r->a2, axis, -1.0); } else { ADD_REAR(r->a, axis, -1.0); ADD_REAR(r->a2, axis, 1.0); }
    cst += Ds; r->slack = sidx(k, pr, i, axis); r->as = -1.0;
  } else if (q == 1) {
    /* pa <= fLB_b - D | pa >= fUB_b + D */
    const double *fl = axis ? fyl_b : fxl_b, *fu = axis ? fyu_b : fxu_b;
    if (first) { ADD_REAR(r->a, axis, 1.0); ADD_FRONT(r->a2, axis, fl, -1.0); } else { ADD_REAR(r->a, axis, -1.0); ADD_FRONT(r->a2, axis, fu, 1.0); }
    cst += D;
  } else if (q == 2) {
    const double *fl = axis ? fyl_a : fxl_a, *fu = axis ? fyu_a : fxu_a;
    if (first) { ADD_REAR(r->a2, axis, 1.0); ADD_FRONT(r->a, axis, fl, -1.0); } else { ADD_REAR(r->a2, axis, -1.0); ADD_FRONT(r->a, axis, fu, 1.0); }
    cst += D;
  } else {
    /* side even: 0 <= fLB_a - (Ds - s) - fUB_b  ->  fUB_b - fLB_a + Ds - s <= 0
       side odd : 0 >= fUB_a + (Ds - s) - fLB_b  ->  fUB_a - fLB_b + Ds - s <= 0 */
    const double *fla = axis ? fyl_a : fxl_a, *fua = axis ? fyu_a : fxu_a, *flb = axis ? fyl_b : fxl_b, *fub = axis ? fyu_b : fxu_b;
    if (first) { ADD_FRONT(r->a2, axis, fub, 1.0); ADD_FRONT(r->a, axis, fla, -1.0); } else { ADD_FRONT(r->a, axis, fua, 1.0); ADD_FRONT(r->a2, axis, flb, -1.0); }
    cst += Ds; r->slack = sidx(k, pr, i, 2 + axis); r->as = -1.0;
  }
#undef ADD_REAR
#undef ADD_FRONT
  r->rhs = -cst;
  /* no normalisation: coefficients are +-1 and O(0.1) */
}

/* ---------------------------------------------------------------------------------- */
typedef struct Node { double bound; int depth; int rank; unsigned char *dec; } Node;

typedef struct Heap { Node *a; int n, cap; int have_inc; } Heap;
static int node_before(const Heap *h, const Node *x, const Node *y) {
  if (!h->have_inc) { if (x->depth != y->depth) return x->depth > y->depth; if (x->rank != y->rank) return x->rank < y->rank; return x->bound < y->bound; }
  if (x->bound != y->bound) return x->bound < y->bound;
  if (x->depth != y->depth) return x->depth > y->depth;
  return x->rank < y->rank;
}
static void heap_push(Heap *h, Node nd) {
  if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 1024; h->a = (Node *)realloc(h->a, sizeof(Node) * (size_t)h->cap); }
  int i = h->n++; h->a[i] = nd;
  while (i > 0) { int pr = (i - 1) / 2; if (node_before(h, &h->a[i], &h->a[pr])) { Node t = h->a[i]; h->a[i] = h->a[pr]; h->a[pr] = t; i = pr; } else break; }
}
static void heap_sift(Heap *h, int i) {
  for (;;) { int l = 2 * i + 1, r = l + 1, b = i; if (l < h->n && node_before(h, &h->a[l], &h->a[b])) b = l; if (r < h->n && node_before(h, &h->a[r], &h->a[b])) b = r; if (b == i) break; Node t = h->a[i]; h->a[i] = h->a[b]; h->a[b] = t; i = b; }
}
static Node heap_pop(Heap *h) { Node t = h->a[0]; h->a[0] = h->a[--h->n]; if (h->n > 0) heap_sift(h, 0); return t; }
static void heap_rebuild(Heap *h) { for (int i = h->n / 2 - 1; i >= 0; --i) heap_sift(h, i); }

/* ---------------------------------------------------------------------------------- */
static void ctx_init(Ctx *k, const OrcProblem *p) {
  memset(k, 0, sizeof *k);
  k->p = p; orc_layout(p, &k->lay);
  k->C = p->C; k->N = p->N; k->R = p->R; k->O = p->O; k->L = p->L; k->E = p->E;
  k->P = p->C * (p->C - 1) / 2;
  const int N = k->N, Nm = N - 1, C = k->C;
  k->nu = 2 * Nm * C; k->ns = 4 * k->P * N; k->n = k->nu + k->ns;
  k->off_mode = 0; k->off_env = k->off_mode + C * N; k->off_obs = k->off_env + 5 * C * N;
  k->off_pair = k->off_obs + 5 * C * k->O * N; k->ndec = k->off_pair + 4 * k->P * N;
  k->feas_tol = 1e-6;
  const double ts = p->ts, c2 = 0.5 * (ts * ts), c3 = (1.0 / 6.0) * ((ts * ts) * ts);
  k->Sp = (double *)calloc((size_t)N * Nm, sizeof(double));
  k->Sv = (double *)calloc((size_t)N * Nm, sizeof(double));
  k->Sa = (double *)calloc((size_t)N * Nm, sizeof(double));
  for (int j = 0; j < Nm; ++j) {
    double P = 0, V = 0, A = 0;
    for (int i = j; i < Nm; ++i) {
      double U = (i == j) ? 1.0 : 0.0;
      double Pn = P + ts * V + c2 * A + c3 * U, Vn = V + ts * A + c2 * U, An = A + ts * U;
      P = Pn; V = Vn; A = An;
      k->Sp[(i + 1) * Nm + j] = P; k->Sv[(i + 1) * Nm + j] = V; k->Sa[(i + 1) * Nm + j] = A;
    }
  }
  k->y0 = (double *)calloc((size_t)C * N * 6, sizeof(double));
  for (int c = 0; c < C; ++c) {
    double *y = k->y0 + (size_t)c * N * 6;
    for (int t = 0; t < 6; ++t) y[t] = p->x0[6 * c + t];
    for (int i = 0; i + 1 < N; ++i)
      for (int ax = 0; ax < 2; ++ax) {
        double P = y[i * 6 + ax * 3], V = y[i * 6 + ax * 3 + 1], A = y[i * 6 + ax * 3 + 2];
        y[(i + 1) * 6 + ax * 3] = P + ts * V + c2 * A;
        y[(i + 1) * 6 + ax * 3 + 1] = V + ts * A;
        y[(i + 1) * 6 + ax * 3 + 2] = A;
      }
  }
  /* objective: sum w (y - ref)^2 -> 1/2 z'Qz + c'z + const */
  const int n = k->n;
  k->Q = (double *)calloc((size_t)n * n, sizeof(double));
  k->cvec = (double *)calloc((size_t)n, sizeof(double));
  k->cconst = 0.0;
  for (int c = 0; c < C; ++c)
    for (int ax = 0; ax < 2; ++ax) {
      double wp = ax ? p->w_pos_y[c] : p->w_pos_x[c], wv = ax ? p->w_vel_y[c] : p->w_vel_x[c], wa = ax ? p->w_acc_y[c] : p->w_acc_x[c];
      double wj = ax ? p->w_jerk_y[c] : p->w_jerk_x[c];
      const double *pref = ax ? p->y_ref : p->x_ref, *vref = ax ? p->vy_ref : p->vx_ref;
      for (int i = 0; i < N; ++i) {
        const double *y0 = k->y0 + ((size_t)c * N + i) * 6 + ax * 3;
        double ep = y0[0] - pref[c * N + i], ev = y0[1] - vref[c * N + i], ea = y0[2];
        k->cconst += wp * ep * ep + wv * ev * ev + wa * ea * ea;
        for (int j = 0; j < i; ++j) {
          int zj = zidx(k, c, ax, j);
          double sp = k->Sp[i * Nm + j], sv = k->Sv[i * Nm + j], sa = k->Sa[i * Nm + j];
          k->cvec[zj] += 2.0 * (wp * ep * sp + wv * ev * sv + wa * ea * sa);
          for (int l = 0; l < i; ++l) {
            int zl = zidx(k, c, ax, l);
            k->Q[(size_t)zj * n + zl] += 2.0 * (wp * sp * k->Sp[i * Nm + l] + wv * sv * k->Sv[i * Nm + l] + wa * sa * k->Sa[i * Nm + l]);
          }
        }
      }
      for (int j = 0; j < Nm; ++j) { int zj = zidx(k, c, ax, j); k->Q[(size_t)zj * n + zj] += 2.0 * wj + 1e-10; }
    }
  for (int s = k->nu; s < n; ++s) k->Q[(size_t)s * n + s] = 2.0 * p->w_slack;
  /* mode alternatives per car: every possible region with its non-dominated half planes */
  k->nalt_mode = (int *)calloc((size_t)C, sizeof(int));
  k->alt_mode = (unsigned char *)calloc((size_t)C * k->R * 4, 1);
  for (int c = 0; c < C; ++c) {
    int na = 0;
    for (int j = 0; j < k->R; ++j) {
      if (p->possible_region[c * k->R + j] != 1) continue;
      /* the wedge is the cone between directions (f0,f1) and (f2,f3); half plane h is useful iff
       * some point of the cone satisfies it, and is dominated if another useful half plane
       * contains cone /\ h.  Sample the two rays. */
      const double *f = p->frac + 4 * j;
      double d1x = f[0], d1y = f[1], d2x = f[2], d2y = f[3];
      int useful[4];
      useful[0] = (d1x > 1e-9 || d2x > 1e-9); useful[1] = (d1y > 1e-9 || d2y > 1e-9);
      useful[2] = (d1x < -1e-9 || d2x < -1e-9); useful[3] = (d1y < -1e-9 || d2y < -1e-9);
      /* dominance: within the cone, |vx| >= |vy| everywhere -> vy half planes dominated by vx, and vice versa */
      int x_dom = (fabs(d1x) >= fabs(d1y) - 1e-9) && (fabs(d2x) >= fabs(d2y) - 1e-9);
      int y_dom = (fabs(d1y) >= fabs(d1x) - 1e-9) && (fabs(d2y) >= fabs(d2x) - 1e-9);
      if (x_dom && (useful[0] || useful[2])) { useful[1] = 0; useful[3] = 0; }
      else if (y_dom && (useful[1] || useful[3])) { useful[0] = 0; useful[2] = 0; }
      for (int h = 0; h < 4; ++h) if (useful[h]) k->alt_mode[c * k->R * 4 + na++] = (unsigned char)(j * 4 + h);
    }
    k->nalt_mode[c] = na;
  }
}
static void ctx_free(Ctx *k) { free(k->Sp); free(k->Sv); free(k->Sa); free(k->y0); free(k->Q); free(k->cvec); free(k->nalt_mode); free(k->alt_mode); }

/* Build the node QP rows.  Returns 0 ok, 1 trivially infeasible.  *penalty gets the constant
 * cost of SOFT decisions. */
static int build_node_qp(Ctx *k, const unsigned char *dec, QP *q, double *penalty) {
  const OrcProblem *p = k->p; const int C = k->C, N = k->N, E = k->E, O = k->O;
  int *jeff = (int *)malloc(sizeof(int) * (size_t)C * N);
  effective_regions(k, dec, jeff);
  SRow r, mr[8];
  double pen = 0.0, cv;
  int infeas = 0;
  q->m = 0;
#define PUSH(row) do { cv = 0.0; if (!condense(k, q, (row), &cv)) { if (cv > 1e-7) infeas = 1; } } while (0)
  for (int c = 0; c < C && !infeas; ++c)
    for (int i = 0; i < N && !infeas; ++i) {
      unsigned char m = dec[k->off_mode + c * N + i];
      int je = jeff[c * N + i];
      double lo[8], hi[8];
      stage_bounds(k, c, i, je, (i > 0 && m == MODE_FROZEN), lo, hi);
      for (int t = (i == 0 ? 6 : 1); t < 8; ++t) {
        if (t == Y_PY) continue;
        if (lo[t] > hi[t] + 1e-12) { infeas = 1; break; }
        if (i == N - 1 && t >= 6) { if (lo[t] > 1e-9 || hi[t] < -1e-9) infeas = 1; continue; }
        if (hi[t] < HUGE_VAL) { srow_clear(&r, c, i); r.a[t] = 1.0; r.rhs = hi[t]; PUSH(&r); }
        if (lo[t] > -HUGE_VAL) { srow_clear(&r, c, i); r.a[t] = -1.0; r.rhs = -lo[t]; PUSH(&r); }
      }
      if (i == 0) continue;
      if (m != UNDEC && m != MODE_FROZEN) { int nr = mode_rows(k, mr, c, i, m >> 2, m & 3); for (int a = 0; a < nr; ++a) PUSH(&mr[a]); }
      /* environment */
      if (E > 0)
        for (int pt = 0; pt < 5; ++pt) {
          int e = (E == 1) ? 0 : dec[k->off_env + (c * N + i) * 5 + pt];
          if (e == UNDEC) continue;
          if (pt > 0 && je < 0) continue;
          for (int ed = p->env_off[e]; ed < p->env_off[e + 1]; ++ed) { edge_row(k, &r, c, i, je, pt, p->env_edges + 4 * ed, -1); PUSH(&r); }
        }
      for (int o = 0; o < O; ++o)
        for (int pt = 0; pt < 5; ++pt) {
          unsigned char d = dec[k->off_obs + ((c * O + o) * N + i) * 5 + pt];
          if (d == UNDEC) continue;
          if (d == OBS_SOFT) { pen += p->w_slack_obs; continue; }
          if (pt > 0 && je < 0) continue;
          edge_row(k, &r, c, i, je, pt, p->obs_edges + 4 * ((o * N + i) * k->L + d), +1); PUSH(&r);
        }
    }
  /* pairs */
  for (int a = 0; a < C - 1 && !infeas; ++a)
    for (int b = a + 1; b < C; ++b) {
      int pr = pair_index(k, a, b);
      for (int i = 0; i < N; ++i) {
        double cap = p->safety_slack[i] < p->maximum_slack ? p->safety_slack[i] : p->maximum_slack;
        for (int sq = 0; sq < 4; ++sq) { /* slack bounds 0 <= s <= cap */
          srow_clear(&r, a, i); r.slack = sidx(k, pr, i, sq); r.as = 1.0; r.rhs = cap; PUSH(&r);
          srow_clear(&r, a, i); r.slack = sidx(k, pr, i, sq); r.as = -1.0; r.rhs = 0.0; PUSH(&r);
        }
        for (int qd = 0; qd < 4; ++qd) {
          unsigned char d = dec[k->off_pair + (pr * N + i) * 4 + qd];
          if (d == UNDEC) continue;
          int ja = jeff[a * N + i], jb = jeff[b * N + i];
          if ((qd == 1 || qd == 3) && jb < 0) continue;
          if ((qd == 2 || qd == 3) && ja < 0) continue;
          pair_row(k, &r, a, b, i, qd, d, ja, jb); PUSH(&r);
        }
      }
    }
#undef PUSH
  free(jeff);
  *penalty = pen;
  return infeas;
}

/* ---------------------------------------------------------------------------------- */
/* Scan: implied alternatives + violations.  Fills `imp` (a complete decision vector that
 * extends dec) and returns the most violated disjunction.  */
typedef struct Branch { int kind; int c, i, o, pt, pr, q; double viol; } Branch; /* kind: 0 none, 1 mode, 2 env, 3 obs, 4 pair */

static double rows_maxviol(const SRow *rows, int n, const double *traj, const double *z, int N) {
  double v = -HUGE_VAL; for (int a = 0; a < n; ++a) { double e = srow_eval(&rows[a], traj, z, N); if (e > v) v = e; } return v;
}

static double mode_alt_violation(const Ctx *k, int c, int i, int alt /* j*4+h or MODE_FROZEN */, int jprev, const double *traj, const double *z) {
  SRow mr[8]; double lo[8], hi[8]; double v = -HUGE_VAL;
  const double *y = traj + ((size_t)c * k->N + i) * 8;
  int j = (alt == MODE_FROZEN) ? jprev : (alt >> 2);
  stage_bounds(k, c, i, j, alt == MODE_FROZEN, lo, hi);
  for (int t = 1; t < 8; ++t) {
    if (t == Y_PY) continue;
    if (i == k->N - 1 && t >= 6) { if (lo[t] > 0 && lo[t] > v) v = lo[t]; if (hi[t] < 0 && -hi[t] > v) v = -hi[t]; continue; }
    if (y[t] - hi[t] > v) v = y[t] - hi[t];
    if (lo[t] - y[t] > v) v = lo[t] - y[t];
  }
  if (alt != MODE_FROZEN) { int nr = mode_rows(k, mr, c, i, alt >> 2, alt & 3); double e = rows_maxviol(mr, nr, traj, z, k->N); if (e > v) v = e; }
  return v;
}

static void scan(Ctx *k, const unsigned char *dec, const double *traj, const double *z, unsigned char *imp, Branch *br) {
  const OrcProblem *p = k->p; const int C = k->C, N = k->N, E = k->E, O = k->O, L = k->L;
  const double tol = k->feas_tol;
  memcpy(imp, dec, (size_t)k->ndec);
  br->kind = 0; br->viol = tol;
  int *jeff = (int *)malloc(sizeof(int) * (size_t)C * N);
  SRow r;
  for (int c = 0; c < C; ++c) {
    int jp = p->initial_region[c] - 1;
    jeff[c * N] = jp;
    int root_undec = -1; /* earliest undecided step that the current effective region depends on, -1 if none */
    for (int i = 1; i < N; ++i) {
      unsigned char m = dec[k->off_mode + c * N + i];
      int j;
      double mviol = 0.0; int blame = i;
      if (m == UNDEC) {
        int best = -1; double bv = HUGE_VAL;
        double vfz = mode_alt_violation(k, c, i, MODE_FROZEN, jp, traj, z);
        if (vfz < bv) { bv = vfz; best = MODE_FROZEN; }
        for (int a = 0; a < k->nalt_mode[c]; ++a) {
          int alt = k->alt_mode[c * k->R * 4 + a];
          double v = mode_alt_violation(k, c, i, alt, jp, traj, z);
          if (v < bv - 1e-12) { bv = v; best = alt; }
        }
        imp[k->off_mode + c * N + i] = (unsigned char)best;
        j = (best == MODE_FROZEN) ? jp : (best >> 2);
        mviol = bv;
        root_undec = (best == MODE_FROZEN && root_undec >= 0) ? root_undec : i;
        if (mviol > br->viol) { br->kind = 1; br->c = c; br->i = i; br->viol = mviol; }
      } else if (m == MODE_FROZEN) {
        j = jp;
        if (root_undec >= 0) {
          /* region of this frozen step is only implied: check its region rows, blame the chain root */
          mviol = mode_alt_violation(k, c, i, MODE_FROZEN, jp, traj, z); blame = root_undec;
          if (mviol > br->viol) { br->kind = 1; br->c = c; br->i = blame; br->viol = mviol; }
        }
      } else { j = m >> 2; root_undec = -1; }
      jeff[c * N + i] = j; jp = j;
      int region_decided = (root_undec < 0);
      int mode_blame = root_undec;
      /* environment points */
      if (E > 0)
        for (int pt = 0; pt < 5; ++pt) {
          unsigned char d = (E == 1) ? 0 : dec[k->off_env + (c * N + i) * 5 + pt];
          int enforced = (d != UNDEC) && (pt == 0 || region_decided);
          if (enforced) continue;
          int best = -1; double bv = HUGE_VAL;
          for (int e = 0; e < E; ++e) {
            if (d != UNDEC && e != d) continue;
            double v = -HUGE_VAL;
            for (int ed = p->env_off[e]; ed < p->env_off[e + 1]; ++ed) { edge_row(k, &r, c, i, j, pt, p->env_edges + 4 * ed, -1); double ev = srow_eval(&r, traj, z, N); if (ev > v) v = ev; }
            if (v < bv) { bv = v; best = e; }
          }
          if (E > 1) imp[k->off_env + (c * N + i) * 5 + pt] = (unsigned char)best;
          if (bv > br->viol) {
            if (pt > 0 && !region_decided) { br->kind = 1; br->c = c; br->i = mode_blame; br->viol = bv; }
            else { br->kind = 2; br->c = c; br->i = i; br->pt = pt; br->viol = bv; }
          }
        }
      for (int o = 0; o < O; ++o)
        for (int pt = 0; pt < 5; ++pt) {
          unsigned char d = dec[k->off_obs + ((c * O + o) * N + i) * 5 + pt];
          if (d == OBS_SOFT) continue;
          int enforced = (d != UNDEC) && (pt == 0 || region_decided);
          if (enforced) continue;
          int ne = p->obs_nedges[o * N + i];
          int best = -1; double bv = HUGE_VAL;
          for (int ed = 0; ed < ne; ++ed) {
            if (d != UNDEC && ed != d) continue;
            edge_row(k, &r, c, i, j, pt, p->obs_edges + 4 * ((o * N + i) * L + ed), +1);
            double v = srow_eval(&r, traj, z, N);
            if (v < bv) { bv = v; best = ed; }
          }
          if (ne == 0) { bv = -1.0; best = 0; }
          imp[k->off_obs + ((c * O + o) * N + i) * 5 + pt] = (unsigned char)best;
          if (bv > br->viol) {
            if (pt > 0 && !region_decided) { br->kind = 1; br->c = c; br->i = mode_blame; br->viol = bv; }
            else { br->kind = 3; br->c = c; br->i = i; br->o = o; br->pt = pt; br->viol = bv; }
          }
        }
    }
  }
  /* pairs */
  for (int a = 0; a < C - 1; ++a)
    for (int b = a + 1; b < C; ++b) {
      int pr = pair_index(k, a, b);
      for (int i = 0; i < N; ++i) {
        int ja = jeff[a * N + i], jb = jeff[b * N + i];
        unsigned char ma = (i == 0) ? 0 : dec[k->off_mode + a * N + i], mb = (i == 0) ? 0 : dec[k->off_mode + b * N + i];
        for (int qd = 0; qd < 4; ++qd) {
          unsigned char d = dec[k->off_pair + (pr * N + i) * 4 + qd];
          int need_a = (qd == 2 || qd == 3), need_b = (qd == 1 || qd == 3);
          int enforced = (d != UNDEC) && (!need_a || ma != UNDEC) && (!need_b || mb != UNDEC);
          if (enforced) continue;
          int best = -1; double bv = HUGE_VAL;
          for (int side = 0; side < 4; ++side) {
            if (d != UNDEC && side != d) continue;
            pair_row(k, &r, a, b, i, qd, side, ja, jb);
            double v = srow_eval(&r, traj, z, N);
            if (v < bv) { bv = v; best = side; }
          }
          imp[k->off_pair + (pr * N + i) * 4 + qd] = (unsigned char)best;
          if (i == 0) continue;
          if (bv > br->viol) {
            if (need_a && ma == UNDEC) { br->kind = 1; br->c = a; br->i = i; br->viol = bv; }
            else if (need_b && mb == UNDEC) { br->kind = 1; br->c = b; br->i = i; br->viol = bv; }
            else { br->kind = 4; br->pr = pr; br->i = i; br->q = qd; br->viol = bv; }
          }
        }
      }
    }
  free(jeff);
}

/* write full column vector from a fully decided node + trajectory */
static void fill_solution(Ctx *k, const unsigned char *dec, const double *traj, const double *z, double *x) {
  const OrcProblem *p = k->p; const OrcLayout *l = &k->lay;
  const int C = k->C, N = k->N, R = k->R, E = k->E, O = k->O, L = k->L, K = l->K;
  memset(x, 0, sizeof(double) * (size_t)l->ncols);
  int *jeff = (int *)malloc(sizeof(int) * (size_t)C * N);
  effective_regions(k, dec, jeff);
  const double vm = p->min_region_change_speed;
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < N; ++i) {
      const double *y = traj + ((size_t)c * N + i) * 8;
      const int blk[8] = {2, 3, 4, 5, 6, 7, 0, 1};
      for (int t = 0; t < 8; ++t) x[(blk[t] * C + c) * N + i] = y[t];
      int j = jeff[c * N + i];
      double fxu[3], fxl[3], fyu[3], fyl[3];
      if (i == 0) {
        double th = atan2(p->x0[6 * c + 4], p->x0[6 * c + 1]);
        double fx = p->x0[6 * c + 0] + cos(th) * p->wheelbase[c], fy = p->x0[6 * c + 3] + sin(th) * p->wheelbase[c];
        x[(8 * C + c) * N] = fx; x[(9 * C + c) * N] = fx; x[(10 * C + c) * N] = fy; x[(11 * C + c) * N] = fy;
      } else {
        front_maps(p, c, j, fxu, fxl, fyu, fyl);
        x[(8 * C + c) * N + i] = y[Y_PX] + (fxu[0] + fxu[1] * y[Y_VX] + fxu[2] * y[Y_VY]);
        x[(9 * C + c) * N + i] = y[Y_PX] + (fxl[0] + fxl[1] * y[Y_VX] + fxl[2] * y[Y_VY]);
        x[(10 * C + c) * N + i] = y[Y_PY] + (fyu[0] + fyu[1] * y[Y_VX] + fyu[2] * y[Y_VY]);
        x[(11 * C + c) * N + i] = y[Y_PY] + (fyl[0] + fyl[1] * y[Y_VX] + fyl[2] * y[Y_VY]);
      }
      x[l->base_ar + (c * N + i) * R + j] = 1.0;
      if (i > 0) {
        unsigned char m = dec[k->off_mode + c * N + i];
        double b[4];
        b[0] = (y[Y_VX] <= vm) ? 1.0 : 0.0;   /* x_positive */
        b[1] = (y[Y_VY] <= vm) ? 1.0 : 0.0;   /* y_positive */
        b[2] = (y[Y_VX] >= -vm) ? 1.0 : 0.0;  /* x_negative */
        b[3] = (y[Y_VY] >= -vm) ? 1.0 : 0.0;  /* y_negative */
        double rho = 0.0;
        if (m == MODE_FROZEN) { b[0] = b[1] = b[2] = b[3] = 1.0; rho = 1.0; }
        else { int h = m & 3; const int map[4] = {0, 1, 2, 3}; b[map[h]] = 0.0; }
        for (int t = 0; t < 4; ++t) x[l->base_rcna + (t * C + c) * N + i] = b[t];
        x[l->base_rcna + (4 * C + c) * N + i] = rho;
      }
      for (int pt = 0; pt < 5; ++pt) {
        /* coordinates of this point (needed at step 0 where nothing is decided: the state is fixed) */
        double PX = (pt == 0) ? y[Y_PX] : x[(((pt == 1 || pt == 3) ? 8 : 9) * C + c) * N + i];
        double PY = (pt == 0) ? y[Y_PY] : x[(((pt == 1 || pt == 2) ? 10 : 11) * C + c) * N + i];
        if (E > 0) {
          int e = (E == 1) ? 0 : dec[k->off_env + (c * N + i) * 5 + pt];
          if (i == 0 && E > 1) {
            double bestv = -HUGE_VAL; e = 0;
            for (int ee = 0; ee < E; ++ee) { double mn = HUGE_VAL;
              for (int ed = p->env_off[ee]; ed < p->env_off[ee + 1]; ++ed) { const double *g = p->env_edges + 4 * ed; double cr = (g[2] - g[0]) * (PY - g[1]) - (PX - g[0]) * (g[3] - g[1]); if (cr < mn) mn = cr; }
              if (mn > bestv) { bestv = mn; e = ee; } }
          }
          for (int ee = 0; ee < E; ++ee) x[l->base_nwe + ((pt * C + c) * E + ee) * N + i] = (ee == e) ? 0.0 : 1.0;
        }
        for (int o = 0; o < O; ++o) {
          unsigned char d = dec[k->off_obs + ((c * O + o) * N + i) * 5 + pt];
          int ne = p->obs_nedges[o * N + i];
          if (i == 0) {
            double bestv = HUGE_VAL; d = 0;
            for (int ed = 0; ed < ne; ++ed) { const double *g = p->obs_edges + 4 * ((o * N + i) * L + ed); double cr = (g[2] - g[0]) * (PY - g[1]) - (PX - g[0]) * (g[3] - g[1]); if (cr < bestv) { bestv = cr; d = (unsigned char)ed; } }
            if (bestv > 1e-9 && p->obs_soft[o] == 1) d = OBS_SOFT;
          }
          /* env point order (UU,LU,UL,LL) -> obstacle front index f = 5 - pt (LL,UL,LU,UU) */
          for (int ed = 0; ed < ne; ++ed) {
            double v = (d == OBS_SOFT || ed != d) ? 1.0 : 0.0;
            if (pt == 0) x[l->base_dcc + ((c * O + o) * N + i) * L + ed] = v;
            else x[l->base_dcf + (((c * O + o) * N + i) * L + ed) * 4 + (4 - pt)] = v;
          }
          if (d == OBS_SOFT) { if (pt == 0) x[l->base_so + (c * O + o) * N + i] = 1.0; else x[l->base_sof + ((c * O + o) * N + i) * 4 + (4 - pt)] = 1.0; }
        }
      }
    }
  for (int a = 0; a < C - 1; ++a)
    for (int b = a + 1; b < C; ++b) {
      int pr = pair_index(k, a, b), k1 = a, k2 = b - 1;
      for (int i = 0; i < N; ++i) {
        for (int qd = 0; qd < 4; ++qd) {
          unsigned char d = dec[k->off_pair + (pr * N + i) * 4 + qd];
          for (int side = 0; side < 4; ++side) x[l->base_c2c + ((k1 * K + k2) * N + i) * 16 + qd * 4 + side] = (side == d) ? 0.0 : 1.0;
        }
        for (int sq = 0; sq < 4; ++sq) { double v = z[sidx(k, pr, i, sq)]; if (v < 0) v = 0; x[l->base_sv + ((k1 * K + k2) * N + i) * 4 + sq] = v; }
      }
    }
  free(jeff);
}

/* decisions from the binaries of a full vector (MIP start / golden vector) */
static void decisions_from_solution(Ctx *k, const double *x, unsigned char *dec) {
  const OrcProblem *p = k->p; const OrcLayout *l = &k->lay;
  const int C = k->C, N = k->N, R = k->R, E = k->E, O = k->O, L = k->L, K = l->K;
  memset(dec, UNDEC, (size_t)k->ndec);
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < N; ++i) {
      int j = 0; for (int jj = 0; jj < R; ++jj) if (x[l->base_ar + (c * N + i) * R + jj] > 0.5) j = jj;
      if (i > 0) {
        double rho = x[l->base_rcna + (4 * C + c) * N + i];
        if (rho > 0.5) dec[k->off_mode + c * N + i] = MODE_FROZEN;
        else {
          int h = -1;
          for (int t = 0; t < 4; ++t) if (x[l->base_rcna + (t * C + c) * N + i] < 0.5) { h = t; break; }
          if (h < 0) h = 0;
          /* use a non-dominated half plane of region j if the stored one is not in the list */
          int found = 0, firsth = -1;
          for (int a = 0; a < k->nalt_mode[c]; ++a) { int alt = k->alt_mode[c * R * 4 + a]; if ((alt >> 2) == j) { if (firsth < 0) firsth = alt & 3; if ((alt & 3) == h) found = 1; } }
          if (!found && firsth >= 0) h = firsth;
          dec[k->off_mode + c * N + i] = (unsigned char)(j * 4 + h);
        }
      }
      for (int pt = 0; pt < 5; ++pt) {
        if (E > 1) { int e = 0; for (int ee = 0; ee < E; ++ee) if (x[l->base_nwe + ((pt * C + c) * E + ee) * N + i] < 0.5) { e = ee; break; } dec[k->off_env + (c * N + i) * 5 + pt] = (unsigned char)e; }
        for (int o = 0; o < O; ++o) {
          int ne = p->obs_nedges[o * N + i]; int d = OBS_SOFT;
          for (int ed = 0; ed < ne; ++ed) {
            double v = (pt == 0) ? x[l->base_dcc + ((c * O + o) * N + i) * L + ed] : x[l->base_dcf + (((c * O + o) * N + i) * L + ed) * 4 + (4 - pt)];
            if (v < 0.5) { d = ed; break; }
          }
          dec[k->off_obs + ((c * O + o) * N + i) * 5 + pt] = (unsigned char)d;
        }
      }
    }
  for (int a = 0; a < C - 1; ++a)
    for (int b = a + 1; b < C; ++b) {
      int pr = pair_index(k, a, b), k1 = a, k2 = b - 1;
      for (int i = 0; i < N; ++i)
        for (int qd = 0; qd < 4; ++qd) {
          int d = 0; for (int side = 0; side < 4; ++side) if (x[l->base_c2c + ((k1 * K + k2) * N + i) * 16 + qd * 4 + side] < 0.5) { d = side; break; }
          dec[k->off_pair + (pr * N + i) * 4 + qd] = (unsigned char)d;
        }
    }
}

static int count_undecided(const Ctx *k, const unsigned char *dec) {
  int n = 0;
  for (int c = 0; c < k->C; ++c) for (int i = 1; i < k->N; ++i) {
    if (dec[k->off_mode + c * k->N + i] == UNDEC) n++;
    if (k->E > 1) for (int pt = 0; pt < 5; ++pt) if (dec[k->off_env + (c * k->N + i) * 5 + pt] == UNDEC) n++;
    for (int o = 0; o < k->O; ++o) for (int pt = 0; pt < 5; ++pt) if (dec[k->off_obs + ((c * k->O + o) * k->N + i) * 5 + pt] == UNDEC) n++;
  }
  for (int pr = 0; pr < k->P; ++pr) for (int i = 1; i < k->N; ++i) for (int q = 0; q < 4; ++q) if (dec[k->off_pair + (pr * k->N + i) * 4 + q] == UNDEC) n++;
  return n;
}

int orc_solve_fixed(const OrcProblem *p, const double *x_bin, double *x_out, double *objective) {
  Ctx k; ctx_init(&k, p);
  unsigned char *dec = (unsigned char *)malloc((size_t)k.ndec);
  decisions_from_solution(&k, x_bin, dec);
  QP q; qp_init(&q, k.n);
  double pen = 0.0, obj = 0.0;
  int rc = build_node_qp(&k, dec, &q, &pen);
  double *z = (double *)calloc((size_t)k.n, sizeof(double));
  if (!rc) rc = qp_solve(&k, &q, z, &obj, NULL);
  if (rc == QP_FEASIBLE_POINT) rc = 0;
  if (!rc) {
    double *traj = (double *)malloc(sizeof(double) * (size_t)k.C * k.N * 8);
    simulate(&k, z, traj);
    fill_solution(&k, dec, traj, z, x_out);
    *objective = obj + pen;
    free(traj);
  }
  free(z); qp_free(&q); free(dec); ctx_free(&k);
  return rc;
}

int orc_solve(const OrcProblem *p, const double *warm, double *x_out, OrcSolveInfo *info, int verbose) {
  Ctx k; ctx_init(&k, p);
  const double t0 = now_s();
  const double gap_tol = p->gap_tol, tlim = p->time_limit;
  Heap heap; memset(&heap, 0, sizeof heap);
  QP q; qp_init(&q, k.n);
  double *z = (double *)calloc((size_t)k.n, sizeof(double));
  double *traj = (double *)malloc(sizeof(double) * (size_t)k.C * k.N * 8);
  unsigned char *imp = (unsigned char *)malloc((size_t)k.ndec);
  double ub = HUGE_VAL; int have_inc = 0;
  double pruned_lb = HUGE_VAL; /* smallest bound among nodes discarded by the gap rule */
  long nodes = 0, uncertified = 0; /* uncertified: nodes closed without optimum or infeasibility certificate */
  memset(info, 0, sizeof *info);

  Node root; root.bound = -HUGE_VAL; root.depth = 0; root.rank = 0;
  root.dec = (unsigned char *)malloc((size_t)k.ndec); memset(root.dec, UNDEC, (size_t)k.ndec);
  if (warm) { /* MIP start: evaluate the fully decided node first (cplex_wrapper.cpp:494-639) */
    Node w; w.bound = -HUGE_VAL; w.depth = 1 << 20; w.rank = 0; w.dec = (unsigned char *)malloc((size_t)k.ndec);
    decisions_from_solution(&k, warm, w.dec);
    heap_push(&heap, w);
  }
  heap_push(&heap, root);

  int timed_out = 0;
  while (heap.n > 0) {
    if (now_s() - t0 > tlim) { timed_out = 1; break; }
    Node nd;
    if (!have_inc && nodes >= 48 && (nodes & 1) && heap.n > 1) {
      /* A dive that has not produced an incumbent after 48 nodes may sit below a wrong early decision (deepest-first
       * backtracking never leaves that subtree): every second node is then taken from the best bound instead. */
      int bi = 0;
      for (int a = 1; a < heap.n; ++a) if (heap.a[a].bound < heap.a[bi].bound || (heap.a[a].bound == heap.a[bi].bound && heap.a[a].depth > heap.a[bi].depth)) bi = a;
      nd = heap.a[bi]; heap.a[bi] = heap.a[--heap.n]; heap_rebuild(&heap);
    } else nd = heap_pop(&heap);
    double cutoff = have_inc ? ub - gap_tol * fabs(ub) : HUGE_VAL;
    if (nd.bound >= cutoff) { if (nd.bound < pruned_lb) pruned_lb = nd.bound; free(nd.dec); continue; }
    nodes++;
    double pen = 0.0, obj = 0.0, qlb = -HUGE_VAL;
    int qs = QP_INFEASIBLE; /* boxes with lo > hi: infeasible by construction */
    if (!build_node_qp(&k, nd.dec, &q, &pen)) qs = qp_solve(&k, &q, z, &obj, &qlb);
    if (qs == QP_INFEASIBLE) { free(nd.dec); continue; }
    if (qs == QP_UNKNOWN) { /* neither solved nor refuted: the node is closed, its bound stays in the books */
      uncertified++; if (nd.bound < pruned_lb) pruned_lb = nd.bound; free(nd.dec); continue;
    }
    /* fval: objective of the point z (an upper bound of the relaxation); obj: lower bound of the node.  They
     * coincide when the iteration converged; a stalled iteration yields the Lagrangian bound of its multipliers. */
    const double fval = obj + pen;
    obj = (qs == QP_OPTIMAL) ? fval : fmax(nd.bound, qlb + pen);
    if (obj < nd.bound) obj = nd.bound; /* numerical monotonicity */
    if (obj >= cutoff) { if (obj < pruned_lb) pruned_lb = obj; free(nd.dec); continue; }
    simulate(&k, z, traj);
    Branch br; scan(&k, nd.dec, traj, z, imp, &br);
    if (verbose > 1) fprintf(stderr, "node %ld depth %d obj %.6f kind %d c%d i%d pt%d viol %.3g open %d ub %.6f\n", nodes, nd.depth, obj, br.kind, br.c, br.i, br.pt, br.viol, heap.n, ub);
    if (br.kind == 0) {
      int und = count_undecided(&k, nd.dec);
      /* a stalled point that satisfies an alternative of every disjunction is still replaced by its completion alone;
       * the dropped completions keep their bound in the books */
      if (und > 0 && qs != QP_OPTIMAL && obj < pruned_lb) pruned_lb = obj;
      if (und == 0) {
        const double inc = fval > obj ? fval : obj;
        if (qs != QP_OPTIMAL && obj < pruned_lb) pruned_lb = obj; /* the leaf's optimum may lie below the stalled point, but not below obj */
        if (!(inc < ub)) { free(nd.dec); continue; }
        ub = inc; have_inc = 1;
        fill_solution(&k, nd.dec, traj, z, x_out);
        if (verbose) fprintf(stderr, "[oracle] incumbent %.8f after %ld nodes, %.3fs\n", ub, nodes, now_s() - t0);
        if (!heap.have_inc) { heap.have_inc = 1; heap_rebuild(&heap); }
        free(nd.dec);
      } else {
        Node ch; ch.bound = obj; ch.depth = nd.depth + 1; ch.rank = 0; ch.dec = nd.dec; memcpy(ch.dec, imp, (size_t)k.ndec);
        heap_push(&heap, ch);
      }
      if (br.kind == 0) continue;
    }
    /* branch */
    int nalt = 0; unsigned char alts[260]; size_t soff = 0;
    if (br.kind == 1) {
      soff = (size_t)(d_mode(&k, nd.dec, br.c, br.i) - nd.dec);
      alts[nalt++] = MODE_FROZEN;
      for (int a = 0; a < k.nalt_mode[br.c]; ++a) alts[nalt++] = k.alt_mode[br.c * k.R * 4 + a];
    } else if (br.kind == 2) {
      soff = (size_t)(d_env(&k, nd.dec, br.c, br.i, br.pt) - nd.dec);
      for (int e = 0; e < k.E; ++e) alts[nalt++] = (unsigned char)e;
    } else if (br.kind == 3) {
      soff = (size_t)(d_obs(&k, nd.dec, br.c, br.o, br.i, br.pt) - nd.dec);
      int ne = p->obs_nedges[br.o * k.N + br.i];
      for (int e = 0; e < ne; ++e) alts[nalt++] = (unsigned char)e;
      if (p->obs_soft[br.o] == 1) alts[nalt++] = OBS_SOFT;
    } else {
      soff = (size_t)(d_pair(&k, nd.dec, br.pr, br.i, br.q) - nd.dec);
      for (int s = 0; s < 4; ++s) alts[nalt++] = (unsigned char)s;
    }
    for (int a = 0; a < nalt; ++a) {
      Node ch; ch.bound = obj; ch.depth = nd.depth + 1; ch.rank = a;
      ch.dec = (unsigned char *)malloc((size_t)k.ndec); memcpy(ch.dec, nd.dec, (size_t)k.ndec);
      ch.dec[soff] = alts[a];
      if (ch.dec[soff] == imp[soff]) ch.rank = -1; /* the least violated alternative first */
      heap_push(&heap, ch);
    }
    free(nd.dec);
  }
  double lb = pruned_lb;
  for (int a = 0; a < heap.n; ++a) { if (heap.a[a].bound < lb) lb = heap.a[a].bound; free(heap.a[a].dec); }
  if (!timed_out && heap.n == 0 && lb == HUGE_VAL) lb = ub; /* tree exhausted */
  if (have_inc && lb > ub) lb = ub;
  info->nodes = nodes; info->qp_solves = k.qp_solves; info->qp_iters = k.qp_iters; info->uncertified = uncertified;
  info->seconds = now_s() - t0;
  if (have_inc) {
    info->status = 0; info->objective = ub; info->best_bound = lb;
    info->gap = fabs(lb - ub) / (1e-10 + fabs(ub));
    info->proven = info->gap <= gap_tol + 1e-15;
    info->max_violation = orc_max_violation(p, x_out, NULL);
    /* objective re-evaluated on the full vector is authoritative */
    info->objective = orc_objective(p, x_out);
  } else {
    info->status = timed_out ? 3 : 1; info->objective = NAN; info->gap = NAN; info->best_bound = lb;
  }
  free(heap.a); qp_free(&q); free(z); free(traj); free(imp); ctx_free(&k);
  return info->status;
}

double orc_complete_assignment(const OrcProblem *p, double *x) {
  /* derive decisions from the trajectory alone: scan with nothing decided, then fill */
  Ctx k; ctx_init(&k, p);
  const int C = k.C, N = k.N;
  double *traj = (double *)malloc(sizeof(double) * (size_t)C * N * 8);
  double *z = (double *)calloc((size_t)k.n, sizeof(double));
  const int blk[8] = {2, 3, 4, 5, 6, 7, 0, 1};
  for (int c = 0; c < C; ++c) for (int i = 0; i < N; ++i) for (int t = 0; t < 8; ++t) traj[((size_t)c * N + i) * 8 + t] = x[(blk[t] * C + c) * N + i];
  unsigned char *dec = (unsigned char *)malloc((size_t)k.ndec), *imp = (unsigned char *)malloc((size_t)k.ndec);
  memset(dec, UNDEC, (size_t)k.ndec);
  Branch br; scan(&k, dec, traj, z, imp, &br);
  /* pair slacks: cheapest feasible value for the implied sides */
  for (int a = 0; a < C - 1; ++a) for (int b = a + 1; b < C; ++b) { int pr = pair_index(&k, a, b);
    for (int i = 0; i < N; ++i) for (int qd = 0; qd < 4; qd += 3) { SRow r; int jdummy = 0; (void)jdummy;
      int side = imp[k.off_pair + (pr * N + i) * 4 + qd]; int *je = (int *)malloc(sizeof(int) * (size_t)C * N); effective_regions(&k, imp, je);
      pair_row(&k, &r, a, b, i, qd, side, je[a * N + i], je[b * N + i]); free(je);
      double v = srow_eval(&r, traj, z, N); if (v > 0) z[r.slack] = v; } }
  fill_solution(&k, imp, traj, z, x);
  double mv = orc_max_violation(p, x, NULL);
  free(traj); free(z); free(dec); free(imp); ctx_free(&k);
  return mv;
}
