// node_qp_multi.cuh -- one CTA solves one node relaxation of a plan with several cars.
//
// Same relaxation as node_qp.cuh (decided disjunctions contribute their rows without big-M,
// undecided ones nothing), for NumCars >= 1: the stage vector is the concatenation of the
// cars' (px,vx,ax,py,vy,ay | ux,uy), coupled inside a stage by the collision rows of
// agent_collision_constraints.mod:38-73 (one row per decided side of a pair quadruple) and
// their slack variables slackvars[k1,k2,i,1..4] (decision_variables.mod:53, cost
// WEIGHTS_SLACK * s^2, objective_function.mod:17-19, bounds [0, min(slack_max(i),
// maximum_slack)]).
//
//     min  sum_i 1/2 z_i' Q_i z_i + c_i' z_i + w_s sum sigma^2
//     s.t. x_{i+1} = A x_i + B u_i (block diagonal over cars), x_0 given, u_{N-1} = 0
//          stage-local rows of every car (bounds, region mode, environment, obstacles)
//          pair rows  a.x_a + a2.x_b - sigma <= rhs,   0 <= sigma <= cap_i
//
// Solver: Mehrotra predictor-corrector interior point; every Newton step is a Riccati sweep
// over the joint state (6C) / control (2C).  A slack variable belongs to exactly one pair
// row, so it is eliminated in closed form inside that row (series combination of the row
// weight and the slack curvature) and recovered after the sweep.
//
// Parallel structure: the CTA executes a sequence of phases; a phase is a loop over
// independent work items (PFOR) followed by a barrier.  Nothing else synchronises, which is
// why the same source also compiles for the host with one thread (MQ_EMULATE, used by the
// CPU tests of the host logic; never by the product).
#pragma once
#include "dev_problem.cuh"

namespace miqp {

#ifdef MQ_EMULATE
#define MQ_FN static inline
#define MQ_MFN inline
#else
#define MQ_FN __device__ __forceinline__
#define MQ_MFN __device__ __forceinline__
#endif
#define MQM_INF HUGE_VAL

#define PFOR(it, n) for (int it = k.tid; it < (n); it += k.nthr)

MQ_FN double m_rcp(double x) {
#ifdef MQ_EMULATE
  return 1.0 / x;
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  r = fma(r, t, r);
  const double e2 = fma(-x, r, 1.0);
  return fma(r, e2, r);
#endif
}

enum { SG_VAL = 0, SG_D, SG_DA, SG_ALPHA, SG_BETA, SG_SIZE };
constexpr int PROW = 12;  // (s, lambda) records of a pair-stage: 4 quadruple rows, 4 x (hi, lo) slack bounds

struct MCtx {
  const DevProb *p;
  const double *D;
  const int *I;
  int tid, nthr;
  int C, N, nx, nu, nz, nxx, nuu, P, kmaxc;
  double *Z, *DZ, *DZA, *GR;  // [N][8C] point, combined step, affine step, gradient of the Newton system
  double *Mxx, *Muu;       // [N][nxx] packed lower, [N][nu]
  double *Kg, *Finv, *kv;  // [N][nu*nx], [N][nuu] packed lower, [N][nu]
  double *Pb, *pb;         // [2][nxx], [2][nx]   value function of the next stage
  double *Gs, *Fs, *phi;   // [nu*nx], [nu*nu], [nx+nu]
  double2 *rows;           // [C][kmaxc][N]
  double2 *prow;           // [P][N][PROW]
  double *sig;             // [P][N][4][SG_SIZE]
  double *red;             // [32] reduction scratch
  int *jeff;               // [C][N] effective region (-1 unknown)
  int *aux;                // [C][3N] scan tables
  double *auxd;            // [C][N]
  unsigned char *dec, *imp;
#ifdef MQ_EMULATE
  void sync() const {}
  double rmax(double v) const { return v; }
  double rsum(double v) const { return v; }
  int rsumi(int v) const { return v; }
  int rany(int v) const { return v; }
#else
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ double rmax(double v) const {
    for (int o = 16; o > 0; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < (nthr >> 5); ++w) r = red[w] > r ? red[w] : r;
    __syncthreads();
    return r;
  }
  __device__ __forceinline__ double rsum(double v) const {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < (nthr >> 5); ++w) r += red[w];
    __syncthreads();
    return r;
  }
  __device__ __forceinline__ int rsumi(int v) const { return (int)(rsum((double)v) + 0.5); }
  __device__ __forceinline__ int rany(int v) const { return __syncthreads_or(v); }
#endif
};

// doubles of workspace one node needs (the launcher places it in shared memory when it fits,
// in a per-CTA slice of HBM otherwise)
struct MLayout {
  long Z, DZ, DZA, GR, Mxx, Muu, Kg, Finv, kv, Pb, pb, Gs, Fs, phi, rows, prow, sig, red, auxd, ints, bytes, total_bytes;
};
__host__ __device__ inline MLayout multi_layout(int C, int N, int P, int kmaxc, int ndec_stride) {
  const int nx = 6 * C, nu = 2 * C, nz = 8 * C, nxx = nx * (nx + 1) / 2, nuu = nu * (nu + 1) / 2;
  MLayout L; long o = 0;
  L.Z = o; o += (long)N * nz; L.DZ = o; o += (long)N * nz; L.DZA = o; o += (long)N * nz; L.GR = o; o += (long)N * nz;
  L.Mxx = o; o += (long)N * nxx; L.Muu = o; o += (long)N * nu;
  L.Kg = o; o += (long)N * nu * nx; L.Finv = o; o += (long)N * nuu; L.kv = o; o += (long)N * nu;
  L.Pb = o; o += 2L * nxx; L.pb = o; o += 2L * nx;
  L.Gs = o; o += (long)nu * nx; L.Fs = o; o += (long)nu * nu; L.phi = o; o += nx + nu;
  L.red = o; o += 32; L.auxd = o; o += (long)C * N;
  L.sig = o; o += (long)P * N * 4 * SG_SIZE;
  o = (o + 1) & ~1L;
  L.rows = o; o += 2L * C * kmaxc * N;
  L.prow = o; o += 2L * P * N * PROW;
  L.ints = o; o += ((long)C * N * 4 + 1) / 2;
  L.bytes = o;
  L.total_bytes = o * 8 + 2L * ndec_stride + 16;
  return L;
}
MQ_FN void multi_bind(MCtx &k, const DevProb *p, double *ws, int ndec_stride) {
  k.p = p; k.C = p->C; k.N = p->N; k.P = p->P; k.kmaxc = p->kmax;
  k.nx = 6 * k.C; k.nu = 2 * k.C; k.nz = 8 * k.C; k.nxx = k.nx * (k.nx + 1) / 2; k.nuu = k.nu * (k.nu + 1) / 2;
  const MLayout L = multi_layout(k.C, k.N, k.P, k.kmaxc, ndec_stride);
  k.Z = ws + L.Z; k.DZ = ws + L.DZ; k.DZA = ws + L.DZA; k.GR = ws + L.GR; k.Mxx = ws + L.Mxx; k.Muu = ws + L.Muu;
  k.Kg = ws + L.Kg; k.Finv = ws + L.Finv; k.kv = ws + L.kv; k.Pb = ws + L.Pb; k.pb = ws + L.pb;
  k.Gs = ws + L.Gs; k.Fs = ws + L.Fs; k.phi = ws + L.phi; k.red = ws + L.red; k.auxd = ws + L.auxd;
  k.sig = ws + L.sig;
  k.rows = reinterpret_cast<double2 *>(ws + L.rows);
  k.prow = reinterpret_cast<double2 *>(ws + L.prow);
  k.jeff = reinterpret_cast<int *>(ws + L.ints);
  k.aux = k.jeff + k.C * k.N;
  k.dec = reinterpret_cast<unsigned char *>(ws + L.bytes);
  k.imp = k.dec + ndec_stride;
}

MQ_FN int tri(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// ---------------------------------------------------------------------------------------
// rows of one car at one stage (same families and slot order as node_qp.cuh:visit_rows)
// ---------------------------------------------------------------------------------------
MQ_FN void m_stage_bounds(const MCtx &k, int c, int i, int je, bool frozen, double lo[8], double hi[8]) {
  const DevProb &p = *k.p;
#pragma unroll
  for (int t = 0; t < 8; ++t) { lo[t] = -MQM_INF; hi[t] = MQM_INF; }
  lo[Y_VX] = p.min_vel; hi[Y_VX] = p.max_vel; lo[Y_VY] = p.min_vel;  // vel_y has no upper bound
  lo[Y_AX] = p.total_min_acc; hi[Y_AX] = p.total_max_acc; lo[Y_AY] = p.total_min_acc; hi[Y_AY] = p.total_max_acc;
  lo[Y_UX] = p.total_min_jerk; hi[Y_UX] = p.total_max_jerk; lo[Y_UY] = p.total_min_jerk; hi[Y_UY] = p.total_max_jerk;
  if (je >= 0) {
    const double *D = k.D;
    const int q = c * p.R + je;
    lo[Y_UX] = fmax(lo[Y_UX], D[p.o_lim[4] + q]); hi[Y_UX] = fmin(hi[Y_UX], D[p.o_lim[5] + q]);
    lo[Y_UY] = fmax(lo[Y_UY], D[p.o_lim[6] + q]); hi[Y_UY] = fmin(hi[Y_UY], D[p.o_lim[7] + q]);
    if (i > 0) {
      lo[Y_AX] = fmax(lo[Y_AX], D[p.o_lim[0] + q]); hi[Y_AX] = fmin(hi[Y_AX], D[p.o_lim[1] + q]);
      lo[Y_AY] = fmax(lo[Y_AY], D[p.o_lim[2] + q]); hi[Y_AY] = fmin(hi[Y_AY], D[p.o_lim[3] + q]);
    }
  }
  if (frozen) {
    const double vm = p.vm;
    lo[Y_VX] = fmax(lo[Y_VX], -vm); hi[Y_VX] = fmin(hi[Y_VX], vm);
    lo[Y_VY] = fmax(lo[Y_VY], -vm); hi[Y_VY] = fmin(hi[Y_VY], vm);
  }
}

// sign=+1: cross(P)/len <= 0 (obstacle, chosen edge); sign=-1: cross(P)/len >= 0 (environment)
// points: 0 rear, 1 (xU,yU), 2 (xL,yU), 3 (xU,yL), 4 (xL,yL)
MQ_FN void m_edge_row(const double *et, const double *ft, int pt, double sign, double a[6], double &rhs) {
  const double ex = et[0], ey = et[1], ec = et[2];
  double xc = 0.0, fx1 = 0.0, fx2 = 0.0, yc = 0.0, fy1 = 0.0, fy2 = 0.0;
  if (pt > 0) {
    const double *fx = (pt == 1 || pt == 3) ? ft : ft + 3;
    const double *fy = (pt == 1 || pt == 2) ? ft + 6 : ft + 9;
    xc = fx[0]; fx1 = fx[1]; fx2 = fx[2];
    yc = fy[0]; fy1 = fy[1]; fy2 = fy[2];
  }
  a[Y_PX] = -sign * ey; a[Y_VX] = sign * (ex * fy1 - ey * fx1); a[Y_AX] = 0.0;
  a[Y_PY] = sign * ex;  a[Y_VY] = sign * (ex * fy2 - ey * fx2); a[Y_AY] = 0.0;
  rhs = -sign * (ex * yc - ey * xc - ec);
}

MQ_FN void m_mode_row(const MCtx &k, int j, int h, int r, double a[6], double &rhs) {
  const DevProb &p = *k.p;
  if (r < 4) {
    const double *t = k.D + p.o_modetab + 20 * j + 5 * r;
    a[Y_PX] = 0.0; a[Y_VX] = t[0]; a[Y_AX] = t[1]; a[Y_PY] = 0.0; a[Y_VY] = t[2]; a[Y_AY] = t[3];
    rhs = t[4];
  } else {
    a[Y_PX] = 0.0; a[Y_AX] = 0.0; a[Y_PY] = 0.0; a[Y_AY] = 0.0;
    a[Y_VX] = (h == 0) ? -1.0 : (h == 2) ? 1.0 : 0.0;   // h: 0 vx>=vm, 1 vy>=vm, 2 vx<=-vm, 3 vy<=-vm
    a[Y_VY] = (h == 1) ? -1.0 : (h == 3) ? 1.0 : 0.0;
    rhs = -p.vm;
  }
}

// Vis::bound(slot, T, sgn, rhs):  sgn * y[T] <= rhs ;  Vis::general(slot, a[6], rhs):  a.x <= rhs
template <class Vis>
MQ_FN void m_car_rows(const MCtx &k, int c, int i, Vis &v) {
  const DevProb &p = *k.p;
  const int N = k.N;
  const unsigned char m = (i > 0) ? k.dec[p.off_mode + c * N + i] : (unsigned char)0;
  const int je = k.jeff[c * N + i];
  double lo[8], hi[8];
  m_stage_bounds(k, c, i, je, i > 0 && m == MODE_FROZEN, lo, hi);
  const bool st = (i > 0), ut = (i < N - 1);
  int slot = 0;
  const int T6[6] = {Y_VX, Y_AX, Y_VY, Y_AY, Y_UX, Y_UY};
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    const int T = T6[b];
    const bool act = (b < 4) ? st : ut;
    if (act && hi[T] < MQM_INF) v.bound(slot, T, 1.0, hi[T]);
    ++slot;
    if (act && lo[T] > -MQM_INF) v.bound(slot, T, -1.0, -lo[T]);
    ++slot;
  }
  if (i == 0) return;
  double a[6], rhs;
  if (m != UNDEC && m != MODE_FROZEN) {
    const int j = m >> 2, h = m & 3;
#pragma unroll 1
    for (int r = 0; r < 5; ++r) { m_mode_row(k, j, h, r, a, rhs); v.general(slot + r, a, rhs); }
  }
  slot += 5;
  const double *ft = k.D + p.o_fronttab + 12 * (c * p.R + (je >= 0 ? je : 0));
  if (p.E > 0) {
#pragma unroll 1
    for (int pt = 0; pt < 5; ++pt) {
      const int e = (p.E == 1) ? 0 : k.dec[p.off_env + (c * N + i) * 5 + pt];
      const bool act = (e != UNDEC) && (pt == 0 || je >= 0);
      const int e0 = act ? k.I[p.o_env_off + e] : 0;
      const int ne = act ? k.I[p.o_env_off + e + 1] - e0 : 0;
#pragma unroll 1
      for (int ed = 0; ed < p.maxEnvEdges; ++ed) {
        if (ed < ne) { m_edge_row(k.D + p.o_envtab + 3 * (e0 + ed), ft, pt, -1.0, a, rhs); v.general(slot, a, rhs); }
        ++slot;
      }
    }
  }
#pragma unroll 1
  for (int o = 0; o < p.O; ++o)
#pragma unroll 1
    for (int pt = 0; pt < 5; ++pt) {
      const unsigned char d = k.dec[p.off_obs + ((c * p.O + o) * N + i) * 5 + pt];
      if (d != UNDEC && d != OBS_SOFT && (pt == 0 || je >= 0)) {
        m_edge_row(k.D + p.o_obstab + 3 * ((o * N + i) * p.L + d), ft, pt, 1.0, a, rhs);
        v.general(slot, a, rhs);
      }
      ++slot;
    }
}

// ---------------------------------------------------------------------------------------
// pair rows (agent_collision_constraints.mod:38-73)
// quadruple q: 0 rear/rear, 1 rear a / front b, 2 rear b / front a, 3 front/front; side 0..3
// (0,1: x axis; 2,3: y axis; even: "a below b").  Row:  ca.x_a + cb.x_b + as*sigma <= rhs.
// ---------------------------------------------------------------------------------------
struct PairRow { double ca[6], cb[6], rhs; int slack; /* 0..3 or -1 */ };

MQ_FN void m_front_map(const MCtx &k, int c, int i, int j, int axis, bool upper, double f[3]) {
  const DevProb &p = *k.p;
  if (i == 0) {  // initial_conditions.mod:20-23: front box collapsed onto the heading
    f[0] = axis ? (k.D[p.o_front0 + 2 * c + 1] - k.D[p.o_x0 + 6 * c + 3]) : (k.D[p.o_front0 + 2 * c] - k.D[p.o_x0 + 6 * c]);
    f[1] = 0.0; f[2] = 0.0;
    return;
  }
  const double *ft = k.D + p.o_fronttab + 12 * (c * p.R + (j >= 0 ? j : 0)) + (axis ? 6 : 0) + (upper ? 0 : 3);
  f[0] = ft[0]; f[1] = ft[1]; f[2] = ft[2];
}

MQ_FN void m_pair_row(const MCtx &k, int a, int b, int i, int q, int side, int ja, int jb, PairRow &r) {
  const DevProb &p = *k.p;
#pragma unroll
  for (int t = 0; t < 6; ++t) { r.ca[t] = 0.0; r.cb[t] = 0.0; }
  r.slack = -1;
  const double RR = k.D[p.o_radius + a] + k.D[p.o_radius + b];
  const double Dd = RR + k.D[p.o_safety + i], Ds = Dd + k.D[p.o_safety_slack + i];
  const int axis = side >> 1;
  const bool first = (side & 1) == 0;
  const int PP = axis ? Y_PY : Y_PX;
  double cst = 0.0, f[3];
  if (q == 0) {
    r.ca[PP] += first ? 1.0 : -1.0; r.cb[PP] += first ? -1.0 : 1.0;
    cst += Ds; r.slack = axis;
  } else if (q == 1) {   // rear a vs front b
    m_front_map(k, b, i, jb, axis, !first, f);
    const double sg = first ? -1.0 : 1.0;
    r.ca[PP] += first ? 1.0 : -1.0;
    r.cb[PP] += sg; r.cb[Y_VX] += sg * f[1]; r.cb[Y_VY] += sg * f[2]; cst += sg * f[0];
    cst += Dd;
  } else if (q == 2) {   // rear b vs front a
    m_front_map(k, a, i, ja, axis, !first, f);
    const double sg = first ? -1.0 : 1.0;
    r.cb[PP] += first ? 1.0 : -1.0;
    r.ca[PP] += sg; r.ca[Y_VX] += sg * f[1]; r.ca[Y_VY] += sg * f[2]; cst += sg * f[0];
    cst += Dd;
  } else {               // front/front worst case
    if (first) {         // fUB_b - fLB_a + Ds - s <= 0
      m_front_map(k, b, i, jb, axis, true, f);
      r.cb[PP] += 1.0; r.cb[Y_VX] += f[1]; r.cb[Y_VY] += f[2]; cst += f[0];
      m_front_map(k, a, i, ja, axis, false, f);
      r.ca[PP] -= 1.0; r.ca[Y_VX] -= f[1]; r.ca[Y_VY] -= f[2]; cst -= f[0];
    } else {             // fUB_a - fLB_b + Ds - s <= 0
      m_front_map(k, a, i, ja, axis, true, f);
      r.ca[PP] += 1.0; r.ca[Y_VX] += f[1]; r.ca[Y_VY] += f[2]; cst += f[0];
      m_front_map(k, b, i, jb, axis, false, f);
      r.cb[PP] -= 1.0; r.cb[Y_VX] -= f[1]; r.cb[Y_VY] -= f[2]; cst -= f[0];
    }
    cst += Ds; r.slack = 2 + axis;
  }
  r.rhs = -cst;
}

MQ_FN double m_slack_cap(const MCtx &k, int i) {
  const double a = k.D[k.p->o_safety_slack + i], b = k.p->maximum_slack;
  return a < b ? a : b;
}
MQ_FN int m_pair_index(int C, int a, int b) { int idx = 0; for (int x = 0; x < a; ++x) idx += C - 1 - x; return idx + (b - a - 1); }

// is quadruple q of pair (a,b) at stage i enforced in this node, and with which side
MQ_FN bool m_pair_active(const MCtx &k, int pr, int a, int b, int i, int q, int &side, int &ja, int &jb) {
  const unsigned char d = k.dec[k.p->off_pair + (pr * k.N + i) * 4 + q];
  ja = k.jeff[a * k.N + i]; jb = k.jeff[b * k.N + i];
  side = d;
  if (d == UNDEC) return false;
  if ((q == 1 || q == 3) && jb < 0) return false;
  if ((q == 2 || q == 3) && ja < 0) return false;
  return true;
}

MQ_FN double dot6m(const double a[6], const double *y) {
  double v = 0.0;
#pragma unroll
  for (int t = 0; t < 6; ++t) v += a[t] * y[t];
  return v;
}

// ---------------------------------------------------------------------------------------
// interior-point row arithmetic (one inequality row with slack s and multiplier lam)
// ---------------------------------------------------------------------------------------
struct MStep { double alpha, sigmu; bool pending; };

// applies the pending Newton step to (s, lam) and returns them
MQ_FN void ip_update(double2 &v, const MStep &sc, double gz, double gdz, double gda, double rhs) {
  if (!sc.pending) return;
  double s = v.x, lam = v.y;
  const double rp_old = (gz - sc.alpha * gdz) + s - rhs;
  const double inv = m_rcp(s);
  const double dsa = -rp_old - gda;
  const double dla = -lam - (lam * inv) * dsa;
  const double ds = -rp_old - gdz;
  const double rc = s * lam + dsa * dla - sc.sigmu;
  const double dl = -(rc + lam * ds) * inv;
  v.x = s + sc.alpha * ds; v.y = lam + sc.alpha * dl;
}
struct IpStat { double rpn, musum, lmax; int m; };
MQ_FN void ip_weight(const double2 &v, double gz, double rhs, double &wgt, double &wr, IpStat &st) {
  const double rp = gz + v.x - rhs;
  wgt = v.y * m_rcp(v.x); wr = wgt * rp;
  st.rpn = fmax(st.rpn, fabs(rp)); st.musum += v.x * v.y; st.lmax = fmax(st.lmax, v.y); ++st.m;
}
struct IpAff { double rmax, s1, s2; };
MQ_FN void ip_affine(const double2 &v, double gz, double gda, double rhs, IpAff &st) {
  const double s = v.x, lam = v.y;
  const double rp = gz + s - rhs;
  const double dsa = -rp - gda;
  const double t = dsa * m_rcp(s);
  const double dla = -lam - lam * t;
  st.rmax = fmax(st.rmax, fmax(-t, 1.0 + t));
  st.s1 += s * dla + lam * dsa; st.s2 += dsa * dla;
}
MQ_FN double ip_corr(const double2 &v, double gz, double gda, double rhs, double sigmu) {
  const double s = v.x, lam = v.y;
  const double inv = m_rcp(s);
  const double rp = gz + s - rhs;
  const double dsa = -rp - gda;
  const double dla = -lam - (lam * inv) * dsa;
  return (lam * rp - (dsa * dla - sigmu)) * inv;
}
MQ_FN double ip_ratio(const double2 &v, double gz, double gdz, double gda, double rhs, double sigmu) {
  const double s = v.x, lam = v.y;
  const double inv = m_rcp(s);
  const double rp = gz + s - rhs;
  const double dsa = -rp - gda;
  const double dla = -lam - (lam * inv) * dsa;
  const double ds = -rp - gdz;
  const double rc = s * lam + dsa * dla - sigmu;
  const double dl = -(rc + lam * ds) * inv;
  return fmax(-ds * inv, -dl * m_rcp(lam));
}

// ---- visitors over the rows of one car-stage ---------------------------------------------
struct MPassInit {
  double2 *rows; int N, i; const double *y;
  MQ_MFN void put(int slot, double gz, double rhs) { const double sl = rhs - gz; rows[slot * N + i] = make_double2(sl > 1.0 ? sl : 1.0, 1.0); }
  MQ_MFN void bound(int slot, int T, double sgn, double rhs) { put(slot, sgn * y[T], rhs); }
  MQ_MFN void general(int slot, const double a[6], double rhs) { put(slot, dot6m(a, y), rhs); }
};
struct MPassA {  // pending step, residuals, Hessian block and predictor gradient of one car-stage
  double2 *rows; int N, i; MStep sc; const double *y, *d, *da;
  double H[21], Huu[2], gx[8], gl[8]; IpStat st;
  MQ_MFN void bound(int slot, int T, double sgn, double rhs) {
    double2 v = rows[slot * N + i];
    ip_update(v, sc, sgn * y[T], sgn * d[T], sgn * da[T], rhs);
    if (sc.pending) rows[slot * N + i] = v;
    double wgt, wr; ip_weight(v, sgn * y[T], rhs, wgt, wr, st);
    if (T < 6) H[T * (T + 1) / 2 + T] += wgt; else Huu[T - 6] += wgt;
    gx[T] += sgn * wr; gl[T] += sgn * v.y;
  }
  MQ_MFN void general(int slot, const double a[6], double rhs) {
    double2 v = rows[slot * N + i];
    const double gz = dot6m(a, y);
    ip_update(v, sc, gz, dot6m(a, d), dot6m(a, da), rhs);
    if (sc.pending) rows[slot * N + i] = v;
    double wgt, wr; ip_weight(v, gz, rhs, wgt, wr, st);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double wa = wgt * a[r];
#pragma unroll
      for (int c = 0; c <= r; ++c) H[r * (r + 1) / 2 + c] += wa * a[c];
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) { gx[t] += a[t] * wr; gl[t] += a[t] * v.y; }
  }
};
struct MPassD {
  double2 *rows; int N, i; const double *y, *da; IpAff st;
  MQ_MFN void bound(int slot, int T, double sgn, double rhs) { ip_affine(rows[slot * N + i], sgn * y[T], sgn * da[T], rhs, st); }
  MQ_MFN void general(int slot, const double a[6], double rhs) { ip_affine(rows[slot * N + i], dot6m(a, y), dot6m(a, da), rhs, st); }
};
struct MPassE {
  double2 *rows; int N, i; double sigmu; const double *y, *da; double gx[8];
  MQ_MFN void bound(int slot, int T, double sgn, double rhs) { gx[T] += sgn * ip_corr(rows[slot * N + i], sgn * y[T], sgn * da[T], rhs, sigmu); }
  MQ_MFN void general(int slot, const double a[6], double rhs) {
    const double cf = ip_corr(rows[slot * N + i], dot6m(a, y), dot6m(a, da), rhs, sigmu);
#pragma unroll
    for (int t = 0; t < 6; ++t) gx[t] += a[t] * cf;
  }
};
struct MPassG {
  double2 *rows; int N, i; double sigmu; const double *y, *d, *da; double rmax;
  MQ_MFN void bound(int slot, int T, double sgn, double rhs) { rmax = fmax(rmax, ip_ratio(rows[slot * N + i], sgn * y[T], sgn * d[T], sgn * da[T], rhs, sigmu)); }
  MQ_MFN void general(int slot, const double a[6], double rhs) { rmax = fmax(rmax, ip_ratio(rows[slot * N + i], dot6m(a, y), dot6m(a, d), dot6m(a, da), rhs, sigmu)); }
};
struct MPassViol {
  double2 *rows; int N, i; const double *y; double worst, lmax;
  MQ_MFN void bound(int slot, int T, double sgn, double rhs) { worst = fmax(worst, sgn * y[T] - rhs); lmax = fmax(lmax, rows[slot * N + i].y); }
  MQ_MFN void general(int slot, const double a[6], double rhs) { worst = fmax(worst, dot6m(a, y) - rhs); lmax = fmax(lmax, rows[slot * N + i].y); }
};
struct MDualAcc { double hl, lgz, mag, fz, usum; };   // h'lambda, lambda'Gz, sum lambda |h|, f(z), sum |gradient| x range of the boxed variables
struct MPassDual {  // G'lambda of one car-stage (multipliers as stored, scaled; any lambda >= 0 gives a valid bound / certificate)
  double2 *rows; int N, i; double scale; const double *y; double gl[8]; MDualAcc *acc;
  MQ_MFN double lam_of(int slot, double gz, double rhs) {
    double lam = rows[slot * N + i].y;
    lam = (lam > 0.0 ? lam : 0.0) * scale;
    acc->hl += lam * rhs; acc->lgz += lam * gz; acc->mag += lam * fabs(rhs);
    return lam;
  }
  MQ_MFN void bound(int slot, int T, double sgn, double rhs) { gl[T] += sgn * lam_of(slot, sgn * y[T], rhs); }
  MQ_MFN void general(int slot, const double a[6], double rhs) {
    const double lam = lam_of(slot, dot6m(a, y), rhs);
#pragma unroll
    for (int t = 0; t < 6; ++t) gl[t] += a[t] * lam;
  }
};

// ---------------------------------------------------------------------------------------
// pair-stage passes.  mode: 0 init, 1 pass A (Hessian), 2 recover affine slack step + pass D,
// 3 pass E (corrector gradient), 4 recover combined slack step + pass G, 5 violation (+ largest multiplier),
// 6 dual information (sigmu = scale of the multipliers, first_iter = Farkas mode: cost gradient left out)
// ---------------------------------------------------------------------------------------
struct PairAcc { IpStat st; IpAff af; double rmax, worst, rd0; MDualAcc du; };

template <int MODE>
MQ_FN void m_pair_stage(const MCtx &k, int i, const MStep &sc, double sigmu, bool first_iter, PairAcc &acc) {
  const DevProb &p = *k.p;
  const int C = k.C, N = k.N, nz = k.nz;
  const double qs = 2.0 * p.w_slack;
  const double cap = m_slack_cap(k, i);
  const bool slack_on = cap > 1e-12;
  double *Mi = k.Mxx + (long)i * k.nxx;
  double *gi = k.GR + (long)i * nz;      // gradient buffer (pass A / E)
  const double *zi = k.Z + (long)i * nz, *di = k.DZ + (long)i * nz, *dai = k.DZA + (long)i * nz;
  int pr = 0;
  for (int a = 0; a < C - 1; ++a)
    for (int b = a + 1; b < C; ++b, ++pr) {
      if (MODE == 1) {  // cross block (b,a) starts from zero in every Hessian pass
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Mi[tri(6 * b + r, 6 * a + c)] = 0.0;
      }
      double2 *pw = k.prow + ((long)pr * N + i) * PROW;
      double *sg = k.sig + ((long)pr * N + i) * 4 * SG_SIZE;
      const double *ya = zi + 8 * a, *yb = zi + 8 * b;
      for (int q = 0; q < 4; ++q) {
        int side, ja, jb;
        if (!m_pair_active(k, pr, a, b, i, q, side, ja, jb)) continue;
        PairRow r; m_pair_row(k, a, b, i, q, side, ja, jb, r);
        const int sk = slack_on ? r.slack : -1;
        double *S = (sk >= 0) ? sg + sk * SG_SIZE : nullptr;
        const double sv = S ? S[SG_VAL] : 0.0;
        const double gz = dot6m(r.ca, ya) + dot6m(r.cb, yb) - sv;
        if (MODE == 0) {
          if (S) {  // start the slack in the interior of [0, cap]
            S[SG_VAL] = 0.5 * cap; S[SG_D] = 0.0; S[SG_DA] = 0.0; S[SG_ALPHA] = 0.0; S[SG_BETA] = 0.0;
            const double g0 = gz - 0.5 * cap, sl = r.rhs - g0;
            pw[q] = make_double2(sl > 1.0 ? sl : 1.0, 1.0);
            pw[4 + 2 * sk] = make_double2(0.5 * cap, 1.0);      // sigma <= cap
            pw[5 + 2 * sk] = make_double2(0.5 * cap, 1.0);      // -sigma <= 0
          } else {
            const double sl = r.rhs - gz;
            pw[q] = make_double2(sl > 1.0 ? sl : 1.0, 1.0);
          }
          continue;
        }
        if (MODE == 5) {
          acc.worst = fmax(acc.worst, gz - r.rhs);
          acc.st.lmax = fmax(acc.st.lmax, pw[q].y);
          if (S) { acc.worst = fmax(acc.worst, sv - cap); acc.worst = fmax(acc.worst, -sv); acc.st.lmax = fmax(acc.st.lmax, fmax(pw[4 + 2 * sk].y, pw[5 + 2 * sk].y)); }
          continue;
        }
        if (MODE == 6) {
          const double scale = sigmu; const bool farkas = first_iter;
          const double lam = fmax(pw[q].y, 0.0) * scale;
          acc.du.hl += lam * r.rhs; acc.du.lgz += lam * gz; acc.du.mag += lam * fabs(r.rhs);
          double *gd = k.DZ + (long)i * nz;
          for (int t = 0; t < 6; ++t) { gd[8 * a + t] += r.ca[t] * lam; gd[8 * b + t] += r.cb[t] * lam; }
          if (S) {   // the slack is a boxed variable of its own: gradient (cost) qs sigma - lam + lam_hi - lam_lo, range within [0, cap]
            const double lh = fmax(pw[4 + 2 * sk].y, 0.0) * scale, ll = fmax(pw[5 + 2 * sk].y, 0.0) * scale;
            acc.du.hl += lh * cap; acc.du.lgz += lh * sv - ll * sv; acc.du.mag += lh * cap;
            const double gs = (farkas ? 0.0 : qs * sv) - lam + lh - ll;
            const double range = farkas ? cap : fmax(fmax(cap - sv, sv), 0.0);
            acc.du.usum += fabs(gs) * range;
            if (!farkas) acc.du.fz += p.w_slack * sv * sv;
          }
          continue;
        }
        const double dxa = dot6m(r.ca, dai + 8 * a) + dot6m(r.cb, dai + 8 * b);   // coefficient part of g.dza
        if (MODE == 1) {
          // the pending step of this row needs g.dz of the PREVIOUS step: x part from DZ (still the
          // step here: the gradient is written after all row passes) and the slack part from SG_D
          const double dxz = dot6m(r.ca, di + 8 * a) + dot6m(r.cb, di + 8 * b);
          const double dsg = S ? S[SG_D] : 0.0, dsga = S ? S[SG_DA] : 0.0;
          double2 v = pw[q];
          ip_update(v, sc, gz, dxz - dsg, dxa - dsga, r.rhs);
          double wr_, w_; ip_weight(v, gz, r.rhs, w_, wr_, acc.st);
          double weff = w_, cf = wr_;
          if (S) {
            double2 vh = pw[4 + 2 * sk], vl = pw[5 + 2 * sk];
            ip_update(vh, sc, sv, dsg, dsga, cap);
            ip_update(vl, sc, -sv, -dsg, -dsga, 0.0);
            double wh, whr, wl, wlr;
            ip_weight(vh, sv, cap, wh, whr, acc.st);
            ip_weight(vl, -sv, 0.0, wl, wlr, acc.st);
            const double dsum = qs + wh + wl + w_;
            const double idd = m_rcp(dsum);
            const double gsig = qs * sv + whr - wlr - wr_;
            const double beta = w_ * idd;
            S[SG_ALPHA] = -gsig * idd; S[SG_BETA] = beta;
            weff = w_ - w_ * beta; cf = wr_ + beta * gsig;
            if (sc.pending) { pw[4 + 2 * sk] = vh; pw[5 + 2 * sk] = vl; }
            if (first_iter) acc.rd0 = fmax(acc.rd0, fabs(qs * sv + vh.y - vl.y - v.y));
          }
          if (sc.pending) pw[q] = v;
          // Hessian: weff * [ca;cb][ca;cb]' into blocks (a,a), (b,b), (b,a); gradient
          for (int r1 = 0; r1 < 6; ++r1) {
            const double wa = weff * r.ca[r1], wb = weff * r.cb[r1];
            if (wa != 0.0) for (int c1 = 0; c1 <= r1; ++c1) Mi[tri(6 * a + r1, 6 * a + c1)] += wa * r.ca[c1];
            if (wb != 0.0) {
              for (int c1 = 0; c1 <= r1; ++c1) Mi[tri(6 * b + r1, 6 * b + c1)] += wb * r.cb[c1];
              for (int c1 = 0; c1 < 6; ++c1) Mi[tri(6 * b + r1, 6 * a + c1)] += wb * r.ca[c1];
            }
          }
          for (int t = 0; t < 6; ++t) { gi[8 * a + t] += r.ca[t] * cf; gi[8 * b + t] += r.cb[t] * cf; }
          if (first_iter) {  // sum g.lambda of the x part, parked in DZA next to the cars' share
            double *gl = k.DZA + (long)i * nz;
            for (int t = 0; t < 6; ++t) { gl[8 * a + t] += r.ca[t] * v.y; gl[8 * b + t] += r.cb[t] * v.y; }
          }
        } else if (MODE == 2) {
          double gda = dxa;
          if (S) { const double dsga = S[SG_ALPHA] + S[SG_BETA] * dxa; S[SG_DA] = dsga; gda = dxa - dsga;
            ip_affine(pw[4 + 2 * sk], sv, dsga, cap, acc.af); ip_affine(pw[5 + 2 * sk], -sv, -dsga, 0.0, acc.af); }
          ip_affine(pw[q], gz, gda, r.rhs, acc.af);
        } else if (MODE == 3) {
          const double gda = dxa - (S ? S[SG_DA] : 0.0);
          double cf = ip_corr(pw[q], gz, gda, r.rhs, sigmu);
          if (S) {
            const double2 vh = pw[4 + 2 * sk], vl = pw[5 + 2 * sk], v = pw[q];
            const double ch = ip_corr(vh, sv, S[SG_DA], cap, sigmu), cl = ip_corr(vl, -sv, -S[SG_DA], 0.0, sigmu);
            const double w_ = v.y * m_rcp(v.x), wh = vh.y * m_rcp(vh.x), wl = vl.y * m_rcp(vl.x);
            const double idd = m_rcp(qs + wh + wl + w_);
            const double gsig = qs * sv + ch - cl - cf;
            const double beta = w_ * idd;
            S[SG_ALPHA] = -gsig * idd; S[SG_BETA] = beta;
            cf += beta * gsig;
          }
          for (int t = 0; t < 6; ++t) { gi[8 * a + t] += r.ca[t] * cf; gi[8 * b + t] += r.cb[t] * cf; }
        } else if (MODE == 4) {
          const double dxz = dot6m(r.ca, di + 8 * a) + dot6m(r.cb, di + 8 * b);
          double gdz = dxz, gda = dxa;
          if (S) {
            const double dsg = S[SG_ALPHA] + S[SG_BETA] * dxz; S[SG_D] = dsg;
            gdz = dxz - dsg; gda = dxa - S[SG_DA];
            acc.rmax = fmax(acc.rmax, ip_ratio(pw[4 + 2 * sk], sv, dsg, S[SG_DA], cap, sigmu));
            acc.rmax = fmax(acc.rmax, ip_ratio(pw[5 + 2 * sk], -sv, -dsg, -S[SG_DA], 0.0, sigmu));
          }
          acc.rmax = fmax(acc.rmax, ip_ratio(pw[q], gz, gdz, gda, r.rhs, sigmu));
        }
      }
    }
}

// ---------------------------------------------------------------------------------------
// Riccati sweeps over the joint state
// ---------------------------------------------------------------------------------------
// column `idx` of [A B] (nx x (nx+nu)): rows row0 .. row0+cnt-1 with coefficients cf[]
MQ_FN void m_ab_column(const DevProb &p, int nx, int idx, int &row0, int &cnt, double cf[3]) {
  if (idx < nx) {
    const int o = idx % 3;
    row0 = idx - o; cnt = o + 1;
    cf[0] = (o == 0) ? 1.0 : (o == 1) ? p.ts : p.c2;
    cf[1] = (o == 1) ? 1.0 : p.ts;
    cf[2] = 1.0;
  } else {
    row0 = (idx - nx) * 3; cnt = 3;   // control (c, axis) drives state block 6c + 3 axis = 3 * (2c + axis)
    cf[0] = p.c3; cf[1] = p.c2; cf[2] = p.ts;
  }
}

// backward factorisation with the gradient in GR (x part: [8c..8c+5], u part: [8c+6, 8c+7])
MQ_FN void m_riccati_factor(const MCtx &k) {
  const DevProb &p = *k.p;
  const int N = k.N, nx = k.nx, nu = k.nu, nxx = k.nxx, nuu = k.nuu, nz = k.nz, nt = nx + nu;
  {  // terminal stage: P = Mxx, p = g_x
    double *Pc = k.Pb + ((N - 1) & 1) * nxx, *pc = k.pb + ((N - 1) & 1) * nx;
    const double *Mn = k.Mxx + (long)(N - 1) * nxx, *gn = k.GR + (long)(N - 1) * nz;
    PFOR(e, nxx) Pc[e] = Mn[e];
    PFOR(t, nx) pc[t] = gn[8 * (t / 6) + t % 6];
  }
  k.sync();
  for (int i = N - 2; i >= 0; --i) {
    const double *Pn = k.Pb + ((i + 1) & 1) * nxx, *pn = k.pb + ((i + 1) & 1) * nx;
    double *Pc = k.Pb + (i & 1) * nxx, *pc = k.pb + (i & 1) * nx;
    const double *Mi = k.Mxx + (long)i * nxx, *Mu = k.Muu + (long)i * nu, *gi = k.GR + (long)i * nz;
    double *Ki = k.Kg + (long)i * nu * nx, *Fi = k.Finv + (long)i * nuu, *kvi = k.kv + (long)i * nu;
    // phase 1: Phi = M + [A B]' P [A B]; phi = g + [A B]' p
    const int nent = nt * (nt + 1) / 2;
    PFOR(e, nent + nt) {
      if (e < nent) {
        int a = 0; while ((a + 1) * (a + 2) / 2 <= e) ++a;
        const int b = e - a * (a + 1) / 2;   // a >= b
        int ra, ca, rb, cb; double fa[3], fb[3];
        m_ab_column(p, nx, a, ra, ca, fa); m_ab_column(p, nx, b, rb, cb, fb);
        double v = 0.0;
        for (int x = 0; x < ca; ++x) for (int y = 0; y < cb; ++y) v += fa[x] * fb[y] * Pn[tri(ra + x, rb + y)];
        if (a < nx) Pc[tri(a, b)] = v + Mi[tri(a, b)];            // Phi_xx (into the P buffer of this stage)
        else if (b < nx) k.Gs[(a - nx) * nx + b] = v;                              // Phi_ux (M_ux = 0)
        else { const double m = (a == b) ? Mu[a - nx] : 0.0; k.Fs[(a - nx) * nu + (b - nx)] = v + m; k.Fs[(b - nx) * nu + (a - nx)] = v + m; }
      } else {
        const int a = e - nent;
        int ra, ca; double fa[3];
        m_ab_column(p, nx, a, ra, ca, fa);
        double v = (a < nx) ? gi[8 * (a / 6) + a % 6] : gi[8 * ((a - nx) / 2) + 6 + (a - nx) % 2];
        for (int x = 0; x < ca; ++x) v += fa[x] * pn[ra + x];
        k.phi[a] = v;
      }
    }
    k.sync();
    // phase 2: Finv = Phi_uu^-1 (Cholesky, one thread; nu <= 16)
    if (k.tid == 0) {
      double *F = k.Fs;
      for (int j = 0; j < nu; ++j) {
        double d = F[j * nu + j];
        for (int x = 0; x < j; ++x) d -= F[j * nu + x] * F[j * nu + x];
        d = sqrt(d);
        F[j * nu + j] = d;
        const double id = 1.0 / d;
        for (int r = j + 1; r < nu; ++r) {
          double s = F[r * nu + j];
          for (int x = 0; x < j; ++x) s -= F[r * nu + x] * F[j * nu + x];
          F[r * nu + j] = s * id;
        }
      }
      // inverse of L (lower) in place, then Finv = L^-T L^-1
      for (int j = 0; j < nu; ++j) {
        F[j * nu + j] = 1.0 / F[j * nu + j];
        for (int r = j + 1; r < nu; ++r) {
          double s = 0.0;
          for (int x = j; x < r; ++x) s -= F[r * nu + x] * F[x * nu + j];
          F[r * nu + j] = s / F[r * nu + r];
        }
      }
      for (int r = 0; r < nu; ++r)
        for (int c = 0; c <= r; ++c) {
          double s = 0.0;
          for (int x = r; x < nu; ++x) s += F[x * nu + r] * F[x * nu + c];
          Fi[tri(r, c)] = s;
        }
    }
    k.sync();
    // phase 3: K = Finv Phi_ux ; k = -Finv phi_u
    PFOR(e, nu * nx + nu) {
      if (e < nu * nx) {
        const int u = e / nx, x = e % nx;
        double v = 0.0;
        for (int y = 0; y < nu; ++y) v += Fi[tri(u, y)] * k.Gs[y * nx + x];
        Ki[e] = v;
      } else {
        const int u = e - nu * nx;
        double v = 0.0;
        for (int y = 0; y < nu; ++y) v += Fi[tri(u, y)] * k.phi[nx + y];
        kvi[u] = -v;
      }
    }
    k.sync();
    // phase 4: P = Phi_xx - Phi_ux' K ; p = phi_x + Phi_ux' k
    PFOR(e, nxx + nx) {
      if (e < nxx) {
        int a = 0; while ((a + 1) * (a + 2) / 2 <= e) ++a;
        const int b = e - a * (a + 1) / 2;
        double v = Pc[e];
        for (int u = 0; u < nu; ++u) v -= k.Gs[u * nx + a] * Ki[u * nx + b];
        Pc[e] = v;
      } else {
        const int a = e - nxx;
        double v = k.phi[a];
        for (int u = 0; u < nu; ++u) v += k.Gs[u * nx + a] * kvi[u];
        pc[a] = v;
      }
    }
    k.sync();
  }
}

// vector-only backward sweep with a new gradient in GR
MQ_FN void m_riccati_vector(const MCtx &k) {
  const DevProb &p = *k.p;
  const int N = k.N, nx = k.nx, nu = k.nu, nz = k.nz, nt = nx + nu;
  {
    double *pc = k.pb + ((N - 1) & 1) * nx;
    const double *gn = k.GR + (long)(N - 1) * nz;
    PFOR(t, nx) pc[t] = gn[8 * (t / 6) + t % 6];
  }
  k.sync();
  for (int i = N - 2; i >= 0; --i) {
    const double *pn = k.pb + ((i + 1) & 1) * nx;
    double *pc = k.pb + (i & 1) * nx;
    const double *gi = k.GR + (long)i * nz;
    const double *Ki = k.Kg + (long)i * nu * nx, *Fi = k.Finv + (long)i * k.nuu;
    double *kvi = k.kv + (long)i * nu;
    PFOR(a, nt) {
      int ra, ca; double fa[3];
      m_ab_column(p, nx, a, ra, ca, fa);
      double v = (a < nx) ? gi[8 * (a / 6) + a % 6] : gi[8 * ((a - nx) / 2) + 6 + (a - nx) % 2];
      for (int x = 0; x < ca; ++x) v += fa[x] * pn[ra + x];
      k.phi[a] = v;
    }
    k.sync();
    // p = phi_x - K' phi_u ; k = -Finv phi_u
    PFOR(e, nx + nu) {
      if (e < nx) {
        double v = k.phi[e];
        for (int u = 0; u < nu; ++u) v -= Ki[u * nx + e] * k.phi[nx + u];
        pc[e] = v;
      } else {
        const int u = e - nx;
        double v = 0.0;
        for (int y = 0; y < nu; ++y) v += Fi[tri(u, y)] * k.phi[nx + y];
        kvi[u] = -v;
      }
    }
    k.sync();
  }
}

// forward sweep: writes the step of every stage into dst (DZ or DZA)
MQ_FN void m_riccati_forward(const MCtx &k, double *dst) {
  const DevProb &p = *k.p;
  const int N = k.N, nx = k.nx, nu = k.nu, nz = k.nz;
  PFOR(t, nx) dst[8 * (t / 6) + t % 6] = 0.0;
  k.sync();
  for (int i = 0; i < N; ++i) {
    double *di = dst + (long)i * nz;
    const double *Ki = k.Kg + (long)i * nu * nx, *kvi = k.kv + (long)i * nu;
    PFOR(u, nu) {
      double v = 0.0;
      if (i < N - 1) {
        v = kvi[u];
        for (int x = 0; x < nx; ++x) v -= Ki[u * nx + x] * di[8 * (x / 6) + x % 6];
      }
      di[8 * (u / 2) + 6 + (u % 2)] = v;
    }
    k.sync();
    if (i < N - 1) {
      double *dn = dst + (long)(i + 1) * nz;
      PFOR(t, nx) {
        const int c = t / 6, ax = (t % 6) / 3, o = t % 3;
        const double *y = di + 8 * c + 3 * ax;
        const double U = di[8 * c + 6 + ax];
        double v;
        if (o == 0) v = y[0] + p.ts * y[1] + p.c2 * y[2] + p.c3 * U;
        else if (o == 1) v = y[1] + p.ts * y[2] + p.c2 * U;
        else v = y[2] + p.ts * U;
        dn[8 * c + 3 * ax + o] = v;
      }
      k.sync();
    }
  }
}

// status: 0 optimal (or, with converged == 0, a feasible point), 1 proven infeasible (empty box or Farkas certificate),
// 4 unknown (no convergence, no feasible point, no certificate).  lb: valid lower bound of the relaxation (obj if converged,
// else the Lagrangian bound of the last multipliers).
struct MQpResult { int status, iters, converged; double obj, lb; long rows; };

// Dual information of the current iterate; same mathematics as dual_check of node_qp.cuh (costate recursion per car, the
// pair rows couple the cars only through the stage gradients; the pair slacks are boxed variables).  Uses DZ as scratch (dead between pass A and the second forward sweep, and after the last iteration).
MQ_FN bool m_dual_check(const MCtx &k, bool farkas, double lmax, double *lb) {
  const DevProb &p = *k.p;
  const double *D = k.D;
  const int N = k.N, C = k.C, nz = k.nz, P = k.P;
  const double scale = farkas ? 1.0 / lmax : 1.0;
  PairAcc acc; acc.du.hl = 0.0; acc.du.lgz = 0.0; acc.du.mag = 0.0; acc.du.fz = 0.0; acc.du.usum = 0.0;
  PFOR(itc, C * N) {
    const int c = itc / N, i = itc % N;
    MPassDual v; v.rows = k.rows + (long)c * k.kmaxc * N; v.N = N; v.i = i; v.scale = scale; v.y = k.Z + (long)i * nz + 8 * c; v.acc = &acc.du;
    for (int t = 0; t < 8; ++t) v.gl[t] = 0.0;
    m_car_rows(k, c, i, v);
    const double *cst = D + p.o_cost + 16 * (c * N + i);
    double *gi = k.DZ + (long)i * nz + 8 * c;
    for (int t = 0; t < 8; ++t) {
      gi[t] = v.gl[t] + (farkas ? 0.0 : cst[t] * v.y[t] + cst[8 + t]);
      if (!farkas) acc.du.fz += (0.5 * cst[t] * v.y[t] + cst[8 + t]) * v.y[t];
    }
  }
  k.sync();
  if (P > 0) { MStep s0; s0.alpha = 0.0; s0.sigmu = 0.0; s0.pending = false; PFOR(i, N) m_pair_stage<6>(k, i, s0, scale, farkas, acc); k.sync(); }
  // costate recursion of every car (the dynamics are block diagonal over the cars)
  double px = 0.0;
  PFOR(c, C) {
    const double ts = p.ts, c2 = p.c2, c3 = p.c3;
    const double Ulo = p.total_min_jerk, Uhi = p.total_max_jerk, Uabs = fmax(fabs(Ulo), fabs(Uhi));
    double pn[6];
    for (int t = 0; t < 6; ++t) pn[t] = k.DZ[(long)(N - 1) * nz + 8 * c + t];
    for (int i = N - 2; i >= 0; --i) {
      const double *g = k.DZ + (long)i * nz + 8 * c, *z = k.Z + (long)i * nz + 8 * c;
      for (int ax = 0; ax < 2; ++ax) {
        const double pp = pn[3 * ax], pv = pn[3 * ax + 1], pa = pn[3 * ax + 2];
        pn[3 * ax] = g[3 * ax] + pp;
        pn[3 * ax + 1] = g[3 * ax + 1] + ts * pp + pv;
        pn[3 * ax + 2] = g[3 * ax + 2] + c2 * pp + ts * pv + pa;
        const double rho = g[6 + ax] + c3 * pp + c2 * pv + ts * pa;
        const double range = farkas ? Uabs : fmax(fmax(Uhi - z[6 + ax], z[6 + ax] - Ulo), 0.0);
        acc.du.usum += fabs(rho) * range;
      }
    }
    for (int t = 0; t < 6; ++t) px += pn[t] * D[p.o_x0 + 6 * c + t];
  }
  const double hl = k.rsum(acc.du.hl), lgz = k.rsum(acc.du.lgz), mag = k.rsum(acc.du.mag), fz = k.rsum(acc.du.fz), usum = k.rsum(acc.du.usum);
  const double pi0x0 = k.rsum(px);
  if (farkas) return (hl - pi0x0 + usum) < -1e-10 * (mag + fabs(pi0x0) + usum) - 1e-13;
  *lb = fz + p.cost_const - (hl - lgz) - usum - 1e-12 * (fabs(fz) + fabs(hl) + fabs(lgz) + usum);
  return false;
}

// Solves the node QP of k.dec (k.jeff filled).  On success Z holds the optimal stage vectors
// and sig[..][SG_VAL] the slacks.
MQ_FN MQpResult m_solve_node_qp(const MCtx &k) {
  const DevProb &p = *k.p;
  const double *D = k.D;
  const int N = k.N, C = k.C, nz = k.nz, P = k.P, nxx = k.nxx;
  MQpResult res; res.status = 1; res.iters = 0; res.converged = 0; res.obj = 0.0; res.lb = -MQM_INF; res.rows = 0;

  // trivially infeasible boxes
  int bad = 0;
  PFOR(it, C * N) {
    const int c = it / N, i = it % N;
    const unsigned char m = (i > 0) ? k.dec[p.off_mode + c * N + i] : (unsigned char)0;
    double lo[8], hi[8];
    m_stage_bounds(k, c, i, k.jeff[c * N + i], i > 0 && m == MODE_FROZEN, lo, hi);
    for (int t = 1; t < 8; ++t) {
      if (t == Y_PY) continue;
      if (i == 0 && t < 6) continue;
      if (lo[t] > hi[t] + 1e-12) bad = 1;
      if (i == N - 1 && t >= 6 && (lo[t] > 1e-9 || hi[t] < -1e-9)) bad = 1;
    }
  }
  if (k.rany(bad)) return res;

  // start: zero jerk (free response); slacks of inactive pair rows stay fixed at 0
  PFOR(e, P * N * 4 * SG_SIZE) k.sig[e] = 0.0;
  double cn = 0.0;
  PFOR(it, C * N) {
    const int c = it / N, i = it % N;
    double *zi = k.Z + (long)i * nz + 8 * c;
    const double *x0 = D + p.o_x0 + 6 * c;
    const double t = i * p.ts;
    for (int ax = 0; ax < 2; ++ax) {
      const double Pp = x0[3 * ax], Vv = x0[3 * ax + 1], A = x0[3 * ax + 2];
      zi[3 * ax] = Pp + t * Vv + 0.5 * t * t * A;
      zi[3 * ax + 1] = Vv + t * A;
      zi[3 * ax + 2] = A;
    }
    zi[6] = 0.0; zi[7] = 0.0;
    for (int t8 = 0; t8 < 8; ++t8) { k.DZ[(long)i * nz + 8 * c + t8] = 0.0; k.DZA[(long)i * nz + 8 * c + t8] = 0.0; }
    const double *cst = D + p.o_cost + 16 * (c * N + i);
    for (int t8 = 0; t8 < 8; ++t8) cn = fmax(cn, fabs(cst[8 + t8]));
  }
  k.sync();
  PFOR(it, C * N) {
    const int c = it / N, i = it % N;
    MPassInit v; v.rows = k.rows + (long)c * k.kmaxc * N; v.N = N; v.i = i; v.y = k.Z + (long)i * nz + 8 * c;
    m_car_rows(k, c, i, v);
  }
  MStep sc; sc.alpha = 0.0; sc.sigmu = 0.0; sc.pending = false;
  if (P > 0) {
    PFOR(i, N) { PairAcc acc; m_pair_stage<0>(k, i, sc, 0.0, false, acc); }
  }
  cn = k.rmax(cn);   // (contains barriers)

  int status = 2, stall = 0, it = 0;
  double rdn = 0.0;
  for (it = 0; it < 100; ++it) {
    const bool first = (it == 0);
    // ---- pass A: cars write their diagonal Hessian block, Muu and gradient ------------------
    PairAcc acc;
    acc.st.rpn = 0.0; acc.st.musum = 0.0; acc.st.lmax = 0.0; acc.st.m = 0; acc.rd0 = 0.0;
    PFOR(itc, C * N) {
      const int c = itc / N, i = itc % N;
      MPassA v; v.rows = k.rows + (long)c * k.kmaxc * N; v.N = N; v.i = i; v.sc = sc;
      v.y = k.Z + (long)i * nz + 8 * c; v.d = k.DZ + (long)i * nz + 8 * c; v.da = k.DZA + (long)i * nz + 8 * c;
      v.st.rpn = 0.0; v.st.musum = 0.0; v.st.lmax = 0.0; v.st.m = 0;
      for (int t = 0; t < 21; ++t) v.H[t] = 0.0;
      v.Huu[0] = v.Huu[1] = 0.0;
      for (int t = 0; t < 8; ++t) { v.gx[t] = 0.0; v.gl[t] = 0.0; }
      m_car_rows(k, c, i, v);
      const double *cst = D + p.o_cost + 16 * (c * N + i);
      double *Mi = k.Mxx + (long)i * nxx;
      for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc <= r; ++cc) Mi[tri(6 * c + r, 6 * c + cc)] = v.H[r * (r + 1) / 2 + cc] + (r == cc ? cst[r] : 0.0);
      k.Muu[(long)i * k.nu + 2 * c] = v.Huu[0] + cst[6] + 1e-10;
      k.Muu[(long)i * k.nu + 2 * c + 1] = v.Huu[1] + cst[7] + 1e-10;
      double *gi = k.GR + (long)i * nz + 8 * c;
      for (int t = 0; t < 8; ++t) gi[t] = cst[t] * v.y[t] + cst[8 + t] + v.gx[t];
      if (first) {  // park q + sum g.lambda of the car rows in DZA (dead until the first forward sweep)
        double *gl = k.DZA + (long)i * nz + 8 * c;
        for (int t = 0; t < 8; ++t) gl[t] = cst[t] * v.y[t] + cst[8 + t] + v.gl[t];
      }
      acc.st.rpn = fmax(acc.st.rpn, v.st.rpn); acc.st.musum += v.st.musum; acc.st.lmax = fmax(acc.st.lmax, v.st.lmax); acc.st.m += v.st.m;
    }
    k.sync();
    if (P > 0) {
      PFOR(i, N) m_pair_stage<1>(k, i, sc, 0.0, first, acc);
      k.sync();
    }
    if (first) {
      double rd0 = acc.rd0;
      PFOR(itc, C * N) {
        const int c = itc / N, i = itc % N;
        double *gl = k.DZA + (long)i * nz + 8 * c;
        for (int t = 0; t < 8; ++t) {
          if ((t < 6 && i > 0) || (t >= 6 && i < N - 1)) rd0 = fmax(rd0, fabs(gl[t]));
          gl[t] = 0.0;
        }
      }
      rdn = k.rmax(rd0);
    }
    sc.pending = false;
    const double rpn = k.rmax(acc.st.rpn), lmax = k.rmax(acc.st.lmax), musum = k.rsum(acc.st.musum);
    const int m = k.rsumi(acc.st.m);
    res.rows += m;
    const double mu = (m > 0) ? musum / m : 0.0;
    if (rpn <= 1e-9 && rdn <= 1e-8 * (1.0 + cn) && mu <= 1e-10) { status = 0; break; }
    if (lmax > 1e13 || !(musum == musum)) { status = 2; break; }   // diverged: infeasible if the multipliers certify it (below)
    if (it >= 3 && rpn > 1e-5 && lmax > 1.0 + cn && m_dual_check(k, true, lmax, nullptr)) { status = 1; break; }   // early exit of infeasible relaxations
    // ---- predictor --------------------------------------------------------------------------
    m_riccati_factor(k);
    m_riccati_forward(k, k.DZA);
    acc.af.rmax = 1.0; acc.af.s1 = 0.0; acc.af.s2 = 0.0;
    PFOR(itc, C * N) {
      const int c = itc / N, i = itc % N;
      MPassD v; v.rows = k.rows + (long)c * k.kmaxc * N; v.N = N; v.i = i;
      v.y = k.Z + (long)i * nz + 8 * c; v.da = k.DZA + (long)i * nz + 8 * c;
      v.st.rmax = 1.0; v.st.s1 = 0.0; v.st.s2 = 0.0;
      m_car_rows(k, c, i, v);
      acc.af.rmax = fmax(acc.af.rmax, v.st.rmax); acc.af.s1 += v.st.s1; acc.af.s2 += v.st.s2;
    }
    if (P > 0) { PFOR(i, N) m_pair_stage<2>(k, i, sc, 0.0, false, acc); }
    const double rmaxa = k.rmax(acc.af.rmax), s1 = k.rsum(acc.af.s1), s2 = k.rsum(acc.af.s2);
    const double amin = 1.0 / rmaxa;
    double sigma = 0.0;
    if (m > 0 && mu > 0.0) {
      const double mu_aff = (musum + amin * s1 + amin * amin * s2) / m;
      const double r = mu_aff / mu;
      sigma = r * r * r;
      if (sigma > 1.0) sigma = 1.0;
      if (!(sigma >= 0.0)) sigma = 0.0;
    }
    const double sigmu = sigma * mu;
    // ---- corrector --------------------------------------------------------------------------
    PFOR(itc, C * N) {
      const int c = itc / N, i = itc % N;
      MPassE v; v.rows = k.rows + (long)c * k.kmaxc * N; v.N = N; v.i = i; v.sigmu = sigmu;
      v.y = k.Z + (long)i * nz + 8 * c; v.da = k.DZA + (long)i * nz + 8 * c;
      for (int t = 0; t < 8; ++t) v.gx[t] = 0.0;
      m_car_rows(k, c, i, v);
      const double *cst = D + p.o_cost + 16 * (c * N + i);
      double *gi = k.GR + (long)i * nz + 8 * c;
      for (int t = 0; t < 8; ++t) gi[t] = cst[t] * v.y[t] + cst[8 + t] + v.gx[t];
    }
    k.sync();
    if (P > 0) { PFOR(i, N) m_pair_stage<3>(k, i, sc, sigmu, false, acc); k.sync(); }
    m_riccati_vector(k);
    m_riccati_forward(k, k.DZ);
    acc.rmax = 0.0;
    int bad_step = 0;
    PFOR(itc, C * N) {
      const int c = itc / N, i = itc % N;
      MPassG v; v.rows = k.rows + (long)c * k.kmaxc * N; v.N = N; v.i = i; v.sigmu = sigmu;
      v.y = k.Z + (long)i * nz + 8 * c; v.d = k.DZ + (long)i * nz + 8 * c; v.da = k.DZA + (long)i * nz + 8 * c;
      v.rmax = 0.0;
      bad_step |= !(fabs(v.d[Y_UX]) + fabs(v.d[Y_UY]) + fabs(v.d[Y_PX]) + fabs(v.d[Y_PY]) < 1e300);
      m_car_rows(k, c, i, v);
      acc.rmax = fmax(acc.rmax, v.rmax);
    }
    if (P > 0) { PFOR(i, N) m_pair_stage<4>(k, i, sc, sigmu, false, acc); }
    const double rmaxg = k.rmax(acc.rmax);
    if (k.rany(bad_step) || !(rmaxg == rmaxg)) { status = 2; break; }
    const double alpha = (rmaxg > 0.995) ? 0.995 / rmaxg : 1.0;
    sc.alpha = alpha; sc.sigmu = sigmu; sc.pending = true;
    rdn *= (1.0 - alpha);
    PFOR(e, N * nz) k.Z[e] += alpha * k.DZ[e];
    PFOR(e, P * N * 4) k.sig[e * SG_SIZE + SG_VAL] += alpha * k.sig[e * SG_SIZE + SG_D];
    k.sync();
    if (alpha < 1e-6) { if (++stall >= 5) { status = 2; break; } } else stall = 0;
  }
  res.iters = it;
  res.converged = (status == 0);
  if (status == 2) {
    // not converged.  A primal feasible point is still usable (upper bound + Lagrangian lower bound); a violated one
    // closes the node only with a Farkas certificate, otherwise the outcome is "unknown"
    PairAcc acc; acc.worst = 0.0; acc.st.lmax = 0.0;
    PFOR(itc, C * N) {
      const int c = itc / N, i = itc % N;
      MPassViol v; v.rows = k.rows + (long)c * k.kmaxc * N; v.N = N; v.i = i; v.worst = 0.0; v.lmax = 0.0; v.y = k.Z + (long)i * nz + 8 * c;
      m_car_rows(k, c, i, v);
      acc.worst = fmax(acc.worst, v.worst); acc.st.lmax = fmax(acc.st.lmax, v.lmax);
    }
    if (P > 0) { PFOR(i, N) m_pair_stage<5>(k, i, sc, 0.0, false, acc); }
    int nan = !(acc.worst == acc.worst) || !(acc.st.lmax == acc.st.lmax);
    const double worst = k.rmax(acc.worst), lmx = k.rmax(acc.st.lmax);
    if (k.rany(nan) || !(lmx < MQM_INF)) status = 4;
    else if (worst <= 1e-7) { status = 0; m_dual_check(k, false, 1.0, &res.lb); }
    else status = (lmx > 0.0 && m_dual_check(k, true, lmx, nullptr)) ? 1 : 4;
  }
  res.status = status;
  if (status == 0) {
    double o = 0.0;
    PFOR(itc, C * N) {
      const int c = itc / N, i = itc % N;
      const double *cst = D + p.o_cost + 16 * (c * N + i);
      const double *z = k.Z + (long)i * nz + 8 * c;
      for (int t = 0; t < 8; ++t) o += (0.5 * cst[t] * z[t] + cst[8 + t]) * z[t];
    }
    PFOR(e, P * N * 4) { const double sv = k.sig[e * SG_SIZE + SG_VAL]; o += p.w_slack * sv * sv; }
    o = k.rsum(o);
    res.obj = o + p.cost_const;
    if (res.converged) res.lb = res.obj;
  }
  return res;
}

}  // namespace miqp
