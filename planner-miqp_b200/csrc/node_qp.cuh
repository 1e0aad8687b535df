// node_qp.cuh -- one CTA (a team of up to four warps) solves one node relaxation of the branch and bound.
//
// A node fixes a subset of the disjunctions of the cplexmodel .mod files (region + low-speed
// mode per car-step, model_region_constraints.mod:43-113 / minimum_speed_constraints.mod;
// environment polygon per point, separating obstacle edge per point,
// obstacle_environment_constraints.mod; collision side per pair quadruple,
// agent_collision_constraints.mod).  The decided alternatives contribute their rows without
// big-M, undecided ones contribute nothing, so the relaxation is a convex QP over the
// trajectory with the triple-integrator dynamics (model_region_constraints.mod:11-19):
//
//     min  sum_i 1/2 z_i' Q_i z_i + c_i' z_i      z_i = (px,vx,ax,py,vy,ay | ux,uy)
//     s.t. x_{i+1} = A x_i + B u_i,  x_0 given,  u_{N-1} = 0
//          G_i z_i <= h_i                          (stage-local rows)
//
// Solver: Mehrotra predictor-corrector interior point in STAGE space.  Every Newton step is
// a Riccati sweep over the banded KKT system.  Everything that is local to a stage (row
// generation, residuals, Hessian accumulation, step lengths) is spread over the whole team:
// sg = 6 .. 1 adjacent lanes share one stage and take its row slots round robin, so that
// N = 40 stages x 3 sub-lanes fill four warps; partial sums are combined with shuffles.  The
// sequential part runs on warp 0 while the others wait at the barrier: lanes = matrix entries
// for the backward Riccati factorisation, and the cheap vector sweeps are done redundantly by
// all lanes without any synchronisation.  Inequality rows are never stored:
// their coefficients are regenerated from the small per-plan tables in every pass; only the
// slack and multiplier (s, lambda) of every row live in shared memory, laid out
// [slot][stage] with the stage stride padded so that the (stage, sub-lane) pattern of a warp is
// bank-conflict free for 16-byte records; residuals, the
// affine step and the pending Newton step are recomputed from the stage vectors instead of
// being stored (3 dot products per row instead of 2 more arrays).
//
// This file handles one car per plan (C == 1); the multi-car kernel couples the cars of a
// stage through the pair rows and is a separate instantiation.
#pragma once
#include "dev_problem.cuh"

namespace miqp {

#define MQ_INF (__longlong_as_double(0x7ff0000000000000LL))
constexpr unsigned FULL = 0xffffffffu;

// per-stage shared-memory records (strides are odd so that lane = stage accesses are
// bank-conflict free for 64-bit words)
// Stage record of the Riccati recursion.  The Hessian block Mxx (pass A) is dead once its stage has been factorised, and the
// factors of the stage (G, Finv: read by the sweeps) take its place: 23 doubles per stage instead of 39 -- 5 kB less shared
// memory per node at N = 40, which is what lets four two-warp teams share an SM.
constexpr int S_STRIDE = 23;  // [0..20] Mxx packed lower (until the stage is factorised), then [0..11] G = Phi_ux (2x6), [12..14] Finv; [21..22] Muu diag
constexpr int S_G = 0, S_FINV = 12, S_MUU = 21;
constexpr int V_STRIDE = 35;  // z[8] g[8] dz[8] dza[8] k[2]
constexpr int V_Z = 0, V_G = 8, V_DZ = 16, V_DZA = 24, V_K = 32;
constexpr int T_P = 0, T_PV = 48, T_PUU = 60, T_PHIU = 63, T_SIZE = 66;  // warp scratch: 2 x P (24), 2 x p (6), Phi_uu (3), phi_u (2)

// 1/x for positive normal x (slacks, multipliers, pivots): MUFU.RCP64H seed (about 20 bits)
// and one third-order Newton step (e + e^2), i.e. 1 MUFU + 3 DFMA instead of the IEEE
// division sequence with its slow path.  Relative error < 2^-52.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  r = fma(r, t, r);
  const double e2 = fma(-x, r, 1.0);   // one more first-order step: exact to rounding
  return fma(r, e2, r);
}

__device__ __forceinline__ int pidx(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

struct WarpCtx {
  const DevProb *p;
  const double *D;
  const int *I;
  double *S;            // [N][S_STRIDE]
  double *V;            // [N][V_STRIDE]
  unsigned char *dec;   // node decisions
  unsigned char *imp;   // implied completion (scan)
  int *jeff;            // effective region per stage (-1 unknown)
  int *aux;             // [3N] scan tables: best alt, region_decided, blame
  double *auxd;         // [N] scan: best non-frozen violation
  double *bnd;          // [12][NB] bound rows of the node: hi / lo of vx, ax, vy, ay, ux, uy per stage (+-INF: no row); filled once per node
  int NB;               // stage stride of bnd
  double *T;            // [T_SIZE] Riccati scratch (value function of the next stage)
  double2 *rows;        // [kmax+1][NP] (s, lambda) per inequality row
  double *red;          // [8][4] team reduction scratch
  double tau_k;
  double *dbgrow;       // [-DMQ_PROF] iteration trace of this CTA
  int lane, N;
  int wid, nw;          // warp of the team, warps per team
  int sg, g, ls, spw;   // sub-lanes per stage, my sub-lane, my stage slot in the warp (idle if >= spw), stages per warp
  int sgmul;            // 65536 / sg + 1: slot % sg without a division
  int NP;               // stage stride of `rows`
  int kmax;             // row slots per stage
  __device__ __forceinline__ bool mine(int slot) const { const int q = (slot * sgmul) >> 16; return slot - q * sg == g; }
};

// sub-lanes per stage for a team of nw warps and the padded stage stride of the row records
__host__ __device__ inline int team_sublanes(int N, int nw) {
  for (int s = 6; s > 1; --s) if (nw * (32 / s) >= N) return s;
  return 1;
}
__host__ __device__ inline int team_row_stride(int N, int sg) {
  const int target = (sg == 4) ? 2 : (sg == 2) ? 4 : (sg == 1) ? (N & 7) : 3;   // sg = 5, 6: two-way conflicts at most
  return N + ((target - N) & 7);
}

// stages of this thread: i = wid * spw + ls, then in steps of nw * spw; every thread runs the same
// number of trips (shuffles and barriers inside the body stay uniform), `act` says whether i is real
#define MQ_FOR_STAGES(i, act)                                                                        \
  for (int i = w.wid * w.spw + w.ls, c0_ = 0; c0_ < N; c0_ += w.nw * w.spw, i += w.nw * w.spw)       \
    if (const bool act = (w.ls < w.spw) && (i < N); true)

__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) { double t = __shfl_xor_sync(FULL, v, o); v = t > v ? t : v; }
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o > 0; o >>= 1) { double t = __shfl_xor_sync(FULL, v, o); v = t < v ? t : v; }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// sums over the sg adjacent lanes that share a stage, for CNT values at once; the results are valid in the leader
// (g == 0).  One straight-line variant per sg: with a run-time trip count the shuffles of the values serialise
// (6.3 k cycles for the 31 sums of pass A instead of a few hundred).
template <int SG, int CNT>
__device__ __forceinline__ void sg_reduce_fixed(double (&a)[CNT]) {
#pragma unroll
  for (int t = 0; t < CNT; ++t) {
    const double v = a[t];
    double r = v;
#pragma unroll
    for (int d = 1; d < SG; ++d) r += __shfl_down_sync(FULL, v, d);
    a[t] = r;
  }
}
template <int CNT>
__device__ __forceinline__ void sg_reduce(const WarpCtx &w, double (&a)[CNT]) {
  switch (w.sg) {   // uniform over the team
    case 2: sg_reduce_fixed<2, CNT>(a); break;
    case 3: sg_reduce_fixed<3, CNT>(a); break;
    case 4: sg_reduce_fixed<4, CNT>(a); break;
    case 5: sg_reduce_fixed<5, CNT>(a); break;
    case 6: sg_reduce_fixed<6, CNT>(a); break;
    default: break;
  }
}
// team-wide {max a, max b, sum c, sum d}; every thread receives bitwise identical results
__device__ __forceinline__ void team_reduce(const WarpCtx &w, double &a, double &b, double &c, double &d) {
  a = warp_max(a); b = warp_max(b); c = warp_sum(c); d = warp_sum(d);
  if (w.lane == 0) { double *r = w.red + 4 * w.wid; r[0] = a; r[1] = b; r[2] = c; r[3] = d; }
  __syncthreads();
  a = w.red[0]; b = w.red[1]; c = w.red[2]; d = w.red[3];
  for (int k = 1; k < w.nw; ++k) { a = fmax(a, w.red[4 * k]); b = fmax(b, w.red[4 * k + 1]); c += w.red[4 * k + 2]; d += w.red[4 * k + 3]; }
  __syncthreads();
}

// bounds of a stage (model_region_constraints.mod:22-39 and the per-region boxes :73-94
// of the effective region; low-speed box of minimum_speed_constraints.mod when frozen)
__device__ __forceinline__ void stage_bounds(const WarpCtx &w, int i, int je, bool frozen, double lo[8], double hi[8]) {
  const DevProb &p = *w.p;
#pragma unroll
  for (int t = 0; t < 8; ++t) { lo[t] = -MQ_INF; hi[t] = MQ_INF; }
  lo[Y_VX] = p.min_vel; hi[Y_VX] = p.max_vel; lo[Y_VY] = p.min_vel;  // vel_y has no upper bound
  lo[Y_AX] = p.total_min_acc; hi[Y_AX] = p.total_max_acc; lo[Y_AY] = p.total_min_acc; hi[Y_AY] = p.total_max_acc;
  lo[Y_UX] = p.total_min_jerk; hi[Y_UX] = p.total_max_jerk; lo[Y_UY] = p.total_min_jerk; hi[Y_UY] = p.total_max_jerk;
  if (je >= 0) {
    const double *D = w.D;
    const int q = je;  // car 0
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

// bound rows of every stage of the node in slot order (hi, lo of vx, ax, vy, ay, ux, uy); a row that does not exist (no state
// rows at stage 0, no input rows at the last stage, an infinite bound) is stored as +-INF.  Called by the whole team once per node.
__device__ __forceinline__ void fill_stage_bounds(const WarpCtx &w) {
  const DevProb &p = *w.p;
  const int N = w.N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const unsigned char m = (i > 0) ? w.dec[p.off_mode + i] : (unsigned char)0;
    double lo[8], hi[8];
    stage_bounds(w, i, w.jeff[i], i > 0 && m == MODE_FROZEN, lo, hi);
    const bool st = (i > 0), ut = (i < N - 1);
    const int T[6] = {Y_VX, Y_AX, Y_VY, Y_AY, Y_UX, Y_UY};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const bool act = (k < 4) ? st : ut;
      w.bnd[(2 * k) * w.NB + i] = act ? hi[T[k]] : MQ_INF;
      w.bnd[(2 * k + 1) * w.NB + i] = act ? lo[T[k]] : -MQ_INF;
    }
  }
}

// row of one polygon edge for point pt of the car in region j:
// sign=+1: cross(P)/len <= 0 (obstacle, chosen edge); sign=-1: cross(P)/len >= 0 (environment)
// points: 0 rear, 1 (xU,yU), 2 (xL,yU), 3 (xU,yL), 4 (xL,yL)
__device__ __forceinline__ void edge_row(const double *et, const double *ft, int pt, double sign, double a[6], double &rhs) {
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

// the five rows of a decided rho=0 mode (j,h): wedge (2), curvature (2), speed half plane (1)
__device__ __forceinline__ void mode_row(const WarpCtx &w, int j, int h, int k, double a[6], double &rhs) {
  const DevProb &p = *w.p;
  if (k < 4) {
    const double *t = w.D + p.o_modetab + 20 * j + 5 * k;
    a[Y_PX] = 0.0; a[Y_VX] = t[0]; a[Y_AX] = t[1]; a[Y_PY] = 0.0; a[Y_VY] = t[2]; a[Y_AY] = t[3];
    rhs = t[4];
  } else {
    a[Y_PX] = 0.0; a[Y_AX] = 0.0; a[Y_PY] = 0.0; a[Y_AY] = 0.0;
    a[Y_VX] = (h == 0) ? -1.0 : (h == 2) ? 1.0 : 0.0;   // h: 0 vx>=vm, 1 vy>=vm, 2 vx<=-vm, 3 vy<=-vm
    a[Y_VY] = (h == 1) ? -1.0 : (h == 3) ? 1.0 : 0.0;
    rhs = -p.vm;
  }
}

// Enumerates the inequality rows of stage i of the node in a fixed slot order; a thread only visits
// the slots of its sub-lane (WarpCtx::mine).
// Vis::bound<T>(slot, sgn, rhs):  sgn * y[T] <= rhs ;  Vis::general(slot, a[6], rhs):  a.x <= rhs
template <class Vis>
__device__ __forceinline__ void visit_rows(const WarpCtx &w, int i, Vis &v) {
  const DevProb &p = *w.p;
  const int N = w.N;
  const unsigned char m = (i > 0) ? w.dec[p.off_mode + i] : (unsigned char)0;
  const int je = w.jeff[i];
  // bound rows: their right-hand sides are constants of the node (fill_stage_bounds), read from shared memory instead of being
  // rebuilt from the per-plan tables in every pass of every iteration
  int slot = 0;
#define MQ_BND(T)                                                                                  \
  {                                                                                                \
    if (w.mine(slot)) { const double h_ = w.bnd[slot * w.NB + i]; if (h_ < MQ_INF) v.template bound<T>(slot, 1.0, h_); }            \
    ++slot;                                                                                        \
    if (w.mine(slot)) { const double l_ = w.bnd[slot * w.NB + i]; if (l_ > -MQ_INF) v.template bound<T>(slot, -1.0, -l_); }         \
    ++slot;                                                                                        \
  }
  MQ_BND(Y_VX) MQ_BND(Y_AX) MQ_BND(Y_VY) MQ_BND(Y_AY) MQ_BND(Y_UX) MQ_BND(Y_UY)
#undef MQ_BND
  if (i == 0) return;
  // General rows: the sub-lanes of a stage walk every family in lockstep (item = k * sg + g), so that all
  // lanes of a warp execute the same code on different rows; a global round robin over the slots would
  // put the sub-lanes on different row kinds and serialise them.
  double a[6], rhs;
  const int sg = w.sg, g = w.g;
  if (m != UNDEC && m != MODE_FROZEN) {
    const int j = m >> 2, h = m & 3;
#pragma unroll 1
    for (int k = g; k < 5; k += sg) { mode_row(w, j, h, k, a, rhs); v.general(slot + k, a, rhs); }
  }
  slot += 5;
  const double *ft = w.D + p.o_fronttab + 12 * (je >= 0 ? je : 0);
  if (p.E > 0) {
    const int ME = p.maxEnvEdges;
    // item = pt * ME + ed, tracked incrementally (no division by the run-time ME); with one environment polygon -- the common
    // case -- its edge range is read once per visit instead of once per row (two dependent table loads less in front of every row)
    const bool single = (p.E == 1);
    int e0s = 0, nes = 0;
    if (single) { e0s = w.I[p.o_env_off]; nes = w.I[p.o_env_off + 1] - e0s; }
    int pt = 0, ed = g;
    while (ed >= ME) { ed -= ME; ++pt; }
#pragma unroll 1
    for (int item = g; item < 5 * ME; item += sg) {
      int e0 = e0s, ne = nes;
      bool act = (pt == 0 || je >= 0);
      if (!single) {
        const int e = w.dec[p.off_env + i * 5 + pt];
        act = act && (e != UNDEC);
        if (act) { e0 = w.I[p.o_env_off + e]; ne = w.I[p.o_env_off + e + 1] - e0; }
      }
      if (act && ed < ne) { edge_row(w.D + p.o_envtab + 3 * (e0 + ed), ft, pt, -1.0, a, rhs); v.edge(slot + item, a, rhs); }
      ed += sg;
      while (ed >= ME) { ed -= ME; ++pt; }
    }
    slot += 5 * ME;
  }
#pragma unroll 1
  for (int item = g; item < 5 * p.O; item += sg) {
    const int o = item / 5, pt = item - 5 * o;
    const unsigned char d = w.dec[p.off_obs + (o * N + i) * 5 + pt];
    if (d != UNDEC && d != OBS_SOFT && (pt == 0 || je >= 0)) {
      edge_row(w.D + p.o_obstab + 3 * ((o * N + i) * p.L + d), ft, pt, 1.0, a, rhs);
      v.edge(slot + item, a, rhs);
    }
  }
}

// ---------------------------------------------------------------------------------------
// row passes
// ---------------------------------------------------------------------------------------
struct StepCtx {  // quantities of the last Newton step, needed to recompute it row by row
  double alpha, sigmu;
  bool pending;
};

// (s, lambda) of the inequality rows of one stage, [slot][stage] in the team's shared memory.
// (A variant with the records in an L2-resident slice of HBM halves the shared memory of a node but ran
// 1.6x slower: profiles/r1_ab_rows_l2_vs_smem.md.)
struct RowIO {
  double2 *base;   // rows + i
  int NP;
  __device__ __forceinline__ void init(double2 *rows, int NP_, int i) { base = rows + i; NP = NP_; }
  __device__ __forceinline__ double2 ld(int slot) { return base[slot * NP]; }
  __device__ __forceinline__ void st(int slot, double2 v) { base[slot * NP] = v; }
};

__device__ __forceinline__ double dot6(const double a[6], const double y[8]) {
  double v = 0.0;
#pragma unroll
  for (int t = 0; t < 6; ++t) v += a[t] * y[t];
  return v;
}

// polygon-edge rows have no acceleration terms (edge_row: a[Y_AX] = a[Y_AY] = 0): four-term products, ten Hessian entries
__device__ __forceinline__ double dot4(const double a[6], const double y[8]) {
  double v = 0.0;
  v += a[Y_PX] * y[Y_PX]; v += a[Y_VX] * y[Y_VX]; v += a[Y_PY] * y[Y_PY]; v += a[Y_VY] * y[Y_VY];
  return v;
}

struct PassInit {  // cold: s = max(h - g.z, 1), lambda = 1; warm: s = max(h - g.z, smin), lambda = mu0 / s; gl = G' lambda
  RowIO io; double y[8]; double gl[8];
  double mu0, smin;   // mu0 <= 0: cold start
  __device__ __forceinline__ double put(int slot, double gz, double rhs) {
    const double sl = rhs - gz;
    double s, lam;
    if (mu0 > 0.0) { s = sl > smin ? sl : smin; lam = mu0 * fast_rcp(s); }
    else { s = sl > 1.0 ? sl : 1.0; lam = 1.0; }
    io.st(slot, make_double2(s, lam));
    return lam;
  }
  template <int T> __device__ __forceinline__ void bound(int slot, double sgn, double rhs) { gl[T] += sgn * put(slot, sgn * y[T], rhs); }
  __device__ __forceinline__ void general(int slot, const double a[6], double rhs) {
    const double lam = put(slot, dot6(a, y), rhs);
#pragma unroll
    for (int t = 0; t < 6; ++t) gl[t] += a[t] * lam;
  }
  __device__ __forceinline__ void edge(int slot, const double a[6], double rhs) { general(slot, a, rhs); }
};

struct PassA {  // apply the pending step, residuals, Hessian and predictor gradient
  RowIO io; StepCtx sc;
  double y[8], d[8], da[8];
  double H[21], Huu[2], gx[8];
  double rpn, musum, lmax; int m;
  // returns (weight, weight*rp) of the row after the update
  __device__ __forceinline__ void core(int slot, double gz, double gdz, double gda, double rhs, double &wgt, double &wr) {
    double2 v = io.ld(slot);
    double s = v.x, lam = v.y;
    if (sc.pending) {
      const double rp_old = (gz - sc.alpha * gdz) + s - rhs;
      const double inv = fast_rcp(s);
      const double dsa = -rp_old - gda;
      const double dla = -lam - (lam * inv) * dsa;
      const double ds = -rp_old - gdz;
      const double rc = s * lam + dsa * dla - sc.sigmu;
      const double dl = -(rc + lam * ds) * inv;
      s += sc.alpha * ds; lam += sc.alpha * dl;
      io.st(slot, make_double2(s, lam));
    }
    const double rp = gz + s - rhs;
    wgt = lam * fast_rcp(s); wr = wgt * rp;
    rpn = fmax(rpn, fabs(rp)); musum += s * lam; lmax = fmax(lmax, lam); ++m;
  }
  template <int T> __device__ __forceinline__ void bound(int slot, double sgn, double rhs) {
    double wgt, wr;
    core(slot, sgn * y[T], sgn * d[T], sgn * da[T], rhs, wgt, wr);
    if (T < 6) H[T * (T + 1) / 2 + T] += wgt; else Huu[T - 6] += wgt;
    gx[T] += sgn * wr;
  }
  __device__ __forceinline__ void general(int slot, const double a[6], double rhs) {
    double wgt, wr;
    core(slot, dot6(a, y), dot6(a, d), dot6(a, da), rhs, wgt, wr);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double wa = wgt * a[r];
#pragma unroll
      for (int c = 0; c <= r; ++c) H[r * (r + 1) / 2 + c] += wa * a[c];
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) gx[t] += a[t] * wr;
  }
  __device__ __forceinline__ void edge(int slot, const double a[6], double rhs) {
    double wgt, wr;
    core(slot, dot4(a, y), dot4(a, d), dot4(a, da), rhs, wgt, wr);
    const double wpx = wgt * a[Y_PX], wvx = wgt * a[Y_VX], wpy = wgt * a[Y_PY], wvy = wgt * a[Y_VY];
    // packed lower triangle, index r (r + 1) / 2 + c with (PX, VX, PY, VY) = (0, 1, 3, 4)
    H[0] += wpx * a[Y_PX];
    H[1] += wvx * a[Y_PX]; H[2] += wvx * a[Y_VX];
    H[6] += wpy * a[Y_PX]; H[7] += wpy * a[Y_VX]; H[9] += wpy * a[Y_PY];
    H[10] += wvy * a[Y_PX]; H[11] += wvy * a[Y_VX]; H[13] += wvy * a[Y_PY]; H[14] += wvy * a[Y_VY];
    gx[Y_PX] += a[Y_PX] * wr; gx[Y_VX] += a[Y_VX] * wr; gx[Y_PY] += a[Y_PY] * wr; gx[Y_VY] += a[Y_VY] * wr;
  }
};

struct PassD {  // affine step: ratios and the three sums that give mu_aff for any step length
  RowIO io; double y[8], da[8];
  double rmax, s1, s2;   // rmax = 1 / (largest feasible affine step), kept >= 1
  __device__ __forceinline__ void row(int slot, double gz, double gda, double rhs) {
    const double2 v = io.ld(slot);
    const double s = v.x, lam = v.y;
    const double rp = gz + s - rhs;
    const double dsa = -rp - gda;
    const double t = dsa * fast_rcp(s);
    const double dla = -lam - lam * t;
    // s + a dsa >= 0  <=>  1/a >= -dsa/s = -t ;  lam + a dla >= 0  <=>  1/a >= -dla/lam = 1 + t
    rmax = fmax(rmax, fmax(-t, 1.0 + t));
    s1 += s * dla + lam * dsa; s2 += dsa * dla;
  }
  template <int T> __device__ __forceinline__ void bound(int slot, double sgn, double rhs) { row(slot, sgn * y[T], sgn * da[T], rhs); }
  __device__ __forceinline__ void general(int slot, const double a[6], double rhs) { row(slot, dot6(a, y), dot6(a, da), rhs); }
  __device__ __forceinline__ void edge(int slot, const double a[6], double rhs) { row(slot, dot4(a, y), dot4(a, da), rhs); }
};

struct PassE {  // corrector gradient
  RowIO io; double sigmu; double y[8], da[8]; double gx[8];
  __device__ __forceinline__ double coef(int slot, double gz, double gda, double rhs) {
    const double2 v = io.ld(slot);
    const double s = v.x, lam = v.y;
    const double inv = fast_rcp(s);
    const double rp = gz + s - rhs;
    const double dsa = -rp - gda;
    const double dla = -lam - (lam * inv) * dsa;
    return (lam * rp - (dsa * dla - sigmu)) * inv;
  }
  template <int T> __device__ __forceinline__ void bound(int slot, double sgn, double rhs) { gx[T] += sgn * coef(slot, sgn * y[T], sgn * da[T], rhs); }
  __device__ __forceinline__ void general(int slot, const double a[6], double rhs) {
    const double cf = coef(slot, dot6(a, y), dot6(a, da), rhs);
#pragma unroll
    for (int t = 0; t < 6; ++t) gx[t] += a[t] * cf;
  }
  __device__ __forceinline__ void edge(int slot, const double a[6], double rhs) {
    const double cf = coef(slot, dot4(a, y), dot4(a, da), rhs);
    gx[Y_PX] += a[Y_PX] * cf; gx[Y_VX] += a[Y_VX] * cf; gx[Y_PY] += a[Y_PY] * cf; gx[Y_VY] += a[Y_VY] * cf;
  }
};

struct PassG {  // step length of the combined step
  RowIO io; double sigmu; double y[8], d[8], da[8];
  double rmax;   // 1 / (largest step that keeps s and lambda non-negative), starts at 0
  __device__ __forceinline__ void row(int slot, double gz, double gdz, double gda, double rhs) {
    const double2 v = io.ld(slot);
    const double s = v.x, lam = v.y;
    const double inv = fast_rcp(s);
    const double rp = gz + s - rhs;
    const double dsa = -rp - gda;
    const double dla = -lam - (lam * inv) * dsa;
    const double ds = -rp - gdz;
    const double rc = s * lam + dsa * dla - sigmu;
    const double dl = -(rc + lam * ds) * inv;
    rmax = fmax(rmax, fmax(-ds * inv, -dl * fast_rcp(lam)));
  }
  template <int T> __device__ __forceinline__ void bound(int slot, double sgn, double rhs) { row(slot, sgn * y[T], sgn * d[T], sgn * da[T], rhs); }
  __device__ __forceinline__ void general(int slot, const double a[6], double rhs) { row(slot, dot6(a, y), dot6(a, d), dot6(a, da), rhs); }
  __device__ __forceinline__ void edge(int slot, const double a[6], double rhs) { row(slot, dot4(a, y), dot4(a, d), dot4(a, da), rhs); }
};

struct PassViol {  // worst primal violation of the current point, largest multiplier
  RowIO io; double y[8]; double worst, lmax;
  template <int T> __device__ __forceinline__ void bound(int slot, double sgn, double rhs) { worst = fmax(worst, sgn * y[T] - rhs); lmax = fmax(lmax, io.ld(slot).y); }
  __device__ __forceinline__ void general(int slot, const double a[6], double rhs) { worst = fmax(worst, dot6(a, y) - rhs); lmax = fmax(lmax, io.ld(slot).y); }
  __device__ __forceinline__ void edge(int slot, const double a[6], double rhs) { general(slot, a, rhs); }
};

struct PassDual {  // G' lambda, h' lambda and lambda' G z of a stage (multipliers after the pending step, scaled)
  RowIO io; StepCtx sc; double scale;
  double y[8], d[8], da[8];
  double gl[8]; double hl, lgz, mag;
  __device__ __forceinline__ double lam_of(int slot, double gz, double gdz, double gda, double rhs) {
    const double2 v = io.ld(slot);
    double lam = v.y;
    if (sc.pending) {   // same update as PassA::core
      const double s = v.x;
      const double rp_old = (gz - sc.alpha * gdz) + s - rhs;
      const double inv = fast_rcp(s);
      const double dsa = -rp_old - gda;
      const double dla = -lam - (lam * inv) * dsa;
      const double ds = -rp_old - gdz;
      const double rc = s * lam + dsa * dla - sc.sigmu;
      const double dl = -(rc + lam * ds) * inv;
      lam += sc.alpha * dl;
    }
    lam = (lam > 0.0 ? lam : 0.0) * scale;
    hl += lam * rhs; lgz += lam * gz; mag += lam * fabs(rhs);
    return lam;
  }
  template <int T> __device__ __forceinline__ void bound(int slot, double sgn, double rhs) { gl[T] += sgn * lam_of(slot, sgn * y[T], sgn * d[T], sgn * da[T], rhs); }
  __device__ __forceinline__ void general(int slot, const double a[6], double rhs) {
    const double lam = lam_of(slot, dot6(a, y), dot6(a, d), dot6(a, da), rhs);
#pragma unroll
    for (int t = 0; t < 6; ++t) gl[t] += a[t] * lam;
  }
  __device__ __forceinline__ void edge(int slot, const double a[6], double rhs) { general(slot, a, rhs); }
};

// ---------------------------------------------------------------------------------------
// Riccati sweeps
// ---------------------------------------------------------------------------------------
// column a of [A B] (6x8): rows 3*axis .. 3*axis+cnt-1 with coefficients cf[]
__device__ __forceinline__ void ab_column(const DevProb &p, int a, int &row0, int &cnt, double cf[3]) {
  if (a < 6) {
    const int o = a % 3;
    row0 = (a / 3) * 3; cnt = o + 1;
    // T = [[1,ts,c2],[0,1,ts],[0,0,1]]: column o = (T[0][o], .., T[o][o])
    cf[0] = (o == 0) ? 1.0 : (o == 1) ? p.ts : p.c2;
    cf[1] = (o == 1) ? 1.0 : p.ts;
    cf[2] = 1.0;
  } else {
    row0 = (a - 6) * 3; cnt = 3;
    cf[0] = p.c3; cf[1] = p.c2; cf[2] = p.ts;
  }
}

struct PhiEntry {  // one entry (a,b) of Phi = M + [A B]' P [A B]: sum_k fa[k] sum_l fb[l] P[pi[3k+l]]
  int a, b;
  int pi[9]; double fa[3], fb[3];
  __device__ __forceinline__ void setup(const DevProb &p, int e) {
    if (e < 21) { a = 0; while ((a + 1) * (a + 2) / 2 <= e) ++a; b = e - a * (a + 1) / 2; }
    else if (e < 33) { a = 6 + (e - 21) / 6; b = (e - 21) % 6; }
    else { a = (e == 33) ? 6 : 7; b = (e == 35) ? 7 : 6; }
    int ra, ca, rb, cb;
    ab_column(p, a, ra, ca, fa); ab_column(p, b, rb, cb, fb);
#pragma unroll
    for (int k = 0; k < 3; ++k) { if (k >= ca) fa[k] = 0.0; if (k >= cb) fb[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int l = 0; l < 3; ++l) pi[k * 3 + l] = (k < ca && l < cb) ? pidx(ra + k, rb + l) : 0;
  }
  // three independent inner chains of three, then one outer chain of three (6 dependent DFMA instead of 9)
  __device__ __forceinline__ double eval(const double *P) const {
    double r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = fma(fb[2], P[pi[3 * k + 2]], fma(fb[1], P[pi[3 * k + 1]], fb[0] * P[pi[3 * k]]));
    return fma(fa[2], r[2], fma(fa[1], r[1], fa[0] * r[0]));
  }
};

// Backward factorisation + predictor vector.  On return S_i holds G_i, Finv_i and V_i holds
// k_i for every stage; the value function (P, p) of the next stage only lives in the warp
// scratch (double buffered).  Straight-line code per stage: every lane evaluates its entry lane of
// Phi (0..31) AND, interleaved, a second entry (32 + lane for lanes 0..3, a zero dummy elsewhere) and
// one component of phi = g + [A B]' p (lanes 0..7); only the stores are predicated.
__device__ __forceinline__ void riccati_factor(const WarpCtx &w, const PhiEntry &e1, const PhiEntry &e2) {
  const DevProb &p = *w.p;
  const int lane = w.lane, N = w.N;
  double *T = w.T;
  // terminal stage: P = Mxx, p = g_x
  {
    const int nb = (N - 1) & 1;
    if (lane < 21) T[T_P + 24 * nb + lane] = w.S[(N - 1) * S_STRIDE + lane];
    if (lane < 6) T[T_PV + 6 * nb + lane] = w.V[(N - 1) * V_STRIDE + V_G + lane];
  }
  __syncwarp();
  int prow0 = 0, pcnt = 0; double pcf[3] = {0, 0, 0};
  if (lane < 8) ab_column(p, lane, prow0, pcnt, pcf);
#pragma unroll
  for (int k = 0; k < 3; ++k) if (k >= pcnt) pcf[k] = 0.0;
  const int la = (lane < 21) ? e1.a : 0, lb = (lane < 21) ? e1.b : 0, l6 = (lane < 6) ? lane : 0;
  for (int i = N - 2; i >= 0; --i) {
    double *Si = w.S + i * S_STRIDE, *Vi = w.V + i * V_STRIDE;
    const double *Pn = T + T_P + 24 * ((i + 1) & 1), *pn = T + T_PV + 6 * ((i + 1) & 1);
    double *Pc = T + T_P + 24 * (i & 1), *pc = T + T_PV + 6 * (i & 1);
    // phase 1: Phi entries and phi
    // unconditional loads from clamped indices + selects (a conditional load compiles to a branch)
    const double m1r = Si[lane < 21 ? lane : 0], m2r = Si[S_MUU + (lane == 3 ? 1 : 0)], gvr = Vi[V_G + (lane < 8 ? lane : 0)];
    const double m1 = (lane < 21) ? m1r : 0.0;                                                 // + Mxx (M_ux = 0)
    const double m2 = (lane == 1 || lane == 3) ? m2r : 0.0;                                    // + Muu on entries 33, 35
    const double gv = (lane < 8) ? gvr : 0.0;
    const double phi1 = m1 + e1.eval(Pn);
    const double phi2 = m2 + e2.eval(Pn);
    const double phiv = fma(pcf[2], pn[prow0 + 2 < 6 ? prow0 + 2 : 5], fma(pcf[1], pn[prow0 + 1 < 6 ? prow0 + 1 : 5], fma(pcf[0], pn[prow0], gv)));
    __syncwarp();                                          // every lane has read its Mxx entry: the record is reused for G
    if (lane >= 21) Si[S_G + lane - 21] = phi1;            // G entries 0..10 (Phi entries 21..31)
    if (lane == 0) Si[S_G + 11] = phi2;                    // G entry 11 (Phi entry 32)
    if (lane >= 1 && lane < 4) T[T_PUU + lane - 1] = phi2; // Phi_uu
    if (lane == 6 || lane == 7) T[T_PHIU + lane - 6] = phiv;
    __syncwarp();
    // phase 2: Finv, P_i, p_i, k_i
    const double f00 = T[T_PUU], f10 = T[T_PUU + 1], f11 = T[T_PUU + 2];
    const double pu0 = T[T_PHIU], pu1 = T[T_PHIU + 1];
    const double g0a = Si[S_G + la], g1a = Si[S_G + 6 + la], g0b = Si[S_G + lb], g1b = Si[S_G + 6 + lb];
    const double h0 = Si[S_G + l6], h1 = Si[S_G + 6 + l6];
    const double idet = fast_rcp(f00 * f11 - f10 * f10);
    const double i00 = f11 * idet, i10 = -f10 * idet, i11 = f00 * idet;
    const double k0 = -(i00 * pu0 + i10 * pu1), k1 = -(i10 * pu0 + i11 * pu1);
    const double w0 = i00 * g0b + i10 * g1b, w1 = i10 * g0b + i11 * g1b;
    if (lane < 21) Pc[lane] = phi1 - (g0a * w0 + g1a * w1);
    if (lane < 6) pc[lane] = phiv + h0 * k0 + h1 * k1;
    if (lane == 31) { Si[S_FINV] = i00; Si[S_FINV + 1] = i10; Si[S_FINV + 2] = i11; Vi[V_K] = k0; Vi[V_K + 1] = k1; }
    __syncwarp();
  }
}

// stage record of the sweeps, fetched one stage ahead of its use (the compiler may not move the
// shared-memory loads of stage i+1 above the stores of stage i by itself)
struct SweepRec {
  double G[12], F[3];
  __device__ __forceinline__ void load(const double *Si) {
#pragma unroll
    for (int t = 0; t < 12; ++t) G[t] = Si[S_G + t];
    F[0] = Si[S_FINV]; F[1] = Si[S_FINV + 1]; F[2] = Si[S_FINV + 2];
  }
};

// vector-only backward sweep with a new gradient (V_G), all lanes redundantly; two stages per trip with
// ping-pong records so that the prefetched values are never copied
__device__ __forceinline__ void riccati_vector_stage(const WarpCtx &w, int i, const SweepRec &rec, const double g[8], double pn[6],
                                                     double ts, double c2, double c3) {
  double phi[8];
#pragma unroll
  for (int ax = 0; ax < 2; ++ax) {
    const double pp = pn[3 * ax], pv = pn[3 * ax + 1], pa = pn[3 * ax + 2];
    phi[3 * ax] = g[3 * ax] + pp;
    phi[3 * ax + 1] = g[3 * ax + 1] + ts * pp + pv;
    phi[3 * ax + 2] = g[3 * ax + 2] + c2 * pp + ts * pv + pa;
    phi[6 + ax] = g[6 + ax] + c3 * pp + c2 * pv + ts * pa;
  }
  const double k0 = -(rec.F[0] * phi[6] + rec.F[1] * phi[7]), k1 = -(rec.F[1] * phi[6] + rec.F[2] * phi[7]);
#pragma unroll
  for (int t = 0; t < 6; ++t) pn[t] = phi[t] + rec.G[t] * k0 + rec.G[6 + t] * k1;
  if (w.lane == 6) { double *Vi = w.V + i * V_STRIDE; Vi[V_K] = k0; Vi[V_K + 1] = k1; }
}
__device__ __forceinline__ void riccati_vector(const WarpCtx &w) {
  const DevProb &p = *w.p;
  const int N = w.N;
  const double ts = p.ts, c2 = p.c2, c3 = p.c3;
  double pn[6];
#pragma unroll
  for (int t = 0; t < 6; ++t) pn[t] = w.V[(N - 1) * V_STRIDE + V_G + t];
  SweepRec ra, rb;
  double ga[8], gb[8];
  ra.load(w.S + (N - 2) * S_STRIDE);
#pragma unroll
  for (int t = 0; t < 8; ++t) ga[t] = w.V[(N - 2) * V_STRIDE + V_G + t];
  for (int i = N - 2; i >= 0; i -= 2) {
    const int i1 = (i > 0) ? i - 1 : 0, i2 = (i > 1) ? i - 2 : 0;
    rb.load(w.S + i1 * S_STRIDE);
#pragma unroll
    for (int t = 0; t < 8; ++t) gb[t] = w.V[i1 * V_STRIDE + V_G + t];
    riccati_vector_stage(w, i, ra, ga, pn, ts, c2, c3);
    if (i == 0) break;
    ra.load(w.S + i2 * S_STRIDE);
#pragma unroll
    for (int t = 0; t < 8; ++t) ga[t] = w.V[i2 * V_STRIDE + V_G + t];
    riccati_vector_stage(w, i - 1, rb, gb, pn, ts, c2, c3);
  }
  __syncwarp();
}

// forward sweep, all lanes redundantly; writes the step of every stage at offset `dst`
__device__ __forceinline__ void riccati_forward_stage(const WarpCtx &w, int i, int dst, bool has_input, const SweepRec &rec, double kc0, double kc1,
                                                      double dx[6], double ts, double c2, double c3) {
  double *Vi = w.V + i * V_STRIDE;
  double du0 = 0.0, du1 = 0.0;
  if (has_input) {
    // two partial chains per row of G dx
    const double t0 = (rec.G[0] * dx[0] + rec.G[1] * dx[1] + rec.G[2] * dx[2]) + (rec.G[3] * dx[3] + rec.G[4] * dx[4] + rec.G[5] * dx[5]);
    const double t1 = (rec.G[6] * dx[0] + rec.G[7] * dx[1] + rec.G[8] * dx[2]) + (rec.G[9] * dx[3] + rec.G[10] * dx[4] + rec.G[11] * dx[5]);
    du0 = kc0 - (rec.F[0] * t0 + rec.F[1] * t1);
    du1 = kc1 - (rec.F[1] * t0 + rec.F[2] * t1);
  }
  {  // lane t < 8 stores component t (one predicated store instead of eight branches)
    double val = dx[0];
    val = (w.lane == 1) ? dx[1] : val; val = (w.lane == 2) ? dx[2] : val; val = (w.lane == 3) ? dx[3] : val;
    val = (w.lane == 4) ? dx[4] : val; val = (w.lane == 5) ? dx[5] : val;
    val = (w.lane == 6) ? du0 : val;   val = (w.lane == 7) ? du1 : val;
    if (w.lane < 8) Vi[dst + w.lane] = val;
  }
#pragma unroll
  for (int ax = 0; ax < 2; ++ax) {
    const double P = dx[3 * ax], Vv = dx[3 * ax + 1], A = dx[3 * ax + 2], U = ax ? du1 : du0;
    dx[3 * ax] = P + ts * Vv + c2 * A + c3 * U;
    dx[3 * ax + 1] = Vv + ts * A + c2 * U;
    dx[3 * ax + 2] = A + ts * U;
  }
}
__device__ __forceinline__ void riccati_forward(const WarpCtx &w, int dst) {
  const DevProb &p = *w.p;
  const int N = w.N;
  const double ts = p.ts, c2 = p.c2, c3 = p.c3;
  double dx[6] = {0, 0, 0, 0, 0, 0};
  SweepRec ra, rb;
  double ka0, ka1, kb0, kb1;
  ra.load(w.S);
  ka0 = w.V[V_K]; ka1 = w.V[V_K + 1];
  for (int i = 0; i < N; i += 2) {
    // stage N-1 has no input: its record is never used, any valid stage may be fetched in its place
    const int i1 = (i + 1 < N - 1) ? i + 1 : 0, i2 = (i + 2 < N - 1) ? i + 2 : 0;
    rb.load(w.S + i1 * S_STRIDE);
    kb0 = w.V[i1 * V_STRIDE + V_K]; kb1 = w.V[i1 * V_STRIDE + V_K + 1];
    riccati_forward_stage(w, i, dst, i < N - 1, ra, ka0, ka1, dx, ts, c2, c3);
    if (i + 1 >= N) break;
    ra.load(w.S + i2 * S_STRIDE);
    ka0 = w.V[i2 * V_STRIDE + V_K]; ka1 = w.V[i2 * V_STRIDE + V_K + 1];
    riccati_forward_stage(w, i + 1, dst, i + 1 < N - 1, rb, kb0, kb1, dx, ts, c2, c3);
  }
  __syncwarp();
}

struct QpResult {
  int status;     // 0 optimal (or, with converged == 0, merely a feasible point), 1 infeasible (Farkas certificate verified or an empty
                  // box), 3 parked (iteration budget used up), 4 unknown (no convergence, no feasible point, no certificate)
  int susp_index; // status 3: slot of the saved state
  int converged;  // 1: obj is the optimum of the relaxation (a valid bound); 0: the iteration stalled, obj is only an upper bound
  int iters;
  double obj;     // without soft-decision penalties
  double lb;      // valid lower bound of the relaxation: obj if converged, else the Lagrangian bound of the last multipliers
  long rows;      // active rows x iterations (work counter)
#ifdef MQ_PROF
  long long c_rows, c_factor, c_sweeps;   // clock64 cycles: row passes / Riccati factorisation / vector + forward sweeps
  long long c_a, c_ared, c_d, c_e, c_g, c_atr;   // split of c_rows: pass A visit, pass A reduce + epilogue + team reduce, D, E, G
#endif
};

// Solves the node QP of w.dec; called by every thread of the team with identical arguments.  zwarm (may be
// null) = relaxed optimum of the parent node, [N][8].  On
// success V_Z holds the optimal stage vectors.  All control decisions derive from team_reduce results,
// which are bitwise identical in every thread, so the barriers inside stay uniform.
// Parked relaxations: state = header (8 doubles: iteration, stall count, alpha, sigma mu, pending, dual residual, cn, -) +
// the stage vectors V + the (s, lambda) records.
constexpr int SUSP_HDR = 8;
struct SuspendIO {
  const double *resume;   // state to continue from (null: fresh solve)
  double *pool;           // pool of this round (null: never park)
  int *counter; int nslots; long stride; int budget;
};

// Dual information of the current iterate (z, lambda); called by every thread of the team, V_DZ is used as scratch.
//
// With g_i = kappa (Q_i z_i + c_i) + G_i' lambda and the costate recursion pi_i = g_x,i + A' pi_{i+1}, every trajectory
// z' of the dynamics satisfies  sum_i g_i' z'_i = pi_0' x_0 + sum_i rho_i' u'_i,  rho_i = g_u,i + B' pi_{i+1},
// and the jerks are confined to the global box [total_min_jerk, total_max_jerk] (model_region_constraints.mod:36-39).
//
// farkas (kappa = 0, lambda scaled by 1 / lmax): a feasible z' has lambda'(G z' - h) <= 0, i.e. sum_i g_i' z'_i <= h' lambda;
//   if pi_0' x_0 - sum |rho_i| U > h' lambda no such z' exists: returns true, the node is PROVEN infeasible.
// else (kappa = 1): Lagrangian bound  f(z') >= f(z) - lambda'(h - G z) - sum_i |rho_i| range_i  for every feasible z'
//   (L(., lambda) is convex, its gradient along the dynamics is rho, |u'_i - u_i| <= range_i), stored in *lb.
// (out of line and with its own copy of the context: it runs once per stalled or infeasible relaxation, and its row pass would
// otherwise sit in the middle of the iteration loop's instruction stream)
__device__ __noinline__ bool dual_check(const WarpCtx w, const StepCtx sc, bool farkas, double lmax, double *lb) {
  const DevProb &p = *w.p;
  const double *D = w.D;
  const int N = w.N;
  const bool lead = (w.g == 0);
  double hl = 0.0, lgz = 0.0, fz = 0.0, mag = 0.0;
  MQ_FOR_STAGES(i, act) {
    PassDual v; v.io.init(w.rows, w.NP, act ? i : 0); v.sc = sc; v.scale = farkas ? 1.0 / lmax : 1.0;
    v.hl = 0.0; v.lgz = 0.0; v.mag = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) v.gl[t] = 0.0;
    if (act) {
      const double *Vi = w.V + i * V_STRIDE;
#pragma unroll
      for (int t = 0; t < 8; ++t) { v.y[t] = Vi[V_Z + t]; v.d[t] = Vi[V_DZ + t]; v.da[t] = Vi[V_DZA + t]; }
      visit_rows(w, i, v);
    }
    sg_reduce(w, v.gl);
    double t3[3] = {v.hl, v.lgz, v.mag};
    sg_reduce(w, t3);
    if (act && lead) {
      double *Vi = w.V + i * V_STRIDE;
      const double *cst = D + p.o_cost + 16 * i;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        Vi[V_DZ + t] = v.gl[t] + (farkas ? 0.0 : cst[t] * v.y[t] + cst[8 + t]);
        if (!farkas) fz += (0.5 * cst[t] * v.y[t] + cst[8 + t]) * v.y[t];
      }
      hl += t3[0]; lgz += t3[1]; mag += t3[2];
    }
  }
  { double m0 = 0.0, m1 = 0.0; team_reduce(w, m0, m1, hl, lgz); }
  { double m0 = 0.0, m1 = 0.0; team_reduce(w, m0, m1, fz, mag); }   // (ends with a barrier: V_DZ is complete)
  if (w.wid == 0) {   // costate recursion, all lanes redundantly
    const double ts = p.ts, c2 = p.c2, c3 = p.c3;
    const double Ulo = p.total_min_jerk, Uhi = p.total_max_jerk, Uabs = fmax(fabs(Ulo), fabs(Uhi));
    double pn[6], usum = 0.0;
#pragma unroll
    for (int t = 0; t < 6; ++t) pn[t] = w.V[(N - 1) * V_STRIDE + V_DZ + t];
    for (int i = N - 2; i >= 0; --i) {
      const double *g = w.V + i * V_STRIDE + V_DZ, *z = w.V + i * V_STRIDE + V_Z;
#pragma unroll
      for (int ax = 0; ax < 2; ++ax) {
        const double pp = pn[3 * ax], pv = pn[3 * ax + 1], pa = pn[3 * ax + 2];
        pn[3 * ax] = g[3 * ax] + pp;
        pn[3 * ax + 1] = g[3 * ax + 1] + ts * pp + pv;
        pn[3 * ax + 2] = g[3 * ax + 2] + c2 * pp + ts * pv + pa;
        const double rho = g[6 + ax] + c3 * pp + c2 * pv + ts * pa;
        const double range = farkas ? Uabs : fmax(fmax(Uhi - z[6 + ax], z[6 + ax] - Ulo), 0.0);
        usum += fabs(rho) * range;
      }
    }
    double px = 0.0;
#pragma unroll
    for (int t = 0; t < 6; ++t) px += pn[t] * D[p.o_x0 + t];
    if (w.lane == 0) { w.red[0] = px; w.red[1] = usum; }
  }
  __syncthreads();
  const double pi0x0 = w.red[0], usum = w.red[1];
  __syncthreads();
  if (farkas) return (hl - pi0x0 + usum) < -1e-10 * (mag + fabs(pi0x0) + usum) - 1e-13;
  *lb = fz + p.cost_const - (hl - lgz) - usum - 1e-12 * (fabs(fz) + fabs(hl) + fabs(lgz) + usum);
  return false;
}

__device__ __forceinline__ QpResult solve_node_qp(const WarpCtx &w, const PhiEntry &e1, const PhiEntry &e2, const double *zwarm, double warm_mu,
                                                  const SuspendIO &sio) {
  const DevProb &p = *w.p;
  const double *D = w.D;
  const int N = w.N;
  const bool lead = (w.g == 0);
  QpResult res; res.status = 1; res.converged = 0; res.iters = 0; res.obj = 0.0; res.lb = -MQ_INF; res.rows = 0; res.susp_index = -1;
#ifdef MQ_PROF
  res.c_rows = res.c_factor = res.c_sweeps = 0; res.c_a = res.c_ared = res.c_d = res.c_e = res.c_g = res.c_atr = 0;
  long long qc0 = 0;
#define MQ_T0 qc0 = clock64();
#define MQ_T1(field) res.field += clock64() - qc0;
  long long pc0 = clock64(), pc1;
#define MQ_TICK(field) { pc1 = clock64(); res.field += pc1 - pc0; pc0 = pc1; }
#else
#define MQ_TICK(field)
#define MQ_T0
#define MQ_T1(field)
#endif

  const int nV = (N * V_STRIDE + 1) & ~1, nRows = (w.kmax + 1) * N;   // stage vectors (doubles), (s, lambda) records (double2)
  double cn = 0.0, rdn = 0.0;
  StepCtx sc; sc.alpha = 0.0; sc.sigmu = 0.0; sc.pending = false;
  int stall = 0, it0 = 0;
  if (sio.resume) {
    // continue a parked relaxation: stage vectors, (s, lambda) records and the scalars of the iteration
    const double *hdr = sio.resume;
    it0 = (int)hdr[0]; stall = (int)hdr[1]; sc.alpha = hdr[2]; sc.sigmu = hdr[3]; sc.pending = hdr[4] != 0.0; rdn = hdr[5]; cn = hdr[6];
    for (int k = threadIdx.x; k < N * V_STRIDE; k += blockDim.x) w.V[k] = hdr[SUSP_HDR + k];
    // records are parked as [slot][N]: the padded stride NP depends on the team size, which may differ between the rounds
    const double2 *src = reinterpret_cast<const double2 *>(hdr + SUSP_HDR + nV);
    for (int k = threadIdx.x; k < nRows; k += blockDim.x) { const int sl = k / N; w.rows[sl * w.NP + (k - sl * N)] = src[k]; }
    __syncthreads();
  } else {
  // trivially infeasible boxes (build_node_qp of the oracle); cn = largest linear cost coefficient
  double bad = 0.0, z0 = 0.0, z1 = 0.0;
  MQ_FOR_STAGES(i, act) {
    if (!act || !lead) continue;
    const unsigned char m = (i > 0) ? w.dec[p.off_mode + i] : (unsigned char)0;
    double lo[8], hi[8];
    stage_bounds(w, i, w.jeff[i], i > 0 && m == MODE_FROZEN, lo, hi);
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      if (t == Y_PY) continue;
      if (i == 0 && t < 6) continue;
      if (lo[t] > hi[t] + 1e-12) bad = 1.0;
      if (i == N - 1 && t >= 6 && (lo[t] > 1e-9 || hi[t] < -1e-9)) bad = 1.0;
    }
    const double *cst = D + p.o_cost + 16 * i;
#pragma unroll
    for (int t8 = 0; t8 < 8; ++t8) cn = fmax(cn, fabs(cst[8 + t8]));
  }
  team_reduce(w, bad, cn, z0, z1);
  if (bad > 0.0) return res;

  // start: zero jerk (free response), s = max(h - g.z, 1), lambda = 1
  const double x0[6] = {D[p.o_x0], D[p.o_x0 + 1], D[p.o_x0 + 2], D[p.o_x0 + 3], D[p.o_x0 + 4], D[p.o_x0 + 5]};
  MQ_FOR_STAGES(i, act) {
    PassInit v; v.io.init(w.rows, w.NP, act ? i : 0);
#pragma unroll
    for (int t = 0; t < 8; ++t) v.gl[t] = 0.0;
    v.mu0 = zwarm ? warm_mu : 0.0; v.smin = zwarm ? sqrt(warm_mu) : 1.0;
    if (act) {
      if (zwarm) {   // the parent's relaxed optimum (satisfies the dynamics, violates the rows this node adds)
#pragma unroll
        for (int t = 0; t < 8; ++t) v.y[t] = zwarm[i * 8 + t];
      } else {
        const double t = i * p.ts;
#pragma unroll
        for (int ax = 0; ax < 2; ++ax) {
          const double P = x0[3 * ax], Vv = x0[3 * ax + 1], A = x0[3 * ax + 2];
          v.y[3 * ax] = P + t * Vv + 0.5 * t * t * A;
          v.y[3 * ax + 1] = Vv + t * A;
          v.y[3 * ax + 2] = A;
        }
        v.y[6] = 0.0; v.y[7] = 0.0;
      }
      visit_rows(w, i, v);
    }
    sg_reduce(w, v.gl);
    if (act && lead) {
      double *Vi = w.V + i * V_STRIDE;
      const double *cst = D + p.o_cost + 16 * i;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        Vi[V_Z + t] = v.y[t]; Vi[V_DZ + t] = 0.0; Vi[V_DZA + t] = 0.0;
        // dual residual at the start (all dynamics multipliers zero); it contracts by (1 - alpha)
        // with every Newton step afterwards
        if ((t < 6 && i > 0) || (t >= 6 && i < N - 1)) rdn = fmax(rdn, fabs(cst[t] * v.y[t] + cst[8 + t] + v.gl[t]));
      }
    }
  }
  { double q0 = 0.0, q1 = 0.0, q2 = 0.0; team_reduce(w, rdn, q0, q1, q2); }
  }   // fresh solve

  int status = 2, it = 0, it_start = it0, next_check = 3;
  double lmax_last = 0.0;
  bool may_park = (sio.pool != nullptr);
  for (;;) {   // interior-point iterations, interrupted when the multipliers are to be tested as a Farkas certificate (status 5)
  status = 2;
  for (it = it_start; it < 100; ++it) {
    if (may_park && it - it0 >= sio.budget) {
      // iteration budget of this round used up: park the relaxation (it continues in the next round) if a slot is left
      if (threadIdx.x == 0) w.red[0] = (double)atomicAdd(sio.counter, 1);
      __syncthreads();
      const int ss = (int)w.red[0];
      __syncthreads();
      if (ss < sio.nslots) {
        double *hdr = sio.pool + (long)ss * sio.stride;
        if (threadIdx.x == 0) { hdr[0] = it; hdr[1] = stall; hdr[2] = sc.alpha; hdr[3] = sc.sigmu; hdr[4] = sc.pending ? 1.0 : 0.0; hdr[5] = rdn; hdr[6] = cn; hdr[7] = 0.0; }
        for (int k = threadIdx.x; k < N * V_STRIDE; k += blockDim.x) hdr[SUSP_HDR + k] = w.V[k];
        double2 *dst = reinterpret_cast<double2 *>(hdr + SUSP_HDR + nV);
        for (int k = threadIdx.x; k < nRows; k += blockDim.x) { const int sl = k / N; dst[k] = w.rows[sl * w.NP + (k - sl * N)]; }
        res.status = 3; res.susp_index = ss; res.iters = it - it0;
        return res;
      }
      may_park = false;   // pool exhausted: run to the end
    }
    // ---- pass A ----
    double rpn = 0.0, musum = 0.0, lmax = 0.0, mcount = 0.0;
    MQ_FOR_STAGES(i, act) {
      PassA v; v.io.init(w.rows, w.NP, act ? i : 0); v.sc = sc;
      v.rpn = 0.0; v.musum = 0.0; v.lmax = 0.0; v.m = 0;
#pragma unroll
      for (int t = 0; t < 21; ++t) v.H[t] = 0.0;
      v.Huu[0] = v.Huu[1] = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) v.gx[t] = 0.0;
      if (act) {
        const double *Vi = w.V + i * V_STRIDE;
#pragma unroll
        for (int t = 0; t < 8; ++t) { v.y[t] = Vi[V_Z + t]; v.d[t] = Vi[V_DZ + t]; v.da[t] = Vi[V_DZA + t]; }
        MQ_T0
        visit_rows(w, i, v);
        MQ_T1(c_a)
      }
      MQ_T0
      sg_reduce(w, v.H); sg_reduce(w, v.Huu); sg_reduce(w, v.gx);
      if (act && lead) {
        double *Vi = w.V + i * V_STRIDE, *Si = w.S + i * S_STRIDE;
        const double *cst = D + p.o_cost + 16 * i;
#pragma unroll
        for (int t = 0; t < 6; ++t) v.H[t * (t + 1) / 2 + t] += cst[t];
#pragma unroll
        for (int t = 0; t < 21; ++t) Si[t] = v.H[t];
        Si[S_MUU] = v.Huu[0] + cst[6] + 1e-10; Si[S_MUU + 1] = v.Huu[1] + cst[7] + 1e-10;
#pragma unroll
        for (int t = 0; t < 8; ++t) Vi[V_G + t] = cst[t] * v.y[t] + cst[8 + t] + v.gx[t];
      }
      rpn = fmax(rpn, v.rpn); musum += v.musum; lmax = fmax(lmax, v.lmax); mcount += (double)v.m;
    }
    sc.pending = false;
    MQ_T1(c_ared)
    MQ_T0
    team_reduce(w, rpn, lmax, musum, mcount);
    MQ_T1(c_atr)
    const long m = (long)mcount;
    res.rows += m;
    const double mu = (m > 0) ? musum / m : 0.0;
    if (rpn <= 1e-9 && rdn <= 1e-8 * (1.0 + cn) && mu <= 1e-10) { status = 0; break; }
    if (lmax > 1e13) { status = 2; break; }   // diverged: infeasible if the multipliers certify it (below)
    // early exit of infeasible relaxations: once the multipliers have grown, test them as a Farkas certificate (below; the
    // iteration resumes here, with this pass A repeated, if they are not one yet)
    lmax_last = lmax;
    if (it >= next_check && rpn > 1e-5 && lmax > 1.0 + cn) { status = 5; break; }
    // ---- predictor ----
    MQ_TICK(c_rows)
    if (w.wid == 0) riccati_factor(w, e1, e2);
    MQ_TICK(c_factor)
    if (w.wid == 0) riccati_forward(w, V_DZA);
    __syncthreads();
    MQ_TICK(c_sweeps)
    double rmax = 1.0, s1 = 0.0, s2 = 0.0, dummy = 0.0;
    MQ_T0
    MQ_FOR_STAGES(i, act) {
      if (!act) continue;
      const double *Vi = w.V + i * V_STRIDE;
      PassD v; v.io.init(w.rows, w.NP, i); v.rmax = 1.0; v.s1 = 0.0; v.s2 = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) { v.y[t] = Vi[V_Z + t]; v.da[t] = Vi[V_DZA + t]; }
      visit_rows(w, i, v);
      rmax = fmax(rmax, v.rmax); s1 += v.s1; s2 += v.s2;
    }
    team_reduce(w, rmax, dummy, s1, s2);
    MQ_T1(c_d)
    double amin = 1.0 / rmax;
    double sigma = 0.0;
    if (m > 0 && mu > 0.0) {
      const double mu_aff = (musum + amin * s1 + amin * amin * s2) / m;
      const double r = mu_aff / mu;
      sigma = r * r * r;
      if (sigma > 1.0) sigma = 1.0;
      if (!(sigma >= 0.0)) sigma = 0.0;
    }
    const double sigmu = sigma * mu;
    // ---- corrector ----
    MQ_T0
    MQ_FOR_STAGES(i, act) {
      PassE v; v.io.init(w.rows, w.NP, act ? i : 0); v.sigmu = sigmu;
#pragma unroll
      for (int t = 0; t < 8; ++t) v.gx[t] = 0.0;
      if (act) {
        const double *Vi = w.V + i * V_STRIDE;
#pragma unroll
        for (int t = 0; t < 8; ++t) { v.y[t] = Vi[V_Z + t]; v.da[t] = Vi[V_DZA + t]; }
        visit_rows(w, i, v);
      }
      sg_reduce(w, v.gx);
      if (act && lead) {
        double *Vi = w.V + i * V_STRIDE;
        const double *cst = D + p.o_cost + 16 * i;
#pragma unroll
        for (int t = 0; t < 8; ++t) Vi[V_G + t] = cst[t] * v.y[t] + cst[8 + t] + v.gx[t];
      }
    }
    __syncthreads();
    MQ_T1(c_e)
    MQ_TICK(c_rows)
    if (w.wid == 0) { riccati_vector(w); riccati_forward(w, V_DZ); }
    __syncthreads();
    MQ_TICK(c_sweeps)
    rmax = 0.0;
    double bad_step = 0.0; s1 = 0.0; s2 = 0.0;
    MQ_T0
    MQ_FOR_STAGES(i, act) {
      if (!act) continue;
      const double *Vi = w.V + i * V_STRIDE;
      PassG v; v.io.init(w.rows, w.NP, i); v.sigmu = sigmu; v.rmax = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) { v.y[t] = Vi[V_Z + t]; v.d[t] = Vi[V_DZ + t]; v.da[t] = Vi[V_DZA + t]; }
      if (!(fabs(v.d[Y_UX]) + fabs(v.d[Y_UY]) + fabs(v.d[Y_PX]) + fabs(v.d[Y_PY]) < 1e300)) bad_step = 1.0;  // NaN / inf step
      visit_rows(w, i, v);
      rmax = fmax(rmax, v.rmax);
    }
    team_reduce(w, rmax, bad_step, s1, s2);
    MQ_T1(c_g)
    if (bad_step > 0.0) { status = 2; break; }  // singular stage system
    // fraction to the boundary: 0.995 far from the solution, -> 1 as mu -> 0 (superlinear end phase)
    const double tau = (w.tau_k > 0.0) ? fmin(fmax(0.995, 1.0 - w.tau_k * mu), 0.9999) : 0.995;
    double alpha = (rmax > tau) ? tau / rmax : 1.0;
    sc.alpha = alpha; sc.sigmu = sigmu; sc.pending = true;
#ifdef MQ_PROF
    if (w.dbgrow && threadIdx.x == 0 && it < 100) { double *q = w.dbgrow + 8 + 5 * it; q[0] = alpha; q[1] = mu; q[2] = rpn; q[3] = sigma; q[4] = lmax; }
#endif
    rdn *= (1.0 - alpha);
    MQ_FOR_STAGES(i, act) {   // z += alpha dz (rows are updated lazily by the next pass A)
      if (!act || !lead) continue;
      double *Vi = w.V + i * V_STRIDE;
#pragma unroll
      for (int t = 0; t < 8; ++t) Vi[V_Z + t] += alpha * Vi[V_DZ + t];
    }
    __syncthreads();
    if (alpha < 1e-6) { if (++stall >= 5) { status = 2; break; } } else stall = 0;
  }
  res.iters = it - it0;
  if (status == 0) { res.converged = 1; break; }
  MQ_TICK(c_rows)
  bool farkas = true; StepCtx dsc; dsc.alpha = 0.0; dsc.sigmu = 0.0; dsc.pending = false;
  double dlm = lmax_last;
  if (status == 2) {
    // not converged.  A primal feasible point is still usable (upper bound + Lagrangian lower bound); a violated one
    // closes the node only with a Farkas certificate, otherwise the outcome is "unknown" and the caller keeps the
    // node's bound in the books.
    double worst = 0.0, lmx = 0.0, nanflag = 0.0, q1 = 0.0;
    MQ_FOR_STAGES(i, act) {
      if (!act) continue;
      PassViol v; v.io.init(w.rows, w.NP, i); v.worst = 0.0; v.lmax = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) v.y[t] = w.V[i * V_STRIDE + V_Z + t];
      visit_rows(w, i, v);
      if (!(v.worst == v.worst) || !(v.lmax == v.lmax)) nanflag = 1.0;
      worst = fmax(worst, v.worst); lmx = fmax(lmx, v.lmax);
    }
    team_reduce(w, worst, lmx, nanflag, q1);
    if (nanflag > 0.0 || !(lmx < MQ_INF) || (worst > 1e-7 && !(lmx > 0.0))) { status = 4; break; }
    farkas = (worst > 1e-7); dsc = sc; dlm = farkas ? lmx : 1.0;
  }
  double lbv = -MQ_INF;
  const bool cert = dual_check(w, dsc, farkas, dlm, &lbv);
  if (status == 5) {
    if (cert) { status = 1; break; }
    it_start = it; next_check = it + 3;   // not a certificate yet: carry on
    continue;
  }
  if (farkas) status = cert ? 1 : 4; else { status = 0; res.lb = lbv; }
  break;
  }   // for (;;)
  res.status = status;
  if (status == 0) {
    double o = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0;
    MQ_FOR_STAGES(i, act) {
      if (!act || !lead) continue;
      const double *cst = D + p.o_cost + 16 * i;
      const double *z = w.V + i * V_STRIDE + V_Z;
#pragma unroll
      for (int t = 0; t < 8; ++t) o += (0.5 * cst[t] * z[t] + cst[8 + t]) * z[t];
    }
    team_reduce(w, q0, q1, o, q2);
    res.obj = o + p.cost_const;
    if (res.converged) res.lb = res.obj;
  }
  return res;
}

}  // namespace miqp
