// formulation_tables.cuh -- derived tables of one plan (unit edge normals, normalised per-region
// rows, front-axle maps, stage costs) and the row / non-zero index prefixes of the big-M model.
// Body of prepare_tables_kernel (formulation.cu), written as two host/device functions so that
// the CPU tests of the host logic can prepare the same tables (never used by the product on
// the host).  Compiled without FMA contraction (bit-exact with the oracle).
#pragma once
#include "dev_problem.cuh"

namespace miqp {

// parallel part: work items strided by (tid, nt)
__host__ __device__ inline void prepare_tables_parallel(DevProb &p, double *D, int *I, int tid, int nt) {
  const int N = p.N, R = p.R, C = p.C, O = p.O, L = p.L;
  // unit edge normals: cross(P)/len = ex*Y - ey*X - ec
  for (int e = tid; e < p.nEnvEdges; e += nt) {
    const double *g = D + p.o_env_edges + 4 * e;
    double dx = g[2] - g[0], dy = g[3] - g[1];
    double len = sqrt(dx * dx + dy * dy);
    if (len > 0.0) { dx /= len; dy /= len; }
    double *t = D + p.o_envtab + 3 * e;
    t[0] = dx; t[1] = dy; t[2] = dx * g[1] - g[0] * dy;
  }
  for (int e = tid; e < O * N * L; e += nt) {
    const double *g = D + p.o_obs_edges + 4 * e;
    double dx = g[2] - g[0], dy = g[3] - g[1];
    double len = sqrt(dx * dx + dy * dy);
    if (len > 0.0) { dx /= len; dy /= len; }
    double *t = D + p.o_obstab + 3 * e;
    t[0] = dx; t[1] = dy; t[2] = dx * g[1] - g[0] * dy;
  }
  // per-region rows of a decided region (model_region_constraints.mod:53-54, :97-104)
  // as  a_vx*vx + a_ax*ax + a_vy*vy + a_ay*ay <= rhs , normalised to unit coefficient norm
  for (int j = tid; j < R; j += nt) {
    const double *f = D + p.o_frac + 4 * j;
    const double *KX = D + p.o_poly[4] + 3 * j, *KN = D + p.o_poly[5] + 3 * j;
    double *t = D + p.o_modetab + 20 * j;
    double sl = (f[1] + f[3]) / (f[0] + f[2]);
    double rows[4][5] = {
        {f[1], 0.0, -f[0], 0.0, 0.0},           // f1*vy >= f2*vx
        {-f[3], 0.0, f[2], 0.0, 0.0},           // f3*vy <= f4*vx
        {-KX[1], -sl, -KX[2], 1.0, KX[0]},      // ay <= KX.[1,vx,vy] + sl*ax
        {KN[1], sl, KN[2], -1.0, -KN[0]}};      // ay >= KN.[1,vx,vy] + sl*ax
    for (int k = 0; k < 4; ++k) {
      double n = 0.0;
      for (int a = 0; a < 4; ++a) n += rows[k][a] * rows[k][a];
      n = (n > 0.0) ? 1.0 / sqrt(n) : 1.0;
      for (int a = 0; a < 5; ++a) t[5 * k + a] = rows[k][a] * n;
    }
  }
  // front axle maps X_front = px + fx[0] + fx[1]*vx + fx[2]*vy  (model_region_constraints.mod:57-70)
  for (int q = tid; q < C * R; q += nt) {
    int c = q / R, j = q % R;
    double wb = D[p.o_wb + c];
    double *t = D + p.o_fronttab + 12 * q;
    for (int a = 0; a < 3; ++a) {
      t[a] = wb * D[p.o_poly[2] + 3 * j + a];      // x UB: POLY_COSS_UB
      t[3 + a] = wb * D[p.o_poly[3] + 3 * j + a];  // x LB: POLY_COSS_LB
      t[6 + a] = wb * D[p.o_poly[0] + 3 * j + a];  // y UB: POLY_SINT_UB
      t[9 + a] = wb * D[p.o_poly[1] + 3 * j + a];  // y LB: POLY_SINT_LB
    }
  }
  // stage costs (objective_function.mod:7-19): w (y-ref)^2 = 1/2 (2w) y^2 - 2 w ref y + w ref^2
  for (int q = tid; q < C * N; q += nt) {
    int c = q / N, i = q % N;
    double *t = D + p.o_cost + 16 * q;
    const int wmap[8] = {0, 1, 2, 3, 4, 5, 6, 7};  // px,vx,ax,py,vy,ay,ux,uy -> o_w index
    double ref[8] = {D[p.o_ref[0] + c * N + i], D[p.o_ref[1] + c * N + i], 0.0,
                     D[p.o_ref[2] + c * N + i], D[p.o_ref[3] + c * N + i], 0.0, 0.0, 0.0};
    for (int a = 0; a < 8; ++a) {
      double w = D[p.o_w[wmap[a]] + c];
      t[a] = 2.0 * w;
      t[8 + a] = -2.0 * w * ref[a];
    }
  }
  // Reach tables of the cost-only LQ problem, per car and axis (triple integrator, weights
  // q = 2w on (p, v, a), r = 2 w_jerk): W_i = covariance of the state at step i under the Gaussian
  // exp(-J), J = 1/2 sum (x'Qx + r u^2).  For any direction g at step i
  //     min { 1/2 d'Q d : d dynamics-consistent, g.d_i <= -v } = 1/2 v^2 / (g' W_i g),
  // the objective increase every trajectory pays for moving by v against g at step i; the branch
  // and bound uses it as a lower bound of a child before the child is solved (bnb.cu).
  // Backward Riccati of the cost (P_{N-1} = Q; F_k = r + B'P_{k+1}B; K_k = B'P_{k+1}A / F_k),
  // then forward W_{k+1} = (A - B K_k) W_k (A - B K_k)' + B B' / F_k, W_0 = 0.
  for (int q = tid; q < 2 * C; q += nt) {
    const int c = q / 2, ax = q % 2;
    const double ts = p.ts, c2 = p.c2, c3 = p.c3;
    const double Q[3] = {2.0 * D[p.o_w[3 * ax] + c], 2.0 * D[p.o_w[3 * ax + 1] + c], 2.0 * D[p.o_w[3 * ax + 2] + c]};
    const double r = 2.0 * D[p.o_w[6 + ax] + c] + 1e-10;
    const double A[3][3] = {{1.0, ts, c2}, {0.0, 1.0, ts}, {0.0, 0.0, 1.0}};
    const double B[3] = {c3, c2, ts};
    double P[3][3] = {{Q[0], 0, 0}, {0, Q[1], 0}, {0, 0, Q[2]}};
    // the gains are parked in the table slots of their stage: K_k (3) and 1/F_k at offsets 0..3
    for (int k = N - 2; k >= 0; --k) {
      double PB[3], PA[3][3];
      for (int a = 0; a < 3; ++a) { PB[a] = 0.0; for (int b = 0; b < 3; ++b) PB[a] += P[a][b] * B[b]; }
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { PA[a][b] = 0.0; for (int e = 0; e < 3; ++e) PA[a][b] += P[a][e] * A[e][b]; }
      double F = r; for (int a = 0; a < 3; ++a) F += B[a] * PB[a];
      double K[3]; for (int b = 0; b < 3; ++b) { K[b] = 0.0; for (int a = 0; a < 3; ++a) K[b] += B[a] * PA[a][b]; K[b] /= F; }
      double Pn[3][3];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        double v = (a == b ? Q[a] : 0.0);
        for (int e = 0; e < 3; ++e) v += A[e][a] * PA[e][b];
        Pn[a][b] = v - K[a] * F * K[b];
      }
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) P[a][b] = 0.5 * (Pn[a][b] + Pn[b][a]);
      double *t = D + p.o_wtab + 14 * (c * N + k) + 7 * ax;
      t[0] = K[0]; t[1] = K[1]; t[2] = K[2]; t[3] = 1.0 / F;
    }
    double W[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < N; ++k) {
      double *t = D + p.o_wtab + 14 * (c * N + k) + 7 * ax;
      double K[3] = {0, 0, 0}, iF = 0.0;
      if (k < N - 1) { K[0] = t[0]; K[1] = t[1]; K[2] = t[2]; iF = t[3]; }
      // table of stage k: W_k packed lower (00,10,11,20,21,22) and var(u_k) = K W K' + 1/F
      double vu = iF;
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) vu += K[a] * W[a][b] * K[b];
      t[0] = W[0][0]; t[1] = W[1][0]; t[2] = W[1][1]; t[3] = W[2][0]; t[4] = W[2][1]; t[5] = W[2][2];
      t[6] = (k < N - 1) ? vu : 0.0;
      if (k < N - 1) {
        double Ac[3][3], T[3][3], Wn[3][3];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Ac[a][b] = A[a][b] - B[a] * K[b];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { T[a][b] = 0.0; for (int e = 0; e < 3; ++e) T[a][b] += Ac[a][e] * W[e][b]; }
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { double v = B[a] * B[b] * iF; for (int e = 0; e < 3; ++e) v += T[a][e] * Ac[b][e]; Wn[a][b] = v; }
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) W[a][b] = 0.5 * (Wn[a][b] + Wn[b][a]);
      }
    }
  }
  // prefix of possible regions per car
  for (int c = tid; c < C; c += nt) {
    int n = 0;
    for (int j = 0; j < R; ++j) { I[p.o_posspre + c * (R + 1) + j] = n; n += (I[p.o_possible + c * R + j] == 1); }
    I[p.o_posspre + c * (R + 1) + R] = n;
  }
}

// serial part (one thread, after a barrier)
__host__ __device__ inline void prepare_tables_serial(DevProb &p, double *D, int *I) {
  const int N = p.N, R = p.R, C = p.C, O = p.O, E = p.E;
    double cc = 0.0;
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < N; ++i)
        for (int a = 0; a < 4; ++a) {
          const int widx[4] = {0, 1, 3, 4};
          double w = D[p.o_w[widx[a]] + c], r = D[p.o_ref[a] + c * N + i];
          cc += w * r * r;
        }
    p.cost_const = cc;
    // obstacle row prefixes
    long srow = 0, snnz = 0;
    for (int i = 0; i < N; ++i) {
      int rr = 0, nn = 0;
      for (int o = 0; o < O; ++o) {
        I[p.o_obsrowpre + i * (O + 1) + o] = rr;
        I[p.o_obsnnzpre + i * (O + 1) + o] = nn;
        int ne = I[p.o_obs_nedges + o * N + i];
        int soft = (I[p.o_obs_soft + o] == 1);
        rr += 5 * ne + 5;
        nn += 15 * ne + 5 * (ne + soft);
      }
      I[p.o_obsrowpre + i * (O + 1) + O] = rr;
      I[p.o_obsnnzpre + i * (O + 1) + O] = nn;
      I[p.o_obsstep_rows + i] = (int)srow;
      I[p.o_obsstep_nnz + i] = (int)snnz;
      srow += (long)C * rr;
      snnz += (long)C * nn;
    }
    I[p.o_obsstep_rows + N] = (int)srow;
    I[p.o_obsstep_nnz + N] = (int)snnz;
    // region rows per car of one step
    long rr = 0, nn = 0;
    for (int c = 0; c < C; ++c) {
      p.region_rows_car[c] = rr; p.region_nnz_car[c] = nn;
      int rp = I[p.o_posspre + c * (R + 1) + R];
      rr += 20L * rp + (R - rp) + 1;
      nn += 76L * rp + (R - rp) + R;
    }
    p.region_rows_car[C] = rr; p.region_nnz_car[C] = nn;
    // family bases
    long fr[NUM_FAM], fn[NUM_FAM];
    fr[FAM_IC1] = 12L * C;                fn[FAM_IC1] = 12L * C;
    fr[FAM_IC2] = 5L * R * C;             fn[FAM_IC2] = 9L * R * C;
    fr[FAM_IC3] = 5L * C;                 fn[FAM_IC3] = 5L * C;
    fr[FAM_DYN] = 6L * C * (N - 1);       fn[FAM_DYN] = 24L * C * (N - 1);
    fr[FAM_BOX] = 12L * C * N;            fn[FAM_BOX] = 12L * C * N;
    fr[FAM_REGION] = rr * (N - 1);        fn[FAM_REGION] = nn * (N - 1);
    fr[FAM_MINSPEED] = 15L * R * C * (N - 1); fn[FAM_MINSPEED] = 35L * R * C * (N - 1);
    fr[FAM_ENV] = (E > 0) ? (long)C * N * (5L * p.nEnvEdges + 5) : 0;
    fn[FAM_ENV] = (E > 0) ? (long)C * N * (15L * p.nEnvEdges + 5L * E) : 0;
    fr[FAM_OBS] = (O > 0) ? srow : 0;     fn[FAM_OBS] = (O > 0) ? snnz : 0;
    long Z = (long)p.K * (p.K - 1) / 2;
    fr[FAM_A2A_ZERO] = (C > 1) ? 20L * Z * N : 0;  fn[FAM_A2A_ZERO] = fr[FAM_A2A_ZERO];
    fr[FAM_A2A] = (C > 1) ? 24L * p.P * N : 0;     fn[FAM_A2A] = (C > 1) ? 76L * p.P * N : 0;
    long ar = 0, an = 0;
    for (int f = 0; f < NUM_FAM; ++f) { p.fam_row[f] = ar; p.fam_nnz[f] = an; ar += fr[f]; an += fn[f]; }
    p.fam_row[NUM_FAM] = ar; p.fam_nnz[NUM_FAM] = an;
}

}  // namespace miqp
