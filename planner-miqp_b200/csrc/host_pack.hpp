// host_pack.hpp -- host-side flattening of a batch of plans (ModelParameters, reference
// src/miqp_planner_data.hpp:99-185) into the device blobs, closed-form model sizes, the list of
// mode alternatives per car and the MIP-start decisions of a full column vector
// (src/cplex_wrapper.cpp:494-639).  Pure C++; used by the host driver (solver.cu).
#pragma once
#include <cstddef>
#include <algorithm>
#include <new>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/miqp_b200.h"
#include "dev_problem.cuh"

namespace miqp {
namespace hostpack {

// Staging memory of the blobs: page-locked when the CUDA runtime is linked (solver.cu defines
// MIQP_PINNED_BLOBS), so that the H2D / D2H copies of a batch are real asynchronous DMA.
#ifdef MIQP_PINNED_BLOBS
template <class T>
struct PinnedAllocator {
  using value_type = T;
  PinnedAllocator() = default;
  template <class U> PinnedAllocator(const PinnedAllocator<U> &) {}
  T *allocate(size_t n) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); throw std::bad_alloc(); }
    return static_cast<T *>(p);
  }
  void deallocate(T *p, size_t) { cudaFreeHost(p); }
  template <class U> bool operator==(const PinnedAllocator<U> &) const { return true; }
  template <class U> bool operator!=(const PinnedAllocator<U> &) const { return false; }
};
using DVec = std::vector<double, PinnedAllocator<double>>;
using IVec = std::vector<int, PinnedAllocator<int>>;
#else
using DVec = std::vector<double>;
using IVec = std::vector<int>;
#endif

template <class DV, class IV>
struct PackedT {
  std::vector<DevProb> probs;
  DV dblob;
  IV iblob;
  long total_rows = 0, total_nnz = 0, total_cols = 0, max_rows = 0;
  int maxN = 0, max_ndec = 0, max_kmax = 0, max_z = 0, maxC = 0;
  // device-filled tables are not staged on the host: they only reserve space behind the inputs of the whole batch
  long dderived = 0, iderived = 0;
  void reset() { probs.clear(); dblob.clear(); iblob.clear(); total_rows = total_nnz = total_cols = max_rows = 0;
                 maxN = max_ndec = max_kmax = max_z = maxC = 0; dderived = iderived = 0; }
};
using Packed = PackedT<DVec, IVec>;                                   // the batch: page-locked staging blobs
using PackedLocal = PackedT<std::vector<double>, std::vector<int>>;   // one worker's share while packing in parallel

template <class V> inline long push_d(V &b, const double *src, size_t n) {
  long off = (long)b.size();
  if (src) b.insert(b.end(), src, src + n); else b.resize(b.size() + n, 0.0);
  return off;
}
template <class V> inline long push_i(V &b, const int *src, size_t n) {
  long off = (long)b.size();
  if (src) b.insert(b.end(), src, src + n); else b.resize(b.size() + n, 0);
  return off;
}

// Offsets of the device-filled tables are counted from 0 while packing (PackedT::dderived / iderived); place_plan moves
// a plan to its final position: inputs by (dshift, ishift), device-filled tables by (dder, ider) = end of all inputs + the
// tables of the plans before it.
inline void place_plan(DevProb &p, long dshift, long ishift, long dder, long ider, long rows, long nnz, long cols) {
  static_assert(offsetof(DevProb, o_poly) + 5 * sizeof(long) - offsetof(DevProb, o_safety) == 34 * sizeof(long), "input offsets of the double blob must be contiguous");
  static_assert(offsetof(DevProb, o_wtab) - offsetof(DevProb, o_envtab) == 5 * sizeof(long), "table offsets of the double blob must be contiguous");
  static_assert(offsetof(DevProb, o_nalt) - offsetof(DevProb, o_initreg) == 6 * sizeof(long), "input offsets of the int blob must be contiguous");
  static_assert(offsetof(DevProb, o_obsstep_nnz) - offsetof(DevProb, o_posspre) == 4 * sizeof(long), "table offsets of the int blob must be contiguous");
  long *d = &p.o_safety;
  for (int k = 0; k < 35; ++k) d[k] += dshift;
  long *dt = &p.o_envtab;
  for (int k = 0; k < 6; ++k) dt[k] += dder;
  long *i = &p.o_initreg;
  for (int k = 0; k < 7; ++k) i[k] += ishift;
  long *it = &p.o_posspre;
  for (int k = 0; k < 5; ++k) it[k] += ider;
  p.row_base += rows; p.nnz_base += nnz; p.x_base += cols;
}
template <class PK> inline long reserve_d(PK &pk, size_t n) { const long off = pk.dderived; pk.dderived += (long)n; return off; }
template <class PK> inline long reserve_i(PK &pk, size_t n) { const long off = pk.iderived; pk.iderived += (long)n; return off; }

inline void layout_of(const MiqpB200Problem &q, MiqpB200Layout &l) {
  l.C = q.C; l.N = q.N; l.R = q.R; l.O = q.O; l.L = q.L; l.E = q.E; l.K = q.C - 1;
  const int C = q.C, N = q.N, R = q.R, O = q.O, L = q.L, E = q.E, K = l.K;
  int b = 12 * C * N;
  l.base_nwe = b;  b += 5 * C * E * N;
  l.base_ar = b;   b += C * N * R;
  l.base_rcna = b; b += 5 * C * N;
  l.base_dcc = b;  b += C * O * N * L;
  l.base_dcf = b;  b += 4 * C * O * N * L;
  l.base_so = b;   b += C * O * N;
  l.base_sof = b;  b += 4 * C * O * N;
  l.base_c2c = b;  b += 16 * K * K * N;
  l.base_sv = b;   b += 4 * K * K * N;
  l.ncols = b;
}

// closed-form sizes of the big-M model (same arithmetic as prepare_tables_kernel)
inline void model_sizes(const MiqpB200Problem &q, long &rows, long &nnz) {
  const long C = q.C, N = q.N, R = q.R, O = q.O, E = q.E;
  long nE = q.E > 0 ? q.env_off[q.E] : 0;
  long rr = 0, nn = 0;
  for (int c = 0; c < C; ++c) {
    long rp = 0;
    for (int j = 0; j < R; ++j) rp += (q.possible_region[c * R + j] == 1);
    rr += 20 * rp + (R - rp) + 1;
    nn += 76 * rp + (R - rp) + R;
  }
  rows = 12 * C + 5 * R * C + 5 * C + 6 * C * (N - 1) + 12 * C * N + rr * (N - 1) + 15 * R * C * (N - 1);
  nnz = 12 * C + 9 * R * C + 5 * C + 24 * C * (N - 1) + 12 * C * N + nn * (N - 1) + 35 * R * C * (N - 1);
  if (E > 0) { rows += C * N * (5 * nE + 5); nnz += C * N * (15 * nE + 5 * E); }
  if (O > 0)
    for (int i = 0; i < N; ++i)
      for (int o = 0; o < O; ++o) {
        long ne = q.obs_nedges[o * N + i], soft = (q.obs_soft[o] == 1);
        rows += C * (5 * ne + 5); nnz += C * (15 * ne + 5 * (ne + soft));
      }
  if (C > 1) {
    long K = C - 1, Z = K * (K - 1) / 2, P = C * (C - 1) / 2;
    rows += 20 * Z * N + 24 * P * N; nnz += 20 * Z * N + 76 * P * N;
  }
}

inline std::string validate(const MiqpB200Problem &q) {
  if (q.N < 2 || q.N > 64) return "NumSteps must be in [2,64]";
  if (q.R < 1 || q.R > 64) return "nr_regions must be in [1,64]";
  if (q.C < 1 || q.C > 8) return "NumCars must be in [1,8]";
  if (q.O < 0 || q.E < 0 || q.L < 0) return "negative dimension";
  if (q.O > 0 && q.L > 16) return "max_lines_obstacles must be <= 16";
  if (q.E > 250 || q.L > 250) return "too many polygons";
  if (!q.x0 || !q.frac || !q.possible_region || !q.initial_region) return "null array";
  for (int c = 0; c < q.C; ++c)
    if (q.initial_region[c] < 1 || q.initial_region[c] > q.R) return "initial_region out of range";
  if (q.O > 0)
    for (int k = 0; k < q.O * q.N; ++k)
      if (q.obs_nedges[k] < 0 || q.obs_nedges[k] > q.L) return "obstacle polygon with more edges than max_lines_obstacles";
  return "";
}

// mode alternatives of a car: every possible region with its non-dominated low-speed half planes
inline void mode_alternatives(const MiqpB200Problem &q, int c, std::vector<int> &out) {
  out.clear();
  for (int j = 0; j < q.R; ++j) {
    if (q.possible_region[c * q.R + j] != 1) continue;
    const double *f = q.frac + 4 * j;
    const double d1x = f[0], d1y = f[1], d2x = f[2], d2y = f[3];
    int useful[4];
    useful[0] = (d1x > 1e-9 || d2x > 1e-9); useful[1] = (d1y > 1e-9 || d2y > 1e-9);
    useful[2] = (d1x < -1e-9 || d2x < -1e-9); useful[3] = (d1y < -1e-9 || d2y < -1e-9);
    const bool x_dom = (std::fabs(d1x) >= std::fabs(d1y) - 1e-9) && (std::fabs(d2x) >= std::fabs(d2y) - 1e-9);
    const bool y_dom = (std::fabs(d1y) >= std::fabs(d1x) - 1e-9) && (std::fabs(d2y) >= std::fabs(d2x) - 1e-9);
    if (x_dom && (useful[0] || useful[2])) { useful[1] = 0; useful[3] = 0; }
    else if (y_dom && (useful[1] || useful[3])) { useful[0] = 0; useful[2] = 0; }
    for (int h = 0; h < 4; ++h) if (useful[h]) out.push_back(j * 4 + h);
  }
}

template <class PK> inline void pack_one(const MiqpB200Problem &q, PK &pk) {
  DevProb p;
  std::memset(&p, 0, sizeof p);
  const int N = q.N, R = q.R, C = q.C, O = q.O, L = q.L, E = q.E;
  p.N = N; p.R = R; p.C = C; p.O = O; p.L = L; p.E = E; p.K = C - 1; p.P = C * (C - 1) / 2;
  p.nEnvEdges = (E > 0) ? q.env_off[E] : 0;
  p.maxEnvEdges = 0;
  for (int e = 0; e < E; ++e) p.maxEnvEdges = std::max(p.maxEnvEdges, q.env_off[e + 1] - q.env_off[e]);
  p.ts = q.ts;
  p.c2 = 0.5 * (q.ts * q.ts);
  p.c3 = (1.0 / 6.0) * ((q.ts * q.ts) * q.ts);
  p.min_vel = q.min_vel; p.max_vel = q.max_vel;
  p.total_min_acc = q.total_min_acc; p.total_max_acc = q.total_max_acc;
  p.total_min_jerk = q.total_min_jerk; p.total_max_jerk = q.total_max_jerk;
  p.maximum_slack = q.maximum_slack; p.w_slack = q.w_slack; p.w_slack_obs = q.w_slack_obs;
  p.vm = q.min_region_change_speed; p.gap_tol = q.gap_tol;
  auto &d = pk.dblob; auto &ib = pk.iblob;
  p.o_safety = push_d(d, q.safety, N);
  p.o_safety_slack = push_d(d, q.safety_slack, N);
  const double *w[8] = {q.w_pos_x, q.w_vel_x, q.w_acc_x, q.w_pos_y, q.w_vel_y, q.w_acc_y, q.w_jerk_x, q.w_jerk_y};
  for (int k = 0; k < 8; ++k) p.o_w[k] = push_d(d, w[k], C);
  p.o_wb = push_d(d, q.wheelbase, C);
  p.o_radius = push_d(d, q.radius, C);
  p.o_x0 = push_d(d, q.x0, 6 * C);
  p.o_front0 = push_d(d, nullptr, 2 * C);
  for (int c = 0; c < C; ++c) {  // initialization.mod:25-29, initial_conditions.mod:20-23
    const double *x0 = q.x0 + 6 * c;
    const double th = std::atan2(x0[4], x0[1]);
    const double ct = std::cos(th), st = std::sin(th), wb = q.wheelbase[c];
    d[p.o_front0 + 2 * c] = x0[0] + ct * wb;
    d[p.o_front0 + 2 * c + 1] = x0[3] + st * wb;
  }
  const double *ref[4] = {q.x_ref, q.vx_ref, q.y_ref, q.vy_ref};
  for (int k = 0; k < 4; ++k) p.o_ref[k] = push_d(d, ref[k], (size_t)C * N);
  const double *lim[8] = {q.min_acc_x, q.max_acc_x, q.min_acc_y, q.max_acc_y, q.min_jerk_x, q.max_jerk_x, q.min_jerk_y, q.max_jerk_y};
  for (int k = 0; k < 8; ++k) p.o_lim[k] = push_d(d, lim[k], (size_t)C * R);
  p.o_obs_edges = push_d(d, O > 0 ? q.obs_edges : nullptr, (size_t)O * N * L * 4);
  p.o_env_edges = push_d(d, p.nEnvEdges > 0 ? q.env_edges : nullptr, (size_t)p.nEnvEdges * 4);
  p.o_frac = push_d(d, q.frac, (size_t)R * 4);
  const double *poly[6] = {q.poly_sint_ub, q.poly_sint_lb, q.poly_coss_ub, q.poly_coss_lb, q.poly_kappa_max, q.poly_kappa_min};
  for (int k = 0; k < 6; ++k) p.o_poly[k] = push_d(d, poly[k], (size_t)R * 3);
  p.o_envtab = reserve_d(pk, (size_t)p.nEnvEdges * 3);
  p.o_obstab = reserve_d(pk, (size_t)O * N * L * 3);
  p.o_modetab = reserve_d(pk, (size_t)R * 20);
  p.o_fronttab = reserve_d(pk, (size_t)C * R * 12);
  p.o_cost = reserve_d(pk, (size_t)C * N * 16);
  p.o_wtab = reserve_d(pk, (size_t)C * N * 14);

  p.o_initreg = push_i(ib, q.initial_region, C);
  p.o_possible = push_i(ib, q.possible_region, (size_t)C * R);
  p.o_obs_nedges = push_i(ib, O > 0 ? q.obs_nedges : nullptr, (size_t)O * N);
  p.o_obs_soft = push_i(ib, O > 0 ? q.obs_soft : nullptr, O);
  p.o_env_off = push_i(ib, q.env_off, E + 1);
  p.o_alt = push_i(ib, nullptr, (size_t)C * 4 * R);
  p.o_nalt = push_i(ib, nullptr, C);
  std::vector<int> alts;
  for (int c = 0; c < C; ++c) {
    mode_alternatives(q, c, alts);
    ib[p.o_nalt + c] = (int)alts.size();
    for (size_t a = 0; a < alts.size(); ++a) ib[p.o_alt + c * 4 * R + a] = alts[a];
  }
  p.o_posspre = reserve_i(pk, (size_t)C * (R + 1));
  p.o_obsrowpre = reserve_i(pk, (size_t)N * (O + 1));
  p.o_obsnnzpre = reserve_i(pk, (size_t)N * (O + 1));
  p.o_obsstep_rows = reserve_i(pk, N + 1);
  p.o_obsstep_nnz = reserve_i(pk, N + 1);

  MiqpB200Layout l; layout_of(q, l);
  p.base_nwe = l.base_nwe; p.base_ar = l.base_ar; p.base_rcna = l.base_rcna; p.base_dcc = l.base_dcc;
  p.base_dcf = l.base_dcf; p.base_so = l.base_so; p.base_sof = l.base_sof; p.base_c2c = l.base_c2c;
  p.base_sv = l.base_sv; p.ncols = l.ncols;

  long rows, nnz; model_sizes(q, rows, nnz);
  p.row_base = pk.total_rows; p.nnz_base = pk.total_nnz; p.x_base = pk.total_cols;
  pk.total_rows += rows; pk.total_nnz += nnz; pk.total_cols += l.ncols;
  pk.max_rows = std::max(pk.max_rows, rows);

  p.off_mode = 0; p.off_env = p.off_mode + C * N; p.off_obs = p.off_env + 5 * C * N;
  p.off_pair = p.off_obs + 5 * C * O * N; p.ndec = p.off_pair + 4 * p.P * N;
  p.ndec_pad = (p.ndec + 15) & ~15;
  p.kmax = 12 + 5 + (E > 0 ? 5 * p.maxEnvEdges : 0) + 5 * O;
  pk.maxN = std::max(pk.maxN, N); pk.max_ndec = std::max(pk.max_ndec, p.ndec_pad);
  pk.max_kmax = std::max(pk.max_kmax, p.kmax); pk.maxC = std::max(pk.maxC, C);
  pk.max_z = std::max(pk.max_z, C * N * 8 + 4 * p.P * N);
  pk.probs.push_back(p);
}

// decisions of a full column vector (MIP start / warm start).  A NaN in the first binary of a
// family at (car, step) leaves that disjunction undecided (partial MIP start: the node is then
// completed by the least violated alternatives, see scan_node).
inline void decisions_from_solution(const MiqpB200Problem &q, const DevProb &p, const std::vector<int> &alts_all,
                             const double *x, unsigned char *dec) {
  const int C = p.C, N = p.N, R = p.R, E = p.E, O = p.O, L = p.L, K = p.K;
  std::memset(dec, UNDEC, (size_t)p.ndec_pad);
  std::vector<int> alts;
  for (int c = 0; c < C; ++c) {
    mode_alternatives(q, c, alts);
    for (int i = 0; i < N; ++i) {
      int j = 0;
      for (int jj = 0; jj < R; ++jj) if (x[col_ar(p, c, i, jj)] > 0.5) j = jj;
      if (i > 0 && !std::isnan(x[col_ar(p, c, i, 0)])) {
        const double rho = x[col_rcna(p, 4, c, i)];
        if (rho > 0.5) dec[p.off_mode + c * N + i] = MODE_FROZEN;
        else {
          int h = -1;
          for (int t = 0; t < 4; ++t) if (x[col_rcna(p, t, c, i)] < 0.5) { h = t; break; }
          if (h < 0) h = 0;
          bool found = false; int firsth = -1;
          for (int alt : alts) if ((alt >> 2) == j) { if (firsth < 0) firsth = alt & 3; if ((alt & 3) == h) found = true; }
          if (!found && firsth >= 0) h = firsth;
          dec[p.off_mode + c * N + i] = (unsigned char)(j * 4 + h);
        }
      }
      for (int pt = 0; pt < 5; ++pt) {
        if (E > 1 && !std::isnan(x[col_nwe(p, pt, c, 0, i)])) {
          int e = 0;
          for (int ee = 0; ee < E; ++ee) if (x[col_nwe(p, pt, c, ee, i)] < 0.5) { e = ee; break; }
          dec[p.off_env + (c * N + i) * 5 + pt] = (unsigned char)e;
        }
        for (int o = 0; o < O; ++o) {
          const int ne = q.obs_nedges[o * N + i];
          if (ne > 0 && std::isnan((pt == 0) ? x[col_dcc(p, c, o, i, 0)] : x[col_dcf(p, c, o, i, 0, 4 - pt)])) continue;
          int dd = OBS_SOFT;
          for (int ed = 0; ed < ne; ++ed) {
            const double v = (pt == 0) ? x[col_dcc(p, c, o, i, ed)] : x[col_dcf(p, c, o, i, ed, 4 - pt)];
            if (v < 0.5) { dd = ed; break; }
          }
          if (dd == OBS_SOFT && q.obs_soft[o] != 1) continue;   // no separating edge marked on a hard obstacle: undecided
          dec[p.off_obs + ((c * O + o) * N + i) * 5 + pt] = (unsigned char)dd;
        }
      }
    }
  }
  int pr = 0;
  for (int a = 0; a < C - 1; ++a)
    for (int b = a + 1; b < C; ++b, ++pr)
      for (int i = 0; i < N; ++i)
        for (int qd = 0; qd < 4; ++qd) {
          if (std::isnan(x[col_c2c(p, a, b - 1, i, qd * 4)])) continue;
          int dd = 0;
          for (int side = 0; side < 4; ++side) if (x[col_c2c(p, a, b - 1, i, qd * 4 + side)] < 0.5) { dd = side; break; }
          dec[p.off_pair + (pr * N + i) * 4 + qd] = (unsigned char)dd;
        }
  (void)alts_all; (void)L; (void)K;
}

}  // namespace hostpack
}  // namespace miqp
