// multi_emu.cpp -- TEST INFRASTRUCTURE.  Compiles the multi-car node processing of the CUDA
// backend (planner-miqp_b200/csrc/bnb_multi_core.cuh, node_qp_multi.cuh) for the host with one
// emulated thread (MQ_EMULATE) and wraps it in a plain sequential branch and bound, so that the
// logic of the device code (row generation, slack elimination, Riccati sweeps, scan, branching)
// can be checked against the CPU oracle without a GPU.  Never loaded by the product.
#define MQ_EMULATE 1
#include <cuda_runtime.h>

#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <queue>
#include <vector>

#include "../../planner-miqp_b200/csrc/host_pack.hpp"
#include "../../planner-miqp_b200/csrc/formulation_tables.cuh"
#include "../../planner-miqp_b200/csrc/bnb_multi_core.cuh"

using namespace miqp;
using namespace miqp::hostpack;

namespace {
struct Node { double bound; int depth, rank; long birth; unsigned long long uid; std::vector<unsigned char> dec; };
struct Cmp {
  bool have_inc; long round; int policy;   // policy 0: deepest first; 1: children of the last round first, else best bound
  bool before(const Node &x, const Node &y) const {
    if (!have_inc && policy >= 1) {
      const bool nx = (x.birth == round - 1), ny = (y.birth == round - 1);
      if (nx != ny) return nx;
      if (nx) { if (x.rank != y.rank) return x.rank < y.rank; if (x.bound != y.bound) return x.bound < y.bound; return x.uid < y.uid; }
      if (x.bound != y.bound) return x.bound < y.bound;
      if (x.depth != y.depth) return x.depth > y.depth;
      return x.uid < y.uid;
    }
    if (have_inc && policy == 3 && (round & 1)) {   // every other round: the preferred children of the last two rounds first (plunging), else best bound
      const bool nx = (x.birth >= round - 2 && x.rank <= 0), ny = (y.birth >= round - 2 && y.rank <= 0);
      if (nx != ny) return nx;
      if (nx) { if (x.depth != y.depth) return x.depth > y.depth; if (x.bound != y.bound) return x.bound < y.bound; return x.uid < y.uid; }
    }
    if (have_inc && policy == 2) {   // best bound with plunging: the children of the last round first (least violated alternative first)
      const bool nx = (x.birth == round - 1), ny = (y.birth == round - 1);
      if (nx != ny) return nx;
      if (nx) { if (x.rank != y.rank) return x.rank < y.rank; if (x.bound != y.bound) return x.bound < y.bound; return x.uid < y.uid; }
    }
    if (!have_inc) { if (x.depth != y.depth) return x.depth > y.depth; if (x.rank != y.rank) return x.rank < y.rank; if (x.bound != y.bound) return x.bound < y.bound; return x.uid < y.uid; }
    if (x.bound != y.bound) return x.bound < y.bound;
    if (x.depth != y.depth) return x.depth > y.depth;
    return x.uid < y.uid;
  }
};
}  // namespace

extern "C" int emu_multi_solve(const MiqpB200Problem *q, double gap_tol, double time_limit, long max_nodes,
                               double *obj_out, double *bound_out, long *nodes_out, long *iters_out,
                               double *traj_out /* [C][N][8] */, double *sig_out /* [P][N][4] */,
                               unsigned char *dec_out /* [ndec_pad] */, int verbose, const double *warm /* full column vector or null */) {
  Packed pk;
  std::string v = validate(*q);
  if (!v.empty()) { fprintf(stderr, "emu: %s\n", v.c_str()); return -2; }
  pack_one(*q, pk);
  DevProb &p = pk.probs[0];
  // the device-filled tables live behind the inputs (host_pack.hpp:place_plan); on the host they need real storage
  place_plan(p, 0, 0, (long)pk.dblob.size(), (long)pk.iblob.size(), 0, 0, 0);
  pk.dblob.resize(pk.dblob.size() + (size_t)pk.dderived, 0.0);
  pk.iblob.resize(pk.iblob.size() + (size_t)pk.iderived, 0);
  prepare_tables_parallel(p, pk.dblob.data(), pk.iblob.data(), 0, 1);
  prepare_tables_serial(p, pk.dblob.data(), pk.iblob.data());
  const int nds = p.ndec_pad;
  const MLayout L = multi_layout(p.C, p.N, p.P, p.kmax, nds);
  std::vector<double> ws((size_t)L.total_bytes / 8 + 8, 0.0);
  MCtx k;
  k.D = pk.dblob.data(); k.I = pk.iblob.data(); k.tid = 0; k.nthr = 1;
  multi_bind(k, &p, ws.data(), nds);
  MShared sh;

  std::vector<Node> open;
  Node root; root.bound = -HUGE_VAL; root.depth = 0; root.rank = 0; root.uid = 1; root.birth = 0; root.dec.assign(nds, UNDEC);
  open.push_back(root);
  if (warm) {  // MIP start: (partially) decided node, evaluated first
    Node wn; wn.bound = -HUGE_VAL; wn.depth = 1 << 20; wn.rank = -1; wn.uid = 0; wn.birth = 0; wn.dec.assign(nds, UNDEC);
    std::vector<int> dummy; decisions_from_solution(*q, p, dummy, warm, wn.dec.data());
    open.push_back(wn);
  }
  double ub = HUGE_VAL, pruned_lb = HUGE_VAL;
  long nodes = 0, iters = 0; unsigned long long next_uid = 2;
  Cmp cmp; cmp.have_inc = false; cmp.round = 0; { const char *ep = std::getenv("EMU_POLICY"); cmp.policy = ep ? std::atoi(ep) : 1; }
  const char *ekd = std::getenv("EMU_KDIVE"); const int Kdive = ekd ? std::atoi(ekd) : 0;
  const char *eh = std::getenv("EMU_HEUR"); const int heur = eh ? std::atoi(eh) : 0;
  const auto t0 = std::chrono::steady_clock::now();
  bool timed_out = false;
  const char *ek = std::getenv("EMU_K");
  const int K = ek ? std::atoi(ek) : 1;   // nodes taken per round with one cutoff snapshot (as the device does)
  long rounds = 0;
  std::vector<Node> batch;
  while (!open.empty() || !batch.empty()) {
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > time_limit || (max_nodes > 0 && nodes >= max_nodes)) { timed_out = true; break; }
    static double cutoff_round = HUGE_VAL;
    if (batch.empty()) {
      ++rounds;
      cutoff_round = (ub < HUGE_VAL) ? ub - gap_tol * fabs(ub) : HUGE_VAL;
      cmp.round = rounds;
      Cmp snap = cmp;
      const int Kr = (!cmp.have_inc && Kdive > 0) ? Kdive : K;
      for (int t = 0; t < Kr && !open.empty(); ++t) {
        size_t bi = 0;
        for (size_t a = 1; a < open.size(); ++a) if (snap.before(open[a], open[bi])) bi = a;
        batch.push_back(open[bi]); open[bi] = open.back(); open.pop_back();
      }
    }
    Node nd = batch.back(); batch.pop_back();
    const double cutoff = cutoff_round;
    if (nd.bound >= cutoff) { pruned_lb = std::min(pruned_lb, nd.bound); continue; }
    ++nodes;
    memcpy(k.dec, nd.dec.data(), nds);
    MNodeOut out = m_process_node(k, &sh, nd.bound, cutoff, nds);
    iters += out.iters;
    if (verbose > 1) fprintf(stderr, "emu node %ld depth %d what %d obj %.6f kind %d c%d i%d q%d viol %.3g open %zu ub %.6f iters %d\n", nodes, nd.depth, out.what, out.obj, sh.br.kind, sh.br.c, sh.br.i, sh.br.q, sh.br.viol, open.size(), ub, out.iters);
    if (out.what == MN_INFEASIBLE) continue;
    if (out.what == MN_UNKNOWN) { pruned_lb = std::min(pruned_lb, nd.bound); if (verbose) fprintf(stderr, "emu uncertified node\n"); continue; }
    if (out.what == MN_PRUNED) { pruned_lb = std::min(pruned_lb, out.obj); continue; }
    if (out.what == MN_INCUMBENT) {
      if (!out.converged) pruned_lb = std::min(pruned_lb, out.obj);
      if (out.fval < ub) {
        ub = out.fval; cmp.have_inc = true;
        for (int c = 0; c < p.C; ++c) for (int i = 0; i < p.N; ++i) for (int t = 0; t < 8; ++t) traj_out[(c * p.N + i) * 8 + t] = k.Z[(long)i * k.nz + 8 * c + t];
        for (int e = 0; e < p.P * p.N * 4; ++e) sig_out[e] = k.sig[e * SG_SIZE + SG_VAL];
        memcpy(dec_out, k.dec, nds);
        if (verbose) fprintf(stderr, "emu incumbent %.8f after %ld nodes\n", ub, nodes);
      }
      continue;
    }
    pruned_lb = std::min(pruned_lb, out.pruned_min);
    if (heur > 0 && !cmp.have_inc && nd.depth < (1 << 20) && out.soff >= 0 && (nd.depth % heur) == 0) {
      // primal heuristic: the completion by the least violated alternatives as one extra, fully decided node (redundant: the
      // children below still cover the node)
      Node hn; hn.bound = out.obj; hn.depth = 1 << 20; hn.rank = -2; hn.uid = next_uid++; hn.birth = rounds; hn.dec.assign(k.imp, k.imp + nds);
      open.push_back(hn);
    }
    const unsigned char *src = out.from_imp ? k.imp : k.dec;
    for (int a = 0; a < out.nalt; ++a) {
      Node ch; ch.bound = (std::getenv("EMU_NO_CHILD_BOUND") ? out.obj : sh.cb[a]); ch.depth = nd.depth + 1; ch.rank = a; ch.uid = next_uid++; ch.birth = rounds;
      ch.dec.assign(src, src + nds);
      if (out.soff >= 0) { ch.dec[out.soff] = sh.alts[a]; if (sh.alts[a] == k.imp[out.soff]) ch.rank = -1; }
      else ch.rank = 0;
      open.push_back(ch);
    }
  }
  double lb = pruned_lb;
  for (const Node &n : open) lb = std::min(lb, n.bound);
  if (!timed_out && open.empty() && lb == HUGE_VAL) lb = ub;
  if (ub < HUGE_VAL && lb > ub) lb = ub;
  *obj_out = ub; *bound_out = lb; *nodes_out = nodes; *iters_out = iters;
  if (verbose) fprintf(stderr, "emu rounds %ld nodes %ld\n", rounds, nodes);
  if (!(ub < HUGE_VAL)) return timed_out ? 3 : 1;
  return 0;
}
