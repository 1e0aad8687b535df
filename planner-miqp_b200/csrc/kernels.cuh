// kernels.cuh -- launcher declarations shared by the host driver (solver.cu).
#pragma once
#include "dev_problem.cuh"

namespace miqp {

// ---- formulation.cu -------------------------------------------------------------------
void launch_prepare_tables(DevProb *probs, double *dblob, int *iblob, int count, cudaStream_t st);
void launch_assemble_rows(const DevProb *probs, const double *dblob, const int *iblob, int count, long max_rows,
                          long *rowptr, int *cols, double *vals, double *lo, double *hi,
                          unsigned long long *nnz_count, cudaStream_t st);
void launch_evaluate(const DevProb *probs, const double *dblob, const int *iblob, int count, long max_rows,
                     const double *xall, double *max_viol, double *objective, cudaStream_t st);

// ---- bnb.cu ---------------------------------------------------------------------------
// Device-resident branch and bound state of a batch of plans.
struct BnbState {
  int count;            // plans
  int cap;              // node slots per plan
  int ndec_stride;      // bytes per node decision vector (max over plans, multiple of 16)
  int zstride;          // doubles per incumbent trajectory (max C*N*8 + 4*P*N)
  int kmax;             // row slots per stage (max over plans)
  int npad;             // stages padded to a multiple of 32 (max over plans)
  int sel_per_plan;     // stride of sel_idx: most node relaxations a plan may take in one round
  int sel_base;         // nodes per plan per round once an incumbent exists
  int sel_dive;         // nodes per plan per round while diving for the first incumbent
  int dive_patience, inc_patience, dive_growth;   // a plan still without (with) incumbent after dive_patience (inc_patience) rounds widens its beam by dive_growth nodes per round
  int teams;            // node relaxations the persistent kernels keep in flight
  int width_mode;       // sched.cuh:round_width
  int max_rounds;       // > 0: a plan stops after this many rounds (acts like an expired time limit)
  int force_multi;      // route every plan to the CTA-per-node kernel (test hook)
  // node pools [count][cap]
  unsigned char *dec;
  double *bound;
  int2 *meta;           // depth, (birth round << 8) | rank + 1
  unsigned long long *uid;
  int *open_idx; int *open_cnt;
  int *free_stack; int *free_cnt;
  int *sel_idx; int *sel_cnt;      // node slots of the plan's current round [count][sel_per_plan]
  unsigned long long *keybuf;      // [count][cap]
  double *zpool;        // [count][cap][zp_stride] relaxed optimum of the parent (warm start of the child's interior-point solve); null = off
  int zp_stride;
  double warm_mu;       // complementarity target of the warm start
  double tau_k;         // step fraction to the boundary = max(0.995, 1 - tau_k * mu); 0 = constant 0.995
  // per plan
  double *ub;           // incumbent objective (inf if none)
  double *cutoff;       // snapshot used by the node kernels during the plan's current round
  double *pruned_lb;
  int *done;            // 0 running, 1 finished (frontier exhausted), 3 stopped by its time limit / round cap
  int *lock;
  double *inc_z;        // [count][zstride]
  unsigned char *inc_dec;  // [count][ndec_stride]
  unsigned long long *inc_uid;  // tie break between equal incumbents (deterministic result)
  unsigned long long *stat_nodes, *stat_iters, *stat_rows;
  unsigned long long *stat_uncert;   // node relaxations closed without optimum, feasible point or infeasibility certificate
  int *overflow;                     // per plan: 1 if its node pool ran out (children dropped, bound kept in pruned_lb)
  double *dbg;                // [-DMQ_PROF] per-CTA iteration traces [ctas + 8][512]; slots ctas.. hold the claimed slow relaxations
  unsigned long long *prof;   // [256] diagnostics (filled only by -DMQ_PROF builds): [it] histogram of IPM iterations per node, [128..] cycles
  // on-device scheduling (sched.cuh): per-plan rounds, ticket counters, ready bitmaps by rank and plan class (0: one car, 1: several)
  int *rd_base, *rd_end, *rd_next, *rd_left, *rd_round;   // [count]
  unsigned *ready[2];         // [32 buckets][(count + 31) / 32] bit r: the plan of rank r has unclaimed tickets (bucket: sched.cuh:prio_bucket)
  int *bucket_cnt;            // [2][32] plans with a set bit per class and bucket
  int prio_mode;
  int *rank_of, *order;       // plan -> rank, rank -> plan
  int *score;                 // [count] disjunctions violated by the root relaxation (hardness estimate: the hard plans go first)
  int *plans_left;            // [2] unfinished plans per class
  int *stop;                  // written by the host (watchdog): every plan stops at its next round boundary
  unsigned long long *t_start;  // %globaltimer when the batch was initialised
  unsigned long long *t_done;   // [count] %globaltimer when the plan finished
  double *tlimit;             // [count] time limit of the plan in seconds (tilim, cplexmodel.mod:8-10)
};

void launch_bnb_init(const BnbState &st, const DevProb *probs, const unsigned char *warm_dec /* [count][ndec_stride] or null */,
                     const int *has_warm, cudaStream_t s);
void launch_bnb_select_all(const BnbState &st, const DevProb *probs, cudaStream_t s);   // first round of every plan
constexpr int NODE_TEAM_WARPS = 4;        // warps that share one node relaxation (bnb_nodes_kernel), 2 teams per SM
#ifndef MQ_TEAMS_PER_SM
#define MQ_TEAMS_PER_SM 2
#endif
constexpr int NODE_TEAMS_PER_SM = MQ_TEAMS_PER_SM;   // 2: 255 registers per thread; 3: 168 registers (spills), A/B in profiles/r1k
constexpr int NODE_TEAM_WARPS_WIDE = 8;   // the same for batches with fewer plans than SMs (latency mode), 1 team per SM
// persistent node kernel of the single-car plans: returns 0 or a cudaError
int launch_bnb_nodes(const BnbState &st, const DevProb *probs, const double *dblob, const int *iblob,
                     int smem_per_node, int warps_per_cta, int ctas, int maxN, cudaStream_t s);
void launch_bnb_finish(const BnbState &st, const DevProb *probs, const double *dblob, const int *iblob,
                       double *xall, double *best_bound, cudaStream_t s);
int node_kernel_smem_per_warp(int maxN, int kmax, int ndec_stride);
int node_kernel_max_ctas(int smem_per_cta, int threads);

// ---- bnb_multi.cu ---------------------------------------------------------------------
// CTA-per-node kernel for plans with several cars.  ws_bytes = workspace of one node (max over the
// plans of the batch); it lives in shared memory when it fits, else in gws[ctas][ws_bytes].
long multi_workspace_bytes(int C, int N, int P, int kmax, int ndec_stride);
int multi_kernel_max_ctas(int smem_bytes, int threads);
int launch_bnb_nodes_multi(const BnbState &st, const DevProb *probs, const double *dblob, const int *iblob,
                           double *gws, long ws_bytes, int use_smem, int threads, int ctas, cudaStream_t s);

}  // namespace miqp
