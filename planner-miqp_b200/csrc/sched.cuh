// sched.cuh -- on-device scheduling of the branch and bound: per-plan rounds without a batch-wide barrier.
//
// What CPLEX's node selection / bounding / pruning does inside cplex.solve() (reference src/cplex_wrapper.cpp:158-185)
// runs here as device code inside the persistent node kernels (bnb.cu, bnb_multi.cu); the host launches one kernel per
// batch and plan class and synchronises once.
//
// Every plan owns a sequence of ROUNDS.  A round is: plan_select() releases the node slots of the previous round,
// snapshots the cutoff from the incumbent, prunes the open list, picks the K best open nodes and publishes them as the
// plan's work list (tickets rd_base .. rd_end-1).  Persistent CTAs claim tickets (acquire_item), solve the node
// relaxation, push children, and report completion (complete_item); the CTA that completes the last ticket of a round
// runs the plan's next plan_select() itself.  Nothing in a plan's search depends on other plans or on timing: K is a
// function of the plan's own round number and of the batch size, incumbents become visible to the plan at its next
// select, ties are broken by node uid -- so a plan returns the same result alone, in any batch, on any number of GPUs.
//
// Which plan a free CTA serves is decided by a bitmap of plans with unclaimed tickets, ordered by the plan's rank:
// the lowest rank wins.  Low ranks therefore run at the latency of their own dependency chain (a dive is a chain of
// node relaxations) while the rest of the batch waits -- instead of all plans advancing one node per batch-wide round --,
// and a hard plan (hundreds of nodes) overlaps with the easy plans behind it instead of forming the tail of the batch.
//
// Memory model: the per-plan state is written by one CTA and read by another inside the same kernel, so every read of
// it goes to L2 (ld.global.cg; L1 is not coherent between SMs) and every hand-over is fenced: writers __threadfence()
// before the barrier that precedes the publishing atomic.
#pragma once
#include "kernels.cuh"
#include "bnb_common.cuh"

namespace miqp {

#ifndef MQ_INF
#define MQ_INF (__longlong_as_double(0x7ff0000000000000LL))
#endif

template <class T> __device__ __forceinline__ T ldg2(const T *p) { return __ldcg(p); }   // L2 load of state shared between CTAs
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// priority key, smaller = earlier.  Without incumbent the search dives: the children created in the last round come first
// (least violated alternative first, then bound); if the dive died (no newborn node) it restarts from the best bound.
// Deepest-first backtracking is deliberately NOT used: it gets trapped below a wrong early decision.  With an incumbent:
// best bound first, then deepest.
__device__ __forceinline__ unsigned long long node_key(double bound, int depth, int rank, unsigned long long uid, bool have_inc, bool newborn) {
  unsigned long long d = 1023 - (unsigned long long)(depth > 1023 ? 1023 : depth);  // 10 bits
  unsigned long long r = (unsigned long long)(rank + 1 > 127 ? 127 : rank + 1);      // 7 bits
  unsigned long long b = ordered_bits(bound) >> 24;                                   // 40 bits
  unsigned long long u = uid & 127ULL;                                                // 7 bits
  if (have_inc) return (b << 24) | (d << 14) | (r << 7) | u;
  if (newborn) return (r << 47) | (b << 7) | u;
  return (1ULL << 63) | (b << 17) | (d << 7) | u;
}
__device__ __forceinline__ int meta_rank(int my) { return (my & 0xff) - 1; }
__device__ __forceinline__ int meta_birth(int my) { return my >> 8; }

// Priority of a plan with unclaimed tickets: bucket first (lower = earlier), then rank.  prio_mode 0: rank only (one bucket);
// 1: plans that have run fewer rounds first (all plans advance together: the hard ones are known, and wide, by the time the
// easy ones are finished); 2: plans that have run more rounds first.
constexpr int PRIO_BUCKETS = 32;
__device__ __forceinline__ int prio_bucket(const BnbState &st, int round) {
  if (st.prio_mode == 0) return 0;
  const int r = round < PRIO_BUCKETS - 1 ? round : PRIO_BUCKETS - 1;
  return st.prio_mode == 1 ? r : PRIO_BUCKETS - 1 - r;
}
constexpr int SEL_MAX_WARPS = 8;
struct SelSmem {
  int warp_tot[SEL_MAX_WARPS];
  int out, free_top, tie, remaining, item[4];
  int hist[256];
  double pruned[SEL_MAX_WARPS];
  unsigned long long prefix;
};

__device__ __forceinline__ int block_excl_scan(int flag, SelSmem &sm, int &total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int pre = __popc(bal & ((1u << lane) - 1));
  if (lane == 0) sm.warp_tot[wid] = __popc(bal);
  __syncthreads();
  int off = 0; total = 0;
  for (int k = 0; k < nw; ++k) { const int t = sm.warp_tot[k]; if (k < wid) off += t; total += t; }
  __syncthreads();
  return off + pre;
}

// Nodes a plan takes in its round `round`.  While more plans are unfinished than node relaxations fit on the machine every
// plan takes one node per round (every extra node is speculative: it may be pruned by an incumbent found in the same
// round); as plans finish, the machine's share of the remaining ones grows (width_mode 1: `teams` / unfinished plans, read
// when the round starts -- the search path then depends on the batch, the result stays within the gap).  A plan that is
// still without incumbent long after the typical plan has finished its dive is a hard one and widens its beam by itself.
// width_mode 0: a function of the plan's own round number only (the search of a plan does not depend on its neighbours).
__device__ __forceinline__ int round_width(const BnbState &st, int round, bool have_inc, int cls) {
  int K = have_inc ? st.sel_base : st.sel_dive;
  if (st.width_mode == 1) {
    const int left = ldg2(&st.plans_left[cls]);
    const int fill = st.teams / (left > 0 ? left : 1);
    const int kf = have_inc ? fill : fill / 2;
    if (kf > K) K = kf;
    if (!have_inc && st.dive_patience > 0 && round > st.dive_patience) { const int kp = (round - st.dive_patience) * st.dive_growth; if (kp > K) K = kp; }
  } else {
    const int patience = have_inc ? st.inc_patience : st.dive_patience;
    if (patience > 0 && round > patience) { const int kp = (round - patience) * st.dive_growth; if (kp > K) K = kp; }
  }
  return K > st.sel_per_plan ? st.sel_per_plan : K;
}

// One round boundary of plan s; executed by every thread of the calling CTA (blockDim.x a multiple of 32, at most 256).
__device__ __forceinline__ void plan_select(const BnbState &st, const DevProb *probs, int s, SelSmem &sm) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const DevProb &p = probs[s];
  const long pb = (long)s * st.cap;
  const int KS = st.sel_per_plan;
  const int cls = (p.C > 1 || st.force_multi) ? 1 : 0;
  const int round = ldg2(&st.rd_round[s]) + 1;
  // 1. release the slots processed in the last round
  const int nsel_prev = ldg2(&st.sel_cnt[s]);
  const int free0 = ldg2(&st.free_cnt[s]);
  if (tid == 0) {
    int f = free0;
    for (int k = 0; k < nsel_prev; ++k) st.free_stack[pb + f++] = ldg2(&st.sel_idx[(long)s * KS + k]);
    sm.free_top = f; sm.out = 0; sm.tie = 0;
  }
  // 2. cutoff snapshot, time limit
  const double ub = ldg2(&st.ub[s]);
  const bool have_inc = ub < MQ_INF;
  const double cutoff = have_inc ? ub - p.gap_tol * fabs(ub) : MQ_INF;
  const unsigned long long now = global_ns();
  const bool expired = (double)(now - ldg2(st.t_start)) * 1e-9 > ldg2(&st.tlimit[s]) || (st.max_rounds > 0 && round > st.max_rounds) || ldg2(st.stop) != 0;
  int K = round_width(st, round, have_inc, cls);
  __syncthreads();
  // 3. prune by bound, compute keys, compact in place
  const int n0 = ldg2(&st.open_cnt[s]);
  double pruned = MQ_INF;
  for (int base = 0; base < n0; base += nt) {
    const int idx = base + tid;
    int slot = -1, keep = 0; unsigned long long key = 0;
    if (idx < n0) {
      slot = ldg2(&st.open_idx[pb + idx]);
      const double b = ldg2(&st.bound[pb + slot]);
      keep = (b < cutoff);
      if (keep) { const int2 m = ldg2(&st.meta[pb + slot]); key = node_key(b, m.x, meta_rank(m.y), ldg2(&st.uid[pb + slot]), have_inc, meta_birth(m.y) == round - 1); }
      else { pruned = fmin(pruned, b); const int pos = atomicAdd(&sm.free_top, 1); st.free_stack[pb + pos] = slot; }
    }
    int total; const int rank = block_excl_scan(keep, sm, total);
    const int out0 = sm.out;
    if (keep) { st.open_idx[pb + out0 + rank] = slot; st.keybuf[pb + out0 + rank] = key; }
    __syncthreads();
    if (tid == 0) sm.out = out0 + total;
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) pruned = fmin(pruned, __shfl_xor_sync(0xffffffffu, pruned, o));
  if ((tid & 31) == 0) sm.pruned[tid >> 5] = pruned;
  __syncthreads();
  const int n1 = sm.out;
  if (tid == 0) {
    double pm = ldg2(&st.pruned_lb[s]);
    for (int k = 0; k < (nt >> 5); ++k) pm = fmin(pm, sm.pruned[k]);
    st.pruned_lb[s] = pm;
  }
  if (expired) K = 0;   // the open list stays as it is: its bounds enter best_bound
  // 4. threshold key of the K best
  unsigned long long T = ~0ULL; int remaining = n1;  // take everything
  if (n1 > K && K > 0) {
    if (tid == 0) { sm.prefix = 0ULL; sm.remaining = K; }
    __syncthreads();
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      for (int k = tid; k < 256; k += nt) sm.hist[k] = 0;
      __syncthreads();
      const unsigned long long prefix = sm.prefix;
      const unsigned long long himask = (pass == 0) ? 0ULL : (~0ULL << (shift + 8));
      for (int idx = tid; idx < n1; idx += nt) {
        const unsigned long long key = ldg2(&st.keybuf[pb + idx]);
        if ((key & himask) == prefix) atomicAdd(&sm.hist[(int)((key >> shift) & 255ULL)], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int rem = sm.remaining, b = 0;
        while (b < 255 && sm.hist[b] < rem) { rem -= sm.hist[b]; ++b; }
        sm.remaining = rem;  // how many to take from bucket b at this digit
        sm.prefix = prefix | ((unsigned long long)b << shift);
      }
      __syncthreads();
    }
    T = sm.prefix; remaining = sm.remaining;
  }
  // 5. hand the selected nodes to the work list, keep the rest
  __syncthreads();
  if (tid == 0) sm.out = 0;
  __syncthreads();
  int nsel_total = 0;
  if (K > 0)
    for (int base = 0; base < n1; base += nt) {
      const int idx = base + tid;
      int slot = -1, sel = 0, keep = 0; unsigned long long key = 0;
      if (idx < n1) {
        slot = ldg2(&st.open_idx[pb + idx]); key = ldg2(&st.keybuf[pb + idx]);
        if (key < T) sel = 1;
        else if (key == T) sel = (atomicAdd(&sm.tie, 1) < remaining);
        keep = !sel;
      }
      int tsel; const int rsel = block_excl_scan(sel, sm, tsel);
      int tkeep; const int rkeep = block_excl_scan(keep, sm, tkeep);
      const int out0 = sm.out;
      if (sel) st.sel_idx[(long)s * KS + nsel_total + rsel] = slot;
      if (keep) { st.open_idx[pb + out0 + rkeep] = slot; st.keybuf[pb + out0 + rkeep] = key; }
      nsel_total += tsel;
      __syncthreads();
      if (tid == 0) sm.out = out0 + tkeep;
      __syncthreads();
    }
  __threadfence();   // sel_idx, open list, free stack: visible before the round is published
  __syncthreads();
  if (tid == 0) {
    st.open_cnt[s] = (K > 0) ? sm.out : n1;
    st.sel_cnt[s] = nsel_total;
    st.free_cnt[s] = sm.free_top;
    st.cutoff[s] = cutoff;
    st.rd_round[s] = round;
    if (nsel_total == 0) {
      // frontier exhausted (everything pruned or solved), or out of time
      st.done[s] = expired && n1 > 0 ? 3 : 1;
      st.t_done[s] = now;
      __threadfence();
      atomicSub(&st.plans_left[cls], 1);
    } else {
      const int base = ldg2(&st.rd_end[s]);   // == rd_next[s]: every ticket of the last round was claimed and completed
      st.rd_base[s] = base; st.rd_left[s] = nsel_total;
      __threadfence();
      atomicExch(&st.rd_end[s], base + nsel_total);
      __threadfence();
      const int rk = st.rank_of[s], bk = prio_bucket(st, round);
      const int nwords = (st.count + 31) >> 5;
      atomicOr(&st.ready[cls][bk * nwords + (rk >> 5)], 1u << (rk & 31));
      atomicAdd(&st.bucket_cnt[cls * PRIO_BUCKETS + bk], 1);
    }
  }
  __syncthreads();
}

// Next (plan, node slot) for this CTA, or plan = -1 when every plan of class `cls` is done.  Called by every thread.
__device__ __forceinline__ int2 acquire_item(const BnbState &st, int cls, SelSmem &sm) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int nwords = (st.count + 31) >> 5;
    int s = -1, slot = -1;
    unsigned backoff = 32;
    for (;;) {
      if (ldg2(&st.plans_left[cls]) <= 0) break;
      // lowest non-empty bucket, then the lowest rank in it
      const int bc = ldg2(&st.bucket_cnt[cls * PRIO_BUCKETS + lane]);
      const unsigned bmask = __ballot_sync(0xffffffffu, bc > 0);
      if (!bmask) { __nanosleep(backoff); if (backoff < 1024) backoff *= 2; continue; }
      const int bk = __ffs(bmask) - 1;
      const unsigned *bm = st.ready[cls] + bk * nwords;
      int found = -1;
      for (int base = 0; base < nwords && found < 0; base += 32) {
        const unsigned wv = (base + lane < nwords) ? ldg2(&bm[base + lane]) : 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, wv != 0u);
        if (bal) { const int src = __ffs(bal) - 1; const unsigned v = __shfl_sync(0xffffffffu, wv, src); found = (base + src) * 32 + (__ffs(v) - 1); }
      }
      if (found < 0) continue;   // the bucket was emptied meanwhile
      const int pl = ldg2(&st.order[found]);
      int got = -1;
      if (lane == 0) {
        const int end = ldg2(&st.rd_end[pl]);
        int t = ldg2(&st.rd_next[pl]);
        while (t < end) { const int old = atomicCAS(&st.rd_next[pl], t, t + 1); if (old == t) { got = t; break; } t = old; }
        if (got >= 0 && got == end - 1) {   // last ticket of the round: the plan leaves the ready set
          atomicAnd(&st.ready[cls][bk * nwords + (found >> 5)], ~(1u << (found & 31)));
          atomicSub(&st.bucket_cnt[cls * PRIO_BUCKETS + bk], 1);
        }
      }
      got = __shfl_sync(0xffffffffu, got, 0);
      if (got >= 0) {
        s = pl;
        slot = ldg2(&st.sel_idx[(long)pl * st.sel_per_plan + (got - ldg2(&st.rd_base[pl]))]);
        break;
      }
    }
    if (lane == 0) { sm.item[0] = s; sm.item[1] = slot; }
  }
  __syncthreads();
  const int2 r = make_int2(sm.item[0], sm.item[1]);
  __syncthreads();
  return r;
}

// Reports the node of plan s as processed; the CTA that completes the last ticket of the round runs the next select.
// Called by every thread after the barrier that follows the node's bookkeeping (whose writers have fenced).
__device__ __forceinline__ void complete_item(const BnbState &st, const DevProb *probs, int s, SelSmem &sm) {
  if (threadIdx.x == 0) { __threadfence(); sm.item[2] = (atomicSub(&st.rd_left[s], 1) == 1) ? 1 : 0; }
  __syncthreads();
  const int last = sm.item[2];
  __syncthreads();
  if (last) plan_select(st, probs, s, sm);
}

}  // namespace miqp
