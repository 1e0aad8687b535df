"""Per-round trace of the device search (work items, active plans, node-kernel ms) plus the
distribution of nodes per plan.  Diagnostic only."""
import argparse, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
a = ap.parse_args()
plans = [obstacle_scenario(k).build() for k in range(a.batch)]
s = P.Solver(verbose=2)
s.upload(plans, gap_tol=1e-4, time_limit=600.0)
s.run()
ms = s.run()
xs, infos = s.fetch()
n = np.array([i.nodes for i in infos]); r = np.array([i.rounds for i in infos]); it = np.array([i.qp_iters for i in infos])
print("ms", ms, "nodes/plan mean", n.mean(), "pcts", np.percentile(n, [50, 90, 99, 100]))
print("rounds/plan pcts", np.percentile(r, [50, 90, 99, 100]))
print("iters/node", it.sum() / n.sum())
prof = s.debug_profile()
if sum(prof):
    h = prof[:101]; tot = sum(h)
    cum = 0; qs = {}
    for k, v in enumerate(h):
        cum += v
        for q in (0.5, 0.9, 0.99, 1.0):
            if q not in qs and cum >= q * tot: qs[q] = k
    print("iters/node quantiles", qs, "nodes", tot, "infeasible", prof[132], "iters of infeasible", prof[133])
    c = prof[128:132]
    print("cycles: rows %.1f%% factor %.1f%% sweeps %.1f%% of node total; per node %.0f cycles; per iter %.0f cycles" % (
        100 * c[0] / c[3], 100 * c[1] / c[3], 100 * c[2] / c[3], c[3] / tot, c[3] / max(sum(k * v for k, v in enumerate(h)), 1)))
    its = max(sum(k * v for k, v in enumerate(h)), 1)
    print("row passes, cycles per iteration: A visit %.0f, A sub-lane reduce+epilogue %.0f, D %.0f, E %.0f, G %.0f, A team reduce (incl. wait) %.0f" % tuple(x / its for x in prof[134:140]))
    hi = prof[150:251]
    print("iterations of infeasible relaxations (iters: count):", {k: v for k, v in enumerate(hi) if v})
    print("iterations of feasible relaxations >= 18:", {k: h[k] - hi[k] for k in range(18, 101) if h[k] - hi[k]})
    tr = s.debug_traces()
    for k in range(8):
        n = int(tr[k, 0])
        if n <= 0: continue
        print(f"slow feasible relaxation {k}: {n} iterations, objective {tr[k,1]:.6f}, depth {int(tr[k,2])}")
        for it in range(min(n, 100)):
            a, mu, rp, sg, lm = tr[k, 8 + 5 * it: 13 + 5 * it]
            print(f"   it {it:2d} alpha {a:.3e} mu {mu:.3e} rp {rp:.3e} sigma {sg:.3e} lmax {lm:.3e}")
        if k >= 2: break
    print("accepted without convergence (stalled iteration, feasible point):", prof[149])
