"""Per-shard statistics of the bench workload: step time, nodes / rounds per plan, the hardest plans, soundness counters.
Diagnostic only.   python tools/shard_stats.py --shards 0 1 2 3 [--batch 2048]"""
import argparse, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--shards", type=int, nargs="+", default=[0, 1])
ap.add_argument("--seeds", type=int, nargs="*", default=None, help="solve these seeds one by one instead")
a = ap.parse_args()
s = P.Solver()
if a.seeds:
    for seed in a.seeds:
        p = obstacle_scenario(seed).build()
        x, i = s.solve(p, gap_tol=1e-4, time_limit=60.0)
        print("seed", seed, i, "ms", s.run_stats()["total_ms"])
        prof = s.debug_profile()
        if sum(prof):
            print("  iteration histogram", {k: v for k, v in enumerate(prof[:101]) if v}, "not converged but feasible", prof[149], "status != 0", prof[132])
            tr = s.debug_traces()
            for k in range(3):
                n = int(tr[k, 0])
                if n <= 0: continue
                print(f"  trace {k}: {n} iterations, objective {tr[k,1]:.6f}, depth {int(tr[k,2])}")
                for it in range(min(n, 100)):
                    al, mu, rp, sg, lm = tr[k, 8 + 5 * it: 13 + 5 * it]
                    print(f"     it {it:2d} alpha {al:.3e} mu {mu:.3e} rp {rp:.3e} sigma {sg:.3e} lmax {lm:.3e}")
    sys.exit(0)
for sh in a.shards:
    plans = [obstacle_scenario(sh * a.batch + k).build() for k in range(a.batch)]
    s.upload(plans, gap_tol=1e-4, time_limit=600.0)
    s.run()
    ms = s.run()
    xs, infos = s.fetch()
    st = s.run_stats()
    n = np.array([i.nodes for i in infos]); r = np.array([i.rounds for i in infos])
    hard = np.argsort(-n)[:6]
    print(f"shard {sh}: {ms:.1f} ms, rounds {st['rounds']}, nodes/plan {n.mean():.1f}, proven {sum(i.proven for i in infos)}, "
          f"uncertified {sum(i.uncertified for i in infos)}, pool exhausted {sum(i.pool_exhausted for i in infos)}, "
          f"hardest (seed, nodes): {[(sh * a.batch + int(k), int(n[k])) for k in hard]}", flush=True)
    print("   nodes of the first 32 plans:", [int(v) for v in n[:32]], "rounds:", [int(v) for v in r[:32]])
    sec = np.array([i.seconds for i in infos]) * 1e3
    print("   finish time (ms) of the plans at rank 10%..100%:", [round(float(sec[int(q * (a.batch - 1))]), 1) for q in np.linspace(0.1, 1.0, 10)])
    late = np.argsort(-sec)[:8]
    print("   last to finish (rank, ms, nodes, rounds):", [(int(k), round(float(sec[k]), 1), int(n[k]), int(r[k])) for k in late])
    print("   node kernel ms", round(st["node_kernel_ms"], 1), "iters/node", round(st["qp_iters"] / max(st["nodes"], 1), 2))
