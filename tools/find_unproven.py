"""Which plans of a bench shard are not proven, and why (gap, nodes closed without certificate, pool overflow)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import planner_miqp_b200 as P  # noqa: E402
from planner_miqp_b200.scenarios import obstacle_scenario  # noqa: E402
shard = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 2048
plans = [obstacle_scenario(shard * B + k).build() for k in range(B)]
s = P.Solver()
xs, infos = s.solve_batch(plans, gap_tol=1e-4, time_limit=600.0)
for k, i in enumerate(infos):
    if not (i.status == 0 and i.proven):
        print("shard", shard, "plan", k, "seed", shard * B + k, i)
        s2 = P.Solver()
        x, j = s2.solve(plans[k], gap_tol=1e-4, time_limit=600.0)
        print("   alone:", j)
        from oracle import oracle as O
        xo, io = O.solve(plans[k], gap_tol=1e-4, time_limit=120.0)
        print("   oracle:", io.status, io.objective, io.best_bound, io.gap, io.proven, io.nodes)
print("done", sum(1 for i in infos if i.status == 0 and i.proven), "proven of", B)
