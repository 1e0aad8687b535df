import sys, os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario
from oracle import oracle as O
plans=[obstacle_scenario(s).build() for s in range(2048)]
s=P.Solver(); xs,infos=s.solve_batch(plans,gap_tol=1e-4,time_limit=600.0)
bad=[k for k,i in enumerate(infos) if abs(i.best_bound-i.objective)>2e-4*abs(i.objective)]
print("batch: plans whose bound and objective differ by more than 2e-4:", [(k, infos[k].objective, infos[k].best_bound, infos[k].nodes) for k in bad])
for k in bad[:3] + [89]:
    x,i=s.solve(plans[k],gap_tol=1e-4,time_limit=60)
    xo,io=O.solve(plans[k],gap_tol=1e-4,time_limit=60)
    print("alone", k, i.objective, i.best_bound, i.nodes, "oracle", io.objective)
