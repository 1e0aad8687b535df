"""Finds the plan of a batch that needs the most nodes and traces its search alone. Diagnostic only."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
plans = [obstacle_scenario(k).build() for k in range(B)]
s = P.Solver()
s.upload(plans, gap_tol=1e-4, time_limit=600.0); s.run(); xs, infos = s.fetch()
n = np.array([i.nodes for i in infos])
order = np.argsort(-n)[:5]
print("hardest", [(int(k), int(n[k])) for k in order])
k = int(order[0])
s2 = P.Solver(verbose=2)
x, info = s2.solve(plans[k], gap_tol=1e-4, time_limit=60)
print("alone:", info)
