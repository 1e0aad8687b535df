"""Small workload for compute-sanitizer (memcheck / racecheck): a few config-2 plans, the reference fixture and a two-car plan."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario, parallel_lanes
from oracle.dat_io import read_dat
plans = [obstacle_scenario(k).build() for k in range(6)] + [read_dat(os.path.join(ROOT, "tests", "golden", "cplexmodel_testcase.dat"))]
s = P.Solver(max_rounds=int(sys.argv[1]) if len(sys.argv) > 1 else 6)
xs, infos = s.solve_batch(plans, gap_tol=1e-4, time_limit=600.0)
print("single:", [(i.status, i.nodes) for i in infos])
xs, infos = s.solve_batch([parallel_lanes(2, nr_steps=5).build()], gap_tol=1e-4, time_limit=600.0)
print("multi:", [(i.status, i.nodes) for i in infos])
