"""Minimal driver for ncu: upload a batch of config-2 plans and run the device solve once
(or a few times).  Never used for reported numbers."""
import argparse
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import planner_miqp_b200 as P  # noqa: E402
from planner_miqp_b200.scenarios import obstacle_scenario  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--runs", type=int, default=1)
ap.add_argument("--nodes-per-round", type=int, default=0)
a = ap.parse_args()
plans = [obstacle_scenario(k).build() for k in range(a.batch)]
s = P.Solver(nodes_per_round=a.nodes_per_round)
s.upload(plans, gap_tol=1e-4, time_limit=600.0)
for _ in range(a.runs):
    ms = s.run()
    st = s.run_stats()
    print(f"run: {ms:.2f} ms, rounds {st['rounds']}, nodes {st['nodes']}, iters {st['qp_iters']}, node kernel {st['node_kernel_ms']:.2f} ms")
xs, infos = s.fetch()
print("proven", sum(i.proven for i in infos), "of", len(infos))
