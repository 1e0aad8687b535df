"""Per-round node-kernel times of one batch (verbose solver output on stderr): A/B of the team sizes."""
import argparse, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import planner_miqp_b200 as P  # noqa: E402
from planner_miqp_b200.scenarios import obstacle_scenario  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
a = ap.parse_args()
plans = [obstacle_scenario(k).build() for k in range(a.batch)]
s = P.Solver(verbose=0)
s.upload(plans, gap_tol=1e-4, time_limit=600.0); s.run()          # warm-up
s.close()
s = P.Solver(verbose=2)
s.upload(plans, gap_tol=1e-4, time_limit=600.0)
ms = s.run()
st = s.run_stats()
print(f"run: {ms:.2f} ms, rounds {st['rounds']}, nodes {st['nodes']}, iters {st['qp_iters']}, node kernel {st['node_kernel_ms']:.2f} ms", file=sys.stderr)
