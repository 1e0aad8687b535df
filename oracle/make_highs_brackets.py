#!/usr/bin/env python
"""HiGHS brackets of the big-M model: an optimum interval [lower, upper] per instance that is independent of both
branch-and-bound codes of this repository (the CPU oracle and the CUDA search share the disjunctive reformulation,
oracle/miqp_oracle_bnb.c:11-21; this script does not use it).  TEST INFRASTRUCTURE.

Method: the rows of the OPL model as the oracle's row generator emits them (oracle.build_rows: big-M rows in OPL order, pinned
to the reference's row / non-zero counts) go to scipy.optimize.milp (HiGHS branch and cut) unchanged.  The objective
(cplexmodel/objective_function.mod) is a separable convex quadratic sum_k 1/2 q_k x_k^2 + c_k x_k + const; q, c and const are
read off the oracle's evaluator (pinned to the CPLEX golden vector) by probing unit vectors, and checked on random vectors.
Every quadratic column gets an epigraph variable t_k >= 1/2 q_k x_k^2, outer-approximated by tangents (Kelley's cutting planes):

    lower = optimum of the MILP with the tangents collected so far      (valid lower bound of the MIQP at any time)
    upper = true objective of the MILP's point (feasible for the big-M model)
    new tangents at the MILP's point; stop when upper - lower <= rel * |upper|.

Writes tests/golden/highs_brackets.json: generator call, lower, upper, iterations, HiGHS gap.
Run:  python oracle/make_highs_brackets.py [--rel 2e-3] [--only NAME]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "highs_brackets.json")

# (name, generator in planner-miqp_b200/scenarios.py, kwargs): small shapes of the configs the GPU tests use
INSTANCES = [
    ("lane_following_n8", "lane_following", dict(seed=0, nr_regions=16, nr_steps=8)),
    ("lane_following_seed3_n10", "lane_following", dict(seed=3, nr_regions=16, nr_steps=10)),
    ("config2_seed0_n10", "obstacle_scenario", dict(seed=0, nr_regions=32, nr_steps=10)),
    ("config2_seed1_n10", "obstacle_scenario", dict(seed=1, nr_regions=32, nr_steps=10)),
    ("config2_seed2_n10", "obstacle_scenario", dict(seed=2, nr_regions=32, nr_steps=10)),
    ("config2_seed5_n12", "obstacle_scenario", dict(seed=5, nr_regions=32, nr_steps=12)),
    ("config2_soft_seed0_n10", "obstacle_scenario", dict(seed=0, nr_regions=32, nr_steps=10, soft=True)),
    ("config2_soft_seed3_n10", "obstacle_scenario", dict(seed=3, nr_regions=32, nr_steps=10, soft=True)),
    ("two_cars_5_steps_close", "parallel_lanes", dict(n_cars=2, nr_steps=5, lane_offset=3.5)),
    ("two_cars_5_steps_slack", "parallel_lanes", dict(n_cars=2, nr_steps=5, lane_offset=4.5)),
    ("two_cars_6_steps_stagger", "parallel_lanes", dict(n_cars=2, nr_steps=6, lane_offset=4.8, stagger=1.0)),
    ("config2_seed7_n14", "obstacle_scenario", dict(seed=7, nr_regions=32, nr_steps=14)),
    ("config2_seed11_n12", "obstacle_scenario", dict(seed=11, nr_regions=32, nr_steps=12)),
    ("three_cars_5_steps", "parallel_lanes", dict(n_cars=3, nr_steps=5, lane_offset=4.5)),
    ("two_cars_8_steps_stagger", "parallel_lanes", dict(n_cars=2, nr_steps=8, lane_offset=4.8, stagger=1.0)),
    ("config2_seed3_n16", "obstacle_scenario", dict(seed=3, nr_regions=32, nr_steps=16)),
    ("config2_seed0_n20", "obstacle_scenario", dict(seed=0, nr_regions=32, nr_steps=20)),
]


def make_problem(gen: str, kwargs: dict):
    import planner_miqp_b200  # noqa: F401
    from planner_miqp_b200 import scenarios
    return getattr(scenarios, gen)(**kwargs).build()


def separable_objective(p):
    """(q, c, const) with objective(x) = sum 1/2 q x^2 + c x + const, read off the evaluator and verified"""
    n = O.layout(p).ncols
    f0 = O.objective(p, np.zeros(n))
    q, c = np.zeros(n), np.zeros(n)
    e = np.zeros(n)
    for k in range(n):
        e[k] = 1.0
        fp = O.objective(p, e)
        e[k] = -1.0
        fm = O.objective(p, e)
        e[k] = 0.0
        q[k] = fp + fm - 2.0 * f0
        c[k] = 0.5 * (fp - fm)
    q[np.abs(q) < 1e-12] = 0.0
    c[np.abs(c) < 1e-12] = 0.0
    rng = np.random.default_rng(1)
    for _ in range(5):
        x = rng.normal(size=n) * 3.0
        want = O.objective(p, x)
        got = float(0.5 * (q * x * x).sum() + (c * x).sum() + f0)
        assert abs(want - got) <= 1e-8 * max(1.0, abs(want)), ("objective is not separable quadratic", want, got)
    assert (q >= 0).all()
    return q, c, f0


def bracket(p, rel=2e-3, max_iter=40, time_limit=900.0, verbose=False):
    rowptr, cols, vals, lo, hi = O.build_rows(p)
    nrows, n = len(lo), O.layout(p).ncols
    A = sp.csr_matrix((vals, cols, rowptr), shape=(nrows, n))
    isb, lb, ub = O.col_info(p)
    q, c, f0 = separable_objective(p)
    quad = np.flatnonzero(q > 0)
    nq = len(quad)
    big = 1e20
    lb = np.where(lb < -big, -np.inf, lb)
    ub = np.where(ub > big, np.inf, ub)
    lo = np.where(lo < -big, -np.inf, lo)
    hi = np.where(hi > big, np.inf, hi)
    # columns: x (n), t (nq)
    cost = np.concatenate([c, np.ones(nq)])
    integrality = np.concatenate([isb.astype(int), np.zeros(nq, dtype=int)])
    bounds = Bounds(np.concatenate([lb, np.zeros(nq)]), np.concatenate([ub, np.full(nq, np.inf)]))
    A_ext = sp.hstack([A, sp.csr_matrix((nrows, nq))]).tocsr()
    # tangent of t_k >= 1/2 q x^2 at x = a:   t_k - q a x_k >= -1/2 q a^2
    tr, tc, tv, tlo = [], [], [], []

    def add_tangents(points):   # points: array [nq] of x values
        for j, k in enumerate(quad):
            a = float(points[j])
            r = len(tlo)
            tr.extend([r, r]); tc.extend([n + j, k]); tv.extend([1.0, -q[k] * a])
            tlo.append(-0.5 * q[k] * a * a)

    # start: tangents at 0 and at +-span of each column (bounds if finite, +-10 otherwise)
    for s in (0.0, 1.0, -1.0):
        pts = np.zeros(nq)
        for j, k in enumerate(quad):
            l_, u_ = lb[k], ub[k]
            span_hi = u_ if np.isfinite(u_) else 10.0
            span_lo = l_ if np.isfinite(l_) else -10.0
            pts[j] = 0.0 if s == 0.0 else (span_hi if s > 0 else span_lo)
        add_tangents(pts)
    best_ub, best_lb, it = np.inf, -np.inf, 0
    x_best = None
    t0 = time.time()
    highs_gap = None
    for it in range(1, max_iter + 1):
        T = sp.csr_matrix((tv, (tr, tc)), shape=(len(tlo), n + nq))
        cons = [LinearConstraint(A_ext, lo, hi), LinearConstraint(T, np.array(tlo), np.inf)]
        res = milp(cost, constraints=cons, integrality=integrality, bounds=bounds,
                   options={"mip_rel_gap": 1e-6, "time_limit": time_limit, "disp": False})
        if res.x is None:
            raise RuntimeError(f"HiGHS: {res.message}")
        x = res.x[:n]
        # HiGHS' dual bound of this MILP is the valid lower bound (its primal value only if solved to optimality)
        mlb = getattr(res, "mip_dual_bound", None)
        lower = (mlb if mlb is not None else res.fun) + f0
        highs_gap = getattr(res, "mip_gap", None)
        best_lb = max(best_lb, lower)
        xr = x.copy()
        xr[isb.astype(bool)] = np.round(xr[isb.astype(bool)])
        viol, _ = O.max_violation(p, xr)
        fx = O.objective(p, xr)
        if viol <= 1e-6 and fx < best_ub:
            best_ub, x_best = fx, xr
        if verbose:
            print(f"  it {it}: lower {best_lb:.8f} upper {best_ub:.8f} (point {fx:.8f}, viol {viol:.1e}) cuts {len(tlo)} {time.time() - t0:.1f}s", flush=True)
        if best_ub - best_lb <= rel * abs(best_ub):
            break
        add_tangents(x[quad])
    return dict(lower=best_lb, upper=best_ub, iterations=it, cuts=len(tlo), highs_mip_gap=highs_gap, seconds=time.time() - t0,
                rows=int(nrows), cols=int(n), binaries=int(isb.sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rel", type=float, default=2e-3)
    ap.add_argument("--only", default=None)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    out = {}
    if os.path.exists(OUT) and args.only:
        with open(OUT) as f:
            out = json.load(f)
    for name, gen, kw in INSTANCES:
        if args.only and name != args.only:
            continue
        p = make_problem(gen, kw)
        print(name, flush=True)
        b = bracket(p, rel=args.rel, verbose=args.verbose)
        b.update(generator=gen, kwargs=kw, rel=args.rel)
        out[name] = b
        print(f"  [{b['lower']:.8f}, {b['upper']:.8f}] after {b['iterations']} MILPs, {b['seconds']:.1f} s", flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
