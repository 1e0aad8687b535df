"""Independent pin of the disjunctive reformulation: tests/golden/highs_brackets.json holds, per small
instance, an optimum interval of the BIG-M model computed by HiGHS branch and cut on the oracle's OPL rows with Kelley
cuts for the separable quadratic objective (oracle/make_highs_brackets.py; neither branch-and-bound code of this repository is
involved).  Both searches work on the disjunctive form (oracle/miqp_oracle_bnb.c:11-21): their proven optima must lie in the
bracket.  CPU part: the oracle; GPU part: the CUDA search through the C ABI."""
import json
import os

import pytest

import planner_miqp_b200  # noqa: F401
from planner_miqp_b200 import scenarios
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "highs_brackets.json")) as f:
    BRACKETS = json.load(f)
GAP = 1e-4


def problem(name):
    b = BRACKETS[name]
    return getattr(scenarios, b["generator"])(**b["kwargs"]).build()


def inside(obj, b):
    # a proven 1e-4 optimum f satisfies  f* <= f <= f* (1 + gap);  lower <= f* <= upper
    tol = 1e-7 * max(1.0, abs(obj))
    return b["lower"] - tol <= obj <= b["upper"] * (1.0 + 2.0 * GAP) + tol


def test_brackets_are_tight_and_cover_the_instance_kinds():
    assert len(BRACKETS) >= 10
    kinds = {b["generator"] for b in BRACKETS.values()}
    assert {"lane_following", "obstacle_scenario", "parallel_lanes"} <= kinds
    assert any(b["kwargs"].get("soft") for b in BRACKETS.values())
    for name, b in BRACKETS.items():
        assert b["lower"] <= b["upper"], name
        assert b["upper"] - b["lower"] <= 2e-3 * abs(b["upper"]), name


@pytest.mark.parametrize("name", sorted(BRACKETS))
def test_oracle_optimum_inside_highs_bracket(name):
    p = problem(name)
    b = BRACKETS[name]
    sz = O.sizes(p)
    assert (sz.nrows, O.layout(p).ncols) == (b["rows"], b["cols"])     # the same model the bracket was computed on
    x, info = O.solve(p, gap_tol=GAP, time_limit=120.0)
    assert info.status == 0 and info.proven
    assert inside(info.objective, b), (info.objective, b["lower"], b["upper"])


@pytest.mark.gpu
def test_device_optima_inside_highs_brackets():
    import planner_miqp_b200 as P
    names = sorted(BRACKETS)
    ps = [problem(n) for n in names]
    s = P.Solver()
    xs, infos = s.solve_batch(ps, gap_tol=GAP, time_limit=120.0)
    s.close()
    for n, p, x, i in zip(names, ps, xs, infos):
        assert i.status == 0 and i.proven, (n, i)
        assert inside(i.objective, BRACKETS[n]), (n, i.objective, BRACKETS[n]["lower"], BRACKETS[n]["upper"])
        viol, _ = O.max_violation(p, x)
        assert viol <= 1e-6, (n, viol)
