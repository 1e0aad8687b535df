"""GPU parity on the synthetic shapes BASELINE.json names, through the solver C ABI:
config 2 (single agent, static + dynamic obstacle, N=40, R=32), the lane-following shape of config 1c,
random single-agent plans of config 4, against the CPU oracle on the same seeded inputs; at the full
bench size through size-independent properties (proven gap, feasibility of every returned vector against
the full big-M model, bound <= objective, independence of a plan's result from its batch).

Tolerance: both solvers stop at a proven relative gap of 1e-4, so two optimal objectives may differ by
2e-4 relative (asserted); constraint violation <= 1e-6 (north_star)."""
import numpy as np
import pytest

import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario, lane_following, random_single_agent
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GAP = 1e-4


@pytest.fixture(scope="module")
def solver():
    s = P.Solver()
    yield s
    s.close()


def _check_batch(solver, plans, label):
    xs, infos = solver.solve_batch(plans, gap_tol=GAP, time_limit=120.0)
    for k, (p, x, info) in enumerate(zip(plans, xs, infos)):
        xo, io = O.solve(p, gap_tol=GAP, time_limit=120.0)
        assert io.status == info.status, (label, k, io.status, info.status)
        if io.status != 0:
            continue
        assert info.proven and io.proven, (label, k)
        assert info.objective == pytest.approx(io.objective, rel=2 * GAP, abs=1e-9), (label, k)
        viol, worst = O.max_violation(p, x)
        assert viol <= 1e-6, (label, k, viol, worst)
        # the bound comes from interior-point solves (1e-9 residuals), the objective from the re-evaluated vector
        assert info.best_bound <= info.objective + 1e-6 * abs(info.objective)


def test_config2_obstacle_plans_match_oracle(solver):
    _check_batch(solver, [obstacle_scenario(s).build() for s in range(20)], "config 2")


def test_config2_soft_obstacles_match_oracle(solver):
    _check_batch(solver, [obstacle_scenario(100 + s, soft=True).build() for s in range(8)], "config 2 soft")


def test_lane_following_and_random_single_agent_match_oracle(solver):
    plans = [lane_following(seed=s).build() for s in range(6)] + [random_single_agent(s).build() for s in range(18)]
    _check_batch(solver, plans, "config 1c / 4")


def test_time_limit_statuses(solver):
    """a time limit that ends the search early: SUCCESS with an unproven gap if an incumbent exists, else FAILED_TIMEOUT
    (reference src/cplex_wrapper.cpp:231-246, test/cplex_wrapper_test.cc:821-842)"""
    s = P.Solver(max_rounds=3)          # the round cap acts like an expired time limit
    p = obstacle_scenario(3).build()
    x, info = s.solve(p, gap_tol=GAP, time_limit=60.0)
    assert info.status in (0, 3)
    if info.status == 0:
        assert not info.proven
        viol, _ = O.max_violation(p, x)
        assert viol <= 1e-6
    else:
        assert np.isnan(info.objective) and np.isnan(info.gap)
    s.close()


def test_full_size_batch_properties(solver):
    """bench-size batch (2048 config-2 plans): every plan proven to 1e-4, every vector feasible for the big-M model"""
    plans = [obstacle_scenario(s).build() for s in range(2048)]
    xs, infos = solver.solve_batch(plans, gap_tol=GAP, time_limit=600.0)
    assert all(i.status == 0 and i.proven for i in infos)
    assert max(i.max_violation for i in infos) <= 1e-6
    assert all(i.gap <= GAP and i.best_bound <= i.objective + 1e-6 * abs(i.objective) for i in infos)
    # device-side evaluation agrees with the oracle's evaluator on a sample
    for k in (0, 511, 1024, 2047):
        assert O.objective(plans[k], xs[k]) == pytest.approx(infos[k].objective, rel=1e-10)
        assert O.max_violation(plans[k], xs[k])[0] <= 1e-6
    # a plan's result does not depend on the batch it is solved in (beyond the gap both runs prove)
    sub = [7, 300, 840, 1500]
    xs2, infos2 = solver.solve_batch([plans[k] for k in sub], gap_tol=GAP, time_limit=600.0)
    for k, i2 in zip(sub, infos2):
        assert i2.objective == pytest.approx(infos[k].objective, rel=2 * GAP)


def test_parked_relaxations_do_not_change_the_result(testcase_problem, monkeypatch):
    """a node relaxation that runs out of its per-round iteration budget is parked in HBM and continues in the next
    round (bnb.cu / node_qp.cuh:SuspendIO); with a tiny budget almost every relaxation is parked several times"""
    plans = [obstacle_scenario(s).build() for s in range(12)] + [testcase_problem]
    results = {}
    for budget in ("0", "3", "10"):
        monkeypatch.setenv("MIQP_SUSP_BUDGET", budget)
        s = P.Solver()
        xs, infos = s.solve_batch(plans, gap_tol=GAP, time_limit=120.0)
        s.close()
        assert all(i.status == 0 and i.proven and i.max_violation <= 1e-6 for i in infos), budget
        results[budget] = [i.objective for i in infos]
    for a, b, c in zip(results["0"], results["3"], results["10"]):
        assert b == pytest.approx(a, rel=2 * GAP) and c == pytest.approx(a, rel=2 * GAP)


def test_time_limit_and_solve_time_are_per_plan():
    """every plan stops at its OWN time limit (max_solution_time) and reports its own solve time, not the batch's
    (reference: one cplex.solve() per plan with CPX_PARAM_TILIM, src/cplex_wrapper.cpp:158-185)"""
    import copy
    from planner_miqp_b200.scenarios import obstacle_scenario, two_agent_merge
    easy = [obstacle_scenario(k).build() for k in range(6)]
    hard = two_agent_merge(0).build()            # does not prove 1e-4 within seconds
    hard.scal = dict(hard.scal); hard.scal["max_solution_time"] = 0.4
    for p in easy:
        p.scal = dict(p.scal); p.scal["max_solution_time"] = 30.0
    s = P.Solver()
    import time
    t0 = time.perf_counter()
    xs, infos = s.solve_batch(easy + [hard], gap_tol=1e-4)
    wall = time.perf_counter() - t0
    s.close()
    assert wall < 5.0                                             # the batch does not wait for the largest limit
    for i in infos[:6]:
        assert i.status == 0 and i.proven and 0.0 < i.seconds < 0.4
    h = infos[6]
    assert not h.proven and 0.4 <= h.seconds < 2.0
    assert h.status in (0, 3)                                     # incumbent found, or FAILED_TIMEOUT without one
    assert len({round(i.seconds, 6) for i in infos}) > 1          # not one number for the whole batch


def test_stalled_leaf_is_resolved_from_the_cold_start():
    """scenario seed 4774 (bench shard 2): the interior-point solve of the optimal leaf stalls when it starts from the parent's
    optimum; it is re-queued and solved from the cold start, so the plan ends at the optimum the oracle proves instead of an
    incumbent 1.5e-3 above it that cannot be proven"""
    p = obstacle_scenario(4774).build()
    s = P.Solver()
    x, info = s.solve(p, gap_tol=1e-4, time_limit=60)
    s.close()
    xo, io = O.solve(p, gap_tol=1e-4, time_limit=60)
    assert io.status == 0 and io.proven
    assert info.status == 0 and info.proven and info.uncertified == 0
    assert info.objective == pytest.approx(io.objective, rel=1e-6)
    viol, _ = O.max_violation(p, x)
    assert viol <= 1e-6
