"""GPU parity for plans with several cars (agent_collision_constraints.mod): the CTA-per-node
kernel (bnb_multi.cu) through the C ABI against the oracle, on horizons where the oracle
proves optimality.  Tolerances as in test_gpu_solve.py."""
import os

import numpy as np
import pytest

import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import parallel_lanes, two_agent_merge, random_scenario
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    s = P.Solver()
    yield s
    s.close()


def check(p, x, info, gap=1e-4):
    xo, io = O.solve(p, gap_tol=gap, time_limit=120)
    assert io.status == 0 and io.proven
    assert info.status == 0 and info.proven, info
    assert info.objective == pytest.approx(io.objective, rel=1e-6, abs=1e-7)
    viol, worst = O.max_violation(p, x)
    assert viol <= 1e-6, (viol, worst)
    assert O.objective(p, x) == pytest.approx(info.objective, rel=1e-9, abs=1e-9)
    return xo, io


@pytest.mark.parametrize("cars,steps,offset,stagger", [(2, 5, 8.0, 0.0), (2, 5, 3.5, 0.0), (2, 5, 4.5, 0.0),
                                                      (2, 6, 4.8, 1.0), (3, 5, 4.5, 0.0)])
def test_parallel_lanes(solver, cars, steps, offset, stagger):
    p = parallel_lanes(cars, steps, offset, stagger=stagger).build()
    x, info = solver.solve(p, gap_tol=1e-4, time_limit=120)
    xo, io = check(p, x, info)
    v, vo = O.block_views(p, x), O.block_views(p, xo)
    for name in ("pos_x", "pos_y"):
        assert np.max(np.abs(v[name] - vo[name])) < 1e-3, name


def test_mixed_batch_single_and_multi(solver, testcase_problem):
    ps = [testcase_problem, parallel_lanes(2, 5, 4.5).build(), parallel_lanes(2, 5, 8.0).build(), testcase_problem]
    xs, infos = solver.solve_batch(ps, gap_tol=1e-4, time_limit=120)
    for p, x, info in zip(ps, xs, infos):
        check(p, x, info)


def test_far_apart_cars_equal_independent_plans(solver):
    """collision rows inactive: the joint optimum is the sum of the single-car optima"""
    b2 = parallel_lanes(2, 12, 12.0)
    p2 = b2.build()
    x, info = solver.solve(p2, gap_tol=1e-4, time_limit=60)
    assert info.status == 0 and info.proven
    tot = 0.0
    for c in range(2):
        b1 = parallel_lanes(1, 12, 12.0)
        b1.cars = [b2.cars[c]]
        p1 = b1.build()
        for k in p1.car:                     # weights of car c in the joint plan (lambda split)
            p1.car[k] = p2.car[k][c:c + 1].copy()
        x1, i1 = solver.solve(p1, gap_tol=1e-4, time_limit=60)
        assert i1.status == 0
        tot += i1.objective
    assert info.objective == pytest.approx(tot, rel=1e-6, abs=1e-7)


def test_time_limited_merge_returns_feasible_incumbent(solver):
    """config 3 shape (N=20): the gap cannot be proven; like the reference at
    its time limit the call reports SUCCESS with an incumbent and an honest gap."""
    p = two_agent_merge(0).build()
    x, info = solver.solve(p, gap_tol=1e-4, time_limit=3.0)
    assert info.status == 0
    viol, worst = O.max_violation(p, x)
    assert viol <= 1e-6, (viol, worst)
    assert info.best_bound <= info.objective + 1e-9
    assert info.seconds < 3.0 + 1.0


def test_single_car_through_multi_kernel(testcase_problem, sos_problem):
    """MIQP_B200_FORCE_MULTI routes single-car plans through the CTA-per-node kernel: same optimum"""
    os.environ["MIQP_B200_FORCE_MULTI"] = "1"
    try:
        s = P.Solver()
        for p in (testcase_problem, sos_problem):
            x, info = s.solve(p, gap_tol=1e-4, time_limit=60)
            check(p, x, info)
        s.close()
    finally:
        os.environ.pop("MIQP_B200_FORCE_MULTI", None)


def test_random_scenarios_of_config4(solver):
    ps = [random_scenario(k, nr_steps=8).build() for k in range(6)]
    xs, infos = solver.solve_batch(ps, gap_tol=1e-4, time_limit=60)
    for p, x, info in zip(ps, xs, infos):
        assert info.status == 0
        viol, _ = O.max_violation(p, x)
        assert viol <= 1e-6
        if info.proven:
            xo, io = O.solve(p, gap_tol=1e-4, time_limit=60)
            if io.proven:
                assert info.objective == pytest.approx(io.objective, rel=1e-6, abs=1e-7)


def test_device_and_oracle_bounds_bracket_each_other_on_unproven_plans(solver):
    """hard multi-car plans stop at the time limit in both searches; sound books mean: nobody's best bound exceeds the other's
    incumbent, and the device's incumbent is feasible for the big-M model"""
    ps = [two_agent_merge(0, nr_steps=8).build(), two_agent_merge(1, nr_steps=10).build(), random_scenario(23, nr_steps=12).build()]
    xs, infos = solver.solve_batch(ps, gap_tol=1e-4, time_limit=2.0)
    for p, x, i in zip(ps, xs, infos):
        xo, io = O.solve(p, gap_tol=1e-4, time_limit=8.0)
        assert i.status == 0 and io.status == 0
        tol = 1e-7 * max(1.0, abs(io.objective))
        assert i.best_bound <= io.objective + tol, (i.best_bound, io.objective)
        assert io.best_bound <= i.objective + tol, (io.best_bound, i.objective)
        viol, worst = O.max_violation(p, x)
        assert viol <= 1e-6, (viol, worst)
        assert O.objective(p, x) == pytest.approx(i.objective, rel=1e-9, abs=1e-9)
