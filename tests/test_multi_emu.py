"""Host logic of the multi-car node processing (planner-miqp_b200/csrc/node_qp_multi.cuh,
bnb_multi_core.cuh), compiled for the CPU with one emulated thread (tests/emu) and wrapped in a
sequential branch and bound, against the oracle.  Objective parity 1e-6 relative where the
oracle proves a 1e-4 gap (north_star tolerance: 1e-4)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu  # noqa: E402
from oracle import oracle as O  # noqa: E402

import planner_miqp_b200  # noqa: F401,E402
from planner_miqp_b200.scenarios import parallel_lanes  # noqa: E402


def both(p, gap=1e-4):
    r = emu.solve(p, gap, 60.0)
    xo, io = O.solve(p, gap_tol=gap, time_limit=60.0)
    return r, io


def test_single_car_fixtures_through_generic_code(testcase_problem, sos_problem):
    for p, pin in ((testcase_problem, 9.57603), (sos_problem, None)):
        r, io = both(p)
        assert r["status"] == 0 and io.status == 0 and io.proven
        assert r["objective"] == pytest.approx(io.objective, rel=1e-6)
        assert abs(r["bound"] - r["objective"]) <= 1e-4 * abs(r["objective"])
        if pin:
            assert abs(r["objective"] - pin) < 1e-5      # test/cplex_wrapper_test.cc:874


@pytest.mark.parametrize("cars,steps,offset,stagger", [(2, 5, 8.0, 0.0), (2, 5, 3.5, 0.0), (2, 5, 4.5, 0.0),
                                                      (2, 6, 4.8, 1.0), (2, 5, 2.0, 6.0)])
def test_collision_rows_and_slacks(cars, steps, offset, stagger):
    p = parallel_lanes(cars, steps, offset, stagger=stagger).build()
    r, io = both(p)
    assert io.status == 0 and io.proven
    assert r["status"] == 0
    assert r["objective"] == pytest.approx(io.objective, rel=1e-6, abs=1e-7)
    assert abs(r["bound"] - r["objective"]) <= 1e-4 * abs(r["objective"]) + 1e-9


def test_completion_heuristic_keeps_results_and_finds_incumbents_early(monkeypatch):
    """multi-car plans without incumbent also try the completion of a node by the least violated alternatives (bnb_multi.cu;
    here the emulation's EMU_HEUR): a redundant, fully decided node -- optima and bounds are unchanged, and a joint plan with
    dozens of violated collision disjunctions gets its first incumbent within a node budget that a plain dive exhausts"""
    from planner_miqp_b200.scenarios import intersection
    for args in ((2, 5, 3.5, 0.0), (2, 6, 4.8, 1.0)):
        p = parallel_lanes(args[0], args[1], args[2], stagger=args[3]).build()
        monkeypatch.setenv("EMU_HEUR", "0")
        r0 = emu.solve(p, 1e-4, 60.0)
        monkeypatch.setenv("EMU_HEUR", "1")
        r1 = emu.solve(p, 1e-4, 60.0)
        assert r0["status"] == r1["status"] == 0
        assert r1["objective"] == pytest.approx(r0["objective"], rel=1e-9)
        assert abs(r1["bound"] - r1["objective"]) <= 1e-4 * abs(r1["objective"]) + 1e-9
    p = intersection(0, n_cars=6, nr_steps=10).build()
    monkeypatch.setenv("EMU_HEUR", "0")
    plain = emu.solve(p, 1e-4, 120.0, max_nodes=60)
    monkeypatch.setenv("EMU_HEUR", "1")
    heur = emu.solve(p, 1e-4, 120.0, max_nodes=60)
    assert heur["status"] == 0 and heur["objective"] < float("inf")
    assert plain["status"] != 0 or plain["objective"] >= heur["objective"] - 1e-9 or plain["nodes"] >= heur["nodes"]


def test_bounds_of_two_independent_searches_bracket_each_other_on_hard_multi_car_plans():
    """plans that neither search proves within its budget (active collision rows, slack-paying merges): if both keep sound books,
    the best bound of one can never exceed the incumbent of the other (the multi-car kernel used to score stalled
    relaxations with an upper bound)"""
    from planner_miqp_b200.scenarios import two_agent_merge, random_scenario
    cases = [two_agent_merge(0, nr_steps=8).build(), two_agent_merge(1, nr_steps=10).build(), random_scenario(23, nr_steps=12).build()]
    for p in cases:
        r = emu.solve(p, 1e-4, 30.0, max_nodes=1200)
        xo, io = O.solve(p, gap_tol=1e-4, time_limit=8.0)
        assert r["status"] == 0 and io.status == 0
        tol = 1e-7 * max(1.0, abs(io.objective))
        assert r["bound"] <= io.objective + tol, (r["bound"], io.objective)
        assert io.best_bound <= r["objective"] + tol, (io.best_bound, r["objective"])
        assert r["bound"] <= r["objective"] + tol and io.best_bound <= io.objective + tol
