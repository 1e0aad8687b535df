"""GPU parity, solver part: the device branch and bound (bnb.cu / node_qp.cuh) against the
CPU oracle on the reference's own instances, through the C ABI.

Tolerances (BASELINE.json north_star): objective within 1e-4 relative gap of the oracle's
proven optimum (we assert 1e-6), constraint violation of the returned vector against the
full big-M model <= 1e-6, identical region sequence / obstacle side where the optimum is
unique."""
import numpy as np
import pytest

import planner_miqp_b200 as P
from oracle import oracle as O
from conftest import golden_vector

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    s = P.Solver()
    yield s
    s.close()


def check_against_oracle(p, x, info, gap=1e-4):
    xo, io = O.solve(p, gap_tol=gap, time_limit=120)
    assert io.status == 0 and io.proven
    assert info.status == 0 and info.proven
    assert info.objective == pytest.approx(io.objective, rel=1e-6, abs=1e-9)
    # independent check of the returned vector with the oracle's evaluator
    viol, worst = O.max_violation(p, x)
    assert viol <= 1e-6, (viol, worst)
    assert info.max_violation == pytest.approx(viol, rel=1e-6, abs=1e-12)
    assert O.objective(p, x) == pytest.approx(info.objective, rel=1e-12)
    assert info.gap <= gap
    return xo, io


def test_testcase_matches_reference_and_oracle(solver, testcase_problem):
    p = testcase_problem
    x, info = solver.solve(p, gap_tol=1e-4, time_limit=60)
    xo, io = check_against_oracle(p, x, info)
    assert abs(info.objective - 9.57603) < 1e-5          # test/cplex_wrapper_test.cc:874
    v, vo = O.block_views(p, x), O.block_views(p, xo)
    seq = v["active_region"][0].argmax(axis=1) + 1
    assert list(seq) == [1, 1, 1, 1] + [32] * 16          # test/cplex_wrapper_test.cc:352-393
    assert v["pos_y"][0, 18] < -1.0                        # passes the obstacle on the -y side
    for name in ("pos_x", "pos_y", "vel_x", "vel_y", "acc_x", "acc_y"):
        assert np.max(np.abs(v[name] - vo[name])) < 1e-4, name
    for name in ("active_region", "deltacc", "region_change_not_allowed_combined"):
        assert np.array_equal(np.round(v[name]), np.round(vo[name])), name


def test_testcase_reference_default_gap(solver, testcase_problem):
    x, info = solver.solve(testcase_problem, time_limit=60)    # gap 0.1 from the .dat
    assert info.status == 0
    assert abs(info.objective - 9.57603) <= 0.1 * 9.57603
    assert info.max_violation <= 1e-6


def test_sos_instance(solver, sos_problem):
    x, info = solver.solve(sos_problem, gap_tol=1e-4, time_limit=60)
    check_against_oracle(sos_problem, x, info)


def test_mip_start_is_used(solver, testcase_problem, golden_solution):
    p = testcase_problem
    xg = golden_vector(p, golden_solution)
    x, info = solver.solve(p, gap_tol=1e-4, time_limit=60, warm=xg)
    assert info.status == 0 and abs(info.objective - 9.57603) < 1e-5
    x0, info0 = solver.solve(p, gap_tol=1e-4, time_limit=60)
    assert info.nodes <= info0.nodes + 1


def test_batch_of_identical_and_perturbed_plans(solver, testcase_problem, sos_problem):
    import copy
    ps = []
    for k in range(6):
        q = copy.deepcopy(testcase_problem if k % 2 == 0 else sos_problem)
        q.x0 = q.x0.copy()
        q.x0[0, 3] += 0.05 * k        # lateral offset
        ps.append(q)
    xs, infos = solver.solve_batch(ps, gap_tol=1e-4, time_limit=120)
    for p, x, info in zip(ps, xs, infos):
        check_against_oracle(p, x, info)


def test_results_are_deterministic(solver, testcase_problem):
    a, ia = solver.solve(testcase_problem, gap_tol=1e-4, time_limit=60)
    b, ib = solver.solve(testcase_problem, gap_tol=1e-4, time_limit=60)
    assert np.array_equal(a, b)
    assert ia.nodes == ib.nodes


def test_infeasible_plan_reports_no_solution(solver, testcase_problem):
    import copy
    q = copy.deepcopy(testcase_problem)
    q.x0 = q.x0.copy()
    q.x0[0, 1] = 200.0               # far above max_vel: no feasible trajectory
    x, info = solver.solve(q, gap_tol=1e-4, time_limit=30)
    assert info.status == 1 and np.isnan(info.objective)
    xo, io = O.solve(q, gap_tol=1e-4, time_limit=30)
    assert io.status == 1


def test_compact_results_equal_the_trajectory_blocks_of_the_full_vector(testcase_problem):
    """miqp_b200_batch_fetch_compact / fetch_vector: [C][N][8] trajectories = the pos / vel / acc / u blocks of the RawResults vector"""
    from planner_miqp_b200.scenarios import obstacle_scenario, parallel_lanes
    from planner_miqp_b200.results import block_views
    ps = [testcase_problem, obstacle_scenario(3).build(), parallel_lanes(2, 5, 4.5).build()]
    s = P.Solver()
    s.upload(ps, gap_tol=1e-4, time_limit=60)
    s.run()
    tr, infos_c = s.fetch_compact()
    xs, infos = s.fetch()
    d2h_full = s.run_stats()["d2h_bytes"]
    for k, (p, x) in enumerate(zip(ps, xs)):
        assert infos[k].status == 0 and infos_c[k].objective == infos[k].objective and infos_c[k].gap == infos[k].gap
        v = block_views(p, x)
        for t, name in enumerate(("pos_x", "vel_x", "acc_x", "pos_y", "vel_y", "acc_y", "u_x", "u_y")):
            assert np.array_equal(tr[k][:, :, t], v[name]), (k, name)
        assert np.array_equal(s.fetch_vector(k), x)
    s.fetch_compact()
    assert s.run_stats()["d2h_bytes"] < d2h_full / 4
    # the one-call form
    prep = s.prepare(ps, gap_tol=1e-4, time_limit=60)
    tr2, infos2 = s.solve_prepared_compact(prep)
    assert all(np.array_equal(a, b) for a, b in zip(tr, tr2))
    s.close()


def test_two_warp_teams_give_the_same_results(testcase_problem, monkeypatch):
    """the throughput variant of the node kernel (two warps per node, four teams per SM) is chosen for rounds with many nodes;
    forced here for every round (MIQP_NARROW_MIN=1) and switched off (MIQP_NO_NARROW_TEAM): identical searches"""
    from planner_miqp_b200.scenarios import obstacle_scenario
    ps = [obstacle_scenario(k).build() for k in range(24)]
    s = P.Solver()
    monkeypatch.setenv("MIQP_NO_NARROW_TEAM", "1")
    xs0, infos0 = s.solve_batch(ps, gap_tol=1e-4, time_limit=60)
    monkeypatch.delenv("MIQP_NO_NARROW_TEAM")
    monkeypatch.setenv("MIQP_NARROW_MIN", "1")
    monkeypatch.setenv("MIQP_NO_WIDE_TEAM", "1")
    xs1, infos1 = s.solve_batch(ps, gap_tol=1e-4, time_limit=60)
    for a, b, xa, xb, p in zip(infos0, infos1, xs0, xs1, ps):
        assert a.status == b.status == 0 and a.proven and b.proven
        assert b.objective == pytest.approx(a.objective, rel=1e-9)
        assert b.nodes == a.nodes                      # the search does not depend on the team size
        np.testing.assert_allclose(xa, xb, atol=1e-6)
        viol, _ = O.max_violation(p, xb)
        assert viol <= 1e-6
    # a different horizon in the same batch: the two-warp layout is not used, results unchanged
    ps2 = ps[:4] + [testcase_problem]
    xs2, infos2 = s.solve_batch(ps2, gap_tol=1e-4, time_limit=60)
    assert [i.objective for i in infos2[:4]] == pytest.approx([i.objective for i in infos0[:4]], rel=1e-9)
    assert abs(infos2[4].objective - 9.57603) < 1e-5
    s.close()
