"""Pins the solver part of the CPU oracle: objective 9.57603 +- 1e-5 of
cplexmodel_testcase.dat (test/cplex_wrapper_test.cc:874), the region sequence and obstacle
side of the reference's pinned solution, SOS on/off-style invariance on test_sos.dat."""
import numpy as np

from oracle import oracle as O
from conftest import golden_vector


def test_testcase_objective_matches_reference(testcase_problem):
    p = testcase_problem
    for gap in (0.1, 1e-4):
        x, info = O.solve(p, gap_tol=gap, time_limit=60)
        assert info.status == 0
        assert abs(info.objective - 9.57603) < 1e-5
        assert info.max_violation < 1e-6
        assert info.gap <= gap
    seq = O.block_views(p, x)["active_region"][0].argmax(axis=1) + 1
    assert list(seq) == [1, 1, 1, 1] + [32] * 16
    # passes the obstacle on the -y side like the pinned CPLEX solution
    assert O.block_views(p, x)["pos_y"][0, 18] < -1.0


def test_fixed_binaries_of_golden_vector_reproduce_objective(testcase_problem, golden_solution):
    p = testcase_problem
    xg = golden_vector(p, golden_solution)
    rc, x, obj = O.solve_fixed(p, xg)
    assert rc == 0
    assert abs(obj - 9.57603) < 1e-5
    # trajectories agree with the 5-digit print of the reference
    for name in ("pos_x", "pos_y", "vel_x", "vel_y"):
        a = O.block_views(p, x)[name]
        b = O.block_views(p, xg)[name]
        assert np.max(np.abs(a - b)) < 2e-3, name


def test_mip_start_is_accepted(testcase_problem, golden_solution):
    p = testcase_problem
    xg = golden_vector(p, golden_solution)
    x, info = O.solve(p, gap_tol=1e-4, time_limit=60, warm=xg)
    assert info.status == 0 and abs(info.objective - 9.57603) < 1e-5


def test_sos_instance_solves(sos_problem):
    x, info = O.solve(sos_problem, gap_tol=1e-4, time_limit=60)
    assert info.status == 0 and info.proven
    assert info.max_violation < 1e-6
    x2, info2 = O.solve(sos_problem, gap_tol=0.1, time_limit=60)
    assert abs(info2.objective - info.objective) <= 0.1 * abs(info.objective)
