"""GPU parity, formulation part: the device row instantiation (rows_kernel) must equal the
oracle's OPL-order rows BIT FOR BIT (integer / index work and coefficient arithmetic), and the
device evaluator must reproduce objective and violation of the reference's pinned CPLEX vector
(test/cplex_wrapper_test.cc:283-457, :857-876)."""
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


@pytest.mark.parametrize("which", ["testcase", "sos"])
def test_rows_bit_exact(solver, testcase_problem, sos_problem, which):
    p = testcase_problem if which == "testcase" else sos_problem
    ref = O.build_rows(p)
    got = solver.assemble(p)
    names = ["rowptr", "cols", "vals", "lo", "hi"]
    for n, a, b in zip(names, ref, got):
        assert a.shape == b.shape, n
        assert np.array_equal(a, b), (n, int(np.flatnonzero(a != b)[0]))


def test_sizes_match_reference_pins(solver, testcase_problem, sos_problem):
    sz = solver.sizes(testcase_problem)
    assert (sz.nrows, sz.nnz, sz.nbin, sz.ncont) == (12361, 29834, 1240, 340)
    sz = solver.sizes(sos_problem)
    assert (sz.nrows, sz.nbin, sz.ncont) == (8944, 420, 240)


def test_rows_bit_exact_multi_car_and_soft(solver, testcase_problem):
    from scenarios_for_tests import multi_car_variant
    p = multi_car_variant(testcase_problem, cars=3)
    ref = O.build_rows(p)
    got = solver.assemble(p)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


def test_evaluate_golden_vector(solver, testcase_problem, golden_solution):
    p = testcase_problem
    x = golden_vector(p, golden_solution)
    obj, viol = solver.evaluate(p, x)
    assert obj == pytest.approx(O.objective(p, x), rel=1e-13)
    v_ref, _ = O.max_violation(p, x)
    assert viol == pytest.approx(v_ref, rel=1e-12)
    assert abs(obj - 9.57603) < 5e-4
