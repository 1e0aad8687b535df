"""Pins the model part of the CPU oracle against the reference's own known answers
(SURVEY.md section 8(c)): problem statistics of cplexmodel_testcase.dat
(test/cplex_wrapper_test.cc:866-871), the full CPLEX solution vector
(test/cplex_wrapper_test.cc:283-457) and the size formulas for test_sos.dat."""
import numpy as np

from oracle import oracle as O
from conftest import golden_vector


def test_testcase_sizes_match_reference_pins(testcase_problem, golden_solution):
    sz = O.sizes(testcase_problem)
    pins = golden_solution["_pins"]
    assert sz.nrows == pins["NrConstraints"] == 12361
    assert sz.nnz == pins["NonZeroCoefficients"] == 29834
    assert sz.nbin == pins["NrBinaryVariables"] == 1240
    assert sz.ncont == pins["NrFloatVariables"] == 340


def test_sos_sizes(sos_problem):
    sz = O.sizes(sos_problem)
    assert (sz.nrows, sz.nbin, sz.ncont) == (8944, 420, 240)


def test_golden_vector_is_feasible_and_has_pinned_objective(testcase_problem, golden_solution):
    p = testcase_problem
    x = golden_vector(p, golden_solution)
    viol, worst = O.max_violation(p, x)
    # the vector is printed with 5 significant digits: the front-axle equalities carry
    # 4.7e-4 of print error, the dynamics 4.2e-5
    assert viol < 6e-4, (viol, worst)
    obj = O.objective(p, x)
    assert abs(obj - 9.57603) < 5e-4  # 9.57584 at print precision
    ar = O.block_views(p, x)["active_region"][0]
    seq = ar.argmax(axis=1) + 1
    assert list(seq) == [1, 1, 1, 1] + [32] * 16


def test_csr_shape_and_bounds(testcase_problem):
    rowptr, cols, vals, lo, hi = O.build_rows(testcase_problem)
    assert rowptr[0] == 0 and rowptr[-1] == len(cols) == len(vals)
    assert np.all(np.diff(rowptr) >= 1)
    assert np.all(lo <= hi)
    lay = O.layout(testcase_problem)
    assert cols.min() >= 0 and cols.max() < lay.ncols
    assert int(np.count_nonzero(vals)) == 29834


def test_dat_roundtrip(tmp_path, testcase_problem):
    from oracle.dat_io import write_dat, read_dat
    f = tmp_path / "rt.dat"
    write_dat(testcase_problem, str(f))
    q = read_dat(str(f))
    a = O.build_rows(testcase_problem)
    b = O.build_rows(q)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
