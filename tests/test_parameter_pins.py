"""Derived-parameter pins of the reference's own unit tests, applied to the Python
mirror (planner-miqp_b200/model_parameters.py) and to the C++ host code (planner-miqp_b200/host/planner_prep.hpp, through
the parameter dump of the planner C ABI):
  mean angles and rotated jerk boxes        common/tests/parameter_preparer_test.cc:17-96
  rotated acc boxes                         cplexmodel/cplexmodel_testcase.dat:56-71 (the tables the reference test pins too)
  region index of a velocity direction      common/tests/regions_test.cc:19-62
  widening of the possible-region mask      common/tests/regions_test.cc:214-288"""
import numpy as np
import pytest

from planner_miqp_b200 import planner_capi as PC
from planner_miqp_b200.model_parameters import (ParameterPreparer, Settings, region_indices, reserve_neighbor_regions)
from oracle.dat_io import read_dat

MEAN_ANGLES_32 = [0.0982, 0.2945, 0.4909, 0.6872, 0.8836, 1.0799, 1.2763, 1.4726, 1.6690, 1.8653, 2.0617, 2.2580, 2.4544,
                  2.6507, 2.8471, 3.0434, 3.2398, 3.4361, 3.6325, 3.8288, 4.0252, 4.2215, 4.4179, 4.6142, 4.8106, 5.0069,
                  5.2033, 5.3996, 5.5960, 5.7923, 5.9887, 6.1850]
JERK_MAX_X_32 = [3.1228, 3.2772, 3.3057, 3.2072, 2.9854, 2.6489, 2.2106, 1.6873] + [1.6873, 2.2106, 2.6489, 2.9854, 3.2072, 3.3057, 3.2772, 3.1228]
JERK_MAX_X_32 = JERK_MAX_X_32 + JERK_MAX_X_32
JERK_MAX_Y_32 = [1.6873, 2.2106, 2.6489, 2.9854, 3.2072, 3.3057, 3.2772, 3.1228] + [3.1228, 3.2772, 3.3057, 3.2072, 2.9854, 2.6489, 2.2106, 1.6873]
JERK_MAX_Y_32 = JERK_MAX_Y_32 + JERK_MAX_Y_32


def _settings32():
    return Settings(nr_regions=32, max_velocity_fitting=20.0, minimum_region_change_speed=2.0)


def test_mean_angles_and_jerk_boxes_python():
    prep = ParameterPreparer(_settings32())
    np.testing.assert_allclose(prep.mean_angles, MEAN_ANGLES_32, atol=1e-3)
    jerk = prep.limits_per_region(prep.jerk)       # rows: min_x, max_x, min_y, max_y
    np.testing.assert_allclose(jerk[1], JERK_MAX_X_32, atol=1e-3)
    np.testing.assert_allclose(jerk[0], -np.array(JERK_MAX_X_32), atol=1e-3)
    np.testing.assert_allclose(jerk[3], JERK_MAX_Y_32, atol=1e-3)
    np.testing.assert_allclose(jerk[2], -np.array(JERK_MAX_Y_32), atol=1e-3)


def test_acc_and_jerk_boxes_equal_the_fixture_tables(testcase_problem):
    prep = ParameterPreparer(_settings32())
    acc, jerk = prep.limits_per_region(prep.acc), prep.limits_per_region(prep.jerk)
    lim = testcase_problem.lim
    for k, row in (("min_acc_x", acc[0]), ("max_acc_x", acc[1]), ("min_acc_y", acc[2]), ("max_acc_y", acc[3]),
                   ("min_jerk_x", jerk[0]), ("max_jerk_x", jerk[1]), ("min_jerk_y", jerk[2]), ("max_jerk_y", jerk[3])):
        np.testing.assert_allclose(row, lim[k][0], atol=6e-5, err_msg=k)      # the fixture prints five digits
    np.testing.assert_allclose(prep.frac, testcase_problem.frac, atol=6e-4)


def test_cpp_host_tables_equal_the_fixture_tables(tmp_path, testcase_problem):
    s = PC.default_settings()
    s.nr_regions = 32
    p = PC.CMiqpPlanner(s)
    p.add_car([0, 5, 0, 0, 0.1, 0], [0, 0, 100, 0], 5, 1)
    path = str(tmp_path / "p.txt")
    assert p.write_parameters(path)
    got = read_dat(path)
    for k in ("min_acc_x", "max_acc_x", "min_acc_y", "max_acc_y", "min_jerk_x", "max_jerk_x", "min_jerk_y", "max_jerk_y"):
        np.testing.assert_allclose(got.lim[k][0], testcase_problem.lim[k][0], atol=6e-5, err_msg=k)
    np.testing.assert_allclose(got.lim["max_jerk_x"][0], JERK_MAX_X_32, atol=1e-3)
    np.testing.assert_allclose(got.frac, testcase_problem.frac, atol=6e-4)
    for k in testcase_problem.poly:
        np.testing.assert_allclose(got.poly[k], testcase_problem.poly[k], atol=6e-5 * max(1.0, np.abs(testcase_problem.poly[k]).max()), err_msg=k)
    # global limits = extreme per-region limit +- 1e-6 (src/miqp_planner.cpp:226-245); the fixture prints 4.2922 / 3.3057
    assert got.scal["total_max_acc"] == pytest.approx(4.2922, abs=1e-4) and got.scal["total_min_acc"] == pytest.approx(-4.2922, abs=1e-4)
    assert got.scal["total_max_jerk"] == pytest.approx(3.3057, abs=1e-4) and got.scal["total_min_jerk"] == pytest.approx(-3.3057, abs=1e-4)
    p.close()


@pytest.mark.parametrize("vx,vy,expect", [(0.1, 0.01, {0}), (0.1951, 0.9808, {6, 7}), (0.1, 0.9, {7}), (-0.1, -0.9, {23}), (0.1, -0.01, {31})])
def test_region_index(vx, vy, expect):
    prep = ParameterPreparer(_settings32())
    assert set(region_indices(prep.frac, vx, vy)) == expect


def test_reserve_neighbor_regions():
    r = np.zeros(32, dtype=np.int32); r[[2, 3]] = 1
    assert reserve_neighbor_regions(r, 1) and list(np.flatnonzero(r)) == [1, 2, 3, 4]
    r = np.zeros(32, dtype=np.int32); r[[0, 1]] = 1
    assert reserve_neighbor_regions(r, 1) and list(np.flatnonzero(r)) == [0, 1, 2, 31]
    r = np.zeros(32, dtype=np.int32); r[[30, 31]] = 1
    assert reserve_neighbor_regions(r, 1) and list(np.flatnonzero(r)) == [0, 29, 30, 31]
    r = np.zeros(32, dtype=np.int32); r[[0, 1]] = 1
    assert reserve_neighbor_regions(r, 2) and list(np.flatnonzero(r)) == [0, 1, 2, 3, 30, 31]


def test_possible_regions_of_a_straight_reference_cpp(tmp_path):
    """a car heading along +x: wedges 0 and R-1 touch the heading, one neighbour on either side is reserved"""
    s = PC.default_settings()
    s.nr_regions = 32
    p = PC.CMiqpPlanner(s)
    p.add_car([0, 5, 0, 0, 0.0, 0], [0, 0, 100, 0], 5, 1)
    path = str(tmp_path / "p.txt")
    assert p.write_parameters(path)
    got = read_dat(path)
    assert list(np.flatnonzero(got.possible_region[0])) == [0, 1, 30, 31]
    p.close()


def test_reference_trajectory_pins_python_and_cpp():
    """common/tests/reference_trajectory_generator_test.cc:22-95: straight centre line, speed ramp over 0.1 m"""
    from planner_miqp_b200.model_parameters import PolyLine, reference_trajectory
    line = PolyLine([[0, 0], [50, 0], [100, 0]], 0.2)
    traj = reference_trajectory(line, 0.0, 0.0, 5.0, 20, 0.2, 10.0, 0.1)      # rows x, y, theta, v; row 0 is filled by the caller
    assert traj[1, 3] == pytest.approx(10.0, abs=1e-3) and traj[-1, 3] == pytest.approx(10.0, abs=1e-3)
    assert traj[1, 0] == pytest.approx(5.0 * 0.2, abs=1e-3) and traj[2, 0] == pytest.approx(5.0 * 0.2 + 10.0 * 0.2, abs=1e-3)
    slow = reference_trajectory(line, 0.0, 0.0, 0.1, 20, 0.2, 10.0, 0.1)      # "starting": from 0.1 m/s
    assert slow[1, 3] > 0.1 and slow[-1, 3] == pytest.approx(10.0, abs=1e-3)
    s = PC.default_settings()
    s.ts = 0.2
    p = PC.CMiqpPlanner(s)
    p.add_car([0, 5, 0, 0, 0, 0], [0, 0, 50, 0, 100, 0], 10.0, 0.1)
    ref = p.last_reference(0)
    v = np.hypot(ref[:, 3], ref[:, 4])
    assert v[0] == pytest.approx(5.0, abs=1e-3) and v[1] == pytest.approx(10.0, abs=1e-3) and v[-1] == pytest.approx(10.0, abs=1e-3)
    assert ref[1, 1] == pytest.approx(1.0, abs=1e-3) and ref[2, 1] == pytest.approx(3.0, abs=1e-3)
    np.testing.assert_allclose(ref[1:, 1], traj[1:, 0], atol=1e-9)             # C++ and Python generators agree point by point
    p.close()


def test_fitting_table_spot_values():
    """common/tests/fitting_polynomial_parameters_test.cc:26-34, :59-76 and test/miqp_planner_test.cc:127-134 (exact values)"""
    from planner_miqp_b200.model_parameters import fitting_tables
    t32 = fitting_tables(32, 20, 2)
    assert t32["POLY_SINT_UB"][0, 0] == 0.18825 and t32["POLY_SINT_UB"][31, 2] == 0.049029
    assert t32["POLY_COSS_LB"][0, 0] == 0.97651
    assert list(t32["POLY_SINT_UB"][7]) == [1.0047, -0.0055247, 0.00032035]
    t16 = fitting_tables(16, 20, 2)
    assert t16["POLY_SINT_UB"][0, 0] == 0.348508875688441 and t16["POLY_SINT_UB"][3, 0] == 1.0037509353218053
    assert fitting_tables(64, 10, 1)["POLY_SINT_UB"].shape == (64, 3)
    for bad in ((64, 20, 2), (128, 10, 1), (32, 15, 2)):          # combinations the reference does not ship throw
        with pytest.raises(ValueError):
            fitting_tables(*bad)


def test_fitting_tables_in_the_cpp_host(tmp_path):
    """the C++ host selects the same tables (generated include of host/planner_prep.hpp): 16 regions, default settings"""
    from planner_miqp_b200.model_parameters import fitting_tables
    p = PC.CMiqpPlanner(PC.default_settings())
    p.add_car([0, 5, 0, 0, 0.1, 0], [0, 0, 100, 0], 5, 1)
    path = str(tmp_path / "p.txt")
    assert p.write_parameters(path)
    got = read_dat(path)
    t16 = fitting_tables(16, 20, 2)
    for k in t16:
        np.testing.assert_allclose(got.poly[k], np.round(t16[k], 10), atol=1e-12, err_msg=k)   # reals enter the model rounded to 10 decimals
    assert got.poly["POLY_SINT_UB"][0, 0] == pytest.approx(0.348508875688441, abs=1e-10)
    p.close()
