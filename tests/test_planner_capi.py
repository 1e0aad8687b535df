"""Planner-level C ABI (libmiqp_planner_c_api.so, include/miqp_planner_c_api.h): the CPU part of the
reference's own C-API tests (test/miqp_planner_c_api_test.cc:14-110, :197-260) plus the check that the
C++ host preparation produces the same problem data as the Python scenario builder."""
import math
import os

import numpy as np
import pytest

from planner_miqp_b200 import planner_capi as PC
from planner_miqp_b200.model_parameters import PlanBuilder, default_settings as py_default_settings
from oracle.dat_io import read_dat

REF = [0, 0, 5, 0, 30, 0]
STATE = [0, 0, 0, 1, 0.01, 0]


def test_exports_every_declared_symbol():
    assert sorted(PC.exported_symbols()) == sorted(PC.C_API_SYMBOLS)
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "miqp_planner_c_api.h")).read()
    for s in PC.C_API_SYMBOLS:
        assert s + "(" in hdr, s


def test_construction_and_getters():
    p = PC.CMiqpPlanner()
    assert p.h
    assert p.N == 20                       # ApolloDefaultSettings().nr_steps (c_api, get_n)
    assert p.ts == pytest.approx(0.25)
    assert p.lib.GetCollisionRadius(p.h) == pytest.approx(1.0)
    p.close()


def test_construction_settings():
    s = PC.default_settings()
    s.nr_steps = 44
    p = PC.CMiqpPlanner(s)
    assert p.N == 44                       # c_api, construction_settings
    p.close()


def test_unknown_fitting_table_combination_is_rejected():
    s = PC.default_settings()
    s.nr_regions = 64                      # 64 regions exist only for (vmax 10, vmin 1)
    with pytest.raises(ValueError):
        PC.CMiqpPlanner(s)
    s.max_velocity_fitting, s.minimum_region_change_speed = 10.0, 1.0
    PC.CMiqpPlanner(s).close()


def test_add_and_update_car():
    p = PC.CMiqpPlanner()
    idx = p.add_car(STATE, REF, 5, 1)
    assert idx == 0                        # c_api, add_car
    p.update_car(idx, [0, 0, 0, -2, 0.01, 0], REF)   # c_api, update_car
    assert p.add_car([0, 5, 0, 4, 0, 0], [0, 4, 50, 4], 5, 1) == 1
    p.close()


def test_add_update_and_remove_obstacle():
    p = PC.CMiqpPlanner()
    box = lambda dy: [[0, 0 + dy], [2, 0 + dy], [2, 4 + dy], [0, 4 + dy]]   # noqa: E731
    assert p.add_obstacle([box(0), box(1)]) == 0     # c_api, add_and_remove_obstacle
    p.update_obstacle(0, [box(1)])                   # c_api, update_obstacle
    p.remove_all_obstacles()
    assert p.add_obstacle([box(0)]) == 0
    p.close()


def test_set_debug_paths(tmp_path):
    p = PC.CMiqpPlanner()
    p.activate_debug_file_write(str(tmp_path), "test_c_api_")
    p.close()


def test_get_raw_reference():
    p = PC.CMiqpPlanner()
    idx = p.add_car(STATE, REF, 5, 1)
    ref = p.last_reference(idx)
    assert ref.shape == (p.N, PC.TRAJECTORY_SIZE)
    assert ref[0, 0] == 0 and ref[1, 0] == 0.25
    vel = np.float32(math.sqrt(np.float32(ref[-1, 3]) ** 2 + np.float32(ref[-1, 4]) ** 2))
    assert vel == 5                        # c_api, get_raw_reference: speed ramp reaches vDes
    assert np.all(ref[:, 5:] == 0)
    p.close()


def test_update_map_convex_and_non_convex():
    p = PC.CMiqpPlanner()
    assert p.update_map([0, 0, 0, 8, 4, 8, 4, 0, 0, 0])          # c_api, update_map (clockwise rectangle, closed)
    # an L-shaped road is decomposed into convex cells (host/convexified_map.hpp; reference ConvexifiedMap::Convert)
    assert p.update_map([0, 0, 10, 0, 10, 4, 4, 4, 4, 10, 0, 10, 0, 0])
    # a degenerate polygon is rejected
    assert not p.update_map([0, 0, 1, 0, 2, 0])
    p.close()


def test_obstacle_outside_the_drivable_area_is_filtered():
    p = PC.CMiqpPlanner()
    assert p.update_map([-10, -10, 60, -10, 60, 10, -10, 10])
    assert p.add_obstacle([[[100, 0], [101, 0], [101, 1], [100, 1]]], is_static=True) == -1
    assert p.add_obstacle([[[10, 0], [11, 0], [11, 1], [10, 1]]], is_static=True) == 0
    p.close()


@pytest.mark.parametrize("state,ref,vdes", [
    ([0, 4, 0, 0, 0.1, 0], [0, 0, 50, 0], 5.0),
    ([0, 5, 0, 0.3, 0.0, 0], [0, 0, 30, 0, 60, 12], 8.0),
    ([2, 0, 0, 1, 3.0, 0.2], [2, -5, 2, 40], 4.0),
])
def test_host_preparation_matches_python_builder(tmp_path, state, ref, vdes):
    """C++ MiqpPlanner::AddCar / AddObstacle / environment -> flattened ModelParameters == the Python mirror
    used by the scenario generators (planner-miqp_b200/model_parameters.py), value by value."""
    s = PC.default_settings()
    p = PC.CMiqpPlanner(s)
    p.add_car(state, ref, vdes, 1.0)
    box = [[[20 + 0.1 * i, -1], [23 + 0.1 * i, -1], [23 + 0.1 * i, 2], [20 + 0.1 * i, 2]] for i in range(p.N)]
    assert p.add_obstacle(box) == 0
    path = str(tmp_path / "parameters.txt")
    assert p.write_parameters(path)
    got = read_dat(path)

    b = PlanBuilder(py_default_settings())
    b.add_car(state, np.asarray(ref, dtype=float).reshape(-1, 2), vdes, 1.0)
    b.obstacles.append(([np.asarray(v, dtype=float) for v in box], False))
    want = b.build()
    want.initial_region[:] = 0            # chosen in Plan(), not in AddCar
    mask_start = want.possible_region.copy()
    assert (got.N, got.R, got.C, got.O, got.L, got.E) == (want.N, want.R, want.C, want.O, want.L, want.E)
    for k in ("ts", "min_vel_x_y", "max_vel_x_y", "total_min_acc", "total_max_acc", "total_min_jerk", "total_max_jerk",
              "maximum_slack", "WEIGHTS_SLACK", "WEIGHTS_SLACK_OBSTACLE", "minimum_region_change_speed",
              "relative_mip_gap_tolerance", "max_solution_time"):
        assert got.scal[k] == pytest.approx(want.scal[k], abs=1e-9), k
    for k in want.car:
        np.testing.assert_allclose(got.car[k], want.car[k], atol=1e-9, err_msg=k)
    np.testing.assert_allclose(got.x0, want.x0, atol=1e-9)
    for k in want.ref:
        np.testing.assert_allclose(got.ref[k], want.ref[k], atol=1e-8, err_msg=k)
    for k in want.lim:
        np.testing.assert_allclose(got.lim[k], want.lim[k], atol=1e-9, err_msg=k)
    np.testing.assert_allclose(got.frac, want.frac, atol=1e-9)
    for k in want.poly:
        np.testing.assert_allclose(got.poly[k], want.poly[k], atol=1e-9, err_msg=k)
    np.testing.assert_allclose(got.obs_edges, want.obs_edges, atol=1e-9)
    assert np.array_equal(got.obs_nedges, want.obs_nedges)
    # the builder forces the start region possible at build time; the planner does it inside Plan()
    diff = np.argwhere(np.asarray(got.possible_region) != mask_start)
    assert len(diff) <= 1
    p.close()


def test_receding_horizon_warm_start_matches_python_mirror(tmp_path):
    """MiqpPlanner::CalculateWarmstart (C++, reference src/miqp_planner.cpp:787-1051) against results.shift_warmstart:
    every family one step to the left, Euler extrapolation of the last step, last step of the binaries undecided"""
    from planner_miqp_b200.capi import ncols_of
    from planner_miqp_b200.results import shift_warmstart, block_views
    p = PC.CMiqpPlanner(PC.default_settings())
    p.add_car([0, 4, 0, 0, 0.1, 0], [0, 0, 50, 0], 5, 1)
    box = [[[20, -1], [23, -1], [23, 2], [20, 2]]] * p.N
    assert p.add_obstacle(box) == 0
    path = str(tmp_path / "p.txt")
    assert p.write_parameters(path)
    prob = read_dat(path)
    n = ncols_of(prob)
    rng = np.random.default_rng(3)
    x = np.zeros(n)
    v = block_views(prob, x)
    for name, a in v.items():
        if a.dtype == np.float64 and name in ("u_x", "u_y", "pos_x", "vel_x", "acc_x", "pos_y", "vel_y", "acc_y",
                                               "pos_x_front_UB", "pos_x_front_LB", "pos_y_front_UB", "pos_y_front_LB"):
            a[...] = rng.normal(size=a.shape) * 3.0 + 5.0
        else:
            a[...] = rng.integers(0, 2, size=a.shape)
    assert p.set_solution(x)
    w_cpp = p.warmstart_vector(n, relax_last_step=True)
    w_py = shift_warmstart(prob, x, relax_last=True)
    vc, vp = block_views(prob, w_cpp), block_views(prob, w_py)
    for name in ("u_x", "u_y", "pos_x", "vel_x", "acc_x", "pos_y", "vel_y", "acc_y"):
        np.testing.assert_allclose(vc[name], vp[name], atol=1e-12, err_msg=name)
    for name in ("deltacc", "deltacc_front", "region_change_not_allowed_x_positive"):
        np.testing.assert_array_equal(np.isnan(vc[name]), np.isnan(vp[name]), err_msg=name)
        np.testing.assert_array_equal(np.nan_to_num(vc[name]), np.nan_to_num(vp[name]), err_msg=name)
    # active region: shifted, last step undecided
    np.testing.assert_array_equal(vc["active_region"][:, :-1], v["active_region"][:, 1:])
    assert np.all(np.isnan(vc["active_region"][:, -1]))
    assert not p.set_solution(x[:-1])          # wrong size is refused
    p.close()


def test_parameter_dump_of_a_two_car_plan_with_environment(tmp_path):
    """the OPL .dat dump written by the C++ host (reference: parameters_<t>.txt, src/cplex_wrapper.cpp:141-155) is read back by
    the oracle's parser: two cars, one drivable cell, a moving obstacle"""
    s = PC.default_settings()
    p = PC.CMiqpPlanner(s)
    assert p.update_map([-50, -50, -50, 50, 50, 50, 50, -50, -50, -50])
    assert p.add_car([0, 4, 0, 1, -0.1, 0], [0, 0, 50, 0], 10, 1) == 0
    assert p.add_car([20, -4, 0, -1, 0.1, 0], [20, 0, -30, 0], 10, 1) == 1
    box = [[[5 + 0.2 * i, 3], [7 + 0.2 * i, 3], [7 + 0.2 * i, 5], [5 + 0.2 * i, 5]] for i in range(p.N)]
    assert p.add_obstacle(box, is_static=False, is_soft=True) == 0
    path = str(tmp_path / "two_cars.txt")
    assert p.write_parameters(path)
    got = read_dat(path)
    assert (got.C, got.N, got.R, got.O, got.L) == (2, 20, 16, 1, 4)
    assert got.obs_soft[0] == 1 and np.all(got.obs_nedges == 4)
    np.testing.assert_allclose(got.obs_edges[0, 3, 0], [5.6, 3, 7.6, 3], atol=1e-9)       # edge 0 of step 3: vertex 0 -> vertex 1
    np.testing.assert_allclose(got.x0, [[0, 4, 0, 1, -0.1, 0], [20, -4, 0, -1, 0.1, 0]], atol=1e-9)
    # ego gets lambda = 0.5 of the weights, the other car the rest (src/miqp_planner.cpp:346-352)
    np.testing.assert_allclose(got.car["WEIGHTS_POS_X"], [1.0, 1.0], atol=1e-9)
    np.testing.assert_allclose(got.car["WEIGHTS_JERK_X"], [0.5, 0.5], atol=1e-9)
    assert got.ref["x_ref"].shape == (2, 20) and got.ref["x_ref"][1, -1] < 20 < got.ref["x_ref"][0, -1] + 25
    # the environment is selected inside Plan() (ResetEnvironment); before the first plan the dump has none
    assert got.E == 0
    p.close()
