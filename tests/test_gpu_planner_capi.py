"""Plans through the planner-level C ABI on the GPU: the GPU half of the reference's C-API tests
(test/miqp_planner_c_api_test.cc:112-196) and planner tests (test/miqp_planner_test.cc:279-308),
every solve cross-checked against the CPU oracle on the parameter dump the planner writes."""
import glob
import os

import numpy as np
import pytest

from planner_miqp_b200 import planner_capi as PC
from oracle import oracle as O
from oracle.dat_io import read_dat

pytestmark = pytest.mark.gpu

REF = [0, 0, 5, 0, 30, 0]
STATE = [0, 0, 0, 1, 0.01, 0]


def _oracle_objective(tmp_path, prefix):
    files = sorted(glob.glob(os.path.join(str(tmp_path), prefix + "parameters_*.txt")))
    assert files, "planner wrote no parameter dump"
    p = read_dat(files[-1])
    x, info = O.solve(p, gap_tol=p.scal["relative_mip_gap_tolerance"], time_limit=60)
    return p, info


def test_c_api_plan1_with_and_without_obstacle(tmp_path):
    p = PC.CMiqpPlanner()
    p.add_car(STATE, REF, 5, 1)
    N = p.N
    box = [[[10, -0.5], [11, -0.5], [11, 0.5], [10, 0.5]]] * N
    assert p.add_obstacle(box, is_static=False, is_soft=False) == 0
    p.activate_debug_file_write(str(tmp_path), "a_")
    assert p.plan(0.0)
    props = p.solution_properties()
    prob, oi = _oracle_objective(tmp_path, "a_")
    assert oi.status == 0
    # both stop at the 10 % gap of the default settings: the objectives bracket each other within it
    assert abs(props["objective"] - oi.objective) <= 0.1 * abs(oi.objective) + 1e-9
    p.remove_all_obstacles()
    assert p.plan(0.0)
    p.close()


def test_c_api_get_raw_traj():
    p = PC.CMiqpPlanner()
    idx = p.add_car(STATE, REF, 5, 1)
    assert p.plan(0.0)
    traj = p.trajectory(idx, 0.0)
    assert traj.shape == (p.N, PC.TRAJECTORY_SIZE)
    assert traj[0, 0] == 0 and traj[0, 1] == 0 and traj[1, 0] == 0.25
    assert traj[1, 1] == pytest.approx(0.005, abs=1e-3)      # test/miqp_planner_c_api_test.cc:192-196
    # the plan obeys the triple-integrator dynamics
    ts = p.ts
    x, vx, ax, ux = traj[:, 1], traj[:, 3], traj[:, 5], traj[:, 7]
    np.testing.assert_allclose(x[1:], x[:-1] + ts * vx[:-1] + ts ** 2 / 2 * ax[:-1] + ts ** 3 / 6 * ux[:-1], atol=1e-6)
    p.close()


def _planner_with_map(gap=1e-4, warm=0, regions=16):
    s = PC.default_settings()
    s.relative_mip_gap_tolerance = gap
    s.warmstartType = warm
    s.nr_regions = regions
    p = PC.CMiqpPlanner(s)
    assert p.update_map([-50, -50, -50, 50, 50, 50, 50, -50, -50, -50])
    return p


def test_planner_plan1_with_environment(tmp_path):
    """miqp_planner_test plan1: square map, straight reference, 1e-4 gap, equal to the oracle"""
    p = _planner_with_map()
    p.add_car([0, 4, 0, 0, 0.1, 0], [0, 0, 50, 0], 5, 1)
    p.activate_debug_file_write(str(tmp_path), "b_")
    assert p.plan(0.0)
    props = p.solution_properties()
    prob, oi = _oracle_objective(tmp_path, "b_")
    assert prob.E == 1 and prob.initial_region[0] == 1       # test/miqp_planner_test.cc:301-304
    assert props["objective"] == pytest.approx(oi.objective, rel=2e-4)
    assert props["gap"] <= 1e-4
    p.close()


def test_start_pose_outside_the_map_fails():
    p = _planner_with_map()
    p.add_car([80, 4, 0, 0, 0.1, 0], [80, 0, 120, 0], 5, 1)
    assert not p.plan(0.0)
    p.close()


def test_plan_batch_equals_single_plans():
    def make(k):
        p = _planner_with_map()
        p.add_car([0, 3 + 0.5 * k, 0, 0.2 * k, 0.1, 0], [0, 0, 50, 0], 5, 1)
        if k % 2:
            p.add_obstacle([[[15, -2], [18, -2], [18, 0.5], [15, 0.5]]] * p.N, is_static=True)
        return p
    singles = [make(k) for k in range(6)]
    objs = []
    for p in singles:
        assert p.plan(0.0)
        objs.append(p.solution_properties()["objective"])
    batch = [make(k) for k in range(6)]
    assert PC.plan_batch(batch, 0.0) == [True] * 6
    for p, o in zip(batch, objs):
        assert p.solution_properties()["objective"] == pytest.approx(o, rel=1e-9)
    for p, q in zip(singles, batch):
        np.testing.assert_allclose(p.trajectory(0), q.trajectory(0), atol=1e-9)
    for p in singles + batch:
        p.close()


def test_receding_horizon_replanning_with_warm_start():
    """UpdateCar with step 1 of the last plan + RECEDING_HORIZON_WARMSTART (config 2 flow through the C ABI)"""
    cold = _planner_with_map(warm=0, regions=32)
    warm = _planner_with_map(warm=1, regions=32)
    ref = [0, 0, 100, 0]
    state = [0, 5, 0, 0.5, 0.1, 0]
    for p in (cold, warm):
        p.add_car(state, ref, 6, 1)
        p.add_obstacle([[[18, -2.5], [21, -2.5], [21, 0.5], [18, 0.5]]] * p.N, is_static=True)
    for cycle in range(4):
        assert cold.plan(cycle * 0.25) and warm.plan(cycle * 0.25)
        pc, pw = cold.solution_properties(), warm.solution_properties()
        assert pw["objective"] == pytest.approx(pc["objective"], rel=2e-4)
        if cycle > 0:
            assert pw["nodes"] <= pc["nodes"] + 2
        t = warm.trajectory(0)
        state = [t[1, 1], t[1, 3], t[1, 5], t[1, 2], t[1, 4], t[1, 6]]
        for p in (cold, warm):
            p.update_car(0, state, ref, (cycle + 1) * 0.25)
    assert state[0] > 1.0      # the car moved along +x
    cold.close(); warm.close()


def test_two_cars_receding_horizon():
    """test/miqp_planner_test.cc:361-432: two cars head-on on the same line pass each other"""
    s = PC.default_settings()
    s.max_solution_time = 5.0
    p = PC.CMiqpPlanner(s)
    assert p.update_map([-50, -50, -50, 50, 50, 50, 50, -50, -50, -50])
    ref1, ref2 = [0, 0, 50, 0], [20, 0, -30, 0]
    s1, s2 = [0, 4, 0, 1, -0.1, 0], [20, -4, 0, -1, 0.1, 0]
    i1 = p.add_car(s1, ref1, 10, 1)
    i2 = p.add_car(s2, ref2, 10, 1)
    x1, x2 = [s1[0]], [s2[0]]
    for k in range(4):
        assert p.plan(0.25 * k)
        for idx, ref, xs in ((i1, ref1, x1), (i2, ref2, x2)):
            t = p.trajectory(idx)
            xs.append(t[1, 1])
            p.update_car(idx, [t[1, 1], t[1, 3], t[1, 5], t[1, 2], t[1, 4], t[1, 6]], ref)
    assert x1[0] < x1[-1] and x2[0] > x2[-1]
    p.close()


def test_batched_receding_horizon_uses_the_device_warm_start():
    """PlanBatch of the same planners, cycle after cycle, with RECEDING_HORIZON_WARMSTART: from the second cycle on the MIP starts
    are the previous incumbents shifted on the device (no host warm-start vectors); results equal cold planners within the gap"""
    def make(k, warm):
        p = _planner_with_map(warm=warm, regions=32)
        p.add_car([0, 4.5 + 0.3 * k, 0, 0.2 * k, 0.1, 0], [0, 0, 100, 0], 6, 1)
        p.add_obstacle([[[18 + k, -2.5], [21 + k, -2.5], [21 + k, 0.5], [18 + k, 0.5]]] * p.N, is_static=True)
        return p
    n = 5
    warm = [make(k, 1) for k in range(n)]
    cold = [make(k, 0) for k in range(n)]
    before = PC.device_warmstart_batches()
    for cycle in range(4):
        assert PC.plan_batch(warm, cycle * 0.25) == [True] * n
        assert PC.plan_batch(cold, cycle * 0.25) == [True] * n
        for k in range(n):
            pw, pc = warm[k].solution_properties(), cold[k].solution_properties()
            assert pw["objective"] == pytest.approx(pc["objective"], rel=2e-4)
            t = warm[k].trajectory(0)
            state = [t[1, 1], t[1, 3], t[1, 5], t[1, 2], t[1, 4], t[1, 6]]
            for p in (warm[k], cold[k]):
                p.update_car(0, state, [0, 0, 100, 0], (cycle + 1) * 0.25)
    # cycles 2-4 of the warm planners went through miqp_b200_batch_upload_replan
    assert PC.device_warmstart_batches() - before == 3
    for p in warm + cold:
        p.close()
