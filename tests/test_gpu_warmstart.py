"""Receding-horizon replanning with the shifted previous solution as (partial) MIP start
(reference: MiqpPlanner::CalculateWarmstart, src/miqp_planner.cpp:787-1051, fed to CPLEX by
initializeWarmstart, src/cplex_wrapper.cpp:494-639).  Config 2 of BASELINE.json."""
import numpy as np
import pytest

import planner_miqp_b200 as P
from planner_miqp_b200.scenarios import obstacle_scenario, advance_obstacle_scenario
from planner_miqp_b200.results import shift_warmstart, block_views
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_replanning_with_shifted_warm_start():
    s = P.Solver()
    b = obstacle_scenario(1)
    p = b.build()
    x, info = s.solve(p, gap_tol=1e-4, time_limit=60)
    assert info.status == 0 and info.proven
    for cycle in range(3):
        w = shift_warmstart(p, x)
        b = advance_obstacle_scenario(b, p, x)
        p = b.build()
        xc, ic = s.solve(p, gap_tol=1e-4, time_limit=60)
        xw, iw = s.solve(p, gap_tol=1e-4, time_limit=60, warm=w)
        assert ic.status == 0 and iw.status == 0 and ic.proven and iw.proven
        # both are optimal to the 1e-4 gap
        assert abs(iw.objective - ic.objective) <= 2e-4 * abs(ic.objective)
        assert iw.max_violation <= 1e-6
        assert iw.nodes <= ic.nodes + 2
        # the plan continues the previous one: step 0 of the new plan is step 1 of the old one
        v_old, v_new = block_views(b.build(), x) if False else None, block_views(p, xw)
        assert v_new["pos_x"][0, 0] == pytest.approx(p.x0[0, 0])
        x = xw
    s.close()


def test_partial_mip_start_with_nan_columns(testcase_problem):
    """NaN in the discrete columns of a step = undecided there (host_pack.hpp:decisions_from_solution)"""
    s = P.Solver()
    p = testcase_problem
    x, info = s.solve(p, gap_tol=1e-4, time_limit=60)
    w = x.copy()
    v = block_views(p, w)
    for name in ("active_region", "region_change_not_allowed_combined", "deltacc", "deltacc_front"):
        a = v[name]
        idx = [slice(None)] * a.ndim
        idx[2 if name.startswith("deltacc") else 1] = slice(p.N - 3, p.N)
        a[tuple(idx)] = np.nan
    xw, iw = s.solve(p, gap_tol=1e-4, time_limit=60, warm=w)
    assert iw.status == 0 and iw.proven
    assert iw.objective == pytest.approx(info.objective, rel=1e-6)
    xo, io = O.solve(p, gap_tol=1e-4, time_limit=60)
    assert iw.objective == pytest.approx(io.objective, rel=1e-6)
    s.close()


def test_device_shifted_warm_start_equals_host_shifted():
    """miqp_b200_batch_upload_replan: the previous incumbents shifted by one step on the device (no solution vector crosses
    PCIe) against the host path (results.shift_warmstart -> MIP start vector -> decisions_from_solution) and a cold solve, over
    four receding-horizon cycles of a small config-2 batch."""
    n = 12
    dev, ref = P.Solver(), P.Solver()
    builders = [obstacle_scenario(k) for k in range(n)]
    plans = [b.build() for b in builders]
    dev.upload(plans, gap_tol=1e-4, time_limit=60)
    dev.run()
    xs, infos = dev.fetch()
    assert all(i.status == 0 and i.proven for i in infos)
    nodes_dev = nodes_host = nodes_cold = 0
    for cycle in range(4):
        warm = [shift_warmstart(p, x) for p, x in zip(plans, xs)]
        builders = [advance_obstacle_scenario(b, p, x) for b, p, x in zip(builders, plans, xs)]
        plans = [b.build() for b in builders]
        dev.upload_replan(plans, gap_tol=1e-4, time_limit=60)
        dev.run()
        xd, idv = dev.fetch()
        xh, ih = ref.solve_batch(plans, gap_tol=1e-4, time_limit=60, warm=warm)
        xc, ic = ref.solve_batch(plans, gap_tol=1e-4, time_limit=60)
        for k in range(n):
            assert idv[k].status == 0 and idv[k].proven and ih[k].proven and ic[k].proven, (cycle, k, idv[k])
            assert abs(idv[k].objective - ic[k].objective) <= 2e-4 * abs(ic[k].objective)
            assert abs(ih[k].objective - ic[k].objective) <= 2e-4 * abs(ic[k].objective)
            assert idv[k].max_violation <= 1e-6
            viol, _ = O.max_violation(plans[k], xd[k])
            assert viol <= 1e-6
        nodes_dev += sum(i.nodes for i in idv); nodes_host += sum(i.nodes for i in ih); nodes_cold += sum(i.nodes for i in ic)
        xs = xd
    # the device shift is as good a start as the host's, and both beat cold starts
    assert nodes_dev <= 1.1 * nodes_host + n
    assert nodes_dev < nodes_cold
    dev.close(); ref.close()


def test_replan_needs_a_previous_run_of_the_same_batch():
    s = P.Solver()
    plans = [obstacle_scenario(k).build() for k in range(3)]
    with pytest.raises(P.MiqpB200Error):
        s.upload_replan(plans)
    s.upload(plans, gap_tol=1e-4, time_limit=60); s.run()
    with pytest.raises(P.MiqpB200Error):
        s.upload_replan(plans[:2])
    # a plan whose shape changed starts cold, the others warm
    other = obstacle_scenario(7, n_static=2).build()
    s.upload(plans, gap_tol=1e-4, time_limit=60); s.run()
    s.upload_replan([plans[0], other, plans[2]], gap_tol=1e-4, time_limit=60); s.run()
    xs, infos = s.fetch()
    assert all(i.status == 0 and i.proven for i in infos)
    xo, io = O.solve(other, gap_tol=1e-4, time_limit=60)
    assert infos[1].objective == pytest.approx(io.objective, rel=2e-4)
    s.close()
