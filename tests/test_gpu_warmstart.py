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
