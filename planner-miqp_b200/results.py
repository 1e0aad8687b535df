"""Views into a solution vector (RawResults families, reference src/miqp_planner_data.hpp:46-97)
and the receding-horizon warm start (MiqpPlanner::CalculateWarmstart,
reference src/miqp_planner.cpp:787-1051)."""
from __future__ import annotations

import numpy as np

from .capi import layout

CORE_BLOCKS = ("u_x", "u_y", "pos_x", "vel_x", "acc_x", "pos_y", "vel_y", "acc_y",
               "pos_x_front_UB", "pos_x_front_LB", "pos_y_front_UB", "pos_y_front_LB")
NWE_NAMES = ("notWithinEnvironmentRear", "notWithinEnvironmentFrontUbUb", "notWithinEnvironmentFrontLbUb",
             "notWithinEnvironmentFrontUbLb", "notWithinEnvironmentFrontLbLb")
RCNA_NAMES = ("region_change_not_allowed_x_positive", "region_change_not_allowed_y_positive",
              "region_change_not_allowed_x_negative", "region_change_not_allowed_y_negative",
              "region_change_not_allowed_combined")


def block_views(p, x: np.ndarray) -> dict:
    """named views into the column vector x (decision_variables.mod order, miqp_b200.h layout)"""
    l = layout(p)
    Cn, N, R, O, L, E, K = l.C, l.N, l.R, l.O, l.L, l.E, l.K
    out = {}
    for b, name in enumerate(CORE_BLOCKS):
        out[name] = x[b * Cn * N:(b + 1) * Cn * N].reshape(Cn, N)
    for k, name in enumerate(NWE_NAMES):
        out[name] = x[l.base_nwe + k * Cn * E * N: l.base_nwe + (k + 1) * Cn * E * N].reshape(Cn, E, N)
    out["active_region"] = x[l.base_ar:l.base_ar + Cn * N * R].reshape(Cn, N, R)
    for k, name in enumerate(RCNA_NAMES):
        out[name] = x[l.base_rcna + k * Cn * N: l.base_rcna + (k + 1) * Cn * N].reshape(Cn, N)
    out["deltacc"] = x[l.base_dcc:l.base_dcf].reshape(Cn, O, N, L)
    out["deltacc_front"] = x[l.base_dcf:l.base_so].reshape(Cn, O, N, L, 4)
    out["slackvarsObstacle"] = x[l.base_so:l.base_sof].reshape(Cn, O, N)
    out["slackvarsObstacle_front"] = x[l.base_sof:l.base_c2c].reshape(Cn, O, N, 4)
    out["car2car_collision"] = x[l.base_c2c:l.base_sv].reshape(K, K, N, 16)
    out["slackvars"] = x[l.base_sv:l.ncols].reshape(K, K, N, 4)
    return out


def shift_warmstart(p, x: np.ndarray, relax_last: bool = True) -> np.ndarray:
    """Warm start of the next planning cycle from the solution x of plan p: every family is shifted
    left by one time step; the last column is extrapolated with forward Euler for the trajectory
    (src/miqp_planner.cpp:887-912) and repeats the previous column for the binaries (:951-963,
    :1029-1046).  With relax_last the discrete columns of the last step are marked NaN, which the
    backend reads as "undecided" (the repeated column is often infeasible for the new last step)."""
    ts = p.scal["ts"]
    w = np.array(x, dtype=np.float64, copy=True)
    v = block_views(p, w)
    time_axis = {name: 1 for name in CORE_BLOCKS}
    time_axis.update({name: 2 for name in NWE_NAMES})
    time_axis.update({name: 1 for name in RCNA_NAMES})
    time_axis.update({"active_region": 1, "deltacc": 2, "deltacc_front": 2, "slackvarsObstacle": 2,
                      "slackvarsObstacle_front": 2, "car2car_collision": 2, "slackvars": 2})
    for name, ax in time_axis.items():
        a = v[name]
        if a.size == 0:
            continue
        a[...] = np.concatenate([np.take(a, range(1, a.shape[ax]), axis=ax), np.take(a, [a.shape[ax] - 1], axis=ax)], axis=ax)
    for axn in ("x", "y"):
        pos, vel, acc, u = v["pos_" + axn], v["vel_" + axn], v["acc_" + axn], v["u_" + axn]
        pos[:, -1] = pos[:, -2] + ts * vel[:, -2]
        vel[:, -1] = vel[:, -2] + ts * acc[:, -2]
        acc[:, -1] = acc[:, -2] + ts * u[:, -2]
        u[:, -1] = 0.0
        u[:, -2] = 0.0 if False else u[:, -2]
    if relax_last:
        for name in NWE_NAMES + RCNA_NAMES + ("active_region", "deltacc", "deltacc_front", "car2car_collision"):
            a = v[name]
            if a.size == 0:
                continue
            idx = [slice(None)] * a.ndim
            idx[time_axis[name]] = -1
            a[tuple(idx)] = np.nan
    return w
